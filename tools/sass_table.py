#!/usr/bin/env python
"""Static SASS evidence for the shipped library: per kernel, the counts of the mnemonics that prove how it moves data and
does its arithmetic (TMA: UTMALDG / UBLKCP; cp.async: LDGSTS; packed fp32x2: FFMA2 / FMUL2 / FADD2; special function: MUFU;
3-input min/max: FMNMX3; tensor cores: UTC*MMA -- expected absent, north_star rules them out), plus registers, spills and
static shared memory from the cubin's resource usage.

    python tools/sass_table.py [--all] [library.so] > profiles/rNN_sass_table.md

By default only the instantiations the benchmarks run (6 and 7 classes) are listed; --all lists every kernel.
"""
import collections
import re
import subprocess
import sys

LIB = next((a for a in sys.argv[1:] if a.endswith(".so")), "uemda_b200/libuem_b200.so")
ALL = "--all" in sys.argv
OPS = ["UTMALDG", "UBLKCP", "LDGSTS", "LDG", "STG", "LDS", "STS", "ATOMS", "ATOMG", "RED", "FFMA2", "FMUL2", "FADD2", "FFMA", "MUFU",
       "FMNMX3", "FMNMX", "SHFL", "MATCH", "SYNCS", "UTC", "HMMA", "LDL", "STL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            d = dict(re.findall(r"(\w+):(\d+)", line))
            usage[cur] = d
            cur = None
    counts = {}
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["_total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + ".") or (o == "UTC" and op.startswith("UTC") and "MMA" in op):
                counts[cur][o] += 1
                break
    names = demangle(list(counts))
    rows = []
    for mangled, c in counts.items():
        nm = names.get(mangled, mangled)
        nm = re.sub(r"\(anonymous namespace\)::", "", nm)
        nm = re.sub(r"^void ", "", nm)
        short = re.sub(r"\(.*$", "", nm)
        if not ALL:
            m = re.search(r"<\(?i?n?t?\)?(\d+)", short)
            if m and int(m.group(1)) not in (6, 7):
                continue
        u = usage.get(mangled, {})
        rows.append((short, c, u))
    rows.sort(key=lambda r: r[0])
    cols = ["UTMALDG", "UBLKCP", "LDGSTS", "LDG", "STG", "LDS", "STS", "FFMA2", "FMUL2", "FADD2", "MUFU", "FMNMX3", "ATOMS", "UTC", "LDL", "STL"]
    print("| kernel | instr | regs | stack B | smem B | " + " | ".join(cols) + " |")
    print("|---|---|---|---|---|" + "---|" * len(cols))
    for short, c, u in rows:
        print("| `%s` | %d | %s | %s | %s | %s |" % (short, c["_total"], u.get("REG", "?"), u.get("STACK", "?"), u.get("SHARED", "?"),
                                                     " | ".join(str(c[o]) for o in cols)))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("\nWhole library (%d kernels): " % len(counts) + ", ".join("%s %d" % (o, tot[o]) for o in OPS) + ".")
    print("`UTC*MMA` / `HMMA` = 0: no tensor-core instruction anywhere (the path is HBM-bound integer / fp32 work, north_star).")
    print("`LDL` / `STL` > 0 marks a kernel with register spills (see the stack column).")


if __name__ == "__main__":
    main()
