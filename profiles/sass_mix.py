"""Opcode mix of one kernel from `ncu -i X.ncu-rep --page source --csv -k regex:NAME` (first kernel instance):
executed warp instructions and stall samples per SASS opcode."""
import collections
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    # several kernel instances may be concatenated: take the first block
    hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    start = hdr_i[0]
    end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
    hdr = rows[start]
    ie, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    mix = collections.Counter()
    samp = collections.Counter()
    tot = 0
    for r in rows[start + 1:end]:
        if len(r) <= ie:
            continue
        src = r[isrc].strip()
        parts = src.split()
        if not parts:
            continue
        op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
        op = op.rstrip(";")
        base = ".".join(op.split(".")[:2]) if op.split(".")[0] in ("LDG", "STG", "LDS", "STS", "LDGSTS", "MUFU", "ATOMG", "RED") else op.split(".")[0]
        n = int(r[ie] or 0)
        mix[base] += n
        samp[base] += int(r[isamp] or 0)
        tot += n
    stot = sum(samp.values()) or 1
    print("total warp instructions: %d (%d SASS lines)" % (tot, end - start - 1))
    for op, n in mix.most_common(top):
        print("%-14s %10d %5.1f%%   samples %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * samp[op] / stot))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
