"""Builds uemda_b200/libuem_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m uemda_b200.build [--force] [--verbose]

The library is plain C-ABI (include/uem_b200.h), links cudart statically and has no torch or
pybind dependency.  Objects are cached under uemda_b200/csrc/.obj keyed by source mtime.
"""
import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, ".obj")
LIB = os.path.join(HERE, "libuem_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libuem_b200.so")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=(), lib=None, obj_dir=None):
    """extra_flags / lib / obj_dir: development A/B builds (e.g. -DUEM_... knobs) into a separate library."""
    nvcc = _nvcc()
    LIB = lib or globals()["LIB"]
    OBJ = obj_dir or globals()["OBJ"]
    os.makedirs(OBJ, exist_ok=True)
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    jobs = []
    objs = []
    for src in sources:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, r in ex.map(run, jobs):
                if verbose or r.returncode:
                    sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
                if r.returncode:
                    raise RuntimeError("nvcc failed for %s" % cmd[-3])
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
               "-cudart", "static", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link of libuem_b200.so failed")
    return LIB


if __name__ == "__main__":
    flags = [a for a in sys.argv[1:] if a.startswith("-D")]
    tag = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--tag=")), None)
    if tag:
        path = build(force=True, verbose="--verbose" in sys.argv, extra_flags=flags,
                     lib=os.path.join(HERE, "libuem_b200_%s.so" % tag), obj_dir=os.path.join(CSRC, ".obj_" + tag))
    else:
        path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, extra_flags=flags)
    print(path)
