"""TEST / BASELINE INFRASTRUCTURE ONLY -- byte-compiles the reference's own hot-path modules into oracle/_ref/.

    python oracle/build_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

/root/reference does not exist on the GPU box, and the reference is pure Python, so "building" it means compiling the
modules its mining path imports (found by import tracing under oracle/ref_shim.py) to bytecode files (``.refbin``), from
the sources where they lie, into the git-ignored ``oracle/_ref/`` -- the same way a C reference would be compiled into a
``.so`` there.  No reference source is copied.  ``oracle/ref_shim.py`` then imports ``uemda.gast.{alignment,
pseudo_generation, balance}`` and ``uemda.utils.tools`` from ``oracle/_ref`` when ``/root/reference`` is absent, which is
how ``bench.py --impl reference`` / ``cpu_baseline`` time the reference's OWN functions on the GPU box's host cores
(``kind: "reference"``).  Same interpreter on both sides (same image): the bytecode's magic number is checked at import.
"""
import os
import py_compile
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("UEM_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
BIN_EXT = ".refbin"

# the modules uemda.gast.{alignment,pseudo_generation,balance} pull in under the shim (everything else is stubbed there)
MODULES = [
    "uemda/__init__.py",
    "uemda/gast/__init__.py",
    "uemda/gast/alignment.py",
    "uemda/gast/balance.py",
    "uemda/gast/class_ware_whiten.py",
    "uemda/gast/coral.py",
    "uemda/gast/pseudo_generation.py",
    "uemda/utils/__init__.py",
    "uemda/utils/tools.py",
]


def build(verbose=False):
    """Returns the output directory, or None when the reference tree is not present (the GPU box: use what is there)."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "uemda", "gast")):
        return None
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    for rel in MODULES:
        src = os.path.join(REFERENCE_ROOT, rel)
        dst = os.path.join(OUT, rel[:-3] + BIN_EXT)   # bytecode, loaded by ref_shim's importer (not named .pyc: snapshot
        #                                               tools commonly drop *.pyc, and these files must travel to the GPU box)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, dfile=rel, doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        if verbose:
            print("compiled", rel, "->", os.path.relpath(dst, HERE))
    with open(os.path.join(OUT, "BUILD_INFO"), "w") as f:
        f.write("python %s\nfrom %s\nmodules %d\n" % (sys.version.split()[0], REFERENCE_ROOT, len(MODULES)))
    return OUT


if __name__ == "__main__":
    out = build(verbose=True)
    print(out or "reference tree not present: nothing built")
