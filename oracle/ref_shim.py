"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Loads the *unmodified* reference modules from /root/reference so that their own
function bodies can be executed on CPU torch.  Only third-party seams are
shimmed:

  * ``torch_scatter.scatter``  -> a small pure-torch restatement (the package is
    not installed here and its source is not in the reference tree; semantics
    restated from its public documentation: index broadcast to ``src``, output
    length ``index.max()+1`` along ``dim``, untouched slots are 0).
    Call sites: uemda/gast/alignment.py:187 (sum, int64) and :245 (max, f32).
  * plotting / IO / model-zoo imports that the hot path never calls
    (matplotlib, ttach, ever, skimage, uemda.viz, uemda.datasets) -> empty stubs.
  * ``Tensor.cuda`` -> identity (the container has no GPU).

It needs the reference's modules: the tree under /root/reference (the build container) or,
where that is absent (the GPU box), their byte-compiled form under oracle/_ref (oracle/build_ref.py).
It is used by ``oracle/gen_golden*.py`` to produce ``tests/golden/*.npz``, by the CPU tests that
cross-check ``oracle/uem_oracle.py`` against the reference, and by ``bench.py``'s CPU legs
(``--impl reference``, ``cpu_baseline``) to time the reference's own functions on the host cores.
"""
import importlib
import os
import sys
import types

import torch

_SRC_ROOT = os.environ.get("UEM_REFERENCE_ROOT", "/root/reference")
_BUILT_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # oracle/build_ref.py (bytecode)
_BIN_EXT = ".refbin"


def _pick_root():
    if os.path.isdir(os.path.join(_SRC_ROOT, "uemda", "gast")):
        return _SRC_ROOT
    if os.path.exists(os.path.join(_BUILT_ROOT, "uemda", "gast", "alignment" + _BIN_EXT)):
        return _BUILT_ROOT
    return _SRC_ROOT


REFERENCE_ROOT = _pick_root()


def reference_available():
    return (os.path.isdir(os.path.join(REFERENCE_ROOT, "uemda", "gast"))
            and any(os.path.exists(os.path.join(REFERENCE_ROOT, "uemda", "gast", "alignment" + ext)) for ext in (".py", _BIN_EXT)))


def reference_kind():
    """'source' (the tree under /root/reference), 'built' (oracle/_ref bytecode of the same files) or None."""
    if not reference_available():
        return None
    return "built" if REFERENCE_ROOT == _BUILT_ROOT else "source"


def _scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    """Restatement of torch_scatter.scatter for the two reference call shapes."""
    if dim < 0:
        dim = src.dim() + dim
    index = index.expand_as(src) if index.shape != src.shape else index
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    res = torch.zeros(shape, dtype=src.dtype, device=src.device)
    if reduce in ("sum", "add"):
        return res.scatter_add_(dim, index, src)
    if reduce == "mean":
        res.scatter_add_(dim, index, src)
        cnt = torch.zeros(shape, dtype=src.dtype, device=src.device)
        cnt.scatter_add_(dim, index, torch.ones_like(src))
        cnt.clamp_(min=1)
        return res / cnt if src.is_floating_point() else torch.div(res, cnt, rounding_mode="floor")
    if reduce == "max":
        return res.scatter_reduce_(dim, index, src, "amax", include_self=False)
    if reduce == "min":
        return res.scatter_reduce_(dim, index, src, "amin", include_self=False)
    raise ValueError(reduce)


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Stub(self.__name__ + "." + name)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return None


def _install_stubs():
    if "torch_scatter" not in sys.modules:
        ts = types.ModuleType("torch_scatter")
        ts.scatter = _scatter
        sys.modules["torch_scatter"] = ts
    for name in ("matplotlib", "matplotlib.pyplot", "ttach", "ever", "skimage", "skimage.io",
                 "audtorch", "audtorch.metrics", "audtorch.metrics.functional"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    sk = sys.modules["skimage.io"]
    if isinstance(sk, _Stub):
        sk.imsave = lambda *a, **k: None
    # uemda.viz / uemda.datasets pull cv2 + dataset stacks the path never uses
    viz = types.ModuleType("uemda.viz")
    viz.VisualizeSegmm = object
    sys.modules.setdefault("uemda.viz", viz)
    ds = types.ModuleType("uemda.datasets")
    ds.__all__ = []
    sys.modules.setdefault("uemda.datasets", ds)
    if not torch.cuda.is_available():
        force_cpu()


_saved_cuda = {}


def force_cpu():
    """The reference moves its state with ``.cuda()`` (alignment.py:48-77, balance.py:25): make that the identity so its
    functions run on the host cores (always the case in the build container; on the GPU box this is what makes the CPU
    arm a CPU arm).  ``restore_cuda()`` undoes it."""
    if not _saved_cuda:
        _saved_cuda["tensor"], _saved_cuda["module"] = torch.Tensor.cuda, torch.nn.Module.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self


def restore_cuda():
    if _saved_cuda:
        torch.Tensor.cuda, torch.nn.Module.cuda = _saved_cuda.pop("tensor"), _saved_cuda.pop("module")


class _BuiltFinder:
    """meta-path finder for the byte-compiled reference modules under oracle/_ref (<module>.refbin / <pkg>/__init__.refbin)"""

    def __init__(self, root):
        self.root = root

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery
        import importlib.util
        if fullname != "uemda" and not fullname.startswith("uemda."):
            return None
        base = os.path.join(self.root, *fullname.split("."))
        pkg = os.path.join(base, "__init__" + _BIN_EXT)
        if os.path.exists(pkg):
            loader = importlib.machinery.SourcelessFileLoader(fullname, pkg)
            return importlib.util.spec_from_file_location(fullname, pkg, loader=loader, submodule_search_locations=[base])
        mod = base + _BIN_EXT
        if os.path.exists(mod):
            loader = importlib.machinery.SourcelessFileLoader(fullname, mod)
            return importlib.util.spec_from_file_location(fullname, mod, loader=loader)
        return None


_loaded = {}


def load_reference():
    """Returns a namespace with the reference's hot-path symbols."""
    if _loaded:
        return _loaded["ns"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT == _BUILT_ROOT:
        if not any(isinstance(f, _BuiltFinder) for f in sys.meta_path):
            sys.meta_path.insert(0, _BuiltFinder(_BUILT_ROOT))
    elif REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    alignment = importlib.import_module("uemda.gast.alignment")
    pseudo = importlib.import_module("uemda.gast.pseudo_generation")
    balance = importlib.import_module("uemda.gast.balance")
    ns = types.SimpleNamespace(
        alignment=alignment, pseudo_generation=pseudo, balance=balance,
        Aligner=alignment.Aligner, DownscaleLabel=alignment.DownscaleLabel,
        pseudo_selection=pseudo.pseudo_selection, pseudo_selection1=pseudo.pseudo_selection1,
        UVEMLoss=balance.UVEMLoss, UPSLoss=balance.UPSLoss, ClassBalance=balance.ClassBalance,
        loss_calc_uvem=balance.loss_calc_uvem, scatter=_scatter,
    )
    _loaded["ns"] = ns
    return ns


class NullLogger:
    def info(self, *a, **k):
        pass
