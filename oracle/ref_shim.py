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

This file only works where /root/reference exists (the build container).  It is
used by ``oracle/gen_golden.py`` to produce ``tests/golden/*.npz`` and by the
CPU tests that cross-check ``oracle/uem_oracle.py`` against the reference when
the tree is present.  Nothing that runs on the GPU box imports it.
"""
import importlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("UEM_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "uemda", "gast"))


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
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self


_loaded = {}


def load_reference():
    """Returns a namespace with the reference's hot-path symbols."""
    if _loaded:
        return _loaded["ns"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
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
