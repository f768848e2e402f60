"""Drop-in for the ``torch_scatter.scatter`` seam used by the mining path.

Reference call sites: uemda/gast/alignment.py:187 (``reduce='sum'`` on an int64 one-hot) and :245
(``reduce='max'`` on fp32 probabilities), both with ``src (b,N,c)``, ``index (b,N,1)``, ``dim=1``.
Semantics kept: index broadcast over the class dim, output length ``index.max()+1`` along ``dim``
(one host sync, as in torch_scatter) unless ``dim_size`` is given, untouched slots are 0.
"""
import torch

from . import ops


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    if out is not None:
        raise NotImplementedError("scatter(out=...) is not used by the reference and not supported")
    if src.dim() != 3:
        raise NotImplementedError("scatter: only the reference call shape src (b,N,c), dim=1 is supported")
    d = dim if dim >= 0 else src.dim() + dim
    if d != 1:
        raise NotImplementedError("scatter: only dim=1 (pixels) is supported")
    if reduce not in ("sum", "add", "max", "mean"):
        raise ValueError("scatter: unknown reduce %r" % (reduce,))
    b, n, c = src.shape
    if index.dim() == 3:
        if index.shape[2] == c and c != 1:
            # an index already expanded over classes must be constant along the class dim
            index = index[:, :, 0]
        else:
            index = index.reshape(b, n)
    assert index.shape == (b, n), "index must be (b,N) or (b,N,1)"
    return ops.region_reduce(src, index, reduce=reduce, dim_size=dim_size, planar=False)
