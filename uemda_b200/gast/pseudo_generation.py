"""Class-wise confidence thresholding -> hard pseudo-labels (drop-in for the hot-path functions of
uemda/gast/pseudo_generation.py).  Same signatures, same return types, CUDA tensors only."""
import torch

from .. import config, ops

__all__ = ["pseudo_selection", "pseudo_selection1"]


def _class_max_checked(mask):
    """per-(b,c) max of `mask`; reproduces the reference's range assert (pseudo_generation.py:36,71)."""
    cached = getattr(mask, "_uem_stats", None)
    if cached is not None and cached.valid_for(mask):
        stats = cached.stats                     # (b, c+2) statistics table raised by the refine kernel
        if config.strict_asserts:
            cmax, imin = ops.class_stats_decode(stats, mask.shape[1])
            host = torch.cat([cmax.reshape(-1), imin]).cpu()
            hi, lo = host[:cmax.numel()].max().item(), host[cmax.numel():].min().item()
            assert hi <= 1 and lo >= 0, print(hi, lo)
        return None, stats
    cmax, cmin, _ = ops.class_max(mask)
    if config.strict_asserts:
        host = torch.stack([cmax, cmin]).cpu()
        hi, lo = host[0].max().item(), host[1].min().item()
        assert hi <= 1 and lo >= 0, print(hi, lo)
    return cmax, None


def _select(mask, cutoff_top, cutoff_low, return_type, ignore_label, variant):
    assert return_type in ["ndarray", "tensor"]
    assert mask.dim() == 4, "mask must be (b, c, h, w)"
    cmax, stats = _class_max_checked(mask)
    if stats is not None and variant == 0:
        ret = ops.pseudo_select_stats(mask, stats, cutoff_top, cutoff_low, ignore_label)
    else:
        if cmax is None:
            cmax = ops.class_max(mask)[0]
        ret = ops.pseudo_select(mask, cmax, cutoff_top, cutoff_low, ignore_label, variant)
    return ret.cpu().numpy() if return_type == "ndarray" else ret


def pseudo_selection(mask, cutoff_top=0.8, cutoff_low=0.6, return_type="ndarray", ignore_label=-1):
    """pseudo_generation.py:59-93.  mask (b,c,h,w) probabilities in [0,1]; a pixel keeps class j iff j is
    the only class with p > max(cutoff_top * max_px p[b,j], cutoff_low); everything else -> ignore_label.
    Returns (b,h,w) int64 (CUDA tensor, or ndarray for return_type='ndarray')."""
    return _select(mask, cutoff_top, cutoff_low, return_type, ignore_label, 0)


def pseudo_selection1(mask, cutoff_top=0.8, cutoff_low=0.6, return_type="ndarray", ignore_label=-1):
    """pseudo_generation.py:24-56: label = argmax (lowest index on ties), ignored iff p_max < thr[label]."""
    return _select(mask, cutoff_top, cutoff_low, return_type, ignore_label, 1)
