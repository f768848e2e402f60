"""uemda_b200 -- B200-native (sm_100a) pseudo-label mining path of StuLiu/UemDA.

Drop-in modules mirror the reference layout:
    uemda_b200.gast.alignment          Aligner, DownscaleLabel        (uemda/gast/alignment.py)
    uemda_b200.gast.pseudo_generation  pseudo_selection(1)            (uemda/gast/pseudo_generation.py)
    uemda_b200.gast.balance            UVEMLoss, UPSLoss, ClassBalance, loss_calc_uvem (uemda/gast/balance.py)
    uemda_b200.scatter                 scatter                        (torch_scatter.scatter seam)
    uemda_b200.mining                  fused refine->select step, batch-sharded multi-GPU driver
All arithmetic of the path runs in libuem_b200.so (include/uem_b200.h); there is no CPU fallback.
"""
from . import config  # noqa: F401

__version__ = "0.1.0"
