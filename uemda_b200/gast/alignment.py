"""Prototype bank + pseudo-label refinement (drop-in for the hot-path part of uemda/gast/alignment.py).

``Aligner`` and ``DownscaleLabel`` keep the reference's constructor and method signatures
(alignment.py:26, :86-126, :175-309, :328-355, :424-509).  Everything on the mining path runs in
hand-written sm_100a kernels through ``uemda_b200.ops``; the grad-carrying alignment losses of the
reference (align_domain/align_class/align_instance/whiten_class_ware, alignment.py:79-170,357-422)
are outside that path and are not provided here.
"""
import torch
import torch.nn as nn

from .. import config, ops
from .pseudo_generation import pseudo_selection  # noqa: F401  (re-exported like the reference module)

__all__ = ["Aligner", "DownscaleLabel"]

_OUT_OF_SCOPE = ("%s is a grad-carrying alignment loss outside the pseudo-label mining path; "
                 "keep using the reference implementation (uemda/gast/alignment.py) for it")


def _cuda_device():
    if not torch.cuda.is_available():
        raise RuntimeError("uemda_b200 needs a CUDA device: there is no CPU fallback for the mining path")
    return torch.device("cuda", torch.cuda.current_device())


class Aligner:

    def __init__(self, logger, feat_channels=64, class_num=7, ignore_label=-1, decay=0.999, topk=32, resume=None):
        dev = _cuda_device()
        self.feat_channels = feat_channels
        self.class_num = class_num
        self.ignore_label = ignore_label
        self.decay = decay
        self.logger = logger
        self.eps = 1e-7
        self.topk = topk
        self.topk_importance = (1 - torch.arange(topk, dtype=torch.float, device=dev) / topk).reshape(1, topk, 1)

        # class prototypes (c,k), computed from source features only (alignment.py:54-60)
        if resume:
            self.prototypes = torch.load(resume, map_location="cpu").to(dev)
            self.logger.info('finish init prototypes!')
            self.logger.info(f'prototypes({self.prototypes.shape})={self.prototypes}')
        else:
            self.prototypes = torch.zeros([class_num, feat_channels], device=dev)

        self.downscale_gt = DownscaleLabel(scale_factor=16, n_classes=class_num, ignore_label=ignore_label,
                                           min_ratio=0.75)
        self._data_sum = torch.zeros([class_num, feat_channels], device=dev)
        self._data_cnt = torch.zeros([class_num, 1], device=dev)

        # Optional, not in the reference: an upper bound on superpixel ids (+1).  When set, label_refine
        # sizes its region table from it instead of reading label_t_sup.max() back (no host sync).
        self.num_regions = None

    # ------------------------------------------------------------------ prototypes (a10-a12)
    def update_prototype(self, feat, label):
        """EMA-update the prototypes from source features and full-resolution labels; returns the
        down-scaled labels (b,1,h,w) (alignment.py:86-90)."""
        label = self.downscale_gt(label)
        self._compute_local_prototypes(feat, label, update=True, decay=self.decay)
        return label

    def update_prototype_bytarget(self, feat_t, label_t_soft):
        """Soft-label weighted mean of target features, then EMA (alignment.py:92-105)."""
        b, k, h, w = feat_t.shape
        sums = ops.proto_accumulate_soft(feat_t, label_t_soft)
        _, new = ops.proto_finalize(sums, None, self.prototypes, eps=self.eps, decay=self.decay, mean_n=b * h * w,
                                    want_local=False)
        self.prototypes = new

    def update_avg(self, feat, label) -> None:
        """Accumulate per-class feature sums / counts over the whole source set (alignment.py:107-119)."""
        labels = self.downscale_gt(label)
        sums, counts = ops.proto_accumulate(feat, labels, self.class_num, self.ignore_label)
        self._data_sum = self._data_sum + sums
        self._data_cnt = self._data_cnt + counts.reshape(self.class_num, 1).to(self._data_cnt.dtype)

    def init_avg(self):
        """prototypes = sum / (count + eps) (alignment.py:121-126); a (c,k) one-off, done with two torch ops."""
        self.prototypes = self._data_sum / (self._data_cnt + self.eps)
        self.logger.info('finish init prototypes!')
        self.logger.info(f'examples cnt({self._data_cnt.shape})={self._data_cnt}')
        self.logger.info(f'prototypes({self.prototypes.shape})={self.prototypes}')

    def _compute_local_prototypes(self, feat, label, update=False, decay=0.999):
        """Masked per-class feature mean within a mini-batch (alignment.py:328-355).
        feat (b,k,h,w), label (b,1,h,w) -> (c,k); classes without pixels keep the global prototype."""
        assert 0 < decay < 1
        sums, counts = ops.proto_accumulate(feat, label, self.class_num, self.ignore_label)
        local, new = ops.proto_finalize(sums, counts, self.prototypes, eps=self.eps, decay=decay if update else None)
        if update:
            self.prototypes = new
        return local

    # ------------------------------------------------------------------ superpixel relabel (a7)
    def superpixel_expand(self, label_t_hard, label_t_sup):
        """Majority class of every superpixel broadcast back to its pixels (alignment.py:175-192).
        label_t_hard (b,H,W), label_t_sup (b,1,H,W) -> (b,H,W) int64; regions without labels -> -1."""
        return ops.superpixel_expand(label_t_hard, label_t_sup, self.class_num, self.ignore_label,
                                     num_regions=self.num_regions)

    # ------------------------------------------------------------------ label_refine (a6)
    def label_refine(self, label_t_sup, feat_t, preds_t, label_t_soft, refine=True, mode='all', temp=2.0):
        """Re-weight the soft pseudo-labels by the prototype / prediction / superpixel views and
        renormalise (alignment.py:194-293).  Returns (b,c,H,W) fp32."""
        assert mode in ['all', 's', 'p', 'n', 'l']
        if not refine:
            return label_t_soft
        if mode == 'n':
            raise NotImplementedError("label_refine(mode='n') (O(n^2) neighbour view, marked 'x' in the reference, "
                                      "alignment.py:260-286) is outside the mining path")
        from .. import mining
        views = ops.MODE_VIEWS[mode]
        pred1 = pred2 = None
        if views & ops.VIEW_PRED:
            if isinstance(preds_t, list):
                assert len(preds_t) == 2
                pred1, pred2 = preds_t
            else:
                pred1 = preds_t
        refined, _ = mining.refine_select(
            views, label_t_soft, temp, feat=feat_t if views & ops.VIEW_PROTO else None, prototypes=self.prototypes,
            pred1=pred1, pred2=pred2, sup=label_t_sup if views & ops.VIEW_SUP else None, num_regions=self.num_regions,
            eps=self.eps, select=None)
        return refined

    def get_prototype_weight_4pixel(self, feats, label_hard, temp=2.0):
        """Prototype-view weight of each pixel's own hard-label class (alignment.py:295-309) -> (b*H*W,)."""
        simi = ops.pearson_dist_nchw(feats, self.prototypes, eps=self.eps, reciprocal=True)
        return ops.proto_weight_4pixel(simi, label_hard, self.ignore_label, self.eps)

    def _pearson_dist(self, feat1, feat2):
        """(n,k),(m,k) -> (n,m) Pearson distance in [0,1] (alignment.py:424-451)."""
        assert feat1.shape[-1] == feat2.shape[-1]
        return ops.pearson_dist_rows(feat1, feat2, eps=self.eps)

    # ------------------------------------------------------------------ small helpers kept for API parity
    @staticmethod
    def _softmax_T(feat, temp=1.0, dim=1):
        assert temp > 0
        return torch.softmax(feat / temp, dim=dim)

    def _logits_norm(self, logits):
        assert len(logits.shape) == 4
        return logits / (torch.sum(logits, dim=1, keepdim=True) + self.eps)

    @staticmethod
    def _ema(history, curr, decay=0.999):
        return (1.0 - decay) * curr + decay * history

    def _index2onehot(self, label):
        """(b,1,h,w)|(b,h,w) int64 -> (b*h*w, c) int64 one-hot, ignore label -> zero row (alignment.py:468-481)."""
        flat = (label if label.dim() == 4 else label.unsqueeze(1)).permute(0, 2, 3, 1).reshape(-1)
        classes = torch.arange(self.class_num, device=flat.device)
        return (flat.unsqueeze(1) == classes.unsqueeze(0)).long()

    def show(self, save_path=None, display=True):
        pass

    # ------------------------------------------------------------------ outside the mining path
    def align_domain(self, feat_s, feat_t):
        raise NotImplementedError(_OUT_OF_SCOPE % "align_domain")

    def align_class(self, feat_s, label_s, feat_t=None, label_t=None):
        raise NotImplementedError(_OUT_OF_SCOPE % "align_class")

    def align_instance(self, feat_s, label_s, feat_t=None, label_t=None):
        raise NotImplementedError(_OUT_OF_SCOPE % "align_instance")

    def whiten_class_ware(self, feat_s, label_s, feat_t=None, label_t=None):
        raise NotImplementedError(_OUT_OF_SCOPE % "whiten_class_ware")


class DownscaleLabel(nn.Module):
    """Block-majority label down-scaling (alignment.py:484-509): a cell keeps its majority class only if
    that class covers >= min_ratio of the scale_factor^2 block and is not the ignore label."""

    def __init__(self, scale_factor=16, n_classes=7, ignore_label=-1, min_ratio=0.75):
        super().__init__()
        assert scale_factor > 1
        self.scale_factor = scale_factor
        self.n_classes = n_classes
        self.ignore_label = ignore_label
        self.min_ratio = min_ratio

    def forward(self, label):
        if label.dim() == 4:
            label = label.squeeze(dim=1)
        assert len(label.shape) == 3
        status = torch.zeros(1, dtype=torch.int32, device=label.device) if config.strict_asserts else None
        out = ops.downscale_label(label, self.scale_factor, self.n_classes, self.ignore_label, self.min_ratio, status)
        if status is not None and int(status.item()) & 1:
            # F.one_hot in the reference (alignment.py:502) rejects these
            raise RuntimeError("Class values must be non-negative and smaller than num_classes.")
        return out
