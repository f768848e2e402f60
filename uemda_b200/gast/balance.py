"""Uncertainty weighting and class balancing (drop-in for the hot-path part of uemda/gast/balance.py).

The detached, per-pixel factors -- entropy of the soft label, the gate u > threshold, the
piecewise-parabolic UVEM weight, the class histogram / frequency EMA / per-pixel class weight -- run in
sm_100a kernels.  Through ``loss_calc_uvem`` (the call the training scripts make, with low-resolution
logits) the cross-entropy is fused with the bilinear up-sampling, forward and backward, in
csrc/uem_uvemloss.cu: no up-sampled logits, no NHWC copy, no autograd
replay.  ``UVEMLoss.forward`` / ``UPSLoss.forward`` called directly on full-resolution logits keep the
PyTorch cross-entropy.
"""
import torch
import torch.nn as nn
import torch.nn.functional as tnf

from .. import ops

__all__ = ["ClassBalance", "UVEMLoss", "UPSLoss", "loss_calc_uvem"]


class ClassBalance(nn.Module):
    """balance.py:15-78: running class frequency (EMA of per-batch label histograms) turned into a
    per-class weight softmax((1-freq)/T)/max and looked up per pixel."""

    def __init__(self, class_num=7, ignore_label=-1, decay=0.99, temperature=0.5):
        super().__init__()
        assert temperature > 0
        if not torch.cuda.is_available():
            raise RuntimeError("uemda_b200 needs a CUDA device: there is no CPU fallback for the mining path")
        self.class_num = class_num
        self.ignore_label = ignore_label
        self.decay = decay
        self.temperature = temperature
        self.eps = 1e-7
        self.freq = torch.ones([class_num], device="cuda").float() / class_num
        self._dist = None   # set by shard(): label histograms are summed over the ranks of a batch-sharded job
        self._group = None

    def shard(self, dist=None, group=None):
        """Batch sharded over ranks (one process per GPU): the per-batch label histogram of balance.py:45-52 is over the
        WHOLE batch, so the local histograms are all-reduced (c+1 int64 values) before the frequency EMA moves."""
        self._dist, self._group = dist, group
        return self

    def get_class_weight_4pixel(self, label):
        self.ema_update(label)  # the EMA moves first, then the lookup uses the new table (balance.py:28)
        return ops.class_weight_lookup(label, self._get_class_wight(), self.ignore_label)

    def ema_update(self, label):
        self.freq = self._ema(self.freq, self._local_freq(label), decay=self.decay)

    def _get_class_wight(self):
        prob = torch.softmax((1.0 - self.freq) / self.temperature, dim=0)
        return prob / (prob.max(dim=0, keepdim=True)[0] + self.eps)

    def _local_freq(self, label):
        hist = ops.class_hist(label, self.class_num, self.ignore_label)  # (c+1,) int64, last = #valid
        if self._dist is not None:
            self._dist.all_reduce(hist, op=self._dist.ReduceOp.SUM, group=self._group)
        return hist[:-1].float() / (hist[-1].float() + self.eps)

    @staticmethod
    def _ema(history, curr, decay):
        return (1.0 - decay) * curr + decay * history

    def _one_hot(self, label):
        flat = label.reshape(-1)
        return (flat.unsqueeze(1) == torch.arange(self.class_num, device=flat.device).unsqueeze(0)).long()

    def __str__(self):
        freq = self.freq.cpu().numpy()
        prob = self._get_class_wight().cpu().numpy()
        return ('class frequency: ' + ', '.join(f'{v:.3f}' for v in freq) +
                ';\tselect probability: ' + ', '.join(f'{v:.3f}' for v in prob))


class _FusedUpsampledCE(torch.autograd.Function):
    """sum over heads of  sum_px coef * CE(upsample(x_m))[target] / (valid + 1e-7)  with its analytic backward."""

    @staticmethod
    def forward(ctx, x1, x2, target, coef, valid):
        sums = ops.uvem_loss_forward(x1, x2, target, coef)
        denom = (valid[0] + 1e-7)  # int64 tensor + Python float -> fp32, like the reference (balance.py:383)
        ctx.save_for_backward(x1, x2, target, coef, denom)
        return (sums.float() / denom).sum()

    @staticmethod
    def backward(ctx, grad_out):
        x1, x2, target, coef, denom = ctx.saved_tensors
        g1, g2 = ops.uvem_loss_backward(x1, x2, target, coef, grad_out.float() / denom)
        return g1, g2, None, None, None


def _per_pixel_ce(preds, targets, ignore_label):
    # same values as the reference's permute+reshape form (balance.py:366-370), without the NHWC copy
    return tnf.cross_entropy(preds, targets, reduction='none', ignore_index=ignore_label).reshape(-1)


class UVEMLoss(nn.Module):
    """balance.py:345-423: cross-entropy gated by the soft label's entropy and weighted by
    get_weight(entropy) (and optionally the class-balance weight), averaged over valid pixels."""

    def __init__(self, m=0.1, threshold=0.7, gamma=8.0, class_balancer=None, class_num=7, ignore_label=-1):
        super().__init__()
        self.m = m
        self.threshold = threshold
        self.gamma = gamma
        self.class_balancer = class_balancer
        self.class_num = class_num
        self.ignore_label = ignore_label

    def forward(self, preds, targets, label_t_soft):
        """preds (b,c,h,w) logits, targets (b,h,w) int64, label_t_soft (b,c,h,w) probabilities -> scalar."""
        targets_ = targets.reshape(-1)
        ce_loss = _per_pixel_ce(preds, targets, self.ignore_label)
        weight_uncer, gate, valid_cnt = ops.uvem_terms(label_t_soft, targets_, self.m, self.threshold, self.gamma,
                                                       use_weight=True, ignore_label=self.ignore_label)
        ce_loss = ce_loss.masked_fill(gate, 0.0)  # gated uncertain example removing (balance.py:373)
        if self.class_balancer is not None:
            weight = weight_uncer * self.class_balancer.get_class_weight_4pixel(targets_)
        else:
            weight = weight_uncer
        return (weight * ce_loss).sum() / (valid_cnt[0] + 1e-7)

    def get_weight(self, uncertainties):
        return ops.uvem_weight(uncertainties, self.m, self.threshold, self.gamma)

    def _coef(self, targets, label_t_soft):
        """detached per-pixel factor of the loss: weight, 0 where gated or ignored; and the valid count"""
        targets_ = targets.reshape(-1)
        weight, gate, valid_cnt = ops.uvem_terms(label_t_soft, targets_, self.m, self.threshold, self.gamma,
                                                 use_weight=True, ignore_label=self.ignore_label)
        if self.class_balancer is not None:
            weight = weight * self.class_balancer.get_class_weight_4pixel(targets_)
        return weight.masked_fill(gate | (targets_ == self.ignore_label), 0.0), valid_cnt

    def fused_heads(self, heads, targets, label_t_soft):
        """sum over `heads` (low-resolution logits) of forward(upsample(head), targets, label_t_soft)."""
        return _fused_heads(self, heads, targets, label_t_soft)


class UPSLoss(nn.Module):
    """balance.py:306-342: the gate without the parabolic weight."""

    def __init__(self, threshold=0.7, class_balancer=None, class_num=7, ignore_label=-1):
        super().__init__()
        self.threshold = threshold
        self.class_balancer = class_balancer
        self.class_num = class_num
        self.ignore_label = ignore_label

    def forward(self, preds, targets, label_t_soft):
        targets_ = targets.reshape(-1)
        ce_loss = _per_pixel_ce(preds, targets, self.ignore_label)
        _, gate, valid_cnt = ops.uvem_terms(label_t_soft, targets_, 0.0, self.threshold, 1.0, use_weight=False,
                                            ignore_label=self.ignore_label)
        ce_loss = ce_loss.masked_fill(gate, 0.0)
        if self.class_balancer is not None:
            ce_loss = self.class_balancer.get_class_weight_4pixel(targets_) * ce_loss
        return ce_loss.sum() / (valid_cnt[0] + 1e-7)

    def _coef(self, targets, label_t_soft):
        targets_ = targets.reshape(-1)
        _, gate, valid_cnt = ops.uvem_terms(label_t_soft, targets_, 0.0, self.threshold, 1.0, use_weight=False,
                                            ignore_label=self.ignore_label)
        if self.class_balancer is not None:
            weight = self.class_balancer.get_class_weight_4pixel(targets_)
        else:
            weight = torch.ones(targets_.numel(), dtype=torch.float32, device=targets_.device)
        return weight.masked_fill(gate | (targets_ == self.ignore_label), 0.0), valid_cnt

    def fused_heads(self, heads, targets, label_t_soft):
        return _fused_heads(self, heads, targets, label_t_soft)


def _fused_heads(loss_fn, heads, targets, label_t_soft):
    """The class balancer's frequency EMA moves once per head in the reference (loss_fn is called per head), so with a
    balancer every head gets its own coefficient map; without one the coefficient is shared and two heads go through
    one launch."""
    targets = targets.long()
    total = 0
    if loss_fn.class_balancer is not None:
        for p in heads:
            coef, valid = loss_fn._coef(targets, label_t_soft)
            total = total + _FusedUpsampledCE.apply(p, None, targets, coef, valid)
        return total
    coef, valid = loss_fn._coef(targets, label_t_soft)
    i = 0
    while i < len(heads):
        pair = heads[i:i + 2]
        x2 = pair[1] if len(pair) == 2 and pair[1].shape == pair[0].shape else None
        total = total + _FusedUpsampledCE.apply(pair[0], x2, targets, coef, valid)
        i += 2 if x2 is not None else 1
    return total


def loss_calc_uvem(pred, label, label_soft, loss_fn, multi=True):
    """balance.py:437-457: apply loss_fn to one head or average it over several, up-sampling the logits
    to the label resolution first."""
    heads = list(pred) if multi is True else [pred]
    if hasattr(loss_fn, "fused_heads") and all(p.dim() == 4 and p.is_cuda and p.size()[-2:] != label.size()[-2:] for p in heads):
        total = loss_fn.fused_heads(heads, label, label_soft)   # up-sampling + CE + weights fused, forward and backward
        return total / len(heads) if multi is True else total
    total = 0
    for p in heads:
        if p.size()[-2:] != label.size()[-2:]:
            p = tnf.interpolate(p, size=label.size()[-2:], mode='bilinear', align_corners=True)
        total = total + loss_fn(p, label.long(), label_soft)
    return total / len(heads) if multi is True else total
