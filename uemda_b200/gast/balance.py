"""Uncertainty weighting and class balancing (drop-in for the hot-path part of uemda/gast/balance.py).

The detached, per-pixel factors -- entropy of the soft label, the gate u > threshold, the
piecewise-parabolic UVEM weight, the class histogram / frequency EMA / per-pixel class weight -- run in
sm_100a kernels.  The cross-entropy itself carries the gradient and stays a PyTorch op, exactly as
SURVEY.md section 2 (row 3) scopes it.
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


def loss_calc_uvem(pred, label, label_soft, loss_fn, multi=True):
    """balance.py:437-457: apply loss_fn to one head or average it over several, up-sampling the logits
    to the label resolution first."""
    heads = list(pred) if multi is True else [pred]
    total = 0
    for p in heads:
        if p.size()[-2:] != label.size()[-2:]:
            p = tnf.interpolate(p, size=label.size()[-2:], mode='bilinear', align_corners=True)
        total = total + loss_fn(p, label.long(), label_soft)
    return total / len(heads) if multi is True else total
