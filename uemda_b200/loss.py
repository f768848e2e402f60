"""Prototype contrastive loss (drop-in for ``uemda.loss.PrototypeContrastiveLoss``, uemda/loss.py:10-47).

Same constructor and ``forward(Proto, feat, labels)`` signature; ``feat`` carries the gradient.  Forward and backward
each make ONE pass over the NCHW feature map in hand-written sm_100a kernels (TMA-tiled streaming, see
csrc/uem_pcl.cu); the reference's permute / boolean-mask / normalize / mm / CrossEntropy chain and its autograd replay
are not materialised.  CUDA tensors only.
"""
import torch
import torch.nn as nn

from . import ops

__all__ = ["PrototypeContrastiveLoss"]


class _PCLFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, feat, proto, labels, temperature, ignore_label):
        loss, coef, ws = ops.pcl_forward(feat, proto, labels, temperature, ignore_label)
        ctx.save_for_backward(feat, coef, ws)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        feat, coef, ws = ctx.saved_tensors
        # the kernels compute in fp32; a half-precision feature map (autocast) gets its gradient back in its own dtype
        return ops.pcl_backward(feat, coef, ws, grad_out).to(feat.dtype), None, None, None, None


class PrototypeContrastiveLoss(nn.Module):

    def __init__(self, temperature=8.0, ignore_label=-1):
        super().__init__()
        self.temperature = temperature
        self.ignore_label = ignore_label

    def forward(self, Proto, feat, labels):
        """Proto (C, A) class prototypes (no grad), feat (b, A, h, w) [or (N, A)] with grad, labels (b,1,h,w) / (N,)."""
        assert not Proto.requires_grad and not labels.requires_grad and feat.requires_grad
        if feat.dim() == 2:
            # (N, A) rows: view as one NCHW image of N pixels (a transposed copy: the kernels stream channel-major maps)
            n, k = feat.shape
            pad = (-n) % 4
            rows = feat.t().reshape(1, k, 1, n)
            lab = labels.reshape(1, n)
            if pad:
                rows = torch.nn.functional.pad(rows, (0, pad))
                lab = torch.nn.functional.pad(lab, (0, pad), value=self.ignore_label)
            return _PCLFunction.apply(rows.contiguous(), Proto, lab, self.temperature, self.ignore_label)
        assert feat.dim() == 4
        b, k, h, w = feat.shape
        if (h * w) % 4:
            # the TMA-tiled kernels need 16-byte aligned pixel rows: odd maps (33x33, 65x65) take the padded row form (all
            # images as one row of b*h*w pixels, padding labelled ignore); the reference accepts any shape
            n = b * h * w
            pad = (-n) % 4
            rows = torch.nn.functional.pad(feat.permute(1, 0, 2, 3).reshape(1, k, 1, n), (0, pad))
            lab = torch.nn.functional.pad(labels.reshape(1, n), (0, pad), value=self.ignore_label)
            return _PCLFunction.apply(rows.contiguous(), Proto, lab, self.temperature, self.ignore_label)
        return _PCLFunction.apply(feat, Proto, labels, self.temperature, self.ignore_label)
