"""IAST class-wise thresholds and sliding-window / TTA accumulation (drop-in for the hot-path part of
uemda/utils/tools.py: ``ias_thresh`` :323-333, the per-batch body of ``generate_pseudo`` :347-371, ``pre_slide`` :61-97,
``tta_predict`` :132-152).  CUDA tensors only; the arithmetic runs in csrc/uem_iast.cu.

The reference builds one Python list of float16 confidences per class and calls ``np.percentile`` on it.  Here the same
sample is an exact histogram over the 65536 half bit patterns built in one pass over the probabilities; the percentile's two
order statistics come off its prefix sum and are interpolated with numpy's fp64 formula, so thresholds and labels are
bit-identical to the reference's (tests/golden/iast_small.npz).  The model forward is the caller's; file IO stays outside.
"""
from math import ceil

import numpy as np
import torch

from .. import _lib as L
from .. import ops

__all__ = ["IASTSelector", "ias_thresh", "pre_slide", "tta_predict", "slide_windows"]


def _qfrac(alpha, w, gamma):
    """q / 100 per class with the reference's own host arithmetic: 100 * (1 - alpha * w ** gamma) (tools.py:332), then
    numpy.percentile's true_divide(q, float64(100))."""
    import ctypes
    w = np.asarray(w, dtype=np.float64)
    q = np.array([100 * (1 - alpha * w[i] ** gamma) for i in range(w.shape[0])], dtype=np.float64)
    qf = np.true_divide(q, np.float64(100))
    if not (np.all(qf >= 0) and np.all(qf <= 1)):
        raise ValueError("Percentiles must be in the range [0, 100]")
    return (ctypes.c_double * len(qf))(*[float(v) for v in qf])


class IASTSelector:
    """The state of ``generate_pseudo`` (tools.py:335-373) across batches: ``cls_thresh`` (float64, starts at 0.9).
    ``step(probs)`` -> uint8 labels (class + 1, 0 = ignored), thresholds updated as the reference does per batch."""

    def __init__(self, n_class=7, pl_alpha=0.2, pl_beta=0.9, pl_gamma=8.0, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("uemda_b200 needs a CUDA device: there is no CPU fallback for the mining path")
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.n_class = n_class
        self.alpha, self.beta, self.gamma = pl_alpha, pl_beta, pl_gamma
        self.cls_thresh = torch.full((n_class,), 0.9, dtype=torch.float64, device=self.device)   # tools.py:346
        self.tmp_thresh = torch.zeros(n_class, dtype=torch.float32, device=self.device)
        self.counts = torch.zeros(n_class, dtype=torch.int64, device=self.device)
        self._hist = torch.empty(int(L.load().uem_iast_hist_bytes(n_class)), dtype=torch.uint8, device=self.device)

    def update_thresholds(self, probs):
        """tools.py:349-361 for one batch: per-class confidence histogram -> percentile -> EMA.  One 8*c-byte read-back of
        the previous thresholds (they enter the percentile rank through w ** gamma, which is host arithmetic upstream)."""
        L.require_cuda(probs)
        probs = L.f32c(probs.detach())
        b, c = probs.shape[:2]
        assert c == self.n_class
        hw = probs[0, 0].numel()
        lib = L.bind(probs)
        st = L.stream_of(probs)
        L.check(lib.uem_iast_conf_hist_f32(L.ptr(probs), b, c, hw, L.ptr(self._hist), st))
        qf = _qfrac(self.alpha, self.cls_thresh.cpu().numpy(), self.gamma)
        L.check(lib.uem_iast_thresholds_f64(L.ptr(self._hist), c, qf, float(self.beta), ops.f32(1 - self.beta), L.ptr(self.cls_thresh),
                                            L.ptr(self.tmp_thresh), L.ptr(self.counts), st))
        return self.cls_thresh

    def labels(self, probs):
        """tools.py:363-372: uint8 (b,H,W): argmax + 1, 0 where the winning probability is below its class threshold."""
        L.require_cuda(probs)
        probs = L.f32c(probs.detach())
        b, c = probs.shape[:2]
        hw = probs[0, 0].numel()
        lib = L.bind(probs)
        out = torch.empty((b,) + tuple(probs.shape[2:]), dtype=torch.uint8, device=probs.device)
        L.check(lib.uem_iast_labels_u8(L.ptr(probs), b, c, hw, L.ptr(self.cls_thresh), L.ptr(out), L.stream_of(probs)))
        return out

    def step(self, probs):
        self.update_thresholds(probs)
        return self.labels(probs)


def ias_thresh(conf_dict, n_class, alpha, w=None, gamma=1.0):
    """tools.py:323-333 with the reference's signature; ``conf_dict[c]`` is a 1-D CUDA tensor (or a list) of confidences or
    None.  Returns float32 ndarray (n_class,).  (The batch path, IASTSelector, never materialises these lists.)"""
    if w is None:
        w = np.ones(n_class)
    out = np.ones(n_class, dtype=np.float32)
    for c in range(n_class):
        if conf_dict[c] is None:
            continue
        v = conf_dict[c]
        if not torch.is_tensor(v):
            v = torch.as_tensor(np.asarray(v, dtype=np.float64))
        v = v.to(torch.float64).reshape(-1)
        if not v.is_cuda:
            raise L.UemLibraryError("uemda_b200 runs on CUDA tensors only; there is no CPU fallback for the mining path")
        q = np.true_divide(np.float64(100 * (1 - alpha * w[c] ** gamma)), np.float64(100))
        if not 0 <= q <= 1:
            raise ValueError("Percentiles must be in the range [0, 100]")
        srt = torch.sort(v)[0]   # generic samples (not float16): the order statistics come from a device sort
        n = srt.numel()
        virt = (n - 1) * q
        if virt >= n - 1:
            out[c] = float(srt[-1])
            continue
        prev = int(np.floor(virt))
        a, bq = (float(x) for x in srt[prev:prev + 2].cpu())
        g = virt - prev
        r = a + (bq - a) * g
        if g >= 0.5:
            r = bq - (bq - a) * (1 - g)
        out[c] = r
    return out


def slide_windows(image_hw, tile_size=(512, 512)):
    """window list of pre_slide (tools.py:62-80): overlap 1/2, windows clamped to the image -> [(y1, x1, y2, x2)]"""
    H, W = image_hw
    stride = ceil(tile_size[0] * (1 - 1 / 2))
    rows = int(ceil((H - tile_size[0]) / stride) + 1)
    cols = int(ceil((W - tile_size[1]) / stride) + 1)
    wins = []
    for r in range(rows):
        for c in range(cols):
            x1, y1 = int(c * stride), int(r * stride)
            x2, y2 = min(x1 + tile_size[1], W), min(y1 + tile_size[0], H)
            x1, y1 = max(int(x2 - tile_size[1]), 0), max(int(y2 - tile_size[0]), 0)
            wins.append((y1, x1, y2, x2))
    return wins


def window_accumulate(full, count, tile, window):
    """full[:, :, y1:y2, x1:x2] += tile[:, :, :y2-y1, :x2-x1]; count[:, :, y1:y2, x1:x2] += 1 (tools.py:94-95), in place."""
    L.require_cuda(full, count, tile)
    b, c, H, W = full.shape
    th, tw = tile.shape[-2:]
    y1, x1, y2, x2 = window
    tile = L.f32c(tile.detach())
    lib = L.bind(full)
    L.check(lib.uem_window_accumulate_f32(L.ptr(full), L.ptr(count), L.ptr(tile), b, c, H, W, th, tw, y1, x1, y2, x2, L.stream_of(full)))


def window_average(full, count):
    """full /= count (tools.py:97), in place."""
    b, c, H, W = full.shape
    lib = L.bind(full)
    L.check(lib.uem_window_average_f32(L.ptr(full), L.ptr(count), b, c, H * W, L.stream_of(full)))
    return full


def pre_slide(model, image, num_classes=7, tile_size=(512, 512), tta=False):
    """tools.py:61-97, same signature: half-overlapping windows through ``model`` (the caller's), their predictions averaged
    per pixel.  The crop / pad of the input is torch slicing; the accumulation and the final division are library kernels."""
    import torch.nn.functional as tnf
    b, _, H, W = image.shape
    full = torch.zeros((b, num_classes, H, W), device=image.device)
    count = torch.zeros((b, 1, H, W), device=image.device)
    for (y1, x1, y2, x2) in slide_windows((H, W), tile_size):
        img = image[:, :, y1:y2, x1:x2]
        rows_missing, cols_missing = tile_size[0] - img.shape[2], tile_size[1] - img.shape[3]
        padded_img = tnf.pad(img, (0, 0, rows_missing, cols_missing), 'constant', 0)   # pad_image, tools.py:54-58
        padded = tta_predict(model, padded_img) if tta is True else model(padded_img)
        window_accumulate(full, count, padded, (y1, x1, y2, x2))
    return window_average(full, count)


def views_mean(views):
    """mean over a list of equally shaped (1, c, h, w) views in list order (tools.py:149-150) -> (1, c, h, w)"""
    stacked = L.f32c(torch.cat([v.detach() for v in views], 0))
    L.require_cuda(stacked)
    n = stacked.shape[0]
    out = torch.empty((1,) + tuple(stacked.shape[1:]), dtype=torch.float32, device=stacked.device)
    lib = L.bind(stacked)
    L.check(lib.uem_views_mean_f32(L.ptr(stacked), n, out.numel(), L.ptr(out), L.stream_of(stacked)))
    return out


def tta_predict(model, img):
    """tools.py:132-152: horizontal flip x rot90 views through ``model`` (the caller's), de-augmented and averaged.  The view
    algebra is what ttach.Compose([HorizontalFlip(), Rotate90([0, 90, 180, 270])]) does, written with torch.flip / rot90."""
    xs = []
    for flip in (False, True):
        for k in (0, 1, 2, 3):
            aug = torch.flip(img, dims=(3,)) if flip else img
            aug = torch.rot90(aug, k, dims=(2, 3))
            x = model(aug)
            x = torch.rot90(x, -k, dims=(2, 3))
            x = torch.flip(x, dims=(3,)) if flip else x
            xs.append(x)
    return views_mean(xs)
