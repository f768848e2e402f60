"""Deterministic synthetic inputs for the mining path.

Shapes follow the reference's data contract: soft labels ``(b,c,H,W)`` f32 loaded from
``.pt`` (uemda/datasets/basedata.py:87), superpixel maps ``(b,1,H,W)`` int64
(basedata.py:77-79) produced by LSC + 7x7 edge shrink (uemda/gast/superpixels.py:129-150),
decoder features ``(b,k,H/s,W/s)`` f32 and two logit heads ``(b,c,H/s,W/s)``
(uemda/models/Encoder.py:150-151), source labels ``(b,H,W)`` int64 in [-1,c).

Everything is generated with a CPU ``torch.Generator`` (seed 2333 = the reference's own
seed, uemda/utils/tools.py:305) so the CPU oracle and the CUDA path see identical bits.
"""
import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

SEED = 2333


@dataclass
class Workload:
    name: str
    b: int
    c: int
    H: int
    W: int
    k: int = 2048
    scale: int = 16
    regions: int = 1024  # superpixels per image (ignore id == regions)

    @property
    def h(self):
        return self.H // self.scale

    @property
    def w(self):
        return self.W // self.scale

    @property
    def pixels(self):
        return self.b * self.H * self.W


# BASELINE.json configs[0..4]
WORKLOADS = {
    "cfg1_cpu_2x6x512": Workload("cfg1_cpu_2x6x512", 2, 6, 512, 512, 2048, 16, 1024),
    "cfg2_isprs_8x6x512": Workload("cfg2_isprs_8x6x512", 8, 6, 512, 512, 2048, 16, 1024),
    "cfg2_isprs_8x6x512_os8": Workload("cfg2_isprs_8x6x512_os8", 8, 6, 512, 512, 2048, 8, 1024),
    "cfg3_loveda_16x7x1024": Workload("cfg3_loveda_16x7x1024", 16, 7, 1024, 1024, 2048, 16, 2048),
    "cfg4_potsdam_tiles": Workload("cfg4_potsdam_tiles", 2016, 6, 512, 512, 2048, 16, 1024),
    "cfg5_sweep_32x6x512": Workload("cfg5_sweep_32x6x512", 32, 6, 512, 512, 2048, 16, 1024),
    "tiny": Workload("tiny", 2, 6, 64, 96, 64, 16, 24),
}


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def voronoi_superpixels(H, W, regions, rng, shrink=True, win=3):
    """Jittered-grid Voronoi region map with ids 0..R-1, then the reference's edge shrink:
    a pixel keeps its id iff its whole (2*win+1)^2 window (clipped to the image) carries the
    same id, else it gets the ignore id R (superpixels.py:129-150, vectorised)."""
    from scipy import ndimage

    gy = max(1, int(round(math.sqrt(regions * H / W))))
    gx = max(1, int(math.ceil(regions / gy)))
    while gy * gx > regions and gx > 1:
        gx -= 1
    R = gy * gx
    cy, cx = H / gy, W / gx
    sy = (np.arange(gy)[:, None] + 0.5 + rng.uniform(-0.35, 0.35, (gy, gx))) * cy
    sx = (np.arange(gx)[None, :] + 0.5 + rng.uniform(-0.35, 0.35, (gy, gx))) * cx
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    iy = np.minimum((yy / cy).astype(np.int64), gy - 1)
    ix = np.minimum((xx / cx).astype(np.int64), gx - 1)
    best = np.full((H, W), np.inf, dtype=np.float32)
    lab = np.zeros((H, W), dtype=np.int64)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            jy = np.clip(iy + dy, 0, gy - 1)
            jx = np.clip(ix + dx, 0, gx - 1)
            d = (yy - sy[jy, jx]) ** 2 + (xx - sx[jy, jx]) ** 2
            upd = d < best
            best = np.where(upd, d, best)
            lab = np.where(upd, jy * gx + jx, lab)
    if shrink:
        size = 2 * win + 1
        mx = ndimage.maximum_filter(lab, size=size, mode="nearest")
        mn = ndimage.minimum_filter(lab, size=size, mode="nearest")
        lab = np.where((mx == lab) & (mn == lab), lab, R)
    return lab.astype(np.int64), R


def make_superpixels(b, H, W, regions, seed=SEED, shrink=True):
    rng = np.random.default_rng(seed)
    maps = []
    R = regions
    for _ in range(b):
        m, R = voronoi_superpixels(H, W, regions, rng, shrink=shrink)
        maps.append(m)
    return torch.from_numpy(np.stack(maps)).unsqueeze(1), R


def make_logits(b, c, H, W, seed=SEED, smooth=16):
    """z = 3*N(0,1)/2 + smooth class field so argmax forms blobs and confidences span (1/c,1)."""
    g = _gen(seed)
    hs, ws = max(1, H // smooth), max(1, W // smooth)
    field = torch.randn(b, c, hs, ws, generator=g) * 4.0
    field = F.interpolate(field, size=(H, W), mode="bilinear", align_corners=True)
    return field + 1.5 * torch.randn(b, c, H, W, generator=g)


def make_inputs(wl: Workload, seed=SEED, b=None, shrink=True, with_source=True):
    """Returns a dict of CPU tensors for one mining step of workload ``wl``."""
    b = wl.b if b is None else b
    c, H, W, k, h, w = wl.c, wl.H, wl.W, wl.k, wl.h, wl.w
    g = _gen(seed + 1)
    z = make_logits(b, c, H, W, seed=seed)
    soft = torch.softmax(z, dim=1)
    z_low = F.adaptive_avg_pool2d(z, (h, w))
    pred1 = z_low + 0.5 * torch.randn(b, c, h, w, generator=g)
    pred2 = z_low + 0.5 * torch.randn(b, c, h, w, generator=g)
    protos = torch.randn(c, k, generator=g)
    cls_low = z_low.argmax(dim=1)  # (b,h,w)
    feat = torch.randn(b, k, h, w, generator=g)
    feat += 0.5 * protos[cls_low].permute(0, 3, 1, 2)
    mu = feat.mean(dim=(2, 3), keepdim=True)
    sd = feat.std(dim=(2, 3), keepdim=True)
    feat = (feat - mu) / (sd + 1e-5)
    sup, R = make_superpixels(b, H, W, wl.regions, seed=seed + 2, shrink=shrink)
    out = dict(soft=soft.contiguous(), pred1=pred1.contiguous(), pred2=pred2.contiguous(),
               feat=feat.contiguous(), sup=sup.contiguous(), prototypes=protos.contiguous(),
               ignore_id=R, logits=z)
    if with_source:
        hs, ws = max(1, H // 32), max(1, W // 32)
        lab = torch.randint(0, c, (b, 1, hs, ws), generator=g).float()
        drop = torch.rand(b, 1, hs, ws, generator=g) < 0.05
        lab[drop] = -1
        lab = F.interpolate(lab, size=(H, W), mode="nearest").squeeze(1).long()
        # sprinkle pixel noise so some 16x16 blocks fall below the 0.75 majority ratio
        noise = torch.rand(b, H, W, generator=g) < 0.12
        rnd = torch.randint(-1, c, (b, H, W), generator=g)
        lab = torch.where(noise, rnd, lab)
        feat_s = torch.randn(b, k, h, w, generator=g)
        out.update(label_s=lab.contiguous(), feat_s=feat_s.contiguous())
    return out
