#!/usr/bin/env python
"""Diagnostic: entropy kernel error vs the fp32 torch reference, bucketed by the dominant probability."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uemda_b200 import ops, _lib
_lib.load()
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(1)
c = 6
z = torch.randn(4, c, 256, 256, generator=g) * torch.linspace(0.2, 12, 256).view(1, 1, 1, 256)
p = torch.softmax(z, 1)
ref32 = -(p * torch.log(p)).sum(1)
ref64 = -(p.double() * torch.log(p.double())).sum(1)
ent, _ = ops.entropy_uvem_weight(p.to(dev), 0.2, 0.7, 4.0)
ent = ent.cpu().reshape(ref32.shape)
vmax = p.max(1)[0]
rel = ((ent.double() - ref32.double()).abs() / ref32.double().abs().clamp_min(1e-30))
rel64 = ((ent.double() - ref64).abs() / ref64.abs().clamp_min(1e-30))
r3264 = ((ref32.double() - ref64).abs() / ref64.abs().clamp_min(1e-30))
edges = [0, 0.3, 0.5, 0.7, 0.9, 0.97, 0.984375, 0.99, 0.999, 0.9999, 0.99999, 1.0000001]
for lo, hi in zip(edges[:-1], edges[1:]):
    m = (vmax >= lo) & (vmax < hi) & torch.isfinite(ref32)
    if m.any():
        print("vmax [%.6f,%.6f) n=%7d  rel vs f32 ref max %.2e  vs f64 max %.2e   (f32 ref vs f64 max %.2e)" % (
            lo, hi, int(m.sum()), float(rel[m].max()), float(rel64[m].max()), float(r3264[m].max())))
m = torch.isfinite(ref32)
i = int(torch.where(m, rel, torch.zeros_like(rel)).argmax())
pv = p.permute(0, 2, 3, 1).reshape(-1, c)[i]
print("worst:", pv.tolist(), "got", float(ent.reshape(-1)[i]), "ref32", float(ref32.reshape(-1)[i]), "ref64", float(ref64.reshape(-1)[i]))
