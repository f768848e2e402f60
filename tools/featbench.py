#!/usr/bin/env python
"""Development aid: feature-map kernels (pearson / proto_accum) vs tensor shape at constant bytes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from uemda_b200 import _lib, ops

dev = torch.device("cuda", 0)
_lib.load()
c, k = 6, 2048
protos = torch.randn(c, k, device=dev)


def timeit(fn, sets, iters=40):
    for s in sets:
        fn(s)
    torch.cuda.synchronize()
    graphs = []
    keep = []
    for s in sets:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep.append(fn(s))
        graphs.append(g)
    for g in graphs:
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        graphs[i % len(graphs)].replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


for (b, h, w) in [(64, 8, 16), (32, 16, 16), (8, 32, 32), (2, 64, 64), (8, 64, 64)]:
    sets = []
    for i in range(4):
        f = torch.randn(b, k, h, w, device=dev)
        lab = torch.randint(-1, c, (b, 1, h, w), device=dev)
        sets.append((f, lab))
    nbytes = b * k * h * w * 4
    t1 = timeit(lambda s: ops.pearson_dist_nchw(s[0], protos, reciprocal=True), sets)
    t2 = timeit(lambda s: ops.proto_accumulate(s[0], s[1], c), sets)
    t3 = timeit(lambda s: s[0].sum(), sets)
    print("feat (%d,%d,%d,%d) %.0f MB: pearson %.1f us (%.0f GB/s)  proto_accum %.1f us (%.0f GB/s)  torch.sum %.1f us (%.0f GB/s)" % (
        b, k, h, w, nbytes / 1e6, t1, nbytes / t1 / 1e3, t2, nbytes / t2 / 1e3, t3, nbytes / t3 / 1e3), flush=True)
