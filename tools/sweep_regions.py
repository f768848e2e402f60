#!/usr/bin/env python
"""BASELINE config 5: superpixel region-count sweep (256..16k regions per image, batch 32, 6 classes, 512x512):
scatter-reduction contention stress for the region-max seam, superpixel_expand and the fused refine+select chain.
Edge-shrunk maps (hot "ignored" id) up to 1k regions, un-shrunk maps beyond (regions smaller than the 7x7 window)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from uemda_b200 import _lib, config, mining, ops  # noqa: E402
from uemda_b200.synth import make_inputs, make_superpixels, Workload  # noqa: E402


def timeit(fn, reps=12):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    keep = []
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            keep.append(fn())
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (3 * reps)


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    _lib.load()
    config.strict_asserts = False
    b, c, H, W = 32, 6, 512, 512
    wl = Workload("cfg5", b, c, H, W, 256, 16, 1024)
    inp = make_inputs(wl, seed=3, with_source=False)
    soft = inp["soft"].to(dev)
    feat, p1, p2, protos = inp["feat"].to(dev), inp["pred1"].to(dev), inp["pred2"].to(dev), inp["prototypes"].to(dev)
    hard = soft.argmax(1)
    P = b * H * W
    print("%-8s %-7s %12s %12s %12s %14s" % ("regions", "shrunk", "region_max", "expand", "refine+select", "chain GB/s(alg)"))
    for regions in (256, 1024, 4096, 16384):
        shrink = regions <= 1024
        sup, R = make_superpixels(b, H, W, regions, seed=5, shrink=shrink)
        sup = sup.to(dev)
        Rcap = int(sup.max()) + 1
        t1 = timeit(lambda: ops.region_reduce(soft, sup, "max", dim_size=Rcap, planar=True))
        t2 = timeit(lambda: ops.superpixel_expand(hard, sup, c, num_regions=Rcap))
        t3 = timeit(lambda: mining.refine_select(7, soft, 2.0, feat=feat, prototypes=protos, pred1=p1, pred2=p2, sup=sup,
                                                 num_regions=Rcap, select=(0.8, 0.6, -1)))
        chain_bytes = P * (8 * c + 8 + 8 + 4 * c) + feat.numel() * 4
        print("%-8d %-7s %9.1f us %9.1f us %10.1f us %14.0f" % (Rcap, shrink, t1, t2, t3, chain_bytes / t3 / 1e3), flush=True)


if __name__ == "__main__":
    main()
