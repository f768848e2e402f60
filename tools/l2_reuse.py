#!/usr/bin/env python
"""Does data one kernel touched stay in L2 for the next one?  (development aid)

Pairs of kernels captured into one CUDA graph, rotating over --sets input copies; "same" = the second kernel reads what
the first one just read / wrote, "other" = the identical kernel pair where the second one works on a set that was last
touched several pairs ago.  The difference is the L2 reuse the chain can get at this batch size.
"""
import argparse
import dataclasses
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from uemda_b200 import _lib, config, ops  # noqa: E402
from uemda_b200.synth import WORKLOADS, make_inputs  # noqa: E402


def timed(fn_list, iters):
    """fn_list: callables captured back to back into one graph; returns us per graph replay / len(fn_list)."""
    for f in fn_list:
        f()
    torch.cuda.synchronize()
    keep = []
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fn_list:
            keep.append(f())
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (iters * len(fn_list))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="2,4,8")
    ap.add_argument("--sets", type=int, default=6)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--opt", action="append", default=[])
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    for kv in args.opt:
        name, val = kv.split("=")
        _lib.check(lib.uem_set_option(name.encode(), int(val)))
    config.strict_asserts = False
    # control: a plain reduction over a hot / cold buffer
    for mib in (16, 32, 64, 96):
        n = mib * 2 ** 20 // 4
        bufs = [torch.randn(n, device=dev) for _ in range(max(2, 512 // mib))]
        hot = timed([lambda: bufs[0].sum() for _ in range(8)], args.iters)
        cold = timed([(lambda b=b: b.sum()) for b in bufs], args.iters)
        print("control sum %3d MiB: hot %.2f us (%.0f GB/s)  cold %.2f us (%.0f GB/s)" % (mib, hot, n * 4 / hot / 1e3, cold, n * 4 / cold / 1e3), flush=True)
        del bufs
    base = WORKLOADS["cfg2_isprs_8x6x512"]
    for b in [int(x) for x in args.batches.split(",")]:
        wl = dataclasses.replace(base, b=b, name="cfg2_b%d" % b)
        inp = make_inputs(wl, seed=2333)
        keys = ("soft", "sup", "feat", "pred1", "pred2")
        S = args.sets
        sets = [{k: torch.roll(inp[k], i, 0).to(dev) for k in keys} for i in range(S)]
        protos = inp["prototypes"].to(dev)
        R = int(inp["ignore_id"]) + 1
        for s in sets:
            s["simi"] = ops.pearson_dist_nchw(s["feat"], protos, reciprocal=True)
            s["rmax"] = ops.region_reduce(s["soft"], s["sup"], "max", dim_size=R, planar=True)
            s["ign"] = ops.i64_minmax(s["sup"])[1:].clone()
            s["refined"], s["stats"] = ops.label_refine(7, s["soft"], 2.0, simi=s["simi"], pred1=s["pred1"], pred2=s["pred2"],
                                                        sup=s["sup"], region_max=s["rmax"], ignored_id=s["ign"])
            s["out"] = torch.empty_like(s["refined"])

        def region(s):
            return ops.region_reduce(s["soft"], s["sup"], "max", dim_size=R, planar=True)

        def refine(s):
            return ops.label_refine(7, s["soft"], 2.0, simi=s["simi"], pred1=s["pred1"], pred2=s["pred2"], sup=s["sup"],
                                    region_max=s["rmax"], ignored_id=s["ign"])

        def select(refined, stats):
            return ops.pseudo_select_stats(refined, stats, 0.8, 0.6, -1, uvem=(0.2, 0.7, 4.0))

        half = S // 2
        res = {}
        res["region->refine same"] = timed([f for i in range(S) for f in (lambda i=i: region(sets[i]), lambda i=i: refine(sets[i]))], args.iters) * 2
        res["region->refine other"] = timed([f for i in range(S) for f in (lambda i=i: region(sets[i]), lambda i=i: refine(sets[(i + half) % S]))], args.iters) * 2

        def ref_sel_same(i):
            r, st = refine(sets[i])
            return select(r, st)

        def ref_sel_other(i):
            r, st = refine(sets[i])
            o = sets[(i + half) % S]
            return r, select(o["refined"], o["stats"])

        res["refine->select same"] = timed([(lambda i=i: ref_sel_same(i)) for i in range(S)], args.iters)
        res["refine->select other"] = timed([(lambda i=i: ref_sel_other(i)) for i in range(S)], args.iters)
        mb = wl.pixels * (4 * wl.c + 8) / 2 ** 20
        print("b=%d (soft+ids %.0f MiB, refined %.0f MiB): " % (b, mb, wl.pixels * 4 * wl.c / 2 ** 20) +
              "  ".join("%s %.1f us" % kv for kv in res.items()), flush=True)
        del sets, inp
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
