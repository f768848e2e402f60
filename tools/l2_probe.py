#!/usr/bin/env python
"""DRAM bytes per kernel of the one-call mining chain with and without the L2 eviction-priority hints (development aid).

Run under `ncu --cache-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` so that
the cache state one kernel leaves is what the next one finds; inputs rotate over three sets (nothing is warm from the
previous call).  First half of the calls: hints off; second half: library defaults.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from uemda_b200 import _lib, config, mining  # noqa: E402
from uemda_b200.synth import WORKLOADS, make_inputs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2_isprs_8x6x512")
    ap.add_argument("--calls", type=int, default=6)
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    config.strict_asserts = False
    inp = make_inputs(wl, seed=2333)
    keys = ("soft", "sup", "feat", "pred1", "pred2")
    sets = [{k: torch.roll(inp[k], i, 0).to(dev) for k in keys} for i in range(3)]
    protos = inp["prototypes"].to(dev)
    R = int(inp["ignore_id"]) + 1
    ws = torch.zeros(lib.uem_mine_ws_bytes(wl.b, wl.c, wl.H, wl.W, wl.h, wl.w, wl.k, R), dtype=torch.uint8, device=dev)

    def chain(s):
        return mining.refine_select(7, s["soft"], 2.0, feat=s["feat"], prototypes=protos, pred1=s["pred1"], pred2=s["pred2"],
                                    sup=s["sup"], num_regions=R, select=(0.8, 0.6, -1), ws=ws, uvem=(0.2, 0.7, 4.0))

    for mode in ("off", "default"):
        for name in ("l2_stream", "l2_last_use"):
            _lib.check(lib.uem_set_option(name.encode(), 0 if mode == "off" else 1))
        for i in range(args.calls):
            chain(sets[i % 3])
        torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
