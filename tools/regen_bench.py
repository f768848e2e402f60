#!/usr/bin/env python
"""BASELINE config 4: pseudo-label regeneration of a whole target set (Potsdam: 2016 tiles of 6x512x512), tile list
sharded over the ranks of a torchrun job (no data-path collective).  Two numbers per run:
  resident : tiles already in HBM (the 12.7 GB of soft labels of the full set fit one B200), refine -> select -> uint8
  staged   : PseudoLabelRegenerator.run on pinned HOST batches (H2D of batch i+1 overlaps batch i, 1 byte/pixel back)
Synthetic tiles: a few distinct batches are cycled over the tile list (host memory), which does not change the traffic.

    python tools/regen_bench.py [--tiles 2016] [--batch 12]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/regen_bench.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from uemda_b200 import _lib, config, mining  # noqa: E402
from uemda_b200.gast.alignment import Aligner  # noqa: E402
from uemda_b200.regen import PseudoLabelRegenerator  # noqa: E402
from uemda_b200.synth import WORKLOADS, Workload, make_inputs  # noqa: E402


class _Log:
    def info(self, *a, **k):
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=2016)
    ap.add_argument("--batch", type=int, default=12)
    ap.add_argument("--distinct", type=int, default=3)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    config.strict_asserts = False
    base = WORKLOADS["cfg4_potsdam_tiles"]
    wl = Workload(base.name, args.batch, base.c, base.H, base.W, base.k, base.scale, base.regions)
    host = []
    for i in range(args.distinct):
        inp = make_inputs(wl, seed=2333 + i)
        host.append({"soft": inp["soft"].pin_memory(), "sup": inp["sup"].pin_memory(), "feat": inp["feat"].pin_memory(),
                     "preds": [inp["pred1"].pin_memory(), inp["pred2"].pin_memory()], "names": None})
    R = int(inp["ignore_id"]) + 1
    al = Aligner(_Log(), feat_channels=wl.k, class_num=wl.c, decay=0.996)
    al.prototypes = inp["prototypes"].to(dev)
    regen = PseudoLabelRegenerator(al, 0.8, 0.6, "all", 2.0, -1, True, num_regions=R)
    nb = (args.tiles + args.batch - 1) // args.batch
    batches = [host[i % args.distinct] for i in range(nb)]
    lo, hi = mining.shard_range(nb, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident: this rank's batches in HBM (cycled distinct batches), timed with CUDA events
    res = [{k: ([t.to(dev) for t in v] if isinstance(v, list) else (v.to(dev) if v is not None else None)) for k, v in h.items()}
           for h in host]
    for r in res:
        regen.process(r["soft"], r["sup"], r["feat"], r["preds"])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(lo, hi):
        r = res[i % args.distinct]
        regen.process(r["soft"], r["sup"], r["feat"], r["preds"])
    e1.record()
    barrier()
    t_res = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    # ---- staged: the public driver on pinned host batches
    regen.run(batches[:2], lambda names, arr: None, 0, 1, dev)
    barrier()
    e0.record()
    n_done, hist = regen.run_sharded(batches, lambda names, arr: None, device=dev, class_num=wl.c)
    e1.record()
    barrier()
    t_st = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_st, op=dist.ReduceOp.MAX)
    if rank == 0:
        px = nb * args.batch * wl.H * wl.W
        in_bytes = sum(t.numel() * t.element_size() for t in [host[0]["soft"], host[0]["sup"], host[0]["feat"]] + host[0]["preds"])
        print(json.dumps({"workload": "cfg4_potsdam_tiles", "tiles": nb * args.batch, "batch": args.batch, "n_gpus": world,
                          "resident_Mpixel_s": px / float(t_res.item()) / 1e3, "resident_ms": float(t_res.item()),
                          "staged_Mpixel_s": px / float(t_st.item()) / 1e3, "staged_ms": float(t_st.item()),
                          "h2d_bytes_per_tile": in_bytes // args.batch, "d2h_bytes_per_tile": wl.H * wl.W,
                          "tiles_rank0": n_done * 1, "label_hist_global": [int(v) for v in hist.cpu()],
                          "label_hist_total_ok": int(hist.sum()) == px,
                          "efficiency_note": "weak-scaling denominator: the same command at 1 GPU"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
