#!/usr/bin/env python
"""Benchmark of the pseudo-label mining step (BASELINE.json metric: Mpixel/s + fraction of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A "step" = one pass of the hot path over one target batch + one source batch
(tools/train_ssl_uem.py:209-216 + the detached part of loss_calc_uvem, balance.py:372-396):
    label_refine(mode='all', temp=2) -> pseudo_selection(0.8, 0.6) -> update_prototype(feat_s, label_s)
    -> entropy + UVEM weight of the refined soft labels.
Default workload = BASELINE.json configs[1]: ISPRS 8x6x512x512, 2048-ch features at 1/16 res.

One JSON line is printed by rank 0 (contract in the task statement):
  value     whole-job Mpixel/s, inputs resident in HBM, K steps replayed as CUDA graphs, CUDA-event timed,
            max over ranks; inputs rotate over several buffer sets so that every step reads cold data
  e2e       same metric through the public drop-in API with HOST (pinned) inputs: H2D of every input and D2H
            of the hard labels inside the timed region
  roofline  fused refine kernel: algorithmic bytes / CUDA-event duration vs the measured HBM copy bandwidth
  cpu_baseline  the CPU oracle (a torch-CPU restatement of the reference, pinned to it by tests/golden)
            on this box's host cores, bounded sample
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

os.environ.pop("NCCL_DEBUG", None)  # NCCL prints its version banner to stdout at WARN/VERSION level: keep stdout to the JSON line

import torch  # noqa: E402

UVEM = (0.2, 0.7, 4.0)      # --uvem-m/-t/-g, tools/train_ssl_uem.py:59-61
CUTOFF = (0.8, 0.6)         # CUTOFF_TOP/LOW, configs/st/uemda/2potsdam.py:24-25
DECAY = 0.996               # tools/train_ssl_uem.py:117
TEMP = 2.0                  # --refine-temp
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="cfg2_isprs_8x6x512")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sets", type=int, default=3, help="rotating input buffer sets (working set > L2)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample-images", type=int, default=4)
    return ap.parse_args()


# --------------------------------------------------------------------------------------------- helpers
def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = []
        with open(self.tmp.name) as f:
            for line in f:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) >= 7:
                    rows.append(parts)
        os.unlink(self.tmp.name)
        sm = []
        reasons = set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out["sm_mhz"] = statistics.median(sm)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_step_fn(wl, n_images):
    from oracle import uem_oracle as O
    from uemda_b200.synth import make_inputs
    inp = make_inputs(wl, seed=2333, b=n_images)
    protos = inp["prototypes"]

    def step():
        return O.mining_step(inp, protos, wl.c, mode="all", temp=TEMP, cutoff_top=CUTOFF[0], cutoff_low=CUTOFF[1],
                             decay=DECAY, uvem=UVEM, scale_factor=wl.scale)
    return step, n_images * wl.H * wl.W


def run_cpu_baseline(wl, n_images, warmup=1, steps=3):
    torch.set_num_threads(os.cpu_count() or 1)
    step, px = cpu_step_fn(wl, n_images)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    t = statistics.median(ts)
    return {"value": px / t / 1e6, "unit": "Mpixel/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d of %d images of %s, %d warm-up + median of %d steps (%.3f s/step), torch CPU oracle "
                      "(oracle/uem_oracle.py)" % (n_images, wl.b, wl.name, warmup, steps, t)}


def run_reference_arm(args, wl):
    """--impl reference: the reference's CPU implementation of the path, i.e. the oracle port (the reference is
    pure Python/torch and cannot travel to the GPU box; the port is pinned to it by tests/golden)."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    n_img = min(wl.b, args.cpu_sample_images)
    step, px = cpu_step_fn(wl, n_img)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    val = px / dt / 1e6
    line = {
        "impl": "reference", "metric": "pseudo-label mining throughput", "value": val, "unit": "Mpixel/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, "b_per_gpu": wl.b, "c": wl.c, "H": wl.H, "W": wl.W, "k": wl.k, "feat_scale": wl.scale,
                   "regions": wl.regions, "step": "label_refine(all)+pseudo_selection+update_prototype+entropy/uvem_weight",
                   "l2_policy": "n/a (host cores)", "cuda_graph": False,
                   "parallelism": "rank 0 only, %d host threads" % torch.get_num_threads()},
        "cpu_baseline": {"value": val, "unit": "Mpixel/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "%d of %d images per step" % (n_img, wl.b)},
        "e2e": {"value": val, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- GPU arm
class _Log:
    def info(self, *a, **k):
        pass


def make_sets(wl, nsets, dev, rank):
    """nsets distinct device copies of one seeded batch (batch-rolled so contents differ); pinned host copy of set 0."""
    from uemda_b200.synth import make_inputs
    inp = make_inputs(wl, seed=2333 + rank)
    keys = ("soft", "sup", "feat", "pred1", "pred2", "label_s", "feat_s")
    host = {k: inp[k].pin_memory() for k in keys}
    sets = []
    for i in range(nsets):
        sets.append({k: torch.roll(host[k], shifts=i, dims=0).to(dev) for k in keys})
    return inp, host, sets


def main():
    args = parse()
    from uemda_b200.synth import WORKLOADS
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py (impl ours) needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from uemda_b200 import _lib, config, mining, ops
    from uemda_b200.gast.alignment import Aligner, DownscaleLabel
    from uemda_b200.gast.pseudo_generation import pseudo_selection
    lib = _lib.load()

    inp, host, sets = make_sets(wl, args.sets, dev, rank)
    R = int(inp["ignore_id"]) + 1
    al = Aligner(_Log(), feat_channels=wl.k, class_num=wl.c, ignore_label=-1, decay=DECAY)
    al.downscale_gt = DownscaleLabel(wl.scale, wl.c, -1, 0.75)
    al.prototypes = inp["prototypes"].to(dev)
    al.num_regions = R
    miner = mining.ShardedMiner(al) if world > 1 else None
    ws = None

    # the target chain is the critical path: it runs on a high-priority stream, the source chain (prototype sums) on a
    # low-priority one so that its CTAs fill the gaps instead of competing for the first wave
    main_stream = torch.cuda.Stream(device=dev, priority=-1)
    side = torch.cuda.Stream(device=dev, priority=0)
    torch.cuda.set_stream(main_stream)
    proto_state = al.prototypes.clone()   # the replicated prototype bank: read by the refine chain, EMA-updated in place
    al.prototypes = proto_state
    n_pack = wl.c * wl.k + wl.c + 1
    # one exchange buffer pair per input set: phase A + all_gather of step i+1 run one step ahead of phase B of step i
    packed_bufs = [torch.zeros(n_pack, dtype=torch.float64, device=dev) for _ in range(args.sets)]
    gathered_bufs = [torch.zeros((world, n_pack), dtype=torch.float64, device=dev) for _ in range(args.sets)]
    ahead = torch.cuda.Stream(device=dev)
    comm = torch.cuda.Stream(device=dev)
    ev_p = [torch.cuda.Event() for _ in range(args.sets)]
    ev_a = [torch.cuda.Event() for _ in range(args.sets)]
    ev_b = [torch.cuda.Event() for _ in range(args.sets)]

    def source_stats(s, fold=True):
        """source-side chain on the second stream: DownscaleLabel -> masked prototype sums (independent of the target chain)"""
        cur = torch.cuda.current_stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            down = al.downscale_gt(s["label_s"])
            return ops.proto_accumulate(s["feat_s"], down, wl.c, -1, fold=fold)

    def target_chain(s, ignored, ws_=None, regions_ready=False):
        return mining.refine_select(7, s["soft"], TEMP, feat=s["feat"], prototypes=proto_state, pred1=s["pred1"],
                                    pred2=s["pred2"], sup=s["sup"], num_regions=R, ignored_id=ignored, eps=al.eps,
                                    select=(CUTOFF[0], CUTOFF[1], -1), ws=ws if ws_ is None else ws_, uvem=UVEM,
                                    regions_ready=regions_ready)

    def step_resident(s):
        """device-resident single-GPU step, no host sync (strict asserts off): source chain and target chain on two
        streams; the EMA writes the prototype bank in place once the target chain (its last reader) is enqueued."""
        nonlocal ws
        cur = torch.cuda.current_stream(dev)
        partials = source_stats(s, fold=False)
        out = target_chain(s, None)
        cur.wait_stream(side)
        ops.proto_fold_finalize(partials, proto_state, eps=al.eps, decay=DECAY, out=proto_state)
        return out

    # multi-GPU: the batch is sharded by image; the only exchange is ONE all_gather of [prototype sums | counts | max id]
    # per step, kept outside the captured graphs (phase A: local statistics; phase B: fold + target chain + EMA)
    # The region half of the target chain needs nothing global, so it runs in phase A, one step ahead: the rank-local
    # max superpixel id falls out of the same pass as on one GPU (no second pass over the ids); one workspace per set.
    ws_sets = [None] * args.sets

    folded = [(torch.zeros((wl.c, wl.k), dtype=torch.float32, device=dev), torch.zeros(wl.c, dtype=torch.int64, device=dev),
               torch.zeros(1, dtype=torch.int64, device=dev)) for _ in range(args.sets)]

    def phase_a(s, j):
        cur = torch.cuda.current_stream(dev)
        partials = source_stats(s, fold=False)
        mx = mining.region_phase(s["soft"], s["sup"], TEMP, R, ws_sets[j], wl.h, wl.w, wl.k)
        cur.wait_stream(side)
        ops.pack_local_partials(partials, mx, out=packed_bufs[j])   # image-order fold + pack in one launch

    # single GPU: the same two phases without the exchange.  Phase A (source statistics + region half of the target chain)
    # reads neither the prototype bank nor anything phase B writes, so phase A of step i+1 overlaps phase B of step i:
    # a software pipeline ACROSS steps (results identical to running the steps back to back).
    partials_sets = [None] * args.sets
    local_ids = [None] * args.sets

    def phase_a1(s, j):
        cur = torch.cuda.current_stream(dev)
        partials_sets[j] = source_stats(s, fold=False)
        local_ids[j] = mining.region_phase(s["soft"], s["sup"], TEMP, R, ws_sets[j], wl.h, wl.w, wl.k)
        cur.wait_stream(side)

    def phase_b1(s, j):
        out = target_chain(s, local_ids[j], ws_sets[j], regions_ready=True)   # one rank: the local max id is the global one
        ops.proto_fold_finalize(partials_sets[j], proto_state, eps=al.eps, decay=DECAY, out=proto_state)
        return out

    def phase_b(s, j):
        # rank-ordered fold of the gathered statistics, one launch, captured with the rest of phase B (an eager launch
        # behind the all_gather would take it off the GPU critical path but costs more host time per step than it saves)
        sums, counts, ignored = ops.fold_gathered(gathered_bufs[j], wl.c, wl.k, out=folded[j])
        out = target_chain(s, ignored, ws_sets[j], regions_ready=True)
        ops.proto_finalize(sums, counts, proto_state, eps=al.eps, decay=DECAY, want_local=False, out=proto_state)
        return out

    def step_sharded(s, j=0):
        phase_a(s, j)
        miner.exchange(packed_bufs[j], out=gathered_bufs[j])
        return phase_b(s, j)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    config.strict_asserts = False
    need = lib.uem_mine_ws_bytes(wl.b, wl.c, wl.H, wl.W, wl.h, wl.w, wl.k, R)
    ws = torch.zeros(need, dtype=torch.uint8, device=dev)
    ws_sets = [torch.zeros(need, dtype=torch.uint8, device=dev) for _ in range(args.sets)]
    step_eager = step_sharded if miner else step_resident

    # ---- warm-up (eager), then graph capture: one graph (two around the exchange when sharded) per buffer set
    for i in range(max(args.warmup, 3)):
        step_eager(sets[i % args.sets])
    barrier()
    graphs = None
    if not args.no_graph:
        try:
            graphs = []
            keep = []
            for j, s in enumerate(sets):
                if miner:
                    ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                    with torch.cuda.graph(ga, stream=main_stream):
                        phase_a(s, j)
                    miner.exchange(packed_bufs[j], out=gathered_bufs[j])
                    with torch.cuda.graph(gb, stream=main_stream):
                        keep.append(phase_b(s, j))
                    graphs.append((ga, gb))
                else:
                    ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                    with torch.cuda.graph(ga, stream=main_stream):
                        phase_a1(s, j)
                    with torch.cuda.graph(gb, stream=main_stream):
                        keep.append(phase_b1(s, j))
                    graphs.append((ga, gb))
            barrier()
        except Exception as e:  # noqa: BLE001
            if rank == 0:
                print("graph capture failed, timing eagerly: %r" % (e,), file=sys.stderr)
            graphs = None
            torch.cuda.synchronize()

    def issue_ahead(i):
        """phase A of step i on the look-ahead stream, its all_gather on a third stream: three pipeline stages
        (A(i+2) | all_gather(i+1) | B(i)) so the collective's latency is not serialised behind the next phase A"""
        j = i % args.sets
        with torch.cuda.stream(ahead):
            ahead.wait_event(ev_b[j])       # phase B that last read this buffer set has finished
            graphs[j][0].replay()
            ev_p[j].record(ahead)
            if not miner:
                ev_a[j].record(ahead)
                return
        with torch.cuda.stream(comm):
            comm.wait_event(ev_p[j])
            miner.exchange(packed_bufs[j], out=gathered_bufs[j])
            ev_a[j].record(comm)

    def run_steps(first, n):
        cur = torch.cuda.current_stream(dev)
        if graphs is None:
            for i in range(first, first + n):
                step_eager(sets[i % args.sets])
        else:
            for e in ev_b:
                e.record(cur)
            # look-ahead depth: with 3 buffer sets phase A + the all_gather run TWO steps ahead, so a late rank has a whole
            # extra step before its contribution is needed (the buffers of step i+2 were last read by phase B of step i-1)
            depth = max(1, min(2, args.sets - 1))
            for i in range(first, min(first + depth, first + n)):
                issue_ahead(i)
            for i in range(first, first + n):
                if i + depth < first + n:
                    issue_ahead(i + depth)
                j = i % args.sets
                cur.wait_event(ev_a[j])
                graphs[j][1].replay()
                ev_b[j].record(cur)


    run_steps(0, args.warmup)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    launches0 = lib.uem_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_steps(args.warmup, args.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches_eager_per_step = None
    if graphs is None:
        launches_eager_per_step = (lib.uem_kernel_launches() - launches0) / args.steps
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())

    # ---- kernel-level timing (eager, event pair recorded inside the C call around the fused refine kernel)
    kern_ms = []
    n_k = min(args.steps, 50)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_k)]
    for a, b2 in evs:  # torch creates the cudaEvent_t lazily on the first record()
        a.record()
        b2.record()
    torch.cuda.synchronize()
    l0 = lib.uem_kernel_launches()
    for i in range(n_k):
        lib.uem_profile_refine_events(evs[i][0].cuda_event, evs[i][1].cuda_event)
        step_eager(sets[i % args.sets])
    torch.cuda.synchronize()
    per_step_launches = (lib.uem_kernel_launches() - l0) / max(n_k, 1)
    for a, b2 in evs:
        try:
            kern_ms.append(a.elapsed_time(b2))
        except Exception:  # noqa: BLE001
            pass
    refine_ms = statistics.mean(kern_ms) if kern_ms else None

    # ---- e2e: public drop-in API, host (pinned) inputs, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        config.strict_asserts = True
        al.num_regions = None  # the un-hinted public path: reads sup.max() back like torch_scatter does
        hard_host = torch.empty((wl.b, wl.H, wl.W), dtype=torch.int64).pin_memory()
        h2d = sum(host[k].numel() * host[k].element_size() for k in host)
        d2h = hard_host.numel() * hard_host.element_size()

        # double-buffered input staging: the H2D copy of step i+1 (copy stream) overlaps the kernels of step i; every
        # step still pays its own full H2D of all inputs and D2H of the hard labels inside the timed region
        copy_stream = torch.cuda.Stream(device=dev)
        dbuf = [{k: torch.empty_like(host[k], device=dev) for k in host} for _ in range(2)]
        copied = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        def issue_copy(j):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[j % 2])
                for k in host:
                    dbuf[j % 2][k].copy_(host[k], non_blocking=True)
                copied[j % 2].record(copy_stream)

        def compute(j):
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(copied[j % 2])
            d = dbuf[j % 2]
            refined = al.label_refine(d["sup"], d["feat"], [d["pred1"], d["pred2"]], d["soft"], refine=True, mode="all", temp=TEMP)
            hard = pseudo_selection(refined, CUTOFF[0], CUTOFF[1], "tensor", -1)
            if miner:
                miner.update_prototype(d["feat_s"], d["label_s"])
            else:
                al.update_prototype(d["feat_s"], d["label_s"])
            ops.entropy_uvem_weight(refined, *UVEM)
            consumed[j % 2].record(cur)
            hard_host.copy_(hard, non_blocking=True)

        def run_e2e(n):
            for e in consumed:
                e.record(torch.cuda.current_stream(dev))
            issue_copy(0)
            for j in range(n):
                if j + 1 < n:
                    issue_copy(j + 1)
                compute(j)

        n_e = max(3, min(args.steps, 30))
        run_e2e(3)
        barrier()
        a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run_e2e(n_e)
        b2.record()
        barrier()
        e_ms = a.elapsed_time(b2) / n_e
        te = torch.tensor([e_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_ms = float(te.item())
        e2e = {"value": world * wl.pixels / e_ms / 1e3, "unit": "Mpixel/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e_ms, "steps": n_e,
               "api": "Aligner.label_refine + pseudo_selection + Aligner.update_prototype + entropy/UVEM weight",
               "staging": "pinned host inputs, double-buffered: H2D of step i+1 overlaps the kernels of step i"}
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        peak, peak_src = peak_hbm()
        alg_bytes = wl.pixels * (4 * wl.c + 8 + 4 * wl.c)  # refine kernel: read soft + sup, write refined
        roof = {"bound": "hbm", "kernel": "refine_col_kernel (fused label_refine, all views)", "achieved": None, "peak": peak,
                "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src, "algorithmic_bytes": alg_bytes,
                "kernel_ms": refine_ms}
        if refine_ms:
            roof["achieved"] = alg_bytes / (refine_ms * 1e-3) / 1e9
            roof["frac"] = roof["achieved"] / peak
        tpath = os.path.join(ROOT, "profiles", "refine_traffic.json")
        if os.path.exists(tpath):
            try:
                roof["traffic"] = json.load(open(tpath)).get(wl.name)
            except Exception:  # noqa: BLE001
                pass
        chain_bytes = wl.pixels * (4 * wl.c + 8 + 4 * wl.c + 8) + wl.b * wl.k * wl.h * wl.w * 4 \
            + wl.pixels * 8 + wl.b * wl.k * wl.h * wl.w * 4 + wl.pixels * (4 * wl.c + 8)
        line = {
            "metric": "pseudo-label mining throughput", "value": world * wl.pixels / ms / 1e3, "unit": "Mpixel/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "b_per_gpu": wl.b, "c": wl.c, "H": wl.H, "W": wl.W, "k": wl.k, "feat_scale": wl.scale,
                       "regions": wl.regions, "step": "label_refine(all)+pseudo_selection+update_prototype+entropy/uvem_weight",
                       "l2_policy": "inputs rotate over %d buffer sets (%.0f MiB each) so each step reads cold data" % (
                           args.sets, sum(v.numel() * v.element_size() for v in sets[0].values()) / 2 ** 20),
                       "cuda_graph": graphs is not None, "parallelism": "batch-sharded x%d" % world,
                       "pipeline": "two graphs per step: A = source statistics + region half of the target chain (reads no "
                                   "prototype), B = pearson + refine + selection + EMA; A of step i+1/i+2 overlaps B of step i"
                                   if graphs is not None else "none"},
            "step_algorithmic_bytes": chain_bytes,
            "step_hbm_frac": chain_bytes / (ms * 1e-3) / 1e9 / peak,
            "roofline": roof,
            "gpu_launches": int(round(per_step_launches * args.steps)),
            "kernels_per_step": per_step_launches,
            "clocks": clocks,
        }
        if e2e:
            line["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = run_cpu_baseline(wl, min(wl.b, args.cpu_sample_images))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
