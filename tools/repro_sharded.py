"""Development aid: the ShardedMiner path (externally supplied ignored id) on one GPU, small workload."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from uemda_b200 import _lib, config, mining, ops
from uemda_b200.gast.alignment import Aligner, DownscaleLabel
from uemda_b200.synth import Workload, make_inputs

dev_index = int(os.environ.get("DEV", "0"))
torch.cuda.set_device(dev_index)
dev = torch.device("cuda", dev_index)
wl = Workload("small", 2, 6, 128, 128, 128, 16, 64)
inp = make_inputs(wl, seed=1)
d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in inp.items()}


class Log:
    def info(self, *a, **k):
        pass


al = Aligner(Log(), feat_channels=wl.k, class_num=wl.c, decay=0.996)
al.downscale_gt = DownscaleLabel(wl.scale, wl.c, -1, 0.75)
al.prototypes = d["prototypes"].clone()
config.strict_asserts = False
miner = mining.ShardedMiner(al)
R = int(inp["ignore_id"]) + 1
for it in range(3):
    ign = miner.global_ignored_id(d["sup"])
    out = mining.refine_select(7, d["soft"], 2.0, feat=d["feat"], prototypes=al.prototypes, pred1=d["pred1"], pred2=d["pred2"],
                               sup=d["sup"], num_regions=R, ignored_id=ign, select=(0.8, 0.6, -1), uvem=(0.2, 0.7, 4.0))
    miner.update_prototype(d["feat_s"], d["label_s"])
    torch.cuda.synchronize()
print("ok", out[1].float().mean().item())

# same step inside a CUDA graph (two streams like bench.py)
side = torch.cuda.Stream(device=dev)
ws = torch.zeros(_lib.load().uem_mine_ws_bytes(wl.b, wl.c, wl.H, wl.W, wl.h, wl.w, wl.k, R), dtype=torch.uint8, device=dev)


def step():
    cur = torch.cuda.current_stream(dev)
    protos = al.prototypes
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        miner.update_prototype(d["feat_s"], d["label_s"])
    ign = miner.global_ignored_id(d["sup"])
    out = mining.refine_select(7, d["soft"], 2.0, feat=d["feat"], prototypes=protos, pred1=d["pred1"], pred2=d["pred2"],
                               sup=d["sup"], num_regions=R, ignored_id=ign, select=(0.8, 0.6, -1), uvem=(0.2, 0.7, 4.0), ws=ws)
    cur.wait_stream(side)
    return out


for _ in range(3):
    step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    keep = step()
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
print("graph ok", keep[1].float().mean().item())
