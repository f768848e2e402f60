"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py

Every array named ``out_*`` in the fixtures is the output of a reference function
(uemda/gast/alignment.py, pseudo_generation.py, balance.py) imported through
``oracle/ref_shim.py``; arrays named ``in_*`` are the seeded inputs it was given.  The only
non-reference arithmetic involved is the ``torch_scatter.scatter`` shim (see ref_shim.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_shim import NullLogger, load_reference  # noqa: E402
from uemda_b200.synth import Workload, make_inputs  # noqa: E402

CASES = {
    # name: (workload, preds as list?, decay)
    "isprs_small": (Workload("isprs_small", 2, 6, 32, 48, 32, 8, 12), True, 0.996),
    "loveda_small": (Workload("loveda_small", 1, 7, 64, 64, 48, 16, 9), False, 0.999),
}


def np_(t):
    return t.detach().cpu().numpy()


def run_case(ref, wl, two_heads, decay, seed):
    torch.manual_seed(seed)
    inp = make_inputs(wl, seed=seed)
    c = wl.c
    out = {}
    for key in ("soft", "pred1", "pred2", "feat", "sup", "prototypes", "label_s", "feat_s", "logits"):
        out["in_" + key] = np_(inp[key])
    out["meta"] = np.array([wl.b, wl.c, wl.H, wl.W, wl.k, wl.scale, int(two_heads), inp["ignore_id"]],
                           dtype=np.int64)
    out["decay"] = np.array([decay], dtype=np.float64)

    al = ref.Aligner(NullLogger(), feat_channels=wl.k, class_num=c, ignore_label=-1, decay=decay)
    al.downscale_gt = ref.DownscaleLabel(scale_factor=wl.scale, n_classes=c, ignore_label=-1, min_ratio=0.75)
    al.prototypes = inp["prototypes"].clone()
    preds = [inp["pred1"], inp["pred2"]] if two_heads else inp["pred1"]

    # a9
    flat = inp["feat"].permute(0, 2, 3, 1).reshape(-1, wl.k)
    out["out_pearson"] = np_(al._pearson_dist(flat, al.prototypes))
    # a6, all modes
    for mode in ("all", "s", "p", "l"):
        out["out_refine_" + mode] = np_(al.label_refine(inp["sup"], inp["feat"], preds, inp["soft"],
                                                       refine=True, mode=mode, temp=2.0))
    refined = torch.from_numpy(out["out_refine_all"])
    # a5
    out["out_select_refined"] = np_(ref.pseudo_selection(refined, 0.8, 0.6, "tensor", -1))
    out["out_select_soft"] = np_(ref.pseudo_selection(inp["soft"].clone(), 0.8, 0.6, "tensor", -1))
    out["out_select_soft_lowcut"] = np_(ref.pseudo_selection(inp["soft"].clone(), 0.5, 0.2, "tensor", -1))
    out["out_select1_soft"] = np_(ref.pseudo_selection1(inp["soft"].clone(), 0.8, 0.6, "tensor", -1))
    hard = torch.from_numpy(out["out_select_soft"])
    # a7
    out["out_expand"] = np_(al.superpixel_expand(hard, inp["sup"]))
    # a8
    out["out_downscale_src"] = np_(al.downscale_gt(inp["label_s"]))
    out["out_downscale_hard"] = np_(al.downscale_gt(hard))
    # a13
    out["out_proto_weight_4pixel"] = np_(al.get_prototype_weight_4pixel(inp["feat"], hard))
    # a10: update_prototype twice (EMA state carried)
    down = al.update_prototype(inp["feat_s"], inp["label_s"])
    out["out_update_down"] = np_(down)
    out["out_proto_after1"] = np_(al.prototypes)
    out["out_local_proto"] = np_(al._compute_local_prototypes(inp["feat_s"], down, update=False))
    al.update_prototype(inp["feat"], inp["label_s"])
    out["out_proto_after2"] = np_(al.prototypes)
    # a12
    al.update_prototype_bytarget(inp["feat"], inp["soft"])
    out["out_proto_bytarget"] = np_(al.prototypes)
    # a11
    al.update_avg(inp["feat_s"], inp["label_s"])
    al.update_avg(inp["feat"], inp["label_s"])
    out["out_avg_sum"] = np_(al._data_sum)
    out["out_avg_cnt"] = np_(al._data_cnt)
    al.init_avg()
    out["out_avg_proto"] = np_(al.prototypes)
    # a1-a4 from logits (train_align_uem.py:158-160)
    import torch.nn.functional as tnf
    x1 = tnf.interpolate(inp["pred1"], (wl.H, wl.W), mode="bilinear", align_corners=True)
    x2 = tnf.interpolate(inp["pred2"], (wl.H, wl.W), mode="bilinear", align_corners=True)
    soft2 = ((x1.softmax(dim=1) + x2.softmax(dim=1)) * 0.5)
    out["out_soft_from_logits"] = np_(soft2)
    # a2/a3 + a14 through the reference loss objects
    cb = ref.ClassBalance(class_num=c, ignore_label=-1, decay=0.99, temperature=0.5)
    loss_fn = ref.UVEMLoss(m=0.2, threshold=0.7, gamma=4.0, class_balancer=cb, class_num=c, ignore_label=-1)
    lts = refined.permute(0, 2, 3, 1).reshape(-1, c)
    unc = torch.sum(-lts * torch.log(lts), dim=1)
    out["out_entropy"] = np_(unc)
    out["out_uvem_weight"] = np_(loss_fn.get_weight(unc))
    grid = torch.linspace(0, 0.99, 100)
    out["in_uvem_grid"] = np_(grid)
    out["out_uvem_weight_grid"] = np_(loss_fn.get_weight(grid))
    hard_r = torch.from_numpy(out["out_select_refined"])
    out["out_cb_weight1"] = np_(cb.get_class_weight_4pixel(hard_r.reshape(-1)))
    out["out_cb_freq1"] = np_(cb.freq)
    out["out_cb_weight2"] = np_(cb.get_class_weight_4pixel(hard.reshape(-1)))
    out["out_cb_freq2"] = np_(cb.freq)
    logits_full = inp["logits"].clone().requires_grad_(True)
    loss = loss_fn(logits_full, hard_r, refined)
    loss.backward()
    out["out_uvem_loss"] = np_(loss).reshape(1)
    out["out_uvem_loss_grad_sum"] = np_(logits_full.grad.abs().sum()).reshape(1)
    out["out_cb_freq3"] = np_(cb.freq)
    ups = ref.UPSLoss(threshold=0.7, class_balancer=None, class_num=c, ignore_label=-1)
    out["out_ups_loss"] = np_(ups(inp["logits"], hard_r, refined)).reshape(1)
    # scatter seam itself (restated on both sides -> "unpinned", kept for regression only)
    rows = inp["soft"].permute(0, 2, 3, 1).reshape(wl.b, -1, c)
    out["out_scatter_max"] = np_(ref.scatter(rows, inp["sup"].reshape(wl.b, -1, 1), dim=1, reduce="max"))
    return out


def main():
    ref = load_reference()
    torch.set_num_threads(1)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for i, (name, (wl, two, decay)) in enumerate(CASES.items()):
        res = run_case(ref, wl, two, decay, seed=2333 + 17 * i)
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **res)
        print(name, "->", path, "%.1f KB" % (os.path.getsize(path) / 1024))
    # edge cases: exact zeros in probabilities (NaN entropy), ties, uniform maps
    c = 6
    p = torch.zeros(1, c, 4, 4)
    p[:, 0] = 1.0
    p[0, :, 1, 1] = torch.tensor([0.5, 0.5, 0, 0, 0, 0])
    p[0, :, 2, 2] = 1.0 / c
    p[0, :, 3, 3] = torch.tensor([0.7, 0.1, 0.1, 0.05, 0.05, 0.0])
    loss_fn = ref.UVEMLoss(m=0.2, threshold=0.7, gamma=4.0, class_num=c)
    lts = p.permute(0, 2, 3, 1).reshape(-1, c)
    unc = torch.sum(-lts * torch.log(lts), dim=1)
    edge = dict(in_prob=np_(p), out_entropy=np_(unc), out_weight=np_(loss_fn.get_weight(unc)),
                out_select=np_(ref.pseudo_selection(p.clone(), 0.8, 0.6, "tensor", -1)),
                out_select1=np_(ref.pseudo_selection1(p.clone(), 0.8, 0.6, "tensor", -1)))
    for m_, t_, g_ in ((0.0, 0.7, 4.0), (0.8, 0.7, 4.0), (0.1, 0.7, 8.0)):
        fn = ref.UVEMLoss(m=m_, threshold=t_, gamma=g_, class_num=c)
        u = torch.linspace(-0.1, 1.2, 66)
        edge["in_u_%g_%g_%g" % (m_, t_, g_)] = np_(u)
        edge["out_w_%g_%g_%g" % (m_, t_, g_)] = np_(fn.get_weight(u))
    path = os.path.join(ROOT, "tests", "golden", "edge_cases.npz")
    np.savez_compressed(path, **edge)
    print("edge_cases ->", path)


if __name__ == "__main__":
    main()
