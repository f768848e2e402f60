"""Generate tests/golden/iast_small.npz by EXECUTING THE REFERENCE's own ``ias_thresh`` / ``pre_slide`` (TEST INFRASTRUCTURE).

    python oracle/gen_golden_iast.py          (build container only: needs /root/reference)

``generate_pseudo`` (uemda/utils/tools.py:335-373) cannot run as a whole (it needs a model, a loader and an undefined
``palette``, :339), so its per-batch body (:349-371) is driven here with the reference's OWN ``ias_thresh`` (:323-333) doing
the percentile step; ``pre_slide`` (:61-97) is executed unmodified with a stand-in "model" that returns seeded tiles.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_shim import load_reference  # noqa: E402


def main():
    load_reference()
    import importlib
    tools = importlib.import_module("uemda.utils.tools")
    out = {}
    g = torch.Generator().manual_seed(2333)
    b, c, H, W = 3, 6, 48, 64
    alpha, beta, gamma = 0.2, 0.9, 8.0          # pl_alpha / pl_beta / pl_gamma of the IAST recipe
    cls_thresh = np.ones(c) * 0.9                # tools.py:346
    for step in range(3):
        z = torch.randn(b, c, H, W, generator=g) * (1.0 + step)
        z[:, 4] -= 6.0 if step == 1 else 0.0     # a class that (almost) never wins in one batch
        probs = torch.softmax(z, dim=1)
        out["in_probs_%d" % step] = probs.numpy()
        out["in_thresh_%d" % step] = cls_thresh.copy()
        max_items = probs.max(dim=1)
        label_pred = max_items[1].data.cpu().numpy()
        logits_pred = max_items[0].data.cpu().numpy()
        d = {k: [cls_thresh[k]] for k in range(c)}
        for k in range(c):
            d[k].extend(logits_pred[label_pred == k].astype(np.float16))
        tmp = tools.ias_thresh(d, c, alpha, w=cls_thresh, gamma=gamma)      # the reference's own function
        out["out_tmp_thresh_%d" % step] = tmp.copy()
        cls_thresh = beta * cls_thresh + (1 - beta) * tmp
        cls_thresh[cls_thresh >= 1] = 0.999
        out["out_thresh_%d" % step] = cls_thresh.copy()
        labs = []
        np_logits = probs.data.cpu().numpy()
        for i in range(b):
            logit = np_logits[i].transpose(1, 2, 0)
            label = np.argmax(logit, axis=2)
            amax = np.amax(logit, axis=2)
            thr = np.apply_along_axis(lambda x: [cls_thresh[_e] for _e in x], 1, label)
            ign = amax < thr
            label += 1
            label[ign] = 0
            labs.append(label.astype(np.uint8))
        out["out_labels_%d" % step] = np.stack(labs)
    out["iast_params"] = np.array([alpha, beta, gamma])

    # pre_slide, unmodified, on a 1 x 3 x 72 x 88 "image" with 32 x 32 tiles; the stand-in model returns seeded tiles
    tiles = []

    def model(x):
        t = torch.rand(x.shape[0], 3, x.shape[2], x.shape[3], generator=g)
        tiles.append(t.numpy().copy())
        return t

    image = torch.zeros(1, 3, 72, 88)
    full = tools.pre_slide(model, image, num_classes=3, tile_size=(32, 32), tta=False)
    out["slide_tiles"] = np.stack(tiles)
    out["slide_out"] = full.numpy()
    out["slide_meta"] = np.array([72, 88, 32, 32])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "iast_small.npz"), **out)
    print("wrote tests/golden/iast_small.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
