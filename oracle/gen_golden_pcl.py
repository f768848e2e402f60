"""Generate tests/golden/pcl_small.npz by EXECUTING THE UNMODIFIED REFERENCE (TEST INFRASTRUCTURE).

    python oracle/gen_golden_pcl.py        (build container only: needs /root/reference)

``uemda.loss.PrototypeContrastiveLoss`` (uemda/loss.py:10-47) imports without any shim; its forward and autograd
backward run on CPU torch.  in_* = seeded inputs, out_loss / out_grad = the reference's outputs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("UEM_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, REF)

from uemda.loss import PrototypeContrastiveLoss  # noqa: E402  (the reference module itself)


def main():
    g = torch.Generator().manual_seed(2333)
    b, k, h, w, c = 2, 64, 8, 12, 6
    feat = torch.randn(b, k, h, w, generator=g)
    protos = torch.randn(c, k, generator=g)
    feat = feat + 0.7 * protos[torch.randint(0, c, (b, h, w), generator=g)].permute(0, 3, 1, 2)
    labels = torch.randint(-1, c, (b, 1, h, w), generator=g)
    out = {"in_feat": feat.numpy(), "in_protos": protos.numpy(), "in_labels": labels.numpy()}
    for temp in (8.0, 0.5):
        f = feat.clone().requires_grad_(True)
        loss = PrototypeContrastiveLoss(temperature=temp, ignore_label=-1)(protos, f, labels)
        loss.backward()
        out["out_loss_t%g" % temp] = loss.detach().numpy()
        out["out_grad_t%g" % temp] = f.grad.numpy()
    path = os.path.join(ROOT, "tests", "golden", "pcl_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k2: v.shape for k2, v in out.items()})


if __name__ == "__main__":
    main()
