"""Generate tests/golden/uvem_loss_small.npz by EXECUTING THE UNMODIFIED REFERENCE (TEST INFRASTRUCTURE).

    python oracle/gen_golden_uvemloss.py        (build container only: needs /root/reference)

``uemda.gast.balance`` (loss_calc_uvem :437-457, UVEMLoss :345-423, UPSLoss :306-342, ClassBalance :15-78) imports
without any shim except ``Tensor.cuda`` -> identity (ClassBalance.__init__ calls .cuda(), balance.py:25); forward and
autograd backward run on CPU torch.  in_* = seeded inputs, out_* = the reference's losses and gradients with respect
to the two low-resolution heads.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("UEM_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, REF)
torch.Tensor.cuda = lambda self, *a, **k: self  # the container has no GPU

from uemda.gast.balance import ClassBalance, UPSLoss, UVEMLoss, loss_calc_uvem  # noqa: E402  (the reference module itself)


def main():
    g = torch.Generator().manual_seed(2333)
    b, c, h, w, H, W = 2, 6, 5, 7, 40, 56
    x1 = torch.randn(b, c, h, w, generator=g) * 2
    x2 = torch.randn(b, c, h, w, generator=g) * 2
    soft = torch.softmax(torch.randn(b, c, H, W, generator=g) * torch.linspace(0.5, 6, W).view(1, 1, 1, W), dim=1)
    label = torch.randint(-1, c, (b, H, W), generator=g)
    out = {"in_x1": x1.numpy(), "in_x2": x2.numpy(), "in_soft": soft.numpy(), "in_label": label.numpy()}

    def run(tag, loss_fn, heads, multi=True):
        xs = [x.clone().requires_grad_(True) for x in heads]
        loss = loss_calc_uvem(xs if multi else xs[0], label, soft, loss_fn, multi=multi)
        loss.backward()
        out["out_%s_loss" % tag] = loss.detach().numpy()
        for i, x in enumerate(xs):
            out["out_%s_grad%d" % (tag, i + 1)] = x.grad.numpy()

    run("uvem", UVEMLoss(m=0.2, threshold=0.7, gamma=4.0, class_num=c), [x1, x2])
    run("uvem_single", UVEMLoss(m=0.2, threshold=0.7, gamma=4.0, class_num=c), [x1], multi=False)
    run("ups", UPSLoss(threshold=0.7, class_num=c), [x1, x2])
    cb = ClassBalance(class_num=c, ignore_label=-1, decay=0.99, temperature=0.5)
    run("uvem_cb", UVEMLoss(m=0.2, threshold=0.7, gamma=4.0, class_balancer=cb, class_num=c), [x1, x2])
    out["out_uvem_cb_freq"] = cb.freq.numpy()
    path = os.path.join(ROOT, "tests", "golden", "uvem_loss_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k2: v.shape for k2, v in out.items()})


if __name__ == "__main__":
    main()
