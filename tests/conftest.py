import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """Thin accessor over a tests/golden/*.npz fixture (outputs of the reference itself)."""

    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        m = self.z["meta"] if "meta" in self.z else None
        if m is not None:
            self.b, self.c, self.H, self.W, self.k, self.scale, two, self.ignore_id = [int(v) for v in m]
            self.two_heads = bool(two)
            self.decay = float(self.z["decay"][0])

    def t(self, key, device="cpu"):
        return torch.from_numpy(self.z[key]).to(device)

    def preds(self, device="cpu"):
        if self.two_heads:
            return [self.t("in_pred1", device), self.t("in_pred2", device)]
        return self.t("in_pred1", device)


@pytest.fixture(params=["isprs_small", "loveda_small"])
def golden(request):
    return Golden(request.param)


@pytest.fixture
def edge_cases():
    return Golden("edge_cases")


def assert_close(a, b, rtol=1e-5, atol=1e-7, what=""):
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    nan_a, nan_b = torch.isnan(a), torch.isnan(b)
    assert torch.equal(nan_a, nan_b), what + ": NaN pattern differs"
    a = torch.where(nan_a, torch.zeros_like(a), a)
    b = torch.where(nan_b, torch.zeros_like(b), b)
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bad.any(), "%s: %d/%d out of tol, max abs %.3e, max rel %.3e" % (
        what, int(bad.sum()), a.numel(), float(err.max()), float((err / (b.abs() + 1e-30)).max()))
