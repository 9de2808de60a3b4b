import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    out = {}
    for k in d.files:
        a = d[k]
        out[k] = torch.from_numpy(np.array(a)) if a.dtype.kind in "fiub" else a
    return out


@pytest.fixture(scope="session")
def cuda_device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def assert_close_obs(actual, expected, rtol=1e-5, atol=1e-6, angle_cols=(), what=""):
    """SURVEY.md section 4 contract for float observations: |d| <= rtol*|ref| + atol, angle columns modulo 2*pi."""
    a, e = actual.double().cpu(), expected.double().cpu()
    d = (a - e).abs()
    if angle_cols:
        two_pi = 2 * np.pi
        for c in angle_cols:
            dc = d[..., c]
            d[..., c] = torch.minimum(dc, (dc - two_pi).abs())
    bad = d > (rtol * e.abs() + atol)
    if bad.any():
        idx = bad.nonzero()[0].tolist()
        raise AssertionError("%s: %d/%d elements out of tolerance, first at %s: got %r want %r" % (
            what, int(bad.sum()), bad.numel(), idx, a[tuple(idx)].item(), e[tuple(idx)].item()))
