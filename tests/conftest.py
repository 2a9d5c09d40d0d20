import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN_DIR, "manifest.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
                cache[name] = {k: z[k] for k in z.files}
        return cache[name]

    return load


def assert_close_f32(got, want, rtol=1e-5, atol=1e-6, what=""):
    """Parity bar for float outputs: identical NaN mask, |got-want| <= atol + rtol*|want|."""
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    gn, wn = np.isnan(got), np.isnan(want)
    assert np.array_equal(gn, wn), f"{what}: NaN masks differ at {int((gn != wn).sum())} px"
    ok = ~wn
    err = np.abs(got[ok].astype(np.float64) - want[ok].astype(np.float64))
    lim = atol + rtol * np.abs(want[ok].astype(np.float64))
    bad = err > lim
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.size} px outside rtol={rtol} atol={atol}; "
                           f"max abs err {err.max():.3e}")
