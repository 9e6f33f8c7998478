import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def split(flat, off):
    return [flat[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def rel_err(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


def assert_close(a, b, rtol=1e-9, atol=0.0, what=""):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape, "%s shape %s vs %s" % (what, a.shape, b.shape)
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    bad = ~(err <= tol)
    assert not bad.any(), "%s: %d/%d outside rtol=%g atol=%g, worst rel %.3e" % (
        what, bad.sum(), bad.size, rtol, atol, float(np.max(err / np.maximum(np.abs(b), 1e-300))))


@pytest.fixture(scope="session")
def gold():
    return golden
