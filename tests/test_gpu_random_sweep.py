"""Randomised parity sweep: seeded random shapes (1..224 points, ragged or equal, 1D and 2D, with and without a mean
function / errors / nugget, uniform and scattered grids, small batches and batches large enough for the two-kernel
route) through the numpy-facing batch layer, against the pinned oracle object by object (relative 1e-9)."""
import numpy as np
import pytest

from conftest import assert_close
from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _case(seed):
    rng = np.random.default_rng(1000 + seed)
    dim = 1 if rng.random() < 0.7 else 2
    big = seed % 7 == 3                                  # >= 2048 objects: factor + grid kernels, uniform-grid variant
    b = int(rng.integers(2048, 2600)) if big else int(rng.integers(1, 25))
    nmax = 64 if big or rng.random() < 0.6 else 224
    if rng.random() < 0.5:
        sizes = np.full(b, int(rng.integers(2, nmax + 1)))
    else:
        sizes = rng.integers(2, nmax + 1, b)
    if big:
        sizes = np.minimum(sizes, 64)
    span = 30.0
    xs = [np.sort(rng.uniform(0, span, n)) if dim == 1 else rng.uniform(0, span, (n, 2)) for n in sizes]
    f = (lambda p: np.sin(p / 4.0)) if dim == 1 else (lambda p: np.sin(p[:, 0] / 4.0) * np.cos(p[:, 1] / 5.0))
    ys = [f(p) + 0.2 * rng.standard_normal(len(p)) for p in xs]
    yes = [rng.uniform(0.05, 0.4, len(p)) for p in xs] if rng.random() < 0.8 else None
    y0s = [0.3 * rng.standard_normal() + 0.1 * np.cos(np.arange(len(p))) for p in xs] if rng.random() < 0.5 else None
    nug = float(rng.choice([0.0, 0.02, 0.3])) if yes is not None else float(rng.choice([0.05, 0.3]))
    if dim == 1:
        hyp = [float(rng.uniform(0.3, 2.0)), float(rng.uniform(0.8, 6.0))]
    else:
        hyp = [float(rng.uniform(0.5, 1.5)), float(rng.uniform(2, 6)), float(rng.uniform(2, 6)), float(rng.uniform(-1.5, 1.5))]
    m = int(rng.integers(1, 70))
    if dim == 1:
        grid = np.linspace(-2, span + 2, m) if rng.random() < 0.6 else np.sort(rng.uniform(-2, span + 2, m))
    else:
        grid = rng.uniform(0, span, (m, 2))
    ny0 = rng.standard_normal((b, m)) if rng.random() < 0.5 else None
    return dim, xs, ys, yes, y0s, hyp, nug, grid, ny0


@pytest.mark.parametrize("seed", range(28))
def test_random_case(seed):
    from cosmogp_b200.batch import DeviceBatch, pack_csr
    dim, xs, ys, yes, y0s, hyp, nug, grid, ny0 = _case(seed)
    kind = "1d" if dim == 1 else "2d"
    x, off = pack_csr(xs, dim); y, _ = pack_csr(ys, 1)
    ye = pack_csr(yes, 1)[0] if yes is not None else None
    y0 = pack_csr(y0s, 1)[0] if y0s is not None else None
    batch = DeviceBatch(x, y, off, y0=y0, y_err=ye, dim=dim)
    tot, ll, info = batch.log_likelihood(hyp, nug)
    assert not info.any()
    mean, var, _ = batch.predict(hyp, nug, grid, new_y0=ny0)
    pred, pvar, pull, resid, _ = batch.loo(hyp, nug)
    b = len(xs)
    pick = sorted(set([0, b - 1] + list(np.random.default_rng(seed).integers(0, b, 4))))
    for i in pick:
        e = yes[i] if yes is not None else None
        m0 = y0s[i] if y0s is not None else None
        assert_close(ll[i], O.log_likelihood(ys[i], xs[i], hyp, nug, e, m0, kind=kind), RTOL, 1e-11, "ll %d" % i)
        mo, vo = O.predict(ys[i], xs[i], hyp, nug, grid, e, 0.0 if m0 is None else m0, 0.0 if ny0 is None else ny0[i],
                           kind=kind, full_cov=False)
        assert_close(mean[i], mo, RTOL, 1e-10, "mean %d" % i); assert_close(var[i], vo, RTOL, 1e-11, "var %d" % i)
        ez = np.zeros(len(ys[i])) if e is None else e
        if m0 is None:
            po = O.loo_closed_form(ys[i], xs[i], hyp, nug, ez, kind=kind)
        else:
            po = O.loo_closed_form(ys[i], xs[i], hyp, nug, ez, mean=m0, diff=0.0, kind=kind)
        s = slice(off[i], off[i + 1])
        assert_close(pred[s], po[0], RTOL, 1e-9, "loo pred %d" % i); assert_close(pull[s], po[2], 1e-8, 1e-8, "pull %d" % i)
    assert_close(tot, float(np.sum(ll)), 1e-12)
