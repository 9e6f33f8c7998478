"""GPU parity tests proper: the CUDA path, called through the C ABI with host buffers,
against (a) fixtures produced by the real reference and (b) the pinned oracle on
seeded inputs.  Tolerance: relative 1e-9 (BASELINE.json north_star) on
log-likelihood, predictions, variances and pulls, with a small absolute floor where
a quantity passes through zero."""
import ctypes as C

import numpy as np
import pytest

from conftest import assert_close, golden, split
from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.fixture(scope="module")
def L():
    from cosmogp_b200 import _lib
    _lib.require_device()
    return _lib


def csr(arrs):
    off = np.zeros(len(arrs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(a) for a in arrs])
    return np.ascontiguousarray(np.concatenate([np.asarray(a, dtype=np.float64) for a in arrs])), off


def run_ll(L, xs, ys, y0s, yes, hyp, nugget, dim=1, floor=0.0, flags=0):
    x, off = csr(xs); y, _ = csr(ys)
    y0 = csr(y0s)[0] if y0s is not None else None
    ye = csr(yes)[0] if yes is not None else None
    b = len(xs)
    ll = np.full(b, -7.0); info = np.full(b, -1, dtype=np.int32); tot = C.c_double(0)
    hyp = np.ascontiguousarray(hyp, dtype=np.float64)
    rc = L.lib().cgp_ll_batched_host(b, L.hptr(off), dim, L.hptr(x), L.hptr(y), L.hptr(y0), L.hptr(ye),
                                     L.hptr(hyp), nugget, floor, flags, L.hptr(ll), L.hptr(info), C.byref(tot))
    L.check(rc, "ll")
    return ll, info, tot.value, rc


def run_predict(L, xs, ys, y0s, yes, hyp, nugget, grid, new_y0=None, dim=1, goff=None, want_var=True, flags=0):
    x, off = csr(xs); y, _ = csr(ys)
    y0 = csr(y0s)[0] if y0s is not None else None
    ye = csr(yes)[0] if yes is not None else None
    b = len(xs)
    grid = np.ascontiguousarray(grid, dtype=np.float64)
    m = len(grid) if goff is None else 0
    nout = b * m if goff is None else int(goff[-1])
    mean = np.full(nout, -7.0); var = np.full(nout, -7.0) if want_var else None
    info = np.full(b, -1, dtype=np.int32)
    hyp = np.ascontiguousarray(hyp, dtype=np.float64)
    ny0 = None if new_y0 is None else np.ascontiguousarray(new_y0, dtype=np.float64)
    rc = L.lib().cgp_predict_batched_host(b, L.hptr(off), dim, L.hptr(x), L.hptr(y), L.hptr(y0), L.hptr(ye),
                                          L.hptr(hyp), nugget, 0.0, flags, L.hptr(grid),
                                          None if goff is None else L.hptr(goff), m, L.hptr(ny0),
                                          L.hptr(mean), L.hptr(var), L.hptr(info))
    L.check(rc, "predict")
    if goff is None:
        return mean.reshape(b, m), (var.reshape(b, m) if want_var else None), info
    return mean, var, info


def run_loo(L, xs, ys, ms, yes, hyp, nugget, mode=0, dim=1):
    x, off = csr(xs); y, _ = csr(ys)
    m = csr(ms)[0] if ms is not None else None
    ye = csr(yes)[0] if yes is not None else None
    b = len(xs); npt = int(off[-1])
    out = [np.full(npt, -7.0) for _ in range(4)]
    info = np.full(b, -1, dtype=np.int32)
    hyp = np.ascontiguousarray(hyp, dtype=np.float64)
    rc = L.lib().cgp_loo_batched_host(b, L.hptr(off), dim, L.hptr(x), L.hptr(y), L.hptr(m), L.hptr(ye),
                                      L.hptr(hyp), nugget, 0.0, 0, mode, *[L.hptr(o) for o in out], L.hptr(info))
    L.check(rc, "loo")
    return out + [info]


# --------------------------------------------------------------------------- golden fixtures
def test_kat_1d(L):
    g = golden("kat_1d")
    ll, info, tot, rc = run_ll(L, [g["x"]], [g["y"]], None, [g["y_err"]], g["hyp"], float(g["nugget"]))
    assert rc == 0 and info[0] == 0
    assert_close(ll[0], g["ll_chol"], RTOL); assert_close(tot, g["ll_chol"], RTOL)
    mean, var, _ = run_predict(L, [g["x"]], [g["y"]], None, [g["y_err"]], g["hyp"], float(g["nugget"]), g["grid"])
    assert_close(mean[0], g["mean"], RTOL); assert_close(var[0], np.diag(g["cov"]), RTOL)
    pred, pvar, pull, resid, _ = run_loo(L, [g["x"]], [g["y"]], None, [g["y_err"]], g["hyp"], float(g["nugget"]))
    assert_close(pull, g["pull"], RTOL, 1e-13); assert_close(pred, g["pred"], RTOL, 1e-13)
    assert_close(resid, g["resid"], RTOL, 1e-13)


def test_kat_2d(L):
    g = golden("kat_2d")
    nug = float(g["nugget"])
    ll, info, _, _ = run_ll(L, [g["x"]], [g["y"]], None, [g["y_err"]], g["hyp"], nug, dim=2)
    assert info[0] == 0
    assert_close(ll[0], g["ll_chol"], RTOL)
    mean, var, _ = run_predict(L, [g["x"]], [g["y"]], None, [g["y_err"]], g["hyp"], nug, g["grid"], dim=2)
    assert_close(mean[0], g["mean"], RTOL); assert_close(var[0], np.diag(g["cov"]), RTOL)
    pred, _, pull, resid, _ = run_loo(L, [g["x"]], [g["y"]], None, [g["y_err"]], g["hyp"], nug, dim=2)
    assert_close(pull, g["pull"], RTOL, 1e-12); assert_close(pred, g["pred"], RTOL, 1e-12)


def test_c1_single_light_curve(L):
    """BASELINE config 1: N=50 epochs, predict on a 500-point grid."""
    g = golden("c1_single")
    nug = float(g["nugget"])
    ll, info, _, _ = run_ll(L, [g["x"]], [g["y"]], None, [g["y_err"]], g["hyp"], nug)
    assert info[0] == 0
    assert_close(ll[0], g["ll_chol"], RTOL)
    mean, var, _ = run_predict(L, [g["x"]], [g["y"]], None, [g["y_err"]], g["hyp"], nug, g["grid"])
    assert_close(mean[0], g["mean"], RTOL, 1e-12); assert_close(var[0], g["cov_diag"], RTOL, 1e-13)
    _, _, pull, resid, _ = run_loo(L, [g["x"]], [g["y"]], None, [g["y_err"]], g["hyp"], nug)
    assert_close(pull, g["pull"], RTOL, 1e-11); assert_close(resid, g["resid"], RTOL, 1e-12)


def test_ragged_1d_shared_mean(L):
    g = golden("ragged_1d")
    off, hyp, nug = g["off"], g["hyp"], float(g["nugget"])
    xs, ys, yes, y0s = (split(g[k], off) for k in ("x", "y", "y_err", "y0"))
    ll, info, tot, rc = run_ll(L, xs, ys, y0s, yes, hyp, nug)
    assert rc == 0 and not info.any()
    assert_close(ll, g["ll_obj"], RTOL); assert_close(tot, g["ll_sum"], RTOL)
    _, _, tot7, _ = run_ll(L, xs, ys, y0s, yes, hyp, 0.07)
    assert_close(tot7, g["ll_sum_nugget007"], RTOL)
    ny0 = np.array([O.return_mean_1d(ys[i], xs[i], g["mean_y"], g["mean_x"], new_x=g["grid"]) for i in range(len(xs))])
    mean, var, _ = run_predict(L, xs, ys, y0s, yes, hyp, nug, g["grid"], new_y0=ny0)
    assert_close(mean, g["mean"], RTOL); assert_close(var, g["var"], RTOL, 1e-13)
    # own-epoch grid (new_binning=None) through the CSR grid path
    m0, v0, _ = run_predict(L, xs[:1], ys[:1], y0s[:1], yes[:1], hyp, 0.05, xs[0], new_y0=y0s[0],
                            goff=np.array([0, len(xs[0])], dtype=np.int64))
    assert_close(m0, g["own_mean0"], RTOL); assert_close(v0, np.diag(g["own_cov0"]), RTOL, 1e-13)
    # pulls: mode B (mean, diff None -> recentred) and mode C (diff given -> plain)
    tmpl = [O.return_mean_1d(ys[i], xs[i], g["mean_y"], g["mean_x"], diff=0.0) for i in range(len(xs))]
    pred, _, pull, resid, _ = run_loo(L, xs, ys, tmpl, yes, hyp, 0.05, mode=1)
    assert_close(pull, g["pullB"], 1e-8, 1e-11); assert_close(pred, g["predB"], RTOL, 1e-11)
    mc = [tmpl[i] + g["diff"][i] for i in range(len(xs))]
    pred, _, pull, resid, _ = run_loo(L, xs, ys, mc, yes, hyp, 0.05, mode=0)
    assert_close(pull, g["pullC"], 1e-8, 1e-11); assert_close(pred, g["predC"], RTOL, 1e-11)


def test_pulls_modes_a_d(L):
    g = golden("pulls_1d")
    hyp, nug = g["hyp"], float(g["nugget"])
    pred, _, pull, resid, info = run_loo(L, list(g["x"]), list(g["y"]), None, list(g["y_err"]), hyp, nug)
    assert not info.any()
    assert_close(pull, g["pullA"], RTOL, 1e-12); assert_close(pred, g["predA"].ravel(), RTOL, 1e-12)
    assert_close(resid, g["residA"], RTOL, 1e-12)
    sticky = np.ones(40) * np.mean(g["yD"][0])
    pred, _, pull, _, _ = run_loo(L, [g["xD"]] * 3, list(g["yD"]), [sticky] * 3, list(g["y_err"][:3]), hyp, nug, mode=1)
    assert_close(pull, g["pullD"], 1e-8, 1e-11); assert_close(pred, g["predD"].ravel(), RTOL, 1e-11)


def test_batch_2d(L):
    g = golden("batch_2d")
    off, hyp, nug = g["off"], g["hyp"], float(g["nugget"])
    xs, ys, yes = split(g["x"], off), split(g["y"], off), split(g["y_err"], off)
    ll, info, tot, _ = _ll2d(L, xs, ys, yes, hyp, nug)
    assert not info.any()
    assert_close(ll, g["ll_obj"], RTOL); assert_close(tot, g["ll_sum"], RTOL)


def _ll2d(L, xs, ys, yes, hyp, nug):
    b = len(xs)
    off = np.zeros(b + 1, dtype=np.int64); off[1:] = np.cumsum([len(v) for v in ys])
    x = np.ascontiguousarray(np.concatenate(xs), dtype=np.float64)      # (sumN, 2) row-major = interleaved
    y = np.ascontiguousarray(np.concatenate(ys)); ye = np.ascontiguousarray(np.concatenate(yes))
    ll = np.zeros(b); info = np.zeros(b, dtype=np.int32); tot = C.c_double(0)
    hyp = np.ascontiguousarray(hyp, dtype=np.float64)
    rc = L.lib().cgp_ll_batched_host(b, L.hptr(off), 2, L.hptr(x), L.hptr(y), None, L.hptr(ye), L.hptr(hyp), nug, 0.0, 0,
                                     L.hptr(ll), L.hptr(info), C.byref(tot))
    L.check(rc, "ll2d")
    return ll, info, tot.value, rc


def test_batch_2d_predict_and_pulls(L):
    g = golden("batch_2d")
    off, hyp, nug = g["off"], g["hyp"], float(g["nugget"])
    xs, ys, yes = split(g["x"], off), split(g["y"], off), split(g["y_err"], off)
    b = 3
    x = np.ascontiguousarray(np.concatenate(xs)); y = np.concatenate(ys); ye = np.concatenate(yes)
    grid = np.ascontiguousarray(g["grid"]); m = len(grid)
    mean = np.zeros(b * m); var = np.zeros(b * m); info = np.zeros(b, dtype=np.int32)
    hp = np.ascontiguousarray(hyp)
    rc = L.lib().cgp_predict_batched_host(b, L.hptr(off), 2, L.hptr(x), L.hptr(y), None, L.hptr(ye), L.hptr(hp), nug, 0.0, 0,
                                          L.hptr(grid), None, m, None, L.hptr(mean), L.hptr(var), L.hptr(info))
    L.check(rc, "predict2d")
    assert_close(mean.reshape(b, m), g["mean"], RTOL, 1e-12); assert_close(var.reshape(b, m), g["var"], RTOL, 1e-12)
    o2 = np.array([0, len(ys[2])], dtype=np.int64)
    x2 = np.ascontiguousarray(xs[2]); outs = [np.zeros(len(ys[2])) for _ in range(4)]
    rc = L.lib().cgp_loo_batched_host(1, L.hptr(o2), 2, L.hptr(x2), L.hptr(np.ascontiguousarray(ys[2])), None,
                                      L.hptr(np.ascontiguousarray(yes[2])), L.hptr(hp), nug, 0.0, 0, 0,
                                      *[L.hptr(o) for o in outs], L.hptr(info))
    L.check(rc, "loo2d")
    assert_close(outs[2], g["pull2"], RTOL, 1e-12); assert_close(outs[0], g["pred2"], RTOL, 1e-12)


def test_notebook_joint_ll(L):
    """docs/notebook/1D_kernel_example_with_noise.ipynb: LL of the 100x60 batch at the fitted optimum."""
    g = golden("notebook_with_noise")
    _, info, tot, _ = run_ll(L, list(g["x"]), list(g["y"]), None, list(g["y_err"]), g["fit_joint"], 0.0)
    assert not info.any()
    assert_close(tot, g["ll_at_joint"], RTOL)


# --------------------------------------------------------------------------- oracle on seeded inputs
@pytest.mark.parametrize("n", [1, 2, 7, 8, 9, 33, 60, 64])
def test_sizes_one_warp(L, n):
    rng = np.random.default_rng(100 + n)
    b = 37
    xs = [np.sort(rng.uniform(-10, 40, n)) for _ in range(b)]
    ys = [rng.standard_normal(n) for _ in range(b)]
    yes = [rng.uniform(0.1, 0.3, n) for _ in range(b)]
    y0s = [rng.standard_normal(n) * 0.1 for _ in range(b)]
    hyp, nug = [0.7, 3.0], 0.05
    ll, info, tot, _ = run_ll(L, xs, ys, y0s, yes, hyp, nug)
    assert not info.any()
    ref = [O.log_likelihood(ys[i], xs[i], hyp, nug, yes[i], y0s[i]) for i in range(b)]
    assert_close(ll, ref, RTOL)
    grid = np.linspace(-12, 42, 45)
    mean, var, _ = run_predict(L, xs, ys, y0s, yes, hyp, nug, grid)
    for i in range(0, b, 6):
        mo, vo = O.predict(ys[i], xs[i], hyp, nug, grid, yes[i], y0s[i], full_cov=False)
        assert_close(mean[i], mo, RTOL, 1e-12); assert_close(var[i], vo, RTOL, 1e-13)
    if n >= 2:
        pred, pvar, pull, resid, _ = run_loo(L, xs, ys, None, yes, hyp, nug)
        po = [O.loo_closed_form(ys[i], xs[i], hyp, nug, yes[i]) for i in range(b)]
        assert_close(pull, np.concatenate([p[2] for p in po]), RTOL, 1e-12)
        assert_close(pvar, np.concatenate([p[1] for p in po]), RTOL, 1e-13)


@pytest.mark.parametrize("n", [65, 100, 128, 129, 200, 224])
def test_sizes_cta_per_object(L, n):
    rng = np.random.default_rng(200 + n)
    b = 5
    xs = [np.sort(rng.uniform(-10, 40, n)) for _ in range(b)]
    ys = [rng.standard_normal(n) for _ in range(b)]
    yes = [rng.uniform(0.2, 0.4, n) for _ in range(b)]
    hyp, nug = [0.7, 1.5], 0.05
    ll, info, _, _ = run_ll(L, xs, ys, None, yes, hyp, nug)
    assert not info.any()
    assert_close(ll, [O.log_likelihood(ys[i], xs[i], hyp, nug, yes[i]) for i in range(b)], RTOL)
    grid = np.linspace(-12, 42, 37)
    mean, var, _ = run_predict(L, xs, ys, None, yes, hyp, nug, grid)
    mo, vo = O.predict(ys[1], xs[1], hyp, nug, grid, yes[1], full_cov=False)
    assert_close(mean[1], mo, RTOL, 1e-12); assert_close(var[1], vo, RTOL, 1e-12)
    pred, pvar, pull, resid, _ = run_loo(L, xs, ys, None, yes, hyp, nug)
    po = O.loo_closed_form(ys[0], xs[0], hyp, nug, yes[0])
    assert_close(pull[:n], po[2], RTOL, 1e-11)


def test_ragged_mixed_sizes_and_empty(L):
    rng = np.random.default_rng(7)
    sizes = [60, 0, 3, 64, 17, 1, 40, 0, 59]
    xs = [np.sort(rng.uniform(0, 30, n)) for n in sizes]
    ys = [rng.standard_normal(n) for n in sizes]
    yes = [rng.uniform(0.1, 0.2, n) for n in sizes]
    hyp, nug = [1.1, 2.5], 0.0
    ll, info, tot, _ = run_ll(L, xs, ys, None, yes, hyp, nug)
    assert not info.any()
    ref = [O.log_likelihood(ys[i], xs[i], hyp, nug, yes[i]) if sizes[i] else 0.0 for i in range(len(sizes))]
    assert_close(ll, ref, RTOL, 1e-300)


def test_not_positive_definite_is_reported(L):
    """Duplicate epochs with zero noise: scipy.linalg.cholesky raises (inv_matrix.py:23);
    the C ABI returns the count, a LAPACK-style info and NaN outputs, other objects unaffected."""
    x_bad = np.array([0.0, 1.0, 1.0, 2.0]); y = np.array([0.1, 0.2, 0.3, 0.4])
    x_ok = np.array([0.0, 1.0, 1.5, 2.0])
    ll, info, tot, rc = run_ll(L, [x_ok, x_bad, x_ok], [y, y, y], None, None, [1.0, 1.0], 0.0)
    assert rc == 1 and info[0] == 0 and info[2] == 0 and info[1] == 3
    assert np.isnan(ll[1]) and np.isfinite(ll[0]) and np.isnan(tot)
    with pytest.raises(np.linalg.LinAlgError):
        O.log_likelihood(y, x_bad, [1.0, 1.0], 0.0)


def test_batched_c2_shape_vs_batched_oracle(L):
    """BASELINE config 2 in miniature (2,000 of the 10^5 light curves x 60 epochs, shared mean)."""
    rng = np.random.default_rng(2)
    b, n = 2000, 60
    x = np.sort(rng.uniform(-10, 40, (b, n)), axis=1)
    ye = np.full((b, n), 0.2)
    y0 = -18 + 2 * np.sin(x / 10) + rng.normal(0, 0.3, (b, 1))
    y = y0 + 0.5 * rng.standard_normal((b, n))
    hyp, nug = [0.5, 2.0], 0.0
    ll, info, tot, _ = run_ll(L, list(x), list(y), list(y0), list(ye), hyp, nug)
    ref = O.ll_batched_1d(x, y, y0, ye, hyp, nug)
    assert not info.any()
    assert_close(ll, ref, RTOL); assert_close(tot, ref.sum(), RTOL)
    grid = np.linspace(-10, 40, 100)
    mean, var, _ = run_predict(L, list(x), list(y), list(y0), list(ye), hyp, nug, grid, new_y0=np.zeros((b, 100)))
    mo, vo = O.predict_batched_1d(x, y, y0, ye, hyp, nug, grid, np.zeros((b, 100)))
    assert_close(mean, mo, RTOL, 1e-12); assert_close(var, vo, RTOL, 1e-13)


def test_c2_full_size_ll_against_batched_oracle(L):
    """BASELINE config 2 at full size: 10^5 light curves x 60 epochs, per-object LL against the
    batched numpy oracle (rel 1e-9), plus the size-independent checks that the per-object values
    do not depend on batch position (a permuted batch gives the permuted result bit for bit)."""
    from cosmogp_b200.batch import DeviceBatch
    rng = np.random.default_rng(2)
    b, n = 100000, 60
    x = np.sort(rng.uniform(-10, 40, (b, n)), axis=1)
    ye = np.full((b, n), 0.2)
    y0 = -18 + 2 * np.sin(x / 10) + rng.normal(0, 0.3, (b, 1))
    y = y0 + 0.5 * rng.standard_normal((b, n))
    hyp, nug = [0.5, 2.0], 0.0
    off = np.arange(b + 1, dtype=np.int64) * n
    batch = DeviceBatch(x.ravel(), y.ravel(), off, y0=y0.ravel(), y_err=ye.ravel())
    tot, ll, info = batch.log_likelihood(hyp, nug)
    assert not info.any()
    ref = O.ll_batched_1d(x, y, y0, ye, hyp, nug)
    assert_close(ll, ref, RTOL); assert_close(tot, np.add.accumulate(ref)[-1], RTOL)
    perm = rng.permutation(b)
    b2 = DeviceBatch(x[perm].ravel(), y[perm].ravel(), off, y0=y0[perm].ravel(), y_err=ye[perm].ravel())
    _, ll2, _ = b2.log_likelihood(hyp, nug)
    assert np.array_equal(ll2, ll[perm])


def test_c2_full_size_prediction(L):
    """BASELINE config 2 at full size, prediction side: 10^5 x 60 on a 100-point grid with a shared mean.  A sample of
    objects against the batched oracle (1e-9); the uniform-grid kernel against the general one on every object;
    a permuted batch gives the permuted result bit for bit; the likelihood emitted by the factor kernel against the
    LL kernel; the pipelined host-resident evaluator gives the resident results bit for bit."""
    import torch
    from cosmogp_b200.batch import DeviceBatch, StreamedEvaluator
    rng = np.random.default_rng(6)
    b, n, m = 100000, 60, 100
    x = np.sort(rng.uniform(-10, 40, (b, n)), axis=1); ye = np.full((b, n), 0.2)
    diff = rng.normal(0, 0.3, b)
    y0 = -18 + 2 * np.sin(x / 10) + diff[:, None]
    y = y0 + 0.5 * rng.standard_normal((b, n))
    grid = np.linspace(-10, 40, m); tmpl = -18 + 2 * np.sin(grid / 10)
    hyp, nug = [0.5, 2.0], 0.0
    off = np.arange(b + 1, dtype=np.int64) * n
    batch = DeviceBatch(x.ravel(), y.ravel(), off, y0=y0.ravel(), y_err=ye.ravel())
    fac = batch.factor_dev(hyp, nug, want_ll=True)
    g = torch.from_numpy(grid).cuda(); packed = torch.from_numpy(np.concatenate([tmpl, diff])).cuda()
    mu, vu, _ = batch.predict_factored_dev(fac, g, None, packed, True, template_mean=True, uniform_grid=True)
    mg, vg, _ = batch.predict_factored_dev(fac, g, None, packed, True, template_mean=True, uniform_grid=False)
    mu = mu.cpu().numpy().reshape(b, m); vu = vu.cpu().numpy().reshape(b, m)
    assert_close(mu, mg.cpu().numpy().reshape(b, m), 1e-10, 1e-10); assert_close(vu, vg.cpu().numpy().reshape(b, m), 1e-10, 1e-11)
    sel = rng.choice(b, 256, replace=False)
    mo, vo = O.predict_batched_1d(x[sel], y[sel], y0[sel], ye[sel], hyp, nug, grid, tmpl[None, :] + diff[sel, None])
    assert_close(mu[sel], mo, RTOL, 1e-11); assert_close(vu[sel], vo, RTOL, 1e-12)
    _, ll, _ = batch.log_likelihood(hyp, nug)
    assert_close(fac["ll"].cpu().numpy(), ll, 1e-13, 1e-11)
    perm = rng.permutation(b)
    b2 = DeviceBatch(x[perm].ravel(), y[perm].ravel(), off, y0=y0[perm].ravel(), y_err=ye[perm].ravel())
    f2 = b2.factor_dev(hyp, nug)
    p2 = torch.from_numpy(np.concatenate([tmpl, diff[perm]])).cuda()
    m2, v2, _ = b2.predict_factored_dev(f2, g, None, p2, True, template_mean=True, uniform_grid=True)
    assert np.array_equal(m2.cpu().numpy().reshape(b, m), mu[perm]) and np.array_equal(v2.cpu().numpy().reshape(b, m), vu[perm])
    ev = StreamedEvaluator(b, n, m, n_chunks=8, n_streams=4, shared_mean=True)
    for k, v in (("x", x), ("y", y), ("y0", y0), ("y_err", ye), ("template", tmpl), ("diff", diff)):
        ev.host(k)[...] = v
    tot, lls, ms, vs, infos = ev.run(hyp, nug, grid)
    assert not infos.any() and np.array_equal(ms, mu) and np.array_equal(vs, vu)
    assert np.array_equal(lls, fac["ll"].cpu().numpy()) and tot == float(np.add.accumulate(lls)[-1])


def test_c5_pulls_against_batched_oracle(L):
    """BASELINE config 5 recipe (N = 40, y_err = 0.1) on 20,000 of the 10^6 objects: closed-form LOO pulls
    against the batched oracle; and pulls of a Gaussian process drawn from the model are ~N(0,1)."""
    from cosmogp_b200.batch import DeviceBatch
    rng = np.random.default_rng(5)
    b, n = 20000, 40
    x = np.sort(rng.uniform(-10, 10, (b, n)), axis=1)
    ye = np.full((b, n), 0.1)
    y = 0.5 * np.sin(x / 2.0 + rng.uniform(0, 6.28, (b, 1))) + 0.1 * rng.standard_normal((b, n))
    hyp, nug = [0.5, 2.0], 0.0
    batch = DeviceBatch(x.ravel(), y.ravel(), np.arange(b + 1, dtype=np.int64) * n, y_err=ye.ravel())
    pred, pvar, pull, resid, info = batch.loo(hyp, nug)
    assert not info.any()
    po = O.loo_batched_1d(x, y, ye, hyp, nug)
    assert_close(pred, po[0].ravel(), RTOL, 1e-12); assert_close(pvar, po[1].ravel(), RTOL, 1e-14)
    assert_close(pull, po[2].ravel(), RTOL, 1e-11); assert_close(resid, po[3].ravel(), RTOL, 1e-12)


def test_factor_once_predict_many(L):
    """cgp_factor_batched_dev + cgp_predict_factored_dev (TMA-staged factor) == fused prediction,
    for two different grids from one factorisation; also through the internal split of large batches."""
    import torch
    from cosmogp_b200.batch import DeviceBatch
    rng = np.random.default_rng(12)
    b, n = 3000, 50                        # >= 2048 objects: cgp_predict_batched_dev takes the two-kernel route
    x = np.sort(rng.uniform(-10, 40, (b, n)), axis=1); y = rng.standard_normal((b, n)); ye = rng.uniform(0.1, 0.3, (b, n))
    y0 = 0.1 * rng.standard_normal((b, n))
    hyp, nug = [0.6, 2.5], 0.04
    batch = DeviceBatch(x.ravel(), y.ravel(), np.arange(b + 1, dtype=np.int64) * n, y0=y0.ravel(), y_err=ye.ravel())
    fac = batch.factor_dev(hyp, nug, want_ll=True)
    ll_fac = fac["ll"].cpu().numpy()                    # the likelihood from the same factorisation
    _, ll_ref, _ = batch.log_likelihood(hyp, nug)
    assert_close(ll_fac, ll_ref, 1e-13, 1e-11)
    for i in (0, 1234, 2999):
        assert_close(ll_fac[i], O.log_likelihood(y[i], x[i], hyp, nug, ye[i], y0[i]), RTOL)
    for m in (37, 100):
        grid = np.linspace(-12, 42, m); ny0 = rng.standard_normal((b, m))
        mean, var, info = batch.predict_factored_dev(fac, torch.from_numpy(grid).cuda(), None, torch.from_numpy(ny0).cuda(), True)
        mean = mean.cpu().numpy().reshape(b, m); var = var.cpu().numpy().reshape(b, m)
        m2, v2, _ = batch.predict(hyp, nug, grid, new_y0=ny0)
        assert_close(mean, m2, 1e-10, 1e-11); assert_close(var, v2, 1e-10, 1e-11)
        for i in (0, 1234, 2999):
            mo, vo = O.predict(y[i], x[i], hyp, nug, grid, ye[i], y0[i], ny0[i], full_cov=False)
            assert_close(mean[i], mo, RTOL, 1e-12); assert_close(var[i], vo, RTOL, 1e-13)
    small = DeviceBatch(x[:5].ravel(), y[:5].ravel(), np.arange(6, dtype=np.int64) * n, y0=y0[:5].ravel(), y_err=ye[:5].ravel())
    grid = np.linspace(-12, 42, 500)
    ms, vs, _ = small.predict(hyp, nug, grid)                       # few objects: fused kernel with grid split
    fs = small.factor_dev(hyp, nug)
    mf, vf, _ = small.predict_factored_dev(fs, torch.from_numpy(grid).cuda(), None, None, True)
    assert_close(mf.cpu().numpy().reshape(5, 500), ms, 1e-13, 1e-13); assert_close(vf.cpu().numpy().reshape(5, 500), vs, 1e-13, 1e-14)


def test_shared_mean_template_flag(L):
    """CGP_MEAN_TEMPLATE: new_y0 = [template on the shared grid | one offset per object] gives exactly the
    results of the materialised (n_obj x M) mean -- fused kernel, generic kernel (N > 64), the two-kernel
    route of large batches, the factored entry point, and 2D; per-object grids are refused."""
    import torch
    from cosmogp_b200.batch import DeviceBatch
    rng = np.random.default_rng(5)
    hyp, nug = [0.6, 2.5], 0.04
    for b, n, m in ((7, 30, 45), (5, 100, 33), (2500, 40, 50)):
        x = np.sort(rng.uniform(-10, 40, (b, n)), axis=1); y = rng.standard_normal((b, n)); ye = rng.uniform(0.1, 0.3, (b, n))
        grid = np.linspace(-12, 42, m); tmpl = np.sin(grid / 7.0); diff = rng.standard_normal(b)
        full = tmpl[None, :] + diff[:, None]
        m1, v1, _ = run_predict(L, list(x), list(y), None, list(ye), hyp, nug, grid, new_y0=full)
        m2, v2, _ = run_predict(L, list(x), list(y), None, list(ye), hyp, nug, grid, new_y0=np.concatenate([tmpl, diff]),
                                flags=L.CGP_MEAN_TEMPLATE)
        assert np.array_equal(m1, m2) and np.array_equal(v1, v2)
        mo, _ = O.predict(y[b // 2], x[b // 2], hyp, nug, grid, ye[b // 2], np.zeros(n), full[b // 2], full_cov=False)
        assert_close(m2[b // 2], mo, RTOL, 1e-12)
        if n <= 64:
            batch = DeviceBatch(x.ravel(), y.ravel(), np.arange(b + 1, dtype=np.int64) * n, y_err=ye.ravel())
            fac = batch.factor_dev(hyp, nug)
            packed = torch.from_numpy(np.concatenate([tmpl, diff])).cuda()
            m3, v3, _ = batch.predict_factored_dev(fac, torch.from_numpy(grid).cuda(), None, packed, True, template_mean=True)
            m4, v4, _ = batch.predict_factored_dev(fac, torch.from_numpy(grid).cuda(), None, torch.from_numpy(full).cuda(), True)
            assert torch.equal(m3, m4) and torch.equal(v3, v4)
            m5, _, _ = batch.predict(hyp, nug, grid, mean_template=(tmpl, diff))
            assert_close(m5, m1, 1e-10, 1e-11)          # (predict() hints a uniform grid: recurrence kernel at >= 2048 objects)
    b, n, m = 9, 20, 30                                   # 2D
    xy = rng.uniform(0, 10, (b, n, 2)); z = rng.standard_normal((b, n)); ze = np.full((b, n), 0.2)
    grid = rng.uniform(0, 10, (m, 2)); tmpl = rng.standard_normal(m); diff = rng.standard_normal(b)
    h2 = [1.0, 2.0, 1.5, 0.3]
    m1, v1, _ = run_predict(L, list(xy), list(z), None, list(ze), h2, 0.05, grid, new_y0=tmpl[None, :] + diff[:, None], dim=2)
    m2, v2, _ = run_predict(L, list(xy), list(z), None, list(ze), h2, 0.05, grid, new_y0=np.concatenate([tmpl, diff]), dim=2,
                            flags=L.CGP_MEAN_TEMPLATE)
    assert np.array_equal(m1, m2) and np.array_equal(v1, v2)
    x = [np.arange(5.0)]; goff = np.array([0, 5], dtype=np.int64)
    with pytest.raises(RuntimeError):
        run_predict(L, x, [np.ones(5)], None, None, hyp, nug, np.arange(5.0), new_y0=np.zeros(6), goff=goff, flags=L.CGP_MEAN_TEMPLATE)


def test_uniform_grid_fast_path(L):
    """CGP_GRID_UNIFORM: the recurrence kernel (two exps per data point and 16 grid rows) against the general
    one -- every tile count, ragged last grid block, the few-objects split, points far outside the grid
    (underflowing anchors), and the fall-backs (non-uniform grid, l < spacing, per-object grids): identical there."""
    import torch
    from cosmogp_b200.batch import DeviceBatch
    rng = np.random.default_rng(17)

    def both(x, y, ye, hyp, nug, grid, ny0=None):
        b, n = x.shape
        batch = DeviceBatch(x.ravel(), y.ravel(), np.arange(b + 1, dtype=np.int64) * n, y_err=ye.ravel())
        fac = batch.factor_dev(hyp, nug)
        g = torch.from_numpy(np.ascontiguousarray(grid)).cuda()
        t = None if ny0 is None else torch.from_numpy(ny0).cuda()
        out = []
        for uni in (False, True):
            m, v, _ = batch.predict_factored_dev(fac, g, None, t, True, uniform_grid=uni)
            out.append((m.cpu().numpy().reshape(b, len(grid)), v.cpu().numpy().reshape(b, len(grid))))
        return out

    for b, n, m in ((3000, 60, 100), (7, 5, 17), (40, 20, 2), (300, 33, 64), (5, 64, 500), (2100, 48, 37)):
        x = np.sort(rng.uniform(-10, 40, (b, n)), axis=1); y = rng.standard_normal((b, n)); ye = rng.uniform(0.1, 0.3, (b, n))
        grid = np.linspace(-12, 42, m)
        (m0, v0), (m1, v1) = both(x, y, ye, [0.6, 2.5], 0.04, grid, rng.standard_normal((b, m)))
        assert_close(m1, m0, 1e-10, 1e-11); assert_close(v1, v0, 1e-10, 1e-11)
        assert not np.array_equal(m1, m0) or grid[1] - grid[0] > 2.5     # the fast path really ran (needs l >= spacing)
        i = b // 2
        mo, vo = O.predict(y[i], x[i], [0.6, 2.5], 0.04, grid, ye[i], full_cov=False)
        (m2, v2), (m3, v3) = both(x, y, ye, [0.6, 2.5], 0.04, grid)
        assert_close(m3[i], mo, RTOL, 1e-12); assert_close(v3[i], vo, RTOL, 1e-13)
    # a descending grid is uniform too (negative spacing)
    x = np.sort(rng.uniform(-10, 40, (2200, 30)), axis=1); y = rng.standard_normal((2200, 30)); ye = np.full((2200, 30), 0.2)
    (m0, v0), (m1, v1) = both(x, y, ye, [0.6, 2.5], 0.04, np.linspace(42, -12, 90))
    assert_close(m1, m0, 1e-10, 1e-11); assert_close(v1, v0, 1e-10, 1e-11); assert not np.array_equal(m1, m0)
    # data far from the grid and a short length scale: anchors underflow, true values are negligible
    b, n, m = 2500, 30, 100
    x = np.sort(rng.uniform(-400, 400, (b, n)), axis=1); y = rng.standard_normal((b, n)); ye = np.full((b, n), 0.2)
    grid = np.linspace(-30, 30, m)                         # spacing 0.606
    (m0, v0), (m1, v1) = both(x, y, ye, [1.0, 0.7], 0.0, grid)
    assert np.isfinite(m1).all() and np.isfinite(v1).all()
    assert_close(m1, m0, 1e-10, 1e-11); assert_close(v1, v0, 1e-10, 1e-11)
    # fall-backs give the general kernel's bits: l < spacing, a grid that is not uniform
    (m0, v0), (m1, v1) = both(x, y, ye, [1.0, 0.5], 0.0, grid)
    assert np.array_equal(m0, m1) and np.array_equal(v0, v1)
    bent = grid.copy(); bent[37] += 1e-9
    (m0, v0), (m1, v1) = both(x, y, ye, [1.0, 2.0], 0.0, bent)
    assert np.array_equal(m0, m1) and np.array_equal(v0, v1)
    # the host entry point takes the hint as well (large batch: two-kernel route inside)
    x = np.sort(rng.uniform(-10, 40, (2500, 40)), axis=1); y = rng.standard_normal((2500, 40)); ye = np.full((2500, 40), 0.2)
    grid = np.linspace(-10, 40, 50)
    ma, va, _ = run_predict(L, list(x), list(y), None, list(ye), [0.6, 2.5], 0.04, grid)
    mb, vb, _ = run_predict(L, list(x), list(y), None, list(ye), [0.6, 2.5], 0.04, grid, flags=L.CGP_GRID_UNIFORM)
    assert_close(mb, ma, 1e-10, 1e-11); assert_close(vb, va, 1e-10, 1e-11); assert not np.array_equal(ma, mb)


def test_degenerate_inputs(L):
    """Empty batches, empty grids, NaN / singular hyperparameters: defined outputs, no crash."""
    rng = np.random.default_rng(3)
    x = np.sort(rng.uniform(0, 10, 20)); y = rng.standard_normal(20); ye = np.full(20, 0.2)
    # no objects at all
    off0 = np.zeros(1, dtype=np.int64); tot_c = C.c_double(7.0)
    hyp = np.array([1.0, 1.0])
    rc = L.lib().cgp_ll_batched_host(0, L.hptr(off0), 1, None, None, None, None, L.hptr(hyp), 0.0, 0.0, 0, None, None, C.byref(tot_c))
    assert rc == 0 and tot_c.value == 0.0
    # empty grid
    mean, var, info = run_predict(L, [x], [y], None, [ye], [1.0, 2.0], 0.0, np.zeros(0))
    assert mean.shape == (1, 0) and info[0] == 0
    # NaN hyperparameter: every pivot is NaN -> reported as not positive definite, outputs NaN
    ll, info, tot, rc = run_ll(L, [x], [y], None, [ye], [np.nan, 2.0], 0.0)
    assert rc == 1 and info[0] == 1 and np.isnan(ll[0])
    # 2D metric that is not positive definite (l_x^2 l_y^2 - l_xy^2 < 0): scipy gives NaN distances
    x2 = rng.uniform(0, 10, (12, 2))
    off = np.array([0, 12], dtype=np.int64); ll2 = np.zeros(1); inf2 = np.zeros(1, dtype=np.int32); t2 = C.c_double(0)
    h2 = np.array([1.0, 1.0, 1.0, 5.0])
    rc = L.lib().cgp_ll_batched_host(1, L.hptr(off), 2, L.hptr(np.ascontiguousarray(x2)), L.hptr(np.ascontiguousarray(y[:12])), None,
                                     L.hptr(np.ascontiguousarray(ye[:12])), L.hptr(h2), 0.0, 0.0, 0, L.hptr(ll2), L.hptr(inf2), C.byref(t2))
    assert rc >= 0                                                          # must not crash; the value is garbage-in
    # huge length scale: K is numerically singular without noise -> flagged; with noise -> finite and equal to the oracle
    ll, info, _, rc = run_ll(L, [x], [y], None, [ye], [1.0, 1e6], 0.0)
    assert info[0] == 0
    assert_close(ll[0], O.log_likelihood(y, x, [1.0, 1e6], 0.0, ye), 1e-7)
    # bad dim / NULL pointers are argument errors, not crashes
    rc = L.lib().cgp_ll_batched_host(1, L.hptr(off), 3, L.hptr(x), L.hptr(y), None, None, L.hptr(hyp), 0.0, 0.0, 0, L.hptr(ll2), L.hptr(inf2), None)
    assert rc < 0 and b"dim" in L.lib().cgp_last_error()
    rc = L.lib().cgp_ll_batched_host(1, L.hptr(off), 1, None, L.hptr(y), None, None, L.hptr(hyp), 0.0, 0.0, 0, L.hptr(ll2), L.hptr(inf2), None)
    assert rc < 0
    # object too large for the shared-memory path: explicit size error
    big = np.zeros(300); offb = np.array([0, 300], dtype=np.int64)
    rc = L.lib().cgp_ll_batched_host(1, L.hptr(offb), 1, L.hptr(big), L.hptr(big), None, None, L.hptr(hyp), 0.0, 0.0, 0, L.hptr(ll2), L.hptr(inf2), None)
    assert rc == -2 and b"exceeds" in L.lib().cgp_last_error()
