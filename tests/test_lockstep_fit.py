"""Lock-step Nelder-Mead (cosmogp_b200.fit) against scipy.optimize.fmin, CPU only, and the
per-object device fits against scipy + oracle (GPU)."""
import numpy as np
import pytest
from scipy.optimize import fmin

from conftest import assert_close, golden
from cosmogp_b200.fit import nelder_mead_lockstep


def test_lockstep_matches_scipy_fmin_exactly():
    """Same decisions, same arithmetic: every object must end where scipy's fmin ends, bit for bit."""
    rng = np.random.default_rng(0)
    b = 40
    a = rng.uniform(0.5, 3.0, b); c = rng.uniform(-2, 2, (b, 3))

    def f_one(x, i):
        return a[i] * (x[0] - c[i, 0]) ** 2 + (x[1] - c[i, 1]) ** 4 + 2.0 * (x[2] - c[i, 2]) ** 2 + np.sin(3 * x[0]) * 0.1

    def fun(X, idx):
        return np.array([f_one(X[k], idx[k]) for k in range(len(idx))])

    x0 = rng.uniform(-1, 1, (b, 3)); x0[3, 1] = 0.0          # a zero coordinate exercises zdelt
    x, fv, its, calls = nelder_mead_lockstep(fun, x0)
    for i in range(b):
        ref = fmin(lambda v: f_one(v, i), x0[i], disp=False, full_output=True)
        assert np.array_equal(x[i], ref[0]), (i, x[i], ref[0])
        assert fv[i] == ref[1] and its[i] == ref[2] and calls[i] == ref[3]


def test_lockstep_handles_inf_and_nan():
    def fun(X, idx):
        f = (X[:, 0] - 1.0) ** 2 + (X[:, 1] + 0.5) ** 2
        f[X[:, 0] < 0] = np.inf
        return f
    x, fv, _, _ = nelder_mead_lockstep(fun, np.array([[0.5, 0.5], [2.0, -1.0]]))
    assert_close(x, [[1.0, -0.5], [1.0, -0.5]], 0, 2e-4)


@pytest.mark.gpu
def test_per_object_fits_match_reference_loop():
    """docs/notebook/1D_kernel_example_with_noise.ipynb cell 13: one fit per object.  Object 0 must
    reproduce the notebook's printed single-object optimum; a sample of the others is checked
    against scipy.fmin on the oracle likelihood (agreement to the optimiser tolerance)."""
    import cosmogp_b200 as cg
    from oracle import gp_oracle as O
    g = golden("notebook_with_noise")
    gp = cg.gaussian_process_nobject(g["y"], g["x"], y_err=g["y_err"])
    gp.find_hyperparameters_per_object(hyperparameter_guess=[0.5, 2])
    assert gp.hyperparameters_per_object.shape == (100, 2)
    assert_close(gp.hyperparameters_per_object[0], g["printed_single"], 1e-4)
    for i in (1, 17, 58, 99):
        ref = np.abs(fmin(lambda h: -O.log_likelihood(g["y"][i], g["x"][i], h, 0.0, g["y_err"][i]), [0.5, 2.0], disp=False))
        assert_close(gp.hyperparameters_per_object[i], ref, 2e-4)
        assert_close(gp.log_likelihood_per_object[i], O.log_likelihood(g["y"][i], g["x"][i], ref, 0.0, g["y_err"][i]), 1e-7)
    gn = cg.gaussian_process_nobject(g["y"][:8], g["x"][:8], y_err=g["y_err"][:8])
    gn.find_hyperparameters_per_object(hyperparameter_guess=[0.5, 2], nugget=True)
    i = 3
    ref = np.abs(fmin(lambda h: -O.log_likelihood(g["y"][i], g["x"][i], h[:2], h[2], g["y_err"][i]), [0.5, 2.0, 1.0], disp=False))
    assert_close(list(gn.hyperparameters_per_object[i]) + [gn.nugget_per_object[i]], ref, 5e-3, 5e-4)


@pytest.mark.gpu
def test_per_object_predict_and_pulls():
    """Each object predicted / pulled with its own hyperparameters (cgp_predict_objhyp_dev,
    cgp_loo_objhyp_dev), through the two-kernel route (>= 2048 objects) and the fused one."""
    import cosmogp_b200 as cg
    from oracle import gp_oracle as O
    rng = np.random.default_rng(21)
    for b in (2500, 40):
        n = 30
        x = np.sort(rng.uniform(0, 20, (b, n)), axis=1); y = rng.standard_normal((b, n)); ye = rng.uniform(0.1, 0.3, (b, n))
        hyp = np.column_stack([rng.uniform(0.4, 1.2, b), rng.uniform(1.0, 4.0, b)]); nug = rng.uniform(0.0, 0.1, b)
        gp = cg.gaussian_process_nobject(y, x, y_err=ye)
        gp.hyperparameters_per_object, gp.nugget_per_object = hyp, nug
        grid = np.linspace(-1, 21, 50)
        gp.get_prediction(new_binning=grid, COV='diag', per_object=True)
        bp = cg.build_pull(y, x, hyp, nugget=nug, y_err=ye)
        bp.compute_pull(svd_method=False)
        for i in (0, b // 2, b - 1):
            mo, vo = O.predict(y[i], x[i], hyp[i], nug[i], grid, ye[i], full_cov=False)
            assert_close(gp.Prediction[i], mo, 1e-9, 1e-12); assert_close(gp.prediction_variance[i], vo, 1e-9, 1e-13)
            po = O.loo_closed_form(y[i], x[i], hyp[i], nug[i], ye[i])
            assert_close(bp._pull[i], po[2], 1e-9, 1e-11)


@pytest.mark.gpu
def test_device_optimiser_equals_host_lockstep():
    """cgp_fit_objects_dev (simplices on the device) takes the same decisions as the numpy
    lock-step replay of scipy's Nelder-Mead: identical parameters, objective, iteration and
    evaluation counts, object by object -- 1D, 1D with a fitted nugget, ragged sizes across the
    one-warp and generic kernels, and 2D; including objects that never become positive definite."""
    import cosmogp_b200 as cg
    rng = np.random.default_rng(33)

    def both(gp, guess, nugget):
        gp.find_hyperparameters_per_object(hyperparameter_guess=guess, nugget=nugget, optimizer='device')
        dev = (gp.hyperparameters_per_object.copy(), gp.nugget_per_object.copy(), gp.log_likelihood_per_object.copy(),
               gp.fit_iterations.copy(), gp.fit_evaluations.copy())
        gp.find_hyperparameters_per_object(hyperparameter_guess=guess, nugget=nugget, optimizer='host')
        host = (gp.hyperparameters_per_object, gp.nugget_per_object, gp.log_likelihood_per_object,
                gp.fit_iterations, gp.fit_evaluations)
        for d, h in zip(dev, host):
            np.testing.assert_array_equal(d, h)
        return dev

    g = golden("notebook_with_noise")
    both(cg.gaussian_process_nobject(g["y"], g["x"], y_err=g["y_err"]), [0.5, 2], False)
    both(cg.gaussian_process_nobject(g["y"][:40], g["x"][:40], y_err=g["y_err"][:40]), [0.5, 2], True)

    sizes = rng.integers(3, 120, 300)                       # ragged, crossing N = 64
    x = [np.sort(rng.uniform(0, 30, n)) for n in sizes]
    y = [np.sin(t / 3.0) + 0.2 * rng.standard_normal(len(t)) for t in x]
    ye = [np.full(len(t), 0.2) for t in x]
    y[7] = np.full(len(y[7]), np.nan)                       # never positive definite: +inf everywhere
    d = both(cg.gaussian_process_nobject(y, x, y_err=ye), [1.0, 3.0], False)
    assert np.isinf(d[2][7]) and d[4][7] >= 400             # never finite: runs into maxfun like scipy would

    b, n = 64, 25                                           # 2D
    xy = [rng.uniform(0, 10, (n, 2)) for _ in range(b)]
    z = [np.sin(p[:, 0]) * np.cos(p[:, 1]) + 0.1 * rng.standard_normal(n) for p in xy]
    ze = [np.full(n, 0.1) for _ in range(b)]
    both(cg.gaussian_process_nobject(z, xy, kernel='RBF2D', y_err=ze), [1.0, 2.0, 2.0, 0.1], False)
    both(cg.gaussian_process_nobject(z[:16], xy[:16], kernel='RBF2D', y_err=ze[:16]), [1.0, 2.0, 2.0, 0.1], True)
