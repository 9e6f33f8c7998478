"""Per-object hyperparameter fits in lock step.

The reference fits one object at a time (`gaussian_process(y[i], x[i]).find_hyperparameters()`
in a Python loop, docs/notebook/1D_kernel_example_with_noise.ipynb cell 13;
cosmogp/Gaussian_process.py:216-253 -> scipy.optimize.fmin).  Here scipy's Nelder-Mead
(scipy/optimize/_optimize.py::_minimize_neldermead, non-adaptive, default tolerances) is
replayed for ALL objects at once on the host: every simplex operation is a vectorised numpy
step and every objective evaluation is ONE batched device launch in which each object uses
its own trial hyperparameters (cgp_ll_objhyp_dev).  Each object follows exactly the decisions
scipy would take for it; objects that have converged drop out of the launches.
"""
import numpy as np


def nelder_mead_lockstep(fun, x0, xatol=1e-4, fatol=1e-4, maxiter=None, maxfun=None):
    """Minimise B independent functions of n variables with scipy's Nelder-Mead rules.

    fun(X, idx) -> f: X is (len(idx), n), idx the object ids (int64); returns the objective of
    object idx[k] at X[k].  x0: (B, n).  Returns (x (B,n), fval (B,), iterations (B,), fcalls (B,)).
    Mirrors `scipy.optimize.fmin(f, x0, disp=False)` object by object, including the initial
    simplex (5 % steps, 0.00025 for zero coordinates), the comparison rules and the
    termination test (max|sim[1:]-sim[0]| <= xatol and max|fsim[0]-fsim[1:]| <= fatol)."""
    x0 = np.array(x0, dtype=np.float64)
    B, n = x0.shape
    rho, chi, psi, sigma = 1.0, 2.0, 0.5, 0.5
    nonzdelt, zdelt = 0.05, 0.00025
    maxiter = n * 200 if maxiter is None else maxiter
    maxfun = n * 200 if maxfun is None else maxfun
    allidx = np.arange(B, dtype=np.int64)

    sim = np.empty((B, n + 1, n))
    sim[:, 0] = x0
    for k in range(n):
        yk = x0.copy()
        yk[:, k] = np.where(yk[:, k] != 0, (1 + nonzdelt) * yk[:, k], zdelt)
        sim[:, k + 1] = yk
    fsim = np.empty((B, n + 1))
    for k in range(n + 1):
        fsim[:, k] = fun(sim[:, k], allidx)
    fcalls = np.full(B, n + 1, dtype=np.int64)
    order = np.argsort(fsim, axis=1, kind="stable")
    fsim = np.take_along_axis(fsim, order, axis=1)
    sim = np.take_along_axis(sim, order[:, :, None], axis=1)
    iterations = np.ones(B, dtype=np.int64)
    active = np.ones(B, dtype=bool)

    def evaluate(points, mask):
        """objective at points[mask] -> full-length array (NaN elsewhere); counts the calls."""
        out = np.full(B, np.nan)
        idx = allidx[mask]
        if len(idx):
            out[mask] = fun(points[mask], idx)
            fcalls[mask] += 1
        return out

    while True:
        active &= (fcalls < maxfun) & (iterations < maxiter)
        done = (np.max(np.abs(sim[:, 1:] - sim[:, :1]), axis=(1, 2)) <= xatol) & \
               (np.max(np.abs(fsim[:, :1] - fsim[:, 1:]), axis=1) <= fatol)
        active &= ~done
        if not active.any():
            break
        xbar = np.add.reduce(sim[:, :-1], axis=1) / n
        xr = (1 + rho) * xbar - rho * sim[:, -1]
        fxr = evaluate(xr, active)
        with np.errstate(invalid="ignore"):
            lt_best = active & (fxr < fsim[:, 0])
            mid = active & ~lt_best & (fxr < fsim[:, -2])
            rest = active & ~lt_best & ~mid
            out_c = rest & (fxr < fsim[:, -1])          # outside contraction
            in_c = rest & ~out_c                         # inside contraction
        # second evaluation: expansion / outside contraction / inside contraction, one launch
        x2 = np.where(lt_best[:, None], (1 + rho * chi) * xbar - rho * chi * sim[:, -1],
                      np.where(out_c[:, None], (1 + psi * rho) * xbar - psi * rho * sim[:, -1],
                               (1 - psi) * xbar + psi * sim[:, -1]))
        need2 = lt_best | out_c | in_c
        f2 = evaluate(x2, need2)
        with np.errstate(invalid="ignore"):
            take_e = lt_best & (f2 < fxr)
            take_r = (lt_best & ~take_e) | mid
            take_oc = out_c & (f2 <= fxr)
            take_ic = in_c & (f2 < fsim[:, -1])
            shrink = (out_c & ~take_oc) | (in_c & ~take_ic)
        take2 = take_e | take_oc | take_ic
        sim[take2, -1] = x2[take2]; fsim[take2, -1] = f2[take2]
        sim[take_r, -1] = xr[take_r]; fsim[take_r, -1] = fxr[take_r]
        if shrink.any():
            for j in range(1, n + 1):
                sim[shrink, j] = sim[shrink, 0] + sigma * (sim[shrink, j] - sim[shrink, 0])
                fj = evaluate(sim[:, j], shrink)
                fsim[shrink, j] = fj[shrink]
        order = np.argsort(fsim[active], axis=1, kind="stable")
        fsim[active] = np.take_along_axis(fsim[active], order, axis=1)
        sim[active] = np.take_along_axis(sim[active], order[:, :, None], axis=1)
        iterations[active] += 1
    return sim[:, 0].copy(), fsim[:, 0].copy(), iterations, fcalls
