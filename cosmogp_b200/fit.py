"""Per-object hyperparameter fits in lock step (host statement; the product path is
cgp_fit_objects_dev, which keeps the simplices on the device and is tested to give identical results).

The reference fits one object at a time (`gaussian_process(y[i], x[i]).find_hyperparameters()`
in a Python loop, docs/notebook/1D_kernel_example_with_noise.ipynb cell 13;
cosmogp/Gaussian_process.py:216-253 -> scipy.optimize.fmin).  Here scipy's Nelder-Mead
(scipy/optimize/_optimize.py::_minimize_neldermead, non-adaptive, default tolerances) is
replayed for ALL objects at once on the host: every simplex operation is a vectorised numpy
step and every objective evaluation is ONE batched device launch in which each object uses
its own trial hyperparameters (cgp_ll_objhyp_dev).  Each object follows exactly the decisions
scipy would take for it; converged objects are dropped from the working set, so the host
arithmetic and the launches shrink as the batch converges.
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

    x_out = np.empty((B, n)); f_out = np.empty(B)
    it_out = np.empty(B, dtype=np.int64); fc_out = np.empty(B, dtype=np.int64)

    # working set (compact): ids, simplex S (k, n+1, n), values F (k, n+1), counters
    ids = np.arange(B, dtype=np.int64)
    S = np.empty((B, n + 1, n))
    S[:, 0] = x0
    for k in range(n):
        yk = x0.copy()
        yk[:, k] = np.where(yk[:, k] != 0, (1 + nonzdelt) * yk[:, k], zdelt)
        S[:, k + 1] = yk
    F = np.empty((B, n + 1))
    for k in range(n + 1):
        F[:, k] = fun(S[:, k], ids)
    fc = np.full(B, n + 1, dtype=np.int64)
    it = np.ones(B, dtype=np.int64)

    def sort_rows(S, F):
        order = np.argsort(F, axis=1, kind="stable")
        return np.take_along_axis(S, order[:, :, None], axis=1), np.take_along_axis(F, order, axis=1)

    S, F = sort_rows(S, F)

    def evaluate(points, mask):
        """objective at points[mask] (NaN elsewhere); counts the calls."""
        out = np.full(len(ids), np.nan)
        if mask.any():
            out[mask] = fun(points[mask], ids[mask])
            fc[mask] += 1
        return out

    while len(ids):
        k = len(ids)
        with np.errstate(invalid="ignore"):                  # inf - inf -> NaN -> never "converged", like scipy
            spread = np.abs(S[:, 1:] - S[:, :1]).reshape(k, -1).max(axis=1)
            fspread = np.abs(F[:, :1] - F[:, 1:]).max(axis=1)
        keep = (fc < maxfun) & (it < maxiter) & ~((spread <= xatol) & (fspread <= fatol))
        if not keep.all():
            fin = ~keep
            x_out[ids[fin]] = S[fin, 0]; f_out[ids[fin]] = F[fin, 0]
            it_out[ids[fin]] = it[fin]; fc_out[ids[fin]] = fc[fin]
            ids, S, F, fc, it = ids[keep], S[keep], F[keep], fc[keep], it[keep]
            if not len(ids):
                break
        everyone = np.ones(len(ids), dtype=bool)
        worst = S[:, -1]
        xbar = np.add.reduce(S[:, :-1], axis=1) / n
        xr = (1 + rho) * xbar - rho * worst
        fxr = evaluate(xr, everyone)
        with np.errstate(invalid="ignore"):
            lt_best = fxr < F[:, 0]
            mid = ~lt_best & (fxr < F[:, -2])
            rest = ~lt_best & ~mid
            out_c = rest & (fxr < F[:, -1])              # outside contraction
            in_c = rest & ~out_c                          # inside contraction
        # second evaluation: expansion / outside contraction / inside contraction, one launch
        x2 = np.where(lt_best[:, None], (1 + rho * chi) * xbar - rho * chi * worst,
                      np.where(out_c[:, None], (1 + psi * rho) * xbar - psi * rho * worst,
                               (1 - psi) * xbar + psi * worst))
        f2 = evaluate(x2, lt_best | out_c | in_c)
        with np.errstate(invalid="ignore"):
            take_e = lt_best & (f2 < fxr)
            take_r = (lt_best & ~take_e) | mid
            take_oc = out_c & (f2 <= fxr)
            take_ic = in_c & (f2 < F[:, -1])
            shrink = (out_c & ~take_oc) | (in_c & ~take_ic)
        take2 = take_e | take_oc | take_ic
        S[take2, -1] = x2[take2]; F[take2, -1] = f2[take2]
        S[take_r, -1] = xr[take_r]; F[take_r, -1] = fxr[take_r]
        if shrink.any():
            for j in range(1, n + 1):
                S[shrink, j] = S[shrink, 0] + sigma * (S[shrink, j] - S[shrink, 0])
                fj = evaluate(S[:, j], shrink)
                F[shrink, j] = fj[shrink]
        S, F = sort_rows(S, F)
        it += 1
    return x_out, f_out, it_out, fc_out
