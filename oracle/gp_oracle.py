"""CPU oracle for the cosmogp GP hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A numpy/scipy restatement of the algorithm of PFLeget/cosmogp (reference tree
/root/reference, citations are file:line into it).  Only tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke() may import it;
the shipped package cosmogp_b200 never does.

Parity status: PINNED.  tests/test_oracle_golden.py checks every function here
against tests/golden/*.json, fixtures produced by running the real reference in
the build container (tests/golden/make_golden.py, via oracle/ref_loader.py), and
against the five notebook numbers the reference reproduces bit for bit
 -- the only known-answer values the reference has: its own
tests/ are `print 'to do'` stubs.

Two layers:
  * per-object functions that follow the reference call for call (same scipy
    routines, same temporaries) -- this is what the cpu_baseline times;
  * `*_batched` helpers (numpy stacked linear algebra) for parity checks at
    10^5..10^6 objects, themselves tested against the per-object layer.
"""
import numpy as np
from scipy import linalg as sla

LOG_2PI = np.log(2.0 * np.pi)


# --------------------------------------------------------------------------- kernels
def noise_diag(n, y_err, nugget, floor):
    """y_err^2 + floor^2 + nugget^2 per point (kernel.py:74-75, :150-151)."""
    ye = np.zeros(n) if y_err is None else np.asarray(y_err, dtype=float)
    return ye * ye + floor ** 2 + nugget ** 2


def rbf_1d(x, hyp, new_x=None, nugget=0.0, floor=0.0, y_err=None):
    """kernel.py:25-77.  Auto-covariance (N,N) with noise diagonal when new_x is
    None, else cross-covariance of shape (len(new_x), len(x)) without it."""
    x = np.asarray(x, dtype=float)
    rows = x if new_x is None else np.asarray(new_x, dtype=float)
    delta = x[None, :] - rows[:, None]                      # kernel.py:71
    cov = hyp[0] ** 2 * np.exp(-0.5 * (delta * delta / hyp[1] ** 2))  # kernel.py:72
    if new_x is None:
        idx = np.arange(len(x))
        cov[idx, idx] += noise_diag(len(x), y_err, nugget, floor)
    return cov


def metric_2d(hyp):
    """Inverse metric (kernel.py:127-130): [[ly^2,-lxy],[-lxy,lx^2]]/(lx^2 ly^2-lxy^2)."""
    lx2, ly2, lxy = hyp[1] ** 2, hyp[2] ** 2, hyp[3]
    scale = 1.0 / (lx2 * ly2 - lxy ** 2)
    return ly2 * scale, -lxy * scale, lx2 * scale          # m00, m01, m11


def rbf_2d(x, hyp, new_x=None, nugget=0.0, floor=0.0, y_err=None, amp_on_autocov=False):
    """kernel.py:80-155 at HEAD.  Cross branch (:137-142): sigma^2 exp(-d^2/2),
    shape (M,N).  Auto branch (:143-151): exp(-d^2/2) with unit diagonal and NO
    sigma^2 plus the noise diagonal.  d is the Mahalanobis
    distance under the inverse metric, written out as the quadratic form."""
    x = np.asarray(x, dtype=float)
    m00, m01, m11 = metric_2d(hyp)
    rows = x if new_x is None else np.asarray(new_x, dtype=float)
    dx = x[None, :, 0] - rows[:, None, 0]
    dy = x[None, :, 1] - rows[:, None, 1]
    q = dx * dx * m00 + 2.0 * dx * dy * m01 + dy * dy * m11
    cov = np.exp(-0.5 * q)
    if new_x is not None:
        return hyp[0] ** 2 * cov
    if amp_on_autocov:
        cov *= hyp[0] ** 2
    idx = np.arange(len(x))
    cov[idx, idx] = (hyp[0] ** 2 if amp_on_autocov else 1.0)
    cov[idx, idx] += noise_diag(len(x), y_err, nugget, floor)
    return cov


def init_rbf(x, y):
    """kernel.py:6-22 including quirk Q6: the loop overwrites L_min/L_max/sigma, so
    only the LAST object's extent and scatter survive; number_point is per object."""
    npts = np.array([len(xi) for xi in x], dtype=float)
    lo, hi, sig = np.min(x[-1]), np.max(x[-1]), np.std(y[-1])
    d = np.mean(np.sqrt((hi - lo) ** 2 / npts))
    return float(sig), float(np.mean([d, hi - lo]))


# --------------------------------------------------------------------------- inverses
def cholesky_inverse(matrix, return_logdet=False):
    """inv_matrix.py:21-31: potrf, general inverse of L, inv(L)^T inv(L);
    logdet = sum 2 log L_ii.  Raises numpy.linalg.LinAlgError when not PD."""
    low = sla.cholesky(matrix, lower=True)
    low_inv = sla.inv(low)
    inv = low_inv.T @ low_inv
    if return_logdet:
        return inv, float(np.sum(2.0 * np.log(np.diag(low))))
    return inv


def svd_inverse(matrix, return_logdet=False):
    """inv_matrix.py:4-18: pseudo-inverse keeping singular values > 1e-15
    (absolute); logdet sums the kept ones only."""
    u, s, vt = sla.svd(matrix)
    keep = s > 1e-15
    inv = vt.T[:, keep] @ (np.diag(1.0 / s[keep]) @ u.T[keep])
    if return_logdet:
        return inv, float(np.sum(np.log(s[keep])))
    return inv


# --------------------------------------------------------------------------- log-likelihood
def log_likelihood(y, x, hyp, nugget, y_err=None, y_mean=None, kind="1d", svd_method=False,
                   amp_on_autocov=False):
    """Gaussian_process.py:13-75 for one object -> python float
    (the reference returns a shape-(1,) array; the value is the same)."""
    y = np.asarray(y, dtype=float)
    r = y - (0.0 if y_mean is None else y_mean)
    if kind == "1d":
        k = rbf_1d(x, hyp, nugget=nugget, y_err=y_err)
    else:
        k = rbf_2d(x, hyp, nugget=nugget, y_err=y_err, amp_on_autocov=amp_on_autocov)
    inv, logdet = (svd_inverse if svd_method else cholesky_inverse)(k, return_logdet=True)
    return float(-0.5 * r @ (inv @ r) - 0.5 * len(y) * LOG_2PI - 0.5 * logdet)


def log_likelihood_sum(ys, xs, hyp, nugget, y_errs=None, y0s=None, kind="1d", svd_method=False):
    """Gaussian_process.py:191-213: plain left-to-right sum over objects."""
    total = 0.0
    for i in range(len(ys)):
        total += log_likelihood(ys[i], xs[i], hyp, nugget,
                                None if y_errs is None else y_errs[i],
                                None if y0s is None else y0s[i], kind, svd_method)
    return total


# --------------------------------------------------------------------------- prediction
def predict(y, x, hyp, nugget, grid, y_err=None, y0=0.0, new_y0=0.0, kind="1d",
            svd_method=False, full_cov=True):
    """Gaussian_process.py:270-361 for one object with an explicit grid.
    mean = H K^-1 (y-y0) + new_y0 (:332-335);
    cov  = K(grid,grid)+nugget^2 I - H K^-1 H^T (:356-361), no y_err on K** (Q13).
    Returns (mean, cov) or (mean, diag(cov)) when full_cov is False."""
    kern = rbf_1d if kind == "1d" else rbf_2d
    k = kern(x, hyp, nugget=nugget, y_err=y_err)
    inv = (svd_inverse if svd_method else cholesky_inverse)(k)
    h = kern(x, hyp, new_x=grid)                            # (M, N)
    mean = h @ (inv @ (np.asarray(y, dtype=float) - y0)) + new_y0
    if full_cov:
        return mean, kern(grid, hyp, nugget=nugget) - h @ (inv @ h.T)
    amp2 = hyp[0] ** 2 if kind == "1d" else 1.0             # 2D auto branch has no sigma^2 (Q2)
    return mean, amp2 + nugget ** 2 - ((h @ inv) * h).sum(axis=1)   # diag(H K^-1 H^T) through BLAS


def predict_grid_chunked(y, x, hyp, nugget, grid, y_err=None, y0=0.0, new_y0=0.0, kind="1d", chunk=2000):
    """mean and diag(covariance) on a grid too long for the M x M covariance (BASELINE config 3: M = 10^5 would
    need 80 GB): the grid goes through Gaussian_process.py:332-335 / :356-361 in slices of `chunk` points with ONE
    inverse (inv_matrix.py:21-31) -- the chunking the reference's own DES notebook applies (cell 9)."""
    kern = rbf_1d if kind == "1d" else rbf_2d
    inv = cholesky_inverse(kern(x, hyp, nugget=nugget, y_err=y_err))
    w = inv @ (np.asarray(y, dtype=float) - y0)
    grid = np.asarray(grid, dtype=float)
    m = len(grid)
    ny0 = np.broadcast_to(np.asarray(new_y0, dtype=float), (m,))
    amp2 = hyp[0] ** 2 if kind == "1d" else 1.0
    mean, var = np.empty(m), np.empty(m)
    for s in range(0, m, chunk):
        h = kern(x, hyp, new_x=grid[s:s + chunk])
        mean[s:s + chunk] = h @ w + ny0[s:s + chunk]
        var[s:s + chunk] = amp2 + nugget ** 2 - ((h @ inv) * h).sum(axis=1)
    return mean, var


# --------------------------------------------------------------------------- leave-one-out pulls
def loo_bruteforce(y, x, hyp, nugget, y_err, kind="1d", svd_method=False):
    """pull.py:66-94, mode A (no mean, no diff): N refits on N-1 points each,
    prediction and |variance| at the left-out point.  O(N^4); small N only."""
    y = np.asarray(y, dtype=float)
    n = len(y)
    pred, var = np.zeros(n), np.zeros(n)
    for t in range(n):
        keep = np.arange(n) != t
        m, c = predict(y[keep], np.asarray(x)[keep], hyp, nugget, np.asarray(x),
                       y_err=np.asarray(y_err)[keep], kind=kind, svd_method=svd_method)
        pred[t], var[t] = m[t], abs(c[t, t])
    resid = pred - y
    return pred, var, resid / np.sqrt(np.asarray(y_err) ** 2 + var + nugget ** 2), resid


def loo_closed_form(y, x, hyp, nugget, y_err, mean=None, diff=None, recenter=False, kind="1d"):
    """Closed form of pull.py:43-102.  With K including
    y_err^2+nugget^2, d = diag(K^-1), loo(v) = v - (K^-1 v)/d:
      mode A  mean None, recenter False : pred = loo(y)
      mode C  diff given                : pred = m + diff + loo(y - m - diff)
      mode B/D mean given, diff None (or recenter with a frozen mean m):
              r = y-m, delta_t = (sum r - r_t)/(N-1),
              pred = m + delta + loo(r) - delta*loo(1)
    pred_var = |1/d - y_err^2|  (K** carries nugget^2 but not y_err^2,
    Gaussian_process.py:357; abs at pull.py:90); pull denominator counts nugget^2
    twice (pull.py:92-93).
    2D at HEAD: the cross-covariance carries sigma^2 but the auto-covariance does not
    (Q2), so with rho = sigma^2 every loo(.) term is scaled by rho and
    pred_var = |1 + nugget^2 - rho^2 (K_tt - 1/d)|; rho = 1 gives the 1D formula."""
    y = np.asarray(y, dtype=float)
    ye = np.asarray(y_err, dtype=float)
    n = len(y)
    kern = rbf_1d if kind == "1d" else rbf_2d
    kmat = kern(x, hyp, nugget=nugget, y_err=ye)
    inv = cholesky_inverse(kmat)
    d = np.diag(inv)
    rho = 1.0 if kind == "1d" else hyp[0] ** 2
    amp_auto = hyp[0] ** 2 if kind == "1d" else 1.0

    def loo(v):
        return rho * (v - (inv @ v) / d)

    if mean is None and not recenter:
        pred = loo(y)
    elif diff is not None:
        m = (0.0 if mean is None else mean) + diff
        pred = m + loo(y - m)
    else:
        m = np.zeros(n) if mean is None else np.asarray(mean, dtype=float)
        r = y - m
        delta = (r.sum() - r) / (n - 1)
        pred = m + delta + loo(r) - delta * loo(np.ones(n))
    var = np.abs(amp_auto + nugget ** 2 - rho ** 2 * (np.diag(kmat) - 1.0 / d))
    resid = pred - y
    return pred, var, resid / np.sqrt(ye * ye + var + nugget ** 2), resid


def norm_fit(values):
    """scipy.stats.norm.fit (pull.py:102) = sample mean and population std."""
    v = np.asarray(values, dtype=float)
    mu = v.mean()
    return float(mu), float(np.sqrt(np.mean((v - mu) ** 2)))


# --------------------------------------------------------------------------- mean function
def return_mean_1d(y, x, mean_y=None, mean_x=None, diff=None, new_x=None):
    """mean.py:70-104, 1D branch: cubic InterpolatedUnivariateSpline of the template
    (mean.py:28-31) at x (and new_x), plus diff; diff None -> mean(y - template)."""
    from scipy.interpolate import InterpolatedUnivariateSpline
    shape = 0.0
    spline = None
    if mean_y is not None:
        spline = InterpolatedUnivariateSpline(mean_x, mean_y)
        shape = spline(x)
    if diff is None:
        diff = np.mean(y - shape)
    y0 = shape + diff
    if new_x is None:
        return y0
    return spline(new_x) + diff if spline is not None else y0


# --------------------------------------------------------------------------- batched helpers
def ll_batched_1d(x, y, y0, y_err, hyp, nugget, chunk=20000):
    """Per-object LL for equal-length objects, x/y/y0/y_err of shape (B,N), by
    stacked Cholesky: LL = -1/2 |L^-1 r|^2 - sum log L_ii - N/2 log 2pi."""
    b, n = x.shape
    out = np.empty(b)
    eye = np.eye(n, dtype=bool)
    for s in range(0, b, chunk):
        xs = x[s:s + chunk]
        d = xs[:, None, :] - xs[:, :, None]
        k = hyp[0] ** 2 * np.exp(-0.5 * (d * d / hyp[1] ** 2))
        k[:, eye] += y_err[s:s + chunk] ** 2 + nugget ** 2
        low = np.linalg.cholesky(k)
        r = (y[s:s + chunk] - y0[s:s + chunk])[:, :, None]
        alpha = np.linalg.solve(k, r)[:, :, 0]
        logdet = 2.0 * np.log(np.diagonal(low, axis1=1, axis2=2)).sum(axis=1)
        out[s:s + chunk] = -0.5 * (r[:, :, 0] * alpha).sum(axis=1) - 0.5 * n * LOG_2PI - 0.5 * logdet
    return out


def predict_batched_1d(x, y, y0, y_err, hyp, nugget, grid, new_y0, chunk=10000):
    """Mean and variance diagonal on a shared grid for equal-length objects."""
    b, n = x.shape
    m = len(grid)
    mean, var = np.empty((b, m)), np.empty((b, m))
    eye = np.eye(n, dtype=bool)
    for s in range(0, b, chunk):
        xs = x[s:s + chunk]
        d = xs[:, None, :] - xs[:, :, None]
        k = hyp[0] ** 2 * np.exp(-0.5 * (d * d / hyp[1] ** 2))
        k[:, eye] += y_err[s:s + chunk] ** 2 + nugget ** 2
        dg = xs[:, None, :] - grid[None, :, None]
        h = hyp[0] ** 2 * np.exp(-0.5 * (dg * dg / hyp[1] ** 2))        # (b, M, N)
        sol = np.linalg.solve(k, np.concatenate(
            [(y[s:s + chunk] - y0[s:s + chunk])[:, :, None], h.transpose(0, 2, 1)], axis=2))
        mean[s:s + chunk] = np.einsum("bmn,bn->bm", h, sol[:, :, 0]) + new_y0[s:s + chunk]
        var[s:s + chunk] = hyp[0] ** 2 + nugget ** 2 - np.einsum("bmn,bnm->bm", h, sol[:, :, 1:])
    return mean, var


def loo_batched_1d(x, y, y_err, hyp, nugget, chunk=20000):
    """Mode-A closed-form LOO for equal-length objects -> pred, var, pull, resid (B,N)."""
    b, n = x.shape
    pred, var = np.empty((b, n)), np.empty((b, n))
    eye = np.eye(n, dtype=bool)
    for s in range(0, b, chunk):
        xs = x[s:s + chunk]
        d = xs[:, None, :] - xs[:, :, None]
        k = hyp[0] ** 2 * np.exp(-0.5 * (d * d / hyp[1] ** 2))
        k[:, eye] += y_err[s:s + chunk] ** 2 + nugget ** 2
        inv = np.linalg.inv(k)
        dg = np.diagonal(inv, axis1=1, axis2=2)
        ys = y[s:s + chunk]
        pred[s:s + chunk] = ys - np.einsum("bij,bj->bi", inv, ys) / dg
        var[s:s + chunk] = np.abs(1.0 / dg - y_err[s:s + chunk] ** 2)
    resid = pred - y
    return pred, var, resid / np.sqrt(y_err ** 2 + var + nugget ** 2), resid
