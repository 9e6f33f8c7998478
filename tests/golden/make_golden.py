"""Generate tests/golden/*.npz by running the REAL reference (PFLeget/cosmogp at
/root/reference, loaded through oracle/ref_loader.py) on seeded inputs.

Run in the build container only:  python tests/golden/make_golden.py
Every fixture stores inputs AND reference outputs, so the tests never need the
reference tree or a particular RNG stream.  Ragged lists are stored flattened
with an `off` (CSR offsets) array.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ref = ref_loader.load()


def flat(lst):
    off = np.zeros(len(lst) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(a) for a in lst])
    return np.concatenate([np.asarray(a, dtype=float).reshape(len(a), -1) for a in lst]).squeeze(), off


def save(name, **kw):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **kw)
    print("wrote", name, {k: np.shape(v) for k, v in kw.items()})


def sample_gp(rng, k):
    return np.linalg.cholesky(k) @ rng.standard_normal(len(k))


def ll_of(gp, hyp, nugget, svd_method):
    gp.nugget = nugget
    gp.fit_nugget = False
    gp.compute_log_likelihood(hyp, svd_method=svd_method)
    return float(np.asarray(gp.log_likelihood).ravel()[0])


def per_object_ll(ys, xs, kernel, hyp, nugget, yerrs, y0s, svd_method=False):
    with ref_loader.quiet():
        return np.array([float(np.asarray(sys.modules["cosmogp.Gaussian_process"].log_likelihood_gp(
            ys[i], xs[i], kernel, hyp, nugget, y_err=yerrs[i], y_mean=y0s[i],
            svd_method=svd_method)).ravel()[0]) for i in range(len(ys))])


# ----------------------------------------------------------------- 1. SURVEY 9.3 1D vector
def kat_1d():
    x = np.linspace(0, 5, 8); y = np.sin(x); ye = 0.1 * np.ones(8)
    hyp = np.array([1.0, 1.0]); nug = 0.05; grid = np.linspace(0, 5, 5)
    gp = ref.gaussian_process(y, x, y_err=ye)
    init = np.array(gp.hyperparameters, dtype=float)
    gp.hyperparameters = hyp
    ll_c = ll_of(gp, hyp, nug, False); ll_s = ll_of(gp, hyp, nug, True)
    gp.get_prediction(new_binning=grid, svd_method=False)
    bp = ref.build_pull([y], [x], hyp, nugget=nug, y_err=[ye]); bp.compute_pull(svd_method=False)
    save("kat_1d", x=x, y=y, y_err=ye, hyp=hyp, nugget=nug, grid=grid, init_rbf=init,
         ll_chol=ll_c, ll_svd=ll_s, kmat=gp.kernel_matrix[0], kinv=gp.inv_kernel_matrix[0],
         mean=gp.Prediction[0], cov=gp.covariance_matrix[0],
         pull=np.array(bp.pull), resid=np.array(bp.residual), pred=bp.prediction[0],
         pull_average=bp.pull_average, pull_std=bp.pull_std)


# ----------------------------------------------------------------- 2. SURVEY 9.3 2D vector
def kat_2d():
    a = np.linspace(0, 5, 3)
    gx, gy = np.meshgrid(a, a)
    c = np.array([gx.ravel(), gy.ravel()]).T
    hyp = np.array([2.0, np.sqrt(5.0), 2.0, 0.1]); ye = 0.3 * np.ones(9); nug = 0.1
    newx = np.array([[1.0, 2.0], [3.0, 0.5]])
    y = np.cos(c[:, 0]) + 0.5 * c[:, 1]
    with ref_loader.quiet():
        k = ref.rbf_kernel_2d(c, hyp, nugget=nug, y_err=ye)
        h = ref.rbf_kernel_2d(c, hyp, new_x=newx)
        ll = float(np.asarray(sys.modules["cosmogp.Gaussian_process"].log_likelihood_gp(y, c, ref.rbf_kernel_2d, hyp, nug, y_err=ye,
                                                    svd_method=False)).ravel()[0])
        gp = ref.gaussian_process(y, c, kernel="RBF2D", y_err=ye)
        init = np.array(gp.hyperparameters, dtype=float)
        gp.hyperparameters = hyp; gp.nugget = nug
        gp.get_prediction(new_binning=newx, svd_method=False)
        bp = ref.build_pull([y], [c], hyp, nugget=nug, y_err=[ye], kernel="RBF2D")
        bp.compute_pull(svd_method=False)
    save("kat_2d", x=c, y=y, y_err=ye, hyp=hyp, nugget=nug, grid=newx, kmat=k, hmat=h, ll_chol=ll,
         init_rbf=init, mean=gp.Prediction[0], cov=gp.covariance_matrix[0],
         pull=np.array(bp.pull), resid=np.array(bp.residual), pred=bp.prediction[0])


# ----------------------------------------------------------------- 3. config C1 (single light curve)
def c1_single():
    rng = np.random.default_rng(1)
    n, m = 50, 500
    x = np.sort(rng.uniform(-12, 42, n)); hyp = np.array([0.5, 8.0]); nug = 0.03
    ye = rng.uniform(0.03, 0.1, n)
    y = sample_gp(rng, ref.rbf_kernel_1d(x, hyp, nugget=nug, y_err=ye))
    grid = np.linspace(-12, 42, m)
    gp = ref.gaussian_process(y, x, y_err=ye)
    gp.hyperparameters = hyp.copy()
    ll_c = ll_of(gp, hyp, nug, False); ll_s = ll_of(gp, hyp, nug, True)
    gp.get_prediction(new_binning=grid, svd_method=False)
    mean, cov = gp.Prediction[0].copy(), gp.covariance_matrix[0].copy()
    bp = ref.build_pull([y], [x], hyp, nugget=nug, y_err=[ye]); bp.compute_pull(svd_method=False)
    gpf = ref.gaussian_process(y, x, y_err=ye)
    gpf.find_hyperparameters(hyperparameter_guess=[0.5, 8.0], svd_method=False)
    fit = np.array(gpf.hyperparameters, dtype=float)
    gpn = ref.gaussian_process(y, x, y_err=ye)
    gpn.find_hyperparameters(hyperparameter_guess=[0.5, 8.0], nugget=True, svd_method=False)
    save("c1_single", x=x, y=y, y_err=ye, hyp=hyp, nugget=nug, grid=grid, ll_chol=ll_c, ll_svd=ll_s,
         mean=mean, cov_diag=np.diag(cov).copy(), cov_block=cov[100:140, 100:140].copy(),
         cov_sum=cov.sum(), pull=np.array(bp.pull), resid=np.array(bp.residual), pred=bp.prediction[0],
         pull_average=bp.pull_average, pull_std=bp.pull_std, fit_hyp=fit,
         fit_hyp_nugget=np.array(gpn.hyperparameters, dtype=float), fit_nugget=float(gpn.nugget))


# ----------------------------------------------------------------- 4. ragged 1D batch with shared mean (C2 in miniature)
def ragged_1d():
    rng = np.random.default_rng(2)
    b = 12
    hyp = np.array([0.5, 2.0]); nug = 0.0
    tmean = np.linspace(-15, 45, 61); ymean = -18 + 2 * np.sin(tmean / 10)
    from scipy.interpolate import InterpolatedUnivariateSpline
    xs, ys, yes = [], [], []
    for i in range(b):
        n = int(rng.integers(5, 70)) if i else 60
        x = np.sort(rng.uniform(-10, 40, n)); ye = 0.2 * np.ones(n) if i % 2 == 0 else rng.uniform(0.1, 0.3, n)
        y = (InterpolatedUnivariateSpline(tmean, ymean)(x) + rng.normal(0, 0.3)
             + sample_gp(rng, ref.rbf_kernel_1d(x, hyp, y_err=ye)))
        xs.append(x); ys.append(y); yes.append(ye)
    grid = np.linspace(-10, 40, 100)
    gp = ref.gaussian_process_nobject(ys, xs, y_err=yes, Mean_Y=ymean, Time_mean=tmean)
    y0 = [np.asarray(v) for v in gp.y0]
    init = np.array(gp.hyperparameters, dtype=float)
    gp.hyperparameters = hyp.copy()
    ll_sum = ll_of(gp, hyp, nug, False)
    ll_sum_nug = ll_of(gp, hyp, 0.07, False)
    ll_obj = per_object_ll(ys, xs, ref.rbf_kernel_1d, hyp, nug, yes, y0)
    gp.nugget = nug
    gp.get_prediction(new_binning=grid, svd_method=False)
    mean = np.array(gp.Prediction); var = np.array([np.diag(c) for c in gp.covariance_matrix])
    # own-epoch prediction (new_binning=None) for object 0 alone: avoids stale-index quirk Q3
    g1 = ref.gaussian_process(ys[0], xs[0], y_err=yes[0], Mean_Y=ymean, Time_mean=tmean)
    g1.hyperparameters = hyp.copy(); g1.nugget = 0.05
    g1.get_prediction(svd_method=False)
    gf = ref.gaussian_process_nobject(ys, xs, y_err=yes, Mean_Y=ymean, Time_mean=tmean)
    gf.find_hyperparameters(hyperparameter_guess=[0.5, 2.0], svd_method=False)
    # pulls, mode B (mean given, diff None) and mode C (diff given)
    bpb = ref.build_pull(ys, xs, hyp, nugget=0.05, y_err=yes, y_mean=ymean, x_axis_mean=tmean)
    bpb.compute_pull(svd_method=False)
    diff = [float(np.mean(ys[i]) + 18.0) for i in range(b)]
    bpc = ref.build_pull(ys, xs, hyp, nugget=0.05, y_err=yes, y_mean=ymean, x_axis_mean=tmean)
    bpc.compute_pull(diff=diff, svd_method=False)
    xf, off = flat(xs); yf, _ = flat(ys); yef, _ = flat(yes); y0f, _ = flat(y0)
    save("ragged_1d", x=xf, y=yf, y_err=yef, off=off, y0=y0f, hyp=hyp, nugget=nug, grid=grid,
         mean_x=tmean, mean_y=ymean, init_rbf=init, ll_sum=ll_sum, ll_sum_nugget007=ll_sum_nug,
         ll_obj=ll_obj, mean=mean, var=var, cov0=gp.covariance_matrix[0], kmat0=gp.kernel_matrix[0],
         kinv0=gp.inv_kernel_matrix[0], own_mean0=g1.Prediction[0], own_cov0=g1.covariance_matrix[0],
         fit_hyp=np.array(gf.hyperparameters, dtype=float),
         pullB=np.array(bpb.pull), residB=np.array(bpb.residual), predB=np.concatenate(bpb.prediction),
         pullB_avg=bpb.pull_average, pullB_std=bpb.pull_std,
         diff=np.array(diff), pullC=np.array(bpc.pull), residC=np.array(bpc.residual),
         predC=np.concatenate(bpc.prediction))


# ----------------------------------------------------------------- 5. pulls, modes A and D, C5 recipe in miniature
def pulls_1d():
    rng = np.random.default_rng(5)
    b, n = 6, 40
    hyp = np.array([0.5, 2.0]); nug = 0.02
    xs, ys, yes = [], [], []
    for i in range(b):
        x = np.sort(rng.uniform(-10, 10, n)); ye = 0.1 * np.ones(n)
        y = 0.5 * np.sin(x / 2.0 + rng.uniform(0, 6.28)) + ye * rng.standard_normal(n)
        xs.append(x); ys.append(y); yes.append(ye)
    bpa = ref.build_pull(ys, xs, hyp, nugget=nug, y_err=yes); bpa.compute_pull(svd_method=False)
    bps = ref.build_pull(ys, xs, hyp, nugget=nug, y_err=yes); bps.compute_pull(svd_method=True)
    # mode D: substract_mean=True without a mean; sticky mean from object 0 (Q8) needs equal x
    # support for the spline, so use one common epoch grid.
    xc = np.sort(rng.uniform(-10, 10, n))
    yd = [1.5 + 0.5 * np.sin(xc / 2.0 + rng.uniform(0, 6.28)) + 0.1 * rng.standard_normal(n) for _ in range(3)]
    bpd = ref.build_pull(yd, [xc] * 3, hyp, nugget=nug, y_err=yes[:3])
    bpd.compute_pull(svd_method=False, substract_mean=True)
    save("pulls_1d", x=np.array(xs), y=np.array(ys), y_err=np.array(yes), hyp=hyp, nugget=nug,
         pullA=np.array(bpa.pull), residA=np.array(bpa.residual), predA=np.array(bpa.prediction),
         pullA_avg=bpa.pull_average, pullA_std=bpa.pull_std, pullA_svd=np.array(bps.pull),
         xD=xc, yD=np.array(yd), pullD=np.array(bpd.pull), residD=np.array(bpd.residual),
         predD=np.array(bpd.prediction))


# ----------------------------------------------------------------- 6. 2D batch (C3 in miniature)
def batch_2d():
    rng = np.random.default_rng(3)
    hyp = np.array([1.0, 30.0, 25.0, 50.0]); nug = 0.05
    xs, ys, yes = [], [], []
    for n in (45, 64, 30):
        x = rng.uniform(-200, 200, (n, 2)); ye = rng.uniform(0.15, 0.25, n)
        y = np.cos(x[:, 0] / 60.0) * np.sin(x[:, 1] / 45.0) + ye * rng.standard_normal(n)
        xs.append(x); ys.append(y); yes.append(ye)
    grid = rng.uniform(-200, 200, (80, 2))
    with ref_loader.quiet():
        gp = ref.gaussian_process_nobject(ys, xs, kernel="RBF2D", y_err=yes)
        init = np.array(gp.hyperparameters, dtype=float)
        gp.hyperparameters = hyp.copy()
        ll_sum = ll_of(gp, hyp, nug, False)
        ll_obj = per_object_ll(ys, xs, ref.rbf_kernel_2d, hyp, nug, yes, [None] * 3)
        gp.nugget = nug
        gp.get_prediction(new_binning=grid, svd_method=False)
        bp = ref.build_pull(ys[2:], xs[2:], hyp, nugget=nug, y_err=yes[2:], kernel="RBF2D")
        bp.compute_pull(svd_method=False)
    xf = np.concatenate(xs); off = np.cumsum([0] + [len(v) for v in ys])
    save("batch_2d", x=xf, y=np.concatenate(ys), y_err=np.concatenate(yes), off=off, hyp=hyp, nugget=nug,
         grid=grid, init_rbf=init, ll_sum=ll_sum, ll_obj=ll_obj, mean=np.array(gp.Prediction),
         var=np.array([np.diag(c) for c in gp.covariance_matrix]), cov2=gp.covariance_matrix[2],
         kmat2=gp.kernel_matrix[2], pull2=np.array(bp.pull), resid2=np.array(bp.residual), pred2=bp.prediction[0])


# ----------------------------------------------------------------- 7. notebook known answers (legacy RNG data stored)
def notebooks():
    # docs/notebook/1D_kernel_example_with_noise.ipynb cells 1,3,7,9,19,21 re-expressed with HEAD keywords
    np.random.seed(1)
    a = 0.2
    grids, ys, yes = [], [], []
    for _ in range(100):
        n_point = int(np.random.uniform(60, 60))
        grid = np.linspace(-10, 40, n_point)
        k = ref.rbf_kernel_1d(grid, np.array([0.5, 2]), nugget=0)
        ys.append(np.random.multivariate_normal(np.zeros_like(grid), k + a * a * np.eye(len(k))))
        grids.append(grid); yes.append(a * np.ones(len(k)))
    gp = ref.gaussian_process(ys[0], grids[0], y_err=yes[0])
    gp.find_hyperparameters(hyperparameter_guess=[0.5, 2], svd_method=False)
    single = np.array(gp.hyperparameters, dtype=float)
    gpn = ref.gaussian_process_nobject(ys, grids, y_err=yes)
    gpn.find_hyperparameters(hyperparameter_guess=[0.5, 2], svd_method=False)
    joint = np.array(gpn.hyperparameters, dtype=float)
    ll_at_joint = ll_of(gpn, joint, 0.0, False)
    save("notebook_with_noise", x=np.array(grids), y=np.array(ys), y_err=np.array(yes),
         fit_single=single, fit_joint=joint, ll_at_joint=ll_at_joint,
         printed_single=np.array([0.57398201063394916, 2.2800929603014994]),
         printed_joint=np.array([0.51115288556575234, 2.0304985324357414]))

    # docs/notebook/1D_kernel_example_with_white_noise.ipynb cells 3,9 (single-object nugget fit)
    np.random.seed(1)
    grids, ys = [], []
    for _ in range(100):
        n_point = int(np.random.uniform(40, 40))          # consumes one draw, like the notebook
        grid = np.linspace(-10, 10, n_point)
        k = ref.rbf_kernel_1d(grid, np.array([0.5, 2]), nugget=0.1)
        ys.append(np.random.multivariate_normal(np.zeros_like(grid), k))
        grids.append(grid)
    gp = ref.gaussian_process(ys[0], grids[0])
    gp.find_hyperparameters(hyperparameter_guess=[0.5, 2], nugget=True, svd_method=True)
    # (the same fit with svd_method=False dies with LinAlgError inside fmin: with no y_err
    #  the simplex walks to nugget ~ 0 where K is numerically singular)
    ll_svd = ll_of(gp, list(gp.hyperparameters), float(gp.nugget), True)
    ll_chol = ll_of(gp, list(gp.hyperparameters), float(gp.nugget), False)
    save("notebook_white_noise", x=np.array(grids), y=np.array(ys),
         fit_single_svd=np.array(list(gp.hyperparameters) + [gp.nugget], dtype=float),
         ll_at_fit_svd=ll_svd, ll_at_fit_chol=ll_chol,
         printed_single=np.array([0.64033549419520619, 2.0650717053156979, 0.0833030775856]))


# ----------------------------------------------------------------- 7a. the reference's DEFAULT path (svd_method=True)
def svd_default():
    # docs/notebook/1D_kernel_example_without_noise.ipynb cells 1,3,7,9,11: noise-free single object, default arguments
    np.random.seed(1)
    grids, ys = [], []
    for _ in range(100):
        n_point = int(np.random.uniform(25, 30))
        grid = np.linspace(-10, 30, n_point)
        k = ref.rbf_kernel_1d(grid, np.array([0.5, 1]), nugget=0)
        ys.append(np.random.multivariate_normal(np.zeros_like(grid), k))
        grids.append(grid)
    gp = ref.gaussian_process(ys[0], grids[0])
    gp.find_hyperparameters(hyperparameter_guess=[0.5, 1])
    fit = np.array(gp.hyperparameters, dtype=float)
    new_grid = np.linspace(-10, 30, 60)
    gp.get_prediction(new_binning=new_grid)
    # a numerically SINGULAR covariance: duplicated epochs, no noise -> Cholesky fails, the default path pseudo-inverts
    rng = np.random.default_rng(12)
    xs = np.repeat(np.sort(rng.uniform(0, 10, 12)), 2)
    yv = np.sin(xs) + 0.01 * rng.standard_normal(len(xs))
    xg = np.sort(rng.uniform(0, 10, 17)); yg = np.sin(xg)           # a well-posed companion object
    hyp = np.array([0.8, 1.5])
    g2 = ref.gaussian_process_nobject([yv, yg], [xs, xg], y_err=[np.zeros(len(xs)), np.full(len(xg), 0.1)])
    g2.hyperparameters = hyp
    ll_svd = ll_of(g2, hyp, 0.0, True)
    with ref_loader.quiet():
        per = per_object_ll([yv, yg], [xs, xg], ref.rbf_kernel_1d, hyp, 0.0, [np.zeros(len(xs)), np.full(len(xg), 0.1)],
                            [0.0, 0.0], svd_method=True)
    grid2 = np.linspace(0, 10, 9)
    with ref_loader.quiet():
        g2.get_prediction(new_binning=grid2, COV=True, svd_method=True)
    chol_fails = False
    try:
        ll_of(g2, hyp, 0.0, False)
    except np.linalg.LinAlgError:
        chol_fails = True
    assert chol_fails
    save("svd_default", x=grids[0], y=ys[0], fit_single=fit, printed_single=np.array([0.54652372907962654, 1.228403962377433]),
         new_grid=new_grid, pred=np.array(gp.Prediction[0]), var=np.diag(gp.covariance_matrix[0]),
         sing_x=xs, sing_y=yv, ok_x=xg, ok_y=yg, sing_hyp=hyp, sing_ll_total=ll_svd, sing_ll_per_object=per,
         sing_grid=grid2, sing_pred=np.array(g2.Prediction), sing_var=np.array([np.diag(c) for c in g2.covariance_matrix]))


# ----------------------------------------------------------------- 7b. joint 2D fit
def fit_2d():
    """find_hyperparameters on three 2D objects (Gaussian_process.py:216-253).  At HEAD the 2D likelihood does not
    depend on sigma (quirk Q2), so only l_x, l_y, l_xy and the likelihood at the optimum are meaningful."""
    g = np.load(os.path.join(HERE, "batch_2d.npz"))
    off = g["off"]
    xs = [g["x"][off[i]:off[i + 1]] for i in range(3)]; ys = [g["y"][off[i]:off[i + 1]] for i in range(3)]
    yes = [g["y_err"][off[i]:off[i + 1]] for i in range(3)]
    guess = [1.0, 40.0, 40.0, 10.0]
    with ref_loader.quiet():
        gp = ref.gaussian_process_nobject(ys, xs, kernel="RBF2D", y_err=yes)
        gp.find_hyperparameters(hyperparameter_guess=guess, svd_method=False)
        fit = np.array(gp.hyperparameters, dtype=float)
        ll = ll_of(gp, fit, 0.0, False)
    save("fit_2d", guess=np.array(guess), fit_hyp=fit, ll_at_fit=ll)


# ----------------------------------------------------------------- 8. constructor mean options
def mean_options():
    """Gaussian_process.py:171-186: substract_mean=True without a template (y0 = mean(y) per object), and a
    template with the offsets `diff` handed in; likelihood and prediction (own epochs and a shared grid)."""
    rng = np.random.default_rng(11)
    b = 7
    hyp = np.array([0.6, 2.5]); nug = 0.04
    xs = [np.sort(rng.uniform(-5, 35, int(n))) for n in rng.integers(8, 40, b)]
    yes = [rng.uniform(0.05, 0.3, len(x)) for x in xs]
    ys = [3.0 + 0.5 * i + np.sin(x / 3.0) + 0.2 * rng.standard_normal(len(x)) for i, x in enumerate(xs)]
    grid = np.linspace(-6, 36, 41)
    out = {}
    with ref_loader.quiet():
        gp = ref.gaussian_process_nobject(ys, xs, y_err=yes, substract_mean=True)
        gp.hyperparameters = hyp
        out["ll_sub"] = ll_of(gp, hyp, nug, False)
        out["y0_sub"] = flat([np.ones(len(xs[i])) * gp.y0[i] for i in range(b)])[0]
        gp.get_prediction(new_binning=grid, svd_method=False)
        out["mean_sub"] = np.array(gp.Prediction, dtype=float)
        out["var_sub"] = np.array([np.diag(c) for c in gp.covariance_matrix])
        tm = np.linspace(-8, 38, 30); ym = 3.0 + np.sin(tm / 3.0)
        diff = list(0.5 * np.arange(b) + 0.05 * rng.standard_normal(b))
        gd = ref.gaussian_process_nobject(ys, xs, y_err=yes, Mean_Y=ym, Time_mean=tm, diff=diff)
        gd.hyperparameters = hyp
        out["ll_diff"] = ll_of(gd, hyp, nug, False)
        gd.get_prediction(new_binning=grid, svd_method=False)
        out["mean_diff"] = np.array(gd.Prediction, dtype=float)
        out["var_diff"] = np.array([np.diag(c) for c in gd.covariance_matrix])
        g1 = ref.gaussian_process(ys[2], xs[2], y_err=yes[2], Mean_Y=ym, Time_mean=tm)     # single object, own epochs
        g1.hyperparameters = hyp
        out["ll_one"] = ll_of(g1, hyp, nug, False)
        g1.get_prediction(new_binning=None, svd_method=False)
        out["mean_one"] = np.array(g1.Prediction[0], dtype=float)
        out["var_one"] = np.diag(g1.covariance_matrix[0])
    x, off = flat(xs)
    save("mean_options", x=x, y=flat(ys)[0], y_err=flat(yes)[0], off=off, hyp=hyp, nugget=nug, grid=grid,
         mean_x=tm, mean_y=ym, diff=np.array(diff), **out)


if __name__ == "__main__":
    np.seterr(all="ignore")
    which = sys.argv[1:] or ["kat_1d", "kat_2d", "c1_single", "ragged_1d", "pulls_1d", "batch_2d", "notebooks", "fit_2d", "mean_options", "svd_default"]
    for w in which:
        globals()[w]()
