"""The drop-in objects, used the way cosmogp's notebooks and pull.py use the reference
(gaussian_process / gaussian_process_nobject / build_pull), against the golden fixtures."""
import numpy as np
import pytest

from conftest import assert_close, golden, split

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cg():
    import cosmogp_b200
    from cosmogp_b200 import _lib
    _lib.require_device()
    return cosmogp_b200


def test_single_object_workflow(cg):
    g = golden("kat_1d")
    gp = cg.gaussian_process(g["y"], g["x"], y_err=g["y_err"])
    assert_close(gp.hyperparameters, g["init_rbf"], 1e-15)
    gp.hyperparameters = g["hyp"]; gp.nugget = float(g["nugget"])
    gp.compute_log_likelihood(g["hyp"], svd_method=False)
    assert gp.log_likelihood.shape == (1,)
    assert_close(gp.log_likelihood[0], g["ll_chol"], 1e-9)
    gp.get_prediction(new_binning=g["grid"], svd_method=False)
    assert_close(gp.Prediction[0], g["mean"], 1e-9)
    assert_close(gp.kernel_matrix[0], g["kmat"], 1e-12)
    assert_close(gp.inv_kernel_matrix[0], g["kinv"], 1e-9, 1e-10)
    assert_close(gp.covariance_matrix[0], g["cov"], 1e-9, 1e-12)
    assert_close(gp.prediction_variance[0], np.diag(g["cov"]), 1e-9)
    with pytest.raises(AssertionError):
        cg.gaussian_process(g["y"], g["x"], kernel="RBF3D")
    with pytest.raises(AssertionError):
        gp.find_hyperparameters(hyperparameter_guess=[1.0])


def test_fit_hyperparameters_c1(cg):
    """find_hyperparameters drives scipy's Nelder-Mead with one device launch per evaluation; the
    optimum must agree with the reference's to the optimiser's own tolerance (xtol = 1e-4)."""
    g = golden("c1_single")
    gp = cg.gaussian_process(g["y"], g["x"], y_err=g["y_err"])
    gp.find_hyperparameters(hyperparameter_guess=[0.5, 8.0], svd_method=False)
    assert_close(gp.hyperparameters, g["fit_hyp"], 1e-4)
    gn = cg.gaussian_process(g["y"], g["x"], y_err=g["y_err"])
    gn.find_hyperparameters(hyperparameter_guess=[0.5, 8.0], nugget=True, svd_method=False)
    assert_close(list(gn.hyperparameters) + [gn.nugget], list(g["fit_hyp_nugget"]) + [float(g["fit_nugget"])], 2e-3, 2e-4)
    gp.hyperparameters = g["hyp"]; gp.nugget = float(g["nugget"])
    gp.get_prediction(new_binning=g["grid"], COV=True, svd_method=False)
    assert_close(gp.Prediction[0], g["mean"], 1e-9, 1e-12)
    cov = gp.covariance_matrix[0]
    assert cov.shape == (500, 500)
    assert_close(np.diag(cov), g["cov_diag"], 1e-9, 1e-13)
    assert_close(cov[100:140, 100:140], g["cov_block"], 1e-9, 1e-12)
    assert_close(cov.sum(), g["cov_sum"], 1e-8)


def test_nobject_shared_mean_workflow(cg):
    g = golden("ragged_1d")
    off = g["off"]
    xs, ys, yes = (split(g[k], off) for k in ("x", "y", "y_err"))
    gp = cg.gaussian_process_nobject(ys, xs, y_err=yes, Mean_Y=g["mean_y"], Time_mean=g["mean_x"])
    assert_close(np.concatenate(gp.y0), g["y0"], 1e-15)
    assert_close(gp.hyperparameters, g["init_rbf"], 1e-15)
    gp.hyperparameters = g["hyp"].copy()
    gp.compute_log_likelihood(g["hyp"], svd_method=False)
    assert_close(gp.log_likelihood[0], g["ll_sum"], 1e-9)
    assert_close(gp.log_likelihood_per_object, g["ll_obj"], 1e-9)
    gp.get_prediction(new_binning=g["grid"], svd_method=False)
    assert_close(np.array(gp.Prediction), g["mean"], 1e-9)
    assert_close(np.array(gp.prediction_variance), g["var"], 1e-9, 1e-13)
    assert_close(gp.covariance_matrix[0], g["cov0"], 1e-9, 1e-12)
    assert_close(gp.kernel_matrix[0], g["kmat0"], 1e-12); assert_close(gp.inv_kernel_matrix[0], g["kinv0"], 1e-9, 1e-9)
    gf = cg.gaussian_process_nobject(ys, xs, y_err=yes, Mean_Y=g["mean_y"], Time_mean=g["mean_x"])
    gf.find_hyperparameters(hyperparameter_guess=[0.5, 2.0], svd_method=False)
    assert_close(gf.hyperparameters, g["fit_hyp"], 1e-4)
    # own-epoch prediction
    g1 = cg.gaussian_process(ys[0], xs[0], y_err=yes[0], Mean_Y=g["mean_y"], Time_mean=g["mean_x"])
    g1.hyperparameters = g["hyp"].copy(); g1.nugget = 0.05
    g1.get_prediction(svd_method=False)
    assert_close(g1.Prediction[0], g["own_mean0"], 1e-9)
    assert_close(g1.covariance_matrix[0], g["own_cov0"], 1e-9, 1e-12)


def test_constructor_mean_options(cg):
    """substract_mean=True without a template, template + given offsets, single object at its own epochs:
    the facade against the real reference's outputs (tests/golden/mean_options.npz)."""
    g = golden("mean_options")
    off, hyp, nug, grid = g["off"], g["hyp"], float(g["nugget"]), g["grid"]
    xs, ys, yes = (split(g[k], off) for k in ("x", "y", "y_err"))
    gp = cg.gaussian_process_nobject(ys, xs, y_err=yes, substract_mean=True)
    assert_close(np.concatenate(gp.y0), g["y0_sub"], 1e-15)
    gp.hyperparameters = hyp.copy(); gp.nugget = nug
    gp.compute_log_likelihood(hyp, svd_method=False)
    assert_close(gp.log_likelihood[0], g["ll_sub"], 1e-9)
    gp.get_prediction(new_binning=grid, svd_method=False)
    assert_close(np.array(gp.Prediction), g["mean_sub"], 1e-9); assert_close(np.array(gp.prediction_variance), g["var_sub"], 1e-9, 1e-13)
    gd = cg.gaussian_process_nobject(ys, xs, y_err=yes, Mean_Y=g["mean_y"], Time_mean=g["mean_x"], diff=list(g["diff"]))
    gd.hyperparameters = hyp.copy(); gd.nugget = nug
    gd.compute_log_likelihood(hyp, svd_method=False)
    assert_close(gd.log_likelihood[0], g["ll_diff"], 1e-9)
    gd.get_prediction(new_binning=grid, COV='diag', svd_method=False)
    assert_close(np.array(gd.Prediction), g["mean_diff"], 1e-9); assert_close(np.array(gd.prediction_variance), g["var_diff"], 1e-9, 1e-13)
    assert_close(np.asarray(gd.warning_pf)[3], O_mean_on_grid(g, 3), 1e-14)
    g1 = cg.gaussian_process(ys[2], xs[2], y_err=yes[2], Mean_Y=g["mean_y"], Time_mean=g["mean_x"])
    g1.hyperparameters = hyp.copy(); g1.nugget = nug
    g1.compute_log_likelihood(hyp, svd_method=False)
    assert_close(g1.log_likelihood[0], g["ll_one"], 1e-9)
    g1.get_prediction(svd_method=False)
    assert_close(g1.Prediction[0], g["mean_one"], 1e-9); assert_close(np.diag(g1.covariance_matrix[0]), g["var_one"], 1e-9, 1e-13)


def O_mean_on_grid(g, i):
    from oracle import gp_oracle as O
    off = g["off"]
    xs, ys = split(g["x"], off), split(g["y"], off)
    return O.return_mean_1d(ys[i], xs[i], g["mean_y"], g["mean_x"], diff=g["diff"][i], new_x=g["grid"])


def test_notebook_fits(cg):
    """docs/notebook/1D_kernel_example_with_noise.ipynb cells 9 and 21."""
    g = golden("notebook_with_noise")
    gp = cg.gaussian_process(g["y"][0], g["x"][0], y_err=g["y_err"][0])
    gp.find_hyperparameters(hyperparameter_guess=[0.5, 2], svd_method=False)
    assert_close(gp.hyperparameters, g["printed_single"], 1e-4)
    gn = cg.gaussian_process_nobject(g["y"], g["x"], y_err=g["y_err"])        # 2-D ndarray input
    gn.find_hyperparameters(hyperparameter_guess=[0.5, 2], svd_method=False)
    assert_close(gn.hyperparameters, g["printed_joint"], 1e-4)


def test_build_pull_modes(cg):
    g = golden("pulls_1d")
    bp = cg.build_pull(list(g["y"]), list(g["x"]), g["hyp"], nugget=float(g["nugget"]), y_err=list(g["y_err"]))
    bp.compute_pull(svd_method=False)
    assert_close(bp.pull, g["pullA"], 1e-9, 1e-12); assert_close(bp.residual, g["residA"], 1e-9, 1e-12)
    assert_close(np.array(bp.prediction), g["predA"], 1e-9, 1e-12)
    assert_close([bp.pull_average, bp.pull_std], [g["pullA_avg"], g["pullA_std"]], 1e-9)
    bd = cg.build_pull(list(g["yD"]), [g["xD"]] * 3, g["hyp"], nugget=float(g["nugget"]), y_err=list(g["y_err"][:3]))
    bd.compute_pull(svd_method=False, substract_mean=True)
    assert_close(bd.pull, g["pullD"], 1e-8, 1e-11)
    r = golden("ragged_1d")
    off = r["off"]
    xs, ys, yes = (split(r[k], off) for k in ("x", "y", "y_err"))
    bb = cg.build_pull(ys, xs, r["hyp"], nugget=0.05, y_err=yes, y_mean=r["mean_y"], x_axis_mean=r["mean_x"])
    bb.compute_pull(svd_method=False)
    assert_close(bb.pull, r["pullB"], 1e-8, 1e-11)
    assert_close([bb.pull_average, bb.pull_std], [r["pullB_avg"], r["pullB_std"]], 1e-8)
    bc = cg.build_pull(ys, xs, r["hyp"], nugget=0.05, y_err=yes, y_mean=r["mean_y"], x_axis_mean=r["mean_x"])
    bc.compute_pull(diff=list(r["diff"]), svd_method=False)
    assert_close(bc.pull, r["pullC"], 1e-8, 1e-11)
    # the reference's lists grow with every compute_pull call (pull.py:94-102): the normal law is refitted on all of them
    bc.compute_pull(svd_method=False)
    assert len(bc.pull) == 2 * len(r["pullC"]) and len(bc._pull) == 2 * len(ys) and len(bc.prediction) == 2 * len(ys)
    assert_close(bc.pull[len(r["pullC"]):], r["pullB"], 1e-8, 1e-11)
    both = np.concatenate([r["pullC"], r["pullB"]])
    assert_close([bc.pull_average, bc.pull_std], [np.mean(both), np.std(both)], 1e-8)
    assert_close(bc._pull[len(ys) + 1], split(r["pullB"], off)[1], 1e-8, 1e-11)


def test_streamed_evaluator_other_shapes(cg):
    """The pipelined evaluator beyond the one-warp kernels: 100-point objects (generic kernel), 2D objects,
    likelihood only (no grid), optional inputs left out, a batch smaller than one chunk."""
    from cosmogp_b200.batch import DeviceBatch, StreamedEvaluator
    rng = np.random.default_rng(9)
    # N = 100 -> one 4-warp CTA per object
    b, n, m = 300, 100, 24
    x = np.sort(rng.uniform(0, 60, (b, n)), axis=1); y = rng.standard_normal((b, n)); ye = np.full((b, n), 0.3)
    grid = np.linspace(0, 60, m)
    ev = StreamedEvaluator(b, n, m, n_chunks=3, n_streams=2)
    for k, v in (("x", x), ("y", y), ("y0", 0.0), ("y_err", ye), ("new_y0", 0.0)):
        ev.host(k)[...] = v
    tot, ll, mean, var, info = ev.run([0.7, 3.0], 0.02, grid)
    batch = DeviceBatch(x.ravel(), y.ravel(), np.arange(b + 1) * n, y_err=ye.ravel())
    t2, ll2, _ = batch.log_likelihood([0.7, 3.0], 0.02)
    m2, v2, _ = batch.predict([0.7, 3.0], 0.02, grid)
    assert np.array_equal(ll, ll2) and tot == t2 and not info.any()
    assert_close(mean, m2, 1e-10, 1e-11); assert_close(var, v2, 1e-10, 1e-11)
    # 2D objects
    b, n, m = 257, 30, 20
    xy = rng.uniform(0, 10, (b, n, 2)); z = rng.standard_normal((b, n)); ze = np.full((b, n), 0.2)
    g2 = rng.uniform(0, 10, (m, 2)); h2 = [1.0, 2.0, 1.5, 0.3]
    ev = StreamedEvaluator(b, n, m, dim=2, n_chunks=2, n_streams=2)
    for k, v in (("x", xy), ("y", z), ("y0", 0.0), ("y_err", ze), ("new_y0", 0.0)):
        ev.host(k)[...] = v
    tot, ll, mean, var, info = ev.run(h2, 0.05, g2)
    batch = DeviceBatch(xy.reshape(-1, 2), z.ravel(), np.arange(b + 1) * n, y_err=ze.ravel(), dim=2)
    _, ll2, _ = batch.log_likelihood(h2, 0.05)
    m2, v2, _ = batch.predict(h2, 0.05, g2)
    assert np.array_equal(ll, ll2) and np.array_equal(mean, m2) and np.array_equal(var, v2)
    # likelihood only
    ev = StreamedEvaluator(b, n, 0, dim=2, n_chunks=2, n_streams=2)
    for k, v in (("x", xy), ("y", z), ("y0", 0.0), ("y_err", ze)):
        ev.host(k)[...] = v
    tot, ll3, mean, var, info = ev.run(h2, 0.05, np.zeros((0, 2)))
    assert np.array_equal(ll3, ll2) and mean.shape == (b, 0)


def test_mean_spline_on_device(cg):
    """cgp_spline_mean_dev = scipy's InterpolatedUnivariateSpline (FITPACK splev) + per-object offset, including
    extrapolation beyond the template; and the streamed evaluator computing y0 from it instead of uploading it."""
    import torch
    import scipy.interpolate as inter
    from cosmogp_b200 import _lib
    from cosmogp_b200.batch import StreamedEvaluator
    rng = np.random.default_rng(4)
    tm = np.sort(rng.uniform(-15, 45, 40)); ym = np.sin(tm / 5) + 0.1 * rng.standard_normal(40)
    spl = inter.InterpolatedUnivariateSpline(tm, ym)
    t, c, k = spl._eval_args
    sizes = rng.integers(0, 70, 500); off = np.zeros(501, dtype=np.int64); off[1:] = np.cumsum(sizes)
    x = np.concatenate([rng.uniform(-30, 60, off[-1] - 42), tm, [tm[0], tm[-1]]]); diff = rng.standard_normal(500)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out = torch.empty(len(x), dtype=torch.float64, device="cuda")
    args = [dev(t), dev(c), dev(x), dev(off), dev(diff)]
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib().cgp_spline_mean_dev(args[0].data_ptr(), args[1].data_ptr(), len(t), args[2].data_ptr(), len(x),
                                              args[3].data_ptr(), 500, args[4].data_ptr(), out.data_ptr(), st), "spline")
    ref = spl(x) + np.repeat(diff, sizes)
    assert np.array_equal(out.cpu().numpy(), ref)               # same operations in the same order as FITPACK
    _lib.check(_lib.lib().cgp_spline_mean_dev(args[0].data_ptr(), args[1].data_ptr(), len(t), args[2].data_ptr(), len(x),
                                              None, 0, None, out.data_ptr(), st), "spline")
    assert np.array_equal(out.cpu().numpy(), spl(x))
    # streamed evaluator: y0 evaluated on the device from the template == y0 computed by scipy and uploaded
    b, n, m = 6000, 30, 40
    xs = np.sort(rng.uniform(-10, 40, (b, n)), axis=1); ys = rng.standard_normal((b, n)); ye = np.full((b, n), 0.2)
    d = rng.standard_normal(b); grid = np.linspace(-10, 40, m)
    res = []
    for on_device in (False, True):
        ev = StreamedEvaluator(b, n, m, n_chunks=2, n_streams=2, shared_mean=True)
        for kk, v in (("x", xs), ("y", ys), ("y_err", ye), ("diff", d), ("template", spl(grid))):
            ev.host(kk)[...] = v
        if on_device:
            ev.set_mean_template(tm, ym)
            ev.host("y0")[...] = np.nan                          # must not be read
        else:
            ev.host("y0")[...] = spl(xs.ravel()).reshape(b, n) + d[:, None]
        tot, ll, mean, var, info = ev.run([0.5, 2.0], 0.03, grid)
        res.append((tot, ll.copy(), mean.copy(), var.copy(), ev.h2d_bytes))
    assert res[0][0] == res[1][0] and all(np.array_equal(res[0][i], res[1][i]) for i in (1, 2, 3))
    assert res[0][4] - res[1][4] == b * n * 8                    # one double per data point less on the wire


def test_moments_on_device(cg):
    """cgp_moments_dev (the norm.fit of the pulls without a download) against numpy, any size and alignment."""
    import torch
    from cosmogp_b200.batch import DeviceBatch
    rng = np.random.default_rng(2)
    b = DeviceBatch(np.arange(4.0), np.arange(4.0), np.array([0, 4]))
    for n in (1, 2, 3, 255, 4097, 65536, 1000003):
        v = rng.standard_normal(n + 1) * 3.0 + 0.7
        d = torch.from_numpy(v).cuda()
        for view, ref in ((d[:n], v[:n]), (d[1:], v[1:])):          # second view: only 8-byte aligned
            s1, s2 = b.moments(view, 0.25)
            assert_close(s1, np.sum(ref - 0.25), 1e-11, 1e-9); assert_close(s2, np.sum((ref - 0.25) ** 2), 1e-12)
    assert b.moments(torch.zeros(0, dtype=torch.float64, device="cuda"), 1.0) == (0.0, 0.0)


def test_2d_objects(cg):
    g = golden("batch_2d")
    off = g["off"]
    xs, ys, yes = split(g["x"], off), split(g["y"], off), split(g["y_err"], off)
    gp = cg.gaussian_process_nobject(ys, xs, kernel="RBF2D", y_err=yes)
    assert_close(gp.hyperparameters, g["init_rbf"], 1e-15)
    gp.hyperparameters = g["hyp"].copy(); gp.nugget = float(g["nugget"])
    gp.compute_log_likelihood(g["hyp"], svd_method=False)
    assert_close(gp.log_likelihood[0], g["ll_sum"], 1e-9)
    gp.get_prediction(new_binning=g["grid"], svd_method=False)
    assert_close(np.array(gp.Prediction), g["mean"], 1e-9, 1e-12)
    assert_close(gp.covariance_matrix[2], g["cov2"], 1e-9, 1e-11)
    assert_close(gp.kernel_matrix[2], g["kmat2"], 1e-12, 1e-300)
    bp = cg.build_pull(ys[2:], xs[2:], g["hyp"], nugget=float(g["nugget"]), y_err=yes[2:], kernel="RBF2D")
    bp.compute_pull(svd_method=False)
    assert_close(bp.pull, g["pull2"], 1e-9, 1e-12)


def test_2d_joint_fit(cg):
    """find_hyperparameters on 2D objects against the real reference's optimum.  The 2D likelihood does not depend on
    sigma at HEAD (quirk Q2), so Nelder-Mead drifts along sigma: the length scales and the likelihood are compared."""
    g, f = golden("batch_2d"), golden("fit_2d")
    off = g["off"]
    xs, ys, yes = split(g["x"], off), split(g["y"], off), split(g["y_err"], off)
    gp = cg.gaussian_process_nobject(ys, xs, kernel="RBF2D", y_err=yes)
    gp.find_hyperparameters(hyperparameter_guess=list(f["guess"]), svd_method=False)
    gp.compute_log_likelihood(gp.hyperparameters, svd_method=False)
    print("2D fit", gp.hyperparameters, gp.log_likelihood[0], "reference", f["fit_hyp"], float(f["ll_at_fit"]))
    assert_close(gp.log_likelihood[0], float(f["ll_at_fit"]), 1e-6)
    assert_close(gp.hyperparameters[1:], f["fit_hyp"][1:], 2e-3)


def test_not_positive_definite_raises_linalgerror(cg):
    x = np.array([0.0, 1.0, 1.0, 2.0]); y = np.array([0.1, 0.2, 0.3, 0.4])
    gp = cg.gaussian_process(y, x)
    with pytest.raises(np.linalg.LinAlgError):
        gp.compute_log_likelihood([1.0, 1.0], svd_method=False)


def test_svd_method_default_on_the_noise_free_notebook(cg):
    """The reference's DEFAULT arguments (svd_method=True) on its noise-free notebook
    (docs/notebook/1D_kernel_example_without_noise.ipynb cells 7-11): fitted hyperparameters as printed there, prediction
    and variance against the fixture made by the real reference.  K is noise free but positive definite along the
    whole simplex path, so the device Cholesky serves every evaluation."""
    g = golden("svd_default")
    gp = cg.gaussian_process(g["y"], g["x"])
    gp.find_hyperparameters(hyperparameter_guess=[0.5, 1])
    assert_close(gp.hyperparameters, g["printed_single"], 1e-4)
    gp.hyperparameters = list(g["fit_single"])
    gp.get_prediction(new_binning=g["new_grid"])
    assert_close(gp.Prediction[0], g["pred"], 1e-7, 1e-9)
    assert_close(gp.prediction_variance[0], g["var"], 1e-6, 1e-9)      # noise-free: variances down to 1e-16 at the data


def test_svd_method_default_on_a_singular_covariance(cg):
    """Duplicated epochs without noise: the reference's default path pseudo-inverts (inv_matrix.py:4-18) and never
    raises; svd_method=False raises LinAlgError (inv_matrix.py:23).  Here the object whose device factorisation
    reports a non-positive pivot is redone by the host SVD reference check when svd_method=True; the well-posed
    companion object keeps its device result."""
    import warnings
    g = golden("svd_default")
    ys, xs = [g["sing_y"], g["ok_y"]], [g["sing_x"], g["ok_x"]]
    yes = [np.zeros(len(g["sing_x"])), np.full(len(g["ok_x"]), 0.1)]
    gp = cg.gaussian_process_nobject(ys, xs, y_err=yes)
    gp.hyperparameters = list(g["sing_hyp"])
    with pytest.raises(np.linalg.LinAlgError):
        gp.compute_log_likelihood(g["sing_hyp"], svd_method=False)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        gp.compute_log_likelihood(g["sing_hyp"])                        # default: svd_method=True
        assert any("SVD reference check" in str(m.message) for m in w)
    assert_close(gp.log_likelihood_per_object[1], g["sing_ll_per_object"][1], 1e-9)
    assert_close(gp.log_likelihood_per_object[0], g["sing_ll_per_object"][0], 1e-6)   # 12 singular values truncated
    assert_close(gp.log_likelihood[0], float(g["sing_ll_total"]), 1e-6)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gp.get_prediction(new_binning=g["sing_grid"], COV='diag')
    assert_close(gp.Prediction[1], g["sing_pred"][1], 1e-9)
    assert_close(gp.prediction_variance[1], g["sing_var"][1], 1e-9)
    assert_close(gp.Prediction[0], g["sing_pred"][0], 1e-5, 1e-7)
    assert_close(gp.prediction_variance[0], g["sing_var"][0], 0.0, 1e-5)     # ~1e-7 numbers made of rounding noise
    with pytest.raises(np.linalg.LinAlgError):
        gp.get_prediction(new_binning=g["sing_grid"], COV='diag', svd_method=False)


def test_streamed_evaluator_matches_resident_batch(cg):
    """The pipelined end-to-end path (chunks over 3 streams) returns bit-identical results."""
    from cosmogp_b200.batch import DeviceBatch, StreamedEvaluator
    rng = np.random.default_rng(0)
    b, n, m = 5003, 60, 100
    x = np.sort(rng.uniform(-10, 40, (b, n)), axis=1); y = rng.standard_normal((b, n)); ye = np.full((b, n), 0.2)
    y0 = 0.1 * rng.standard_normal((b, n)); ny0 = rng.standard_normal((b, m))
    ev = StreamedEvaluator(b, n, m, n_chunks=7)
    for k, v in (("x", x), ("y", y), ("y0", y0), ("y_err", ye), ("new_y0", ny0)):
        ev.host(k)[...] = v
    grid = np.linspace(-10, 40, m)
    tot, ll, mean, var, info = ev.run([0.5, 2.0], 0.03, grid)
    batch = DeviceBatch(x.ravel(), y.ravel(), np.arange(b + 1) * n, y0=y0.ravel(), y_err=ye.ravel())
    t2, ll2, _ = batch.log_likelihood([0.5, 2.0], 0.03)
    m2, v2, _ = batch.predict([0.5, 2.0], 0.03, grid, new_y0=ny0)
    assert not info.any() and tot == t2
    assert np.array_equal(ll, ll2)
    # (batch.predict hints a uniform grid and, at >= 2048 objects, runs the recurrence kernel; the small chunks here do not)
    assert_close(mean, m2, 1e-10, 1e-11); assert_close(var, v2, 1e-10, 1e-11)
    assert ev.h2d_bytes == (b * (4 * n + m) + m) * 8 and ev.d2h_bytes == b * (8 + 16 * m + 4)     # + the grid itself
    # shared mean (template + offset per object): same results from M + B mean values instead of B x M
    tmpl, diff = np.cos(grid / 5.0), rng.standard_normal(b)
    ev2 = StreamedEvaluator(b, n, m, n_chunks=7, shared_mean=True)
    for k, v in (("x", x), ("y", y), ("y0", y0), ("y_err", ye), ("template", tmpl), ("diff", diff)):
        ev2.host(k)[...] = v
    tot3, ll3, mean3, var3, info3 = ev2.run([0.5, 2.0], 0.03, grid)
    m4, v4, _ = batch.predict([0.5, 2.0], 0.03, grid, new_y0=tmpl[None, :] + diff[:, None])
    assert tot3 == t2
    assert_close(mean3, m4, 1e-10, 1e-11); assert_close(var3, v4, 1e-10, 1e-11)
    assert ev2.h2d_bytes == (b * (4 * n + 1) + m + m) * 8                    # template and offsets once per run, grid once
    # large chunks: ramped schedule (2048, 3276, ... up to the buffer capacity, then down again), two-kernel route
    b, n, m = 40001, 20, 16
    x = np.sort(rng.uniform(-10, 40, (b, n)), axis=1); y = rng.standard_normal((b, n)); ye = np.full((b, n), 0.2)
    grid = np.linspace(-10, 40, m); tmpl, diff = np.cos(grid / 5.0), rng.standard_normal(b)
    ev3 = StreamedEvaluator(b, n, m, n_chunks=4, n_streams=3, shared_mean=True)
    for k, v in (("x", x), ("y", y), ("y0", 0.0), ("y_err", ye), ("template", tmpl), ("diff", diff)):
        ev3.host(k)[...] = v
    tot5, ll5, mean5, var5, info5 = ev3.run([0.5, 2.0], 0.03, grid)
    big = DeviceBatch(x.ravel(), y.ravel(), np.arange(b + 1) * n, y_err=ye.ravel())
    t6, ll6, _ = big.log_likelihood([0.5, 2.0], 0.03)
    m6, v6, _ = big.predict([0.5, 2.0], 0.03, grid, mean_template=(tmpl, diff))
    # (the likelihood comes from the factor kernel here: same pivots, z = inv(L) r instead of a forward substitution)
    assert_close(ll5, ll6, 1e-13, 1e-11); assert_close(tot5, t6, 1e-12)
    assert np.array_equal(mean5, m6) and np.array_equal(var5, v6) and not info5.any()
    assert ev3.d2h_bytes == b * (8 + 16 * m + 4)


def test_large_objects_through_the_facade(cg):
    """Objects beyond the shared-memory path (N > 224) are routed to the blocked HBM factorisation."""
    from oracle import gp_oracle as O
    rng = np.random.default_rng(8)
    n = 400
    x = rng.uniform(-100, 100, (n, 2)); ye = rng.uniform(0.15, 0.3, n)
    y = np.cos(x[:, 0] / 40) + 0.2 * rng.standard_normal(n)
    hyp, nug = [1.0, 30.0, 25.0, 50.0], 0.05
    gp = cg.gaussian_process(y, x, kernel="RBF2D", y_err=ye)
    gp.compute_log_likelihood(hyp, svd_method=False)
    gp.nugget = nug
    gp.compute_log_likelihood(hyp, svd_method=False)
    assert_close(gp.log_likelihood[0], O.log_likelihood(y, x, hyp, nug, ye, kind="2d"), 1e-9)
    grid = rng.uniform(-100, 100, (150, 2))
    gp.hyperparameters = np.array(hyp)
    gp.get_prediction(new_binning=grid, COV='diag')
    mo, vo = O.predict(y, x, hyp, nug, grid, ye, kind="2d", full_cov=False)
    assert_close(gp.Prediction[0], mo, 1e-9, 1e-11); assert_close(gp.prediction_variance[0], vo, 1e-9, 1e-12)
    # 1D, with a mean template, own-epoch prediction
    n1 = 300
    x1 = np.sort(rng.uniform(0, 100, n1)); ye1 = np.full(n1, 0.2)
    tm = np.linspace(-5, 105, 40); ym = np.sin(tm / 15)
    y1 = np.sin(x1 / 15) + 0.3 + 0.2 * rng.standard_normal(n1)
    g1 = cg.gaussian_process(y1, x1, y_err=ye1, Mean_Y=ym, Time_mean=tm)
    g1.hyperparameters = np.array([0.5, 3.0]); g1.nugget = 0.0
    g1.get_prediction(COV='diag')
    y0 = O.return_mean_1d(y1, x1, ym, tm)
    mo, vo = O.predict(y1, x1, [0.5, 3.0], 0.0, x1, ye1, y0, y0, full_cov=False)
    assert_close(g1.Prediction[0], mo, 1e-9, 1e-11); assert_close(g1.prediction_variance[0], vo, 1e-9, 1e-12)


def test_bulk_covariance_writer(cg):
    """covariance_matrix of many objects at once (cgp_covariance_batched_dev) against the reference fixture and the
    oracle: shared grid, per-object epochs (new_binning=None), 1D and 2D."""
    from conftest import golden
    from oracle import gp_oracle as O
    from cosmogp_b200.batch import DeviceBatch, pack_csr
    g = golden("kat_1d")
    gp = cg.gaussian_process(g["y"], g["x"], y_err=g["y_err"])
    gp.hyperparameters = g["hyp"]; gp.nugget = float(g["nugget"])
    gp.get_prediction(new_binning=g["grid"], svd_method=False)
    assert_close(gp.covariance_matrix[0], g["cov"], 1e-9, 1e-12)
    rng = np.random.default_rng(17)
    sizes = rng.integers(3, 64, 300)
    xs = [np.sort(rng.uniform(-10, 40, n)) for n in sizes]
    ys = [rng.standard_normal(n) for n in sizes]; yes = [rng.uniform(0.1, 0.3, n) for n in sizes]
    hyp, nug = [0.6, 2.5], 0.07
    grid = np.linspace(-12, 42, 77)
    gn = cg.gaussian_process_nobject(ys, xs, y_err=yes)
    gn.hyperparameters = np.array(hyp); gn.nugget = nug
    gn.get_prediction(new_binning=grid, COV=True, svd_method=False)
    for i in (0, 150, 299):
        _, co = O.predict(ys[i], xs[i], hyp, nug, grid, yes[i])
        assert_close(gn.covariance_matrix[i], co, 1e-9, 1e-11)
        assert_close(np.diag(gn.covariance_matrix[i]), gn.prediction_variance[i], 1e-9, 1e-12)
    gn.get_prediction(COV=True, svd_method=False)                      # every object on its own epochs
    for i in (3, 200):
        _, co = O.predict(ys[i], xs[i], hyp, nug, xs[i], yes[i])
        assert gn.covariance_matrix[i].shape == (sizes[i], sizes[i])
        assert_close(gn.covariance_matrix[i], co, 1e-9, 1e-11)
    # all matrices of a batch in one call, 2D objects
    x2 = [rng.uniform(-50, 50, (n, 2)) for n in sizes[:40]]
    xf, off = pack_csr(x2, 2); yf, _ = pack_csr(ys[:40], 1); ef, _ = pack_csr(yes[:40], 1)
    b2 = DeviceBatch(xf, yf, off, y_err=ef, dim=2)
    h2 = [1.1, 20.0, 15.0, 30.0]; g2 = rng.uniform(-50, 50, (33, 2))
    mats, info = b2.covariance(h2, 0.05, g2)
    assert mats.shape == (40, 33, 33) and not info.any()
    for i in (0, 39):
        _, co = O.predict(ys[i], x2[i], h2, 0.05, g2, yes[i], kind="2d")
        assert_close(mats[i], co, 1e-9, 1e-11)
