"""Single-process multi-GPU context (cgp_ctx_*): sharded likelihood, shared-grid prediction and pulls equal the
one-GPU results and the oracle.  With one GPU the context has one shard (still the cgp_ctx code path); on a box
with 2+ GPUs every available GPU is used and both output routes are checked (per-GPU PCIe, NCCL gather over NVLink)."""
import numpy as np
import pytest

from conftest import assert_close
from oracle import gp_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    import cosmogp_b200 as cg
    from cosmogp_b200 import _lib
    _lib.require_device()
    return cg, torch.cuda.device_count()


def _ragged(rng, b, lo=5, hi=60):
    sizes = rng.integers(lo, hi + 1, b)
    xs = [np.sort(rng.uniform(-10, 40, n)) for n in sizes]
    ys = [0.5 * np.sin(x / 3.0 + rng.uniform(0, 6)) + 0.2 * rng.standard_normal(len(x)) for x in xs]
    yes = [rng.uniform(0.15, 0.25, len(x)) for x in xs]
    return xs, ys, yes


@pytest.mark.parametrize("gather", [False, True])
def test_sharded_facade_matches_single_device(env, gather):
    cg, ngpu = env
    rng = np.random.default_rng(21)
    xs, ys, yes = _ragged(rng, 3001)
    tmean = np.linspace(-15, 45, 61); ymean = 0.3 * np.sin(tmean / 10)
    grid = np.linspace(-10, 40, 50)
    one = cg.gaussian_process_nobject(ys, xs, y_err=yes, Mean_Y=ymean, Time_mean=tmean)
    many = cg.gaussian_process_nobject(ys, xs, y_err=yes, Mean_Y=ymean, Time_mean=tmean, devices="all")
    many.gather_over_nvlink = gather
    hyp = [0.5, 2.5]
    one.compute_log_likelihood(hyp, svd_method=False); many.compute_log_likelihood(hyp, svd_method=False)
    assert many.batch.n_devices == ngpu and sum(b - a for a, b in many.batch.ranges) == 3001
    assert_close(many.log_likelihood[0], one.log_likelihood[0], 1e-12)
    assert_close(many.log_likelihood_per_object, one.log_likelihood_per_object, 0.0, 0.0)      # same kernels: bit identical
    ref = O.log_likelihood(ys[7], xs[7], hyp, 0.0, yes[7], one.y0[7])
    assert_close(many.log_likelihood_per_object[7], ref, 1e-9)
    for gp in (one, many):
        gp.hyperparameters = np.array(hyp); gp.nugget = 0.05
        gp.get_prediction(new_binning=grid, COV='diag', svd_method=False)
    # a shard below 2048 objects takes the general prediction kernel, the full batch the uniform-grid recurrence: ~1e-10 apart
    assert_close(np.asarray(many.Prediction), np.asarray(one.Prediction), 1e-9, 1e-11)
    assert_close(np.asarray(many.prediction_variance), np.asarray(one.prediction_variance), 1e-9, 1e-12)
    mo, vo = O.predict(ys[7], xs[7], hyp, 0.05, grid, yes[7], one.y0[7], np.asarray(one.warning_pf)[7], full_cov=False)
    assert_close(many.Prediction[7], mo, 1e-9, 1e-11); assert_close(many.prediction_variance[7], vo, 1e-9, 1e-12)
    many.find_hyperparameters(hyperparameter_guess=[0.4, 3.0], svd_method=False)
    one.find_hyperparameters(hyperparameter_guess=[0.4, 3.0], svd_method=False)
    assert_close(many.hyperparameters, one.hyperparameters, 1e-6)


@pytest.mark.parametrize("gather", [False, True])
def test_sharded_pulls_match_single_device_and_oracle(env, gather):
    cg, ngpu = env
    rng = np.random.default_rng(5)
    b, n = 5000, 40
    x = np.sort(rng.uniform(-10, 10, (b, n)), axis=1)
    y = 0.5 * np.sin(x / 2.0 + rng.uniform(0, 6, (b, 1))) + 0.1 * rng.standard_normal((b, n))
    ye = np.full((b, n), 0.1)
    one = cg.build_pull(y, x, [0.5, 2.0], nugget=0.03, y_err=ye)
    one.compute_pull(svd_method=False)
    many = cg.build_pull(y, x, [0.5, 2.0], nugget=0.03, y_err=ye, devices="all", gather_over_nvlink=gather)
    many.compute_pull(svd_method=False)
    assert len(many.shard_ranges) == ngpu
    for name in ("pull", "residual"):
        assert_close(getattr(many, name), getattr(one, name), 0.0, 0.0)
    assert_close(np.asarray(many.prediction), np.asarray(one.prediction), 0.0, 0.0)
    assert_close(many.prediction_variance, one.prediction_variance, 0.0, 0.0)
    assert_close(many.pull_average, one.pull_average, 1e-9, 1e-12); assert_close(many.pull_std, one.pull_std, 1e-9)
    po = O.loo_batched_1d(x[:200], y[:200], ye[:200], [0.5, 2.0], 0.03)
    assert_close(many.pull[:200 * n], po[2].ravel(), 1e-8, 1e-10)
    # recentred mode (template mean, offset re-estimated on the kept points) through the sharded path
    tm = np.linspace(-11, 11, 23); ym = 0.1 * np.cos(tm / 4)
    o2 = cg.build_pull(y[:700], x[:700], [0.5, 2.0], nugget=0.03, y_err=ye[:700], y_mean=ym, x_axis_mean=tm)
    m2 = cg.build_pull(y[:700], x[:700], [0.5, 2.0], nugget=0.03, y_err=ye[:700], y_mean=ym, x_axis_mean=tm, devices="all",
                       gather_over_nvlink=gather)
    o2.compute_pull(svd_method=False); m2.compute_pull(svd_method=False)
    assert_close(m2.pull, o2.pull, 0.0, 0.0)


def test_context_collectives(env):
    """cgp_ctx_gather_f64 / cgp_ctx_allreduce_sum_f64 on per-device buffers."""
    import ctypes as C
    import torch
    cg, ngpu = env
    from cosmogp_b200 import _lib, multi
    L = _lib.lib()
    path = multi._nccl_library()
    if path:
        L.cgp_set_nccl_library(path.encode())
    ctx = C.c_void_p()
    _lib.check(L.cgp_ctx_create(0, None, C.byref(ctx)), "cgp_ctx_create")
    nd, nccl = C.c_int(0), C.c_int(0)
    L.cgp_ctx_info(ctx, C.byref(nd), C.byref(nccl))
    assert nd.value == ngpu and (ngpu == 1 or nccl.value == 1)
    counts = np.array([1000 + 37 * d for d in range(ngpu)], dtype=np.int64)
    bufs = [torch.full((int(counts[d]),), float(d + 1), dtype=torch.float64, device="cuda:%d" % d) for d in range(ngpu)]
    recv = torch.zeros(int(counts.sum()), dtype=torch.float64, device="cuda:0")
    torch.cuda.synchronize()
    ptrs = (C.c_void_p * ngpu)(*[b.data_ptr() for b in bufs])
    _lib.check(L.cgp_ctx_gather_f64(ctx, ptrs, counts.ctypes.data, recv.data_ptr(), 0), "cgp_ctx_gather_f64")
    want = np.concatenate([np.full(int(counts[d]), d + 1.0) for d in range(ngpu)])
    assert np.array_equal(recv.cpu().numpy(), want)
    red = [torch.full((8,), float(d + 1), dtype=torch.float64, device="cuda:%d" % d) for d in range(ngpu)]
    torch.cuda.synchronize()
    rptrs = (C.c_void_p * ngpu)(*[b.data_ptr() for b in red])
    _lib.check(L.cgp_ctx_allreduce_sum_f64(ctx, rptrs, 8), "cgp_ctx_allreduce_sum_f64")
    for d in range(ngpu):
        assert np.array_equal(red[d].cpu().numpy(), np.full(8, ngpu * (ngpu + 1) / 2.0))
    L.cgp_ctx_destroy(ctx)
