"""Timings of the BASELINE configs other than the bench.py workload: C1 (single curve, N=50,
M=500) and C5 (leave-one-out pulls, N=40) on one B200, with the CPU port beside them.
   python tests/perf_configs.py [--c5 1000000]"""
import argparse, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cosmogp_b200 as cg
from cosmogp_b200 import _lib
from cosmogp_b200.batch import DeviceBatch
from oracle import gp_oracle as O            # CPU baseline leg only

ap = argparse.ArgumentParser(); ap.add_argument("--c5", type=int, default=1000000); args = ap.parse_args()
out = {"dmma_peak_tflops": _lib.fp64_peak(1)}

def ev_time(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best

# ---- C1
rng = np.random.default_rng(1)
n, m = 50, 500
x = np.sort(rng.uniform(-12, 42, n)); ye = rng.uniform(0.03, 0.1, n); hyp = [0.5, 8.0]; nug = 0.03
y = np.linalg.cholesky(O.rbf_1d(x, hyp, nugget=nug, y_err=ye)) @ rng.standard_normal(n)
grid = np.linspace(-12, 42, m)
gp = cg.gaussian_process(y, x, y_err=ye); gp.hyperparameters = np.array(hyp); gp.nugget = nug
gp.get_prediction(new_binning=grid, COV='diag')
t0 = time.perf_counter()
for _ in range(20): gp.get_prediction(new_binning=grid, COV='diag')
wall_pred = (time.perf_counter() - t0) / 20
gw = cg.gaussian_process(y, x, y_err=ye); gw.find_hyperparameters([0.5, 8.0], svd_method=False)      # warm: module load, pinned buffers
t0 = time.perf_counter(); gf = cg.gaussian_process(y, x, y_err=ye); gf.find_hyperparameters([0.5, 8.0], svd_method=False); wall_fit = time.perf_counter() - t0
b1 = gp.batch; g_dev = torch.from_numpy(grid).cuda()
k_ll = ev_time(lambda: b1.ll_dev(hyp, nug)); k_pr = ev_time(lambda: b1.predict_dev(hyp, nug, g_dev, None, None, True))
t0 = time.perf_counter()
for _ in range(20): O.predict(y, x, hyp, nug, grid, ye, full_cov=True)
cpu_pred = (time.perf_counter() - t0) / 20
t0 = time.perf_counter()
for _ in range(200): O.log_likelihood(y, x, hyp, nug, ye)
cpu_ll = (time.perf_counter() - t0) / 200
out["c1"] = {"n": n, "m": m, "facade_predict_wall_ms": wall_pred * 1e3, "facade_fit_wall_ms": wall_fit * 1e3,
             "ll_kernel_us": k_ll * 1e3, "predict_kernel_us": k_pr * 1e3, "cpu_port_predict_fullcov_ms": cpu_pred * 1e3,
             "cpu_port_ll_us": cpu_ll * 1e6, "fit_hyp": [float(v) for v in gf.hyperparameters]}

# ---- C5
b, n = args.c5, 40
rng = np.random.default_rng(5)
x = np.sort(rng.uniform(-10, 10, (b, n)), axis=1); ye = np.full((b, n), 0.1)
y = 0.5 * np.sin(x / 2.0 + rng.uniform(0, 6.28, (b, 1))) + 0.1 * rng.standard_normal((b, n))
hyp, nug = [0.5, 2.0], 0.0
batch = DeviceBatch(x.ravel(), y.ravel(), np.arange(b + 1, dtype=np.int64) * n, y_err=ye.ravel())
outs = [torch.empty(b * n, dtype=torch.float64, device="cuda") for _ in range(4)]
h = np.ascontiguousarray(hyp, dtype=np.float64); st = torch.cuda.current_stream().cuda_stream
def loo():
    _lib.check(_lib.lib().cgp_loo_batched_dev(b, batch.off.data_ptr(), n, 1, batch.x.data_ptr(), batch.y.data_ptr(), None,
               batch.y_err.data_ptr(), h.ctypes.data, nug, 0.0, 0, 0, *[o.data_ptr() for o in outs], batch._info.data_ptr(), st), "loo")
k_loo = ev_time(loo, reps=3)
cg.build_pull(y, x, hyp, nugget=nug, y_err=ye).compute_pull(svd_method=False)          # warm: the allocator's first 2 GB
torch.cuda.synchronize()
t0 = time.perf_counter(); bp = cg.build_pull(y, x, hyp, nugget=nug, y_err=ye); bp.compute_pull(svd_method=False); wall = time.perf_counter() - t0
ns = 48
t0 = time.perf_counter()
for i in range(ns): O.loo_bruteforce(y[i], x[i], hyp, nug, ye[i])
cpu_bf = (time.perf_counter() - t0) / ns
fl = 2.0 * n ** 3 / 3.0 + 6.0 * n ** 2
out["c5"] = {"objects": b, "n": n, "loo_kernel_ms": k_loo, "objects_per_s_kernel": b / k_loo * 1e3,
             "achieved_tflops": fl * b / k_loo * 1e-9, "build_pull_wall_s": wall, "pull_average": bp.pull_average, "pull_std": bp.pull_std,
             "cpu_port_bruteforce_ms_per_object_1core": cpu_bf * 1e3, "cpu_sample": "%d objects, reference algorithm (N refits), 1 core" % ns}
print(json.dumps(out))
