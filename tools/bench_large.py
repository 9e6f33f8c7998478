"""Timing of the large-object configs (BASELINE configs 3 and 4) on one B200.
   python tools/bench_large.py [--n4 20000] [--m3 100000]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosmogp_b200 import _lib, dense


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


ap = argparse.ArgumentParser()
ap.add_argument("--n4", type=int, default=20000)
ap.add_argument("--n3", type=int, default=2000)
ap.add_argument("--m3", type=int, default=100000)
args = ap.parse_args()
out = {"dmma_peak_tflops": _lib.fp64_peak(1)}
st = torch.cuda.current_stream().cuda_stream
L = _lib.lib()

# raw GEMM rate (the trailing-update kernel), 8192 x 8192 x 128 and x 2048
for k in (128, 2048):
    m = 8192
    a = torch.randn(m, k, dtype=torch.float64, device="cuda"); c = torch.zeros(m, m, dtype=torch.float64, device="cuda")
    ms = timed(lambda: L.cgp_gemm_nt_dev(a.data_ptr(), k, a.data_ptr(), k, c.data_ptr(), m, m, m, k, -1.0, 1.0, 0, st))
    out["gemm_nt_8192x8192x%d_tflops" % k] = 2.0 * m * m * k / ms * 1e-9
    del a, c

# C4: N = 20000 2D, build + factor + solve
rng = np.random.default_rng(4)
n = args.n4
x = rng.uniform(0, 1000, (n, 2)); ye = np.full(n, 0.3)
y = np.cos(x[:, 0] / 100) * np.sin(x[:, 1] / 80) + 0.3 * rng.standard_normal(n)
obj = dense.LargeObject(x, y, ye, None, dim=2)
hyp = [1.0, 30.0, 30.0, 0.0]
t0 = time.perf_counter(); ll = obj.factor(hyp, 0.0); torch.cuda.synchronize(); wall = time.perf_counter() - t0
h = np.ascontiguousarray(hyp, dtype=np.float64)
scal = torch.zeros(2, dtype=torch.float64, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
def build():
    L.cgp_cov_matrix_dev(2, obj.x.data_ptr(), n, None, 0, obj.y_err.data_ptr(), h.ctypes.data, 0.0, 0.0, 0, obj.a.data_ptr(), obj.n_pad, obj.n_pad, obj.n_pad, st)
def factor():
    build(); L.cgp_potrf_dev(obj.a.data_ptr(), obj.n_pad, obj.n_pad, scal.data_ptr(), info.data_ptr(), st)
ms_build = timed(build)
ms_fac = timed(factor, reps=2) - ms_build
ms_solve = timed(lambda: L.cgp_large_solve_dev(obj.a.data_ptr(), n, obj.n_pad, obj.n_pad, obj.y.data_ptr(), None, obj.alpha.data_ptr(), scal[1:].data_ptr(), st))
out["c4"] = {"n": n, "ll": ll, "build_ms": ms_build, "potrf_ms": ms_fac, "solve_ms": ms_solve,
             "potrf_tflops": n ** 3 / 3.0 / ms_fac * 1e-9, "first_call_wall_s": wall,
             "build_gbs": obj.n_pad ** 2 * 8 / ms_build * 1e-6}
del obj
torch.cuda.empty_cache()

# C3: N = 2000 stars, predict on 10^5 grid points
rng = np.random.default_rng(3)
n, m = args.n3, args.m3
x = rng.uniform(-200, 200, (n, 2)); ye = np.full(n, 0.2)
y = np.cos(x[:, 0] / 60) * np.sin(x[:, 1] / 45) + 0.2 * rng.standard_normal(n)
grid = rng.uniform(-200, 200, (m, 2))
obj = dense.LargeObject(x, y, ye, None, dim=2)
hyp = [1.0, 30.0, 25.0, 50.0]
ms_fac = timed(lambda: obj.factor(hyp, 0.0))
g = torch.from_numpy(grid).cuda(); mean = torch.empty(m, dtype=torch.float64, device="cuda"); var = torch.empty_like(mean)
chunk = 131072 if m > 100000 else (m + 127) // 128 * 128
vwork = torch.empty((chunk, obj.n_pad), dtype=torch.float64, device="cuda")
h = np.ascontiguousarray(hyp, dtype=np.float64)
ms_mean = timed(lambda: L.cgp_large_predict_dev(obj.a.data_ptr(), n, obj.n_pad, obj.n_pad, 2, obj.x.data_ptr(), obj.alpha.data_ptr(), h.ctypes.data, 0.0, 0, g.data_ptr(), m, None, mean.data_ptr(), None, None, 0, st))
ms_all = timed(lambda: L.cgp_large_predict_dev(obj.a.data_ptr(), n, obj.n_pad, obj.n_pad, 2, obj.x.data_ptr(), obj.alpha.data_ptr(), h.ctypes.data, 0.0, 0, g.data_ptr(), m, None, mean.data_ptr(), var.data_ptr(), vwork.data_ptr(), chunk, st))
out["c3"] = {"n": n, "m": m, "factor_incl_host_sync_ms": ms_fac, "mean_ms": ms_mean, "mean_var_ms": ms_all,
             "var_tflops": (m * float(n) ** 2) / (ms_all - ms_mean) * 1e-9, "mean_gexp_per_s": m * n / ms_mean * 1e-6}
print(json.dumps(out))
