"""Batched 2D objects (e.g. independent PSF exposures of 60 stars each): LL, factor and grid kernels, resident."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cosmogp_b200 import _lib
from cosmogp_b200.batch import DeviceBatch
B, N, M = 100000, 60, 100
rng = np.random.default_rng(0)
xy = rng.uniform(0, 10, (B, N, 2)); z = rng.standard_normal((B, N)); ze = np.full((B, N), 0.2)
grid = rng.uniform(0, 10, (M, 2)); hyp = [1.0, 2.0, 1.5, 0.3]; nug = 0.05
batch = DeviceBatch(xy.reshape(-1, 2), z.ravel(), np.arange(B + 1, dtype=np.int64) * N, y_err=ze.ravel(), dim=2)
g = torch.from_numpy(grid).cuda()
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
fac = batch.factor_dev(hyp, nug)
out = {"objects": B, "points": N, "grid_points": M,
       "ll_ms": timed(lambda: batch.ll_dev(hyp, nug)),
       "factor_ms": timed(lambda: batch.factor_dev(hyp, nug)),
       "grid_ms": timed(lambda: batch.predict_factored_dev(fac, g, None, None, True))}
out["step_objects_per_s"] = B / ((out["ll_ms"] + out["factor_ms"] + out["grid_ms"]) * 1e-3)
print(json.dumps(out))
