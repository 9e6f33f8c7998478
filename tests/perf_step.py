"""Resident C2 step, kernel variants side by side (CUDA events): separate LL kernel, factor(+LL) + grid kernel,
and the single fused kernel (cgp_step_batched_dev).  python tests/perf_step.py [objects] [reps]"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from cosmogp_b200 import _lib, mean as M
from cosmogp_b200.batch import DeviceBatch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
x, y, ye, tmean, ymean = bench.make_c2(B, 2)
off = np.arange(B + 1, dtype=np.int64) * bench.N_EPOCH
y0, d = M.batched_mean(x.ravel(), y.ravel(), off, 1, ymean, tmean, None)
grid = np.linspace(-10, 40, bench.M_GRID)
tmpl = M.template_on_grid(grid, 1, ymean, tmean)
batch = DeviceBatch(x.ravel(), y.ravel(), off, y0=y0, y_err=ye.ravel(), dim=1)
g = torch.from_numpy(grid).cuda(); ny0 = torch.from_numpy(np.concatenate([tmpl, d])).cuda()
HYP, NUG = bench.HYP, bench.NUGGET


def timed(fn):
    for _ in range(3):
        out = fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def split():
    fac = batch.factor_dev(HYP, NUG, want_ll=True)
    m, v, _ = batch.predict_factored_dev(fac, g, None, ny0, True, template_mean=True, uniform_grid=True)
    return fac["ll"], m, v

res = {}
res["ll_ms"], (ll0, _) = timed(lambda: batch.ll_dev(HYP, NUG))
ll0 = ll0.clone()
res["split_ms"], s = timed(split)
s = [t.clone() for t in s]
res["fused_ms"], f = timed(lambda: batch.step_dev(HYP, NUG, g, None, ny0, True, template_mean=True, uniform_grid=True))
rel = lambda a, b: float(((a - b).abs() / b.abs()).max())
res["fused_vs_split"] = [rel(f[0], s[0]), rel(f[1], s[1]), rel(f[2], s[2])]
res["fused_ll_vs_ll_kernel"] = rel(f[0], ll0)
res["objects"] = B
# never report a time for wrong numbers: the first objects against the numpy oracle (checker)
from oracle import gp_oracle as O
k = 16
y0m = y0.reshape(B, bench.N_EPOCH)
ll_o = O.ll_batched_1d(x[:k], y[:k], y0m[:k], ye[:k], HYP, NUG)
mo, vo = O.predict_batched_1d(x[:k], y[:k], y0m[:k], ye[:k], HYP, NUG, grid, tmpl[None, :] + d[:k, None])
npy = lambda t, n: t[:n].cpu().numpy()
res["parity_vs_oracle"] = max(float(np.max(np.abs(npy(f[0], k) - ll_o) / np.abs(ll_o))),
                              float(np.max(np.abs(npy(f[1], k * bench.M_GRID).reshape(k, -1) - mo) / np.abs(mo))),
                              float(np.max(np.abs(npy(f[2], k * bench.M_GRID).reshape(k, -1) - vo) / np.abs(vo))),
                              float(np.max(np.abs(npy(ll0, k) - ll_o) / np.abs(ll_o))))
assert res["parity_vs_oracle"] < 1e-9, res
print(json.dumps(res))
