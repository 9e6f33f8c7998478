"""ns per object of the resident fused step (factor+LL kernel, uniform-grid prediction kernel) vs batch size:
how much a chunk-sized launch loses (explains the end-to-end leg's compute time).  Buffers are allocated once."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cosmogp_b200 import _lib
from cosmogp_b200.batch import DeviceBatch
L = _lib.lib()
out = {}
M = bench.M_GRID
g = torch.from_numpy(np.linspace(-10, 40, M)).cuda()
hyp = np.ascontiguousarray(bench.HYP, dtype=np.float64)
BMAX = 100000
x, y, ye, tmean, ymean = bench.make_c2(BMAX, 2)
full = DeviceBatch(x.ravel(), y.ravel(), np.arange(BMAX + 1, dtype=np.int64) * bench.N_EPOCH, y_err=ye.ravel(), dim=1)
stride = int(L.cgp_factor_ws_doubles(bench.N_EPOCH))
ws = torch.empty(BMAX * stride, dtype=torch.float64, device="cuda")
ll = torch.empty(BMAX, dtype=torch.float64, device="cuda"); info = torch.empty(BMAX, dtype=torch.int32, device="cuda")
mean = torch.empty(BMAX * M, dtype=torch.float64, device="cuda"); var = torch.empty_like(mean)
st = torch.cuda.current_stream().cuda_stream
p = lambda t: t.data_ptr()
for B in (2048, 4096, 8192, 16384, 32768, 65536, 100000):
    def factor():
        _lib.check(L.cgp_factor_batched_dev(B, p(full.off), bench.N_EPOCH, 1, p(full.x), p(full.y), None, p(full.y_err),
                                            _lib.hptr(hyp), bench.NUGGET, 0.0, 0, p(ws), p(ll), p(info), st), "factor")
    def grid(flags):
        _lib.check(L.cgp_predict_factored_dev(B, p(full.off), bench.N_EPOCH, 1, p(full.x), _lib.hptr(hyp), bench.NUGGET, flags,
                                              p(ws), p(info), p(g), None, M, None, p(mean), p(var), st), "grid")
    res = {}
    for name, fn in (("factor", factor), ("grid_general", lambda: grid(0)), ("grid_uniform", lambda: grid(_lib.CGP_GRID_UNIFORM))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        res[name + "_ns_per_object"] = e0.elapsed_time(e1) / reps * 1e6 / B
    out[B] = res
print(json.dumps(out))
