"""Do concurrent PCIe copies slow the small-object kernels?  Resident factor+grid step alone vs with
H2D and D2H traffic running on two other streams."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cosmogp_b200 import mean as M
from cosmogp_b200.batch import DeviceBatch
B = 100000
x, y, ye, tmean, ymean = bench.make_c2(B, 2)
off = np.arange(B + 1, dtype=np.int64) * bench.N_EPOCH
y0, d = M.batched_mean(x.ravel(), y.ravel(), off, 1, ymean, tmean, None)
grid = np.linspace(-10, 40, bench.M_GRID)
batch = DeviceBatch(x.ravel(), y.ravel(), off, y0=y0, y_err=ye.ravel(), dim=1)
g = torch.from_numpy(grid).cuda()
n = 64 * 1024 * 1024 // 8
hin = torch.empty(n, dtype=torch.float64, pin_memory=True).fill_(1.0); hout = torch.empty(n, dtype=torch.float64, pin_memory=True)
din = torch.empty(n, dtype=torch.float64, device="cuda"); dout = torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def step():
    fac = batch.factor_dev(bench.HYP, bench.NUGGET, want_ll=True)
    return batch.predict_factored_dev(fac, g, None, None, True)
def timed(traffic, reps=10):
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if traffic:
        for _ in range(40):                       # ~2.7 GB each way: outlasts the timed kernels
            with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
            with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
    e0.record()
    for _ in range(reps): step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print(json.dumps({"factor_plus_grid_ms_alone": timed(False), "with_pcie_traffic_both_ways": timed(True), "alone_again": timed(False)}))
