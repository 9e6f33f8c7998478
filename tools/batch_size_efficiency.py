"""ns per object of the resident fused step (factor+LL kernel, uniform-grid prediction kernel) vs batch size:
how much a chunk-sized launch loses to kernel tails (explains the end-to-end leg's compute time)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cosmogp_b200.batch import DeviceBatch
out = {}
g = torch.from_numpy(np.linspace(-10, 40, bench.M_GRID)).cuda()
for B in (2048, 4096, 8192, 16384, 32768, 65536, 100000):
    x, y, ye, tmean, ymean = bench.make_c2(B, 2)
    batch = DeviceBatch(x.ravel(), y.ravel(), np.arange(B + 1, dtype=np.int64) * bench.N_EPOCH, y_err=ye.ravel(), dim=1)
    def step():
        fac = batch.factor_dev(bench.HYP, bench.NUGGET, want_ll=True)
        return batch.predict_factored_dev(fac, g, None, None, True, uniform_grid=False)
    for _ in range(3): step()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    reps = 10
    tf = tp = 0.0
    for _ in range(reps):
        e[0].record(); fac = batch.factor_dev(bench.HYP, bench.NUGGET, want_ll=True); e[1].record()
        batch.predict_factored_dev(fac, g, None, None, True, uniform_grid=False); e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tp += e[1].elapsed_time(e[2])
    out[B] = {"factor_ns_per_object": tf / reps * 1e6 / B, "grid_general_ns_per_object": tp / reps * 1e6 / B}
print(json.dumps(out))
