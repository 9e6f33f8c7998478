"""End-to-end step (StreamedEvaluator.run, C2) for several chunk counts / stream counts.  python tools/e2e_sweep.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from cosmogp_b200 import mean as M
from cosmogp_b200.batch import StreamedEvaluator
B = 100000
x, y, ye, tmean, ymean = bench.make_c2(B, 2)
off = np.arange(B + 1, dtype=np.int64) * bench.N_EPOCH
y0, d = M.batched_mean(x.ravel(), y.ravel(), off, 1, ymean, tmean, None)
grid = np.linspace(-10, 40, bench.M_GRID)
tmpl = M.template_on_grid(grid, 1, ymean, tmean)
out = {}
for nc, ns in ((4, 4), (6, 6), (8, 8), (12, 8), (16, 8), (24, 8), (8, 4), (8, 2)):
    ev = StreamedEvaluator(B, bench.N_EPOCH, bench.M_GRID, n_chunks=nc, n_streams=ns, shared_mean=True)
    for name, arr in (("x", x), ("y", y), ("y_err", ye), ("template", tmpl), ("diff", d)):
        ev.host(name)[...] = arr
    ev.set_mean_template(tmean, ymean)
    for _ in range(3):
        ev.run(bench.HYP, bench.NUGGET, grid)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(8):
        ev.run(bench.HYP, bench.NUGGET, grid)
    torch.cuda.synchronize()
    out["%d chunks %d streams" % (nc, ns)] = (time.perf_counter() - t0) / 8 * 1e3
    del ev
print(json.dumps(out))
