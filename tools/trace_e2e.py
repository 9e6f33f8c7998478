"""Timeline of one end-to-end step (CGP_STREAM_TRACE): python tools/trace_e2e.py [n_chunks] [n_streams]"""
import os, sys
os.environ["CGP_STREAM_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from cosmogp_b200 import mean as M
from cosmogp_b200.batch import StreamedEvaluator
B = 100000
nc = int(sys.argv[1]) if len(sys.argv) > 1 else 40
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 10
x, y, ye, tmean, ymean = bench.make_c2(B, 2)
off = np.arange(B + 1, dtype=np.int64) * bench.N_EPOCH
y0, d = M.batched_mean(x.ravel(), y.ravel(), off, 1, ymean, tmean, None)
grid = np.linspace(-10, 40, bench.M_GRID)
tmpl = M.template_on_grid(grid, 1, ymean, tmean)
ev = StreamedEvaluator(B, bench.N_EPOCH, bench.M_GRID, n_chunks=nc, n_streams=ns, shared_mean=True)
for name, arr in (("x", x), ("y", y), ("y_err", ye), ("template", tmpl), ("diff", d)):
    ev.host(name)[...] = arr
ev.set_mean_template(tmean, ymean)          # y0 from the template spline, on the device
for _ in range(3):
    sys.stderr.write("--- run\n")
    ev.run(bench.HYP, bench.NUGGET, grid)
