"""Generic shared-memory path (65..224 points, one 4-warp CTA per object): LL / predict / LOO timings."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosmogp_b200 import _lib
from cosmogp_b200.batch import DeviceBatch
out = {"dmma_peak_tflops": _lib.fp64_peak(1)}
def ev_time(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
rng = np.random.default_rng(0)
for n, b in ((96, 20000), (128, 20000), (200, 8000), (224, 8000)):
    x = np.sort(rng.uniform(0, n / 2.0, (b, n)), axis=1); y = rng.standard_normal((b, n)); ye = np.full((b, n), 0.2)
    bt = DeviceBatch(x.ravel(), y.ravel(), np.arange(b + 1) * n, y_err=ye.ravel())
    g = torch.from_numpy(np.linspace(0, n / 2.0, 100)).cuda()
    t_ll = ev_time(lambda: bt.ll_dev([0.7, 2.0], 0.0))
    t_pr = ev_time(lambda: bt.predict_dev([0.7, 2.0], 0.0, g, None, None, True))
    fl_ll = n ** 3 / 3 + 3.5 * n * n; fl_pr = n ** 3 / 3 + 4 * n * n + 100 * (n * n + 8 * n)
    out["n%d" % n] = {"objects": b, "ll_ms": t_ll, "ll_tflops": fl_ll * b / t_ll * 1e-9, "predict_ms": t_pr,
                      "predict_tflops": fl_pr * b / t_pr * 1e-9, "ll_objects_per_s": b / t_ll * 1e3}
print(json.dumps(out))
