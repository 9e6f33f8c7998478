"""Wall clock of the numpy-in / numpy-out facade at C2 (10^5 x 60, shared mean) and at C1 (one object of 50 points):
construction, one likelihood evaluation, a full find_hyperparameters, get_prediction.  python tests/perf_facade.py"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import cosmogp_b200 as cg
import torch

def wall(fn, reps=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
x, y, ye, tmean, ymean = bench.make_c2(B, 2)
grid = np.linspace(-10, 40, bench.M_GRID)
res = {"objects": B}
res["construct_s"], gp = wall(lambda: cg.gaussian_process_nobject(y, x, y_err=ye, Mean_Y=ymean, Time_mean=tmean))
gp.compute_log_likelihood([0.5, 2.0], svd_method=False)          # uploads the batch
res["ll_eval_ms"] = wall(lambda: gp.compute_log_likelihood([0.5, 2.0], svd_method=False), 20)[0] * 1e3
res["fit_s"], _ = wall(lambda: gp.find_hyperparameters(hyperparameter_guess=[0.4, 3.0], svd_method=False))
res["fit_hyp"] = [float(v) for v in gp.hyperparameters]
for _ in range(3):                                         # the pinned download buffers are recycled from the third call on
    gp.get_prediction(new_binning=grid, COV='diag', svd_method=False)
res["predict_ms"] = wall(lambda: gp.get_prediction(new_binning=grid, COV='diag', svd_method=False), 5)[0] * 1e3
res["prediction_shape"] = list(np.asarray(gp.Prediction).shape)
# C1: a single light curve of 50 epochs, fit + prediction on 500 points
rng = np.random.default_rng(1)
x1 = np.sort(rng.uniform(-12, 42, 50)); ye1 = rng.uniform(0.03, 0.1, 50)
k = 0.25 * np.exp(-0.5 * (x1[:, None] - x1[None, :]) ** 2 / 64.0) + np.diag(ye1 ** 2 + 0.03 ** 2)
y1 = np.linalg.cholesky(k) @ rng.standard_normal(50)
g1 = cg.gaussian_process(y1, x1, y_err=ye1); g1.nugget = 0.03
g1.compute_log_likelihood([0.5, 8.0], svd_method=False)
res["c1_ll_eval_us"] = wall(lambda: g1.compute_log_likelihood([0.5, 8.0], svd_method=False), 200)[0] * 1e6
res["c1_fit_ms"] = wall(lambda: g1.find_hyperparameters(hyperparameter_guess=[0.5, 8.0], svd_method=False), 3)[0] * 1e3
res["c1_predict_ms"] = wall(lambda: g1.get_prediction(new_binning=np.linspace(-12, 42, 500), COV='diag', svd_method=False), 10)[0] * 1e3
# the same single-object fit by the CPU oracle port (checker; not part of the product path)
from oracle import gp_oracle as O
from scipy.optimize import fmin
t0 = time.perf_counter()
for _ in range(3):
    fmin(lambda h: -O.log_likelihood(y1, x1, h, 0.03, ye1), [0.5, 8.0], disp=False)
res["c1_fit_cpu_port_ms"] = (time.perf_counter() - t0) / 3 * 1e3
print(json.dumps(res))
