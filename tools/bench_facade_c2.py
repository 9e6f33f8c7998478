"""Wall-clock of the numpy-in/numpy-out facade at C2 (10^5 x 60, shared mean, M = 100): what a cosmogp user sees."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import cosmogp_b200 as cg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
x, y, ye, tmean, ymean = bench.make_c2(B, 2)
grid = np.linspace(-10, 40, bench.M_GRID)
out = {"objects": B}
def wall(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); best = min(best, time.perf_counter() - t0)
    return best, r
out["construct_s"], gp = wall(lambda: cg.gaussian_process_nobject(y, x, y_err=ye, Mean_Y=ymean, Time_mean=tmean))
gp.hyperparameters = np.array(bench.HYP); gp.nugget = bench.NUGGET
gp.compute_log_likelihood(bench.HYP)
out["compute_log_likelihood_s"], _ = wall(lambda: gp.compute_log_likelihood(bench.HYP))
out["get_prediction_diag_s"], _ = wall(lambda: gp.get_prediction(new_binning=grid, COV='diag'))
out["get_prediction_mean_only_s"], _ = wall(lambda: gp.get_prediction(new_binning=grid, COV=False))
t0 = time.perf_counter(); gp.find_hyperparameters(hyperparameter_guess=[0.4, 3.0], svd_method=False); out["find_hyperparameters_s"] = time.perf_counter() - t0
out["fit_hyp"] = [float(v) for v in gp.hyperparameters]
print(json.dumps(out))
