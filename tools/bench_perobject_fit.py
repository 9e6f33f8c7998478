"""10^5 per-object hyperparameter fits: Nelder-Mead with the simplices on the device vs on the host."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import cosmogp_b200 as cg
from cosmogp_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
x, y, ye, tmean, ymean = bench.make_c2(B, 2)
gp = cg.gaussian_process_nobject(y, x, y_err=ye, Mean_Y=ymean, Time_mean=tmean)
gp.batch
out = {"objects": B}
for opt in ("device", "host"):
    for rep in range(2):                      # second pass: warm pools and buffers
        l0 = _lib.lib().cgp_launch_count()
        t0 = time.perf_counter()
        gp.find_hyperparameters_per_object(hyperparameter_guess=[0.5, 2.0], optimizer=opt)
        dt = time.perf_counter() - t0
    h = gp.hyperparameters_per_object
    out[opt] = {"wall_s": dt, "fits_per_s": B / dt, "launches": int(_lib.lib().cgp_launch_count() - l0),
                "median_hyp": [float(np.median(h[:, 0])), float(np.median(h[:, 1]))],
                "mean_iterations": float(gp.fit_iterations.mean()), "mean_evaluations": float(gp.fit_evaluations.mean()),
                "max_evaluations": int(gp.fit_evaluations.max())}
print(json.dumps(out))
