import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
import cosmogp_b200 as cg
from cosmogp_b200 import _lib
B = 100000
x, y, ye, tmean, ymean = bench.make_c2(B, 2)
grid = np.linspace(-10, 40, bench.M_GRID)
gp = cg.gaussian_process_nobject(y, x, y_err=ye, Mean_Y=ymean, Time_mean=tmean)
gp.hyperparameters = np.array([0.5, 2.0])
for _ in range(3):
    gp.get_prediction(new_binning=grid, COV='diag', svd_method=False)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    gp.get_prediction(new_binning=grid, COV='diag', svd_method=False)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
