"""Launched under torchrun with 2+ GPUs: the sharded facade must give the same joint fit and the same
predictions as the single-process object (run by rank 0 on the full data)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cosmogp_b200 as cg
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(7)                                       # same data on every rank
sizes = rng.integers(20, 60, 301)
xs = [np.sort(rng.uniform(-10, 40, n)) for n in sizes]
ys = [0.5 * np.sin(x / 3.0 + rng.uniform(0, 6)) + 0.2 * rng.standard_normal(len(x)) for x in xs]
yes = [np.full(len(x), 0.2) for x in xs]
gp = cg.gaussian_process_nobject.sharded(ys, xs, y_err=yes)
gp.find_hyperparameters(hyperparameter_guess=[0.5, 2.0], svd_method=False)
grid = np.linspace(-10, 40, 64)
gp.get_prediction(new_binning=grid, COV='diag')
pred = np.array(gp.gather(gp.Prediction))
if rank == 0:
    ref = cg.gaussian_process_nobject(ys, xs, y_err=yes)
    ref.find_hyperparameters(hyperparameter_guess=[0.5, 2.0], svd_method=False)
    ref.get_prediction(new_binning=grid, COV='diag')
    dh = np.max(np.abs(np.array(gp.hyperparameters) - np.array(ref.hyperparameters)) / np.array(ref.hyperparameters))
    dp = np.max(np.abs(pred - np.array(ref.Prediction)))
    print("ranges", gp.local_range, "hyp", gp.hyperparameters, "rel diff hyp %.2e  max diff pred %.2e  shape %s" % (dh, dp, pred.shape))
    assert pred.shape == (301, 64) and dh < 1e-6 and dp < 1e-6
    print("sharded facade ok")
dist.barrier()
dist.destroy_process_group()
