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
# a batch smaller than the number of ranks: the last rank owns the single object, the others an EMPTY shard; every rank must
# still take part in the exchanges (likelihood total, gather) and end with the same numbers
g1 = cg.gaussian_process_nobject.sharded(ys[:1], xs[:1], y_err=yes[:1])
g1.compute_log_likelihood([0.5, 2.0], svd_method=False)
g1.hyperparameters = np.array([0.5, 2.0])
g1.get_prediction(new_binning=grid, COV='diag')
p1 = g1.gather(g1.Prediction)
root_only = g1.gather(g1.Prediction, root=0)
assert (root_only is None) == (rank != 0)
r1 = cg.gaussian_process_nobject(ys[:1], xs[:1], y_err=yes[:1])
r1.compute_log_likelihood([0.5, 2.0], svd_method=False)
r1.hyperparameters = np.array([0.5, 2.0]); r1.get_prediction(new_binning=grid, COV='diag')
assert abs(g1.log_likelihood[0] - r1.log_likelihood[0]) < 1e-12 * abs(r1.log_likelihood[0]), (g1.log_likelihood, r1.log_likelihood)
assert np.array(p1).shape == (1, 64) and np.max(np.abs(np.array(p1) - np.array(r1.Prediction))) < 1e-12
# objects beyond the shared-memory path on one rank only: every rank must take the large-object branch (global sizes)
big = [np.sort(rng.uniform(0, 100, 300)), np.sort(rng.uniform(0, 100, 40))]
yb = [np.sin(b / 5.0) for b in big]; eb = [np.full(len(b), 0.2) for b in big]
gl = cg.gaussian_process_nobject.sharded(yb, big, y_err=eb)
gl.compute_log_likelihood([0.7, 3.0], svd_method=False)
rl = cg.gaussian_process_nobject(yb, big, y_err=eb); rl.compute_log_likelihood([0.7, 3.0], svd_method=False)
assert abs(gl.log_likelihood[0] - rl.log_likelihood[0]) < 1e-10 * abs(rl.log_likelihood[0]), (gl.log_likelihood, rl.log_likelihood)
if rank == 0:
    print("empty shard and mixed large/small shards ok", g1.local_range, gl.local_range)
dist.barrier()
dist.destroy_process_group()
