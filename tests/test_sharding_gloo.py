"""N > 1 host logic on CPU: two gloo ranks shard a ragged batch by sum N^3, all-reduce the
likelihood sum and gather per-object outputs; results must equal the unsharded ones.
(The per-object arithmetic is stood in for by the oracle: sharding must not change it.)"""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from cosmogp_b200 import sharding


def test_balanced_ranges_cover_and_balance():
    rng = np.random.default_rng(0)
    sizes = rng.integers(1, 200, 5000)
    for world in (1, 2, 4, 8):
        r = sharding.balanced_ranges(sizes, world)
        assert r[0][0] == 0 and r[-1][1] == len(sizes)
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        cost = [float((sizes[a:b].astype(float) ** 3).sum()) for a, b in r]
        assert max(cost) < 1.1 * sum(cost) / world + 200.0 ** 3
    assert sharding.balanced_ranges([], 4) == [(0, 0)] * 4
    assert sharding.balanced_ranges([5], 4)[-1][1] == 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import gp_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)                       # same data on every rank
    sizes = rng.integers(3, 40, 23)
    off = np.zeros(len(sizes) + 1, dtype=np.int64); off[1:] = np.cumsum(sizes)
    x = np.concatenate([np.sort(rng.uniform(0, 30, n)) for n in sizes]); y = rng.standard_normal(off[-1])
    ye = rng.uniform(0.1, 0.2, off[-1])
    a, b = sharding.my_range(sizes)
    (xs, ys, yes), loff = sharding.shard_csr([x, y, ye], off, a, b)
    ll = np.array([O.log_likelihood(ys[loff[i]:loff[i + 1]], xs[loff[i]:loff[i + 1]], [0.7, 3.0], 0.05,
                                    yes[loff[i]:loff[i + 1]]) for i in range(b - a)])
    total = sharding.allreduce_sum(ll.sum())
    counts = [r[1] - r[0] for r in sharding.balanced_ranges(sizes, world)]
    allll = sharding.gather_ragged(ll, counts)
    # the final gather on one rank only (what build_pull / C5 uses): point-to-point, nothing padded or replicated
    rooted = sharding.gather_to_root(ll, counts, root=0)
    assert (rooted is None) == (rank != 0)
    if rank == 0:
        assert np.array_equal(rooted, allll)
    t, nb = sharding.allreduce_sums([ll.sum(), float(rank + 1)])
    assert nb == 3.0 and abs(t - total) <= 1e-12 * abs(total)
    import torch
    dev_g = sharding.gather_to_root_dev(torch.from_numpy(ll), counts, root=0)
    assert (dev_g is None) == (rank != 0) and (rank != 0 or np.array_equal(dev_g.numpy(), allll))
    if rank == 0:
        ref = np.array([O.log_likelihood(y[off[i]:off[i + 1]], x[off[i]:off[i + 1]], [0.7, 3.0], 0.05,
                                         ye[off[i]:off[i + 1]]) for i in range(len(sizes))])
        q.put((total, ref.sum(), np.abs(allll - ref).max(), counts))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, ref, err, counts = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(counts) == 23 and min(counts) > 0
    assert abs(total - ref) < 1e-9 * abs(ref)
    assert err == 0.0
