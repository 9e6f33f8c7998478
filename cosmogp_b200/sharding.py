"""Sharding of independent objects over the GPUs of one box (one process per GPU,
torch.distributed; NCCL on GPUs, gloo in the CPU tests).

Objects never interact except through the scalar sum of their log-likelihoods
(cosmogp/Gaussian_process.py:205-213), so the data path has NO collective: each rank owns a
contiguous range of objects, balanced by sum N^3 (the factorisation cost), and keeps its
inputs and outputs local.  Exchanges: one all-reduced double per likelihood evaluation and
one final gather of the per-object outputs.
"""
import numpy as np
import torch
import torch.distributed as dist


def balanced_ranges(sizes, world):
    """Contiguous ranges [start, stop) per rank with near-equal sum of N^3 (at least N for empty-cost safety)."""
    sizes = np.asarray(sizes, dtype=np.float64)
    cost = np.cumsum(sizes ** 3 + 1.0)
    total = cost[-1] if len(cost) else 0.0
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cost, total * r / world, side="left")) if len(cost) else 0)
    cuts.append(len(sizes))
    cuts = np.maximum.accumulate(cuts)
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def my_range(sizes, rank=None, world=None):
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    return balanced_ranges(sizes, world)[rank]


def shard_csr(arrays, off, start, stop):
    """Slice flat per-point arrays (and their CSR offsets) to objects [start, stop)."""
    o0, o1 = int(off[start]), int(off[stop])
    return [None if a is None else a[o0:o1] for a in arrays], np.asarray(off[start:stop + 1]) - o0


def allreduce_sum(value, device=None):
    """Sum of one double over ranks, in rank order on every rank (deterministic: gather, then add)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    total = 0.0
    for p in parts:
        total += float(p.item())
    return total


def allreduce_sums(values, device=None):
    """allreduce_sum for a few doubles at once (one exchange): -> list of sums, identical on every rank."""
    vals = [float(v) for v in values]
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return vals
    t = torch.tensor(vals, dtype=torch.float64, device=device or "cpu")
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    rows = torch.stack(parts).cpu().numpy()
    out = [0.0] * len(vals)
    for r in rows:                                          # rank order
        for k in range(len(vals)):
            out[k] += float(r[k])
    return out


def gather_to_root(local, counts, root=0, device=None):
    """Final gather of per-rank 1-D float64 arrays (lengths `counts`, known everywhere) on `root` only: every other
    rank sends its own slice once (point-to-point over NVLink with NCCL), nothing is padded or replicated.
    -> the concatenation on root, None elsewhere."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return np.asarray(local)
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device or "cpu"
    mine = torch.as_tensor(np.ascontiguousarray(local, dtype=np.float64)).to(dev)
    assert mine.numel() == int(counts[rank]), "rank %d: %d values, expected %d" % (rank, mine.numel(), counts[rank])
    if rank != root:
        if mine.numel():
            dist.send(mine, dst=root)
        return None
    starts = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    out = torch.empty(int(starts[-1]), dtype=torch.float64, device=dev)
    out[starts[root]:starts[root + 1]] = mine
    for r in range(world):
        if r != root and counts[r]:
            dist.recv(out[starts[r]:starts[r + 1]], src=r)
    return out.cpu().numpy()


def gather_to_root_dev(local, counts, root=0):
    """gather_to_root for a DEVICE tensor, result left on the root's device (None elsewhere): the final gather of
    per-object outputs over NVLink (NCCL point-to-point; every rank sends its slice exactly once)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    assert local.numel() == int(counts[rank])
    if rank != root:
        if local.numel():
            dist.send(local.contiguous(), dst=root)
        return None
    starts = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    out = torch.empty(int(starts[-1]), dtype=local.dtype, device=local.device)
    out[starts[root]:starts[root + 1]] = local
    for r in range(world):
        if r != root and counts[r]:
            dist.recv(out[starts[r]:starts[r + 1]], src=r)
    return out


def gather_ragged(local, counts, device=None):
    """Concatenate per-rank 1-D float64 arrays (lengths `counts`, known everywhere) on every rank."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return np.asarray(local)
    world = dist.get_world_size()
    width = int(max(counts)) if len(counts) else 0
    buf = torch.zeros(width, dtype=torch.float64, device=device or "cpu")
    buf[:len(local)] = torch.as_tensor(np.asarray(local, dtype=np.float64)).to(buf.device)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return np.concatenate([parts[r][:int(counts[r])].cpu().numpy() for r in range(world)])
