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
