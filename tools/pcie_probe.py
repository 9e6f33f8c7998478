"""PCIe ceilings of the box: pinned H2D, D2H, and both at once (the end-to-end leg's bound).
One GPU:   python tools/pcie_probe.py
N GPUs at the same time (what bounds the end-to-end leg of `bench.py --gpus N`):
           python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/pcie_probe.py
Every rank moves 200 MB blocks over its own link concurrently with the others (barrier before and after); rank 0
prints the per-rank and the aggregate GB/s."""
import json, os, time, torch
import torch.distributed as dist
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 200 * 1024 * 1024 // 8
h_in = torch.empty(n, dtype=torch.float64, pin_memory=True).fill_(1.0)
h_out = torch.empty(n, dtype=torch.float64, pin_memory=True)
d_in = torch.empty(n, dtype=torch.float64, device=dev)
d_out = torch.ones(n, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def run(up, down, reps=10, pieces=1):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    k = n // pieces
    for _ in range(reps):
        for p in range(pieces):
            if up:
                with torch.cuda.stream(s1): d_in[p * k:(p + 1) * k].copy_(h_in[p * k:(p + 1) * k], non_blocking=True)
            if down:
                with torch.cuda.stream(s2): h_out[p * k:(p + 1) * k].copy_(d_out[p * k:(p + 1) * k], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    if world > 1:                                   # the slowest rank bounds the step
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return n * 8 / dt / 1e9
out = {"ranks": world, "host_cores": os.cpu_count()}
for name, a in (("h2d", (True, False)), ("d2h", (False, True)), ("both_each_direction", (True, True))):
    run(*a, reps=2)
    per = run(*a)
    out[name + "_GBps_per_rank"] = per
    out[name + "_GBps_aggregate"] = per * world * (2 if name.startswith("both") else 1)
    out[name + "_40pieces_GBps_per_rank"] = run(*a, pieces=40)
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
