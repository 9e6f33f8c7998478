"""PCIe ceilings of the box: pinned H2D, D2H, and both at once (the end-to-end leg's bound)."""
import json, time, torch
dev = torch.device("cuda", 0)
n = 200 * 1024 * 1024 // 8
h_in = torch.empty(n, dtype=torch.float64, pin_memory=True).fill_(1.0)
h_out = torch.empty(n, dtype=torch.float64, pin_memory=True)
d_in = torch.empty(n, dtype=torch.float64, device=dev)
d_out = torch.ones(n, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def run(up, down, reps=10, pieces=1):
    torch.cuda.synchronize()
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
    return n * 8 / dt / 1e9
out = {}
for name, a in (("h2d", (True, False)), ("d2h", (False, True)), ("both_each_direction", (True, True))):
    run(*a, reps=2)
    out[name + "_GBps"] = run(*a)
    out[name + "_40pieces_GBps"] = run(*a, pieces=40)
print(json.dumps(out))
