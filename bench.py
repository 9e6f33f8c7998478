#!/usr/bin/env python
"""Benchmark of the cosmogp GP hot path on B200 (contract: see the task statement).

Workload (BASELINE.json configs[1], "C2"): 10^5 synthetic light curves x 60 epochs with a
shared mean function.  One STEP = one pass of the hot path over the batch:
  one log-likelihood evaluation of all objects (what scipy's optimiser triggers per
  simplex point, cosmogp/Gaussian_process.py:191-213) + one prediction of mean and
  variance on a shared 100-point grid (:270-361; grid of the multi-object notebook cell 22).
metric = objects put through LL+predict per second ("fits/s"), whole job.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--objects B]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_EPOCH = 60
M_GRID = 100
HYP = (0.5, 2.0)
NUGGET = 0.0
YERR = 0.2


def flops_ll(n):            # build + POTRF + forward solve + norms
    return n ** 3 / 3.0 + 3.5 * n ** 2


def flops_predict(n, m):    # build + POTRF + 2 TRSV + per grid point (h build, mean dot, v = L^-1 h, |v|^2)
    return n ** 3 / 3.0 + 4.0 * n ** 2 + m * (n ** 2 + 8.0 * n)


def make_c2(n_obj, seed):
    """SURVEY.md section 8(d) C2 recipe: x = sort(U(-10,40)), truth sigma=0.5 l=2, y_err=0.2,
    template -18+2 sin(t/10) on linspace(-15,45,61), per-object offset N(0,0.3), y = mean + offset + L z."""
    rng = np.random.default_rng(seed)
    tmean = np.linspace(-15, 45, 61)
    ymean = -18 + 2 * np.sin(tmean / 10)
    x = np.empty((n_obj, N_EPOCH)); y = np.empty((n_obj, N_EPOCH))
    from scipy.interpolate import InterpolatedUnivariateSpline
    spline = InterpolatedUnivariateSpline(tmean, ymean)
    eye = np.eye(N_EPOCH)
    for s in range(0, n_obj, 10000):
        e = min(n_obj, s + 10000)
        xs = np.sort(rng.uniform(-10, 40, (e - s, N_EPOCH)), axis=1)
        d = xs[:, None, :] - xs[:, :, None]
        k = HYP[0] ** 2 * np.exp(-0.5 * d * d / HYP[1] ** 2) + YERR ** 2 * eye
        low = np.linalg.cholesky(k)
        z = rng.standard_normal((e - s, N_EPOCH, 1))
        x[s:e] = xs
        y[s:e] = spline(xs) + rng.normal(0, 0.3, (e - s, 1)) + (low @ z)[:, :, 0]
    ye = np.full((n_obj, N_EPOCH), YERR)
    return x, y, ye, tmean, ymean


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    """One host core: the oracle port of the reference per-object path
    (log_likelihood_gp with cholesky_inverse + get_prediction mean/variance diagonal)."""
    x, y, y0, ye, grid, ny0 = args
    from threadpoolctl import threadpool_limits
    from oracle import gp_oracle as O
    with threadpool_limits(limits=1):
        t0 = time.perf_counter()
        acc = 0.0
        for i in range(len(x)):
            acc += O.log_likelihood(y[i], x[i], HYP, NUGGET, ye[i], y0[i])
            m, v = O.predict(y[i], x[i], HYP, NUGGET, grid, ye[i], y0[i], ny0[i], full_cov=False)
            acc += m[0] + v[0]
        return time.perf_counter() - t0, acc


def cpu_pass(x, y, y0, ye, grid, ny0, pool, cores):
    """All host cores, one process each, objects split evenly.  Returns objects/s."""
    parts = np.array_split(np.arange(len(x)), cores)
    t0 = time.perf_counter()
    pool.map(_cpu_worker, [(x[p], y[p], y0[p], ye[p], grid, ny0[p]) for p in parts])
    return len(x) / (time.perf_counter() - t0)


def run_reference(args, rank, world):
    if rank != 0:
        return
    import multiprocessing as mp
    from cosmogp_b200 import mean as M
    cores = os.cpu_count() or 1
    per_step = args.cpu_objects or 400 * cores
    x, y, ye, tmean, ymean = make_c2(per_step, 2)
    off = np.arange(per_step + 1, dtype=np.int64) * N_EPOCH
    y0, d = M.batched_mean(x.ravel(), y.ravel(), off, 1, ymean, tmean, None)
    y0 = y0.reshape(per_step, N_EPOCH)
    grid = np.linspace(-10, 40, M_GRID)
    ny0 = M.template_on_grid(grid, 1, ymean, tmean)[None, :] + d[:, None]
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_pass(x, y, y0, ye, grid, ny0, pool, cores)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_pass(x, y, y0, ye, grid, ny0, pool, cores)
        dt = time.perf_counter() - t0
    val = per_step * args.steps / dt
    sample = "%d of the 10^5 objects per step (N=%d, M=%d), oracle port, %d processes x 1 BLAS thread" % (
        per_step, N_EPOCH, M_GRID, cores)
    print(json.dumps({
        "impl": "reference", "metric": "gp_fits_per_sec", "value": val, "unit": "objects/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(per_step, 1),
        "cpu_baseline": {"value": val, "unit": "objects/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "objects/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def workload_config(n_obj, world):
    return {"workload": "C2: batched 1D light curves, shared mean; step = 1 LL evaluation + predict(mean,var) on shared grid",
            "objects_per_gpu": n_obj, "epochs": N_EPOCH, "grid_points": M_GRID, "kernel": "RBF1D",
            "hyp": list(HYP), "nugget": NUGGET, "y_err": YERR, "sharding": "objects x%d, no data-path collective "
            "(LL sum all-reduced)" % world, "l2_policy": "inputs+outputs per step (%.0f MB) exceed the 126 MB L2"
            % (n_obj * (4 * N_EPOCH + 3 * M_GRID) * 8 / 1e6)}


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--objects", type=int, default=100000, help="objects per GPU")
    ap.add_argument("--cpu-objects", type=int, default=0, help="objects per CPU baseline pass")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from cosmogp_b200 import _lib, mean as M
    from cosmogp_b200.batch import DeviceBatch, pinned_like

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.require_device()

    B = args.objects
    x, y, ye, tmean, ymean = make_c2(B, 2 + rank)
    off = np.arange(B + 1, dtype=np.int64) * N_EPOCH
    y0, d = M.batched_mean(x.ravel(), y.ravel(), off, 1, ymean, tmean, None)      # host prep, outside the path
    grid = np.linspace(-10, 40, M_GRID)
    # shared mean: template on the grid + one offset per object (the reference's Mean_Y / diff, mean.py:92-101);
    # the device adds them, so M + B doubles travel instead of B x M
    tmpl = M.template_on_grid(grid, 1, ymean, tmean)
    ny0 = tmpl[None, :] + d[:, None]                                    # the same thing materialised, for the CPU leg
    packed_mean = np.concatenate([tmpl, d])
    # pinned host staging (what a caller that cares about PCIe hands us)
    (xp, _k1), (yp, _k2), (y0p, _k3), (yep, _k4), (ny0p, _k5) = (pinned_like(a) for a in (x.ravel(), y.ravel(), y0, ye.ravel(), packed_mean))

    batch = DeviceBatch(xp, yp, off, y0=y0p, y_err=yep, dim=1)
    g_dev = torch.from_numpy(grid).to(dev)
    ny0_dev = torch.from_numpy(ny0p).to(dev)
    peak_dmma = _lib.fp64_peak(1)
    peak_dfma = _lib.fp64_peak(0)

    def step():
        ll, info = batch.ll_dev(HYP, NUGGET)
        if world > 1:
            tot = ll.sum()
            dist.all_reduce(tot)          # the one exchange a likelihood evaluation needs
        e_mid.record()
        # prediction = factorisation kernel + grid kernel (what cgp_predict_batched_dev runs internally
        # for large batches; called as two entry points here so that each kernel is timed on its own)
        fac = batch.factor_dev(HYP, NUGGET)
        e_mid2.record()
        mean, var, _ = batch.predict_factored_dev(fac, g_dev, None, ny0_dev, True, template_mean=True, uniform_grid=True)
        return ll, mean, var

    e_mid = torch.cuda.Event(enable_timing=True)
    e_mid2 = torch.cuda.Event(enable_timing=True)
    for _ in range(args.warmup):
        out = step()
    torch.cuda.synchronize()
    ll_first = out[0][:64].cpu().numpy()         # checked against the CPU leg's oracle values below

    sampler = ClockSampler(local); sampler.start()
    time.sleep(0.3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    launches0 = _lib.lib().cgp_launch_count()
    for k in range(args.steps):
        ev[k][0].record()
        e_mid, e_mid2 = ev[k][1], ev[k][2]
        step()
        ev[k][3].record()
    torch.cuda.synchronize()
    launches = _lib.lib().cgp_launch_count() - launches0
    if world > 1:
        dist.barrier()
    total_ms = ev[0][0].elapsed_time(ev[-1][3])
    ll_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    fa_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    pr_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in ev]))
    clocks = sampler.stop()
    # the same outputs from ONE factorisation per object (the factor kernel also emits the likelihood): what the
    # end-to-end path runs; reported beside the step, whose separate LL launch is what a fit repeats ~60 times
    def fused_step():
        fac = batch.factor_dev(HYP, NUGGET, want_ll=True)
        return batch.predict_factored_dev(fac, g_dev, None, ny0_dev, True, template_mean=True, uniform_grid=True)
    fused_step(); fused_step()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    f0.record()
    for _ in range(max(3, min(args.steps, 10))):
        fused_step()
    f1.record()
    torch.cuda.synchronize()
    fused_ms = f0.elapsed_time(f1) / max(3, min(args.steps, 10))
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())

    # ---- end to end through the numpy-in/numpy-out layer: every step uploads the step's inputs from
    # pinned host memory, runs LL + predict and downloads ll/mean/var; chunks of objects (ramping up from
    # 2048 to a quarter of the batch and down again) are pipelined: one upload stream, one download stream,
    # kernels of consecutive chunks on 8 compute streams -- one native call per step (cgp_streamer_run)
    from cosmogp_b200.batch import StreamedEvaluator
    ev_e2e = StreamedEvaluator(B, N_EPOCH, M_GRID, dim=1, n_chunks=int(os.environ.get("CGP_E2E_CHUNKS", "8")), n_streams=int(os.environ.get("CGP_E2E_STREAMS", "8")), shared_mean=True)
    for name, arr in (("x", x), ("y", y), ("y_err", ye), ("template", tmpl), ("diff", d)):
        ev_e2e.host(name)[...] = arr
    # the mean at the epochs (template spline + offset, cosmogp/mean.py:84-90) is evaluated on the device from the
    # template itself -- the reference's own inputs (Mean_Y, Time_mean, diff) -- instead of uploading y0
    ev_e2e.set_mean_template(tmean, ymean)

    def e2e_step():
        tot, ll_h, mean, var, info = ev_e2e.run(HYP, NUGGET, grid)
        return ev_e2e.h2d_bytes, ev_e2e.d2h_bytes, tot
    e2e_steps = max(3, min(args.steps, 10))
    e2e_step(); e2e_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h2d, d2h, _ = e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    traffic = None                      # dram bytes per launch of the predict kernel, from the committed ncu capture
    try:
        if B == 100000:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["predict_f"]["traffic"]
    except Exception:
        pass
    value = B * world * args.steps / (total_ms * 1e-3)
    # dominant kernel = the grid kernel: per grid point h build + mean dot + v = L^-1 h + |v|^2
    fl_grid = M_GRID * (N_EPOCH ** 2 + 8.0 * N_EPOCH)
    fl_factor = flops_predict(N_EPOCH, M_GRID) - fl_grid
    achieved = fl_grid * B / (pr_ms * 1e-3) * 1e-12
    line = {
        "metric": "gp_fits_per_sec", "value": value, "unit": "objects/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(B, world),
        "clocks": clocks,
        "e2e": {"value": B * world * e2e_steps / e2e_s, "unit": "objects/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s / e2e_steps * 1e3, "steps": e2e_steps,
                "path": "StreamedEvaluator.run -> cgp_streamer_run: pinned host in, pinned host out; per chunk one "
                        "factorisation serves the likelihood and the prediction (the timed resident step above "
                        "keeps the separate LL kernel)"},
        "gpu_launches": int(launches),
        "fused_step": {"ms_per_step": fused_ms, "objects_per_s_per_gpu": B / (fused_ms * 1e-3),
                       "what": "LL + predict from one factorisation per object (factor kernel emits LL), resident"},
        "roofline": {"bound": "tensor", "kernel": "gp64_kernel<1,PREDICT_FU,8>: predictive mean+variance on the (uniform) grid "
                     "from the TMA-staged factor (FP64 tensor pipe, DMMA.8x8x4)", "achieved": achieved,
                     "peak": peak_dmma, "unit": "TFLOP/s", "frac": achieved / peak_dmma, "traffic": traffic,
                     "peak_source": "FP64 DMMA m8n8k4 ceiling measured in this run (cgp_fp64_peak); MEASURED_PEAKS.json "
                                    "has no FP64 entry; DFMA ceiling %.2f" % peak_dfma,
                     "flop_per_object": fl_grid, "ms_per_launch": pr_ms,
                     "factor_kernel": {"ms_per_launch": fa_ms, "flop_per_object": fl_factor,
                                       "achieved": fl_factor * B / (fa_ms * 1e-3) * 1e-12},
                     "ll_kernel": {"ms_per_launch": ll_ms, "flop_per_object": flops_ll(N_EPOCH),
                                   "achieved": flops_ll(N_EPOCH) * B / (ll_ms * 1e-3) * 1e-12},
                     "whole_step": {"flop_per_object": flops_ll(N_EPOCH) + flops_predict(N_EPOCH, M_GRID),
                                    "achieved": (flops_ll(N_EPOCH) + flops_predict(N_EPOCH, M_GRID)) * B * args.steps
                                    / (total_ms * 1e-3) * 1e-12}},
    }
    if not args.no_cpu and world == 1:
        import multiprocessing as mp
        cores = os.cpu_count() or 1
        n_cpu = args.cpu_objects or 1000 * cores
        with mp.get_context("fork").Pool(cores) as pool:
            y0m = y0.reshape(B, N_EPOCH)
            cpu_pass(x[:4 * cores], y[:4 * cores], y0m[:4 * cores], ye[:4 * cores], grid, ny0[:4 * cores], pool, cores)
            v = cpu_pass(x[:n_cpu], y[:n_cpu], y0m[:n_cpu], ye[:n_cpu], grid, ny0[:n_cpu], pool, cores)
        # the CPU leg doubles as the checker: the timed GPU step's likelihoods of the first objects against the port
        from oracle import gp_oracle as O
        ref = O.ll_batched_1d(x[:64], y[:64], y0m[:64], ye[:64], HYP, NUGGET)
        parity = float(np.max(np.abs(ll_first - ref) / np.abs(ref)))
        assert parity < 1e-9, "parity check failed: %g" % parity
        line["parity_max_rel_err_ll"] = parity
        line["cpu_baseline"] = {"value": v, "unit": "objects/s", "cores": cores, "kind": "port",
                                "sample": "first %d of the %d objects, one pass (LL + predict mean/var-diag per object), "
                                          "oracle port, %d processes x 1 BLAS thread" % (n_cpu, B, cores)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
