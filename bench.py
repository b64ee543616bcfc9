#!/usr/bin/env python
"""Benchmark of the LS-SPA hot path (BASELINE.json metric: permutations/sec and time-to-tolerance
at p=100, N=M=1e6, 1/2/4/8 B200 vs host CPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one whole LS-SPA job on the C4 workload: tall-skinny
reduction of [X_train|y_train] and [X_test|y_test] (p=100, N=M=10^6 rows in total), then
2^16 antithetic permutohedron samples PER GPU (= 2^17 permutation evaluations per GPU; one
"permutation" = one square_shapley evaluation, an antithetic pair counts 2), estimator and
epilogue.  `value` = permutations evaluated by all ranks / max-over-ranks device time with the
inputs resident in HBM; `e2e` = the same job through ls_spa_b200.ls_spa() from pinned host
buffers (host->device copies and the result read-back inside the timed region).

`time_to_tolerance` (the second half of the metric) runs the same job with a sample budget large
enough that the error estimate, not the budget, stops it, at tolerance 1e-2, 1e-3 and the
configuration's 1e-4; under torchrun that job is STRONG-scaled (fixed tolerance, rows and samples
sharded), and the CPU figure next to it is samples-needed / CPU rate + the CPU reduction.

Multi-GPU: rows of the reduction are sharded (N=10^6 in total, fixed), permutations are
weak-scaled (2^16 samples per GPU).  Before timing, a small fixed job is compared on every rank
with the numpy oracle (`result_check.multi_gpu_parity`).

`--impl reference` / `cpu_baseline` time the numpy oracle port of the reference (the same LAPACK
calls as the reference, which is pure Python and cannot travel to the GPU box) on the host cores,
with BLAS threads pinned explicitly in child processes, so the numbers do not depend on the
OMP_NUM_THREADS=1 that torchrun exports.
"""

from __future__ import annotations

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

P = 100
N_ROWS = 1_000_000
M_ROWS = 1_000_000
REG = 1e-2
BATCH = 128
BATCHES_PER_GPU = 512          # 2^16 antithetic samples = 2^17 permutation evaluations per GPU
TOL = 1e-4
SEED = 42
FLOP_PER_PERM = 7.0 / 3.0 * P ** 3   # 4/3 p^3 Householder + p^3 triangular solve
TTT_BATCHES = 1 << 17          # sample budget of the time-to-tolerance jobs: 2^24 pairs (never reached)
METRIC = "permutations/sec (LS-SPA, p=100, N=M=1e6 rows, reduction + permutohedron samples + estimator)"


# --------------------------------------------------------------------------- helpers
def read_peaks():
    peaks = {"hbm_gbs": 6650.0, "which": "fallback"}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            peaks.update(json.load(f))
        peaks["which"] = "measured"
    return peaks


def ncu_capture(route):
    """(dram bytes per permutation evaluation, file) of the lift kernel of that route from the newest
    committed `ncu --set full` capture (profiles/r0*_lifts*_ncu_summary.json); (None, None) if absent."""
    names = (["r02_lifts_chol_ncu_summary.json", "r01_lifts_chol_ncu_summary.json"] if route == "cholesky"
             else ["r02_lifts_ncu_summary.json", "r01_lifts_ncu_summary.json"])
    for name in names:
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                d = json.load(f)
            return float(d["dram_bytes_per_launch"]) / float(d["permutation_evaluations_per_launch"]), name
        except Exception:
            continue
    return None, None


def fp64_peak_tflops():
    """FP64 pipe peak: live run of tools/bin/fp64_peak (DFMA and DMMA loops), else the
    committed measurement in profiles/."""
    exe = os.path.join(ROOT, "tools", "bin", "fp64_peak")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
        d = json.loads(out.strip().splitlines()[-1])
        return max(d["dfma_tflops"], d["dmma_tflops"]), "measured live (tools/fp64_peak.cu)"
    except Exception:
        pass
    try:
        with open(os.path.join(ROOT, "profiles", "r01_fp64_peak.json")) as f:
            d = json.load(f)
        return max(d["dfma_tflops"], d["dmma_tflops"]), "profiles/r01_fp64_peak.json"
    except Exception:
        return 37.0, "nominal"


class ClockSampler:
    """nvidia-smi polled every 100 ms from before the warm-up; only the samples whose timestamp falls
    inside the timed region (mark_start .. mark_end) are reported (all samples if none does)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def summarise(rows):
            sm, mx, reasons = [], [], set()
            for _, r in rows:
                try:
                    sm.append(float(r[1]))
                    mx.append(float(r[2]))
                    for n, v in zip(names, r[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(n)
                except Exception:
                    continue
            return sm, mx, reasons

        inside = [x for x in self.rows if self.t0 is not None and self.t0 - 0.05 <= x[0] <= self.t1 + 0.15]
        sm, mx, reasons = summarise(inside)
        scope = "timed region"
        if not sm:
            sm, mx, reasons = summarise(self.rows)
            scope = "whole run (no sample fell inside the timed region)"
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "scope": scope}


# --------------------------------------------------------------------------- GPU arm
def synth_on_device(torch, dev, p, rows_train, rows_test, seed, dist=None, n_train_total=None):
    """Medium-experiment recipe (reference experiments/ground_truth_medium.py:74-106) drawn on
    the device: unit-diagonal covariance A A^T + I with max(p/20, 1) latent factors, (p+1)//10 active
    coefficients equal to 2, SNR 5; centred with the TRAIN means (of all ranks' rows when sharded)."""
    g = torch.Generator(device=dev).manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g, device=dev, dtype=torch.float64)
    A = torch.randn(p, max(p // 20, 1), generator=torch.Generator().manual_seed(1234), dtype=torch.float64).to(dev)
    cov = A @ A.T + torch.eye(p, device=dev, dtype=torch.float64)
    d = cov.diagonal().sqrt()
    cov = cov / d.outer(d)
    Lc = torch.linalg.cholesky(cov)
    theta = torch.zeros(p, dtype=torch.float64, device=dev)
    theta[torch.randperm(p, generator=torch.Generator().manual_seed(99))[: max((p + 1) // 10, 1)].to(dev)] = 2.0
    std = float(torch.sqrt((cov.diagonal() * theta ** 2).sum() / 5.0))
    Xtr = rn(rows_train, p) @ Lc.T
    ytr = Xtr @ theta + std * rn(rows_train)
    Xte = rn(rows_test, p) @ Lc.T
    yte = Xte @ theta + std * rn(rows_test)
    sums = torch.cat([Xtr.sum(0), ytr.sum().reshape(1)])
    total = float(rows_train)
    if dist is not None:
        dist.all_reduce(sums)
        total = float(n_train_total)
    mu, ymu = (sums[:p] / total).unsqueeze(0), sums[p] / total
    return Xtr - mu, Xte - mu, ytr - ymu, yte - ymu


def multi_gpu_parity(L, torch, dist, dev):
    """One small fixed job (p=40, 384 antithetic samples, tolerance 0) on every rank against the numpy
    oracle on the same permutohedron stream: attribution, theta, r_squared, attribution history."""
    from oracle import lsspa_oracle as lo
    from oracle import samplers_oracle as so
    Xtr, Xte, ytr, yte, _, _ = so.gen_data(np.random.default_rng(11), 40, 3000, 2500)
    got = L.ls_spa(Xtr, Xte, ytr, yte, reg=1e-3, method="permutohedron", batch_size=16, num_batches=24,
                   tolerance=0.0, seed=5, antithetical=True, return_history=True)
    perms = so.perms_permutohedron(40, 384, 5)[0]
    want = lo.ls_spa_reference_loop(Xtr, Xte, ytr, yte, reg=1e-3, perms=list(perms), tolerance=0.0, batch_size=16,
                                    antithetical=True, return_attribution_history=True)
    sc = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(np.asarray(b))))
    errs = [sc(got.attribution, want.attribution), sc(got.theta, want.theta),
            abs(float(got.r_squared) - float(want.r_squared)), sc(got.attribution_history, want.attribution_history)]
    shapes_ok = got.error_history.shape == want.error_history.shape
    worst = torch.tensor([max(errs), 0.0 if shapes_ok else 1.0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    # every rank must hold bit-identical results (replicated estimator state)
    same = torch.from_numpy(np.ascontiguousarray(got.attribution)).to(dev)
    lo_, hi_ = same.clone(), same.clone()
    if dist is not None:
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
    return {"job": "p=40, N=3000, M=2500, reg=1e-3, permutohedron, 24 batches x 16 antithetic samples, tolerance 0",
            "checked": "attribution, theta, r_squared, attribution_history vs the numpy oracle on every rank",
            "max_scaled_err_over_ranks": float(worst[0].item()), "tolerance": 1e-9,
            "error_history_shape_ok": bool(worst[1].item() == 0.0),
            "ranks_bit_identical": bool(torch.equal(lo_, hi_)),
            "ok": bool(worst[0].item() < 1e-9 and worst[1].item() == 0.0 and torch.equal(lo_, hi_))}


def side_configs(L, torch, dev):
    """Throughput of the other BASELINE.json configurations on one GPU (device-resident synthetic
    inputs, wall clock around ls_spa() after one warm-up call); parity of each is a -m gpu test."""
    out = {}

    def run(tag, note, p, n, perms, **kw):
        Xtr, Xte, ytr, yte = synth_on_device(torch, dev, p, n, n, 77)
        L.ls_spa(Xtr, Xte, ytr, yte, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = L.ls_spa(Xtr, Xte, ytr, yte, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[tag] = {"workload": note, "seconds": dt, "permutations": perms, "permutations_per_s": perms / dt,
                    "sum_attribution_minus_r2": float(abs(r.attribution.sum() - r.r_squared)),
                    "overall_error": float(r.overall_error)}
        del Xtr, Xte, ytr, yte
        torch.cuda.empty_cache()

    run("C2", "p=10, N=M=1e5, method=exact (all 10! = 3628800 permutations)", 10, 100_000, 3_628_800, method="exact")
    run("C3", "p=100, N=M=1e5, method=argsort, 2^7 x 2^7 samples, no antithetic pairs", 100, 100_000, 1 << 14,
        method="argsort", batch_size=128, num_batches=128, tolerance=0.0, antithetical=False)
    # the guarded fallback, measured: singular values spread over 1e5 in random directions, so the condition
    # guard refuses the Cholesky route: two-pass CholeskyQR reduction + Householder lift kernel (lifts_mma.cu)
    g = torch.Generator(device=dev).manual_seed(5)
    Q, _ = torch.linalg.qr(torch.randn(P, P, generator=g, device=dev, dtype=torch.float64))
    mixm = (Q * torch.logspace(0, -5, P, device=dev, dtype=torch.float64)) @ Q.T
    n = 100_000
    Xtr = torch.randn(n, P, generator=g, device=dev, dtype=torch.float64) @ mixm
    Xte = torch.randn(n, P, generator=g, device=dev, dtype=torch.float64) @ mixm
    th = torch.randn(P, generator=g, device=dev, dtype=torch.float64)
    ytr = Xtr @ th + 1e-3 * torch.randn(n, generator=g, device=dev, dtype=torch.float64)
    yte = Xte @ th + 1e-3 * torch.randn(n, generator=g, device=dev, dtype=torch.float64)
    kw = dict(method="permutohedron", batch_size=128, num_batches=128, tolerance=0.0, antithetical=True)
    from ls_spa_b200 import ops as _ops
    L.ls_spa(Xtr, Xte, ytr, yte, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = L.ls_spa(Xtr, Xte, ytr, yte, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["ill_conditioned_p100"] = {
        "workload": "p=100, N=M=1e5, cond(X) = 1e5 in random directions, permutohedron, 2^14 antithetic pairs, tolerance 0",
        "route": _ops.LIFT_ROUTE, "seconds": dt, "permutations": 2 << 14, "permutations_per_s": (2 << 14) / dt,
        "sum_attribution_minus_r2": float(abs(r.attribution.sum() - r.r_squared))}
    return out


def c5_line(L, torch, dist, dev, world, rank):
    """BASELINE config 5 itself: p=1000, N=M=10^6 rows sharded over the ranks, method='random', 2^16 permutations
    (no antithetic pairs) sharded over the ranks; at world 1 the same job with 1/8 of the rows and permutations
    (one GPU's share).  Wall clock around ls_spa() after one warm-up call, max over ranks."""
    p, share = 1000, (1 if world > 1 else 8)
    rows = 1_000_000 // (world * share)
    nperm = (1 << 16) // share
    Xtr, Xte, ytr, yte = synth_on_device(torch, dev, p, rows, rows, 2000 + rank, dist if world > 1 else None, rows * world)
    kw = dict(method="random", batch_size=128, num_batches=nperm // 128, tolerance=0.0, antithetical=False,
              row_sharded=world > 1)
    L.ls_spa(Xtr, Xte, ytr, yte, **kw)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = L.ls_spa(Xtr, Xte, ytr, yte, **kw)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    del Xtr, Xte, ytr, yte
    torch.cuda.empty_cache()
    peak, _ = fp64_peak_tflops() if rank == 0 else (None, None)
    tf = (4.0 / 3.0) * p ** 3 * nperm / dt / 1e12
    return {"workload": f"p=1000, N=M={rows * world} rows ({rows} per GPU), method=random, {nperm} permutations, no antithetic pairs"
                        + ("" if world > 1 else " (one GPU's share of C5: 1/8 of the rows and permutations)"),
            "seconds": dt, "permutations": nperm, "permutations_per_s": nperm / dt,
            "fp64_tflops_executed": tf, "fp64_frac_executed_per_gpu": (tf / world / peak) if peak else None,
            "sum_attribution_minus_r2": float(abs(r.attribution.sum() - r.r_squared)), "overall_error": float(r.overall_error),
            "note": "whole job: Gram reduction (gram_big.cu), blocked Cholesky, wide lift kernels (lifts_big.cu), estimator"}


def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the run totals exchanged per super-batch are 0.9 MB per rank: NCCL's default pick for that size is the
        # low-latency LL protocol, whose bandwidth (flag word per 8 bytes) makes the exchange ~3x slower
        # than the Simple protocol over NVLink (measured: 10.74 -> 11.03 M perm/s on 2 GPUs)
        os.environ.setdefault("NCCL_PROTO", "Simple")
        # NCCL prints its version banner to stdout on first use: keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    import ls_spa_b200 as L
    from ls_spa_b200 import engine, ops

    parity = multi_gpu_parity(L, torch, dist if world > 1 else None, dev) if world > 1 else None

    rows_tr = N_ROWS // world + (1 if rank < N_ROWS % world else 0)
    rows_te = M_ROWS // world + (1 if rank < M_ROWS % world else 0)
    Xtr, Xte, ytr, yte = synth_on_device(torch, dev, P, rows_tr, rows_te, 1000 + rank,
                                         dist if world > 1 else None, N_ROWS)
    num_batches = BATCHES_PER_GPU * world
    perms_per_step = 2 * BATCH * num_batches          # antithetic pair = 2 evaluations
    kw = dict(reg=REG, method="permutohedron", batch_size=BATCH, num_batches=num_batches, tolerance=TOL,
              seed=SEED, antithetical=True, row_sharded=world > 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    last = {}

    def step_device():
        last["res"] = L.ls_spa(Xtr, Xte, ytr, yte, **kw)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    ops.LIFT_TRACE = []
    launches0 = ops.LAUNCHES
    sampler.mark_start()
    ms_total = timed(step_device, args.steps)
    sampler.mark_end()
    launches = ops.LAUNCHES - launches0
    trace, ops.LIFT_TRACE = ops.LIFT_TRACE, None
    clocks = sampler.stop() if rank == 0 else None
    lift_ms = sum(a.elapsed_time(b) for a, b, _ in trace)
    lift_perms = sum(n for _, _, n in trace)
    ms_step = ms_total / args.steps
    value = perms_per_step / (ms_step * 1e-3)

    # stand-alone pass over the reduction for its own roofline line (this rank's rows, collectives included)
    backend, coll = engine.CudaBackend(dev), engine.Collective(None)
    red_fn = lambda: engine.reduce_problem(backend, coll, Xtr, Xte, ytr, yte, REG, P, n_train_global=N_ROWS)
    red_fn()
    red_ms = timed(red_fn, 5) / 5
    red_bytes = 8.0 * (rows_tr + rows_te) * (P + 1)
    red_flop = float(rows_tr + rows_te) * (P + 1) ** 2          # symmetric Gram count

    # end to end: pinned host buffers -> ls_spa() -> host results
    host = [t.cpu().pin_memory() for t in (Xtr, Xte, ytr, yte)]
    h2d = sum(t.numel() * 8 for t in host)

    def step_e2e():
        r = L.ls_spa(host[0], host[1], host[2], host[3], **kw)
        last["e2e"] = r

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = perms_per_step / float(e2e_s.item())
    r = last["e2e"]
    d2h = 8 * (r.attribution.size + r.theta.size + r.attribution_errors.size + r.error_history.size + 2)
    # the optional fp32 TRANSFER mode: the same values rounded to float32 in pinned host memory cross the
    # link as float32 (half the bytes) and are widened on the device; the arithmetic stays fp64
    host32 = [t.float().cpu().pin_memory() for t in (Xtr, Xte, ytr, yte)]

    def step_e2e32():
        last["e2e32"] = L.ls_spa(host32[0], host32[1], host32[2], host32[3], **kw)

    step_e2e32()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e32()
    barrier()
    e2e32_s = torch.tensor([(time.perf_counter() - t0) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e32_s, op=dist.ReduceOp.MAX)
    e2e32_val = perms_per_step / float(e2e32_s.item())
    e2e32_diff = float(np.max(np.abs(last["e2e32"].attribution - last["e2e"].attribution)) / np.max(np.abs(last["e2e"].attribution)))
    del host32
    # the link alone: the same pinned buffers copied once more (reported next to e2e)
    dst = torch.empty_like(Xtr)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dst.copy_(host[0], non_blocking=True)
    torch.cuda.synchronize()
    h2d_gbs = host[0].numel() * 8 / (time.perf_counter() - t0) / 1e9
    del dst, host

    # time to tolerance (the second half of BASELINE.json's metric): the same job, device-resident
    # inputs, a sample budget that is never reached (2^24 pairs), stopped by the error estimate.
    # Strong-scaled under torchrun: the tolerance is fixed, every rank takes 1/world of each round.
    ttt = []
    for tol in (1e-2, 1e-3, TOL):
        kw_t = dict(kw, tolerance=tol, num_batches=TTT_BATCHES)
        L.ls_spa(Xtr, Xte, ytr, yte, **kw_t)
        barrier()
        t0 = time.perf_counter()
        rt = L.ls_spa(Xtr, Xte, ytr, yte, **kw_t)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        nb = int(rt.error_history.size)
        ttt.append({"tolerance": tol, "seconds": float(dt.item()), "estimated_error_at_stop": float(rt.overall_error),
                    "batches": nb, "pairs": nb * BATCH, "permutation_evaluations": 2 * nb * BATCH,
                    "stopped_by": "tolerance" if nb < TTT_BATCHES else "budget",
                    "scaling": "strong (fixed tolerance, rows and samples sharded over the ranks)",
                    "timer": "wall clock around ls_spa() with device-resident inputs, max over ranks"})

    c5 = None
    if not args.no_cpu:
        try:
            c5 = c5_line(L, torch, dist if world > 1 else None, dev, world, rank)
        except Exception as exc:      # a side line must never cost the headline
            c5 = {"error": repr(exc)}

    if rank == 0:
        peaks = read_peaks()
        fp64_peak, fp64_src = fp64_peak_tflops()
        ach = FLOP_PER_PERM * lift_perms / (lift_ms * 1e-3) / 1e12 if lift_ms > 0 else 0.0
        # `achieved` counts SURVEY 8(d)'s algorithmic figure (7/3 p^3: Householder R + triangular
        # solve).  The Cholesky route executes 1/3 p^3 + p^3; frac_executed rates the pipe on that.
        chol = ops.LIFT_ROUTE == "cholesky"
        exec_flop = (4.0 / 3.0 if chol else 7.0 / 3.0) * P ** 3
        bpp, cap_file = ncu_capture(ops.LIFT_ROUTE)
        traffic = bpp * lift_perms / max(len(trace), 1) if bpp is not None else None
        red_gbs = red_bytes / (red_ms * 1e-3) / 1e9
        red_tf = red_flop / (red_ms * 1e-3) / 1e12
        out = {
            "metric": METRIC,
            "value": value, "unit": "permutations/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C4: p=100, N=M=10^6 (rows sharded over GPUs), reg=1e-2, method=permutohedron, "
                                   "antithetic, batch 128 x 512 batches per GPU (2^17 permutation evaluations per "
                                   "GPU per step), tolerance 1e-4",
                       "permutations_per_step": perms_per_step, "rows_per_gpu": rows_tr,
                       "l2": "inputs (2 x 808 MB / n_gpus) larger than L2; no explicit flush"},
            "e2e": {"value": e2e_val, "unit": "permutations/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "timer": "wall clock around ls_spa() incl. H2D/D2H, max over ranks",
                    "float32_inputs": {"value": e2e32_val, "unit": "permutations/s", "h2d_bytes_per_step": h2d // 2,
                                       "note": "optional fp32 transfer mode: float32 pinned host inputs, widened on the "
                                               "device, fp64 arithmetic", "max_scaled_attribution_diff_vs_f64_inputs": e2e32_diff},
                    "h2d_link_gbs_measured": h2d_gbs,
                    "link_floor_permutations_per_s": perms_per_step / (h2d / (h2d_gbs * 1e9))},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "pipe": "fp64 (DFMA/DMMA share one pipe; tcgen05 has no fp64)",
                         "kernel": "lifts_chol_kernel" if chol else "lifts_mma_kernel",
                         "route": ops.LIFT_ROUTE, "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": ach / fp64_peak if fp64_peak else None,
                         "executed_flop_per_permutation": exec_flop,
                         "frac_executed": (ach * exec_flop / FLOP_PER_PERM) / fp64_peak if fp64_peak else None,
                         "traffic": traffic,
                         "traffic_note": f"dram read+write bytes per launch from the committed ncu --set full capture "
                                         f"profiles/{cap_file}, scaled to this launch size (not re-measured in this run); "
                                         f"algorithmic HBM bytes/launch = 1200 B x evaluations / 2",
                         "peak_source": fp64_src, "kernel_ms_per_step": lift_ms / args.steps,
                         "kernel_share_of_step": lift_ms / ms_total,
                         "algorithmic_flop_per_permutation": FLOP_PER_PERM},
            # intensity (p+1)/8 = 12.6 flop/B is above the FP64 ridge (~5.7 flop/B), so the
            # binding roof of the reduction at p=100 is the FP64 pipe; both fractions are reported
            "roofline_reduce": {"bound": "tensor", "binding": "fp64 pipe (12.6 flop/B > ridge 5.7 flop/B)",
                                "kernel": "gram_tma_kernel (CholeskyQR Gram pass, DMMA) + gram_tail_kernel",
                                "ms": red_ms, "algorithmic_bytes": red_bytes, "algorithmic_flop": red_flop,
                                "achieved": red_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                                "frac": red_tf / fp64_peak if fp64_peak else None,
                                "achieved_gbs": red_gbs, "peak_gbs": peaks["hbm_gbs"], "frac_hbm": red_gbs / peaks["hbm_gbs"],
                                "peak_source": f"fp64: {fp64_src}; hbm: MEASURED_PEAKS.json ({peaks['which']})",
                                "rows_this_rank": rows_tr + rows_te},
            "clocks": clocks,
            "time_to_tolerance": ttt,
            "result_check": {"sum_attribution_minus_r2": float(abs(last["res"].attribution.sum() - last["res"].r_squared)),
                             "overall_error": float(last["res"].overall_error),
                             "multi_gpu_parity": parity},
        }
        if c5 is not None:
            out["configs"] = {"C5" if world > 1 else "C5_per_gpu": c5}
        if world == 1 and not args.no_cpu:
            try:
                out["configs"].update(side_configs(L, torch, dev))
            except Exception as exc:      # a side line must never cost the headline
                out["configs"]["error"] = repr(exc)
            out["cpu_baseline"] = cpu_baseline(sample_perms_per_core=1024, full=True)
            cb = out["cpu_baseline"]
            for t in ttt:
                t["cpu_seconds_extrapolated"] = (cb["reduce_data_s"] + t["permutation_evaluations"]
                                                 / cb["loop_only_permutations_per_s"])
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- CPU arm
def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def _cpu_worker(job):
    from threadpoolctl import threadpool_limits
    from oracle import lsspa_oracle as lo
    R_tr, R_te, c_tr, c_te, ynsq, perms = job
    with threadpool_limits(limits=1):
        t0 = time.perf_counter()
        lo.mean_of_lifts(R_tr, R_te, c_tr, c_te, ynsq, perms, antithetical=True)
        return time.perf_counter() - t0


def cheap_data(rows, seed=SEED):
    """The medium-experiment recipe without its SVD-based multivariate_normal (too slow at 10^6 rows):
    X = Z chol(cov)^T.  Same distribution, used only for timing the CPU path."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((P, P // 20))
    cov = a @ a.T + np.eye(P)
    v = np.sqrt(np.diag(cov))
    Lc = np.linalg.cholesky(cov / np.outer(v, v))
    theta = np.zeros(P)
    theta[rng.permutation(P)[: (P + 1) // 10]] = 2.0
    std = np.sqrt(np.sum(theta ** 2) / 5.0)
    out = []
    for _ in range(2):
        X = rng.standard_normal((rows, P)) @ Lc.T
        out.append((X, X @ theta + std * rng.standard_normal(rows)))
    (Xtr, ytr), (Xte, yte) = out
    mu, ymu = Xtr.mean(0, keepdims=True), ytr.mean()
    return Xtr - mu, Xte - mu, ytr - ymu, yte - ymu


def child_main(args):
    """Body of the child processes of the CPU arm (BLAS threads were pinned through the environment
    before numpy was imported).  Prints one JSON line."""
    from oracle import lsspa_oracle as lo
    from oracle import samplers_oracle as so
    what, rows, count = args.child, args.child_rows, args.child_count
    Xtr, Xte, ytr, yte = cheap_data(rows)
    t0 = time.perf_counter()
    fac = lo.reduce_data(Xtr, Xte, ytr, yte, REG)
    t_red = time.perf_counter() - t0
    out = {"what": what, "rows": rows, "reduce_data_s": t_red,
           "blas_threads": os.environ.get("OPENBLAS_NUM_THREADS")}
    if what == "loop":
        # the reference's loop as shipped (ls_spa/ls_spa.py:196-236: square_shapley x 2, merge_sample_cov,
        # merge_sample_mean, error_estimates every batch), explicit permutohedron stream
        perms = so.perms_permutohedron(P, count, SEED)[0]
        t0 = time.perf_counter()
        res = lo.ls_spa_reference_loop(Xtr, Xte, ytr, yte, reg=REG, perms=list(perms), tolerance=0.0,
                                       batch_size=BATCH, antithetical=True)
        t_all = time.perf_counter() - t0
        out.update(pairs=count, total_s=t_all, loop_s=max(t_all - t_red, 1e-9),
                   permutations_per_s=2 * count / max(t_all - t_red, 1e-9), r_squared=float(res.r_squared))
    elif what == "factors":
        np.savez(args.child_out, R_tr=fac[0], R_te=fac[1], c_tr=fac[2], c_te=fac[3], ynsq=float(yte @ yte))
    print(json.dumps(out), flush=True)


def run_child(what, rows, count, threads, out=None, timeout=900):
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OPENBLAS_NUM_THREADS=str(threads),
               MKL_NUM_THREADS=str(threads), PYTHONDONTWRITEBYTECODE="1")
    cmd = [sys.executable, os.path.abspath(__file__), "--child", what, "--child-rows", str(rows),
           "--child-count", str(count)] + (["--child-out", out] if out else [])
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    if res.returncode != 0:
        raise RuntimeError(f"cpu child {what} failed: {res.stderr[-500:]}")
    return json.loads(res.stdout.strip().splitlines()[-1])


_REDUCE_CACHE = {}


def cpu_reduce(rows, cores):
    """reduce_data of the oracle (np.linalg.qr x 2, as the reference) on `rows` train and test rows with
    `cores` BLAS threads, in a child process; also leaves the reduced factors for the loop workers."""
    key = (rows, cores)
    if key not in _REDUCE_CACHE:
        path = os.path.join("/tmp", f"lsspa_bench_factors_{os.getpid()}_{rows}.npz")
        info = run_child("factors", rows, 0, cores, out=path)
        z = np.load(path)
        os.remove(path)
        _REDUCE_CACHE[key] = (info["reduce_data_s"], tuple(z[k] for k in ("R_tr", "R_te", "c_tr", "c_te")) + (float(z["ynsq"]),))
    return _REDUCE_CACHE[key]


def cpu_baseline(sample_perms_per_core=256, full=True):
    """Oracle port (same LAPACK calls as the reference) on all host cores: P worker processes x 1
    BLAS thread over slices of the same permutohedron stream (variant 3);
    reduce_data timed ONCE at the full N=M=10^6 with all cores as BLAS threads.  With full=True also the
    reference's loop as shipped (variants 1 and 2: default BLAS threads / 1 thread) on a short stream."""
    import multiprocessing as mp
    from oracle import samplers_oracle as so
    cores = host_cores()
    t_red, fac = cpu_reduce(N_ROWS, cores)
    perms = so.perms_permutohedron(P, sample_perms_per_core * cores, SEED)[0]
    jobs = [fac + (perms[i::cores],) for i in range(cores)]
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_worker, [fac + (perms[:8],)] * cores)      # spin the workers up (imports, page-in)
        t0 = time.perf_counter()
        pool.map(_cpu_worker, jobs)
        t_loop = time.perf_counter() - t0
    evals = 2 * len(perms)                         # antithetic pairs
    loop_rate = evals / t_loop
    job_perms = 2 * BATCH * BATCHES_PER_GPU
    whole_job = job_perms / (t_red + job_perms / loop_rate)
    out = {"value": whole_job, "unit": "permutations/s", "cores": cores, "kind": "port",
           "loop_only_permutations_per_s": loop_rate, "reduce_data_s": t_red,
           "reduce_data_rows": N_ROWS, "reduce_data_blas_threads": cores,
           "host": {"cpu_count": os.cpu_count(), "affinity": cores, "numpy": np.__version__},
           "sample": f"{evals} permutation evaluations (p=100, {len(perms)} antithetic permutohedron samples) "
                     f"split over {cores} processes x 1 BLAS thread; reduce_data timed once at the full N=M={N_ROWS} rows "
                     f"with {cores} BLAS threads (child process, threads pinned by environment); "
                     f"value = 2^17 / (reduce + 2^17 / loop rate)"}
    if full:
        try:
            one = run_child("loop", 10_000, 256, 1)
            dflt = run_child("loop", 10_000, 32, cores)
            out["as_shipped"] = {
                "what": "the reference's own loop (square_shapley x 2 per pair, merge_sample_cov, merge_sample_mean, "
                        "error_estimates per batch; oracle restatement of ls_spa/ls_spa.py:196-236), one process, "
                        "N=M=10^4 rows (the loop does not depend on N)",
                "one_blas_thread": {"permutations_per_s": one["permutations_per_s"], "pairs": one["pairs"]},
                "default_blas_threads": {"permutations_per_s": dflt["permutations_per_s"], "pairs": dflt["pairs"],
                                         "threads": cores},
                "whole_job_one_thread_permutations_per_s": job_perms / (t_red + job_perms / one["permutations_per_s"]),
                "whole_job_default_threads_permutations_per_s": job_perms / (t_red + job_perms / dflt["permutations_per_s"]),
            }
        except Exception as exc:
            out["as_shipped"] = {"error": repr(exc)}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    steps, warm = args.steps, args.warmup
    vals = []
    for i in range(warm + steps):
        b = cpu_baseline(sample_perms_per_core=128, full=(i == warm + steps - 1))
        if i >= warm:
            vals.append(b)
    v = float(np.mean([b["value"] for b in vals]))
    b = vals[-1]
    b["value"] = v
    out = {"impl": "reference", "metric": METRIC,
           "value": v, "unit": "permutations/s", "n_gpus": world, "steps": steps, "warmup": warm,
           "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": "C4: p=100, N=M=10^6, reg=1e-2, permutohedron, antithetic, 2^17 permutation "
                                  "evaluations per step (each step a bounded sample of the loop; reduce_data timed once "
                                  "at full size; see cpu_baseline.sample)"},
           "cpu_baseline": b,
           "e2e": {"value": v, "unit": "permutations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline and side-configuration legs")
    ap.add_argument("--child", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--child-rows", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--child-count", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--child-out", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.child:
        child_main(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
