#!/usr/bin/env python
"""Benchmark of the LS-SPA hot path (BASELINE.json metric: permutations/sec at p=100, N=M=1e6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one whole LS-SPA job on the C4 workload: tall-skinny
reduction of [X_train|y_train] and [X_test|y_test] (p=100, N=M=10^6 rows in total), then
2^16 antithetic permutohedron samples PER GPU (= 2^17 permutation evaluations per GPU; one
"permutation" = one square_shapley evaluation, an antithetic pair counts 2), estimator and
epilogue.  `value` = permutations evaluated by all ranks / max-over-ranks device time with the
inputs resident in HBM; `e2e` = the same job through ls_spa_b200.ls_spa() from pinned host
buffers (host->device copies and the result read-back inside the timed region).

Multi-GPU: rows of the reduction are sharded (N=10^6 in total, fixed), permutations are
weak-scaled (2^16 samples per GPU).  `--impl reference` times the numpy oracle port of the
reference (same LAPACK calls as the reference) on all host cores, on a bounded sample.
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


# --------------------------------------------------------------------------- helpers
def read_peaks():
    peaks = {"hbm_gbs": 6650.0, "which": "fallback"}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            peaks.update(json.load(f))
        peaks["which"] = "measured"
    return peaks


def lifts_dram_bytes_per_perm(route=None):
    """DRAM bytes (read + write) per permutation evaluation of the lift kernel, from the committed
    `ncu --set full` capture of the kernel of that route (profiles/r01_lifts*_ncu_summary.json);
    None if absent."""
    name = "r01_lifts_chol_ncu_summary.json" if route == "cholesky" else "r01_lifts_ncu_summary.json"
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            d = json.load(f)
        return float(d["dram_bytes_per_launch"]) / float(d["permutation_evaluations_per_launch"])
    except Exception:
        return None


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
def synth_on_device(torch, dev, rows_train, rows_test, seed):
    """Medium-experiment recipe (reference experiments/ground_truth_medium.py:74-106) drawn on
    the device: unit-diagonal covariance A A^T + I with p/20 latent factors, (p+1)//10 active
    coefficients equal to 2, SNR 5; centred with the train means."""
    g = torch.Generator(device=dev).manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g, device=dev, dtype=torch.float64)
    A = torch.randn(P, P // 20, generator=torch.Generator().manual_seed(1234), dtype=torch.float64).to(dev)
    cov = A @ A.T + torch.eye(P, device=dev, dtype=torch.float64)
    d = cov.diagonal().sqrt()
    cov = cov / d.outer(d)
    Lc = torch.linalg.cholesky(cov)
    theta = torch.zeros(P, dtype=torch.float64, device=dev)
    theta[torch.randperm(P, generator=torch.Generator().manual_seed(99))[: (P + 1) // 10].to(dev)] = 2.0
    std = float(torch.sqrt((cov.diagonal() * theta ** 2).sum() / 5.0))
    Xtr = rn(rows_train, P) @ Lc.T
    ytr = Xtr @ theta + std * rn(rows_train)
    Xte = rn(rows_test, P) @ Lc.T
    yte = Xte @ theta + std * rn(rows_test)
    # the driver centres with the train means; a per-shard mean is close enough for a benchmark
    mu, ymu = Xtr.mean(0, keepdim=True), ytr.mean()
    return Xtr - mu, Xte - mu, ytr - ymu, yte - ymu


def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
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

    rows_tr = N_ROWS // world + (1 if rank < N_ROWS % world else 0)
    rows_te = M_ROWS // world + (1 if rank < M_ROWS % world else 0)
    Xtr, Xte, ytr, yte = synth_on_device(torch, dev, rows_tr, rows_te, 1000 + rank)
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

    # stand-alone pass over the reduction (HBM-bound stage) for its own roofline line
    backend, coll = engine.CudaBackend(dev), engine.Collective(None)
    red_ms = timed(lambda: engine.reduce_problem(backend, coll, Xtr, Xte, ytr, yte, REG, P,
                                                 n_train_global=N_ROWS), 3) / 3
    red_bytes = 8.0 * (rows_tr + rows_te) * (P + 1)

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

    # time to tolerance (the second half of BASELINE.json's metric): the same job, device-resident
    # inputs, stopped by the error estimate at the reference's default tolerance 1e-2 and at 2e-3
    # (outside the timed region; every rank runs it, the stop decision is collective)
    ttt = []
    for tol in (1e-2, 2e-3):
        kw_t = dict(kw, tolerance=tol)
        L.ls_spa(Xtr, Xte, ytr, yte, **kw_t)
        barrier()
        t0 = time.perf_counter()
        rt = L.ls_spa(Xtr, Xte, ytr, yte, **kw_t)
        barrier()
        ttt.append({"tolerance": tol, "seconds": time.perf_counter() - t0,
                    "estimated_error_at_stop": float(rt.overall_error),
                    "batches": int(rt.error_history.size), "pairs": int(rt.error_history.size) * BATCH})

    if rank == 0:
        peaks = read_peaks()
        fp64_peak, fp64_src = fp64_peak_tflops()
        ach = FLOP_PER_PERM * lift_perms / (lift_ms * 1e-3) / 1e12 if lift_ms > 0 else 0.0
        # `achieved` counts SURVEY 8(d)'s algorithmic figure (7/3 p^3: Householder R + triangular
        # solve).  The Cholesky route executes 1/3 p^3 + p^3; frac_executed rates the pipe on that.
        chol = ops.LIFT_ROUTE == "cholesky"
        exec_flop = (4.0 / 3.0 if chol else 7.0 / 3.0) * P ** 3
        bpp = lifts_dram_bytes_per_perm(ops.LIFT_ROUTE)
        traffic = bpp * lift_perms / max(len(trace), 1) if bpp is not None else None
        out = {
            "metric": "permutations/sec (LS-SPA, p=100, N=M=1e6 rows, reduction + permutohedron samples + estimator)",
            "value": value, "unit": "permutations/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C4: p=100, N=M=10^6 (rows sharded over GPUs), reg=1e-2, method=permutohedron, "
                                   "antithetic, batch 128 x 512 batches per GPU (2^17 permutation evaluations per "
                                   "GPU per step), tolerance 1e-4",
                       "permutations_per_step": perms_per_step, "rows_per_gpu": rows_tr,
                       "l2": "inputs (2 x 808 MB / n_gpus) larger than L2; no explicit flush"},
            "e2e": {"value": e2e_val, "unit": "permutations/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "timer": "wall clock around ls_spa() incl. H2D/D2H, max over ranks"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "pipe": "fp64 (DFMA/DMMA share one pipe; tcgen05 has no fp64)",
                         "kernel": "lifts_chol_kernel" if chol else "lifts_mma_kernel",
                         "route": ops.LIFT_ROUTE, "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": ach / fp64_peak if fp64_peak else None,
                         "executed_flop_per_permutation": exec_flop,
                         "frac_executed": (ach * exec_flop / FLOP_PER_PERM) / fp64_peak if fp64_peak else None,
                         "traffic": traffic,
                         "traffic_note": "dram read+write bytes per launch from the committed ncu --set full capture, "
                                         "scaled to this launch size; algorithmic HBM bytes/launch = 1200 B x evaluations / 2",
                         "peak_source": fp64_src, "kernel_ms_per_step": lift_ms / args.steps,
                         "kernel_share_of_step": lift_ms / ms_total,
                         "algorithmic_flop_per_permutation": FLOP_PER_PERM},
            "roofline_reduce": {"bound": "hbm", "kernel": "gram_rows_kernel (CholeskyQR, second pass only when cond > 1e3) + chol_factor_kernel",
                                "achieved": red_bytes / (red_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                                "unit": "GB/s", "frac": red_bytes / (red_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                "peak_source": peaks["which"], "ms": red_ms, "algorithmic_bytes": red_bytes},
            "clocks": clocks,
            "time_to_tolerance": ttt,
            "result_check": {"sum_attribution_minus_r2": float(abs(last["res"].attribution.sum() - last["res"].r_squared)),
                             "overall_error": float(last["res"].overall_error)},
        }
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(sample_perms_per_core=1024)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- CPU arm
def _cpu_worker(job):
    from threadpoolctl import threadpool_limits
    from oracle import lsspa_oracle as lo
    R_tr, R_te, c_tr, c_te, ynsq, perms = job
    with threadpool_limits(limits=1):
        t0 = time.perf_counter()
        lo.mean_of_lifts(R_tr, R_te, c_tr, c_te, ynsq, perms, antithetical=True)
        return time.perf_counter() - t0


_CPU_CACHE = {}


def cpu_data(rows):
    if rows not in _CPU_CACHE:
        from oracle import samplers_oracle as so
        rng = np.random.default_rng(SEED)
        _CPU_CACHE[rows] = so.gen_data(rng, P, rows, rows)[:4]
    return _CPU_CACHE[rows]


def cpu_baseline(sample_perms_per_core=256):
    """Oracle port (same LAPACK calls as the reference) on all host cores: P worker processes x 1
    BLAS thread over slices of the same permutohedron stream; reduce_data timed on a row sample
    with default BLAS threads and scaled linearly to 10^6 rows."""
    import multiprocessing as mp
    from oracle import lsspa_oracle as lo
    from oracle import samplers_oracle as so
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    sample_rows = 100_000
    Xtr, Xte, ytr, yte = cpu_data(sample_rows)
    t0 = time.perf_counter()
    fac = lo.reduce_data(Xtr, Xte, ytr, yte, REG)
    t_red = time.perf_counter() - t0
    ynsq = float(yte @ yte)
    perms = so.perms_permutohedron(P, sample_perms_per_core * cores, SEED)[0]
    jobs = [(fac[0], fac[1], fac[2], fac[3], ynsq, perms[i::cores]) for i in range(cores)]
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_worker, jobs[:cores])          # spin the workers up (imports, page-in)
        t0 = time.perf_counter()
        pool.map(_cpu_worker, jobs)
        t_loop = time.perf_counter() - t0
    evals = 2 * len(perms)                         # antithetic pairs
    loop_rate = evals / t_loop
    job_perms = 2 * BATCH * BATCHES_PER_GPU
    t_red_full = t_red * (N_ROWS / sample_rows)
    whole_job = job_perms / (t_red_full + job_perms / loop_rate)
    return {"value": whole_job, "unit": "permutations/s", "cores": cores, "kind": "port",
            "loop_only_permutations_per_s": loop_rate, "reduce_data_s_extrapolated_1e6_rows": t_red_full,
            "sample": f"{evals} permutation evaluations (p=100, {len(perms)} antithetic permutohedron samples) "
                      f"split over {cores} processes x 1 BLAS thread; reduce_data timed on N=M={sample_rows} rows "
                      f"(default BLAS threads) and scaled x{N_ROWS // sample_rows}; value = 2^17 / (reduce + 2^17 / loop rate)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    steps, warm = args.steps, args.warmup
    vals = []
    for i in range(warm + steps):
        b = cpu_baseline(sample_perms_per_core=256)
        if i >= warm:
            vals.append(b)
    v = float(np.mean([b["value"] for b in vals]))
    b = vals[-1]
    b["value"] = v
    out = {"impl": "reference", "metric": "permutations/sec (LS-SPA, p=100, N=M=1e6 rows, reduction + permutohedron samples + estimator)",
           "value": v, "unit": "permutations/s", "n_gpus": world, "steps": steps, "warmup": warm,
           "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": "C4: p=100, N=M=10^6, reg=1e-2, permutohedron, antithetic, 2^17 permutation "
                                  "evaluations per step (bounded sample, extrapolated; see cpu_baseline.sample)"},
           "cpu_baseline": b,
           "e2e": {"value": v, "unit": "permutations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
