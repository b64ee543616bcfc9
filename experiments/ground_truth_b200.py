"""Ground-truth experiment on the medium shape (the role of the reference's
experiments/ground_truth_medium.py:74-119 and notebooks/medium_experiment.py:340-603, without the
plots): correlated synthetic data drawn on the device after that recipe, a ground truth from 2^19
antithetic pairs of the PCG64 stream, then every sampler (random, permutohedron, argsort) with and
without antithetic pairs for 2^13 evaluations, and the true error
||attribution_history[k] - ground truth||_2 against the number of samples next to the estimated
error.  Everything runs through ls_spa_b200.ls_spa on the GPU; results go to experiments/data/ as
.npy / .json.

    python experiments/ground_truth_b200.py [--p 100] [--n 100000] [--m 100000] [--gt-log2 19]
                                            [--samples-log2 13]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ls_spa_b200 as L  # noqa: E402

STN_RATIO = 5.0
CONDITIONING = 20.0


def gen_data_device(seed, p, n, m, dev):
    """Factor-model covariance with unit diagonal, a tenth of the features relevant, signal-to-noise
    ratio 5 (the reference's recipe), drawn on the device with torch generators: nothing crosses
    PCIe."""
    g = torch.Generator(device=dev).manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g, device=dev, dtype=torch.float64)
    A = rn(p, max(int(p / CONDITIONING), 1))
    cov = A @ A.T + torch.eye(p, device=dev, dtype=torch.float64)
    d = cov.diagonal().sqrt()
    cov = cov / d.outer(d)
    Lc = torch.linalg.cholesky(cov)
    theta = torch.zeros(p, dtype=torch.float64, device=dev)
    k = max((p + 1) // 10, 1)
    theta[torch.randperm(p, generator=torch.Generator().manual_seed(seed + 1))[:k].to(dev)] = 2.0
    std = float(torch.sqrt((cov.diagonal() * theta ** 2).sum() / STN_RATIO))
    Xtr = rn(n, p) @ Lc.T
    ytr = Xtr @ theta + std * rn(n)
    Xte = rn(m, p) @ Lc.T
    yte = Xte @ theta + std * rn(m)
    mu, ymu = Xtr.mean(0, keepdim=True), ytr.mean()
    return Xtr - mu, Xte - mu, ytr - ymu, yte - ymu


def run(p, n, m, gt_log2, samples_log2, seed=42, out_dir=None, quiet=False):
    dev = torch.device("cuda")
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    Xtr, Xte, ytr, yte = gen_data_device(seed, p, n, m, dev)
    t_data = time.perf_counter() - t0

    # ground truth: the device continues the PCG64 stream of this numpy generator (the reference script
    # draws its permutations from the generator that made the data)
    t0 = time.perf_counter()
    gt = L.ls_spa(Xtr, Xte, ytr, yte, max_samples=2 ** gt_log2, batch_size=2 ** 10, tolerance=0.0, seed=rng,
                  antithetical=True)
    torch.cuda.synchronize()
    t_gt = time.perf_counter() - t0
    gt_attr = gt.attribution * gt.r_squared / np.sum(gt.attribution)   # reference :117

    total = 2 ** samples_log2
    curves, timings = {}, {}
    for method in ("random", "permutohedron", "argsort"):
        for anti in (False, True):
            name = ("a" if anti else "") + method
            nsamp = total // 2 if anti else total          # equal numbers of permutation evaluations
            t0 = time.perf_counter()
            r = L.ls_spa(Xtr, Xte, ytr, yte, method=method, batch_size=2 ** 7, num_batches=nsamp // 2 ** 7,
                         tolerance=0.0, seed=seed, antithetical=anti, return_history=True)
            torch.cuda.synchronize()
            timings[name] = time.perf_counter() - t0
            err = np.linalg.norm(r.attribution_history - gt_attr, axis=1)      # notebook :597-603
            ks = [2 ** e for e in range(5, int(np.log2(nsamp)) + 1)]
            curves[name] = dict(samples=ks, evaluations=[k * (2 if anti else 1) for k in ks],
                                true_error=[float(err[k - 1]) for k in ks],
                                estimated_error=[float(v) for v in r.error_history],
                                estimated_every=2 ** 7, final_true_error=float(err[-1]),
                                r_squared=float(r.r_squared))
    out = dict(p=p, n=n, m=m, ground_truth_pairs=2 ** gt_log2, evaluations_per_method=total,
               ground_truth_error_estimate=float(gt.overall_error), r_squared=float(gt.r_squared),
               seconds=dict(data=t_data, ground_truth=t_gt, **timings), curves=curves)
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
        np.save(os.path.join(out_dir, "gt_Medium.npy"), gt.attribution)
        with open(os.path.join(out_dir, "errors_Medium.json"), "w") as f:
            json.dump(out, f, indent=1)
    if not quiet:
        print(f"p={p} N={n} M={m}: data {t_data:.2f} s, ground truth 2^{gt_log2} pairs {t_gt:.2f} s "
              f"(estimated error {gt.overall_error:.2e}), R^2 {gt.r_squared:.4f}")
        print(f"{'evaluations':>12}" + "".join(f"{k:>16}" for k in curves))
        evs = curves["random"]["evaluations"]
        for ev in evs:
            row = f"{ev:>12}"
            for k, c in curves.items():
                row += f"{c['true_error'][c['evaluations'].index(ev)]:>16.3e}" if ev in c["evaluations"] else f"{'':>16}"
            print(row)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--p", type=int, default=100)
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--m", type=int, default=100_000)
    ap.add_argument("--gt-log2", type=int, default=19)
    ap.add_argument("--samples-log2", type=int, default=13)
    a = ap.parse_args()
    run(a.p, a.n, a.m, a.gt_log2, a.samples_log2, out_dir=os.path.join(ROOT, "experiments", "data"))
