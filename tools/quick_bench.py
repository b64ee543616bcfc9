"""Kernel-level timings (CUDA events) of the hot-path stages; development aid."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from ls_spa_b200 import ops, samplers  # noqa: E402


def ev_time(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def synth_problem(p, dev, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = torch.randn(4 * p, p, generator=g, dtype=torch.float64)
    mix = torch.randn(p, max(p // 20, 1), generator=g, dtype=torch.float64)
    cov = mix @ mix.T + torch.eye(p, dtype=torch.float64)
    Lc = torch.linalg.cholesky(cov / cov.diagonal().sqrt().outer(cov.diagonal().sqrt()))
    X = A @ Lc.T
    R = torch.linalg.qr(X / np.sqrt(4 * p), mode="r").R
    R2 = torch.linalg.qr((torch.randn(4 * p, p, generator=g, dtype=torch.float64) @ Lc.T), mode="r").R
    c1 = torch.randn(p, generator=g, dtype=torch.float64)
    c2 = torch.randn(p, generator=g, dtype=torch.float64)
    return ops.ReducedProblem(R.to(dev), c1.to(dev), R2.to(dev), c2.to(dev), float(c2 @ c2) * 1.3)


def main():
    dev = torch.device("cuda")
    out = {}
    for p, count in ((10, 1 << 18), (33, 1 << 16), (100, 1 << 15), (117, 1 << 14), (160, 1 << 11), (256, 1 << 10)):
        prob = synth_problem(p, dev)
        src = samplers.ArgsortSource(p, 42, None, dev)
        t_gen = ev_time(lambda: (setattr(src, "position", 0), src.take(count)))
        src.position = 0
        perms = src.take(count)
        buf = torch.empty((count, p), dtype=torch.float64, device=dev)
        t = ev_time(lambda: ops.lifts(prob, perms, False, out=buf))
        flops = 7.0 / 3.0 * p ** 3 * count
        out[f"lifts_p{p}"] = dict(count=count, ms=t, perms_per_s=count / t * 1e3, tflops=flops / t / 1e9,
                                  gen_argsort_ms=t_gen)
        print(p, out[f"lifts_p{p}"], flush=True)
    for p, n in ((10, 1 << 20), (100, 1 << 20), (100, 1 << 17)):
        X = torch.randn(n, p, dtype=torch.float64, device=dev)
        y = torch.randn(n, dtype=torch.float64, device=dev)
        t_rows = ev_time(lambda: ops.tsqr_rows(X, y, 3.0))
        parts = ops.tsqr_rows(X, y, 3.0)
        t_merge = ev_time(lambda: ops.tsqr_merge(parts, p))
        gb = 8.0 * n * (p + 1) / 1e9
        out[f"tsqr_p{p}_n{n}"] = dict(rows_ms=t_rows, merge_ms=t_merge, parts=int(parts.shape[0]),
                                      gbps=gb / ((t_rows + t_merge) * 1e-3))
        print(p, n, out[f"tsqr_p{p}_n{n}"], flush=True)
    for p, count in ((100, 1 << 15),):
        for name, cls in (("random", samplers.RandomSource), ("permutohedron", samplers.PermutohedronSource)):
            src = cls(p, 42, None, dev)
            t = ev_time(lambda: src.take(count))
            out[f"gen_{name}_p{p}"] = dict(count=count, ms=t, perms_per_s=count / t * 1e3)
            print(name, out[f"gen_{name}_p{p}"], flush=True)
    # estimator: 32 batches of 256 rows at p=100
    p = 100
    rows = torch.randn(32 * 256, p, dtype=torch.float64, device=dev)
    est = ops.Estimator(p, 4096, 0.0, 1, True, dev)
    desc = [(b * 256, 256, b * 256) for b in range(32)]
    t_part = ev_time(lambda: est.partials(rows, desc))
    part = est.partials(rows, desc)
    t_upd = ev_time(lambda: est.update(part, 32))
    out["estimator_p100_32x256"] = dict(partials_ms=t_part, update_ms=t_upd)
    print(out["estimator_p100_32x256"], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
