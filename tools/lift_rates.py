"""Lift-kernel rates at a few widths, Cholesky route, antithetic pairs (development aid).
LSSPA_CHOL_PACKED=0 selects the eight-warp kernel for A/B timing."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from quick_bench import synth_problem, ev_time
from ls_spa_b200 import ops, samplers
dev = torch.device("cuda")
widths = [int(a) for a in sys.argv[1:]] or [56, 64, 80, 100, 104, 112, 128]
for p in widths:
    prob = synth_problem(p, dev)
    prob.use_chol = True
    count = 1 << 15
    perms = samplers.PermutohedronSource(p, 42, None, dev).take(count)
    buf = torch.empty((count, p), dtype=torch.float64, device=dev)
    ms = ev_time(lambda: ops.lifts(prob, perms, True, out=buf), reps=3, warm=1)
    evals = 2 * count
    print(f"p={p}: {ms:.2f} ms for {evals} evaluations -> {evals / ms / 1e3:.2f} M evals/s, "
          f"{4.0 / 3.0 * p ** 3 * evals / ms / 1e9:.2f} TFLOP/s executed, {7.0 / 3.0 * p ** 3 * evals / ms / 1e9:.2f} on (7/3)p^3", flush=True)
