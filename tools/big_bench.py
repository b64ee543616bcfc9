"""Wide-problem lift kernels (lifts_big.cu): throughput at a few widths (development aid)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from quick_bench import synth_problem, ev_time
from ls_spa_b200 import ops, samplers
dev = torch.device("cuda")
widths = [int(a) for a in sys.argv[1:]] or [256, 512, 1000]
for p in widths:
    prob = synth_problem(p, dev)
    count = 592 if p >= 512 else 2368
    perms = samplers.RandomSource(p, 42, None, dev).take(count)
    print(f"p={p} cond_estimate={prob.cond_estimate:.1f} use_chol={prob.use_chol} big={prob.train.big}", flush=True)
    for anti in (False,):
        ms = ev_time(lambda: ops.lifts(prob, perms, anti), reps=2, warm=1)
        prob.check()
        evals = count * (2 if anti else 1)
        T = (p + 64) // 64
        flop = 4.0 / 3.0 * p ** 3
        print(f"   anti={anti}: {ms:.2f} ms for {evals} evaluations -> {evals / ms * 1e3:.0f} evals/s, "
              f"{flop * evals / ms / 1e9:.2f} TFLOP/s on (4/3)p^3, route {ops.LIFT_ROUTE}", flush=True)
