"""Split Cholesky route (factor | eliminate) against the fused kernel; timings.  Development aid."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from quick_bench import ev_time, synth_problem
from ls_spa_b200 import ops, samplers

dev = torch.device("cuda")
for p in (49, 56, 72, 100, 104, 128):
    prob = synth_problem(p, dev, seed=p)
    for count, anti in ((5, False), (700, True)):
        perms = samplers.ArgsortSource(p, 9, None, dev).take(count)
        prob.use_chol = True
        a = ops.lifts(prob, perms, anti)
        fac = ops.lifts_factor(prob.train, perms, anti)
        b = ops.lifts_eliminate(prob, fac, perms, anti)
        err = float((a - b).abs().max() / a.abs().max())
        print("p", p, "count", count, "anti", anti, "split-vs-fused", f"{err:.2e}", "nan", bool(torch.isnan(b).any()), flush=True)
p = 100
prob = synth_problem(p, dev)
perms = samplers.PermutohedronSource(p, 42, None, dev).take(1 << 14)
n = 2 * perms.shape[0]
t0 = ev_time(lambda: ops.lifts(prob, perms, True))
fac = ops.lifts_factor(prob.train, perms, True)
t1 = ev_time(lambda: ops.lifts_factor(prob.train, perms, True))
t2 = ev_time(lambda: ops.lifts_eliminate(prob, fac, perms, True))
print(f"fused {t0:.3f} ms ({n / t0 / 1e3:.2f} M evals/s) | factor {t1:.3f} ms ({n / t1 / 1e3:.2f} M/s) | eliminate {t2:.3f} ms ({n / t2 / 1e3:.2f} M/s)")
