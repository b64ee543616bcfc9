"""Every width 49..128: Cholesky route against the Householder route, small and odd launch sizes."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from quick_bench import synth_problem
from ls_spa_b200 import ops, samplers

dev = torch.device("cuda")
worst = 0.0
for p in range(17, 153):
    prob = synth_problem(p, dev, seed=p)
    for count, anti in ((1, False), (3, True), (301, True), (700, False)):
        perms = samplers.ArgsortSource(p, p + count, None, dev).take(count)
        prob.use_chol = True
        a = ops.lifts(prob, perms, anti)
        prob.use_chol = False
        b = ops.lifts(prob, perms, anti)
        err = float((a - b).abs().max() / b.abs().max())
        bad = bool(torch.isnan(a).any())
        worst = max(worst, err)
        if err > 1e-11 or bad:
            print("MISMATCH p", p, "count", count, "anti", anti, err, bad, flush=True)
print("worst scaled difference over p = 17..152:", worst)
from quick_bench import ev_time
for p in (33, 136, 152):
    prob = synth_problem(p, dev, seed=1)
    perms = samplers.ArgsortSource(p, 3, None, dev).take(1 << 13)
    buf = torch.empty((perms.shape[0], p), dtype=torch.float64, device=dev)
    for name, flag in (("scalar/householder", False), ("chol", True)):
        prob.use_chol = flag
        t = ev_time(lambda: ops.lifts(prob, perms, True, out=buf))
        print(f"p={p} {name}: {2 * perms.shape[0] / t * 1e3 / 1e6:.2f} M evals/s", flush=True)
