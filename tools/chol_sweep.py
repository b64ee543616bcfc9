"""Every width 49..128: Cholesky route against the Householder route, small and odd launch sizes."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from quick_bench import synth_problem
from ls_spa_b200 import ops, samplers

dev = torch.device("cuda")
worst = 0.0
for p in range(49, 129):
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
print("worst scaled difference over p = 49..128:", worst)
