import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from quick_bench import synth_problem
from ls_spa_b200 import ops
dev = torch.device("cuda")
for p in (20, 50, 100, 128, 140):
    prob = synth_problem(p, dev, seed=p)
    base = (p + 1) ** 2
    info = prob.gram[base:base + 4].cpu().numpy()
    R = prob.R_tr_cm.t().cpu().numpy()
    Rp = R / np.linalg.norm(R, axis=0)
    print(p, "cond2", round(float(np.linalg.cond(Rp)), 1), "info", [round(float(v), 2) for v in info], "frob numpy", round(float(np.linalg.norm(Rp) * np.linalg.norm(np.linalg.inv(Rp))), 2))
