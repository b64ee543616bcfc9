"""Small end-to-end runs for compute-sanitizer (racecheck / memcheck): every kernel family once."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import ls_spa_b200 as L
from oracle import samplers_oracle as so
rng = np.random.default_rng(0)
for p, n, m in ((12, 300, 200), (60, 400, 300), (100, 500, 300)):
    Xtr, Xte, ytr, yte, _, _ = so.gen_data(rng, p, n, m, conditioning=max(p / 5.0, 1.0))
    for method in ("random", "argsort", "permutohedron"):
        r = L.ls_spa(Xtr, Xte, ytr, yte, reg=1e-3, method=method, batch_size=4, num_batches=2, tolerance=0.0,
                     return_history=True)
        assert abs(r.attribution.sum() - r.r_squared) < 1e-9
Xtr, Xte, ytr, yte, _, _ = so.gen_data(rng, 5, 100, 80, conditioning=5.0)
r = L.ls_spa(Xtr, Xte, ytr, yte)
print("ok", r.r_squared)
