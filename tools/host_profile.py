"""cProfile of one device-resident C4 job: where the host time of ls_spa() goes (development aid)."""
import cProfile, pstats, sys, time
import torch
sys.path.insert(0, ".")
import bench
from ls_spa_b200 import api

dev = torch.device("cuda")
Xtr, Xte, ytr, yte = bench.synth_on_device(torch, dev, 1_000_000, 1_000_000, 1000)
kw = dict(reg=1e-2, method="permutohedron", batch_size=128, num_batches=512, tolerance=1e-4, seed=42, antithetical=True)
for _ in range(3):
    api.ls_spa(Xtr, Xte, ytr, yte, **kw)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    api.ls_spa(Xtr, Xte, ytr, yte, **kw)
torch.cuda.synchronize()
print("wall per job ms", (time.perf_counter() - t0) / 3 * 1e3)
pr = cProfile.Profile()
pr.enable()
api.ls_spa(Xtr, Xte, ytr, yte, **kw)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
