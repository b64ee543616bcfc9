"""Short driver for ncu / timing: the reduction of one C4-sized side (10^6 x 100 rows), a few times."""
import sys, time
import torch
sys.path.insert(0, ".")
from ls_spa_b200 import engine, ops
dev = torch.device("cuda")
n, p = 1_000_000, int(sys.argv[1]) if len(sys.argv) > 1 else 100
g = torch.Generator(device=dev).manual_seed(0)
X = torch.randn(n, p, generator=g, device=dev, dtype=torch.float64)
y = X @ torch.randn(p, generator=g, device=dev, dtype=torch.float64) + torch.randn(n, generator=g, device=dev, dtype=torch.float64)
backend, coll = engine.CudaBackend(dev), engine.Collective(None)
for _ in range(3):
    prob = engine.reduce_problem(backend, coll, X, X, y, y, 1e-2, p, n_train_global=n)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    prob = engine.reduce_problem(backend, coll, X, X, y, y, 1e-2, p, n_train_global=n)
e1.record()
torch.cuda.synchronize()
print(f"reduce_problem (both sides, p={p}, N=M={n}): {e0.elapsed_time(e1) / 5:.3f} ms; cond {prob.cond_estimate:.1f} chol {prob.use_chol}")
base = (p + 1) * (p + 1)
if prob.gram is not None:
    print("chol_factor phase cycles (factor, inverse, bounds):", prob.gram[base + 4: base + 7].tolist())
