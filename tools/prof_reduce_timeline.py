import sys, torch
sys.path.insert(0, ".")
from ls_spa_b200 import engine
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda"); n, p = 1_000_000, 100
g = torch.Generator(device=dev).manual_seed(0)
X = torch.randn(n, p, generator=g, device=dev, dtype=torch.float64)
y = X @ torch.randn(p, generator=g, device=dev, dtype=torch.float64) + torch.randn(n, generator=g, device=dev, dtype=torch.float64)
backend, coll = engine.CudaBackend(dev), engine.Collective(None)
for _ in range(3):
    prob = engine.reduce_problem(backend, coll, X, X, y, y, 1e-2, p, n_train_global=n)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        prob = engine.reduce_problem(backend, coll, X, X, y, y, 1e-2, p, n_train_global=n)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type.name == "CUDA"]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev[: len(ev) // 3]:
    print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f}  {e.name[:70]}")
