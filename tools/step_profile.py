"""Kernel totals of one device-resident bench step (C4: p=100, N=M=10^6, permutohedron, antithetic,
2^17 evaluations, tolerance 1e-4), from torch.profiler -- which kernels the step is made of."""
import sys, torch
sys.path.insert(0, ".")
import bench
import ls_spa_b200 as L
dev = torch.device("cuda")
Xtr, Xte, ytr, yte = bench.synth_on_device(torch, dev, 100, 1_000_000, 1_000_000, bench.SEED)
kw = dict(reg=bench.REG, method="permutohedron", batch_size=bench.BATCH, num_batches=512, tolerance=bench.TOL,
          seed=bench.SEED, antithetical=True)
for _ in range(2):
    L.ls_spa(Xtr, Xte, ytr, yte, **kw)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    L.ls_spa(Xtr, Xte, ytr, yte, **kw)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=24, max_name_column_width=60))
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
t0 = min(e.time_range.start for e in evs); t1 = max(e.time_range.end for e in evs)
print(f"GPU span {(t1 - t0) / 1e3:.2f} ms, kernel-busy {sum(e.time_range.end - e.time_range.start for e in evs) / 1e3:.2f} ms")
