"""Kernel totals of one GPU's share of BASELINE config 5 (p=1000, N=M=125000, 8192 random permutations)."""
import sys, torch
sys.path.insert(0, ".")
import bench
import ls_spa_b200 as L
dev = torch.device("cuda")
Xtr, Xte, ytr, yte = bench.synth_on_device(torch, dev, 1000, 125_000, 125_000, 2000)
kw = dict(method="random", batch_size=128, num_batches=64, tolerance=0.0, antithetical=False)
L.ls_spa(Xtr, Xte, ytr, yte, **kw)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    L.ls_spa(Xtr, Xte, ytr, yte, **kw)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
t0 = min(e.time_range.start for e in evs); t1 = max(e.time_range.end for e in evs)
print(f"GPU span {(t1 - t0) / 1e3:.1f} ms, kernel-busy {sum(e.time_range.end - e.time_range.start for e in evs) / 1e3:.1f} ms")
