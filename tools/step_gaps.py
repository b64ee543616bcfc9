"""Idle gaps (> 15 us) between consecutive GPU activities of one device-resident bench step (development aid)."""
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
evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
t0, end, tot = evs[0].time_range.start, evs[0].time_range.end, 0.0
prev = evs[0]
for e in evs[1:]:
    gap = e.time_range.start - end
    if gap > 15:
        tot += gap
        print(f"{(end - t0) / 1e3:8.3f} ms  gap {gap:7.1f} us  after {prev.name.replace('(anonymous namespace)::', '').split('(')[0][-40:]:40s} before {e.name.replace('(anonymous namespace)::', '').split('(')[0][-40:]}")
    if e.time_range.end > end:
        end, prev = e.time_range.end, e
print(f"span {(end - t0) / 1e3:.2f} ms, gaps > 15 us total {tot / 1e3:.2f} ms")
