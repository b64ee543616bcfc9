"""Timeline of one end-to-end (pinned host inputs) C4 job: copies and kernels in start order, merged into
runs of the same name (development aid).  Usage: python tools/e2e_timeline.py [f32]"""
import sys, torch
sys.path.insert(0, ".")
import bench
import ls_spa_b200 as L
dev = torch.device("cuda")
Xtr, Xte, ytr, yte = bench.synth_on_device(torch, dev, 100, 1_000_000, 1_000_000, bench.SEED)
f32 = len(sys.argv) > 1 and sys.argv[1] == "f32"
host = [(t.float() if f32 else t).cpu().pin_memory() for t in (Xtr, Xte, ytr, yte)]
del Xtr, Xte, ytr, yte
kw = dict(reg=bench.REG, method="permutohedron", batch_size=bench.BATCH, num_batches=512, tolerance=bench.TOL,
          seed=bench.SEED, antithetical=True)
for _ in range(2):
    L.ls_spa(*host, **kw)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    L.ls_spa(*host, **kw)
    torch.cuda.synchronize()
evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
runs = []
for e in evs:
    name = e.name.replace("(anonymous namespace)::", "").split("(")[0][-48:]
    if runs and runs[-1][0] == name and e.time_range.start - runs[-1][2] < 50:
        runs[-1][2] = max(runs[-1][2], e.time_range.end); runs[-1][3] += 1; runs[-1][4] += e.time_range.end - e.time_range.start
    else:
        runs.append([name, e.time_range.start, e.time_range.end, 1, e.time_range.end - e.time_range.start])
for name, a, b, n, busy in runs:
    if b - a > 100 or busy > 100:
        print(f"{(a - t0) / 1e3:8.2f} -> {(b - t0) / 1e3:8.2f} ms  {n:4d} x  busy {busy / 1e3:7.2f} ms  {name}")
print(f"GPU span {(evs[-1].time_range.end - t0) / 1e3:.2f} ms")
