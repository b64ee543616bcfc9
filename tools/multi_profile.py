"""Kernel / collective totals of one multi-GPU bench step on rank 0 (torch.profiler; development aid).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_profile.py"""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, ".")
import bench
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
import ls_spa_b200 as L
rows = bench.N_ROWS // world
Xtr, Xte, ytr, yte = bench.synth_on_device(torch, dev, bench.P, rows, rows, 1000 + rank, dist if world > 1 else None, bench.N_ROWS)
kw = dict(reg=bench.REG, method="permutohedron", batch_size=bench.BATCH, num_batches=bench.BATCHES_PER_GPU * world,
          tolerance=bench.TOL, seed=bench.SEED, antithetical=True, row_sharded=world > 1)
for _ in range(3):
    L.ls_spa(Xtr, Xte, ytr, yte, **kw)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        L.ls_spa(Xtr, Xte, ytr, yte, **kw)
    torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    if evs:
        t0 = min(e.time_range.start for e in evs); t1 = max(e.time_range.end for e in evs)
        busy = sum(e.time_range.end - e.time_range.start for e in evs)
        print(f"GPU span {(t1 - t0) / 2e3:.2f} ms per step, kernel-busy {busy / 2e3:.2f} ms per step")
if world > 1:
    dist.destroy_process_group()
