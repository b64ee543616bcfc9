import sys, time, torch
sys.path.insert(0, ".")
import bench
import ls_spa_b200 as L
dev = torch.device("cuda")
Xtr, Xte, ytr, yte = bench.synth_on_device(torch, dev, 100, 1_000_000, 1_000_000, bench.SEED)
kw = dict(reg=bench.REG, method="permutohedron", batch_size=bench.BATCH, num_batches=512, tolerance=bench.TOL, seed=bench.SEED, antithetical=True)
host = [t.cpu().pin_memory() for t in (Xtr, Xte, ytr, yte)]
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); L.ls_spa(*host, **kw); torch.cuda.synchronize(); print("f64", i, round((time.perf_counter() - t0) * 1e3, 2))
host32 = [t.float().cpu().pin_memory() for t in (Xtr, Xte, ytr, yte)]
for i in range(7):
    torch.cuda.synchronize(); t0 = time.perf_counter(); L.ls_spa(*host32, **kw); torch.cuda.synchronize(); print("f32", i, round((time.perf_counter() - t0) * 1e3, 2), torch.cuda.memory_reserved() >> 20)
