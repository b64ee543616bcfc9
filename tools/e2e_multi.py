import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, ".")
import bench
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import ls_spa_b200 as L
rows = bench.N_ROWS // world
Xtr, Xte, ytr, yte = bench.synth_on_device(torch, dev, bench.P, rows, rows, 1000 + rank, dist, bench.N_ROWS)
host = [t.cpu().pin_memory() for t in (Xtr, Xte, ytr, yte)]
kw = dict(reg=bench.REG, method="permutohedron", batch_size=bench.BATCH, num_batches=bench.BATCHES_PER_GPU * world,
          tolerance=bench.TOL, seed=bench.SEED, antithetical=True, row_sharded=True)
for _ in range(2): L.ls_spa(*host, **kw)
for rep in range(2):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(8): L.ls_spa(*host, **kw)
    dist.barrier(); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 8
    if rank == 0: print(f"HOST_OVERLAP={os.environ.get('LSSPA_HOST_OVERLAP','1')} e2e {dt*1e3:.2f} ms/step = {2*131072/dt/1e6:.2f} M/s", flush=True)
dist.destroy_process_group()
