"""Phase timing of one C4 job (device-resident inputs) with CUDA events; development aid."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from ls_spa_b200 import engine, ops, samplers, api

dev = torch.device("cuda")
Xtr, Xte, ytr, yte = bench.synth_on_device(torch, dev, 1_000_000, 1_000_000, 1000)
kw = dict(reg=1e-2, method="permutohedron", batch_size=128, num_batches=512, tolerance=1e-4, seed=42, antithetical=True)
for _ in range(2):
    api.ls_spa(Xtr, Xte, ytr, yte, **kw)
torch.cuda.synchronize()

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

backend, coll = engine.CudaBackend(dev), engine.Collective(None)
t0 = time.perf_counter(); e0 = ev()
prob = engine.reduce_problem(backend, coll, Xtr, Xte, ytr, yte, 1e-2, 100)
e1 = ev()
src = samplers.make_source("permutohedron", 100, 42, 65536, dev)
cfg = engine.JobConfig(p=100, batch_size=128, max_samples=65536, tolerance=1e-4, seed=42, antithetical=True,
                       estimate_errors=True, return_history=False)
est = backend.make_estimator(cfg)
tp = tl = tpa = tab = tsync = 0.0
pos = 0
for sb in range(8):
    a = ev(); perms = src.take(8192); b = ev()
    rows = backend.lifts(prob, perms, True); c = ev()
    desc = [(i * 128, 128, pos + i * 128) for i in range(64)]
    part = est.partials(rows, desc); d = ev()
    ov, ft = est.absorb(part, list(range(64)), [128] * 64, own=(0, 64), emit=True); e = ev()
    errs = ov.cpu().numpy(); f = ev()
    torch.cuda.synchronize()
    tp += a.elapsed_time(b); tl += b.elapsed_time(c); tpa += c.elapsed_time(d); tab += d.elapsed_time(e); tsync += e.elapsed_time(f)
    pos += 8192
e2 = ev()
theta, r2 = backend.theta_r2(prob)
e3 = ev(); torch.cuda.synchronize()
print(f"reduce {e0.elapsed_time(e1):.2f} ms | perms {tp:.2f} lifts {tl:.2f} partials {tpa:.2f} absorb+quant {tab:.2f} errs->host {tsync:.2f} | theta {e2.elapsed_time(e3):.2f} | wall {1e3*(time.perf_counter()-t0):.1f} ms")
t0 = time.perf_counter(); api.ls_spa(Xtr, Xte, ytr, yte, **kw); torch.cuda.synchronize(); print("full job wall", 1e3*(time.perf_counter()-t0))
