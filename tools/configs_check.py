"""Runs the BASELINE.json configs C1-C3 at full size and a reduced C5; prints wall times (development aid)."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import ls_spa_b200 as L
from oracle import samplers_oracle as so
from oracle import lsspa_oracle as lo

def timed(f):
    torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize(); return r, time.perf_counter() - t

rng = np.random.default_rng(42)
# C2: p=10, N=M=1e5, exact (10! permutations)
Xtr, Xte, ytr, yte, _, _ = so.gen_data(rng, 10, 100_000, 100_000, conditioning=10.0)
r, t = timed(lambda: L.ls_spa(Xtr, Xte, ytr, yte, method="exact"))
r, t = timed(lambda: L.ls_spa(Xtr, Xte, ytr, yte, method="exact"))
print(f"C2 p=10 exact 3628800 perms: {t:.3f} s  -> {3628800/t/1e6:.2f} M perm/s; sum-attr - r2 = {r.attribution.sum()-r.r_squared:.2e}", flush=True)
sub = so.perms_exact(10, 4000, first=1234567)
want = lo.mean_of_lifts(*lo.reduce_data(Xtr, Xte, ytr, yte, 0.0), float(yte @ yte), sub)
got = L.ls_spa(Xtr, Xte, ytr, yte, perms=sub, antithetical=False, tolerance=0.0)
print("   C2 subset parity (4000 perms):", np.max(np.abs(got.attribution - want)) / np.max(np.abs(want)), flush=True)
# C3: p=100, N=M=1e5, argsort 2^7 x 2^7
Xtr, Xte, ytr, yte, _, _ = so.gen_data(np.random.default_rng(42), 100, 100_000, 100_000)
r, t = timed(lambda: L.ls_spa(Xtr, Xte, ytr, yte, method="argsort", batch_size=128, num_batches=128, tolerance=0.0, antithetical=False))
r, t = timed(lambda: L.ls_spa(Xtr, Xte, ytr, yte, method="argsort", batch_size=128, num_batches=128, tolerance=0.0, antithetical=False))
print(f"C3 p=100 argsort 16384 perms (host inputs): {t:.3f} s  overall_error {r.overall_error:.3e} r2 {r.r_squared:.6f}", flush=True)
perms = so.perms_argsort(100, 512, 42)[0]
want = lo.mean_of_lifts(*lo.reduce_data(Xtr, Xte, ytr, yte, 0.0), float(yte @ yte), perms)
got = L.ls_spa(Xtr, Xte, ytr, yte, method="argsort", batch_size=128, num_batches=4, tolerance=0.0, antithetical=False)
print("   C3 parity on the first 512 perms:", np.max(np.abs(got.attribution - want)) / np.max(np.abs(want)), flush=True)
# random method throughput at p=100
r, t = timed(lambda: L.ls_spa(Xtr, Xte, ytr, yte, max_samples=2**15, batch_size=256, tolerance=0.0))
print(f"random p=100 2^15 antithetic samples: {t:.3f} s", flush=True)
# reduced C5: p=1000
Xtr, Xte, ytr, yte, _, _ = so.gen_data(np.random.default_rng(1), 1000, 6000, 5000)
r, t = timed(lambda: L.ls_spa(Xtr, Xte, ytr, yte, method="random", batch_size=4, num_batches=2, tolerance=0.0, antithetical=False))
print(f"C5-reduced p=1000 N=6000 8 perms: {t:.3f} s", flush=True)
perms = so.perms_random(1000, 8, 42)
fac = lo.reduce_data(Xtr, Xte, ytr, yte, 0.0)
want = lo.mean_of_lifts(*fac, float(yte @ yte), perms, lift_fn=lo.square_shapley_lean)
print("   C5-reduced parity:", np.max(np.abs(r.attribution - want)) / np.max(np.abs(want)), "theta", np.max(np.abs(r.theta - np.linalg.lstsq(fac[0], fac[2], rcond=None)[0])), flush=True)
