import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from ls_spa_b200 import ops
from quick_bench import ev_time
dev = torch.device("cuda")
def serr(a, b): return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
for n, p in ((500, 12), (4097, 64), (3000, 100), (20000, 100), (1000, 119), (257, 33)):
    rng = np.random.default_rng(1)
    X = rng.standard_normal((n, p)) @ (rng.standard_normal((p, p)) * 0.3 + np.eye(p)); y = rng.standard_normal(n)
    Xd, yd = torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev)
    slot, info = ops.cholqr2_factor([(Xd, yd)], p, 2.0)
    q = p + 1
    T = slot[:q*q].view(q, q).cpu().numpy()
    Z = np.column_stack([X, y]) / 2.0
    ref = np.linalg.qr(Z, mode="r")
    print(f"n={n} p={p} gram err {serr(T.T @ T, Z.T @ Z):.2e} |R| err {serr(np.abs(T), np.abs(ref)):.2e} lower {np.abs(np.tril(T,-1)).max():.1e} ysq {float(slot[q*q]):.6f} {(y@y)/4:.6f} info {info.cpu().numpy().ravel()} cond {np.linalg.cond(Z):.1f}")
n, p = 1 << 20, 100
X = torch.randn(n, p, dtype=torch.float64, device=dev); y = torch.randn(n, dtype=torch.float64, device=dev)
t = ev_time(lambda: ops.cholqr2_factor([(X, y)], p, 3.0))
t1 = 0.0
print(f"cholqr2 1M x 100: {t:.3f} ms (pass 1 alone {t1:.3f} ms)")
