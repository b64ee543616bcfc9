import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from ls_spa_b200 import ops, engine, samplers
from oracle import samplers_oracle as so, lsspa_oracle as lo
dev = torch.device("cuda")
def serr(a, b): return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
for p, n, m, reg in ((33, 600, 500, 1e-2), (100, 1500, 1200, 0.0), (33, 600, 500, 0.0), (40, 128, 64, 0.0), (40, 129, 64, 0.0)):
    rng = np.random.default_rng(1)
    X = rng.standard_normal((n, p)); y = rng.standard_normal(n)
    Xd, yd = torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev)
    parts = ops.tsqr_rows(Xd, yd, 2.0)
    q = p + 1
    Z = np.column_stack([X, y]) / 2.0
    nparts = parts.shape[0]
    per = -(-(-(-n // nparts)) // 32) * 32
    errs = []
    for c in range(nparts):
        T = parts[c, :q*q].view(q, q).cpu().numpy()
        Zc = Z[c*per:min((c+1)*per, n)]
        errs.append(serr(T.T @ T, Zc.T @ Zc) if len(Zc) else 0.0)
    print(f"p={p} n={n} nparts={nparts} per={per} part errs max={max(errs):.2e}", ["%.1e" % e for e in errs][:12])
    slot = ops.tsqr_merge(parts, p)
    T = slot[:q*q].view(q, q).cpu().numpy()
    print("   merged err", serr(T.T @ T, Z.T @ Z), "ysq", float(slot[q*q]), (y @ y) / 4.0)
    if reg:
        both = torch.stack([slot, ops.ridge_factor(p, reg, dev)], 0)
        s2 = ops.tsqr_merge(both, p, group=2)
        T2 = s2[:q*q].view(q, q).cpu().numpy()
        D = np.zeros((q, q)); D[:p, :p] = reg * np.eye(p)
        print("   ridge err", serr(T2.T @ T2, Z.T @ Z + D), "ridge diag", ops.ridge_factor(p, reg, dev)[:q*q].view(q,q).diagonal()[:3].cpu().numpy())

# job pieces at p=100
rng = np.random.default_rng(42)
Xtr, Xte, ytr, yte, _, _ = so.gen_data(rng, 100, 1500, 1200)
backend = engine.CudaBackend(dev); coll = engine.Collective(None)
prob = engine.reduce_problem(backend, coll, Xtr, Xte, ytr, yte, 0.0, 100)
R_tr, R_te, c_tr, c_te = lo.reduce_data(Xtr, Xte, ytr, yte, 0.0)
print("ynsq dev", prob.y_norm_sq, "ref", float(yte @ yte))
perms = so.perms_argsort(100, 32, 42)[0]
pd = torch.from_numpy(perms.astype(np.int32)).to(dev)
rows = ops.lifts(prob, pd, False).cpu().numpy()
want = np.array([lo.square_shapley(R_tr, R_te, c_tr, c_te, float(yte @ yte), pm) for pm in perms])
print("lifts (device-reduced) err", serr(rows, want), "rows sum", rows.sum(1)[:3], want.sum(1)[:3])
est = ops.Estimator(100, 8, 0.0, 1, True, dev)
rd = torch.from_numpy(rows).to(dev)
desc = [(0, 8, 0), (8, 8, 8), (16, 8, 16), (24, 8, 24)]
part = est.partials(rd, desc)
pv = part[:, :8].cpu().numpy(); print("partial hdr", pv[:, 0])
print("partial mean err", serr(part[0, 8:108].cpu().numpy(), rows[:8].mean(0)))
est.update(part, 4)
out = est.read(want_cov=True)
print("est count", out["count"], "mean err", serr(out["mean"], rows.mean(0)), "cov err", serr(out["cov"], np.cov(rows, rowvar=False, bias=True)), "hist", out["error_history"])
th, r2 = ops.theta_r2(prob)
thr = np.linalg.lstsq(R_tr, c_tr, rcond=None)[0]
print("theta err", serr(th.cpu().numpy(), thr), "r2", float(r2))
