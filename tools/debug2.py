import sys, numpy as np, torch
sys.path.insert(0, ".")
from ls_spa_b200 import ops
dev = torch.device("cuda")
def serr(a, b): return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
for n, p in ((4097, 64), (1000, 17), (4097, 100), (9000, 33), (257, 130), (2000, 300)):
    rng = np.random.default_rng(1)
    X = rng.standard_normal((n, p)); y = rng.standard_normal(n)
    parts = ops.tsqr_rows(torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev), 2.0)
    q = p + 1
    Z = np.column_stack([X, y]) / 2.0
    nparts = parts.shape[0]
    rb = 64 if q <= 256 else 32
    per = -(-(-(-n // nparts)) // rb) * rb
    errs = []
    for c in range(nparts):
        T = parts[c, :q*q].view(q, q).cpu().numpy()
        Zc = Z[c*per:min((c+1)*per, n)]
        errs.append(serr(T.T @ T, Zc.T @ Zc) if len(Zc) else float(np.abs(T).max()))
    print(f"p={p} n={n} nparts={nparts} per={per} part errs max={max(errs):.2e}", ["%.1e" % e for e in errs][:20])
    pp = parts
    lvl = 0
    while pp.shape[0] > 1:
        cnt = pp.shape[0]
        nout = (cnt + 7) // 8
        out = torch.empty((nout, pp.shape[1]), dtype=torch.float64, device=dev)
        from ls_spa_b200 import _cabi
        _cabi.check(_cabi.load().lsspa_tsqr_merge(pp.data_ptr(), cnt, 8, p, out.data_ptr(), 0))
        torch.cuda.synchronize()
        errs = []
        for g in range(nout):
            T = out[g, :q*q].view(q, q).cpu().numpy()
            S = sum(pp[t, :q*q].view(q, q).cpu().numpy().T @ pp[t, :q*q].view(q, q).cpu().numpy() for t in range(8*g, min(8*g+8, cnt)))
            errs.append(serr(T.T @ T, S))
        lvl += 1
        print(f"   merge level {lvl}: {cnt} -> {nout} errs", ["%.1e" % e for e in errs])
        pp = out
