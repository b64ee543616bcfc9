import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from quick_bench import ev_time
from ls_spa_b200 import ops
dev = torch.device("cuda")
n, p = 1 << 20, 100
X = torch.randn(n, p, dtype=torch.float64, device=dev); y = torch.randn(n, dtype=torch.float64, device=dev)
t_rows = ev_time(lambda: ops.tsqr_rows(X, y, 3.0))
parts = ops.tsqr_rows(X, y, 3.0)
t_merge = ev_time(lambda: ops.tsqr_merge(parts, p))
print(f"parts {parts.shape[0]} rows {t_rows:.3f} ms merge {t_merge:.3f} ms total {t_rows+t_merge:.3f}")
