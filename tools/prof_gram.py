import sys, torch
sys.path.insert(0, ".")
from ls_spa_b200 import ops
dev = torch.device("cuda")
n, p = 1 << 19, 100
X = torch.randn(n, p, dtype=torch.float64, device=dev); y = torch.randn(n, dtype=torch.float64, device=dev)
for _ in range(2):
    slot, info = ops.cholqr2_factor([(X, y)], p, 3.0)
torch.cuda.synchronize()
print("ok", info.cpu().numpy().ravel())
