"""Short driver for ncu: a few launches of the lift kernel at p=100 (8192 permutation evaluations each)."""
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from quick_bench import synth_problem  # noqa: E402
from ls_spa_b200 import ops, samplers  # noqa: E402

p = int(sys.argv[1]) if len(sys.argv) > 1 else 100
count = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda")
prob = synth_problem(p, dev)
perms = samplers.PermutohedronSource(p, 42, None, dev).take(count)
out = torch.empty((count, p), dtype=torch.float64, device=dev)
for _ in range(3):
    ops.lifts(prob, perms, True, out=out)
torch.cuda.synchronize()
print("ok", float(out.sum()))
