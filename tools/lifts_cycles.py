"""Cycle counters of one CTA of the DMMA lift kernel (development aid).
Needs a library built with the probes: LSSPA_EXTRA_NVCC_FLAGS=-DLSSPA_LIFTS_TIMING python -m ls_spa_b200.build --force"""
import ctypes, sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from quick_bench import synth_problem
from ls_spa_b200 import ops, samplers, _cabi
dev = torch.device("cuda")
p = int(sys.argv[1]) if len(sys.argv) > 1 else 100
prob = synth_problem(p, dev)
if len(sys.argv) > 2:
    prob.use_chol = sys.argv[2] == "chol"
print("route:", "chol" if prob.use_chol else "householder")
perms = samplers.PermutohedronSource(p, 42, None, dev).take(4096)
buf = torch.zeros(64, dtype=torch.int64, device=dev)
lib = _cabi.load()
lib.lsspa_debug_set_lifts_counters.argtypes = [ctypes.c_void_p]
ops.lifts(prob, perms, True)
lib.lsspa_debug_set_lifts_counters(buf.data_ptr())
ops.lifts(prob, perms, True)
torch.cuda.synchronize()
lib.lsspa_debug_set_lifts_counters(None)
d = buf.cpu().numpy().reshape(8, 8).copy()
print("warp  gather   panel|diag  trailing|accumulate  barrier-wait  phase1  phase1.5+2   (packed kernel: gather, work, wait1, scale+finish, wait2, phase1, phase2, tail-wait)")
for w in range(8):
    print(w, d[w, :8])
