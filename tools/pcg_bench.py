"""Throughput of the PCG64 permutation source (numpy Generator.permutation stream), development aid."""
import sys
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tools")
from quick_bench import ev_time
from ls_spa_b200 import samplers

dev = torch.device("cuda")
for p, count in ((100, 8192), (100, 65536), (33, 65536), (1000, 2048)):
    src = samplers.RandomSource(p, 42, None, dev)
    got = src.take(count).cpu().numpy()
    rng = np.random.default_rng(42)
    want = np.stack([rng.permutation(p) for _ in range(min(count, 2000))])
    ok = np.array_equal(got[: want.shape[0]], want)
    # chained calls continue the stream
    nxt = src.take(5).cpu().numpy()
    rng2 = np.random.default_rng(42)
    ref_all = np.stack([rng2.permutation(p) for _ in range(count + 5)]) if count <= 8192 else None
    ok2 = ref_all is None or (np.array_equal(got, ref_all[:count]) and np.array_equal(nxt, ref_all[count:]))
    src2 = samplers.RandomSource(p, 7, None, dev)
    t = ev_time(lambda: src2.take(count))
    print(f"p={p} count={count}: bit-exact {ok} chained {ok2}  {t:.3f} ms  {count / t * 1e3 / 1e6:.2f} M perms/s", flush=True)
