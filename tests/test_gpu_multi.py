"""Multi-GPU (NCCL, one process per GPU): sharded rows + sharded batches must reproduce the
single-GPU result.  Skipped when fewer than two GPUs are visible."""

import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import ls_spa_b200 as L
        from oracle import samplers_oracle as so
        rng = np.random.default_rng(11)
        Xtr, Xte, ytr, yte, _, _ = so.gen_data(rng, 40, 3000, 2500)
        out = {}
        for method, anti in (("permutohedron", True), ("random", False), ("argsort", True)):
            r = L.ls_spa(Xtr, Xte, ytr, yte, reg=1e-3, method=method, batch_size=16, num_batches=24,
                         tolerance=0.0, seed=5, antithetical=anti, return_history=True)
            out[method] = (r.attribution, r.theta, float(r.r_squared), r.error_history, r.attribution_history)
        # pre-sharded rows: every rank hands in only its slice
        per_tr, per_te = -(-3000 // world), -(-2500 // world)
        sl_tr = slice(rank * per_tr, min((rank + 1) * per_tr, 3000))
        sl_te = slice(rank * per_te, min((rank + 1) * per_te, 2500))
        r = L.ls_spa(Xtr[sl_tr], Xte[sl_te], ytr[sl_tr], yte[sl_te], reg=1e-3, method="argsort", batch_size=16,
                     num_batches=24, tolerance=0.0, seed=5, antithetical=True, row_sharded=True)
        out["presharded"] = (r.attribution, r.theta, float(r.r_squared), r.error_history, None)
        # early stop must be taken identically on every rank
        tol = float(np.sort(out["argsort"][3])[::-1][5]) * 1.000001
        r = L.ls_spa(Xtr, Xte, ytr, yte, reg=1e-3, method="argsort", batch_size=16, num_batches=24,
                     tolerance=tol, seed=5, antithetical=True)
        out["stopped"] = (r.attribution, r.theta, float(r.r_squared), r.error_history, None)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_two_gpus_match_one():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import ls_spa_b200 as L
    from conftest import scaled_err
    from oracle import samplers_oracle as so
    rng = np.random.default_rng(11)
    Xtr, Xte, ytr, yte, _, _ = so.gen_data(rng, 40, 3000, 2500)
    single = {}
    for method, anti in (("permutohedron", True), ("random", False), ("argsort", True)):
        single[method] = L.ls_spa(Xtr, Xte, ytr, yte, reg=1e-3, method=method, batch_size=16, num_batches=24,
                                  tolerance=0.0, seed=5, antithetical=anti, return_history=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29731, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for method, ref in single.items():
        for rank in (0, 1):
            a, th, r2, eh, hist = got[rank][method]
            assert scaled_err(a, ref.attribution) < 1e-11, (method, rank)
            assert scaled_err(th, ref.theta) < 1e-11
            assert abs(r2 - ref.r_squared) < 1e-12
            assert eh.shape == ref.error_history.shape == (24,)
            np.testing.assert_allclose(eh, ref.error_history, rtol=1e-6)
            assert scaled_err(hist, ref.attribution_history) < 1e-11
        assert np.array_equal(got[0][method][0], got[1][method][0])      # replicated state
    ref = single["argsort"]
    for rank in (0, 1):
        assert scaled_err(got[rank]["presharded"][0], ref.attribution) < 1e-11
        assert got[rank]["stopped"][3].shape == got[0]["stopped"][3].shape
    stop_hist = got[0]["stopped"][3]
    assert 1 <= len(stop_hist) < 24 and np.array_equal(got[0]["stopped"][0], got[1]["stopped"][0])
