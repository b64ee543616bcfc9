"""CPU, world_size 2, gloo: the sharding / collective logic of ls_spa_b200.engine
(row-sharded reduction, per-rank runs of batches, gathered partial moments, replicated
stop decision) with an ORACLE-backed stand-in for the CUDA backend.  The product has no CPU
compute path; this stand-in exists only here, to exercise the host logic without a GPU."""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ls_spa_b200 import engine  # noqa: E402
from ls_spa_b200.samplers import ExplicitSource, PermutationSource  # noqa: E402
from oracle import lsspa_oracle as lo  # noqa: E402
from oracle import samplers_oracle as so  # noqa: E402


class FakeProblem:
    pass


class FakeEstimator:
    """Same interface as ops.Estimator (partials / snapshot / restore / absorb / read)."""

    def __init__(self, p):
        self.p = p
        self.count, self.mean, self.cov = 0, np.zeros(p), np.zeros((p, p))
        self.partial_doubles = 1 + p + p * p

    def partials(self, rows, desc):
        out = torch.zeros((len(desc), self.partial_doubles), dtype=torch.float64)
        r = rows.numpy()
        for b, (first, n, _) in enumerate(desc):
            blk = r[first:first + n]
            if n:
                out[b, 0] = n
                out[b, 1:1 + self.p] = torch.from_numpy(blk.mean(0))
                out[b, 1 + self.p:] = torch.from_numpy(np.cov(blk, rowvar=False, bias=True).reshape(-1))
        return out

    def snapshot(self):
        return self.count, self.mean.copy(), self.cov.copy()

    def scratch(self):
        if not hasattr(self, "_scratch"):
            self._scratch = FakeEstimator(self.p)
        return self._scratch

    def reset(self):
        self.count, self.mean, self.cov = 0, np.zeros(self.p), np.zeros((self.p, self.p))

    def copy_from(self, other):
        self.count, self.mean, self.cov = other.count, other.mean.copy(), other.cov.copy()

    def export_block(self):
        blk = torch.zeros(self.partial_doubles, dtype=torch.float64)
        blk[0] = self.count
        blk[1:1 + self.p] = torch.from_numpy(self.mean)
        blk[1 + self.p:] = torch.from_numpy(self.cov.reshape(-1))
        return blk

    def restore(self, snap):
        self.count, self.mean, self.cov = snap[0], snap[1].copy(), snap[2].copy()

    def absorb(self, partials, slots, counts, own=(0, 0), emit=False):
        flat = partials.reshape(-1, self.partial_doubles).numpy()
        errs = []
        for b, slot in enumerate(slots):
            blk = flat[slot]
            n2, m2, c2 = int(blk[0]), blk[1:1 + self.p], blk[1 + self.p:].reshape(self.p, self.p)
            assert n2 == counts[b]
            self.cov = lo.merge_sample_cov(self.mean, m2, self.cov, c2, self.count, n2)
            self.mean = lo.merge_sample_mean(self.mean, m2, self.count, n2)
            self.count += n2
            if emit and own[0] <= b < own[1]:
                errs.append(float(np.sqrt(np.trace(self.cov) / max(self.count - 1, 1))))   # deterministic stand-in
        if emit and own[1] > own[0]:
            return torch.tensor(errs, dtype=torch.float64), torch.zeros((len(errs), self.p), dtype=torch.float64)
        return None, None

    def read(self):
        return dict(count=self.count, mean=self.mean.copy())


class OracleBackend:
    """Same interface as engine.CudaBackend, numpy/oracle arithmetic, CPU tensors."""
    device = torch.device("cpu")

    def _slot(self, T, ysq, p):
        q = p + 1
        s = torch.zeros(q * q + 8, dtype=torch.float64)
        s[:q * q] = torch.from_numpy(np.ascontiguousarray(T)).reshape(-1)
        s[q * q] = ysq
        return s

    def reduce_rows(self, X, y, lo_, hi, p, divisor):
        Z = np.column_stack([np.asarray(X)[lo_:hi], np.asarray(y)[lo_:hi]]) / divisor
        T = np.zeros((p + 1, p + 1))
        if len(Z):
            r = np.linalg.qr(Z, mode="r")
            T[:r.shape[0]] = r
        return self._slot(T, float((Z[:, p] ** 2).sum()), p)

    def merge_factors(self, factors, p):
        q = p + 1
        stack = np.vstack([f[:q * q].numpy().reshape(q, q) for f in factors])
        r = np.linalg.qr(stack, mode="r")
        return self._slot(r, float(sum(f[q * q] for f in factors)), p)

    def ridge(self, p, reg):
        T = np.zeros((p + 1, p + 1))
        T[:p, :p] = np.sqrt(reg) * np.eye(p)
        return self._slot(T, 0.0, p)

    def make_problem(self, tr, te, p):
        q = p + 1
        pr = FakeProblem()
        Ttr, Tte = tr[:q * q].numpy().reshape(q, q), te[:q * q].numpy().reshape(q, q)
        pr.R_tr, pr.c_tr, pr.R_te, pr.c_te, pr.ynsq = Ttr[:p, :p], Ttr[:p, p], Tte[:p, :p], Tte[:p, p], float(te[q * q])
        return pr

    def lifts(self, prob, perms, anti):
        out = []
        for pm in perms.numpy():
            l = lo.square_shapley(prob.R_tr, prob.R_te, prob.c_tr, prob.c_te, prob.ynsq, pm)
            if anti:
                l = (l + lo.square_shapley(prob.R_tr, prob.R_te, prob.c_tr, prob.c_te, prob.ynsq, pm[::-1])) / 2
            out.append(l)
        return torch.from_numpy(np.array(out).reshape(-1, perms.shape[1]))

    def lifts_eliminate(self, prob, factors, perms, anti):
        assert factors == "factored"               # stand-in for the stored factor blocks
        return self.lifts(prob, perms, anti)

    def make_estimator(self, cfg):
        return FakeEstimator(cfg.p)

    def prefix_means(self, rows, carry_sum, carry_count):
        c = carry_sum.numpy() + np.cumsum(rows.numpy(), 0)
        out = c / (carry_count + np.arange(1, len(c) + 1))[:, None]
        if len(c):
            carry_sum.copy_(torch.from_numpy(c[-1]))
        return torch.from_numpy(out)

    def zeros(self, *shape):
        return torch.zeros(shape, dtype=torch.float64)


class RangeSource(PermutationSource):
    """Random-access stand-in (like the exact / Sobol sources): permutation k of a fixed table."""
    method, random_access = "table", True

    def __init__(self, table):
        super().__init__(table.shape[1], table.shape[0])
        self.table = torch.from_numpy(table.astype(np.int32))

    def take(self, count):
        count = self._clip(count)
        out = self.table[self.position:self.position + count]
        self.position += count
        return out


def problem(seed=3, p=12, n=240, m=200):
    rng = np.random.default_rng(seed)
    return so.gen_data(rng, p, n, m, conditioning=4.0)[:4]


def job(rank, world, port, kind, anti, tol, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Xtr, Xte, ytr, yte = problem()
        p = Xtr.shape[1]
        perms = so.perms_random(p, 45, 9)
        coll = engine.Collective(None)
        assert coll.world == world and coll.rank == rank
        backend = OracleBackend()
        prob = engine.reduce_problem(backend, coll, Xtr, Xte, ytr, yte, 0.05, p)
        src = RangeSource(perms) if kind == "table" else ExplicitSource(p, iter(list(perms)), torch.device("cpu"))
        cfg = engine.JobConfig(p=p, batch_size=4, max_samples=None, tolerance=tol, seed=1, antithetical=anti,
                               estimate_errors=True, return_history=True)
        import ls_spa_b200.engine as E
        E.target_samples = lambda p: 8            # small super-batches: several rounds of gathers
        res, hist, done = engine.run_samples(backend, coll, prob, src, cfg)
        if rank == 0:
            out.put((res["mean"], res["count"], res["error_history"], hist, done))
        else:
            out.put((res["mean"], res["count"]))
    finally:
        dist.destroy_process_group()


def run_world(world, kind, anti, tol, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=job, args=(r, world, port, kind, anti, tol, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    return got


@pytest.mark.parametrize("kind,anti", [("table", True), ("stream", False)])
def test_two_ranks_match_single_process_oracle(kind, anti):
    Xtr, Xte, ytr, yte = problem()
    perms = so.perms_random(12, 45, 9)
    want, lifts = lo.ls_spa_reference_loop(Xtr, Xte, ytr, yte, reg=0.05, perms=list(perms), tolerance=0.0,
                                           batch_size=4, antithetical=anti, return_attribution_history=True,
                                           return_lifts=True)
    got = run_world(2, kind, anti, 0.0, 29611 + (1 if anti else 0))
    full = [g for g in got if len(g) == 5][0]
    other = [g for g in got if len(g) == 2][0]
    mean, count, err_hist, hist, done = full
    assert count == 45 and done == 45
    assert np.max(np.abs(mean - want.attribution)) < 1e-12
    assert np.max(np.abs(other[0] - mean)) == 0.0          # replicated state: identical on both ranks
    assert err_hist.shape == want.error_history.shape == (12,)     # 11 full batches + the short one
    assert np.max(np.abs(hist - want.attribution_history)) < 1e-12


def test_two_ranks_early_stop_is_replicated():
    got = run_world(2, "table", False, 0.0, 29621)
    errs = [g for g in got if len(g) == 5][0][2]
    tol = float(np.sort(errs)[::-1][4]) * 1.000001
    stop_at = int(np.argmax(errs < tol)) + 1
    got = run_world(2, "table", False, tol, 29622)
    full = [g for g in got if len(g) == 5][0]
    other = [g for g in got if len(g) == 2][0]
    assert full[1] == other[1] == 4 * stop_at
    assert len(full[2]) == stop_at and full[3].shape[0] == 4 * stop_at


class FakePre:
    """Stand-in for engine.Prefactor: permutations of the first super-batches drawn ahead of the
    loop with the same plan (superbatch_geometry) the loop will follow."""

    def __init__(self, cfg, source, rounds):
        self.table, self.source = {}, source
        limit, _, size = engine.superbatch_geometry(cfg, 1, source.total)
        pos = 0
        for index in range(rounds):
            if pos >= limit:
                break
            want = min(size(index), limit - pos)
            source.position = pos
            self.table[pos] = (source.take(want), "factored")
            pos += want


@pytest.mark.parametrize("rounds,tol", [(0, 0.0), (2, 0.0), (99, 0.0), (3, None)])
def test_prefactored_superbatches_follow_the_plan(rounds, tol):
    """The split route prepares super-batches ahead of the sample loop; the loop must find them at
    exactly the positions and sizes it would have asked for (ramp-up included), use the rest of
    the stream normally, and produce the same result as without preparation."""
    Xtr, Xte, ytr, yte = problem()
    p = Xtr.shape[1]
    perms = so.perms_random(p, 150, 9)
    backend, coll = OracleBackend(), engine.Collective(None)
    prob = engine.reduce_problem(backend, coll, Xtr, Xte, ytr, yte, 0.05, p)
    import ls_spa_b200.engine as E
    old = E.target_samples
    E.target_samples = lambda p: 32
    try:
        base_cfg = dict(p=p, batch_size=4, max_samples=None, seed=1, antithetical=True, estimate_errors=True,
                        return_history=True)
        ref = engine.run_samples(backend, coll, prob, RangeSource(perms),
                                 engine.JobConfig(tolerance=0.0, **base_cfg))
        if tol is None:                               # an early stop inside the prepared range
            tol = float(np.sort(ref[0]["error_history"])[::-1][6]) * 1.000001
        cfg = engine.JobConfig(tolerance=tol, **base_cfg)
        want = engine.run_samples(backend, coll, prob, RangeSource(perms), cfg)
        src = RangeSource(perms)
        pre = FakePre(cfg, src, rounds)
        got = engine.run_samples(backend, coll, prob, src, cfg, pre=pre)
        assert got[2] == want[2] and got[0]["count"] == want[0]["count"]
        assert np.array_equal(got[0]["mean"], want[0]["mean"])
        assert np.array_equal(got[0]["error_history"], want[0]["error_history"])
        assert np.array_equal(got[1], want[1])
    finally:
        E.target_samples = old
