"""CPU: the device permutation algorithms (modelled in tests/device_models.py exactly as
ls_spa_b200/csrc/perms.cu implements them) against numpy / scipy / itertools streams."""

import itertools

import numpy as np
import pytest
from scipy.stats.qmc import MultivariateNormalQMC, Sobol

import device_models as dm
from conftest import load_golden


@pytest.mark.parametrize("p,count", [(9, 64), (37, 64), (100, 40), (2, 16), (3, 50), (66, 20)])
def test_pcg64_scan_matches_numpy(p, count):
    st = dm.state_from_numpy(42)
    got, iters = dm.pcg64_perms(p, count, st)
    rng = np.random.default_rng(42)
    want = np.array([rng.permutation(p) for _ in range(count)])
    assert np.array_equal(got, want)
    assert iters <= 33
    # chained call continues the same stream (state hand-over incl. the buffered half word)
    got2, _ = dm.pcg64_perms(p, 7, st)
    want2 = np.array([rng.permutation(p) for _ in range(7)])
    assert np.array_equal(got2, want2)
    after = rng.bit_generator.state
    assert st[4] == after["has_uint32"]
    assert ((st[0] << 64) | st[1]) == after["state"]["state"]
    if st[4]:
        assert st[5] == after["uinteger"]


@pytest.mark.parametrize("p,count", [(9, 300), (37, 64), (100, 40), (2, 16), (3, 50)])
def test_pcg64_fsm_scan_matches_numpy(p, count):
    """The parallel automaton version (pcg64_fsm / group / prefix / emit / finish kernels)."""
    st = dm.state_from_numpy(42)
    budget = 2 * (count * 3 * p // 2 + 700)   # several chunks of 256 draws and a partial one
    got, _ = dm.pcg64_perms(p, count, st, fsm=True, budget=budget)
    rng = np.random.default_rng(42)
    want = np.array([rng.permutation(p) for _ in range(count)])
    assert np.array_equal(got, want)
    got2, _ = dm.pcg64_perms(p, 7, st, fsm=True)
    want2 = np.array([rng.permutation(p) for _ in range(7)])
    assert np.array_equal(got2, want2)
    after = rng.bit_generator.state
    assert st[4] == after["has_uint32"]
    assert ((st[0] << 64) | st[1]) == after["state"]["state"]
    if st[4]:
        assert st[5] == after["uinteger"]


def test_pcg64_golden_p100_and_p1000():
    g = load_golden("streams")
    st = dm.state_from_numpy(42)
    got, _ = dm.pcg64_perms(100, 24, st)
    assert np.array_equal(got, g["random_p100_seed42"][:24])
    st = dm.state_from_numpy(42)
    got, _ = dm.pcg64_perms(1000, 3, st)
    assert np.array_equal(got, g["random_p1000_seed42"][:3])


def test_exact_unranking():
    for p in (1, 3, 5):
        want = list(itertools.permutations(range(p)))
        got = [tuple(dm.exact_perm(p, r)) for r in range(len(want))]
        assert got == want
    g = load_golden("streams")
    got = np.array([dm.exact_perm(10, 3_000_000 + r) for r in range(64)])
    assert np.array_equal(got, g["exact_p10_at_3000000"][:64])
    # p > 20: leading positions stay the identity
    row = dm.exact_perm(23, 5)
    assert row[:3] == [0, 1, 2] and sorted(row) == list(range(23))


def test_sobol_argsort_model():
    g = load_golden("streams")
    for p, seed, key in ((100, 42, "argsort_p100_seed42"), (10, 42, "argsort_p10_seed42"),
                         (1000, 42, "argsort_p1000_seed42")):
        eng = Sobol(p, seed=seed)
        sv, shift = eng._sv.copy(), eng._shift.copy()
        n = min(len(g[key]), 128)
        for k in list(range(8)) + [n - 1]:
            assert np.array_equal(dm.sobol_argsort_perm(sv, shift, eng.bits, k), g[key][k])
    # counting rank == stable argsort
    keys = [5, 3, 5, 1, 3]
    assert dm.rank_scatter(keys) == list(np.argsort(keys, kind="stable"))


def test_permutohedron_model():
    g = load_golden("streams")
    for p, key, n in ((100, "permutohedron_p100_seed42", 1024), (10, "permutohedron_p10_seed42", 256),
                      (11, "permutohedron_p11_seed42", 64), (1000, "permutohedron_p1000_seed42", 16)):
        eng = MultivariateNormalQMC(np.zeros(p - 1), seed=42, inv_transform=False).engine
        sv, shift = eng._sv.copy(), eng._shift.copy()
        assert sv.shape[0] == 2 * ((p - 1 + 1) // 2)
        bad = 0
        for k in range(n):
            perm, _ = dm.permutohedron_perm(p, sv, shift, eng.bits, k)
            bad += not np.array_equal(perm, g[key][k])
        assert bad == 0


@pytest.mark.parametrize("p", [9, 20, 29])
def test_lane_level_lift_model_matches_oracle(p):
    """tests/lifts_v2_model.py mirrors the DMMA lift kernel lane by lane (fragments, tile accesses,
    C->A reuse, the residual product with the triangular c matrix, the three-shuffle row sums);
    it must reproduce the oracle's lifts."""
    import lifts_v2_model as lm
    from oracle import lsspa_oracle as lo
    from oracle import samplers_oracle as so
    rng = np.random.default_rng(p)
    Xtr, Xte, ytr, yte, _, _ = so.gen_data(rng, p, 40 * p, 30 * p)
    R_tr, R_te, c_tr, c_te = lo.reduce_data(Xtr, Xte, ytr, yte, 1e-3)
    ynsq = float(yte @ yte)
    for _ in range(2):
        perm = rng.permutation(p)
        got = lm.lifts_one(R_tr, c_tr, R_te, c_te, ynsq, perm)
        want = lo.square_shapley(R_tr, R_te, c_tr, c_te, ynsq, perm)
        assert np.max(np.abs(got - want)) < 1e-12 * max(np.max(np.abs(want)), 1e-300) + 1e-15


def test_fast_sobol_tables_equal_scipy_engines():
    """samplers.sobol_tables (vectorised rebuild of scipy's LMS scrambling) against the engines scipy
    itself constructs: Sobol(d, seed) and the engine inside MultivariateNormalQMC (spawned stream)."""
    import warnings
    from scipy.stats.qmc import MultivariateNormalQMC, Sobol
    from ls_spa_b200.samplers import sobol_tables
    for d, seed in ((1, 0), (5, 3), (37, 123), (100, 42), (100, 7), (1000, 5)):
        e = Sobol(d, seed=seed)
        sv, shift = sobol_tables(d, seed)
        assert sv.dtype == np.uint32 and sv.shape == (d, 30)
        assert np.array_equal(sv, e._sv.astype(np.uint32)) and np.array_equal(shift, e._shift.astype(np.uint32)), (d, seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for p, seed in ((2, 1), (10, 42), (11, 5), (100, 42), (1000, 9)):
            q = MultivariateNormalQMC(np.zeros(p - 1), seed=seed, inv_transform=False)
            sv, shift = sobol_tables(q.engine.d, seed, spawn=True)
            assert q.engine.bits == 30 and q.engine.d == 2 * -(-(p - 1) // 2)
            assert np.array_equal(sv, q.engine._sv.astype(np.uint32)), (p, seed)
            assert np.array_equal(shift, q.engine._shift.astype(np.uint32)), (p, seed)
