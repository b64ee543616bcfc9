"""GPU parity at the BASELINE.json configurations (C2, C3, the C4-shaped
reduction, C5 reduced), the Cholesky lift kernel against the reference at every tile count, an
ill-conditioned problem on the Householder kernels, long sampler streams, and the robustness
cases of ADVICE.md (float32 device tensors, invalid permutations, one-sample estimates, tiny
batch sizes).  Everything goes through the C ABI; the oracle is only the checker.

Tolerances: permutations bit-exact; lifts / attribution / theta / r_squared
max|d| <= 1e-9 * max|ref| (fp64, BASELINE.json north_star); Gram invariants of the reduction 1e-11.
"""

import os

import numpy as np
import pytest

from conftest import load_golden, scaled_err

pytestmark = pytest.mark.gpu

TOL = 1e-9


@pytest.fixture(scope="module")
def T():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("gpu tests need a CUDA device")
    return torch


@pytest.fixture(scope="module")
def L():
    import ls_spa_b200
    return ls_spa_b200


def _problem(T, R_tr, c_tr, R_te, c_te, ynsq):
    from ls_spa_b200 import ops
    dev = T.device("cuda")
    f = lambda a: T.from_numpy(np.ascontiguousarray(a)).to(dev)
    return ops.ReducedProblem(f(R_tr), f(c_tr), f(R_te), f(c_te), float(ynsq))


# ------------------------------------------------------------------ C2: p = 10, N = M = 1e5, exact
def test_c2_exact_full_size(T, L):
    from oracle import lsspa_oracle as lo
    from oracle import samplers_oracle as so
    rng = np.random.default_rng(42)
    Xtr, Xte, ytr, yte, _, _ = so.gen_data(rng, 10, 100_000, 100_000, conditioning=10.0)
    fac = lo.reduce_data(Xtr, Xte, ytr, yte, 0.0)
    # reduction invariants at N = 1e5 against the oracle's LAPACK QR
    R_tr, R_te, c_tr, c_te = L.reduce_data(Xtr, Xte, ytr, yte, 0.0)
    for ours, ref in ((R_tr.T @ R_tr, fac[0].T @ fac[0]), (R_tr.T @ c_tr, fac[0].T @ fac[2]),
                      (R_te.T @ R_te, fac[1].T @ fac[1]), (R_te.T @ c_te, fac[1].T @ fac[3])):
        assert scaled_err(ours, ref) < 1e-11
    # a 4096-permutation slice of the lexicographic enumeration against the oracle
    sub = so.perms_exact(10, 4096, first=1_234_567)
    want = lo.mean_of_lifts(*fac, float(yte @ yte), sub)
    got = L.ls_spa(Xtr, Xte, ytr, yte, perms=sub, antithetical=False, tolerance=0.0)
    assert scaled_err(got.attribution, want) < TOL
    # all 10! permutations: the exact Shapley values; efficiency axiom: they sum to the full-model R^2
    full = L.ls_spa(Xtr, Xte, ytr, yte, method="exact")
    theta = np.linalg.lstsq(fac[0], fac[2], rcond=None)[0]
    r2 = (fac[3] @ fac[3] - np.sum((fac[3] - fac[1] @ theta) ** 2)) / float(yte @ yte)
    assert abs(full.r_squared - r2) < 1e-11 and scaled_err(full.theta, theta) < TOL
    assert abs(full.attribution.sum() - full.r_squared) < 1e-11
    assert full.overall_error == 0.0 and full.error_history.size == 0
    # symmetry axiom as a size-independent check of the enumeration: averaging over all permutations
    # equals averaging over all permutations composed with a fixed transposition of positions
    swapped = sub[:, [1, 0] + list(range(2, 10))]
    again = L.ls_spa(Xtr, Xte, ytr, yte, perms=np.vstack([sub, swapped]), antithetical=False, tolerance=0.0)
    both = lo.mean_of_lifts(*fac, float(yte @ yte), np.vstack([sub, swapped]))
    assert scaled_err(again.attribution, both) < TOL


# ------------------------------------------------------------------ C3: p = 100, N = M = 1e5, argsort
def test_c3_argsort_full_size(T, L):
    from oracle import lsspa_oracle as lo
    from oracle import samplers_oracle as so
    Xtr, Xte, ytr, yte, _, _ = so.gen_data(np.random.default_rng(42), 100, 100_000, 100_000)
    fac = lo.reduce_data(Xtr, Xte, ytr, yte, 0.0)
    R_tr, R_te, c_tr, c_te = L.reduce_data(Xtr, Xte, ytr, yte, 0.0)
    for ours, ref in ((R_tr.T @ R_tr, fac[0].T @ fac[0]), (R_tr.T @ c_tr, fac[0].T @ fac[2]),
                      (R_te.T @ R_te, fac[1].T @ fac[1]), (R_te.T @ c_te, fac[1].T @ fac[3])):
        assert scaled_err(ours, ref) < 1e-11
    perms = so.perms_argsort(100, 512, 42)[0]
    want = lo.mean_of_lifts(*fac, float(yte @ yte), perms)
    got = L.ls_spa(Xtr, Xte, ytr, yte, method="argsort", batch_size=128, num_batches=4, tolerance=0.0,
                   antithetical=False)
    assert scaled_err(got.attribution, want) < TOL
    assert got.error_history.shape == (4,)
    # the whole configuration: 2^7 x 2^7 samples
    full = L.ls_spa(Xtr, Xte, ytr, yte, method="argsort", batch_size=128, num_batches=128, tolerance=0.0,
                    antithetical=False)
    assert full.error_history.shape == (128,) and abs(full.attribution.sum() - full.r_squared) < 1e-10
    assert np.max(np.abs(full.attribution - got.attribution)) < 5 * got.overall_error


# ------------------------------------------------------------------ C4-shaped reduction: N = 1e6, p = 100
def test_c4_reduction_full_size(T, L):
    """[X | y] with 10^6 rows on the device against a host Gram matrix accumulated block-wise in
    extended precision (10^4-row float64 BLAS blocks summed in long double)."""
    dev = T.device("cuda")
    p, n = 100, 1_000_000
    g = T.Generator(device=dev).manual_seed(7)
    mix = T.randn(p, p, generator=g, device=dev, dtype=T.float64) * 0.15 + T.eye(p, device=dev, dtype=T.float64)
    X = T.randn(n, p, generator=g, device=dev, dtype=T.float64) @ mix
    y = X @ T.randn(p, generator=g, device=dev, dtype=T.float64) + T.randn(n, generator=g, device=dev, dtype=T.float64)
    reg = 1e-2
    R_tr, R_te, c_tr, c_te = L.reduce_data(X, X[: n // 2], y, y[: n // 2], reg)
    Xh, yh = X.cpu().numpy(), y.cpu().numpy()
    del X, y
    Z = np.column_stack([Xh, yh])

    def gram(rows):
        acc = np.zeros((p + 1, p + 1), dtype=np.longdouble)
        for lo_ in range(0, rows, 10_000):
            blk = Z[lo_: min(lo_ + 10_000, rows)]
            acc += (blk.T @ blk).astype(np.longdouble)
        return acc

    G = gram(n)
    G_tr = np.asarray(G / np.longdouble(n), dtype=np.float64)
    G_tr[:p, :p] += reg * np.eye(p)
    assert scaled_err(R_tr.T @ R_tr, G_tr[:p, :p]) < 1e-11
    assert scaled_err(R_tr.T @ c_tr, G_tr[:p, p]) < 1e-11
    G_te = np.asarray(gram(n // 2), dtype=np.float64)        # the test side is not scaled (:315)
    assert scaled_err(R_te.T @ R_te, G_te[:p, :p]) < 1e-11
    assert scaled_err(R_te.T @ c_te, G_te[:p, p]) < 1e-11
    assert np.allclose(np.tril(R_tr, -1), 0.0) and np.allclose(np.tril(R_te, -1), 0.0)


# ------------------------------------------------------------------ C5 reduced: p = 1000
def test_c5_reduced_p1000(T, L):
    """p = 1000 (N = 6000, M = 5000 so that the host oracle finishes in seconds): 8 random
    permutations against the reference's square_shapley, theta against lstsq."""
    from oracle import lsspa_oracle as lo
    from oracle import samplers_oracle as so
    Xtr, Xte, ytr, yte, _, _ = so.gen_data(np.random.default_rng(1), 1000, 6000, 5000)
    got = L.ls_spa(Xtr, Xte, ytr, yte, method="random", batch_size=4, num_batches=2, tolerance=0.0,
                   antithetical=False)
    fac = lo.reduce_data(Xtr, Xte, ytr, yte, 0.0)
    want = lo.mean_of_lifts(*fac, float(yte @ yte), so.perms_random(1000, 8, 42))
    assert scaled_err(got.attribution, want) < TOL
    theta = np.linalg.lstsq(fac[0], fac[2], rcond=None)[0]
    assert scaled_err(got.theta, theta) < 1e-8          # cond(R_tr) ~ 1e2 at N = 6 p
    assert abs(got.attribution.sum() - got.r_squared) < 1e-10
    assert got.error_history.shape == (2,)


@pytest.mark.parametrize("p", [200, 256, 333])
def test_large_p_lifts_against_oracle(T, p):
    """Widths above the single-SM tile kernels (p > 152) against the oracle's square_shapley,
    both conditionings, antithetic and plain, launch sizes that are not multiples of anything."""
    from ls_spa_b200 import ops
    from oracle import lsspa_oracle as lo
    from oracle import samplers_oracle as so
    Xtr, Xte, ytr, yte, _, _ = so.gen_data(np.random.default_rng(p), p, 5 * p, 4 * p)
    R_tr, R_te, c_tr, c_te = lo.reduce_data(Xtr, Xte, ytr, yte, 1e-3)
    ynsq = float(yte @ yte)
    prob = _problem(T, R_tr, c_tr, R_te, c_te, ynsq)
    perms = so.perms_random(p, 11, 3)
    want = np.array([lo.square_shapley(R_tr, R_te, c_tr, c_te, ynsq, pm) for pm in perms])
    rev = np.array([lo.square_shapley(R_tr, R_te, c_tr, c_te, ynsq, pm[::-1]) for pm in perms])
    dperms = T.from_numpy(perms.astype(np.int32)).cuda()
    assert prob.use_chol and prob.train.big and prob.cond_estimate < 1e3, prob.cond_estimate
    got = ops.lifts(prob, dperms, False).cpu().numpy()
    prob.check()
    assert ops.LIFT_ROUTE == "cholesky"
    assert scaled_err(got, want) < TOL, scaled_err(got, want)
    anti = ops.lifts(prob, dperms, True).cpu().numpy()
    assert scaled_err(anti, 0.5 * (want + rev)) < TOL
    # several passes over a small workspace (3 samples at a time) give the same rows
    prob._big_ws, prob.BIG_WS_BUDGET = None, 3 * 2 * int(prob._big_ws.numel() // (2 * 11))
    again = ops.lifts(prob, dperms, True).cpu().numpy()
    assert np.array_equal(again, anti)
    # the Householder kernels (scalar at these widths) agree
    prob.use_chol = False
    hh = ops.lifts(prob, dperms[:4].contiguous(), False).cpu().numpy()
    assert scaled_err(hh, want[:4]) < TOL


# ------------------------------------------------------------------ Cholesky lift kernel, every tile count
def test_cholesky_route_against_reference_every_tile_count(T):
    """lifts of the reference's square_shapley (tests/golden/widths.npz, made by oracle/make_golden.py)
    at 18 widths covering every row-tile count 3..19 of lifts_chol_kernel, and the oracle's
    square_shapley live on more permutations; both routes of the per-permutation core."""
    from ls_spa_b200 import ops
    from oracle import lsspa_oracle as lo
    from oracle import samplers_oracle as so
    g = load_golden("widths")
    assert len(so.WIDTHS) >= 12 and {(p + 7) // 8 for p in so.WIDTHS} == set(range(3, 20))
    for p in so.WIDTHS:
        R_tr, R_te, c_tr, c_te, ynsq, perms = so.width_problem(p)
        prob = _problem(T, R_tr, c_tr, R_te, c_te, ynsq)
        assert prob.gram is not None and prob.use_chol and prob.cond_estimate < 1e3, (p, prob.cond_estimate)
        dperms = T.from_numpy(perms.astype(np.int32)).cuda()
        for route in (True, False):
            prob.use_chol = route
            got = ops.lifts(prob, dperms, False).cpu().numpy()
            assert scaled_err(got, g[f"lifts_p{p}"]) < TOL, (p, route, scaled_err(got, g[f"lifts_p{p}"]))
        more = so.perms_random(p, 5, p)
        want = np.array([0.5 * (lo.square_shapley(R_tr, R_te, c_tr, c_te, ynsq, pm)
                                + lo.square_shapley(R_tr, R_te, c_tr, c_te, ynsq, pm[::-1])) for pm in more])
        prob.use_chol = True
        got = ops.lifts(prob, T.from_numpy(more.astype(np.int32)).cuda(), True).cpu().numpy()
        assert scaled_err(got, want) < TOL, (p, scaled_err(got, want))


# ------------------------------------------------------------------ ill-conditioned problem
def test_ill_conditioned_takes_householder_kernels(T, L, monkeypatch):
    """tests/golden/ill_p64.npz: singular values spread over 1e5 in random directions (equilibrated
    cond 9.5e4).  The condition guard must send it to the Householder lift kernel; the reduction is
    checked on both of its routes.  Both the reference (LAPACK Householder) and this path are
    backward stable, so they differ by O(eps * cond) = 1e-11 relative; the tolerance is eps*cond*100."""
    from ls_spa_b200 import ops
    g = load_golden("ill_p64")
    cond = float(g["cond_equilibrated"])
    tol = 2.2e-16 * cond * 100
    assert 5e4 < cond < 2e5 and tol < 5e-9
    prob = _problem(T, g["R_tr"], g["c_tr"], g["R_te"], g["c_te"], float(g["y_norm_sq"]))
    assert not prob.use_chol and prob.cond_estimate > 1e4
    perms = g["perms_random"].astype(np.int32)
    got = ops.lifts(prob, T.from_numpy(perms).cuda(), False).cpu().numpy()
    assert ops.LIFT_ROUTE == "householder"
    assert scaled_err(got, g["lifts_random"]) < tol, scaled_err(got, g["lifts_random"])
    Xtr, Xte, ytr, yte = (g[k].astype(np.float64) for k in ("X_train", "X_test", "y_train", "y_test"))
    for reduce_route in ("cholqr2", "householder"):
        monkeypatch.setenv("LSSPA_REDUCE", reduce_route)
        for anti in (0, 1):
            res = L.ls_spa(Xtr, Xte, ytr, yte, perms=list(g["perms_random"].astype(np.int64)), tolerance=0.0,
                           batch_size=4, antithetical=bool(anti), return_attribution_history=True)
            pre = f"random_anti{anti}_"
            assert ops.LIFT_ROUTE == "householder"
            assert scaled_err(res.attribution, g[pre + "attribution"]) < tol, (reduce_route, anti)
            assert abs(res.r_squared - float(g[pre + "r_squared"])) < tol
            assert scaled_err(res.attribution_history, g[pre + "attribution_history"]) < tol
            # theta = R^-1 c amplifies by cond once more
            assert scaled_err(res.theta, g[pre + "theta"]) < tol * 1e3


# ------------------------------------------------------------------ long sampler streams
def test_long_qmc_streams_bit_exact(T):
    """2^17 argsort and permutohedron permutations at p = 100 against scipy on the host.  The device
    sorts by rank counting with the index as tie-break = a stable sort; numpy's default argsort
    (introsort) may order exact ties differently, so a mismatching row is accepted only if the host
    row has tied keys and the device row equals the STABLE argsort of the same keys."""
    from ls_spa_b200 import samplers
    from oracle import samplers_oracle as so
    dev = T.device("cuda")
    n, p = 1 << 17, 100
    for name, make, src in (("argsort", so.perms_argsort, samplers.ArgsortSource(p, 42, None, dev)),
                            ("permutohedron", so.perms_permutohedron, samplers.PermutohedronSource(p, 42, None, dev))):
        want, keys = make(p, n, 42)
        got = np.vstack([src.take(c).cpu().numpy() for c in (1, 4095, n - 4096)])
        bad = np.nonzero((got != want).any(axis=1))[0]
        for r in bad:
            srt = np.sort(keys[r])
            assert (srt[1:] == srt[:-1]).any(), (name, int(r), "differs without a tie")
            assert np.array_equal(got[r], np.argsort(keys[r], kind="stable")), (name, int(r))
        print(f"{name}: {len(bad)} of {n} rows differ from numpy's unstable tie order")
        assert len(bad) <= 8, (name, len(bad))


# ------------------------------------------------------------------ ADVICE.md robustness cases
def test_float32_cuda_tensor_inputs(T, L):
    """Device tensors of torch's default dtype must be widened, not read as raw float64 storage."""
    from oracle import samplers_oracle as so
    Xtr, Xte, ytr, yte, _, _ = so.gen_data(np.random.default_rng(9), 24, 3000, 2000)
    f32 = [T.from_numpy(a.astype(np.float32)).cuda() for a in (Xtr, Xte, ytr, yte)]
    f64 = [t.double() for t in f32]
    kw = dict(method="argsort", batch_size=16, num_batches=4, tolerance=0.0, seed=3)
    a = L.ls_spa(*f32, **kw)
    b = L.ls_spa(*f64, **kw)
    assert a.attribution.dtype == np.float64
    assert scaled_err(a.attribution, b.attribution) < 1e-13 and abs(a.r_squared - b.r_squared) < 1e-13
    # non-contiguous columns (a transposed view) and a strided row slice
    Xt = f64[0].t().contiguous().t()
    assert Xt.stride(1) != 1
    c = L.ls_spa(Xt, f64[1], f64[2], f64[3], **kw)
    assert scaled_err(c.attribution, b.attribution) < 1e-13
    R = L.reduce_data(f32[0], f32[1], f32[2], f32[3], 0.0)
    R64 = L.reduce_data(f64[0], f64[1], f64[2], f64[3], 0.0)
    assert scaled_err(R[0].T @ R[0], R64[0].T @ R64[0]) < 1e-13


def test_invalid_permutations_are_rejected(T, L):
    from ls_spa_b200 import ops
    g = load_golden("syn_p33")
    Xtr, Xte, ytr, yte = (g[k].astype(np.float64) for k in ("X_train", "X_test", "y_train", "y_test"))
    good = g["perms_random"].astype(np.int64)
    ops.perms_validate(T.from_numpy(good.astype(np.int32)).cuda())
    for mutate in ("duplicate", "too_big", "negative"):
        bad = good.copy()
        if mutate == "duplicate":
            bad[5, 7] = bad[5, 8]
        elif mutate == "too_big":
            bad[2, 0] = 33
        else:
            bad[0, 3] = -1
        with pytest.raises(ValueError):
            L.ls_spa(Xtr, Xte, ytr, yte, perms=bad, tolerance=0.0)
        with pytest.raises(ValueError):
            L.ls_spa(Xtr, Xte, ytr, yte, perms=[row for row in bad], tolerance=0.0)
        with pytest.raises(ValueError):
            L.ls_spa(Xtr, Xte, ytr, yte, perms=T.from_numpy(bad.astype(np.int32)).cuda(), tolerance=0.0)
    with pytest.raises(ValueError):
        L.square_shapley(g["R_tr"], g["R_te"], g["c_tr"], g["c_te"], float(g["y_norm_sq"]), np.zeros(33, dtype=int))
    # a wide permutation matrix (p = 1000) goes through the same kernel
    wide = np.array([np.random.default_rng(1).permutation(1000) for _ in range(3)], dtype=np.int32)
    ops.perms_validate(T.from_numpy(wide).cuda())
    wide[1, 999] = wide[1, 0]
    with pytest.raises(ValueError):
        ops.perms_validate(T.from_numpy(wide).cuda())


def test_one_sample_estimate_never_stops(T, L):
    """An error estimate after a single sample is 0/0 in the reference (it returns nan and nan <
    tolerance is False): batch_size = 1 must not stop after one sample with a zero error."""
    g = load_golden("syn_p33")
    Xtr, Xte, ytr, yte = (g[k].astype(np.float64) for k in ("X_train", "X_test", "y_train", "y_test"))
    r = L.ls_spa(Xtr, Xte, ytr, yte, max_samples=6, batch_size=1, tolerance=1e-2, antithetical=False)
    assert r.error_history.size >= 2 and np.isnan(r.error_history[0]) and np.all(np.isfinite(r.error_history[1:]))
    r = L.ls_spa(Xtr, Xte, ytr, yte, max_samples=2, batch_size=4, tolerance=1e-2)   # quirk: estimate at i = 1
    assert np.isnan(r.error_history[0]) and r.error_history.size == 2


def test_tiny_batches_bounded_memory(T, L):
    """batch_size = 2 with the default sample budget (the reference's own tests use it): the
    super-batch is cut so that the per-batch partial blocks stay within the byte budget."""
    from ls_spa_b200 import engine
    g = load_golden("syn_p100")
    Xtr, Xte, ytr, yte = (g[k].astype(np.float64) for k in ("X_train", "X_test", "y_train", "y_test"))
    assert engine.batch_cap(100, True) * 2 < engine.target_samples(100)       # the cap binds here
    T.cuda.reset_peak_memory_stats()
    base = T.cuda.memory_allocated()
    r = L.ls_spa(Xtr, Xte, ytr, yte, max_samples=4096, batch_size=2, tolerance=0.0)
    assert r.error_history.shape == (2049,)         # 2048 batches + the max_samples - 1 quirk
    assert T.cuda.max_memory_allocated() - base < 5 << 30
    big = L.ls_spa(Xtr, Xte, ytr, yte, max_samples=4096, batch_size=512, tolerance=0.0)
    assert scaled_err(r.attribution, big.attribution) < 1e-12


def test_error_estimates_export(T, L):
    """api.error_estimates against numpy's draws from the same covariance (statistical: 1024 draws)."""
    from oracle import lsspa_oracle as lo
    rng = np.random.default_rng(4)
    p = 40
    A = rng.standard_normal((p, p - 3)) * rng.uniform(0.2, 2.0, p)[:, None]
    cov = A @ A.T / 50.0                                  # rank p - 3: singular, like the lift covariance
    feat, overall = L.error_estimates(np.random.default_rng(0), cov)
    want_feat, want_overall = lo.error_estimates(np.random.default_rng(1), cov)
    assert feat.shape == (p,) and np.isfinite(overall)
    np.testing.assert_allclose(feat, 1.959964 * np.sqrt(np.diag(cov)), rtol=0.12)
    np.testing.assert_allclose(feat, want_feat, rtol=0.2)
    np.testing.assert_allclose(overall, want_overall, rtol=0.08)


def test_gram_reduction_ragged_shapes(T, L):
    """Every branch of the Gram route on awkward shapes: one row, fewer rows than a chunk, row counts that
    are not multiples of 32, odd p (no TMA: the leading dimension is not 16-byte aligned), p + 1 = 120 (the
    widest single-CTA tile set), wide problems with few rows, a strided view, fewer rows than features
    (train side regularised; the singular test-side Gram matrix takes the Householder fallback)."""
    rng = np.random.default_rng(0)
    for n, p in ((1, 3), (31, 5), (33, 6), (1000, 17), (4097, 64), (257, 118), (300, 119), (500, 130), (700, 200),
                 (64, 100), (90, 100)):
        X, y = rng.standard_normal((n, p)), rng.standard_normal(n)
        Xt, yt = rng.standard_normal((n + 3, p)), rng.standard_normal(n + 3)
        reg = 0.5 if n < p else 0.0
        R_tr, R_te, c_tr, c_te = L.reduce_data(X, Xt, y, yt, reg)
        N = len(X)
        G = X.T @ X / N + reg * np.eye(p)
        assert scaled_err(R_tr.T @ R_tr, G) < 1e-11, (n, p)
        assert scaled_err(R_tr.T @ c_tr, X.T @ y / N) < 1e-11, (n, p)
        m = min(len(Xt), p)
        assert R_te.shape == (m, p)
        assert scaled_err(R_te.T @ R_te, Xt.T @ Xt) < 1e-11, (n, p)
        assert scaled_err(R_te.T @ c_te, Xt.T @ yt) < 1e-11, (n, p)
    # a strided device view (leading dimension 2 p) and a row slice starting at an odd row
    n, p = 3001, 40
    big = T.from_numpy(rng.standard_normal((n, 2 * p))).cuda()
    yv = T.from_numpy(rng.standard_normal(n)).cuda()
    Xv = big[1:, :p]
    R_tr, _, c_tr, _ = L.reduce_data(Xv, Xv, yv[1:], yv[1:], 0.0)
    Xh, yh = Xv.cpu().numpy(), yv[1:].cpu().numpy()
    assert scaled_err(R_tr.T @ R_tr, Xh.T @ Xh / (n - 1)) < 1e-12
    assert scaled_err(R_tr.T @ c_tr, Xh.T @ yh / (n - 1)) < 1e-12


@pytest.mark.parametrize("p", [111, 112, 120, 128, 129, 152, 153])
def test_whole_jobs_around_the_kernel_boundaries(T, L, p):
    """Widths where the reduction (single-CTA Cholesky tail <-> blocked wide factorisation at p + 1 = 112/113)
    and the lift kernels (packed four-warp <-> eight-warp at 128/129, shared-memory <-> tile workspace at
    152/153) change hands: whole jobs against the oracle's reference loop on the same permutations."""
    from oracle import lsspa_oracle as lo
    from oracle import samplers_oracle as so
    Xtr, Xte, ytr, yte, _, _ = so.gen_data(np.random.default_rng(p), p, 6 * p, 5 * p)
    perms = so.perms_random(p, 12, 1)
    got = L.ls_spa(Xtr, Xte, ytr, yte, reg=1e-3, perms=perms, tolerance=0.0, batch_size=4, antithetical=True,
                   return_attribution_history=True)
    want = lo.ls_spa_reference_loop(Xtr, Xte, ytr, yte, reg=1e-3, perms=list(perms), tolerance=0.0, batch_size=4,
                                    antithetical=True, return_attribution_history=True)
    assert scaled_err(got.attribution, want.attribution) < TOL, scaled_err(got.attribution, want.attribution)
    assert scaled_err(got.theta, want.theta) < 1e-8
    assert abs(got.r_squared - want.r_squared) < TOL
    assert scaled_err(got.attribution_history, want.attribution_history) < TOL
    assert got.error_history.shape == want.error_history.shape


def test_float32_host_inputs_travel_narrow(T, L, monkeypatch):
    """The optional fp32 mode is a TRANSFER format: float32 host data (numpy or pinned tensors) crosses the
    link as float32 and is widened on the device, so the fp64 arithmetic sees exactly the values a host-side
    astype(float64) would give (bit-identical results), and against the fp64 oracle on the ORIGINAL fp64 data
    the attribution agrees to the 1e-4 that BASELINE.json's north_star asks of an fp32 mode."""
    from oracle import lsspa_oracle as lo
    from oracle import samplers_oracle as so
    from ls_spa_b200 import engine
    Xtr, Xte, ytr, yte, _, _ = so.gen_data(np.random.default_rng(3), 40, 6000, 5000)
    kw = dict(reg=1e-3, method="permutohedron", batch_size=16, num_batches=8, tolerance=0.0, seed=4)
    narrow = [a.astype(np.float32) for a in (Xtr, Xte, ytr, yte)]
    wide = [a.astype(np.float64) for a in narrow]
    seen = []
    real = T.Tensor.copy_

    def spy(self, src, *a, **k):
        if self.is_cuda and not src.is_cuda:
            seen.append(src.dtype)
        return real(self, src, *a, **k)
    monkeypatch.setattr(T.Tensor, "copy_", spy)
    a = L.ls_spa(*narrow, **kw)
    assert seen and all(d == T.float32 for d in seen), seen          # nothing was widened on the host
    seen.clear()
    pinned = [T.from_numpy(x).pin_memory() for x in narrow]
    c = L.ls_spa(*pinned, **kw)
    assert seen and all(d == T.float32 for d in seen)
    monkeypatch.undo()
    b = L.ls_spa(*wide, **kw)
    assert a.attribution.dtype == np.float64
    assert np.array_equal(a.attribution, b.attribution) and np.array_equal(c.attribution, b.attribution)
    assert a.r_squared == b.r_squared and np.array_equal(a.theta, b.theta)
    perms = so.perms_permutohedron(40, 128, 4)[0]
    want = lo.ls_spa_reference_loop(Xtr, Xte, ytr, yte, reg=1e-3, perms=list(perms), tolerance=0.0, batch_size=16)
    assert scaled_err(a.attribution, want.attribution) < 1e-4
    assert abs(a.r_squared - want.r_squared) < 1e-4 and scaled_err(a.theta, want.theta) < 1e-4
