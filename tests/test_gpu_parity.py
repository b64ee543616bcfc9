"""GPU parity tests: the CUDA path (through the C ABI) against the numpy oracle and the
reference's golden outputs.  Tolerances: permutations bit-exact; lifts / attribution /
theta / r_squared max|d| <= 1e-9 * max|ref| (BASELINE.json north_star, fp64); Monte-Carlo
error estimates statistical only (hard parts)."""

import itertools

import numpy as np
import pytest

from conftest import load_golden, scaled_err

pytestmark = pytest.mark.gpu

TOL = 1e-9
SYN = ["syn_p10", "syn_p33", "syn_p100", "syn_p100_reg", "syn_p160"]


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


def regen(g):
    """The stored float32 inputs, widened to float64 (exact)."""
    return tuple(g[k].astype(np.float64) for k in ("X_train", "X_test", "y_train", "y_test"))


def device_problem(T, g):
    from ls_spa_b200 import ops
    dev = T.device("cuda")
    f = lambda a: T.from_numpy(np.ascontiguousarray(a)).to(dev)
    return ops.ReducedProblem(f(g["R_tr"]), f(g["c_tr"]), f(g["R_te"]), f(g["c_te"]), float(g["y_norm_sq"]))


# ------------------------------------------------------------------ permutation sources
def test_perm_sources_bit_exact(T):
    from ls_spa_b200 import samplers
    g = load_golden("streams")
    dev = T.device("cuda")
    for p, key in ((9, "random_p9_seed42"), (37, "random_p37_seed42"), (100, "random_p100_seed42"),
                   (1000, "random_p1000_seed42")):
        src = samplers.RandomSource(p, 42, None, dev)
        n = len(g[key])
        first = src.take(n // 2).cpu().numpy()          # two chained calls: state hand-over
        second = src.take(n - n // 2).cpu().numpy()
        src.check()
        assert np.array_equal(np.vstack([first, second]), g[key]), key
    src = samplers.RandomSource(100, 7, None, dev)
    assert np.array_equal(src.take(128).cpu().numpy(), g["random_p100_seed7"])
    for p, seed, key in ((10, 42, "argsort_p10_seed42"), (100, 42, "argsort_p100_seed42"),
                         (100, 7, "argsort_p100_seed7"), (1000, 42, "argsort_p1000_seed42")):
        src = samplers.ArgsortSource(p, seed, None, dev)
        assert np.array_equal(src.take(len(g[key])).cpu().numpy(), g[key]), key
    for p, seed, key in ((10, 42, "permutohedron_p10_seed42"), (11, 42, "permutohedron_p11_seed42"),
                         (100, 42, "permutohedron_p100_seed42"), (100, 7, "permutohedron_p100_seed7"),
                         (1000, 42, "permutohedron_p1000_seed42")):
        src = samplers.PermutohedronSource(p, seed, None, dev)
        got = src.take(len(g[key])).cpu().numpy()
        assert np.array_equal(got, g[key]), key
    src = samplers.ExactSource(5, dev)
    assert np.array_equal(src.take(1000).cpu().numpy(), g["exact_p5"])
    src = samplers.ExactSource(10, dev)
    assert np.array_equal(src.take(512).cpu().numpy(), g["exact_p10_first"])
    src.position = 3_000_000
    assert np.array_equal(src.take(512).cpu().numpy(), g["exact_p10_at_3000000"])


def test_random_source_long_stream(T):
    """4096 permutations of p=100 in uneven chunks == numpy's stream (bit-exact)."""
    from ls_spa_b200 import samplers
    src = samplers.RandomSource(100, 123, None, T.device("cuda"))
    got = np.vstack([src.take(n).cpu().numpy() for n in (1, 1000, 37, 3058)])
    src.check()
    rng = np.random.default_rng(123)
    want = np.array([rng.permutation(100) for _ in range(4096)])
    assert np.array_equal(got, want)


# ------------------------------------------------------------------ per-permutation core
def test_toy_lifts(T):
    from ls_spa_b200 import ops
    g = load_golden("toy")
    prob = device_problem(T, g)
    perms = T.from_numpy(g["perms"].astype(np.int32)).cuda()
    got = ops.lifts(prob, perms, False).cpu().numpy()
    assert scaled_err(got, g["lifts"]) < TOL
    # lifts of (2,0,1)
    np.testing.assert_allclose(got[4], [1.248311838792743, 0.10654150291296915, -0.43105313286637437],
                               rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", SYN)
def test_lifts_match_reference(T, name):
    from ls_spa_b200 import ops
    g = load_golden(name)
    prob = device_problem(T, g)
    methods = ["random", "argsort", "permutohedron"] + (["exact"] if f"perms_exact" in g.files else [])
    # both routes of the per-permutation core: Householder always, Cholesky where it is offered
    routes = [False] + ([True] if prob.gram is not None else [])
    if name.startswith("syn_p100"):
        assert prob.use_chol and prob.cond_estimate < 1e3, prob.cond_estimate   # the golden data are benign
    for method, route in itertools.product(methods, routes):
        prob.use_chol = route
        perms = g[f"perms_{method}"].astype(np.int32)
        got = ops.lifts(prob, T.from_numpy(perms).cuda(), False).cpu().numpy()
        want = g[f"lifts_{method}"]
        assert scaled_err(got, want) < TOL, (name, method, scaled_err(got, want))
        # invariant: every lift vector sums to the full-model R^2
        np.testing.assert_allclose(got.sum(axis=1), float(g[f"{method}_anti0_r_squared"]), atol=1e-10)
        # antithetic rows = mean of the lifts of perm and perm[::-1]
        rev = ops.lifts(prob, T.from_numpy(np.ascontiguousarray(perms[:, ::-1])).cuda(), False).cpu().numpy()
        anti = ops.lifts(prob, T.from_numpy(perms).cuda(), True).cpu().numpy()
        assert scaled_err(anti, 0.5 * (got + rev)) < 1e-13


def test_lifts_bitwise_reproducible(T):
    """compute-sanitizer is closed on this pool, so races are hunted the cheap way: the same launch
    repeated, and the same permutations in differently sized launches (different CTA <-> sample
    mapping, different co-residency), must give bit-identical rows."""
    from ls_spa_b200 import ops, samplers
    dev = T.device("cuda")
    for name in ("syn_p100", "syn_p33", "syn_p160"):
        g = load_golden(name)
        p = int(g["p"])
        prob = device_problem(T, g)
        perms = samplers.PermutohedronSource(p, 3, None, dev).take(1500)
        for route in [False] + ([True] if prob.gram is not None else []):
            prob.use_chol = route
            base = ops.lifts(prob, perms, True).cpu().numpy()
            for _ in range(3):
                assert np.array_equal(ops.lifts(prob, perms, True).cpu().numpy(), base), (name, route)
            parts = [ops.lifts(prob, perms[a:b].contiguous(), True).cpu().numpy()
                     for a, b in ((0, 7), (7, 300), (300, 1500))]
            assert np.array_equal(np.vstack(parts), base), (name, route)
            np.testing.assert_allclose(base.sum(axis=1), float(g["argsort_anti0_r_squared"]), atol=1e-10)


def test_lift_routes_agree_every_width(T):
    """Every instantiation of the Cholesky lift kernel (p = 17..152: tile geometry is a template
    parameter) against the Householder route (DMMA kernel for 49 <= p <= 128, scalar kernel elsewhere), with
    launch sizes below, at and above the grid."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from quick_bench import synth_problem
    from ls_spa_b200 import ops, samplers
    dev = T.device("cuda")
    for p in range(17, 153):
        prob = synth_problem(p, dev, seed=p)
        assert prob.gram is not None
        for count, anti in ((1, False), (3, True), (301, True)):
            perms = samplers.ArgsortSource(p, p + count, None, dev).take(count)
            prob.use_chol = True
            a = ops.lifts(prob, perms, anti)
            prob.use_chol = False
            b = ops.lifts(prob, perms, anti)
            assert not bool(T.isnan(a).any())
            err = float((a - b).abs().max() / b.abs().max())
            assert err < 1e-11, (p, count, anti, err)


def test_lift_route_selection(T):
    """The Cholesky route (error ~ eps cond^2) is only taken for well-conditioned train factors;
    both routes agree there for every tile count, and ill-conditioned or singular factors keep
    to Householder."""
    from ls_spa_b200 import ops, samplers
    dev = T.device("cuda")
    rng = np.random.default_rng(5)
    f = lambda a: T.from_numpy(np.ascontiguousarray(a)).to(dev)
    for p in (49, 56, 64, 71, 100, 104, 120, 128):
        R1 = np.linalg.qr(rng.standard_normal((3 * p, p)) / np.sqrt(3 * p), mode="r")
        R2 = np.linalg.qr(rng.standard_normal((3 * p, p)), mode="r")
        c1, c2 = rng.standard_normal(p), rng.standard_normal(p)
        prob = ops.ReducedProblem(f(R1), f(c1), f(R2), f(c2), float(c2 @ c2) * 1.5)
        # the estimate is a rigorous bound on the condition number of the column-equilibrated factor
        assert prob.use_chol and prob.cond_estimate >= np.linalg.cond(R1 / np.linalg.norm(R1, axis=0)) * (1 - 1e-12)
        perms = samplers.ArgsortSource(p, 11, None, dev).take(600)
        for anti in (False, True):
            prob.use_chol = True
            a = ops.lifts(prob, perms, anti).cpu().numpy()
            prob.use_chol = False
            b = ops.lifts(prob, perms, anti).cpu().numpy()
            assert scaled_err(a, b) < 1e-11, (p, anti, scaled_err(a, b))
    # badly scaled features (units): the Cholesky route equilibrates the columns, so it is still taken,
    # and the lifts equal those of the unscaled problem (they are invariant under feature scaling)
    p = 100
    R1 = np.linalg.qr(rng.standard_normal((3 * p, p)) / np.sqrt(3 * p), mode="r")
    R2 = np.linalg.qr(rng.standard_normal((3 * p, p)), mode="r")
    c1, c2 = rng.standard_normal(p), rng.standard_normal(p)
    scale = 10.0 ** rng.uniform(-4, 4, p)
    plain = ops.ReducedProblem(f(R1), f(c1), f(R2), f(c2), float(c2 @ c2) * 1.5)
    scaled = ops.ReducedProblem(f(R1 * scale), f(c1), f(R2 * scale), f(c2), float(c2 @ c2) * 1.5)
    assert np.linalg.cond(R1 * scale) > 1e6 and scaled.use_chol and scaled.cond_estimate < 1e3
    perms = samplers.ArgsortSource(p, 3, None, dev).take(500)
    plain.use_chol = False
    want = ops.lifts(plain, perms, True).cpu().numpy()
    got = ops.lifts(scaled, perms, True).cpu().numpy()
    assert scaled_err(got, want) < 1e-10, scaled_err(got, want)

    R2 = np.linalg.qr(rng.standard_normal((3 * p, p)), mode="r")
    c1, c2 = rng.standard_normal(p), rng.standard_normal(p)
    U, _, Vt = np.linalg.svd(rng.standard_normal((p, p)))
    for smin in (1e-6, 0.0):
        ill = np.linalg.qr((U * np.geomspace(1.0, max(smin, 1e-300), p)) @ Vt, mode="r")
        if smin == 0.0:
            ill[-1, -1] = 0.0
        prob = ops.ReducedProblem(f(ill), f(c1), f(R2), f(c2), float(c2 @ c2) * 1.5)
        assert not prob.use_chol and prob.cond_estimate > 1e5, (smin, prob.cond_estimate)


def test_square_shapley_export(T, L):
    g = load_golden("syn_p33")
    perm = g["perms_random"][3].astype(np.int64)
    got = L.square_shapley(g["R_tr"], g["R_te"], g["c_tr"], g["c_te"], float(g["y_norm_sq"]), perm)
    assert scaled_err(got, g["lifts_random"][3]) < TOL


# ------------------------------------------------------------------ reduction
@pytest.mark.parametrize("name", SYN + ["exact_p7"])
def test_reduce_matches_reference(T, L, name):
    g = load_golden(name)
    Xtr, Xte, ytr, yte = regen(g)
    R_tr, R_te, c_tr, c_te = L.reduce_data(Xtr, Xte, ytr, yte, float(g["reg"]))
    assert R_tr.shape == g["R_tr"].shape and R_te.shape == g["R_te"].shape
    assert np.allclose(np.tril(R_tr, -1), 0.0)
    for ours, ref in ((R_tr.T @ R_tr, g["R_tr"].T @ g["R_tr"]), (R_tr.T @ c_tr, g["R_tr"].T @ g["c_tr"]),
                      (R_te.T @ R_te, g["R_te"].T @ g["R_te"]), (R_te.T @ c_te, g["R_te"].T @ g["c_te"])):
        assert scaled_err(ours, ref) < 1e-11


def test_reduce_ragged_rows(T):
    """Row counts that are not multiples of the block size, one row, fewer rows than CTAs."""
    from ls_spa_b200 import ops
    rng = np.random.default_rng(0)
    for n, p in ((1, 3), (31, 5), (33, 5), (1000, 17), (4097, 64), (257, 130)):
        X, y = rng.standard_normal((n, p)), rng.standard_normal(n)
        slot = ops.tsqr_merge(ops.tsqr_rows(T.from_numpy(X).cuda(), T.from_numpy(y).cuda(), 2.0), p)
        R, c, ysq = (t.cpu().numpy() for t in ops.split_factor(slot, p))
        assert scaled_err(R.T @ R, X.T @ X / 4.0) < 1e-12, (n, p)
        assert scaled_err(R.T @ c, X.T @ y / 4.0) < 1e-12, (n, p)
        assert abs(float(ysq) - y @ y / 4.0) <= 1e-12 * max(y @ y, 1.0)


# ------------------------------------------------------------------ whole jobs
def test_toy_hello_world(T, L):
    g = load_golden("toy")
    res = L.ls_spa(g["X_train"], g["X_test"], g["y_train"], g["y_test"])
    assert isinstance(res, L.ShapleyResults)
    assert scaled_err(res.attribution, g["default_attribution"]) < TOL
    assert scaled_err(res.theta, g["default_theta"]) < TOL
    assert abs(res.r_squared - float(g["default_r_squared"])) < TOL
    assert res.overall_error == 0.0 and res.error_history.size == 0
    assert np.all(res.attribution_errors == 0.0) and res.attribution_history is None
    assert repr(res) == str(g["default_repr"])
    res = L.ls_spa(g["X_train"], g["X_test"], g["y_train"], g["y_test"], reg=0.1)
    assert scaled_err(res.attribution, g["reg01_attribution"]) < TOL
    assert scaled_err(res.theta, g["reg01_theta"]) < TOL
    res = L.ls_spa(g["X_train"], g["X_test"], g["y_train"], g["y_test"], return_attribution_history=True)
    assert scaled_err(res.attribution_history, g["hist_attribution_history"]) < TOL
    # README keywords: p <= 10 -> exact
    res = L.ls_spa(g["X_train"], g["X_test"], g["y_train"], g["y_test"], method="exact", return_history=True)
    assert scaled_err(res.attribution, g["default_attribution"]) < TOL
    assert res.attribution_history.shape == (6, 3)


def test_exact_default_p7(T, L):
    g = load_golden("exact_p7")
    Xtr, Xte, ytr, yte = regen(g)
    res = L.ls_spa(Xtr, Xte, ytr, yte, reg=float(g["reg"]))
    assert scaled_err(res.attribution, g["default_attribution"]) < TOL
    assert scaled_err(res.theta, g["default_theta"]) < TOL
    assert abs(res.r_squared - float(g["default_r_squared"])) < TOL
    assert abs(res.attribution.sum() - res.r_squared) < 1e-10


@pytest.mark.parametrize("name", SYN)
@pytest.mark.parametrize("anti", [0, 1])
def test_job_with_explicit_perms(T, L, name, anti):
    g = load_golden(name)
    Xtr, Xte, ytr, yte = regen(g)
    for method in ("random", "argsort", "permutohedron"):
        perms = g[f"perms_{method}"].astype(np.int64)
        k = len(perms)
        res = L.ls_spa(Xtr, Xte, ytr, yte, reg=float(g["reg"]), perms=list(perms), tolerance=0.0,
                       batch_size=max(k // 4, 2), antithetical=bool(anti), return_attribution_history=True)
        pre = f"{method}_anti{anti}_"
        assert scaled_err(res.attribution, g[pre + "attribution"]) < TOL, (name, method)
        assert scaled_err(res.theta, g[pre + "theta"]) < TOL
        assert abs(res.r_squared - float(g[pre + "r_squared"])) < TOL
        assert scaled_err(res.attribution_history, g[pre + "attribution_history"]) < TOL
        if int(g["p"]) >= 9:
            ref_hist = g[pre + "error_history"]
            assert res.error_history.shape == ref_hist.shape
            # Monte-Carlo estimate from 1024 draws: statistical agreement only
            np.testing.assert_allclose(res.error_history, ref_hist, rtol=0.2)
            np.testing.assert_allclose(res.overall_error, float(g[pre + "overall_error"]), rtol=0.2)
            ref_fe = g[pre + "attribution_errors"]
            big = ref_fe > 0.05 * ref_fe.max()
            np.testing.assert_allclose(res.attribution_errors[big], ref_fe[big], rtol=0.3)


def test_device_generators_inside_job(T, L):
    """method= uses the device generators; the same permutations handed to the oracle
    through perms= must give the same attribution."""
    from oracle import lsspa_oracle as lo
    from oracle import samplers_oracle as so
    g = load_golden("syn_p33")
    Xtr, Xte, ytr, yte = regen(g)
    p, k = 33, 48
    streams = {"random": so.perms_random(p, k, 11), "argsort": so.perms_argsort(p, k, 11)[0],
               "permutohedron": so.perms_permutohedron(p, k, 11)[0]}
    for method, perms in streams.items():
        for anti in (False, True):
            res = L.ls_spa(Xtr, Xte, ytr, yte, reg=1e-2, method=method, batch_size=16, num_batches=3,
                           tolerance=0.0, seed=11, antithetical=anti)
            want = lo.ls_spa_reference_loop(Xtr, Xte, ytr, yte, reg=1e-2, perms=list(perms), tolerance=0.0,
                                            batch_size=16, antithetical=anti)
            assert scaled_err(res.attribution, want.attribution) < TOL, (method, anti)
            assert res.error_history.shape == want.error_history.shape == (3,)


def test_estimator_state_matches_merge_formulas(T):
    """count / mean / biased covariance after several uneven batches == the reference's
    sequential merge_sample_mean / merge_sample_cov (ls_spa/ls_spa.py:212-216)."""
    from ls_spa_b200 import ops
    from oracle import lsspa_oracle as lo
    rng = np.random.default_rng(5)
    p, n = 23, 300
    rows = rng.standard_normal((n, p)) * rng.uniform(0.1, 3.0, p) + rng.standard_normal(p)
    dev = T.device("cuda")
    est = ops.Estimator(p, 99, True, dev)
    d = T.from_numpy(rows).to(dev)
    cuts = [(0, 100, 0), (100, 37, 100), (137, 163, 137)]
    overall, feat = est.absorb(est.partials(d, cuts), [0, 1, 2], [100, 37, 163], own=(0, 3), emit=True)
    out = est.read(want_cov=True)
    out["attribution_errors"] = feat[-1].cpu().numpy()
    out["overall_error"] = float(overall[-1].item())
    mean, cov = np.zeros(p), np.zeros((p, p))
    for i, r in enumerate(rows, 1):
        cov = lo.merge_sample_cov(mean, r, cov, np.zeros((p, p)), i - 1, 1)
        mean = lo.merge_sample_mean(mean, r, i - 1, 1)
    assert out["count"] == n and overall.shape == (3,)
    assert scaled_err(out["mean"], mean) < 1e-12
    assert scaled_err(out["cov"], cov) < 1e-11
    # the error draws have covariance unbiased_cov / n: compare with the analytic quantiles
    sd = np.sqrt(np.diag(cov) * n / (n - 1) / n)
    np.testing.assert_allclose(out["attribution_errors"], 1.959964 * sd, rtol=0.12)
    want_overall = lo.error_estimates(np.random.default_rng(0), cov * n / (n - 1) / n)[1]
    np.testing.assert_allclose(out["overall_error"], want_overall, rtol=0.1)


@pytest.mark.parametrize("p", [5, 23, 100, 130])
def test_fused_error_fold_matches_two_kernel_route(T, p):
    """lsspa_estimator_absorb_errors (draw fold + norms in one kernel, per-feature errors of the last batch
    only) against lsspa_estimator_absorb + lsspa_estimator_quantiles (every squared draw written out):
    same state, same overall errors for every batch, same per-feature errors on the last one --
    uneven batches, an empty one, a one-sample start (NaN estimate), two folds in a row."""
    from ls_spa_b200 import ops
    rng = np.random.default_rng(p)
    sizes = [1, 40, 0, 17, 64, 3]
    n = sum(sizes)
    rows = rng.standard_normal((2 * n, p)) * rng.uniform(0.1, 3.0, p) + rng.standard_normal(p)
    dev = T.device("cuda")
    d = T.from_numpy(rows).to(dev)
    outs = []
    for every in (True, False):
        est = ops.Estimator(p, 7, True, dev)
        res = []
        for fold in range(2):
            cuts, at = [], fold * n
            for sz in sizes:
                cuts.append((at, sz, at))
                at += sz
            part = est.partials(d, cuts)
            own = (0, len(sizes)) if fold == 0 else (2, 5)
            overall, feat = est.absorb(part, list(range(len(sizes))), sizes, own=own, emit=True, every_feature=every)
            res.append((overall.cpu().numpy(), feat[-1].cpu().numpy()))
        outs.append((res, est.state.cpu().numpy().copy(), est.count))
    (ra, sa, ca), (rb, sb, cb) = outs
    assert ca == cb == 2 * n
    np.testing.assert_allclose(sb, sa, rtol=1e-11, atol=1e-12)
    for (oa, fa), (ob, fb) in zip(ra, rb):
        assert oa.shape == ob.shape
        np.testing.assert_allclose(ob, oa, rtol=1e-10, equal_nan=True)
        np.testing.assert_allclose(fb, fa, rtol=1e-10, equal_nan=True)
    assert np.isnan(ra[0][0][0]) and np.isnan(rb[0][0][0])      # one sample: 0/0, never below a tolerance


def test_early_stop_matches_history_semantics(T, L):
    """Stops at the first batch whose estimated error is below the tolerance; later batches
    are not folded in (reference `break`, ls_spa/ls_spa.py:229)."""
    g = load_golden("syn_p33")
    Xtr, Xte, ytr, yte = regen(g)
    full = L.ls_spa(Xtr, Xte, ytr, yte, method="argsort", batch_size=32, num_batches=8, tolerance=0.0,
                    return_history=True)
    assert full.error_history.shape == (8,) and full.attribution_history.shape == (256, 33)
    tol = float(np.sort(full.error_history)[::-1][3]) * 1.0000001     # 4th largest -> stops early
    part = L.ls_spa(Xtr, Xte, ytr, yte, method="argsort", batch_size=32, num_batches=8, tolerance=tol,
                    return_history=True)
    stop_at = int(np.argmax(full.error_history < tol)) + 1
    assert part.error_history.shape == (stop_at,)
    np.testing.assert_allclose(part.error_history, full.error_history[:stop_at], rtol=1e-9)
    assert part.attribution_history.shape == (32 * stop_at, 33)
    assert scaled_err(part.attribution, full.attribution_history[32 * stop_at - 1]) < 1e-12
    assert part.overall_error < tol


# ------------------------------------------------------------------ the reference's own tests
class TestReferenceSuite:
    """test/test_ls_spa.py of the reference, restated against this package."""

    @pytest.fixture(scope="class")
    def data(self):
        rng = np.random.default_rng(128)
        n = 100
        A = rng.standard_normal((n, n))
        X, _ = np.linalg.qr(A)
        d = {"diag": np.sqrt(np.diag(np.arange(1, n + 1)))}
        d["Xtr_e"] = X @ d["diag"]
        d["Xte_e"] = d["Xtr_e"].copy()
        d["ytr_e"] = X[:, 0]
        d["yte_e"] = d["ytr_e"].copy()
        theta = rng.standard_normal(n)
        Xh = rng.multivariate_normal(np.zeros(n), A @ A.T, n)
        d["Xtr_h"] = Xh - np.mean(Xh, axis=0, keepdims=True)
        Xt = rng.multivariate_normal(np.zeros(n), A @ A.T, n)
        d["Xte_h"] = Xt - np.mean(Xh, axis=0, keepdims=True)
        yh = d["Xtr_h"] @ theta + rng.standard_normal(n)
        d["ytr_h"] = yh - np.mean(yh)
        yt = d["Xte_h"] @ theta + rng.standard_normal(n)
        d["yte_h"] = yt - np.mean(yt)
        return d

    def test_merge_functions(self, T, L):
        rng = np.random.default_rng(128)
        n = 100
        A = rng.standard_normal((n, 3 * n))
        X = rng.multivariate_normal(np.zeros(n), A @ A.T, 5 * n)
        b1, b2 = X[:2 * n], X[2 * n:]
        np.testing.assert_almost_equal(L.merge_sample_mean(b1.mean(0), b2.mean(0), 2 * n, 3 * n), X.mean(0))
        c = L.merge_sample_cov(b1.mean(0), b2.mean(0), np.cov(b1, rowvar=False, bias=True),
                               np.cov(b2, rowvar=False, bias=True), 2 * n, 3 * n)
        np.testing.assert_almost_equal(c, np.cov(X, rowvar=False, bias=True))

    def test_return_type(self, T, L, data):
        r = L.ls_spa(data["Xtr_e"], data["Xte_e"], data["ytr_e"], data["yte_e"])
        assert isinstance(r, L.ShapleyResults)

    def test_linear_regression(self, T, L, data):
        for s in ("e", "h"):
            theta = np.linalg.lstsq(data["Xtr_" + s], data["ytr_" + s], rcond=None)[0]
            r = L.ls_spa(data["Xtr_" + s], data["Xte_" + s], data["ytr_" + s], data["yte_" + s],
                         max_samples=4, batch_size=2)
            np.testing.assert_almost_equal(theta, r.theta)

    def test_rsquared(self, T, L, data):
        theta = np.linalg.lstsq(data["Xtr_h"], data["ytr_h"], rcond=None)[0]
        rss = np.sum((data["yte_h"] - data["Xte_h"] @ theta) ** 2)
        r2 = 1 - rss / np.sum(data["yte_h"] ** 2)
        r = L.ls_spa(data["Xtr_h"], data["Xte_h"], data["ytr_h"], data["yte_h"], max_samples=4, batch_size=2)
        np.testing.assert_almost_equal(r2, r.r_squared)
        # the reference's extra estimate at i == max_samples - 1: 2, 3, 4 -> three entries
        assert r.error_history.shape == (3,)

    def test_regularization(self, T, L, data):
        N, p = data["Xtr_h"].shape
        Xr = np.vstack((data["Xtr_h"] / np.sqrt(N), np.sqrt(0.1) * np.eye(p)))
        yr = np.concatenate((data["ytr_h"] / np.sqrt(N), np.zeros(p)))
        theta = np.linalg.lstsq(Xr, yr, rcond=None)[0]
        r = L.ls_spa(data["Xtr_h"], data["Xte_h"], data["ytr_h"], data["yte_h"], reg=0.1,
                     max_samples=4, batch_size=2)
        np.testing.assert_almost_equal(theta, r.theta)

    def test_random_seed_consistency(self, T, L, data):
        a = L.ls_spa(data["Xtr_h"], data["Xte_h"], data["ytr_h"], data["yte_h"], seed=42, max_samples=4, batch_size=2)
        b = L.ls_spa(data["Xtr_h"], data["Xte_h"], data["ytr_h"], data["yte_h"], seed=42, max_samples=4, batch_size=2)
        np.testing.assert_almost_equal(a.attribution, b.attribution)

    def test_correctness_easy(self, T, L, data):
        p = data["Xtr_e"].shape[1]
        proposal = np.zeros(p)
        tss = np.sum(data["yte_e"] ** 2)
        prev = 0.0
        for i in range(p):
            th = np.linalg.lstsq(data["Xtr_e"][:, :i + 1], data["ytr_e"], rcond=None)[0]
            r2 = 1 - np.sum((data["yte_e"] - data["Xte_e"][:, :i + 1] @ th) ** 2) / tss
            proposal[i] = r2 - prev
            prev = r2
        r = L.ls_spa(data["Xtr_e"], data["Xte_e"], data["ytr_e"], data["yte_e"],
                     max_samples=256 * 256, batch_size=256)
        np.testing.assert_almost_equal(proposal, r.attribution)
        assert r.error_history.size == 1          # variance ~ 0 -> stops after the first batch


def test_pandas_and_float32_inputs(T, L):
    import pandas as pd
    g = load_golden("toy")
    res = L.ls_spa(pd.DataFrame(g["X_train"]), pd.DataFrame(g["X_test"]), pd.Series(g["y_train"]),
                   pd.Series(g["y_test"]))
    assert scaled_err(res.attribution, g["default_attribution"]) < TOL
    res = L.ls_spa(g["X_train"].astype(np.float32), g["X_test"].astype(np.float32),
                   g["y_train"].astype(np.float32), g["y_test"].astype(np.float32))
    assert res.attribution.dtype == np.float64
    assert scaled_err(res.attribution, g["default_attribution"]) < 1e-5
    # device-resident inputs skip the host copy
    f = lambda a: T.from_numpy(a).cuda()
    res = L.ls_spa(f(g["X_train"]), f(g["X_test"]), f(g["y_train"]), f(g["y_test"]))
    assert scaled_err(res.attribution, g["default_attribution"]) < TOL


def test_size_incompatible(T, L):
    x = np.zeros((5, 3))
    with pytest.raises(L.SizeIncompatible):
        L.ls_spa(x, np.zeros((5, 4)), np.zeros(5), np.zeros(5))
    with pytest.raises(L.SizeIncompatible):
        L.ls_spa(np.zeros((2, 3)), np.zeros((5, 3)), np.zeros(2), np.zeros(5))


def test_full_size_properties(T, L):
    """p = 100 at scale (2^14 permutations): size-independent properties only.
    sum(attribution) == r_squared (telescoping), antithetic on/off agree statistically,
    and the mean of all lifts equals the attribution returned."""
    from oracle import samplers_oracle as so
    rng = np.random.default_rng(42)
    Xtr, Xte, ytr, yte, _, _ = so.gen_data(rng, 100, 20000, 20000)
    a = L.ls_spa(Xtr, Xte, ytr, yte, method="permutohedron", batch_size=128, num_batches=64,
                 tolerance=0.0, antithetical=True)
    b = L.ls_spa(Xtr, Xte, ytr, yte, method="argsort", batch_size=128, num_batches=128,
                 tolerance=0.0, antithetical=False)
    assert abs(a.attribution.sum() - a.r_squared) < 1e-10
    assert abs(b.attribution.sum() - b.r_squared) < 1e-10
    assert abs(a.r_squared - b.r_squared) < 1e-12
    assert np.max(np.abs(a.attribution - b.attribution)) < 5 * max(a.overall_error, b.overall_error)
    assert a.error_history.shape == (64,) and b.error_history.shape == (128,)
    assert a.error_history[-1] < a.error_history[0]


def test_generator_continuation(T, L):
    """seed=<numpy Generator>: the device PCG64 stream continues exactly where the host generator
    stands (the reference's experiment scripts draw data and permutations from one generator), and
    the host generator ends up where the reference's would."""
    from oracle import samplers_oracle as so
    rng = np.random.default_rng(5)
    p = 24
    Xtr, Xte, ytr, yte, _, _ = so.gen_data(rng, p, 900, 700)
    rng.integers(0, 10, 3)                      # leaves a buffered half word behind
    twin = np.random.default_rng(0)
    twin.bit_generator.state = rng.bit_generator.state
    got = L.ls_spa(Xtr, Xte, ytr, yte, max_samples=96, batch_size=16, tolerance=0.0, seed=rng, antithetical=False)
    perms = [twin.permutation(p) for _ in range(96)]
    want = L.ls_spa(Xtr, Xte, ytr, yte, perms=perms, batch_size=16, tolerance=0.0, antithetical=False)
    # same permutations, same lifts; the batch cuts differ by the reference's max_samples - 1 quirk
    assert scaled_err(got.attribution, want.attribution) < 1e-14
    a, b = rng.bit_generator.state, twin.bit_generator.state
    assert a["state"] == b["state"] and a["has_uint32"] == b["has_uint32"]
    assert not a["has_uint32"] or a["uinteger"] == b["uinteger"]    # stale when no half word is buffered
    assert np.array_equal(rng.permutation(p), twin.permutation(p))
    assert np.array_equal(rng.integers(0, 1000, 5), twin.integers(0, 1000, 5))


def test_ground_truth_experiment_script(T):
    """experiments/ground_truth_b200.py on a small shape: the true error of every
    sampler falls with the number of samples and the antithetic quasi-Monte-Carlo samplers beat
    plain Monte Carlo at equal cost."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "experiments"))
    import ground_truth_b200 as gtm
    out = gtm.run(p=20, n=4000, m=3000, gt_log2=14, samples_log2=10, quiet=True)
    assert out["ground_truth_error_estimate"] < 2e-3
    finals = {k: c["final_true_error"] for k, c in out["curves"].items()}
    for k, c in out["curves"].items():
        assert all(np.isfinite(c["true_error"])) and c["true_error"][-1] < c["true_error"][0], k
        assert abs(c["r_squared"] - out["r_squared"]) < 1e-12
    assert finals["apermutohedron"] < finals["random"] and finals["aargsort"] < finals["random"], finals
    assert max(finals.values()) < 0.05, finals


def test_naive_comparator_against_device(T, L):
    """the device path against the naive method on the unreduced data (small N)."""
    from oracle import naive_oracle as no
    from oracle import samplers_oracle as so
    rng = np.random.default_rng(3)
    p = 12
    Xtr, Xte, ytr, yte, _, _ = so.gen_data(rng, p, 400, 300)
    perms = [rng.permutation(p) for _ in range(24)]
    for reg in (0.0, 1e-2):
        want = no.naive_attribution(Xtr, Xte, ytr, yte, perms, reg=reg)
        got = L.ls_spa(Xtr, Xte, ytr, yte, reg=reg, perms=perms, tolerance=0.0, antithetical=False)
        assert scaled_err(got.attribution, want) < 1e-9, (reg, scaled_err(got.attribution, want))


def test_split_route_host_inputs(T, L, monkeypatch):
    """Host-resident inputs on the Cholesky route: the factorisations run ahead (train side only,
    overlapping the copy of the test rows) and the sample loop only eliminates.  Same results as the
    fused kernel, for every device sampler, with and without early stop."""
    from ls_spa_b200 import ops
    from oracle import samplers_oracle as so
    rng = np.random.default_rng(21)
    p = 64
    Xtr, Xte, ytr, yte, _, _ = so.gen_data(rng, p, 5000, 4000)
    calls = {"n": 0}
    real = ops.lifts_eliminate

    def counting(*a, **k):
        calls["n"] += 1
        return real(*a, **k)
    monkeypatch.setattr(ops, "lifts_eliminate", counting)
    from ls_spa_b200 import engine
    monkeypatch.setattr(engine.Prefactor, "FACTOR_RATE", 1e15)   # tiny test rows: lift the overlap-window cap
    for method, anti, tol in (("permutohedron", True, 0.0), ("argsort", False, 0.0), ("random", True, 0.0),
                              ("permutohedron", True, 5e-3)):
        kw = dict(reg=1e-3, method=method, batch_size=32, num_batches=40, tolerance=tol, seed=9, antithetical=anti,
                  return_history=True)
        calls["n"] = 0
        monkeypatch.setenv("LSSPA_SPLIT_ROUTE", "1")
        a = L.ls_spa(Xtr, Xte, ytr, yte, **kw)
        assert calls["n"] > 0, "the split route was not taken"
        monkeypatch.setenv("LSSPA_SPLIT_ROUTE", "0")
        calls["n"] = 0
        b = L.ls_spa(Xtr, Xte, ytr, yte, **kw)
        assert calls["n"] == 0
        assert scaled_err(a.attribution, b.attribution) < 1e-13, (method, anti, tol)
        assert a.error_history.shape == b.error_history.shape
        np.testing.assert_allclose(a.error_history, b.error_history, rtol=1e-9)
        assert scaled_err(a.attribution_history, b.attribution_history) < 1e-13
        # device-resident inputs never take it
        monkeypatch.setenv("LSSPA_SPLIT_ROUTE", "1")
        dev = [T.from_numpy(np.ascontiguousarray(v)).cuda() for v in (Xtr, Xte, ytr, yte)]
        calls["n"] = 0
        c = L.ls_spa(*dev, **kw)
        assert calls["n"] == 0 and scaled_err(c.attribution, b.attribution) < 1e-12


def test_torch_custom_ops_match_direct_calls(T):
    """torch.ops.ls_spa_b200.* (torch.library wrappers of the C ABI) against the golden lifts and the
    engine's direct calls."""
    from ls_spa_b200 import ops, samplers
    g = load_golden("syn_p100")
    prob = device_problem(T, g)
    perms = T.from_numpy(g["perms_random"].astype(np.int32)).cuda()
    a = T.ops.ls_spa_b200.lifts_chol(prob.gram, prob.R_te_scaled_cm, prob.c_te, prob.y_norm_sq, perms, False)
    b = T.ops.ls_spa_b200.lifts(prob.R_tr_cm, prob.c_tr, prob.R_te_cm, prob.c_te, prob.y_norm_sq, perms, False)
    assert scaled_err(a.cpu().numpy(), g["lifts_random"]) < TOL and scaled_err(b.cpu().numpy(), g["lifts_random"]) < TOL
    tr = T.ops.ls_spa_b200.theta_r2(prob.R_tr_cm, prob.c_tr, prob.R_te_cm, prob.c_te, prob.y_norm_sq).cpu().numpy()
    assert scaled_err(tr[:100], g["random_anti0_theta"]) < TOL and abs(tr[100] - float(g["random_anti0_r_squared"])) < TOL
    src = samplers.PermutohedronSource(100, 42, None, T.device("cuda"))
    got = T.ops.ls_spa_b200.perms_permutohedron(src.sv, src.shift, src.bits, 100, 0, 64).cpu().numpy()
    assert np.array_equal(got, load_golden("streams")["permutohedron_p100_seed42"][:64])
    Xtr, _, ytr, _ = regen(g)
    X, y = T.from_numpy(Xtr).cuda(), T.from_numpy(ytr).cuda()
    G = T.ops.ls_spa_b200.gram_reduce(X, y, 2.0).cpu().numpy().reshape(104, 104)
    Z = np.column_stack([Xtr, ytr]) / 2.0
    assert scaled_err(np.triu(G[:101, :101]), np.triu(Z.T @ Z)) < 1e-13
