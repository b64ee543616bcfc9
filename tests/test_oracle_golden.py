"""CPU: the numpy oracle against the reference's golden outputs (tests/golden,
made by oracle/make_golden.py from the unmodified reference) and against the
numbers quoted in SURVEY.md section 8c."""


import numpy as np
import pytest

from conftest import load_golden, scaled_err
from oracle import lsspa_oracle as lo
from oracle import samplers_oracle as so

SYN = ["syn_p10", "syn_p33", "syn_p100", "syn_p100_reg", "syn_p160"]


def regen(g):
    """The stored float32 inputs, widened to float64 (exact)."""
    return tuple(g[k].astype(np.float64) for k in ("X_train", "X_test", "y_train", "y_test"))


def test_survey_toy_numbers():
    g = load_golden("toy")
    # SURVEY.md 8c, measured on the reference
    np.testing.assert_allclose(g["default_attribution"],
                               [0.5967131862126389, 0.47096034709822954, -0.14387332447153078], rtol=0, atol=1e-15)
    np.testing.assert_allclose(g["default_theta"],
                               [2.070837486284558, 1.365338023490087, 0.07234202811499152], rtol=0, atol=1e-14)
    assert abs(float(g["default_r_squared"]) - 0.9238002088393378) < 1e-15
    res = lo.ls_spa_reference_loop(g["X_train"], g["X_test"], g["y_train"], g["y_test"])
    assert scaled_err(res.attribution, g["default_attribution"]) < 1e-13
    assert scaled_err(res.theta, g["default_theta"]) < 1e-13
    assert abs(res.r_squared - float(g["default_r_squared"])) < 1e-14
    assert res.overall_error == 0.0 and res.error_history.size == 0 and res.attribution_history is None
    res = lo.ls_spa_reference_loop(g["X_train"], g["X_test"], g["y_train"], g["y_test"], reg=0.1)
    assert scaled_err(res.attribution, g["reg01_attribution"]) < 1e-13


def test_toy_lifts_both_formulations():
    g = load_golden("toy")
    for fn in (lo.square_shapley, lo.square_shapley_lean):
        got = np.array([fn(g["R_tr"], g["R_te"], g["c_tr"], g["c_te"], float(g["y_norm_sq"]), pm)
                        for pm in g["perms"]])
        assert scaled_err(got, g["lifts"]) < 1e-13


@pytest.mark.parametrize("name", SYN)
def test_reduce_and_lifts(name):
    g = load_golden(name)
    Xtr, Xte, ytr, yte = regen(g)
    R_tr, R_te, c_tr, c_te = lo.reduce_data(Xtr, Xte, ytr, yte, float(g["reg"]))
    # factors are unique only up to row signs: compare invariants
    assert scaled_err(R_tr.T @ R_tr, g["R_tr"].T @ g["R_tr"]) < 1e-12
    assert scaled_err(R_tr.T @ c_tr, g["R_tr"].T @ g["c_tr"]) < 1e-12
    assert scaled_err(R_te.T @ R_te, g["R_te"].T @ g["R_te"]) < 1e-12
    for method in ("random", "argsort", "permutohedron"):
        perms = g[f"perms_{method}"].astype(np.int64)
        for fn, tol in ((lo.square_shapley, 1e-13), (lo.square_shapley_lean, 1e-10)):
            got = np.array([fn(g["R_tr"], g["R_te"], g["c_tr"], g["c_te"], float(g["y_norm_sq"]), pm)
                            for pm in perms[:8]])
            assert scaled_err(got, g[f"lifts_{method}"][:8]) < tol
            # telescoping invariant: lifts sum to the full-model R^2
            np.testing.assert_allclose(got.sum(axis=1), float(g[f"{method}_anti0_r_squared"]), atol=1e-12)


@pytest.mark.parametrize("name", ["syn_p10", "syn_p33", "syn_p100_reg"])
@pytest.mark.parametrize("anti", [0, 1])
def test_driver_loop(name, anti):
    g = load_golden(name)
    Xtr, Xte, ytr, yte = regen(g)
    perms = g["perms_argsort"].astype(np.int64)
    k = len(perms)
    res = lo.ls_spa_reference_loop(Xtr, Xte, ytr, yte, reg=float(g["reg"]), perms=list(perms),
                                   tolerance=0.0, batch_size=max(k // 4, 2), antithetical=bool(anti),
                                   return_attribution_history=True)
    pre = f"argsort_anti{anti}_"
    assert scaled_err(res.attribution, g[pre + "attribution"]) < 1e-12
    assert scaled_err(res.theta, g[pre + "theta"]) < 1e-12
    assert abs(res.r_squared - float(g[pre + "r_squared"])) < 1e-13
    assert scaled_err(res.attribution_history, g[pre + "attribution_history"]) < 1e-12
    assert res.error_history.shape == g[pre + "error_history"].shape
    # Monte-Carlo quantity: same generator, same draws -> should agree closely unless the
    # Cholesky/SVD coin flip lands differently; keep it statistical
    np.testing.assert_allclose(res.error_history, g[pre + "error_history"], rtol=0.25)


def test_exact_default_path_p7():
    g = load_golden("exact_p7")
    Xtr, Xte, ytr, yte = regen(g)
    res = lo.ls_spa_reference_loop(Xtr, Xte, ytr, yte, reg=float(g["reg"]))
    assert scaled_err(res.attribution, g["default_attribution"]) < 1e-12
    assert res.overall_error == 0.0


def test_streams_match_generators():
    g = load_golden("streams")
    # default_rng(42).permutation(100)[:12], next call, and p=9
    assert list(g["random_p100_seed42"][0][:12]) == [59, 21, 56, 18, 33, 42, 50, 27, 92, 71, 2, 25]
    assert list(g["random_p100_seed42"][1][:12]) == [34, 26, 42, 46, 38, 7, 79, 45, 40, 56, 20, 6]
    assert list(g["random_p9_seed42"][0]) == [3, 0, 7, 2, 4, 6, 1, 5, 8]
    assert np.array_equal(so.perms_random(37, 64, 42), g["random_p37_seed42"])
    assert np.array_equal(so.perms_argsort(100, 256, 7)[0], g["argsort_p100_seed7"])
    assert np.array_equal(so.perms_argsort(10, 64, 42, one_at_a_time=True)[0], g["argsort_p10_seed42"][:64])
    assert np.array_equal(so.perms_permutohedron(100, 256, 7)[0], g["permutohedron_p100_seed7"])
    assert np.array_equal(so.perms_exact(10, 512, first=3_000_000), g["exact_p10_at_3000000"])


def test_online_stats_merge():
    # reference test/test_ls_spa.py:20-44 restated on the oracle
    rng = np.random.default_rng(128)
    n = 40
    a = rng.standard_normal((n, 3 * n))
    x = rng.multivariate_normal(np.zeros(n), a @ a.T, 5 * n)
    b1, b2 = x[:2 * n], x[2 * n:]
    m = lo.merge_sample_mean(b1.mean(0), b2.mean(0), 2 * n, 3 * n)
    np.testing.assert_almost_equal(m, x.mean(0))
    c = lo.merge_sample_cov(b1.mean(0), b2.mean(0), np.cov(b1, rowvar=False, bias=True),
                            np.cov(b2, rowvar=False, bias=True), 2 * n, 3 * n)
    np.testing.assert_almost_equal(c, np.cov(x, rowvar=False, bias=True))


@pytest.mark.parametrize("name", ["syn_p10", "syn_p33", "syn_p100_reg"])
def test_naive_comparator_matches_reference_lifts(name):
    """the naive method (p full-data least-squares fits per permutation, reference
    notebooks/medium_experiment.py:251-312) reproduces the reference's own per-permutation lifts
    stored in the goldens -- an independent check of the reduction trick."""
    from oracle import naive_oracle as no
    g = load_golden(name)
    Xtr, Xte, ytr, yte = regen(g)
    reg = float(g["reg"])
    perms = g["perms_random"][:3]
    for k, perm in enumerate(perms):
        got = no.naive_lifts(Xtr, Xte, ytr, yte, perm, reg=reg)
        assert scaled_err(got, g["lifts_random"][k]) < 1e-9, (name, k)
