"""Permutation sources and the synthetic-data recipe, as the reference's own
drivers build them (TEST INFRASTRUCTURE ONLY).

Each generator here is a *pure stream*: it owns its RNG/QMC engine and is not
interleaved with the error estimator's draws (SURVEY.md section 0.4 explains
why the reference's ``perms=None`` stream is not well defined beyond the first
batch).  The results are what one hands to the reference through ``perms=``.

Third-party arithmetic (not under /root/reference): numpy 2.3.5
``Generator.permutation`` (PCG64 + masked-rejection Fisher-Yates) and scipy 1.18.1
``scipy.stats.qmc.Sobol`` / ``MultivariateNormalQMC``.
"""

from __future__ import annotations

import itertools

import numpy as np
from scipy.stats.qmc import MultivariateNormalQMC, Sobol


def perms_random(p, count, seed):
    """``default_rng(seed).permutation(p)`` repeated.   ls_spa/ls_spa.py:168,175"""
    rng = np.random.default_rng(seed)
    return np.array([rng.permutation(p) for _ in range(count)], dtype=np.int64).reshape(count, p)


def perms_exact(p, count=None, first=0):
    """``itertools.permutations(range(p))`` = lexicographic.   ls_spa/ls_spa.py:171"""
    it = itertools.permutations(range(p))
    it = itertools.islice(it, first, None if count is None else first + count)
    return np.array(list(it), dtype=np.int64).reshape(-1, p)


def perms_argsort(p, count, seed, one_at_a_time=False):
    """``np.argsort(Sobol(p, seed).random(n), axis=1)``.
    experiments/ground_truth_medium.py:70-71; notebooks/medium_experiment.py:390-391
    (the notebook draws one point per call; Sobol' is order-consistent so both agree)."""
    eng = Sobol(p, seed=seed)
    if one_at_a_time:
        pts = np.vstack([eng.random(1) for _ in range(count)])
    else:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pts = eng.random(count)
    return np.argsort(pts, axis=1).astype(np.int64), pts


def permutohedron_matrix(p):
    """The (p-1) x p projection used by ``permutohedron_samples``.
    experiments/ground_truth_medium.py:61-65"""
    lower = np.tril(np.ones((p - 1, p)))
    upper = np.diag(-np.arange(1, p), 1)[:-1]
    u = lower + upper
    return u / np.linalg.norm(u, axis=1, keepdims=True)


def perms_permutohedron(p, count, seed):
    """experiments/ground_truth_medium.py:56-67 with
    ``MultivariateNormalQMC(zeros(p-1), seed=seed, inv_transform=False)``
    (notebooks/medium_experiment.py:433-435)."""
    import warnings
    qmc = MultivariateNormalQMC(np.zeros(p - 1), seed=seed, inv_transform=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        z = qmc.random(count)
    z = z / np.linalg.norm(z, axis=1, keepdims=True)                       # :59
    proj = z @ permutohedron_matrix(p)                                     # :66
    return np.argsort(proj, axis=1).astype(np.int64), proj                 # :67


def gen_data(rng, p, n_train, n_test, stn_ratio=5.0, conditioning=20.0):
    """Synthetic regression problem of the medium experiment.
    experiments/ground_truth_medium.py:74-106 (same draw order from ``rng``)."""
    a = rng.standard_normal((p, int(p / conditioning)))                    # :77
    cov = a @ a.T + np.eye(p)                                              # :78
    v = np.sqrt(np.diag(cov))
    cov = cov / np.outer(v, v)                                             # :79-80
    X_train = rng.multivariate_normal(np.zeros(p), cov, (n_train,), method="svd")
    X_test = rng.multivariate_normal(np.zeros(p), cov, (n_test,), method="svd")
    k = max((p + 1) // 10, 1)
    theta_vals = np.zeros(p)
    theta_vals[:k] = 2.0                                                   # :89-90
    theta_true = rng.permutation(theta_vals)                               # :91
    std = np.sqrt(np.sum(np.diag(cov) * theta_true ** 2) / stn_ratio)      # :94
    y_train = X_train @ theta_true + std * rng.standard_normal(n_train)    # :95
    x_mean = np.mean(X_train, axis=0, keepdims=True)
    X_train = X_train - x_mean                                             # :97-98
    y_mean = np.mean(y_train)
    y_train = y_train - y_mean                                             # :99-100
    y_test = X_test @ theta_true + std * rng.standard_normal(n_test)       # :102
    X_test = X_test - x_mean                                               # :103
    y_test = y_test - y_mean                                               # :104
    return X_train, X_test, y_train, y_test, theta_true, cov


WIDTHS = (17, 24, 31, 40, 48, 56, 64, 65, 80, 88, 96, 104, 105, 120, 128, 129, 144, 152)


def width_problem(p):
    """Reduced problem of width p built WITHOUT LAPACK (seeded numpy draws only), so that the test
    regenerates it bit for bit on any host: well-conditioned upper-triangular factors."""
    rng = np.random.default_rng(1000 + p)
    R_tr = np.triu(np.eye(p) + 0.6 * rng.standard_normal((p, p)) / np.sqrt(p))
    R_te = np.triu(1.5 * np.eye(p) + 0.9 * rng.standard_normal((p, p)) / np.sqrt(p))
    c_tr, c_te = rng.standard_normal(p), rng.standard_normal(p)
    perms = np.array([rng.permutation(p) for _ in range(4)])
    return R_tr, R_te, c_tr, c_te, float(c_te @ c_te) * 1.3, perms
