"""NumPy restatement of the reference LS-SPA hot path (TEST INFRASTRUCTURE ONLY).

Every function names the reference lines it follows (paths are relative to
``/root/reference``).  The restatement issues the *same* LAPACK/BLAS calls as the
reference (``np.linalg.qr``, ``scipy.linalg.solve_triangular``, dense ``@``), so
that timing it is a faithful stand-in for timing the reference on the same host
(``bench.py`` ``cpu_baseline.kind == "port"``).

The third-party arithmetic the reference reaches but does not vendor is
numpy 2.3.5 (PCG64 ``Generator``, LAPACK through OpenBLAS 0.3.30) and
scipy 1.18.1 (``scipy.stats.qmc``); those are the pinned oracle versions.

Parity pinned by ``tests/golden/*.npz`` (made by ``oracle/make_golden.py`` from
the unmodified reference) -- see ``tests/test_oracle_golden.py``.
"""

from __future__ import annotations

import itertools
from dataclasses import dataclass

import numpy as np
import scipy.linalg

ERR_DRAWS = 1 << 10      # ls_spa/ls_spa.py:334  (size=2 ** 10)
ERR_QUANTILE = 0.95      # ls_spa/ls_spa.py:339-340


@dataclass
class OracleResults:
    """Same seven fields, same order, as ``ShapleyResults`` (ls_spa/ls_spa.py:34-42)."""

    attribution: np.ndarray
    theta: np.ndarray
    overall_error: float
    attribution_errors: np.ndarray
    r_squared: float
    error_history: np.ndarray
    attribution_history: np.ndarray | None


# --------------------------------------------------------------------------
# stage 1: tall-skinny reduction                     ls_spa/ls_spa.py:290-318
# --------------------------------------------------------------------------
def reduce_data(X_train, X_test, y_train, y_test, reg):
    """(N,p),(M,p),(N,),(M,) -> R_tr (p,p), R_te (min(M,p),p), c_tr, c_te.

    Train side is scaled by 1/sqrt(N) and padded with sqrt(reg)*I rows
    (:309-312); test side is used as is (:315).  Reduced-mode QR of each, then
    the targets are rotated by Q^T (:314-317).
    """
    n_train, p = X_train.shape
    root_n = np.sqrt(n_train)
    stacked_X = np.vstack((X_train / root_n, np.sqrt(reg) * np.eye(p)))   # :309-310
    stacked_y = np.concatenate((y_train / root_n, np.zeros(p)))            # :311-312
    q_tr, r_tr = np.linalg.qr(stacked_X)                                   # :314
    q_te, r_te = np.linalg.qr(X_test)                                      # :315
    return r_tr, r_te, q_tr.T @ stacked_y, q_te.T @ y_test                 # :316-318


# --------------------------------------------------------------------------
# stage 2: one permutation -> lift vector            ls_spa/ls_spa.py:256-287
# --------------------------------------------------------------------------
def square_shapley(R_tr, R_te, c_tr, c_te, y_norm_sq, perm):
    """Lift vector of one permutation, literal operation order of the reference."""
    perm = np.asarray(perm)
    p = R_tr.shape[0]
    q, r = np.linalg.qr(R_tr[:, perm])                                     # :275
    x_te = R_te[:, perm]                                                   # :276
    rhs = np.triu(q.T @ np.tile(c_tr, (p, 1)).T)                           # :278
    coef = scipy.linalg.solve_triangular(r, rhs)                           # :279
    coef = np.hstack((np.zeros((p, 1)), coef))                             # :280
    resid = x_te @ coef - np.tile(c_te, (p + 1, 1)).T                      # :282-283
    costs = np.sum(resid ** 2, axis=0)                                     # :283
    r_sq = (np.linalg.norm(c_te) ** 2 - costs) / y_norm_sq                 # :284
    return np.ediff1d(r_sq)[np.argsort(perm)]                              # :285


def square_shapley_lean(R_tr, R_te, c_tr, c_te, y_norm_sq, perm):
    """Algebraically identical, cheaper formulation.

    Used only to cross-check the structure of the CUDA kernel on the CPU:
    QR of [R_tr[:,perm] | c_tr] without forming Q, prefix coefficients from
    R^-1, residual recurrence.  Not the timed baseline.
    """
    perm = np.asarray(perm)
    p = R_tr.shape[0]
    rfac = np.linalg.qr(np.column_stack((R_tr[:, perm], c_tr)), mode="r")
    r, c = rfac[:p, :p], rfac[:p, p]
    w = scipy.linalg.solve_triangular(r, R_te[:, perm].T, trans="T").T     # W = X R^-1
    resid = c_te.astype(float).copy()
    costs = np.empty(p + 1)
    costs[0] = resid @ resid
    for k in range(p):
        resid = resid - c[k] * w[:, k]
        costs[k + 1] = resid @ resid
    r_sq = (c_te @ c_te - costs) / y_norm_sq
    out = np.empty(p)
    out[perm] = np.diff(r_sq)
    return out


# --------------------------------------------------------------------------
# stage 3: online statistics                          ls_spa/ls_spa.py:103-119
# --------------------------------------------------------------------------
def merge_sample_mean(old_mean, new_mean, old_N, new_N):
    tot = old_N + new_N                                                    # :105
    return (old_N / tot) * old_mean + (new_N / tot) * new_mean             # :106-108


def merge_sample_cov(old_mean, new_mean, old_cov, new_cov, old_N, new_N):
    tot = old_N + new_N                                                    # :114
    gap = old_mean - new_mean                                              # :115
    cross = (old_N / tot) * (new_N / tot) * np.outer(gap, gap)             # :118
    return (old_N / tot) * old_cov + (new_N / tot) * new_cov + cross       # :116-119


def error_estimates(rng, cov):
    """1024 draws of N(0, cov); 95 % quantiles.        ls_spa/ls_spa.py:321-341

    Cholesky first; on *any* failure the reference redraws with SVD (:333-336).
    """
    p = cov.shape[0]
    try:
        z = rng.multivariate_normal(np.zeros(p), cov, size=ERR_DRAWS, method="cholesky")
    except Exception:  # the reference uses a bare except (:335)
        z = rng.multivariate_normal(np.zeros(p), cov, size=ERR_DRAWS, method="svd")
    per_feature = np.quantile(np.abs(z), ERR_QUANTILE, axis=0)             # :337,339
    overall = np.quantile(np.linalg.norm(z, axis=1), ERR_QUANTILE)          # :338,340
    return per_feature, overall


# --------------------------------------------------------------------------
# driver                                              ls_spa/ls_spa.py:122-253
# --------------------------------------------------------------------------
def ls_spa_reference_loop(X_train, X_test, y_train, y_test, reg=0.0,
                          max_samples=2 ** 13, batch_size=2 ** 8, tolerance=1e-2,
                          seed=42, perms=None, antithetical=True,
                          return_attribution_history=False, lift_fn=square_shapley,
                          return_lifts=False):
    """The reference's estimator loop with its own keyword set (:122-133).

    ``lift_fn`` lets tests swap in ``square_shapley_lean``; ``return_lifts``
    additionally returns the per-sample lift vectors (pair-averaged when
    antithetical) for kernel-level parity checks.
    """
    X_train, X_test = np.array(X_train), np.array(X_test)                  # :158-159
    y_train, y_test = np.array(y_train), np.array(y_test)                  # :160-161
    p = X_train.shape[1]
    rng = np.random.default_rng(seed)                                      # :168
    if perms is None:
        if p < 9:                                                          # :170-173
            perms = itertools.permutations(range(p))
            batch_size, antithetical = 2 ** 8, False
        else:                                                              # :175
            perms = (rng.permutation(p) for _ in range(max_samples))
    else:
        max_samples = 2 ** 100                                             # :177

    y_norm_sq = np.linalg.norm(y_test) ** 2                                # :180
    R_tr, R_te, c_tr, c_te = reduce_data(X_train, X_test, y_train, y_test, reg)

    mean = np.zeros(p)
    cov = np.zeros((p, p))
    feat_err = np.full(p, 0.0)
    overall = 0.0
    err_hist = np.zeros(0)
    hist = [] if return_attribution_history else None
    kept = []
    count, pending = 0, False
    for i, perm in enumerate(perms, 1):                                    # :197
        count, pending = i, True
        perm = np.array(perm)
        lift = lift_fn(R_tr, R_te, c_tr, c_te, y_norm_sq, perm)            # :203-204
        if antithetical:                                                   # :205-208
            lift = (lift + lift_fn(R_tr, R_te, c_tr, c_te, y_norm_sq, perm[::-1])) / 2
        if return_lifts:
            kept.append(lift)
        cov = merge_sample_cov(mean, lift, cov, np.zeros((p, p)), i - 1, 1)  # :212-214
        mean = merge_sample_mean(mean, lift, i - 1, 1)                     # :215-216
        if hist is not None:
            hist.append(mean.copy())                                       # :217-219
        if (i % batch_size == 0 or i == max_samples - 1) and p >= 9:       # :222
            feat_err, overall = error_estimates(rng, cov * i / (i - 1) / i)  # :223-224
            err_hist = np.append(err_hist, overall)                        # :225
            pending = False
            if overall < tolerance:                                        # :229
                break
    if p >= 9 and pending:                                                 # :233-236
        feat_err, overall = error_estimates(rng, cov * count / (count - 1) / count)
        err_hist = np.append(err_hist, overall)

    theta = np.linalg.lstsq(R_tr, c_tr, rcond=None)[0]                     # :240
    r_squared = ((np.linalg.norm(c_te) ** 2
                  - np.linalg.norm(c_te - R_te @ theta) ** 2) / y_norm_sq)  # :241-243
    res = OracleResults(mean, theta, overall, feat_err, r_squared, err_hist,
                        None if hist is None else np.array(hist).reshape(-1, p))
    if return_lifts:
        return res, np.array(kept).reshape(-1, p)
    return res


def mean_of_lifts(R_tr, R_te, c_tr, c_te, y_norm_sq, perms, antithetical=False,
                  lift_fn=square_shapley):
    """Plain average of lift vectors over an explicit permutation list; the CPU
    baseline's inner loop (one unit = one ``square_shapley`` call)."""
    perms = np.asarray(perms)
    acc = np.zeros(R_tr.shape[0])
    for perm in perms:
        lift = lift_fn(R_tr, R_te, c_tr, c_te, y_norm_sq, perm)
        if antithetical:
            lift = (lift + lift_fn(R_tr, R_te, c_tr, c_te, y_norm_sq, perm[::-1])) / 2
        acc += lift
    return acc / max(len(perms), 1)
