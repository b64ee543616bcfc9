"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the reference's naive comparator.

Restates notebooks/medium_experiment.py:251-312 of the reference: for every permutation, p
full-data least-squares fits (one per prefix of the permutation) on the UNREDUCED data, the
out-of-sample R^2 of each fit, and the lift of a feature = the R^2 gained when it joins.  It shares
nothing with the reduction trick (no QR compression, no triangular solves), which makes it an
independent check of reduce_data + square_shapley at small N.

Parity: pinned against the unmodified reference package in tests/test_oracle_golden.py through
the identity naive lifts == ls_spa.square_shapley(reduce_data(...)) on the golden inputs (the
reference's naive method lives in a notebook cell and cannot be imported).
"""
import numpy as np


def naive_lifts(X_train, X_test, y_train, y_test, perm, reg=0.0):
    """Lift vector of one permutation (notebook :262-276).  reg > 0 adds the ridge rows
    sqrt(reg) I of reduce_data (ls_spa/ls_spa.py:309-311: train rows scaled by 1/sqrt(N))."""
    perm = np.asarray(perm)
    p = X_train.shape[1]
    n = X_train.shape[0]
    ysq = float(np.sum(y_test ** 2))
    lift = np.zeros(p)
    baseline = 0.0
    for j in range(1, p + 1):
        cols = perm[:j]
        A, b = X_train[:, cols], y_train
        if reg > 0.0:
            A = np.vstack([A / np.sqrt(n), np.sqrt(reg) * np.eye(j)])
            b = np.concatenate([b / np.sqrt(n), np.zeros(j)])
        theta = np.linalg.lstsq(A, b, rcond=None)[0]
        cost = float(np.sum((X_test[:, cols] @ theta - y_test) ** 2))
        r_sq = (ysq - cost) / ysq
        lift[perm[j - 1]] = r_sq - baseline
        baseline = r_sq
    return lift


def naive_attribution(X_train, X_test, y_train, y_test, perms, reg=0.0):
    """Mean of the naive lift vectors (notebook :277-283 without the bookkeeping)."""
    perms = np.asarray(perms)
    return np.mean([naive_lifts(X_train, X_test, y_train, y_test, q, reg) for q in perms], axis=0)
