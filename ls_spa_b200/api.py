"""Drop-in public surface of the reference package (``from ls_spa import *`` exposes
exactly these names, reference ls_spa/__init__.py:1):

    ls_spa, ShapleyResults, SizeIncompatible, validate_data, merge_sample_mean,
    merge_sample_cov, square_shapley, reduce_data, error_estimates

``ls_spa`` accepts both keyword sets that exist for it:
  * the one the reference code ships (ls_spa/ls_spa.py:122-133):
    ``reg, max_samples, batch_size, tolerance, seed, perms, antithetical,
    return_attribution_history``;
  * the one its README documents (README.md:96-106) and BASELINE.json names:
    ``reg, method, batch_size, num_batches, tolerance, seed, return_history``.
Mixing ``max_samples`` with ``num_batches`` (or the two history flags) is rejected.

Everything numeric runs on the GPU through ``libls_spa_b200.so``; without a CUDA
device the calls raise ``LsSpaCudaError`` (there is no CPU fallback).
"""

from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np
import torch

from . import engine, ops
from ._cabi import LsSpaCudaError
from .samplers import make_source

_SOURCE_POOL = None


def _seed_int(seed) -> int:
    """Seed of the device's own streams (error-estimate Gaussians).  A numpy Generator handed in as
    ``seed`` (the permutation stream continues from it) contributes a value derived from its state
    without consuming from it."""
    if isinstance(seed, np.random.Generator):
        return int(seed.bit_generator.state["state"]["state"] & ((1 << 62) - 1))
    return int(seed)


def _source_pool():
    global _SOURCE_POOL
    if _SOURCE_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _SOURCE_POOL = ThreadPoolExecutor(max_workers=1, thread_name_prefix="lsspa-source")
    return _SOURCE_POOL

METHODS = ("random", "permutohedron", "argsort", "exact")


def _fmt_head(values) -> str:
    flat = np.asarray(values).flatten()
    shown = ", ".join("{:.2f}".format(v) for v in flat[:5])
    return "(" + shown + (", ...)" if len(flat) > 5 else ")")


@dataclass
class ShapleyResults:
    """Same fields, same order as the reference's record (ls_spa/ls_spa.py:34-42)."""

    attribution: np.ndarray
    theta: np.ndarray
    overall_error: float
    attribution_errors: np.ndarray
    r_squared: float
    error_history: np.ndarray | None
    attribution_history: np.ndarray | None

    def __repr__(self):
        # same dashboard text as the reference prints (ls_spa/ls_spa.py:44-70)
        lines = [
            "",
            "        p = {}".format(len(np.asarray(self.attribution).flatten())),
            "        Out-of-sample R^2 with all features: {:.2f}".format(self.r_squared),
            "",
            "        Shapley attribution: {}".format(_fmt_head(self.attribution)),
            "        Estimated error in Shapley attribution: {:.2E}".format(self.overall_error),
            "",
            "        Fitted coeficients with all features: {}".format(_fmt_head(self.theta)),
            "        ",
        ]
        return "\n".join(lines)


class SizeIncompatible(Exception):
    """Shapes of the data do not fit together (reference ls_spa/ls_spa.py:73-78)."""

    def __init__(self, message):
        self.message = message
        super().__init__(self.message)


def validate_data(X_train, X_test, y_train, y_test):
    """The reference's four shape checks, in its order (ls_spa/ls_spa.py:81-100)."""
    checks = (
        (X_train.shape[1] != X_test.shape[1],
         "X_train and X_test should have the same number of columns (features)."),
        (X_train.shape[0] != y_train.shape[0],
         "X_train should have the same number of rows as y_train has entries (observations)."),
        (X_test.shape[0] != y_test.shape[0],
         "X_test should have the same number of rows as y_test has entries (observations)."),
        (X_train.shape[1] > X_train.shape[0],
         "The function works only if the number of features is at most the number of observations."),
    )
    for failed, msg in checks:
        if failed:
            raise SizeIncompatible(msg)


def _coerce(a, two_d: bool):
    """np.array(...) as the reference does (:158-161), but CUDA / CPU tensors pass through."""
    if isinstance(a, torch.Tensor):
        return a
    arr = np.array(a)
    if arr.dtype not in (np.float64, np.float32):      # float32 stays: it is widened on the device, after the copy
        arr = arr.astype(np.float64)
    if not two_d and arr.ndim != 1:
        raise ValueError("y_train / y_test must be one-dimensional")
    return arr


def _to_dev(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float64)
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64))).to(device)


# ---------------------------------------------------------------------------
# the entry point
# ---------------------------------------------------------------------------
def ls_spa(X_train, X_test, y_train, y_test, reg: float = 0.0, method: str | None = None,
           batch_size: int | None = None, num_batches: int | None = None, tolerance: float = 1e-2,
           seed: int = 42, return_history: bool | None = None, *, max_samples: int | None = None,
           perms=None, antithetical: bool | None = None,
           return_attribution_history: bool | None = None, process_group=None,
           row_sharded: bool = False) -> ShapleyResults:
    """Estimate the Shapley attribution of the out-of-sample R^2 of a least-squares fit.

    Reference-code keywords (ls_spa/ls_spa.py:122-133) select the reference's behaviour:
    p < 9 -> all p! permutations, no antithetic pairs, no error estimates (:170-173);
    otherwise up to ``max_samples`` (default 2**13) samples of
    ``default_rng(seed).permutation(p)`` (:175) in batches of ``batch_size`` (default 2**8),
    antithetic pairs on (:132), stop when the estimated error < ``tolerance`` (:229).

    README keywords (README.md:96-106): ``method`` in {'random', 'permutohedron', 'argsort',
    'exact'} (None -> 'argsort' if p > 10 else 'exact'), ``batch_size`` (2**7) x
    ``num_batches`` (2**7) samples, ``return_history``.  'exact' enumerates all p!
    permutations without antithetic pairs and reports a zero error.

    With an initialised ``torch.distributed`` NCCL group (or ``process_group=``) the rows of
    the reduction and the sample batches are sharded across the ranks; every rank returns
    the same result.  ``row_sharded=True`` means the inputs are already the local row shard.
    """
    if max_samples is not None and num_batches is not None:
        raise TypeError("pass either max_samples (reference code) or num_batches (README), not both")
    if return_history is not None and return_attribution_history is not None:
        raise TypeError("pass either return_history or return_attribution_history, not both")
    if method is not None and method not in METHODS:
        raise ValueError(f"unknown method {method!r}; expected one of {METHODS}")
    if method is not None and perms is not None:
        raise TypeError("pass either method or perms, not both")
    want_history = bool(return_history) or bool(return_attribution_history)

    X_train, X_test = _coerce(X_train, True), _coerce(X_test, True)
    y_train, y_test = _coerce(y_train, False), _coerce(y_test, False)
    validate_data(X_train, X_test, y_train, y_test)
    p = int(X_train.shape[1])

    device = ops.require_cuda()
    backend = engine.CudaBackend(device)
    coll = engine.Collective(process_group)

    readme_mode = method is not None or num_batches is not None
    penultimate = False
    if perms is not None:
        # reference :176-177 -- run the given stream to exhaustion (or tolerance)
        bs = 2 ** 8 if batch_size is None else int(batch_size)
        anti = True if antithetical is None else bool(antithetical)
        total, estimate, meth = None, p >= 9, None
    elif readme_mode:
        meth = method if method is not None else ("argsort" if p > 10 else "exact")
        bs = 2 ** 7 if batch_size is None else int(batch_size)
        nb = 2 ** 7 if num_batches is None else int(num_batches)
        total = bs * nb
        anti = True if antithetical is None else bool(antithetical)
        estimate = p >= 9
        if meth == "exact":
            total, anti, estimate = None, False, False
    else:
        # reference-code mode (:169-175)
        bs = 2 ** 8 if batch_size is None else int(batch_size)
        total = 2 ** 13 if max_samples is None else int(max_samples)
        anti = True if antithetical is None else bool(antithetical)
        if p < 9:
            meth, total, anti, estimate, bs = "exact", None, False, False, 2 ** 8
        else:
            meth, estimate, penultimate = "random", True, True
    if bs < 1:
        raise ValueError("batch_size must be positive")

    # The Sobol-based generators spend a few ms of host time in scipy building their scrambled
    # direction numbers: do that on a helper thread while the reduction runs on the device.
    if perms is None and meth in ("argsort", "permutohedron"):
        def _build_source():
            torch.cuda.set_device(device)
            src = make_source(meth, p, seed, total, device, perms=None)
            done = torch.cuda.Event()
            done.record()                      # the helper thread uploads on its own current stream
            return src, done
        pending = _source_pool().submit(_build_source)

        def get_source():
            src, done = pending.result()
            torch.cuda.current_stream().wait_event(done)
            return src
    else:
        ready = make_source(meth, p, seed, total, device, perms=perms)
        get_source = lambda: ready
    cfg = engine.JobConfig(p=p, batch_size=bs, max_samples=total, tolerance=float(tolerance),
                           seed=_seed_int(seed), antithetical=anti, estimate_errors=estimate,
                           return_history=want_history, penultimate_check=penultimate)

    # Host-resident test rows, single process, device permutation source, bounded job: factor the
    # permutations of the first super-batches (train side only) while the test rows cross PCIe.
    pre = None
    test_on_host = not (isinstance(X_test, torch.Tensor) and X_test.is_cuda)
    if (coll.world == 1 and test_on_host and perms is None and meth != "exact" and total is not None
            and ops.split_route_supported(p) and os.environ.get("LSSPA_SPLIT_ROUTE", "1") != "0"):
        esize = 4 if getattr(X_test, "dtype", None) in (np.float32, torch.float32) else 8
        pre = engine.Prefactor(backend, cfg, get_source, esize * int(X_test.shape[0]) * (p + 1))
    # host-side preparation that needs nothing from the reduction runs while the device is still busy with it
    # (reduce_problem calls this right before it waits for the reduction's flags)
    ready_early = {}

    def prepare():
        ready_early["est"] = backend.make_estimator(cfg)
        if pre is None:
            ready_early["source"] = get_source()

    prob = engine.reduce_problem(backend, coll, X_train, X_test, y_train, y_test, float(reg), p,
                                 row_sharded=row_sharded, prefactor=pre, prepare=prepare)
    if pre is not None and pre.source is not None:
        source = pre.source
    else:
        source = ready_early["source"] if "source" in ready_early else get_source()
    # the epilogue (theta, R^2) depends on the reduced problem only: its kernel goes in ahead of the sample
    # loop and is read after it, instead of one more launch-and-wait at the end of the job
    early = backend.theta_r2_start(prob) if (hasattr(backend, "theta_r2_start") and engine.HOST_OVERLAP) else None
    res, history, done = engine.run_samples(backend, coll, prob, source, cfg, pre=pre, est=ready_early.get("est"))
    if getattr(source, "host_generator", None) is not None:
        source.sync_generator()      # the caller's generator moves past the permutations drawn
    if done == 0 and p >= 9:
        raise ValueError("no permutations were supplied")
    theta, r2 = backend.theta_r2_finish(early) if early is not None else backend.theta_r2(prob)

    never = res["n_history"] == 0
    return ShapleyResults(
        attribution=res["mean"],
        theta=theta,
        overall_error=0.0 if never else res["overall_error"],
        attribution_errors=np.zeros(p) if never else res["attribution_errors"],
        r_squared=np.float64(r2),
        error_history=res["error_history"],
        attribution_history=history if want_history else None,
    )


# ---------------------------------------------------------------------------
# the helper functions the reference also exports, each backed by the device kernels
# ---------------------------------------------------------------------------
def reduce_data(X_train, X_test, y_train, y_test, reg):
    """Device TSQR twin of reference reduce_data (ls_spa/ls_spa.py:290-318).
    Returns (R_tr, R_te, c_tr, c_te) as numpy arrays; R factors are unique up to row signs."""
    X_train, X_test = _coerce(X_train, True), _coerce(X_test, True)
    y_train, y_test = _coerce(y_train, False), _coerce(y_test, False)
    p = int(X_train.shape[1])
    backend = engine.CudaBackend(ops.require_cuda())
    prob = engine.reduce_problem(backend, engine.Collective(None), X_train, X_test, y_train, y_test,
                                 float(reg), p)
    m = min(int(X_test.shape[0]), p)
    return (prob.R_tr_cm.t().cpu().numpy(), prob.R_te_cm.t()[:m].cpu().numpy(),
            prob.c_tr.cpu().numpy(), prob.c_te[:m].cpu().numpy())


def square_shapley(X_train, X_test, y_train, y_test, y_norm_sq, perm):
    """Device twin of reference square_shapley (ls_spa/ls_spa.py:256-287): lift vector of one
    permutation given the *reduced* factors."""
    device = ops.require_cuda()
    prob = ops.ReducedProblem(_to_dev(X_train, device), _to_dev(y_train, device), _to_dev(X_test, device),
                              _to_dev(y_test, device), float(y_norm_sq))
    pt = torch.as_tensor(np.asarray(perm).astype(np.int32)).reshape(1, -1).to(device)
    if pt.shape[1] != prob.p:
        raise ValueError("perm must have p entries")
    ops.perms_validate(pt)
    return ops.lifts(prob, pt, False)[0].cpu().numpy()


def merge_sample_mean(old_mean, new_mean, old_N, new_N):
    """Device twin of reference merge_sample_mean (ls_spa/ls_spa.py:103-108)."""
    device = ops.require_cuda()
    m = _to_dev(old_mean, device).clone()
    ops.merge_moments(m, None, old_N, _to_dev(new_mean, device), None, new_N)
    return m.cpu().numpy()


def merge_sample_cov(old_mean, new_mean, old_cov, new_cov, old_N, new_N):
    """Device twin of reference merge_sample_cov (ls_spa/ls_spa.py:111-119)."""
    device = ops.require_cuda()
    m = _to_dev(old_mean, device).clone()
    c = _to_dev(old_cov, device).contiguous().clone()
    ops.merge_moments(m, c, old_N, _to_dev(new_mean, device), _to_dev(new_cov, device).contiguous(), new_N)
    return c.cpu().numpy()


def error_estimates(rng, cov):
    """Device twin of reference error_estimates (ls_spa/ls_spa.py:321-341): 0.95 quantiles of
    |z| per feature and of |z|_2 over 1024 draws z ~ N(0, cov).

    cov is factorised on the device by a Cholesky that skips vanishing pivots (it may be singular:
    the reference falls back from Cholesky to SVD there) and the draws come from the device's
    counter-based Gaussian stream, keyed by one integer taken from ``rng``: the values agree with the
    reference statistically, not bit for bit."""
    device = ops.require_cuda()
    seed = int(rng.integers(0, 2 ** 63 - 1)) if hasattr(rng, "integers") else int(rng)
    feat, overall = ops.error_draws_quantiles(_to_dev(cov, device), seed)
    return feat.cpu().numpy(), float(overall.item())
