"""CPU: host-side logic and the C-ABI surface (no compute calls; there is no GPU here)."""

import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from ls_spa_b200 import _cabi
    header = open(os.path.join(ROOT, "include", "lsspa.h")).read()
    declared = set(re.findall(r"LSSPA_API[^;(]*?\b(lsspa_\w+)\s*\(", header))
    assert len(declared) >= 25
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _cabi.load().lsspa_abi_version() == 1
    assert b"bad argument" in _cabi.load().lsspa_status_string(-1)
    # pure size queries work without a device
    assert _cabi.load().lsspa_tsqr_slot_doubles(100) == 101 * 101 + 8
    assert _cabi.load().lsspa_estimator_partial_doubles(3) == 8 + 3 + 9 + 1024 + 3 * 1024


def test_no_cpu_fallback():
    import torch
    import ls_spa_b200 as L
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    x = np.random.default_rng(0).standard_normal((20, 3))
    with pytest.raises(L.LsSpaCudaError):
        L.ls_spa(x, x, x[:, 0], x[:, 0])
    with pytest.raises(L.LsSpaCudaError):
        L.merge_sample_mean(np.zeros(3), np.ones(3), 1, 1)


def test_validation_happens_before_any_device_work():
    import ls_spa_b200 as L
    with pytest.raises(L.SizeIncompatible) as e:
        L.ls_spa(np.zeros((5, 3)), np.zeros((5, 4)), np.zeros(5), np.zeros(5))
    assert "same number of columns" in e.value.message
    with pytest.raises(L.SizeIncompatible):
        L.ls_spa(np.zeros((5, 3)), np.zeros((5, 3)), np.zeros(4), np.zeros(5))
    with pytest.raises(L.SizeIncompatible):
        L.ls_spa(np.zeros((5, 3)), np.zeros((5, 3)), np.zeros(5), np.zeros(6))
    with pytest.raises(L.SizeIncompatible):
        L.ls_spa(np.zeros((2, 3)), np.zeros((5, 3)), np.zeros(2), np.zeros(5))
    with pytest.raises(TypeError):
        L.ls_spa(np.zeros((5, 3)), np.zeros((5, 3)), np.zeros(5), np.zeros(5), max_samples=4, num_batches=2)
    with pytest.raises(ValueError):
        L.ls_spa(np.zeros((5, 3)), np.zeros((5, 3)), np.zeros(5), np.zeros(5), method="sobol")
    with pytest.raises(TypeError):
        L.ls_spa(np.zeros((5, 3)), np.zeros((5, 3)), np.zeros(5), np.zeros(5), return_history=True,
                 return_attribution_history=True)


def test_exported_names_match_reference_package():
    import ls_spa_b200 as L
    for name in ("ls_spa", "ShapleyResults", "SizeIncompatible", "validate_data", "merge_sample_mean",
                 "merge_sample_cov", "square_shapley", "reduce_data", "error_estimates"):
        assert hasattr(L, name), name
    import dataclasses
    assert [f.name for f in dataclasses.fields(L.ShapleyResults)] == [
        "attribution", "theta", "overall_error", "attribution_errors", "r_squared", "error_history",
        "attribution_history"]


def test_repr_matches_reference_dashboard():
    import ls_spa_b200 as L
    from conftest import load_golden
    g = load_golden("toy")
    r = L.ShapleyResults(g["default_attribution"], g["default_theta"], 0.0, np.zeros(3),
                         float(g["default_r_squared"]), np.zeros(0), None)
    assert repr(r) == str(g["default_repr"])
    w = L.ShapleyResults(np.arange(7) / 3.0, -np.arange(7) / 7.0, 1.25e-3, np.zeros(7), 0.5, np.zeros(0), None)
    assert repr(w) == str(g["wide_repr"])


def test_split_batches_and_runs():
    from ls_spa_b200.engine import contiguous_runs, split_batches, target_samples
    assert split_batches(0, 10, 4) == [(0, 4), (4, 4), (8, 2)]
    assert split_batches(6, 10, 4) == [(6, 2), (8, 4), (12, 4)]
    # the reference's extra estimate at i == max_samples - 1 (ls_spa/ls_spa.py:222)
    assert split_batches(0, 4, 2, extra_cut=3) == [(0, 2), (2, 1), (3, 1)]
    assert split_batches(0, 2048, 256, extra_cut=2047)[-2:] == [(1792, 255), (2047, 1)]
    assert len(split_batches(0, 2048, 256, extra_cut=2047)) == 9
    runs, per = contiguous_runs(10, 4)
    assert per == 3 and runs == [(0, 3), (3, 6), (6, 9), (9, 10)]
    runs, per = contiguous_runs(2, 4)
    assert runs == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert target_samples(100) == 32768 and target_samples(1000) == 592 and target_samples(150) == 14563 and target_samples(3) == 131072


def test_explicit_source_on_cpu_tensors():
    import torch
    from ls_spa_b200.samplers import ExplicitSource
    perms = [np.random.default_rng(i).permutation(6) for i in range(7)]
    src = ExplicitSource(6, iter(perms), torch.device("cpu"))
    assert src.total is None
    a = src.take(4)
    b = src.take(4)
    assert a.shape == (4, 6) and b.shape == (3, 6) and src.total == 7 and src.exhausted
    assert np.array_equal(torch.cat([a, b]).numpy(), np.array(perms))
    src = ExplicitSource(6, np.array(perms), torch.device("cpu"))
    assert src.total == 7 and src.take(100).shape == (7, 6)


def test_superbatch_plan():
    """Jobs that can stop early ramp their super-batches up (2048 samples per rank, x4 per round);
    jobs that cannot (tolerance 0 or no error estimates) start at full size; the plan is what the
    split route prepares its factorisations against, so it must be a pure function of the config."""
    from ls_spa_b200.engine import JobConfig, superbatch_geometry, target_samples
    base = dict(p=100, batch_size=128, max_samples=65536, seed=1, antithetical=True, return_history=False)
    full = target_samples(100)
    limit, bs, size = superbatch_geometry(JobConfig(tolerance=1e-4, estimate_errors=True, **base), 1, None)
    assert (limit, bs) == (65536, 128)
    assert [size(i) for i in range(4)] == [2048, 8192, full, full]
    assert [size(i) for i in range(3)] == [2048, 8192, full]          # calling it again changes nothing
    _, _, size8 = superbatch_geometry(JobConfig(tolerance=1e-4, estimate_errors=True, **base), 8, None)
    assert [size8(i) for i in range(3)] == [8 * 2048, 8 * 8192, 8 * full]
    _, _, flat = superbatch_geometry(JobConfig(tolerance=0.0, estimate_errors=True, **base), 1, None)
    assert flat(0) == flat(5) == full
    limit, bs, nosplit = superbatch_geometry(JobConfig(tolerance=1e-2, estimate_errors=False, **base), 1, 5000)
    assert limit == 5000 and bs == 1024 and nosplit(0) == full


def test_superbatch_memory_cap():
    """Tiny batches (the reference's tests use batch_size=2) must not blow the per-batch partial blocks
    up: the batches of one super-batch are capped by a byte budget."""
    from ls_spa_b200.engine import JobConfig, batch_cap, superbatch_geometry, target_samples
    cap = batch_cap(100, True)
    per_batch = 8 * (8 + 100 + 100 * 100 + 1024 + 1024 * 100) + 8 * 101 * 1024
    assert cap * per_batch <= 2 << 30 < (cap + 1) * per_batch
    assert batch_cap(100, True, free_bytes=1 << 30) == (1 << 28) // per_batch
    assert batch_cap(30000, True) == 1                      # never zero
    cfg = JobConfig(p=100, batch_size=2, max_samples=1 << 20, tolerance=0.0, seed=1, antithetical=True,
                    estimate_errors=True, return_history=False)
    _, bs, size = superbatch_geometry(cfg, 1, None)
    assert bs == 2 and size(0) == 2 * cap < target_samples(100)
    _, _, size4 = superbatch_geometry(cfg, 4, None)
    assert size4(0) == 4 * 2 * cap                          # per rank
    _, _, small = superbatch_geometry(cfg, 1, None, max_batches=7)
    assert small(0) == 14
    cfg10 = JobConfig(p=10, batch_size=1, max_samples=None, tolerance=0.0, seed=1, antithetical=False,
                      estimate_errors=True, return_history=False)
    _, _, s10 = superbatch_geometry(cfg10, 1, None)
    assert s10(0) == min(target_samples(10), batch_cap(10, True))


def test_torch_custom_ops_registered():
    import pytest
    import torch
    """The tensor-level entry points exist as PyTorch custom ops (torch.library) with CUDA-only kernels."""
    import ls_spa_b200  # noqa: F401
    from ls_spa_b200 import torch_ops
    for name in torch_ops.OPS:
        op = getattr(torch.ops.ls_spa_b200, name)
        assert op is not None and "ls_spa_b200::" + name in str(op.default._schema)
    with pytest.raises((NotImplementedError, RuntimeError)):      # no CPU kernel: there is no CPU fallback
        torch.ops.ls_spa_b200.theta_r2(torch.eye(3, dtype=torch.float64), torch.ones(3, dtype=torch.float64),
                                       torch.eye(3, dtype=torch.float64), torch.ones(3, dtype=torch.float64), 1.0)
