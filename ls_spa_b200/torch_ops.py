"""``torch.ops.ls_spa_b200.*`` -- the tensor-level entry points of the hot path as PyTorch custom ops
(torch.library), each a thin wrapper of one C-ABI call of include/lsspa.h (BASELINE.json north_star:
"the host side is Python, calling PyTorch custom ops backed by a thin C-ABI").

    torch.ops.ls_spa_b200.gram_reduce(X, y, divisor)                      -> scaled Gram matrix of [X | y]
    torch.ops.ls_spa_b200.perms_exact / perms_sobol_argsort / perms_permutohedron(...) -> int32 (count, p)
    torch.ops.ls_spa_b200.lifts(R_tr_cm, c_tr, R_te_cm, c_te, y_norm_sq, perms, antithetical)   Householder route
    torch.ops.ls_spa_b200.lifts_chol(gram, R_te_scaled_cm, c_te, y_norm_sq, perms, antithetical)  Cholesky route
    torch.ops.ls_spa_b200.theta_r2(R_tr_cm, c_tr, R_te_cm, c_te, y_norm_sq)                      -> (p + 1,)

All tensors are CUDA tensors; there is no CPU kernel registered (a CPU tensor raises).  The engine
(engine.py / ops.py) calls the same C ABI directly through ctypes: a dispatcher round trip costs ~10 us
and a benchmark step issues ~60 launches; results are identical (tests/test_gpu_parity.py).
"""

from __future__ import annotations

import torch

from . import ops as _ops
from ._cabi import LsSpaCudaError, check

_lib_def = torch.library.Library("ls_spa_b200", "DEF")
_lib_def.define("gram_reduce(Tensor X, Tensor y, float divisor) -> Tensor")
_lib_def.define("perms_exact(int p, int first_rank, int count, Tensor like) -> Tensor")
_lib_def.define("perms_sobol_argsort(Tensor sv, Tensor shift, int bits, int p, int first_index, int count) -> Tensor")
_lib_def.define("perms_permutohedron(Tensor sv, Tensor shift, int bits, int p, int first_index, int count) -> Tensor")
_lib_def.define("lifts(Tensor R_tr_cm, Tensor c_tr, Tensor R_te_cm, Tensor c_te, float y_norm_sq, Tensor perms, "
                "bool antithetical) -> Tensor")
_lib_def.define("lifts_chol(Tensor gram, Tensor R_te_scaled_cm, Tensor c_te, float y_norm_sq, Tensor perms, "
                "bool antithetical) -> Tensor")
_lib_def.define("theta_r2(Tensor R_tr_cm, Tensor c_tr, Tensor R_te_cm, Tensor c_te, float y_norm_sq) -> Tensor")

_impl = torch.library.Library("ls_spa_b200", "IMPL", "CUDA")


def _f64(t, name):
    if t.dtype != torch.float64 or not t.is_cuda:
        raise LsSpaCudaError(f"{name} must be a float64 CUDA tensor")
    return t.contiguous()


def _gram_reduce(X, y, divisor):
    p = int(X.shape[1])
    if not _ops.gram_supported(p):
        raise LsSpaCudaError("gram_reduce: p + 1 <= 112 (wide problems: ops.GramBig)")
    fac = _ops.CholQR2(p, divisor, X.device)
    fac.add_chunk(X, y)
    return fac.gram()


def _perms_exact(p, first_rank, count, like):
    return _ops.perms_exact(p, first_rank, count, like.device)


def _perms_sobol_argsort(sv, shift, bits, p, first_index, count):
    return _ops.perms_sobol_argsort(p, sv, shift, bits, first_index, count)


def _perms_permutohedron(sv, shift, bits, p, first_index, count):
    return _ops.perms_permutohedron(p, sv, shift, bits, first_index, count)


def _lifts(R_tr_cm, c_tr, R_te_cm, c_te, y_norm_sq, perms, antithetical):
    lib = _ops._lib()
    R_tr_cm, c_tr, R_te_cm, c_te = (_f64(t, n) for t, n in ((R_tr_cm, "R_tr_cm"), (c_tr, "c_tr"), (R_te_cm, "R_te_cm"),
                                                               (c_te, "c_te")))
    perms = perms.contiguous()
    count, p = perms.shape
    out = torch.empty((count, p), dtype=torch.float64, device=perms.device)
    nbytes = lib.lsspa_lifts_workspace_bytes(p, count)
    ws = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=perms.device)
    check(lib.lsspa_lifts(p, R_tr_cm.data_ptr(), c_tr.data_ptr(), R_te_cm.data_ptr(), c_te.data_ptr(), float(y_norm_sq),
                          perms.data_ptr(), count, 1 if antithetical else 0, out.data_ptr(), ws.data_ptr(), int(nbytes),
                          _ops._stream()), "lsspa_lifts")
    return out


def _lifts_chol(gram, R_te_scaled_cm, c_te, y_norm_sq, perms, antithetical):
    lib = _ops._lib()
    gram, R_te_scaled_cm, c_te = _f64(gram, "gram"), _f64(R_te_scaled_cm, "R_te_scaled_cm"), _f64(c_te, "c_te")
    perms = perms.contiguous()
    count, p = perms.shape
    if not lib.lsspa_lifts_chol_supported(p):
        raise LsSpaCudaError("lifts_chol: 17 <= p <= 152 (wide problems: ops.lifts with a ReducedProblem)")
    out = torch.empty((count, p), dtype=torch.float64, device=perms.device)
    check(lib.lsspa_lifts_chol(p, gram.data_ptr(), R_te_scaled_cm.data_ptr(), c_te.data_ptr(), float(y_norm_sq),
                               perms.data_ptr(), count, 1 if antithetical else 0, out.data_ptr(), _ops._stream()),
          "lsspa_lifts_chol")
    return out


def _theta_r2(R_tr_cm, c_tr, R_te_cm, c_te, y_norm_sq):
    lib = _ops._lib()
    p = int(c_tr.numel())
    out = torch.empty(p + 1, dtype=torch.float64, device=c_tr.device)
    nbytes = lib.lsspa_theta_r2_workspace_bytes(p)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=c_tr.device)
    check(lib.lsspa_theta_r2(p, _f64(R_tr_cm, "R_tr_cm").data_ptr(), _f64(c_tr, "c_tr").data_ptr(),
                             _f64(R_te_cm, "R_te_cm").data_ptr(), _f64(c_te, "c_te").data_ptr(), float(y_norm_sq),
                             out.data_ptr(), ws.data_ptr(), nbytes, _ops._stream()), "lsspa_theta_r2")
    return out


_impl.impl("gram_reduce", _gram_reduce)
_impl.impl("perms_exact", _perms_exact)
_impl.impl("perms_sobol_argsort", _perms_sobol_argsort)
_impl.impl("perms_permutohedron", _perms_permutohedron)
_impl.impl("lifts", _lifts)
_impl.impl("lifts_chol", _lifts_chol)
_impl.impl("theta_r2", _theta_r2)

OPS = ("gram_reduce", "perms_exact", "perms_sobol_argsort", "perms_permutohedron", "lifts", "lifts_chol", "theta_r2")
