"""ctypes binding of libls_spa_b200.so (the C ABI declared in include/lsspa.h).

There is deliberately no fallback: if the library is missing or cannot be loaded
the import of any compute entry point raises, and on a machine without a CUDA
device every compute call raises ``LsSpaCudaError``.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libls_spa_b200.so")

c_i32, c_i64, c_u64, c_f64 = C.c_int, C.c_int64, C.c_uint64, C.c_double
vp, sz = C.c_void_p, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/lsspa.h declares
SIGNATURES = {
    "lsspa_abi_version": (c_i32, []),
    "lsspa_status_string": (C.c_char_p, [c_i32]),
    "lsspa_device_sm_count": (c_i32, []),
    "lsspa_device_smem_optin": (c_i32, []),
    "lsspa_tsqr_slot_doubles": (c_i64, [c_i32]),
    "lsspa_tsqr_num_parts": (c_i32, [c_i32, c_i64]),
    "lsspa_tsqr_rows": (c_i32, [vp, c_i64, vp, c_i64, c_i32, c_f64, vp, c_i32, vp]),
    "lsspa_tsqr_merge": (c_i32, [vp, c_i32, c_i32, c_i32, vp, vp]),
    "lsspa_gram_supported": (c_i32, [c_i32]),
    "lsspa_gram_slot_doubles": (c_i64, [c_i32]),
    "lsspa_gram_rinv_doubles": (c_i64, [c_i32]),
    "lsspa_gram_num_parts": (c_i32, [c_i32, c_i64, c_i32]),
    "lsspa_gram_rows": (c_i32, [vp, c_i64, vp, c_i64, c_i32, vp, vp, c_i32, vp]),
    "lsspa_gram_finish": (c_i32, [vp, c_i32, c_i32, c_f64, vp, vp]),
    "lsspa_chol_factor": (c_i32, [vp, c_i32, vp, vp, vp, vp]),
    "lsspa_gram_add_ridge": (c_i32, [vp, c_i32, c_f64, vp]),
    "lsspa_gram_big_supported": (c_i32, [c_i32]),
    "lsspa_gram_big_num_splits": (c_i32, [c_i32, c_i64]),
    "lsspa_gram_big_part_doubles": (c_i64, [c_i32]),
    "lsspa_gram_big_rows": (c_i32, [vp, c_i64, vp, c_i64, c_i32, vp, c_i32, vp]),
    "lsspa_gram_big_accumulate": (c_i32, [vp, c_i32, c_i32, vp, vp]),
    "lsspa_gram_big_factor_workspace_bytes": (sz, [c_i32]),
    "lsspa_gram_big_factor": (c_i32, [vp, c_i32, c_f64, c_f64, vp, vp, sz, vp, vp]),
    "lsspa_chol_factor_gram": (c_i32, [vp, c_i32, vp, vp, vp, vp, vp]),
    "lsspa_tri_product": (c_i32, [vp, vp, c_i32, vp, vp, vp]),
    "lsspa_perms_exact": (c_i32, [c_i32, c_u64, c_i64, vp, vp]),
    "lsspa_perms_pcg64_workspace_bytes": (sz, [c_i32, c_i64]),
    "lsspa_perms_pcg64": (c_i32, [c_i32, vp, c_i64, vp, vp, sz, vp, vp]),
    "lsspa_perms_sobol_argsort": (c_i32, [c_i32, vp, vp, c_i32, c_u64, c_i64, vp, vp]),
    "lsspa_perms_permutohedron": (c_i32, [c_i32, vp, vp, c_i32, c_u64, c_i64, vp, vp]),
    "lsspa_perms_validate": (c_i32, [c_i32, vp, c_i64, vp, vp]),
    "lsspa_lifts_workspace_bytes": (sz, [c_i32, c_i64]),
    "lsspa_lifts": (c_i32, [c_i32, vp, vp, vp, vp, c_f64, vp, c_i64, c_i32, vp, vp, sz, vp]),
    "lsspa_lifts_chol_supported": (c_i32, [c_i32]),
    "lsspa_lifts_gram_doubles": (c_i64, [c_i32]),
    "lsspa_lifts_gram": (c_i32, [c_i32, vp, vp, vp, vp]),
    "lsspa_lifts_chol": (c_i32, [c_i32, vp, vp, vp, c_f64, vp, c_i64, c_i32, vp, vp]),
    "lsspa_lifts_big_supported": (c_i32, [c_i32]),
    "lsspa_lifts_big_workspace_bytes": (sz, [c_i32, c_i64, c_i32, sz]),
    "lsspa_lifts_big": (c_i32, [c_i32, vp, vp, vp, c_f64, vp, c_i64, c_i32, vp, vp, sz, vp, vp]),
    "lsspa_lifts_chol_factor_doubles": (c_i64, [c_i32]),
    "lsspa_lifts_chol_factor": (c_i32, [c_i32, vp, vp, c_i64, c_i32, vp, vp]),
    "lsspa_lifts_chol_eliminate": (c_i32, [c_i32, vp, vp, vp, c_f64, vp, c_i64, c_i32, vp, vp]),
    "lsspa_estimator_state_bytes": (sz, [c_i32]),
    "lsspa_estimator_partial_doubles": (c_i64, [c_i32]),
    "lsspa_estimator_max_batches": (c_i32, [c_i32]),
    "lsspa_estimator_partials": (c_i32, [c_i32, vp, vp, c_i32, c_u64, c_i32, vp, vp]),
    "lsspa_estimator_absorb": (c_i32, [vp, c_i32, c_i32, c_f64, vp, vp, c_i32, c_i32, c_i32, vp, c_i32, vp]),
    "lsspa_estimator_block_total": (c_i32, [c_i32, vp, c_i32, c_i32, vp, vp]),
    "lsspa_estimator_quantiles": (c_i32, [c_i32, vp, c_i32, vp, vp, vp]),
    "lsspa_estimator_errors_workspace_doubles": (c_i64, [c_i32, c_i32]),
    "lsspa_estimator_absorb_errors": (c_i32, [vp, c_i32, c_i32, c_f64, vp, vp, c_i32, c_i32, c_i32, c_i32, vp, vp, vp, vp]),
    "lsspa_error_draws": (c_i32, [c_i32, vp, c_u64, vp, vp, vp]),
    "lsspa_prefix_means": (c_i32, [c_i32, vp, c_i64, vp, c_f64, vp, vp]),
    "lsspa_merge_moments": (c_i32, [c_i32, vp, vp, c_f64, vp, vp, c_f64, vp]),
    "lsspa_theta_r2_workspace_bytes": (sz, [c_i32]),
    "lsspa_theta_r2": (c_i32, [c_i32, vp, vp, vp, vp, c_f64, vp, vp, sz, vp]),
}


class LsSpaCudaError(RuntimeError):
    """A C-ABI call returned a non-zero status, or the CUDA library is unusable."""


_lib = None


def load():
    """Load the shared library (once) and attach the signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LsSpaCudaError(
            f"{LIB_PATH} is missing: build it with `python -m ls_spa_b200.build` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.lsspa_abi_version() != 1:
        raise LsSpaCudaError("libls_spa_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().lsspa_status_string(status).decode()
        raise LsSpaCudaError(f"{what or 'lsspa call'} failed: {msg} (status {status})")
