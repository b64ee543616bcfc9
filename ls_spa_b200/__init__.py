"""B200-native LS-SPA (least-squares Shapley performance attribution).

Exports the same names as the reference package (ls_spa/__init__.py:1 star-exports
ls_spa/ls_spa.py), so ``from ls_spa_b200 import ls_spa, ShapleyResults`` is a drop-in
for ``from ls_spa import ...``.  The numeric path is hand-written sm_100a CUDA behind
the C ABI in include/lsspa.h; importing the package does not need a GPU, calling
into it does.
"""

from ._cabi import LsSpaCudaError
from . import torch_ops  # noqa: F401  (registers torch.ops.ls_spa_b200.*)
from .api import (ShapleyResults, SizeIncompatible, error_estimates, ls_spa, merge_sample_cov,
                  merge_sample_mean, reduce_data, square_shapley, validate_data)

__all__ = [
    "ls_spa", "ShapleyResults", "SizeIncompatible", "validate_data", "merge_sample_mean",
    "merge_sample_cov", "square_shapley", "reduce_data", "error_estimates", "LsSpaCudaError",
]
