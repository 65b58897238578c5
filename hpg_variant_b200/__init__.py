"""hpg_variant_b200 -- B200-native exhaustive epistasis search (MDR + k-fold CV),
a drop-in for the `hpg-var-gwas epi` hot path of opencb/hpg-variant.

The product is the CUDA library (csrc/ -> libhpgv_epi.so, C-ABI in
include/hpgv_epi.h) and the C host API mirroring the reference
(include/hpgv_epi_compat.h).  This package is the thin Python face used by the
tests and bench.py.  Importing the engine requires the built library; there is
no CPU fallback.
"""
from ._lib import (EVAL_BA, EVAL_CA, EVAL_CA_TRUE, EVAL_GAMMA, EVAL_TAU_B, EVAL_WBA, MODEL_DTYPE, SUBSET_TESTING,  # noqa: F401
                   SUBSET_TRAINING, UINT64_MAX, HpgvError)
from .engine import EpistasisEngine, k_folds, num_combinations  # noqa: F401
