"""ctypes binding of include/hpgv_epi.h.  Loading fails loudly when the CUDA
library has not been built: there is no CPU fallback behind this package."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# HPGV_EPI_LIB: development override (A/B runs of kernel variants built side by side); never a CPU path
LIB_PATH = os.environ.get("HPGV_EPI_LIB") or os.path.join(HERE, "libhpgv_epi.so")

MODEL_DTYPE = np.dtype([("accuracy", "<f8"), ("snp", "<i4", (3,)), ("risky_mask", "<u4"), ("conf", "<u4", (4,))])
assert MODEL_DTYPE.itemsize == 40

SUBSET_TESTING, SUBSET_TRAINING = 0, 1
# enum eval_function (model.h:84) + the documented CA formula as an extension (include/hpgv_epi.h)
EVAL_CA, EVAL_BA, EVAL_WBA, EVAL_GAMMA, EVAL_TAU_B, EVAL_CA_TRUE = 0, 1, 2, 3, 4, 5
UINT64_MAX = (1 << 64) - 1


class Layout(C.Structure):
    _fields_ = [("num_folds", C.c_int), ("num_segments", C.c_int), ("num_blocks", C.c_int), ("block_words", C.c_int),
                ("count_bits", C.c_int), ("plane_bytes", C.c_int64), ("words_per_class_row", C.c_int),
                ("num_chunks", C.c_int), ("chunk_blocks", C.c_int), ("row_words", C.c_int)]


class HpgvError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"hpgv_epi error {code}: {msg}")
        self.code = code


_lib = None

# every symbol include/hpgv_epi.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "hpgv_epi_create", "hpgv_epi_destroy", "hpgv_epi_last_error", "hpgv_epi_set_stream", "hpgv_epi_launch_count",
    "hpgv_epi_load_dataset_host", "hpgv_epi_load_dataset_device", "hpgv_epi_load_dataset_file", "hpgv_epi_dataset_dims",
    "hpgv_epi_set_folds", "hpgv_epi_k_folds", "hpgv_epi_search", "hpgv_epi_search_device", "hpgv_epi_merge_device",
    "hpgv_epi_num_combinations", "hpgv_epi_eval", "hpgv_epi_unpack_masks", "hpgv_epi_run_host", "hpgv_epi_layout",
    "hpgv_epi_pipe_peak", "hpgv_epi_last_search_ms", "hpgv_epi_search_times",
    "hpgv_epi_set_eval_function", "hpgv_epi_confusion", "hpgv_epi_high_risk", "hpgv_epi_evaluate",
    "hpgv_epi_debug_counters", "hpgv_epi_merge_host", "hpgv_epi_device_count",
]


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m hpg_variant_b200.build` (needs nvcc). "
            "hpg_variant_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
    lib.hpgv_epi_create.argtypes = [i32, C.POINTER(vp)]
    lib.hpgv_epi_destroy.argtypes = [vp]
    lib.hpgv_epi_destroy.restype = None
    lib.hpgv_epi_last_error.argtypes = [vp]
    lib.hpgv_epi_last_error.restype = C.c_char_p
    lib.hpgv_epi_set_stream.argtypes = [vp, vp]
    lib.hpgv_epi_launch_count.argtypes = [vp]
    lib.hpgv_epi_launch_count.restype = i64
    lib.hpgv_epi_load_dataset_host.argtypes = [vp, vp, i64, i32, i32]
    lib.hpgv_epi_load_dataset_device.argtypes = [vp, vp, i64, i32, i32]
    lib.hpgv_epi_load_dataset_file.argtypes = [vp, C.c_char_p]
    lib.hpgv_epi_dataset_dims.argtypes = [vp, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]
    lib.hpgv_epi_set_folds.argtypes = [vp, i32, vp]
    lib.hpgv_epi_k_folds.argtypes = [i32, i32, i32, C.c_long, vp, vp]
    lib.hpgv_epi_search.argtypes = [vp, i32, i32, i32, u64, u64, vp]
    lib.hpgv_epi_search_device.argtypes = [vp, i32, i32, i32, u64, u64, vp]
    lib.hpgv_epi_merge_device.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp]
    lib.hpgv_epi_num_combinations.argtypes = [i64, i32]
    lib.hpgv_epi_num_combinations.restype = u64
    lib.hpgv_epi_eval.argtypes = [vp, i32, i32, i64, vp, vp, vp, vp, vp, vp]
    lib.hpgv_epi_unpack_masks.argtypes = [vp, i64, vp]
    lib.hpgv_epi_run_host.argtypes = [vp, vp, i64, i32, i32, i32, vp, i32, i32, i32, u64, u64, vp]
    lib.hpgv_epi_layout.argtypes = [vp, C.POINTER(Layout)]
    lib.hpgv_epi_last_search_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(i32)]
    lib.hpgv_epi_search_times.argtypes = [vp, i32, C.POINTER(C.c_float)]
    lib.hpgv_epi_pipe_peak.argtypes = [vp, i32, i32, C.POINTER(C.c_double)]
    lib.hpgv_epi_set_eval_function.argtypes = [vp, i32]
    lib.hpgv_epi_confusion.argtypes = [vp, i32, i32, i64, vp, vp, vp, vp]
    lib.hpgv_epi_high_risk.argtypes = [vp, vp, vp, i64, i32, i32, vp]
    lib.hpgv_epi_evaluate.argtypes = [vp, i32, i64, vp, vp]
    lib.hpgv_epi_debug_counters.argtypes = [vp, vp, i32]
    lib.hpgv_epi_merge_host.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp]
    lib.hpgv_epi_device_count.argtypes = []
    _lib = lib
    return lib
