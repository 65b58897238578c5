"""Python face of the C-ABI (include/hpgv_epi.h): one `EpistasisEngine` per GPU.

This is plumbing for tests and bench.py -- the product is the CUDA library and
the C host API (include/hpgv_epi_compat.h).  Names follow the reference's
domain: dataset, folds, combinations, models.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import MODEL_DTYPE, SUBSET_TESTING, SUBSET_TRAINING, UINT64_MAX, HpgvError, Layout


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


class EpistasisEngine:
    def __init__(self, device=-1):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.hpgv_epi_create(int(device), C.byref(h))
        if rc != 0:
            raise HpgvError(rc, self.lib.hpgv_epi_last_error(None).decode())
        self.h = h
        self.num_folds = None
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.lib.hpgv_epi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise HpgvError(rc, self.lib.hpgv_epi_last_error(self.h).decode())

    # -- plumbing ---------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        self._ck(self.lib.hpgv_epi_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    @property
    def launch_count(self):
        return int(self.lib.hpgv_epi_launch_count(self.h))

    # -- dataset ------------------------------------------------------------------
    def load_dataset(self, genotypes, num_affected, num_unaffected):
        """genotypes: uint8 [num_variants, A+U] host array (cases first)."""
        g = np.ascontiguousarray(genotypes, dtype=np.uint8)
        assert g.ndim == 2 and g.shape[1] == num_affected + num_unaffected
        self._keep = [g]
        self._ck(self.lib.hpgv_epi_load_dataset_host(self.h, _ptr(g), g.shape[0], num_affected, num_unaffected))

    def load_dataset_device(self, dev_ptr, num_variants, num_affected, num_unaffected):
        self._ck(self.lib.hpgv_epi_load_dataset_device(self.h, C.c_void_p(dev_ptr), num_variants, num_affected, num_unaffected))

    def load_dataset_file(self, path):
        self._ck(self.lib.hpgv_epi_load_dataset_file(self.h, str(path).encode()))

    def dims(self):
        nv, a, u = C.c_int64(), C.c_int(), C.c_int()
        self._ck(self.lib.hpgv_epi_dataset_dims(self.h, C.byref(nv), C.byref(a), C.byref(u)))
        return nv.value, a.value, u.value

    # -- folds ----------------------------------------------------------------------
    def set_folds(self, num_folds, fold_of_sample):
        f = np.ascontiguousarray(fold_of_sample, dtype=np.int32)
        self._ck(self.lib.hpgv_epi_set_folds(self.h, int(num_folds), _ptr(f)))
        self.num_folds = int(num_folds)

    def layout(self):
        lay = Layout()
        self._ck(self.lib.hpgv_epi_layout(self.h, C.byref(lay)))
        return {k: getattr(lay, k) for k, _ in Layout._fields_}

    # -- search ---------------------------------------------------------------------
    def search(self, order, eval_subset=SUBSET_TRAINING, rank_size=50, first=0, last=UINT64_MAX):
        out = np.zeros((self.num_folds, rank_size), MODEL_DTYPE)
        self._ck(self.lib.hpgv_epi_search(self.h, order, eval_subset, rank_size, first, last, _ptr(out)))
        return out

    def search_device(self, order, eval_subset, rank_size, first, last, d_out_ptr):
        self._ck(self.lib.hpgv_epi_search_device(self.h, order, eval_subset, rank_size, first, last, C.c_void_p(d_out_ptr)))

    def merge_device(self, order, eval_subset, num_lists, rank_size, d_lists_ptr, d_out_ptr):
        self._ck(self.lib.hpgv_epi_merge_device(self.h, order, eval_subset, num_lists, self.num_folds, rank_size,
                                                C.c_void_p(d_lists_ptr), C.c_void_p(d_out_ptr)))

    def run_host(self, genotypes, num_affected, num_unaffected, num_folds, fold_of_sample, order,
                 eval_subset=SUBSET_TRAINING, rank_size=50, first=0, last=UINT64_MAX, out=None):
        """Whole path with HOST buffers in and out (what bench.py's e2e times)."""
        g = genotypes
        f = fold_of_sample
        if out is None:
            out = np.zeros((num_folds, rank_size), MODEL_DTYPE)
        self._ck(self.lib.hpgv_epi_run_host(self.h, _ptr(g), g.shape[0], num_affected, num_unaffected, num_folds, _ptr(f),
                                            order, eval_subset, rank_size, first, last, _ptr(out)))
        self.num_folds = num_folds
        return out

    # -- parity hooks -----------------------------------------------------------------
    def eval(self, order, combs, eval_subset=SUBSET_TRAINING):
        combs = np.ascontiguousarray(combs, dtype=np.int32).reshape(-1, order)
        n, F, Cc = combs.shape[0], self.num_folds, 3 ** order
        ca = np.zeros((n, F, Cc), np.int32)
        cu = np.zeros((n, F, Cc), np.int32)
        mask = np.zeros((n, F), np.uint32)
        conf = np.zeros((n, F, 4), np.uint32)
        ba = np.zeros((n, F), np.float64)
        self._ck(self.lib.hpgv_epi_eval(self.h, order, eval_subset, n, _ptr(combs), _ptr(ca), _ptr(cu), _ptr(mask), _ptr(conf), _ptr(ba)))
        return dict(counts_aff=ca, counts_unaff=cu, risky_mask=mask, conf=conf, ba=ba)

    def set_eval_function(self, eval_function):
        """Which of evaluate_model's functions (model.c:462-479) ranks the models; default BA."""
        self._ck(self.lib.hpgv_epi_set_eval_function(self.h, int(eval_function)))

    def confusion(self, order, combs, risky_mask, eval_subset=SUBSET_TRAINING):
        """confusion_matrix (model.c:337-460) for risky cells given by the caller: risky_mask [n, F]."""
        combs = np.ascontiguousarray(combs, dtype=np.int32).reshape(-1, order)
        n, F = combs.shape[0], self.num_folds
        mask = np.ascontiguousarray(risky_mask, dtype=np.uint32).reshape(n, F)
        conf = np.zeros((n, F, 4), np.uint32)
        val = np.zeros((n, F), np.float64)
        self._ck(self.lib.hpgv_epi_confusion(self.h, order, eval_subset, n, _ptr(combs), _ptr(mask), _ptr(conf), _ptr(val)))
        return conf, val

    def high_risk(self, counts_aff, counts_unaff, num_affected, num_unaffected):
        """The device's high-risk rule (mdr.c:45-75) on explicit count pairs."""
        ca = np.ascontiguousarray(counts_aff, dtype=np.int32).ravel()
        cu = np.ascontiguousarray(counts_unaff, dtype=np.int32).ravel()
        flags = np.zeros(ca.size, np.int32)
        self._ck(self.lib.hpgv_epi_high_risk(self.h, _ptr(ca), _ptr(cu), ca.size, int(num_affected), int(num_unaffected), _ptr(flags)))
        return flags.astype(bool)

    def evaluate(self, conf, eval_function=1):
        """evaluate_model (model.c:462-479) on the device: conf [n, 4] = {TP, FN, FP, TN}."""
        m = np.ascontiguousarray(conf, dtype=np.uint32).reshape(-1, 4)
        out = np.zeros(m.shape[0], np.float64)
        self._ck(self.lib.hpgv_epi_evaluate(self.h, int(eval_function), m.shape[0], _ptr(m), _ptr(out)))
        return out

    def unpack_masks(self, variant):
        _, a, u = self.dims()
        s_pad = 16 * ((a + 15) // 16) + 16 * ((u + 15) // 16)
        out = np.zeros((3, s_pad), np.uint8)
        self._ck(self.lib.hpgv_epi_unpack_masks(self.h, int(variant), _ptr(out)))
        return out

    def last_search_ms(self):
        ms, grid = C.c_float(), C.c_int()
        self._ck(self.lib.hpgv_epi_last_search_ms(self.h, C.byref(ms), C.byref(grid)))
        return ms.value, grid.value

    def search_times(self, n):
        """Device durations (ms) of the last n <= 32 search launches, oldest first; one synchronisation."""
        buf = (C.c_float * max(1, n))()
        got = self.lib.hpgv_epi_search_times(self.h, n, buf)
        if got < 0:
            self._ck(got)
        return [buf[i] for i in range(got)]

    def debug_counters(self, n=8 + 2 * 148):
        out = np.zeros(n, np.uint64)
        got = self.lib.hpgv_epi_debug_counters(self.h, _ptr(out), n)
        if got < 0:
            self._ck(got)
        return out[:got]

    def pipe_peak(self, kind, iters=2000):
        v = C.c_double()
        self._ck(self.lib.hpgv_epi_pipe_peak(self.h, kind, iters, C.byref(v)))
        return v.value


def k_folds(num_affected, num_unaffected, num_folds, seed):
    """Stratified folds with the reference's algorithm and an explicit seed (host only)."""
    lib = _lib.load()
    fos = np.zeros(num_affected + num_unaffected, np.int32)
    sizes = np.zeros(3 * num_folds, np.uint32)
    rc = lib.hpgv_epi_k_folds(num_affected, num_unaffected, num_folds, seed, _ptr(fos), _ptr(sizes))
    if rc != 0:
        raise HpgvError(rc, "k_folds")
    return fos, sizes.reshape(num_folds, 3)


def num_combinations(num_variants, order):
    return int(_lib.load().hpgv_epi_num_combinations(num_variants, order))
