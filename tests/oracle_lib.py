"""ctypes bindings for the CPU checkers under oracle/ (TEST INFRASTRUCTURE).

`Checker("oracle")` loads oracle/liboracle.so (plain-C restatement, symbols
oracle_*); `Checker("ref")` loads oracle/_ref/libhpgref.so (the reference's
own sources compiled by oracle/Makefile, symbols ref_*).  Both expose the same
driver entry points (oracle/epi_driver.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import
this module; nothing under hpg_variant_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

MODEL_DTYPE = np.dtype(
    [("ba", "<f8"), ("snp", "<i4", (3,)), ("risky_mask", "<u4"), ("conf", "<u4", (4,))]
)
assert MODEL_DTYPE.itemsize == 40


def _clean_env():
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    return env


def build_oracle(target="oracle"):
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, target], check=True, env=_clean_env())


def lib_path(kind):
    return os.path.join(ORACLE_DIR, "liboracle.so") if kind == "oracle" else os.path.join(ORACLE_DIR, "_ref", "libhpgref.so")


def available(kind):
    return os.path.exists(lib_path(kind))


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class Checker:
    def __init__(self, kind="oracle"):
        assert kind in ("oracle", "ref")
        self.kind = kind
        path = lib_path(kind)
        if kind == "oracle" and not os.path.exists(path):
            build_oracle("oracle")
        self.lib = C.CDLL(path)
        self.pfx = "oracle_" if kind == "oracle" else "ref_"
        f = self._fn("eval")
        f.restype = C.c_int
        f = self._fn("search")
        f.restype = C.c_int
        self._fn("num_combinations").restype = C.c_uint64
        self._fn("evaluate").restype = C.c_double
        self._fn("enumerate_blocked").restype = C.c_int64

    def _fn(self, name):
        return getattr(self.lib, self.pfx + name)

    # ---- per-combination dump -------------------------------------------------
    def eval(self, geno, A, U, order, fold_of_sample, subset, combs):
        """geno: uint8 [nv, A+U]; combs: int32 [n, order]; subset: 1=training, 0=testing.
        Returns dict of counts_aff/unaff [n,F,C] (TRAINING counts), risky_mask [n,F], conf [n,F,4], ba [n,F]."""
        geno = np.ascontiguousarray(geno, dtype=np.uint8)
        fos = np.ascontiguousarray(fold_of_sample, dtype=np.int32)
        combs = np.ascontiguousarray(combs, dtype=np.int32).reshape(-1, order)
        nv = geno.shape[0]
        F = int(fos.max()) + 1 if fos.size else 1
        n = combs.shape[0]
        Cc = 3 ** order
        ca = np.zeros((n, F, Cc), np.int32)
        cu = np.zeros((n, F, Cc), np.int32)
        mask = np.zeros((n, F), np.uint32)
        conf = np.zeros((n, F, 4), np.uint32)
        ba = np.zeros((n, F), np.float64)
        rc = self._fn("eval")(
            _p(geno, C.c_uint8), C.c_int(nv), C.c_int(A), C.c_int(U), C.c_int(order), C.c_int(F),
            _p(fos, C.c_int32), C.c_int(subset), C.c_int64(n), _p(combs, C.c_int32),
            _p(ca, C.c_int32), _p(cu, C.c_int32), _p(mask, C.c_uint32), _p(conf, C.c_uint32), _p(ba, C.c_double))
        if rc != 0:
            raise RuntimeError(f"{self.pfx}eval failed: {rc}")
        return dict(counts_aff=ca, counts_unaff=cu, risky_mask=mask, conf=conf, ba=ba)

    # ---- exhaustive canonical top-N ---------------------------------------------
    def search(self, geno, A, U, order, fold_of_sample, subset, topn, first=0, last=None, threads=1, num_folds=None):
        geno = np.ascontiguousarray(geno, dtype=np.uint8)
        fos = np.ascontiguousarray(fold_of_sample, dtype=np.int32)
        nv = geno.shape[0]
        F = num_folds if num_folds is not None else int(fos.max()) + 1
        total = int(self._fn("num_combinations")(C.c_int(nv), C.c_int(order)))
        if last is None:
            last = total
        out = np.zeros((F, topn), MODEL_DTYPE)
        n_out = np.zeros(F, np.int32)
        rc = self._fn("search")(
            _p(geno, C.c_uint8), C.c_int(nv), C.c_int(A), C.c_int(U), C.c_int(order), C.c_int(F),
            _p(fos, C.c_int32), C.c_int(subset), C.c_int(topn), C.c_uint64(first), C.c_uint64(last),
            C.c_int(threads), out.ctypes.data_as(C.c_void_p), _p(n_out, C.c_int32))
        if rc != 0:
            raise RuntimeError(f"{self.pfx}search failed: {rc}")
        return out, n_out

    def set_eval_function(self, fn):
        """evaluate_model's function code for eval()/search() (1 = BA, the default; 5 = documented CA)."""
        self._fn("set_eval_function")(C.c_int(fn))

    # ---- leaf helpers -----------------------------------------------------------
    def k_folds(self, A, U, k, seed):
        fos = np.full(A + U, -1, np.int32)
        sizes = np.zeros(3 * k, np.uint32)
        self._fn("k_folds")(C.c_int(A), C.c_int(U), C.c_int(k), C.c_long(seed), _p(fos, C.c_int32), _p(sizes, C.c_uint32))
        return fos, sizes.reshape(k, 3)

    def fold_masks(self, A, U, F, fold_of_sample):
        fos = np.ascontiguousarray(fold_of_sample, dtype=np.int32)
        s_pad = 16 * ((A + 15) // 16) + 16 * ((U + 15) // 16)
        out = np.zeros((F, s_pad), np.uint8)
        rc = self._fn("fold_masks")(C.c_int(A), C.c_int(U), C.c_int(F), _p(fos, C.c_int32), _p(out, C.c_uint8))
        if rc != 0:
            raise RuntimeError(f"fold_masks failed: {rc}")
        return out

    def high_risk(self, ca, cu, A, U):
        ca = np.ascontiguousarray(ca, dtype=np.int32)
        cu = np.ascontiguousarray(cu, dtype=np.int32)
        flags = np.zeros(ca.size, np.int32)
        self._fn("high_risk")(_p(ca, C.c_int32), _p(cu, C.c_int32), C.c_int(ca.size), C.c_uint(A), C.c_uint(U), _p(flags, C.c_int32))
        return flags.astype(bool)

    def evaluate(self, conf, function=1):
        m = np.ascontiguousarray(conf, dtype=np.uint32)
        return float(self._fn("evaluate")(_p(m, C.c_uint32), C.c_int(function)))

    def enumerate_blocked(self, nv, order, stride, cap=1 << 22):
        out = np.zeros((cap, order), np.int32)
        n = int(self._fn("enumerate_blocked")(C.c_int(nv), C.c_int(order), C.c_int(stride), _p(out, C.c_int32), C.c_int64(cap)))
        return out[: min(n, cap)], n

    def unrank(self, nv, order, idx):
        comb = np.zeros(order, np.int32)
        self._fn("unrank")(C.c_int(nv), C.c_int(order), C.c_uint64(idx), _p(comb, C.c_int32))
        return comb

    # ---- reference only: merge_rankings + epistasis_report on given per-fold rankings ------------
    def merge_rankings(self, order, models, mode):
        """models: structured [F, rank] (MODEL_DTYPE, empty slots snp[0] < 0); mode 0 = CV_C, 1 = CV_A.
        Returns the rows in the order the reference's report lists them: (snp tuple, cv_count, cv_accuracy, risky cells)."""
        assert self.kind == "ref"
        m = np.ascontiguousarray(models)
        F, rank = m.shape
        cap = F * rank
        snp = np.zeros((cap, 3), np.int32)
        cnt = np.zeros(cap, np.int32)
        acc = np.zeros(cap, np.float64)
        nr = np.zeros(cap, np.int32)
        gts = np.zeros((cap, 81), np.uint8)
        f = self._fn("merge_rankings")
        f.restype = C.c_int
        n = f(C.c_int(order), C.c_int(F), C.c_int(rank), m.ctypes.data_as(C.c_void_p), C.c_int(mode), C.c_int(cap),
              _p(snp, C.c_int32), _p(cnt, C.c_int32), _p(acc, C.c_double), _p(nr, C.c_int32), _p(gts, C.c_uint8))
        rows = []
        for r in range(n):
            cells = [int(sum(int(gts[r, q * order + p]) * 3 ** (order - 1 - p) for p in range(order))) for q in range(nr[r])]
            rows.append((tuple(int(x) for x in snp[r, :order]), int(cnt[r]), float(acc[r]), cells))
        return rows

    def report(self, order, models, mode, subset, cv_repetition, max_ranking_size, path):
        assert self.kind == "ref"
        m = np.ascontiguousarray(models)
        F, rank = m.shape
        f = self._fn("report")
        f.restype = C.c_int
        rc = f(C.c_int(order), C.c_int(F), C.c_int(rank), m.ctypes.data_as(C.c_void_p), C.c_int(mode), C.c_int(subset),
               C.c_int(cv_repetition), C.c_int(max_ranking_size), str(path).encode())
        if rc != 0:
            raise RuntimeError("ref_report failed")
        with open(path) as fh:
            return fh.read()

    # ---- reference only: its own end-to-end runner ------------------------------
    def run_epistasis(self, dataset, outdir, order, stride, num_folds, reps, rank_size, subset, mode, threads):
        assert self.kind == "ref"
        f = self._fn("run_epistasis")
        f.restype = C.c_int
        return f(dataset.encode(), outdir.encode(), C.c_int(order), C.c_int(stride), C.c_int(num_folds), C.c_int(reps),
                 C.c_int(rank_size), C.c_int(subset), C.c_int(mode), C.c_int(threads))
