"""Small searches for compute-sanitizer (racecheck / memcheck): a many-unit order-2 search (score histogram, first-unit wait,
list locks), an order-2 search with global-memory lists, and the two order-3 kernels.  Results are checked against the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hpg_variant_b200 as h
from hpg_variant_b200 import synth
import oracle_lib
orc = oracle_lib.Checker("oracle")
eng = h.EpistasisEngine(0)
def case(tag, order, nv, A, U, F, rank, subset=1, env=None):
    for k, v in (env or {}).items():
        os.environ[k] = v
    g = synth.make_dataset(nv, A, U, seed=nv + order, order=order, missing=0.01, planted=2)
    fos, _ = h.k_folds(A, U, F, seed=5)
    eng.load_dataset(g, A, U); eng.set_folds(F, fos)
    got = eng.search(order, subset, rank)
    want, _ = orc.search(g, A, U, order, fos, subset, rank, threads=8, num_folds=F)
    ok = np.array_equal(got["snp"][..., :order], want["snp"][..., :order]) and np.array_equal(got["accuracy"], want["ba"], equal_nan=True)
    print(f"{tag}: order {order}, {nv} SNPs x {A}+{U} samples, {F} folds, rank {rank}: {'matches the oracle' if ok else 'MISMATCH'}", flush=True)
    for k in (env or {}):
        os.environ.pop(k, None)
    return ok
ok = True
ok &= case("order 2, tri layout, ~500 units (histogram bound, first-unit wait, array lists)", 2, 640, 100, 100, 2, 50)
ok &= case("order 2, 8-word single-block layout, marginals", 2, 200, 720, 720, 3, 30)
ok &= case("order 2, multi-block 16-bit counters, unbalanced, lists in global memory", 2, 120, 1500, 1100, 4, 700)
ok &= case("order 3, resident tiles (search3v2)", 3, 40, 2000, 2000, 5, 30)
ok &= case("order 3, plain kernel", 3, 30, 700, 700, 5, 30, env={"HPGV_SEARCH3_V2": "0"})
sys.exit(0 if ok else 1)
