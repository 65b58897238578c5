import os, sys, numpy as np
import hpg_variant_b200 as h
from hpg_variant_b200 import synth
nv, A, F = 20000, 1000, 10
g = synth.make_dataset(nv, A, A, 1002)
fos, _ = h.k_folds(A, A, F, 1)
eng = h.EpistasisEngine(0)
eng.load_dataset(g, A, A); eng.set_folds(F, fos)
total = h.num_combinations(nv, 2)
def t(lo, hi):
    for rep in range(3):
        eng.search(2, h.SUBSET_TRAINING, 50, lo, hi)
    return eng.last_search_ms()[0]
tag = sys.argv[1]
print(tag, "quarters", [round(t(total * q // 4, total * (q + 1) // 4), 3) for q in range(4)])
print(tag, "sixteenths of the last quarter (same pair count each)", [round(t(total * (12 + q) // 16, total * (13 + q) // 16), 3) for q in range(4)])
