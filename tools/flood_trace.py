"""Which tuples reach the lists in the slow part of the pair range (development trace, HPGV_DEBUG_COUNTERS=1)."""
import os, sys, collections, numpy as np
os.environ["HPGV_DEBUG_COUNTERS"] = "1"
import hpg_variant_b200 as h
from hpg_variant_b200 import synth
nv, A, F = 20000, 1000, 10
g = synth.make_dataset(nv, A, A, 1002)
fos, _ = h.k_folds(A, A, F, 1)
eng = h.EpistasisEngine(0)
eng.load_dataset(g, A, A); eng.set_folds(F, fos)
total = h.num_combinations(nv, 2)
lo, hi = total * 13 // 16, total * 14 // 16
res = eng.search(2, h.SUBSET_TRAINING, 50, lo, hi)
c = eng.debug_counters(8 + 2 * 1024 + 8 + 2 * (1 << 20))
n = int(c[8 + 2 * 1024]); n = min(n, 1 << 20)
tr = c[8 + 2 * 1024 + 8:][: 2 * n].reshape(n, 2)
i = (tr[:, 0] >> np.uint64(32)).astype(np.int64); j = (tr[:, 0] & np.uint64(0xffffffff)).astype(np.int64)
f = (tr[:, 1] >> np.uint64(56)).astype(np.int64); t = ((tr[:, 1] >> np.uint64(28)) & np.uint64(0xfffffff)).astype(np.int64) // 900; b = (tr[:, 1] & np.uint64(0xfffffff)).astype(np.int64) // 900
print("offered", n, "final bound t of fold 0:", res["conf"][0, -1], "ba", res["accuracy"][0, -1])
print("by fold", np.bincount(f, minlength=F))
ci = collections.Counter(i.tolist()).most_common(8); cj = collections.Counter(j.tolist()).most_common(8)
print("most frequent i:", ci); print("most frequent j:", cj)
sel = f == 0
print("fold 0: t quantiles", np.quantile(t[sel], [0, .1, .5, .9, 1]), "bound-at-offer quantiles", np.quantile(b[sel], [0, .1, .5, .9, 1]))
for k in range(0, n, max(1, n // 12)):
    print(k, i[k], j[k], f[k], t[k], b[k])
