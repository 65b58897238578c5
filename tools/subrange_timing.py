"""Search-kernel time of sub-ranges of the pair space, with the development counters of the kernel
(HPGV_DEBUG_COUNTERS=1): where a multi-GPU rank's time goes.  usage: subrange_timing.py <label> [nv]"""
import os, sys, numpy as np
os.environ["HPGV_DEBUG_COUNTERS"] = "1"
import hpg_variant_b200 as h
from hpg_variant_b200 import synth
label = sys.argv[1]
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
A, F = 1000, 10
g = synth.make_dataset(nv, A, A, 1002)
fos, _ = h.k_folds(A, A, F, 1)
eng = h.EpistasisEngine(0)
eng.load_dataset(g, A, A); eng.set_folds(F, fos)
total = h.num_combinations(nv, 2)
def t(lo, hi, tag):
    ms = []
    for rep in range(4):
        eng.search(2, h.SUBSET_TRAINING, 50, lo, hi)
        ms.append(eng.last_search_ms()[0])
    c = eng.debug_counters()
    clk = c[8:].reshape(-1, 2).astype(np.int64)
    dur = (clk[:, 1] - clk[:, 0]) / 1.965e6
    print(f"{label} {tag:>14s} pairs={hi - lo:>11d} ms={min(ms[1:]):.3f} prefilter_warps={c[0]} offered={c[1]} after_root={c[2]} stored={c[3]} spins={c[4]} "
          f"cta_ms min/med/max={dur.min():.3f}/{np.median(dur):.3f}/{dur.max():.3f} slowest CTAs {np.argsort(-dur)[:4].tolist()} {np.sort(dur)[::-1][:4].round(3).tolist()}")
t(0, total, "all")
for q in range(4):
    t(total * q // 4, total * (q + 1) // 4, f"quarter{q}")
for d in (8, 16, 32, 64):
    t(0, total // d, f"first 1/{d}")
for q in range(4):
    t(total * (12 + q) // 16, total * (13 + q) // 16, f"16th {12 + q}")
