#!/bin/bash
# round-1 (session 9): per-unit descriptor list -- parity, N=1 A/B, sub-range timing (the last quarter of the pairs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload"
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 300 $B $wl > gpurun_out/bench_s_$name.json 2>gpurun_out/bench_s_$name.err; echo "$name rc=$?"; tail -2 gpurun_out/bench_s_$name.err; cat gpurun_out/bench_s_$name.json | python tools/bench_short.py; }
run c2_desc c2 HPGV_UNIT_DESC=1
run c2_nodesc c2 HPGV_UNIT_DESC=0
python - <<'PY'
import numpy as np, torch, time
import hpg_variant_b200 as h
from hpg_variant_b200 import synth
nv, A, F = 20000, 1000, 10
g = synth.make_dataset(nv, A, A, 1002)
fos, _ = h.k_folds(A, A, F, 1)
eng = h.EpistasisEngine(0)
eng.load_dataset(g, A, A); eng.set_folds(F, fos)
total = h.num_combinations(nv, 2)
for q in range(4):
    lo, hi = total * q // 4, total * (q + 1) // 4
    for rep in range(3):
        eng.search(2, h.SUBSET_TRAINING, 50, lo, hi)
    print("quarter", q, "search kernel ms", eng.last_search_ms())
PY
