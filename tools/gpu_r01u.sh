#!/bin/bash
# round-1 (session 9): ncu capture of the order-2 search over the last quarter of the pair range (nv = 20000)
mkdir -p gpurun_out
cat > gpurun_out/q3.py <<'PY'
import numpy as np
import hpg_variant_b200 as h
from hpg_variant_b200 import synth
nv, A, F = 20000, 1000, 10
g = synth.make_dataset(nv, A, A, 1002)
fos, _ = h.k_folds(A, A, F, 1)
eng = h.EpistasisEngine(0)
eng.load_dataset(g, A, A); eng.set_folds(F, fos)
total = h.num_combinations(nv, 2)
for rep in range(3):
    r = eng.search(2, h.SUBSET_TRAINING, 50, total * 3 // 4, total)
print(eng.last_search_ms(), r["snp"][0, :5, :2].tolist(), r["accuracy"][0, :5].tolist(), r["accuracy"][0, 45:].tolist())
for rep in range(2):
    r = eng.search(2, h.SUBSET_TRAINING, 50, 0, total // 4)
print(eng.last_search_ms(), r["snp"][0, :5, :2].tolist(), r["accuracy"][0, :5].tolist(), r["accuracy"][0, 45:].tolist())
PY
PYTHONPATH=. timeout 600 ncu --set full --clock-control none --import-source on -k regex:search2 -s 2 -c 1 -f -o gpurun_out/prof_q3 python gpurun_out/q3.py > gpurun_out/q3.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/q3.log
