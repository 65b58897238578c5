#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/s.py <<'PY'
import numpy as np, hpg_variant_b200 as h
from hpg_variant_b200 import synth
eng = h.EpistasisEngine(0)
nv, A, U, F, rank = 48, 120, 136, 5, 20
g = synth.make_dataset(nv, A, U, seed=19, order=2, missing=0.01, planted=1)
fos, _ = h.k_folds(A, U, F, seed=99)
eng.load_dataset(g, A, U); eng.set_folds(F, fos)
try:
    got = eng.search(2, h.SUBSET_TRAINING, rank)
    print(got[0, :3])
except Exception as e:
    print("ERR", e)
PY
timeout 600 env PYTHONPATH=$PWD compute-sanitizer --tool memcheck --print-limit 5 python /tmp/s.py > gpurun_out/sanitize.txt 2>&1
echo rc=$?
grep -v "^=========     Host Frame\|^=========         in \|^=========                in" gpurun_out/sanitize.txt | head -80
