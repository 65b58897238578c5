#!/bin/bash
# round-1 (session 9): three-stage ring A/B + full ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload"
timeout 200 $B c2 > gpurun_out/bench_c2_k.json 2> gpurun_out/bench_c2_k.err; echo "c2 rc=$?"; tail -3 gpurun_out/bench_c2_k.err; cat gpurun_out/bench_c2_k.json | python tools/bench_short.py
HPGV_STAGES=2 timeout 200 $B c2 > gpurun_out/bench_c2_k_ns2.json 2>/dev/null; echo "c2 ns2 rc=$?"; cat gpurun_out/bench_c2_k_ns2.json | python tools/bench_short.py
HPGV_STAGGER=0 timeout 200 $B c2 > gpurun_out/bench_c2_k_nostagger.json 2>/dev/null; echo "c2 nostagger rc=$?"; cat gpurun_out/bench_c2_k_nostagger.json | python tools/bench_short.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search -s 3 -c 1 -f -o gpurun_out/prof_c2_k \
    python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c2.log 2>&1
echo "full capture rc=$?"
ls gpurun_out/ | head -30
