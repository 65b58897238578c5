#!/bin/bash
# round-1 (session 9): 20-warp tri kernel and stagger modes, A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload"
run() { name=$1; shift; env "$@" timeout 200 $B c2 > gpurun_out/bench_c2_l_$name.json 2>/dev/null; echo "$name rc=$?"; cat gpurun_out/bench_c2_l_$name.json | python tools/bench_short.py; }
run w20_s1 HPGV_STAGGER=1
run w20_s2 HPGV_STAGGER=2
run w20_s0 HPGV_STAGGER=0
run w16_s1 HPGV_TRI_WARPS=16 HPGV_STAGGER=1
run w16_s2 HPGV_TRI_WARPS=16 HPGV_STAGGER=2
run w20_s2_ns2 HPGV_STAGGER=2 HPGV_STAGES=2
