#!/bin/bash
# round-1 (session 9): derived genotype row for the single-block layouts (c3), A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload"
run() { name=$1; wl=$2; shift; shift; env "$@" timeout 300 $B $wl > gpurun_out/bench_o_$name.json 2>gpurun_out/bench_o_$name.err; echo "$name rc=$?"; tail -2 gpurun_out/bench_o_$name.err; cat gpurun_out/bench_o_$name.json | python tools/bench_short.py; }
run c3_derive c3 HPGV_TRI_DERIVE=1

run c3_w16 c3 HPGV_TRI_WARPS=16
run c2 c2 HPGV_TRI_DERIVE=1
