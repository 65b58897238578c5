#!/bin/bash
# round-1 (session 9): tri layout with the derived genotype row (marginals), A/B + ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload"
run() { name=$1; shift; env "$@" timeout 200 $B c2 > gpurun_out/bench_c2_m_$name.json 2>gpurun_out/bench_c2_m_$name.err; echo "$name rc=$?"; tail -2 gpurun_out/bench_c2_m_$name.err; cat gpurun_out/bench_c2_m_$name.json | python tools/bench_short.py; }
run derive HPGV_TRI_DERIVE=1
run noderive HPGV_TRI_DERIVE=0
run derive_s0 HPGV_STAGGER=0
run derive_w16 HPGV_TRI_WARPS=16
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search -s 3 -c 1 -f -o gpurun_out/prof_c2_m \
    python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c2.log 2>&1
echo "full capture rc=$?"
