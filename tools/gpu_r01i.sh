#!/bin/bash
# round-1 (session 9): staged packer, suspended barrier waits, edge-row range checks -- parity + bench + launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload"
timeout 200 $B c2 > gpurun_out/bench_c2_j.json 2> gpurun_out/bench_c2_j.err; echo "c2 rc=$?"; tail -3 gpurun_out/bench_c2_j.err; cat gpurun_out/bench_c2_j.json | python tools/bench_short.py
HPGV_PACK_WARP=1 timeout 200 $B c2 > gpurun_out/bench_c2_j_oldpack.json 2>/dev/null; echo "c2 oldpack rc=$?"; cat gpurun_out/bench_c2_j_oldpack.json | python tools/bench_short.py
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c2.csv \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_c2.log 2>&1
echo "launch list rc=$?"
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload"
timeout 300 $B c3 > gpurun_out/bench_c3_j.json 2>/dev/null; echo "c3 rc=$?"; cat gpurun_out/bench_c3_j.json | python tools/bench_short.py
ls gpurun_out/ | head -30
