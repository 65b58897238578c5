#!/bin/bash
# round-1 (session 9): packer offset tables -- parity, bench, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_q.json 2>gpurun_out/bench_c2_q.err; echo "c2 rc=$?"; tail -2 gpurun_out/bench_c2_q.err; cat gpurun_out/bench_c2_q.json | python tools/bench_short.py
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c2_q.csv \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_c2.log 2>&1
echo "launch list rc=$?"; grep -E "pack_rows|merge_kernel" gpurun_out/launches_c2_q.csv | tail -4 | cut -d, -f1,5,15-
timeout 300 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_q.json 2>/dev/null; echo "c3 rc=$?"; cat gpurun_out/bench_c3_q.json | python tools/bench_short.py
