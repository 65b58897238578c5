#!/bin/bash
# round-1 (session 9): published list root (seqlock) -- parity, sub-range timings, c2 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
PYTHONPATH=. python gpurun_out/subrange.py rootfilter
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_v.json 2>gpurun_out/bench_c2_v.err; echo "c2 rc=$?"; tail -2 gpurun_out/bench_c2_v.err; cat gpurun_out/bench_c2_v.json | python tools/bench_short.py
