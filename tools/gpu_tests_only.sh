#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python tools/bench_short.py
