#!/bin/bash
# round-1 (session 9): score-histogram thresholds -- parity, A/B bench on c2, ncu full capture
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.txt
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
B="python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline"
timeout 200 $B > gpurun_out/bench_c2_hist.json 2> gpurun_out/bench_c2_hist.err; echo "hist rc=$?"; tail -3 gpurun_out/bench_c2_hist.err; cat gpurun_out/bench_c2_hist.json | python tools/bench_short.py
HPGV_HIST=0 timeout 200 $B > gpurun_out/bench_c2_nohist.json 2>/dev/null; echo "nohist rc=$?"; cat gpurun_out/bench_c2_nohist.json | python tools/bench_short.py
HPGV_NO_TRI=1 timeout 200 $B > gpurun_out/bench_c2_bw4_hist.json 2>/dev/null; echo "bw4+hist rc=$?"; cat gpurun_out/bench_c2_bw4_hist.json | python tools/bench_short.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search -s 3 -c 1 -f -o gpurun_out/prof_c2_hist \
    python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c2.log 2>&1
echo "full capture rc=$?"
timeout 300 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_hist.json 2>/dev/null; echo "c3 rc=$?"; cat gpurun_out/bench_c3_hist.json | python tools/bench_short.py
ls -la gpurun_out/
