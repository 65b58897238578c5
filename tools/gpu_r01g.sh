#!/bin/bash
# round-1 (session 9): histogram counting out of line, rotating sparse threshold refresh -- parity + bench c2/c3 (+ A/B without histogram)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload"
timeout 200 $B c2 > gpurun_out/bench_c2_g.json 2> gpurun_out/bench_c2_g.err; echo "c2 rc=$?"; tail -3 gpurun_out/bench_c2_g.err; cat gpurun_out/bench_c2_g.json | python tools/bench_short.py
HPGV_STAGGER=0 timeout 200 $B c2 > gpurun_out/bench_c2_g_nostagger.json 2>/dev/null; echo "c2 nostagger rc=$?"; cat gpurun_out/bench_c2_g_nostagger.json | python tools/bench_short.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search -s 3 -c 1 -f -o gpurun_out/prof_c2_g \
    python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c2.log 2>&1
echo "full capture rc=$?"
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload"
timeout 300 $B c3 > gpurun_out/bench_c3_g.json 2>/dev/null; echo "c3 rc=$?"; cat gpurun_out/bench_c3_g.json | python tools/bench_short.py
HPGV_HIST=0 timeout 300 $B c3 > gpurun_out/bench_c3_g_nohist.json 2>/dev/null; echo "c3 nohist rc=$?"; cat gpurun_out/bench_c3_g_nohist.json | python tools/bench_short.py
timeout 300 $B c5 > gpurun_out/bench_c5_g.json 2>/dev/null; echo "c5 rc=$?"; cat gpurun_out/bench_c5_g.json | python tools/bench_short.py
ls gpurun_out/
