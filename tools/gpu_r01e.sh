#!/bin/bash
# round-1 (session 9): tri layout -- parity, A/B bench on c2 (tri / no stagger / 4-word blocks), ncu full capture of the tri kernel
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.txt
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
B="python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline"
timeout 200 $B > gpurun_out/bench_c2_tri.json 2> gpurun_out/bench_c2_tri.err; echo "tri rc=$?"; cat gpurun_out/bench_c2_tri.json | python tools/bench_short.py
HPGV_STAGGER=0 timeout 200 $B > gpurun_out/bench_c2_tri_nostagger.json 2>/dev/null; echo "tri nostagger rc=$?"; cat gpurun_out/bench_c2_tri_nostagger.json | python tools/bench_short.py
HPGV_NO_TRI=1 timeout 200 $B > gpurun_out/bench_c2_bw4.json 2>/dev/null; echo "bw4 rc=$?"; cat gpurun_out/bench_c2_bw4.json | python tools/bench_short.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search -s 3 -c 1 -f -o gpurun_out/prof_c2_tri \
    python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c2.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
