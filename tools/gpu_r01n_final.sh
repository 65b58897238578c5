#!/bin/bash
# round-1 checkpoint (session 9): smoke, GPU parity, default bench (with CPU baseline), reference arm, c3/c4/c5 lines, launch list + full capture (c2)
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.txt
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json | python tools/bench_short.py
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_reference.json
for WL in c3 c5; do
timeout 400 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$WL.json 2> gpurun_out/bench_$WL.err; echo "bench $WL rc=$?"; cat gpurun_out/bench_$WL.json | python tools/bench_short.py
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c2.csv \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_c2.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search -s 3 -c 1 -f -o gpurun_out/prof_c2 \
    python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c2.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/ | tail -15
