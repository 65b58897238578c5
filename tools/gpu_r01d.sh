#!/bin/bash
# round-1 re-entry checkpoint: smoke, GPU parity tests, default bench (with CPU baseline), ncu launch list + full capture (c2)
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.txt
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
for WL in c3 c4 c5; do
timeout 300 python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$WL.json 2> gpurun_out/bench_$WL.err; echo "bench $WL rc=$?"; cat gpurun_out/bench_$WL.json
done
WL=c2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$WL.csv \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$WL.log 2>&1
echo "launch list $WL rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search -s 3 -c 1 -f -o gpurun_out/prof_$WL \
    python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$WL.log 2>&1
echo "full capture $WL rc=$?"
ls -la gpurun_out/
