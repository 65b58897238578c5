#!/bin/bash
# ncu: launch list of one bench command + full capture of the search kernel
mkdir -p gpurun_out
WL=${1:-c2}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$WL.csv \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$WL.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search -s 3 -c 1 -f -o gpurun_out/prof_$WL \
    python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$WL.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
