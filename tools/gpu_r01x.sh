#!/bin/bash
# round-1 (session 9): A/B of three builds (previous commit, current, current without the root filter) on c2, c3, c5 and the slow quarter
mkdir -p gpurun_out
for V in prev norf cur; do
  if [ $V = cur ]; then unset HPGV_EPI_LIB; else export HPGV_EPI_LIB=$PWD/hpg_variant_b200/libhpgv_epi_$V.so; fi
  timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x_c2_$V.json 2>/dev/null; echo "$V c2:"; cat gpurun_out/bench_x_c2_$V.json | python tools/bench_short.py
  timeout 200 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x_c3_$V.json 2>/dev/null; echo "$V c3:"; cat gpurun_out/bench_x_c3_$V.json | python tools/bench_short.py
  timeout 200 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x_c5_$V.json 2>/dev/null; echo "$V c5:"; cat gpurun_out/bench_x_c5_$V.json | python tools/bench_short.py
  PYTHONPATH=. python tools/subrange_timing.py $V 2>&1 | head -1
done
