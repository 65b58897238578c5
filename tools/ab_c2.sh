#!/bin/bash
# A/B of library builds / switches on c2 and on the quarters of the 4-GPU weak data set (the slow rank of round 1).
# usage: tools/ab_c2.sh "<label>:<env assignments>" ...   e.g.  "r01:HPGV_EPI_LIB=$PWD/hpg_variant_b200/libhpgv_epi_r01.so" "cur:"
mkdir -p gpurun_out
for spec in "$@"; do
  label=${spec%%:*}; envs=${spec#*:}
  for rep in 1 2; do
    env $envs python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-parity-check 2>/dev/null | python tools/bench_short.py | sed "s/^/$label c2 /"
  done
  env $envs PYTHONPATH=$PWD python tools/subrange_timing.py $label
done
