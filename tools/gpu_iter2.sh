#!/bin/bash
mkdir -p gpurun_out
run() { WL=$1; shift
  timeout 900 python bench.py --workload $WL --no-cpu-baseline "$@" > gpurun_out/bench_$WL.json 2> gpurun_out/bench_$WL.err; echo "bench $WL rc=$?"
  tail -2 gpurun_out/bench_$WL.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$WL.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$WL value %.3e e2e %.3e ms/step %.2f kernel_ms %.2f frac %.3f share %.2f layout %s clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["kernel_share_of_step"], d["config"]["layout"], d["clocks"]))
except Exception as e:
    print("no bench line:", e)
PY
}
run c3 --steps 3 --warmup 3
run c5 --steps 3 --warmup 3
run c4 --steps 3 --warmup 3 --snps 1200
