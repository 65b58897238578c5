#!/bin/bash
# quick iteration: GPU parity tests + short benches (no CPU baseline)
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
for WL in ${@:-c2}; do
  timeout 600 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$WL.json 2> gpurun_out/bench_$WL.err; echo "bench $WL rc=$?"
  tail -2 gpurun_out/bench_$WL.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$WL.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$WL value %.3e e2e %.3e ms/step %.2f kernel_ms %.2f frac %.3f share %.2f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["kernel_share_of_step"], d["clocks"]))
except Exception as e:
    print("no bench line:", e)
PY
done
