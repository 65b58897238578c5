#!/bin/bash
# iteration checkpoint: GPU parity tests, pipe micro-benchmark, device benches of the four synthetic workloads, ncu capture of c2
mkdir -p gpurun_out
TAG=${1:-it}
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_$TAG.txt
[ -x tools/pipebench ] && timeout 120 tools/pipebench > gpurun_out/pipebench.txt 2>&1
run() { WL=$1; shift
  timeout 900 python bench.py --workload $WL --no-cpu-baseline "$@" > gpurun_out/bench_${WL}_$TAG.json 2> gpurun_out/bench_${WL}_$TAG.err; echo "bench $WL rc=$?"
  tail -2 gpurun_out/bench_${WL}_$TAG.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${WL}_$TAG.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$WL value %.3e e2e %.3e ms/step %.2f kernel_ms %.2f frac %.3f share %.2f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["kernel_share_of_step"], d["clocks"]))
except Exception as e:
    print("no bench line:", e)
PY
}
run c2 --steps 10 --warmup 3
run c3 --steps 2 --warmup 3
run c5 --steps 2 --warmup 3
run c4 --steps 2 --warmup 3 --snps 1200
if [ "$2" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search -s 3 -c 1 -f -o gpurun_out/prof_c2_$TAG \
    python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_c2_$TAG.log 2>&1
echo "full capture c2 rc=$?"
fi
ls gpurun_out/ | head -50
