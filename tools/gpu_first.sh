#!/bin/bash
# first GPU contact: smoke, pipe peaks, parity tests, a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.txt
tail -5 gpurun_out/smoke.txt
timeout 120 python - <<'PY' > gpurun_out/pipe_peak.txt 2>&1
import hpg_variant_b200 as h
e = h.EpistasisEngine(0)
for kind, name in ((0, "POPC"), (1, "LOP3"), (2, "MIX 2LOP3+1POPC")):
    for it in (500, 2000):
        v = e.pipe_peak(kind, it)
        print(f"{name:18s} iters={it:5d}  {v/1e12:8.3f} Tops/s  = {v/148/1.965e9:7.2f} ops/clk/SM @1965MHz")
PY
cat gpurun_out/pipe_peak.txt
timeout 500 python -m pytest tests -m gpu -x -q --timeout 60 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.txt
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
