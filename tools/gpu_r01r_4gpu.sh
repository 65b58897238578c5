#!/bin/bash
# round-1 (session 9): where a multi-GPU step spends its time (N=4, N=2; phases in the bench line)
mkdir -p gpurun_out
for N in 4 2; do
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N"
timeout 300 $T bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_c2_r_n$N.json 2> gpurun_out/bench_c2_r_n$N.err; echo "n$N rc=$?"; grep -v "OMP_NUM\|\*\*\*" gpurun_out/bench_c2_r_n$N.err | tail -3; grep '^{' gpurun_out/bench_c2_r_n$N.json | python tools/bench_short.py
grep '^{' gpurun_out/bench_c2_r_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps(d.get('phases_ms'))); print(d['clocks'])"
done
timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_r_n1.json 2>/dev/null; echo "n1 rc=$?"; grep '^{' gpurun_out/bench_c2_r_n1.json | python tools/bench_short.py
