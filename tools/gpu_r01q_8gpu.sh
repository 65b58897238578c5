#!/bin/bash
# round-1 (session 9): the bench contract at N=8 and N=4 (torchrun, NCCL, weak scaling)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 8 4; do
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
timeout 300 $T bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_c2_n$N.json 2> gpurun_out/bench_c2_n$N.err; echo "n$N rc=$?"; grep -v "OMP_NUM\|\*\*\*" gpurun_out/bench_c2_n$N.err | tail -3; grep '^{' gpurun_out/bench_c2_n$N.json | python tools/bench_short.py
done
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_n1b.json 2>/dev/null; echo "n1 rc=$?"; grep '^{' gpurun_out/bench_c2_n1b.json | python tools/bench_short.py
