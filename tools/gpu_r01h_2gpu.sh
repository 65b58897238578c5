#!/bin/bash
# round-1 (session 9): the bench contract at N=2 (torchrun, NCCL all-gather of the per-rank lists, weak scaling) + reference arm plumbing
mkdir -p gpurun_out
nvidia-smi -L
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err; echo "n2 rc=$?"; tail -5 gpurun_out/bench_c2_n2.err; grep '^{' gpurun_out/bench_c2_n2.json | python tools/bench_short.py
timeout 300 $T bench.py --gpus 2 --steps 10 --warmup 3 --strong > gpurun_out/bench_c2_n2_strong.json 2> gpurun_out/bench_c2_n2_strong.err; echo "n2 strong rc=$?"; grep '^{' gpurun_out/bench_c2_n2_strong.json | python tools/bench_short.py
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_n1.json 2>/dev/null; echo "n1 rc=$?"; grep '^{' gpurun_out/bench_c2_n1.json | python tools/bench_short.py
timeout 300 $T bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?"; tail -3 gpurun_out/bench_ref_n2.err; cat gpurun_out/bench_ref_n2.json | cut -c1-400
timeout 300 python -m pytest tests/test_sharding_gloo.py -x -q 2>&1 | tail -2
