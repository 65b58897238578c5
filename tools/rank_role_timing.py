"""One GPU plays each rank of an N-rank weak-scaled c2 step in turn (pack + search of that rank's pair range) and times the
segment with CUDA events next to the library's own search-kernel time: where a rank's time outside the search kernel goes.
Run it under `ncu --metrics gpu__time_duration.sum` for the per-kernel list.  usage: rank_role_timing.py [N] [reps]"""
import sys
import numpy as np
import torch
import bench
import hpg_variant_b200 as h
from hpg_variant_b200 import sharding, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
import argparse
w = bench.workload(argparse.Namespace(gpus=N, steps=1, warmup=1, impl="ours", workload="c2", snps=0, no_cpu_baseline=True, strong=False), N)
nv, A, U, F = w["nv"], w["A"], w["U"], w["folds"]
g = torch.empty((nv, A + U), dtype=torch.uint8).pin_memory()
synth.make_dataset(nv, A, U, w["seed"], order=2, out=g.numpy())
fos, _ = h.k_folds(A, U, F, bench.FOLD_SEED)
eng = h.EpistasisEngine(0)
stream = torch.cuda.current_stream()
eng.set_stream(stream.cuda_stream)
d_raw = g.cuda()
d_out = torch.empty(F * bench.RANK_SIZE * 40, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
total = h.num_combinations(nv, 2)
for r in range(N):
    first, last = sharding.shard_range(total, r, N)
    seg, pk = [], []
    for k in range(reps):
        flush.fill_(k)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(stream)
        eng.load_dataset_device(d_raw.data_ptr(), nv, A, U)
        eng.set_folds(F, fos)
        e[1].record(stream)
        eng.search_device(2, h.SUBSET_TRAINING, bench.RANK_SIZE, first, last, d_out.data_ptr())
        e[2].record(stream)
        torch.cuda.synchronize()
        pk.append(e[0].elapsed_time(e[1]))
        seg.append(e[1].elapsed_time(e[2]))
    km = eng.search_times(reps - 2)
    print(f"role {r}/{N} nv={nv} pairs={last - first}: pack segment {np.median(pk[2:]):.3f} ms, reset+search+list-merge segment {np.median(seg[2:]):.3f} ms, "
          f"search kernel {np.mean(km):.3f} ms -> outside the kernel {np.median(pk[2:]) + np.median(seg[2:]) - np.mean(km):.3f} ms", flush=True)
