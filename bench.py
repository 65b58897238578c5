#!/usr/bin/env python
"""bench.py -- throughput of the exhaustive epistasis search (MDR + k-fold CV) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c5]

One "step" = one pass of the hot path over one synthetic dataset: pack the bit planes for the
fold assignment, evaluate every SNP combination for every fold, rank, merge.  Metric (BASELINE.json):
SNP-combinations x folds evaluated per second.

  value      device-resident: raw genotype bytes already in HBM, CUDA-event time of pack+search+merge
             (+ NCCL all-gather + merge of the per-rank top-N when N > 1), max over ranks
  e2e        the same through hpgv_epi_run_host with HOST (pinned) buffers: H2D of the genotype bytes,
             pack, search, merge, D2H of the ranked models -- wall clock between device synchronisations
  roofline   the search kernel alone (CUDA events recorded around its launch by the library) against the
             POPC-pipe peak measured by a micro-benchmark in the same run; HBM figures alongside
  cpu_baseline  the reference's own run_epistasis (oracle/_ref, OpenMP, all host cores) on a bounded
             SNP-prefix sample of the same workload

  parity_check  outside the timed region: every returned model re-scored by the CPU oracle, a random sample of
             tuples evaluated by the oracle (bounded) and by the GPU's per-combination hook (10^6) must not rank
             before the last model of its fold unless it is in the list; at N > 1 all ranks must hold the same
             bytes and a small search must equal the 1-rank search of the same data

N = 1 runs BASELINE.json configs[1] ("c2": 10k SNPs x 2k samples, order 2, 10 folds).  N > 1 is weak
scaling: the c2 sample shape with 10k*sqrt(N) SNPs (N x the combinations), the linear combination
index space split into N contiguous ranges, one per rank; only the per-rank top-N lists cross NVLink.
At N > 1 the line also carries `named_configs`: the configs BASELINE.json names for several GPUs, run
strong-scaled (total work fixed) on the same ranks -- c3 always, c4 and c5 with --named all.
"""
import argparse
import json
import math
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SNP-combinations x folds evaluated/sec"
UNIT = "comb*folds/s"
RANK_SIZE = 50
FOLD_SEED = 20261017


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--snps", type=int, default=0, help="override the SNP count of the workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the post-run check of the returned models against the oracle")
    ap.add_argument("--strong", action="store_true", help="keep the total work fixed when N > 1")
    ap.add_argument("--named", default="c3", help="N > 1: named BASELINE configs to run strong-scaled after the main measurement "
                                                  "(comma-separated subset of c3,c4,c5; 'all'; 'none')")
    return ap.parse_args()


def workload(args, world):
    from hpg_variant_b200 import synth
    nv, A, U, order, folds, seed = synth.CONFIGS[args.workload]
    if args.snps:
        nv = args.snps
    scaling = "weak"
    if world > 1:
        if args.strong:
            scaling = "strong"
        else:
            nv = int(round(nv * world ** (1.0 / order)))
    name = {"c2": "c2: synthetic 10k SNPs x 2k samples (1k cases/1k controls), order 2, 10-fold CV",
            "c3": "c3: synthetic 100k SNPs x 4k samples, order 2, 10-fold CV",
            "c4": "c4: synthetic 5k SNPs x 4k samples, order 3, 5-fold CV",
            "c5": "c5: synthetic 20k SNPs x 50k samples, order 2, 10-fold CV"}[args.workload]
    return dict(name=name, nv=nv, A=A, U=U, order=order, folds=folds, seed=seed, scaling=scaling)


# ---------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        if not shutil.which("nvidia-smi"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Samples taken inside [t0, t1] (perf_counter); nvidia-smi needs about a second to start, so it is launched
        before the warm-up and the window is cut out afterwards."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for (t, r) in self.rows if t0 is None or (t0 <= t <= t1 + 0.15)]
        window = "timed region"
        if not rows and self.rows:           # region shorter than the sampling period: the samples next to it
            mid = 0.5 * (t0 + t1)
            rows = [r for (t, r) in sorted(self.rows, key=lambda x: abs(x[0] - mid))[:3]]
            window = "nearest samples (the timed region is shorter than the sampling period)"
        sm, smax, reasons = [], None, set()
        for r in rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                smax = float(p[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm),
                "window": window}


# ---------------------------------------------------------------------------------------------
# CPU baseline = the reference's own run_epistasis (or the oracle port when _ref is not built)
# ---------------------------------------------------------------------------------------------
def cpu_reference_run(g, A, U, order, folds, nv_sub, threads):
    """Times one reference run on the first nv_sub SNPs; returns (comb*folds/s, seconds, kind)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from hpg_variant_b200 import synth
    sub = np.ascontiguousarray(g[:nv_sub])
    ncomb = math.comb(nv_sub, order)
    if oracle_lib.available("ref"):
        ref = oracle_lib.Checker("ref")
        tmp = tempfile.mkdtemp(prefix="hpgv_bench_")
        try:
            path = os.path.join(tmp, "sub.bin")
            synth.write_dataset(path, sub, A, U)
            # order 3: the reference only enumerates all triples with a single block => stride = nv_sub
            stride = 100 if order == 2 else nv_sub
            t0 = time.perf_counter()
            ref.run_epistasis(path, os.path.join(tmp, "out"), order, stride, folds, 1, RANK_SIZE, 1, 1, threads)
            dt = time.perf_counter() - t0
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        used = threads if order == 2 else 1     # one block = one OpenMP task
        return ncomb * folds / dt, dt, "reference", used
    import hpg_variant_b200 as h
    orc = oracle_lib.Checker("oracle")
    fos, _ = h.k_folds(A, U, folds, FOLD_SEED)
    t0 = time.perf_counter()
    orc.search(sub, A, U, order, fos, 1, RANK_SIZE, threads=threads, num_folds=folds)
    dt = time.perf_counter() - t0
    return ncomb * folds / dt, dt, "port", threads


def sized_sample(order, nv, target_seconds, rate_comb_per_s):
    want = max(1.0, target_seconds * rate_comb_per_s)
    if order == 2:
        n = int((1 + math.sqrt(1 + 8 * want)) / 2)
    else:
        n = int(round((6 * want) ** (1 / 3))) + 2
    return max(order + 14, min(nv, n))


def cpu_baseline(g, w, target_seconds=12.0):
    threads = os.cpu_count() or 1
    order = w["order"]
    # calibration on a tiny prefix, then the bounded sample
    n0 = sized_sample(order, w["nv"], 1.0, 20000.0 * (threads if order == 2 else 1))
    rate0, _, kind, used = cpu_reference_run(g, w["A"], w["U"], order, w["folds"], n0, threads)
    n1 = sized_sample(order, w["nv"], target_seconds, rate0 / w["folds"])
    rate, dt, kind, used = cpu_reference_run(g, w["A"], w["U"], order, w["folds"], n1, threads)
    return {"value": rate, "unit": UNIT, "cores": used, "kind": kind,
            "sample": f"first {n1} SNPs of the workload ({math.comb(n1, order)} combinations x {w['folds']} folds), "
                      f"{'run_epistasis of the reference (OpenMP, stride ' + ('100' if order == 2 else str(n1)) + ')' if kind == 'reference' else 'oracle port (OpenMP)'}, "
                      f"{dt:.1f} s on {used} threads"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from hpg_variant_b200 import synth
    w = workload(args, 1)
    threads = os.cpu_count() or 1
    # only a SNP prefix is ever timed: generate just that much
    nv_gen = min(w["nv"], 6000 if w["order"] == 2 else 400)
    g = synth.make_dataset(nv_gen, w["A"], w["U"], w["seed"], order=w["order"])
    w_gen = dict(w, nv=nv_gen)
    order = w["order"]
    n0 = sized_sample(order, nv_gen, 1.0, 20000.0 * (threads if order == 2 else 1))
    rate0, _, kind, used = cpu_reference_run(g, w["A"], w["U"], order, w["folds"], n0, threads)
    total_budget = 150.0
    per_step = max(1.0, min(8.0, total_budget / max(1, args.steps + args.warmup)))
    n1 = sized_sample(order, nv_gen, per_step, rate0 / w["folds"])
    for _ in range(args.warmup):
        cpu_reference_run(g, w["A"], w["U"], order, w["folds"], n1, threads)
    times, rates = [], []
    for _ in range(args.steps):
        r, dt, kind, used = cpu_reference_run(g, w["A"], w["U"], order, w["folds"], n1, threads)
        times.append(dt)
        rates.append(r)
    ncomb = math.comb(n1, order)
    value = ncomb * w["folds"] * len(times) / sum(times)
    sample = (f"first {n1} SNPs of the workload ({ncomb} combinations x {w['folds']} folds) per step, "
              f"{'the reference run_epistasis (OpenMP)' if kind == 'reference' else 'oracle port (OpenMP)'} on {used} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": w["name"], "order": order, "num_variants": w["nv"], "num_affected": w["A"], "num_unaffected": w["U"],
                   "num_folds": w["folds"], "rank_size": RANK_SIZE, "eval_subset": "training", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def kernel_counts(workload_name):
    """Per-launch DRAM bytes and executed POPC / ALU-pipe instructions per combination of the dominant kernel, from the
    committed `ncu --set full` captures (profiles/kernel_counts.json names the capture each figure comes from)."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_counts.json")) as fh:
            return json.load(fh).get(workload_name)
    except Exception:
        return None


def parity_check(eng, w, g, fos, got, first, last, budget_s=45.0, gpu_samples=1_000_000, seed=7):
    """Outside the timed region.  `got` = this rank's final F x N models (after the cross-rank merge when N > 1).
    (1) every model re-scored by the CPU ORACLE (risky cells, confusion matrix, accuracy: equality);
    (2) lists in canonical order, no duplicates;
    (3) random tuples -- a bounded sample through the oracle, 10^6 through the GPU's per-combination hook (a different,
        simple kernel that the tests pin to the oracle) -- none may rank before the last model of its fold unless it is in
        the list;
    (4) how many of the planted tuples made the lists (reported, not asserted: a strong single SNP can outrank them)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    import hpg_variant_b200 as h
    orc = oracle_lib.Checker("oracle")
    nv, A, U, order, F = w["nv"], w["A"], w["U"], w["order"], w["folds"]
    t0 = time.perf_counter()
    res = {"models_rescored": 0, "oracle_samples": 0, "gpu_samples": 0, "violations": 0, "ok": True}
    rank_n = got.shape[1]
    for f in range(F):
        combs = np.ascontiguousarray(got["snp"][f, :, :order])
        ov = orc.eval(g, A, U, order, fos, 1, combs)
        same = (np.array_equal(ov["risky_mask"][:, f], got["risky_mask"][f]) and np.array_equal(ov["conf"][:, f], got["conf"][f])
                and np.array_equal(ov["ba"][:, f], got["accuracy"][f]))
        keys = [(-got["accuracy"][f, r],) + tuple(int(x) for x in got["snp"][f, r, :order]) for r in range(rank_n)]
        if not same or keys != sorted(keys) or len(set(keys)) != rank_n:
            res["ok"] = False
            res["violations"] += 1
        res["models_rescored"] += rank_n
    rng = np.random.default_rng(seed)

    def draw(n):
        c = np.sort(rng.integers(0, nv, (n, order)), axis=1).astype(np.int32)
        return np.ascontiguousarray(c[(np.diff(c, axis=1) > 0).all(axis=1)])

    def check(combs, ba):
        bad = 0
        for f in range(F):
            lastm = got[f, rank_n - 1]
            inlist = {tuple(int(x) for x in t) for t in got["snp"][f, :, :order]}
            for n in np.nonzero(ba[:, f] >= lastm["accuracy"])[0]:
                t = tuple(int(x) for x in combs[n])
                if (ba[n, f] > lastm["accuracy"] or t < tuple(int(x) for x in lastm["snp"][:order])) and t not in inlist:
                    bad += 1
        return bad

    # oracle sample, sized from a calibration batch so that the whole check stays inside its budget
    cal = draw(64)
    t1 = time.perf_counter()
    ov = orc.eval(g, A, U, order, fos, 1, cal)
    per = (time.perf_counter() - t1) / max(1, cal.shape[0])
    left = budget_s - (time.perf_counter() - t0)
    n_or = int(max(0, min(50_000, 0.5 * left / max(per, 1e-9))))
    if n_or > 0:
        combs = draw(n_or)
        ov = orc.eval(g, A, U, order, fos, 1, combs)
        res["violations"] += check(combs, ov["ba"])
        res["oracle_samples"] = int(combs.shape[0])
    done = 0
    while done < gpu_samples:
        combs = draw(min(250_000, gpu_samples - done) + 64)
        ev = eng.eval(order, combs, h.SUBSET_TRAINING)
        res["violations"] += check(combs, ev["ba"])
        done += combs.shape[0]
    res["gpu_samples"] = int(done)
    if w.get("planted"):
        res["planted"] = [list(t) for t in w["planted"]]
        res["planted_in_lists"] = [int(sum(tuple(t) in {tuple(int(x) for x in m) for m in got["snp"][f, :, :order]} for f in range(F))) for t in w["planted"]]
    res["ok"] = bool(res["ok"] and res["violations"] == 0)
    res["seconds"] = round(time.perf_counter() - t0, 1)
    res["checker"] = "oracle/liboracle.so (re-score + sample) and hpgv_epi_eval (10^6 sample)"
    return res


class Runner:
    """One rank's GPU, engine and process group; runs workloads on them."""

    def __init__(self, args):
        import torch
        import hpg_variant_b200 as h
        self.torch, self.h, self.args = torch, h, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- hpg_variant_b200 has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        self.eng = h.EpistasisEngine(self.local_rank)
        self.stream = torch.cuda.current_stream()
        self.eng.set_stream(self.stream.cuda_stream)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def reduce(self, values, op):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return t.tolist()

    def run(self, w, steps, warmup, sampler=None, e2e=True):
        """Times `steps` steps of workload w (pack + search [+ all-gather + merge]); returns the measurements."""
        torch, h, eng, dist, world, rank = self.torch, self.h, self.eng, self.dist, self.world, self.rank
        from hpg_variant_b200 import sharding, synth
        nv, A, U, order, F = w["nv"], w["A"], w["U"], w["order"], w["folds"]
        S = A + U
        total = h.num_combinations(nv, order)
        first, last = sharding.shard_range(total, rank, world)
        g_pinned = torch.empty((nv, S), dtype=torch.uint8).pin_memory()
        g = g_pinned.numpy()
        planted = []
        synth.make_dataset(nv, A, U, w["seed"], order=order, out=g, planted_out=planted)
        w = dict(w, planted=planted)
        fos, _ = h.k_folds(A, U, F, FOLD_SEED)
        d_raw = g_pinned.cuda(non_blocking=False)
        shard = sharding.ShardedSearch(eng, dist, rank, world, F, RANK_SIZE, "cuda")
        d_local, d_all, d_final = shard.d_local, shard.d_all, shard.d_final
        h_out = np.zeros((F, RANK_SIZE), h.MODEL_DTYPE)
        stream = self.stream
        phase_ev = []

        def step_device(record=False):
            """pack + search (+ all-gather + merge) with the genotype bytes resident in HBM"""
            eng.load_dataset_device(d_raw.data_ptr(), nv, A, U)
            eng.set_folds(F, fos)
            if world == 1 or not record:
                shard.run(order, h.SUBSET_TRAINING, total)     # search [-> all-gather -> merge]
                return
            # the same calls as ShardedSearch.run, with events between them (where a multi-GPU step spends its time)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            eng.search_device(order, h.SUBSET_TRAINING, RANK_SIZE, first, last, d_local.data_ptr())
            e[0].record(stream)
            sharding.all_gather_models(dist, d_local, world, out=d_all)
            e[1].record(stream)
            eng.merge_device(order, h.SUBSET_TRAINING, world, RANK_SIZE, d_all.data_ptr(), d_final.data_ptr())
            e[2].record(stream)
            phase_ev.append(e)

        # first call of a shape: the host-built work list (unit descriptors) is made and uploaded once, then cached
        self.barrier()
        t0 = time.perf_counter()
        step_device()
        self.barrier()
        first_call_s = time.perf_counter() - t0
        for _ in range(max(warmup, 3) - 1):
            step_device()
        self.barrier()

        # ---- timed region: K steps, CUDA events per step, L2 flushed between steps ----
        launches0 = eng.launch_count
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        t_wall0 = time.perf_counter()
        for k in range(steps):
            self.flush.fill_(k & 0xFF)            # > L2 (126 MB): next step starts with a cold L2
            ev[k][0].record(stream)
            step_device(record=True)
            ev[k][1].record(stream)
        self.barrier()
        search_ms = eng.search_times(min(steps, 32))           # the library's CUDA events around each search launch
        t_wall = time.perf_counter() - t_wall0
        clocks = sampler.stop(t_wall0, t_wall0 + t_wall) if sampler else None
        launches = eng.launch_count - launches0
        my_dev_ms = sum(a.elapsed_time(b) for a, b in ev)
        dev_ms = self.reduce([my_dev_ms], "MAX")[0]
        k_ms = float(np.mean(search_ms))
        kmin, kmax = self.reduce([k_ms], "MIN")[0], self.reduce([k_ms], "MAX")[0]
        phases = None
        if world > 1 and phase_ev:
            # per-rank means: pack + search, wait for the slowest rank + all-gather, final merge; search kernel alone
            p = [np.mean([ev[k][0].elapsed_time(phase_ev[k][0]) for k in range(steps)]),
                 np.mean([phase_ev[k][0].elapsed_time(phase_ev[k][1]) for k in range(steps)]),
                 np.mean([phase_ev[k][1].elapsed_time(phase_ev[k][2]) for k in range(steps)]), k_ms]
            tmax, tmin = self.reduce(p, "MAX"), self.reduce(p, "MIN")
            names = ["pack+search", "all_gather(incl. wait for the slowest rank)", "merge", "search_kernel"]
            phases = {n: {"min_over_ranks": float(a), "max_over_ranks": float(b)} for n, a, b in zip(names, tmin, tmax)}
        dump = os.environ.get("HPGV_BENCH_DUMP_STEPS")          # development: per-step times of every rank, one file per rank
        if dump:
            os.makedirs(dump, exist_ok=True)
            rec = {"rank": rank, "step_ms": [a.elapsed_time(b) for a, b in ev], "search_kernel_ms": [float(x) for x in search_ms]}
            if world > 1 and phase_ev:
                rec["pack_search_ms"] = [ev[k][0].elapsed_time(phase_ev[k][0]) for k in range(steps)]
                rec["gather_ms"] = [phase_ev[k][0].elapsed_time(phase_ev[k][1]) for k in range(steps)]
            with open(os.path.join(dump, f"steps_{w['name'].split(':')[0]}_n{world}_rank{rank}.json"), "w") as fh:
                json.dump(rec, fh)
        out = dict(w=w, g=g, fos=fos, total=total, first=first, last=last, dev_ms=dev_ms, steps=steps, search_ms=search_ms, k_ms=k_ms,
                   k_ms_min=kmin, k_ms_max=kmax, my_step_ms=my_dev_ms / steps, phases=phases, launches=launches, clocks=clocks,
                   t_wall=t_wall, first_call_s=first_call_s, value=total * F * steps / (dev_ms * 1e-3), lay=eng.layout(),
                   final=shard.result().copy())

        # every rank must hold the same final ranking (the all-gathered lists and the merge are deterministic)
        if world > 1:
            mine = torch.from_numpy(out["final"].view(np.uint8).reshape(-1).copy()).cuda()
            allb = torch.empty(world * mine.numel(), dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(allb, mine)
            allb = allb.cpu().numpy().reshape(world, -1)
            out["ranks_identical"] = bool((allb == allb[0]).all())

        # ---- e2e: host buffers in, host result out ----
        if e2e:
            # N = 1: the C-ABI call hpgv_epi_run_host (host pointers in and out).  N > 1: ShardedSearch.run_from_host -- every
            # rank uploads 1/N of the SNP rows from its pinned host copy, the slices are all-gathered over NVLink, then
            # pack + search + all-gather of the lists + merge, and the final ranking comes back to the host of every rank.
            def e2e_step():
                if world == 1:
                    eng.run_host(g, A, U, F, fos, order, h.SUBSET_TRAINING, RANK_SIZE, first, last, out=h_out)
                else:
                    shard.run_from_host(g_pinned, A, U, F, fos, order, h.SUBSET_TRAINING, total)
            for _ in range(2):
                e2e_step()
            self.barrier()
            e2e_steps = max(3, min(steps, 10))
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            self.barrier()
            e2e_s = self.reduce([time.perf_counter() - t0], "MAX")[0]
            out["e2e_value"] = total * F * e2e_steps / e2e_s
            out["e2e_steps"] = e2e_steps
            # leave the engine on the device-resident data set (the parity check evaluates tuples on it)
            eng.load_dataset_device(d_raw.data_ptr(), nv, A, U)
            eng.set_folds(F, fos)
        out["keep"] = (g_pinned, d_raw)
        return out

    def small_equals_one_rank(self, w):
        """N > 1: a small search of w's sample shape, sharded + all-gathered + merged, must equal the 1-rank search."""
        torch, h, eng, dist, world, rank = self.torch, self.h, self.eng, self.dist, self.world, self.rank
        from hpg_variant_b200 import sharding, synth
        nv = 2500 if w["order"] == 2 else 160
        A, U, order, F = w["A"], w["U"], w["order"], w["folds"]
        g = synth.make_dataset(nv, A, U, w["seed"] + 1, order=order)
        fos, _ = h.k_folds(A, U, F, FOLD_SEED)
        eng.load_dataset(g, A, U)
        eng.set_folds(F, fos)
        total = h.num_combinations(nv, order)
        shard = sharding.ShardedSearch(eng, dist, rank, world, F, RANK_SIZE, "cuda")
        shard.run(order, h.SUBSET_TRAINING, total)
        merged = shard.result().copy()
        full = eng.search(order, h.SUBSET_TRAINING, RANK_SIZE)
        ok = merged.tobytes() == full.tobytes()
        return bool(self.reduce([1.0 if ok else 0.0], "MIN")[0] == 1.0)


def roofline_of(runner, r, name, peaks, popc_peak, alu_peak):
    w, lay = r["w"], r["lay"]
    order = w["order"]
    my_combs = r["last"] - r["first"]
    Wwords = lay["words_per_class_row"]
    popc_per_comb = (3 ** order) * Wwords                 # algorithmic POPC32 per combination, all folds together
    k_ms = r["k_ms"]
    achieved = my_combs * popc_per_comb / (k_ms * 1e-3)
    plane_bytes = lay["plane_bytes"]
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    kc = kernel_counts(name) if not runner.args.snps else None
    roof = {
        "bound": "int_popc", "kernel": f"search{order}_kernel", "achieved": achieved / 1e12, "peak": popc_peak / 1e12, "unit": "TPOPC32/s",
        "frac": achieved / popc_peak, "traffic": kc.get("dram_bytes_per_launch") if kc and runner.world == 1 else None,
        "traffic_source": (kc.get("traffic_source") or kc.get("source")) if kc and runner.world == 1 else None,
        "algorithmic": f"3^{order} x W = {popc_per_comb} POPC32 per combination (W = {Wwords} words), {my_combs} combinations per launch",
        "peak_source": "POPC micro-benchmark in this run (hpgv_epi_pipe_peak), all SMs",
        "kernel_ms": k_ms, "kernel_share_of_step": k_ms / r["my_step_ms"],
        "hbm": {"compulsory_bytes": plane_bytes, "achieved_gbs": plane_bytes / (k_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s", "frac": plane_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak},
    }
    if kc and kc.get("popc_warp_inst_per_comb"):
        # what the kernel EXECUTES (counted by ncu in the committed capture) against the measured pipe peaks: the distance to the machine
        xu = my_combs * kc["popc_warp_inst_per_comb"] * 32 / (k_ms * 1e-3)
        alu = my_combs * kc["alu_warp_inst_per_comb"] * 32 / (k_ms * 1e-3)
        roof["executed"] = {"popc_frac_of_xu_peak": xu / popc_peak, "alu_frac_of_alu_peak": alu / alu_peak if alu_peak else None,
                            "popc32_per_comb": kc["popc_warp_inst_per_comb"] * 32, "alu_ops_per_comb": kc["alu_warp_inst_per_comb"] * 32,
                            "alu_peak_tops": alu_peak / 1e12 if alu_peak else None, "source": kc.get("source")}
    return roof


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    runner = Runner(args)
    world, rank = runner.world, runner.rank
    eng = runner.eng
    w = workload(args, world)
    sampler = ClockSampler(runner.local_rank)
    if rank == 0 and os.environ.get("HPGV_BENCH_NO_SAMPLER") != "1":    # rank 0's GPU is the one whose clocks the line reports
        sampler.start()
    r = runner.run(w, args.steps, args.warmup, sampler=sampler, e2e=True)
    w = r["w"]
    nv, A, U, order, F = w["nv"], w["A"], w["U"], w["order"], w["folds"]
    S = A + U
    lay = r["lay"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    popc_peak = max(eng.pipe_peak(0, 2000), eng.pipe_peak(0, 4000))
    alu_peak = max(eng.pipe_peak(1, 2000), eng.pipe_peak(1, 4000))
    roofline = roofline_of(runner, r, args.workload, peaks, popc_peak, alu_peak)

    # ---- parity of what was just timed (outside the timed region) ----
    parity = None
    if not args.no_parity_check:
        parity = {}
        if world > 1:
            parity["ranks_identical"] = r.get("ranks_identical")
            parity["small_sharded_equals_one_rank"] = runner.small_equals_one_rank(w)
        if rank == 0:
            try:
                runner.eng.load_dataset_device(r["keep"][1].data_ptr(), nv, A, U)
                runner.eng.set_folds(F, r["fos"])
                parity.update(parity_check(eng, w, r["g"], r["fos"], r["final"], r["first"], r["last"]))
            except Exception as e:
                parity.update({"ok": False, "error": repr(e)})
        runner.barrier()
        if world > 1 and rank == 0:
            parity["ok"] = bool(parity.get("ok") and parity["ranks_identical"] and parity["small_sharded_equals_one_rank"])

    # ---- named configs of BASELINE.json on the same ranks (strong scaling: the total work is fixed) ----
    named = {}
    want = [] if args.named == "none" else (["c3", "c4", "c5"] if args.named == "all" else [x for x in args.named.split(",") if x])
    if args.snps or args.strong or args.workload != "c2":
        want = []
    r_keep = r.pop("keep")
    del r_keep
    for name in want:
        from hpg_variant_b200 import synth
        cnv, cA, cU, corder, cF, cseed = synth.CONFIGS[name]
        cw = dict(name=name, nv=cnv, A=cA, U=cU, order=corder, folds=cF, seed=cseed, scaling="strong")
        nsteps = 3 if name != "c4" else 2
        try:
            cr = runner.run(cw, nsteps, 1, sampler=None, e2e=False)
            roof = roofline_of(runner, cr, name, peaks, popc_peak, alu_peak)
            entry = {"value": cr["value"], "unit": UNIT, "ms_per_step": cr["dev_ms"] / nsteps, "steps": nsteps, "scaling": "strong",
                     "combinations": cr["total"], "per_rank_kernel_ms": {"min": cr["k_ms_min"], "max": cr["k_ms_max"]},
                     "roofline_frac_rank0": roof["frac"], "executed": roof.get("executed"), "first_call_s": cr["first_call_s"],
                     "ranks_identical": cr.get("ranks_identical")}
            if not args.no_parity_check and rank == 0:
                try:
                    pc = parity_check(eng, cr["w"], cr["g"], cr["fos"], cr["final"], cr["first"], cr["last"], budget_s=25.0, gpu_samples=200_000)
                    entry["parity_check"] = pc
                except Exception as e:
                    entry["parity_check"] = {"ok": False, "error": repr(e)}
            named[name + "_strong"] = entry
            del cr
        except Exception as e:                     # a named config must never take the headline line down
            named[name + "_strong"] = {"error": repr(e)}
        runner.barrier()

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            try:
                cpu = cpu_baseline(r["g"], w)
            except Exception as e:        # the baseline is a reported figure; never let it kill the bench line
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": f"failed: {e}"}
        line = {
            "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": r["dev_ms"] / args.steps, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": w["name"] if world == 1 else w["name"] + f" -- weak-scaled to {nv} SNPs for {world} GPUs",
                       "order": order, "num_variants": nv, "num_affected": A, "num_unaffected": U, "num_folds": F,
                       "rank_size": RANK_SIZE, "eval_subset": "training", "combinations": r["total"],
                       "sharding": f"{world} contiguous combination-index ranges", "l2": "flushed between timed steps (256 MiB write)",
                       "layout": lay},
            "clocks": r["clocks"],
            "e2e": {"value": r["e2e_value"], "unit": UNIT,
                    "h2d_bytes_per_step": int(-(-nv // world) * S + lay["num_blocks"] * (4 if lay["block_words"] == 3 else lay["block_words"]) * 32 * 4),
                    "h2d_note": "per rank: its 1/N slice of the genotype rows (all-gathered over NVLink when N > 1) + the fold permutation",
                    "d2h_bytes_per_step": F * RANK_SIZE * 40, "steps": r["e2e_steps"],
                    "first_call_s": r["first_call_s"],
                    "first_call_note": "first pack+search of this shape, once per run: builds and uploads the work list (unit descriptors), sizes the buffers; later calls reuse them"},
            "gpu_launches": int(r["launches"]),
            "roofline": roofline,
            "wall_s_timed_region": r["t_wall"],
        }
        if r["phases"]:
            line["phases_ms"] = r["phases"]
        if parity is not None:
            line["parity_check"] = parity
        if named:
            line["named_configs"] = named
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        runner.dist.barrier()
        runner.dist.destroy_process_group()


if __name__ == "__main__":
    main()
