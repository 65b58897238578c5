"""bench.py's host-side helpers (no GPU): workload shapes and weak scaling, the bounded CPU sample, the clock sampler's
window and throttle-reason parsing, and the reference arm's argument handling."""
import argparse
import math
import numpy as np
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def _args(**kw):
    d = dict(gpus=1, steps=5, warmup=3, impl="ours", workload="c2", snps=0, no_cpu_baseline=True, strong=False)
    d.update(kw)
    return argparse.Namespace(**d)


def test_workload_shapes_follow_baseline_configs():
    w = bench.workload(_args(), 1)
    assert (w["nv"], w["A"], w["U"], w["order"], w["folds"]) == (10000, 1000, 1000, 2, 10) and w["scaling"] == "weak"
    assert bench.workload(_args(workload="c3"), 1)["nv"] == 100000
    w4 = bench.workload(_args(workload="c4"), 1)
    assert (w4["nv"], w4["order"], w4["folds"]) == (5000, 3, 5)
    assert bench.workload(_args(workload="c5"), 1)["A"] == 25000


def test_weak_scaling_keeps_the_work_per_rank():
    base = math.comb(10000, 2)
    for n in (2, 4, 8):
        w = bench.workload(_args(gpus=n), n)
        per_rank = math.comb(w["nv"], 2) / n
        assert abs(per_rank / base - 1.0) < 0.01 and w["scaling"] == "weak"
    w3 = bench.workload(_args(workload="c4", gpus=8), 8)
    assert abs(math.comb(w3["nv"], 3) / 8 / math.comb(5000, 3) - 1.0) < 0.01
    assert bench.workload(_args(gpus=4, strong=True), 4) == dict(bench.workload(_args(), 1), scaling="strong")


def test_sized_sample_is_bounded_by_the_workload():
    assert bench.sized_sample(2, 10000, 10.0, 1e9) == 10000                   # never more SNPs than the workload has
    n = bench.sized_sample(2, 10000, 10.0, 1e5)
    assert 16 <= n <= 10000 and abs(math.comb(n, 2) - 1e6) / 1e6 < 0.01       # ~ target_seconds * rate combinations
    n3 = bench.sized_sample(3, 5000, 10.0, 1e4)
    assert 17 <= n3 <= 5000 and 0.5e5 < math.comb(n3, 3) < 2e5


def test_clock_sampler_window_and_reasons():
    s = bench.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None, "wait": lambda self, timeout=None: 0, "kill": lambda self: None})()
    ok = "1965, 1965, Not Active, Not Active, Not Active, Not Active, 512.3"
    cap = "1800, 1965, Not Active, Not Active, Not Active, Active, 998.0"
    hot = "1200, 1965, Not Active, Active, Not Active, Not Active, 700.0"
    s.rows = [(0.5, hot), (1.1, ok), (1.2, cap), (1.3, ok), (2.0, hot)]
    c = s.stop(1.0, 1.35)
    assert c["samples"] == 3 and c["sm_mhz"] == 1965.0 and c["sm_max_mhz"] == 1965.0
    assert c["reasons"] == ["sw_power_cap"] and c["window"] == "timed region"
    s.rows = [(0.5, hot), (2.0, ok)]
    c = s.stop(1.0, 1.1)                                                      # region between two samples: the nearest ones
    assert c["samples"] == 2 and "nearest" in c["window"] and "hw_thermal_slowdown" in c["reasons"]


def test_reference_arm_runs_on_rank_zero_only(monkeypatch, capsys):
    monkeypatch.setenv("RANK", "3")
    bench.run_reference_arm(_args(impl="reference", gpus=8))                  # other ranks: no work, no output
    assert capsys.readouterr().out == ""


class _OracleAsEngine:
    """bench.parity_check takes the GPU engine for its 10^6-tuple sample; on the CPU the oracle answers in its place."""

    def __init__(self, oracle, g, A, U, fos):
        self.o, self.g, self.A, self.U, self.fos = oracle, g, A, U, fos

    def eval(self, order, combs, subset):
        return self.o.eval(self.g, self.A, self.U, order, self.fos, subset, combs)


def test_parity_check_accepts_a_correct_result_and_flags_a_wrong_one(oracle):
    """bench.py's parity_check (oracle re-score + sample check of the models a bench run returns) must pass on the true
    ranking and fail when a model is altered, dropped from the top, or listed out of order."""
    import bench
    from hpg_variant_b200 import synth
    from hpg_variant_b200._lib import MODEL_DTYPE
    nv, A, U, F, rank = 60, 80, 80, 3, 8
    g = synth.make_dataset(nv, A, U, seed=5, missing=0.01, planted=2)
    fos = (np.concatenate([np.arange(A), np.arange(U)]) % F).astype(np.int32)
    want, _ = oracle.search(g, A, U, 2, fos, 1, rank, threads=2, num_folds=F)
    good = np.zeros((F, rank), MODEL_DTYPE)
    good["accuracy"], good["snp"], good["risky_mask"], good["conf"] = want["ba"], want["snp"], want["risky_mask"], want["conf"]
    w = dict(nv=nv, A=A, U=U, order=2, folds=F)
    eng = _OracleAsEngine(oracle, g, A, U, fos)
    res = bench.parity_check(eng, w, g, fos, good, 0, nv * (nv - 1) // 2, budget_s=20.0, gpu_samples=3000)
    assert res["ok"] and res["violations"] == 0 and res["models_rescored"] == F * rank and res["gpu_samples"] >= 2000
    bad = good.copy()
    bad["conf"][1, 2, 0] += 1                                   # a confusion matrix that does not re-score
    assert not bench.parity_check(eng, w, g, fos, bad, 0, 1, budget_s=5.0, gpu_samples=200)["ok"]
    bad = good.copy()
    bad[0, :-1] = good[0, 1:]                                   # the best model of fold 0 is missing from its list
    bad[0, -1] = good[0, -1]
    bad[0, -2] = good[0, -1]
    assert not bench.parity_check(eng, w, g, fos, bad, 0, 1, budget_s=20.0, gpu_samples=4000)["ok"]
    bad = good.copy()
    bad[2, [0, 1]] = good[2, [1, 0]]                            # out of order
    assert not bench.parity_check(eng, w, g, fos, bad, 0, 1, budget_s=5.0, gpu_samples=200)["ok"]
