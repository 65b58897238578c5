"""GPU parity against the reference's OWN answers (tests/golden/epi_golden.json, produced by running the reference's
code in tests/golden/make_golden.py) and at the scales bench.py runs.

  * every golden eval / top-N case through hpgv_epi_eval / hpgv_epi_search, incl. BASELINE configs[0] =
    test/epistasis_dataset.bin (`fixture_order2/3`, loaded through hpgv_epi_load_dataset_file with its legacy header)
    and the Appendix-D anchor rows;
  * the DEVICE high-risk function (epi_device.cuh high_risk, the one the search kernels call) on the golden (ca, cu) grids
    of mdr_high_risk_combinations2, A = 1900 / U = 2100 (SURVEY F5 float32 ties) included, and around the decision
    boundary of unbalanced cohorts against the oracle's float32 sequence;
  * evaluate_model's functions on the device against the golden values and the oracle, and searches ranked by them;
  * merge_rankings / epistasis_report of the reference on GPU-found rankings;
  * scale: c2 in full against the oracle's exhaustive search, c3/c4/c5-shaped searches re-scored and sample-checked.
"""
import os
import struct

import numpy as np
import pytest

import hpg_variant_b200 as h
from hpg_variant_b200 import synth
from golden_util import eval_case_arrays, load_golden, unb64

pytestmark = pytest.mark.gpu
GOLD = load_golden()


# ---- golden eval / top-N through the CUDA path -----------------------------------------------------------------------
@pytest.mark.parametrize("rec", GOLD["eval"], ids=lambda r: r["name"])
def test_eval_matches_reference_golden(engine, rec):
    g, A, U, F, order, fos, combs, want = eval_case_arrays(rec)
    engine.load_dataset(g, A, U)
    engine.set_folds(F, fos)
    for subset, sname in ((h.SUBSET_TRAINING, "training"), (h.SUBSET_TESTING, "testing")):
        got = engine.eval(order, combs, subset)
        assert np.array_equal(got["counts_aff"], want["counts_aff"])
        assert np.array_equal(got["counts_unaff"], want["counts_unaff"])
        assert np.array_equal(got["risky_mask"], want["risky_mask"])
        assert np.array_equal(got["conf"], want["conf_" + sname])
        assert np.array_equal(got["ba"], want["ba_" + sname], equal_nan=True)


@pytest.mark.parametrize("rec", GOLD["topn"], ids=lambda r: f'{r["name"]}-{r["subset"]}')
def test_topn_matches_reference_golden(engine, rec):
    case = next(c for c in GOLD["eval"] if c["name"] == rec["name"])
    g, A, U, F, order, fos, _, _ = eval_case_arrays(case)
    want = unb64(rec["models"], h.MODEL_DTYPE, (F, rec["rank"]))
    engine.load_dataset(g, A, U)
    engine.set_folds(F, fos)
    got = engine.search(order, rec["subset"], rec["rank"])
    assert np.array_equal(got["snp"][..., :order], want["snp"][..., :order])
    assert np.array_equal(got["risky_mask"], want["risky_mask"]) and np.array_equal(got["conf"], want["conf"])
    assert np.array_equal(got["accuracy"], want["accuracy"], equal_nan=True)


def test_c1_fixture_file_through_the_cuda_path(engine, tmp_path):
    """BASELINE.json configs[0]: test/epistasis_dataset.bin as shipped (size_t + 2 x uint32 header) loaded by
    hpgv_epi_load_dataset_file, order 2, 10 folds: Appendix D's anchor rows and the reference's top-N."""
    case = next(c for c in GOLD["eval"] if c["name"] == "fixture_order2")
    g, A, U, F, order, fos, combs, want = eval_case_arrays(case)
    path = tmp_path / "epistasis_dataset.bin"
    with open(path, "wb") as fh:                       # the fixture's own layout (test/test_epistasis_dataset.c:162-178)
        fh.write(struct.pack("<QII", g.shape[0], A, U))
        fh.write(g.tobytes())
    engine.load_dataset_file(path)
    assert engine.dims() == (4, 49, 98)
    engine.set_folds(F, fos)
    got = engine.eval(2, combs, h.SUBSET_TRAINING)
    # pair (0,1) totals and fold 0 / fold 9 rows
    assert (got["counts_aff"][0].sum(0) // (F - 1)).tolist() == [0, 0, 1, 0, 0, 13, 0, 0, 34]
    assert (got["counts_unaff"][0].sum(0) // (F - 1)).tolist() == [0, 0, 4, 0, 0, 20, 0, 1, 70]
    assert got["risky_mask"][0, 0] == 0x020 and got["conf"][0, 0].tolist() == [11, 33, 17, 71] and got["ba"][0, 0] == 0.52840909090909083
    assert got["risky_mask"][0, 9] == 0x020 and got["conf"][0, 9].tolist() == [13, 32, 18, 71] and got["ba"][0, 9] == 0.54332084893882637
    assert got["risky_mask"][1, 7] == 0x011 and got["conf"][1, 7].tolist() == [12, 32, 15, 73] and got["ba"][1, 7] == 0.55113636363636365
    t = engine.eval(2, combs, h.SUBSET_TESTING)
    assert t["conf"][0, 8].tolist() == [1, 4, 9, 0] and t["ba"][0, 8] == 0.10000000000000001
    for rec in (r for r in GOLD["topn"] if r["name"] == "fixture_order2"):
        top = engine.search(2, rec["subset"], rec["rank"])
        ref = unb64(rec["models"], h.MODEL_DTYPE, (F, rec["rank"]))
        assert top.tobytes() == ref.tobytes() or (np.array_equal(top["snp"][..., :2], ref["snp"][..., :2]) and np.array_equal(top["conf"], ref["conf"])
                                                    and np.array_equal(top["accuracy"], ref["accuracy"], equal_nan=True))


# ---- the device high-risk function itself ------------------------------------------------------------------------------
@pytest.mark.parametrize("rec", GOLD["risk"], ids=lambda r: f'A{r["A"]}U{r["U"]}')
def test_device_high_risk_matches_reference_grid(engine, rec):
    lim = rec["lim"]
    ca, cu = np.meshgrid(np.arange(lim), np.arange(lim), indexing="ij")
    want = np.unpackbits(unb64(rec["flags"], np.uint8))[: lim * lim].astype(bool)
    got = engine.high_risk(ca.ravel(), cu.ravel(), rec["A"], rec["U"])
    assert np.array_equal(got, want)


def test_device_high_risk_vector_of_test_mdr(engine):
    r = GOLD["risk_test_mdr"]                          # test/test_mdr.c:52-66
    assert engine.high_risk(r["ca"], r["cu"], r["A"], r["U"]).astype(int).tolist() == r["flags"]


@pytest.mark.parametrize("A,U", [(1900, 2100), (1000, 3000), (333, 777), (49, 98), (25000, 24999), (7, 65000), (60001, 3)])
def test_device_high_risk_near_the_decision_boundary(engine, oracle, A, U):
    """Unbalanced cohorts: every cu in range with the ca values around cu * A / U (where the float32 sequence and the exact
    ratio can disagree) plus random pairs -- the device's band test + float32 replay against the oracle's
    restatement of mdr.c:45-75."""
    rng = np.random.default_rng(A * 7 + U)
    cu = np.repeat(np.arange(0, min(U, 40000) + 1), 7)
    ca = np.clip(np.round(cu * (A / U)).astype(np.int64) + np.tile(np.arange(-3, 4), cu.size // 7), 0, A)
    ca = np.concatenate([ca, rng.integers(0, A + 1, 200000)])
    cu = np.concatenate([cu, rng.integers(0, U + 1, 200000)])
    got = engine.high_risk(ca, cu, A, U)
    want = oracle.high_risk(ca, cu, A, U)
    assert np.array_equal(got, want)


# ---- evaluation functions ------------------------------------------------------------------------------------------------
def test_device_evaluate_matches_reference_formulas(engine, oracle):
    codes = {"CA": h.EVAL_CA, "BA": h.EVAL_BA, "GAMMA": h.EVAL_GAMMA, "TAU_B": h.EVAL_TAU_B}
    for rec in GOLD["formulas"]:                       # test/test_epistasis_model.c:513-534, values from the reference
        for name, want in zip(rec["functions"], rec["values"]):
            assert engine.evaluate([rec["conf"]], codes[name])[0] == want
    assert engine.evaluate([[40, 2, 4, 10]], h.EVAL_CA_TRUE)[0] == 50 / 56
    rng = np.random.default_rng(5)
    conf = rng.integers(0, 3000, (20000, 4)).astype(np.uint32)
    conf[:50] = rng.integers(0, 2, (50, 4))            # zero rows and columns: NaN / inf like the reference
    for fn in (h.EVAL_CA, h.EVAL_BA, h.EVAL_GAMMA, h.EVAL_TAU_B):
        got = engine.evaluate(conf, fn)
        with np.errstate(all="ignore"):
            want = np.array([oracle.evaluate(m, fn) for m in conf[:3000]])
        assert np.array_equal(got[:3000], want, equal_nan=True)
    with pytest.raises(h.HpgvError):
        engine.evaluate(conf[:4], h.EVAL_WBA)


@pytest.mark.parametrize("fn", [h.EVAL_GAMMA, h.EVAL_TAU_B, h.EVAL_CA_TRUE, h.EVAL_CA])
@pytest.mark.parametrize("order,nv,A,U,F", [(2, 60, 300, 300, 5), (2, 50, 230, 460, 4), (3, 16, 200, 200, 3), (2, 40, 2500, 2500, 2)])
def test_search_ranked_by_other_evaluation_functions(engine, oracle, fn, order, nv, A, U, F):
    """CA / gamma / tau-b rank the models (model.h:84, model.c:462-479); the oracle evaluates the same
    confusion matrices with the reference's evaluate_model."""
    rng = np.random.default_rng(nv + fn)
    g = synth.make_dataset(nv, A, U, seed=nv * 3 + fn, order=order, missing=0.01, planted=2)
    fos = np.concatenate([rng.permutation(A) % F, rng.permutation(U) % F]).astype(np.int32)
    engine.load_dataset(g, A, U)
    engine.set_folds(F, fos)
    try:
        engine.set_eval_function(fn)
        oracle.set_eval_function(fn)
        for subset in (h.SUBSET_TRAINING, h.SUBSET_TESTING):
            got = engine.search(order, subset, 25)
            want, _ = oracle.search(g, A, U, order, fos, subset, 25, threads=8, num_folds=F)
            assert np.array_equal(got["snp"][..., :order], want["snp"][..., :order])
            assert np.array_equal(got["conf"], want["conf"]) and np.array_equal(got["risky_mask"], want["risky_mask"])
            assert np.array_equal(got["accuracy"], want["ba"], equal_nan=True)
    finally:
        engine.set_eval_function(h.EVAL_BA)
        oracle.set_eval_function(h.EVAL_BA)


def test_confusion_matrix_with_given_risky_cells(engine, oracle):
    """hpgv_epi_confusion = confusion_matrix (model.c:337-460) for the caller's risky cells: feeding back the masks of
    hpgv_epi_eval reproduces its matrices, and arbitrary masks count exactly the samples of the chosen cells."""
    nv, A, U, F = 9, 120, 150, 4
    g = synth.make_dataset(nv, A, U, seed=4, missing=0.03, planted=1)
    fos = np.concatenate([np.arange(A) % F, np.arange(U) % F]).astype(np.int32)
    engine.load_dataset(g, A, U)
    engine.set_folds(F, fos)
    combs = np.array([(i, j) for i in range(nv) for j in range(i + 1, nv)], np.int32)
    for subset in (h.SUBSET_TRAINING, h.SUBSET_TESTING):
        ev = engine.eval(2, combs, subset)
        conf, val = engine.confusion(2, combs, ev["risky_mask"], subset)
        assert np.array_equal(conf, ev["conf"]) and np.array_equal(val, ev["ba"], equal_nan=True)
    mask = np.full((combs.shape[0], F), 0b100010001, np.uint32)          # cells (0,0), (1,1), (2,2)
    conf, _ = engine.confusion(2, combs, mask, h.SUBSET_TESTING)
    for n, (i, j) in enumerate(combs[:6]):
        for f in range(F):
            sel = (fos == f) & (g[i] == g[j]) & (g[i] <= 2)
            tp, fp = int(sel[:A].sum()), int(sel[A:].sum())
            assert conf[n, f].tolist() == [tp, int((fos[:A] == f).sum()) - tp, fp, int((fos[A:] == f).sum()) - fp]


# ---- a16 / a17 on GPU-found rankings ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("idx", range(len(GOLD["merge"])))
def test_gpu_rankings_through_reference_merge_and_report(engine, tmp_path, idx):
    """search on the GPU == the rankings the reference's merge_rankings / epistasis_report were run on (golden `merge`),
    and the host merge + report of those GPU rankings reproduce the reference's rows and .epi bytes."""
    import test_host_api as hostapi
    case = GOLD["merge"][idx]
    order, nv, A, U, F, rank = case["order"], case["nv"], case["A"], case["U"], case["F"], case["rank"]
    g = unb64(case["genotypes"], np.uint8, (nv, A + U))
    fos = unb64(case["fold_of_sample"], np.int32)
    want = unb64(case["models"], h.MODEL_DTYPE, (F, rank))
    engine.load_dataset(g, A, U)
    engine.set_folds(F, fos)
    got = engine.search(order, case["subset"], rank)
    assert np.array_equal(got["snp"][..., :order], want["snp"][..., :order])
    assert np.array_equal(got["accuracy"], want["accuracy"], equal_nan=True) and np.array_equal(got["risky_mask"], want["risky_mask"])
    hostapi.check_against_reference_merge(hostapi.load_host(), tmp_path, case, np.ascontiguousarray(got))


# ---- scale -------------------------------------------------------------------------------------------------------------------
def sample_check(oracle, g, A, U, F, fos, order, got, rank, rng, nsample):
    """Size-independent checks of a search result the CPU cannot reproduce in full: (1) every returned model re-scores to
    itself through the ORACLE; (2) lists are in canonical order; (3) no tuple of a random sample ranks before the last
    model of its fold without being in the list."""
    nv = g.shape[0]
    for f in range(F):
        combs = np.ascontiguousarray(got["snp"][f, :, :order])
        ov = oracle.eval(g, A, U, order, fos, 1, combs)
        assert np.array_equal(ov["risky_mask"][:, f], got["risky_mask"][f])
        assert np.array_equal(ov["conf"][:, f], got["conf"][f])
        assert np.array_equal(ov["ba"][:, f], got["accuracy"][f])
        keys = [(-got["accuracy"][f, r],) + tuple(got["snp"][f, r, :order]) for r in range(rank)]
        assert keys == sorted(keys) and len(set(keys)) == rank
    combs = np.sort(rng.integers(0, nv, (nsample, order)), axis=1).astype(np.int32)
    combs = combs[(np.diff(combs, axis=1) > 0).all(axis=1)]
    ov = oracle.eval(g, A, U, order, fos, 1, combs)
    for f in range(F):
        last = got[f, rank - 1]
        inlist = {tuple(t) for t in got["snp"][f, :, :order]}
        better = np.nonzero(ov["ba"][:, f] >= last["accuracy"])[0]
        for n in better:
            t = tuple(int(x) for x in combs[n])
            if ov["ba"][n, f] > last["accuracy"] or t < tuple(last["snp"][:order]):
                assert t in inlist, (f, t, ov["ba"][n, f], last)
    return combs.shape[0]


def test_c2_full_search_equals_the_exhaustive_reference(engine, ref):
    """BASELINE configs[1] in full: 49 995 000 pairs x 10 folds, the GPU's 10 x 50 models against an exhaustive search
    driven over the REFERENCE's own leaf functions (oracle/_ref/libhpgref.so: set_genotypes_masks ->
    combination_counts_all_folds -> choose_high_risk_combinations2 -> confusion_matrix -> evaluate_model, SSE and all;
    about a minute of CPU on 16 threads -- the scalar restatement would need an hour)."""
    oracle = ref
    nv, A, U, order, F, seed = synth.CONFIGS["c2"]
    if os.environ.get("HPGV_FAST_TESTS") == "1":
        nv = 5000          # a quarter of the pairs (17 s of CPU on the GPU box instead of about 70)
    g = synth.make_dataset(nv, A, U, seed, order=order)
    fos, _ = h.k_folds(A, U, F, 20261017)
    engine.load_dataset(g, A, U)
    engine.set_folds(F, fos)
    got = engine.search(2, h.SUBSET_TRAINING, 50)
    want, _ = oracle.search(g, A, U, 2, fos, 1, 50, threads=os.cpu_count() or 8, num_folds=F)
    assert np.array_equal(got["snp"][..., :2], want["snp"][..., :2])
    assert np.array_equal(got["risky_mask"], want["risky_mask"]) and np.array_equal(got["conf"], want["conf"])
    assert np.array_equal(got["accuracy"], want["ba"])


@pytest.mark.parametrize("name,nv", [("c3", 30000), ("c4", 600), ("c5", 5000)])
def test_benchmark_shapes_sample_checked(engine, oracle, name, nv):
    """c3-, c4- and c5-shaped searches (sample axis of the BASELINE config, fewer SNPs so that the GPU part takes
    seconds): re-scored by the oracle and sample-checked."""
    _, A, U, order, F, seed = synth.CONFIGS[name]
    g = synth.make_dataset(nv, A, U, seed, order=order)
    fos, _ = h.k_folds(A, U, F, 20261017)
    engine.load_dataset(g, A, U)
    engine.set_folds(F, fos)
    got = engine.search(order, h.SUBSET_TRAINING, 50)
    n = sample_check(oracle, g, A, U, F, fos, order, got, 50, np.random.default_rng(nv), {"c3": 6000, "c4": 4000, "c5": 500}[name])
    assert n > 400
