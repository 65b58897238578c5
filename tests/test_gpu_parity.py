"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on the same
seeded inputs.  Bit-exact for counts, risk masks, confusion matrices and selected
models; balanced accuracy to 1e-12 (it is computed from the same integers in
double on both sides, so the tests actually demand equality)."""
import numpy as np
import pytest

import hpg_variant_b200 as h
from hpg_variant_b200 import synth

pytestmark = pytest.mark.gpu

BA_TOL = 1e-12


def random_folds(rng, A, U, F):
    fos = np.concatenate([rng.permutation(A) % F, rng.permutation(U) % F]).astype(np.int32)
    return fos


def all_combs(nv, order):
    import itertools
    return np.array(list(itertools.combinations(range(nv), order)), np.int32)


def check_eval(engine, oracle, g, A, U, F, fos, order, combs, subset):
    engine.load_dataset(g, A, U)
    engine.set_folds(F, fos)
    got = engine.eval(order, combs, subset)
    want = oracle.eval(g, A, U, order, fos, subset, combs)
    for k in ("counts_aff", "counts_unaff", "risky_mask", "conf"):
        assert np.array_equal(got[k], want[k]), k
    assert np.array_equal(np.isnan(got["ba"]), np.isnan(want["ba"]))
    ok = ~np.isnan(want["ba"])
    assert np.max(np.abs(got["ba"][ok] - want["ba"][ok]), initial=0.0) <= BA_TOL
    assert np.array_equal(got["ba"][ok], want["ba"][ok])   # same integers, same double ops
    return got


SHAPES = [
    # nv, A, U, F   (layout exercised)
    (12, 49, 98, 10),      # BW=4, unbalanced (float32 rule slow path), fixture-like sizes
    (10, 64, 64, 4),       # BW=4, balanced, segment = 16
    (9, 700, 700, 5),      # BW=8 single block (140 per segment), u8 counters
    (8, 1300, 900, 3),     # BW=8, 2 blocks per segment (434, 300), u16 counters, unbalanced
    (7, 5000, 5000, 2),    # many blocks per segment, u16
    (6, 33, 17, 2),        # ragged tiny
    (10, 120, 115, 3),     # tri layout, eight blocks, tails empty
    (9, 360, 345, 3),      # BW=4 without the tri layout (115..120 per segment), unbalanced
    (8, 397, 400, 4),      # tri layout with the 4-bit tails in use (99..100 per segment)
]


@pytest.mark.parametrize("nv,A,U,F", SHAPES)
@pytest.mark.parametrize("order", [2, 3])
@pytest.mark.parametrize("subset", [h.SUBSET_TRAINING, h.SUBSET_TESTING])
def test_eval_matches_oracle(engine, oracle, nv, A, U, F, order, subset):
    rng = np.random.default_rng(nv * 1000 + A + F + order)
    g = synth.make_dataset(nv, A, U, seed=A + U + nv, order=order, missing=0.02, planted=1)
    fos = random_folds(rng, A, U, F)
    check_eval(engine, oracle, g, A, U, F, fos, order, all_combs(nv, order), subset)


def compare_models(got, want, order):
    assert got.shape == want.shape
    assert np.array_equal(got["snp"][..., :order], want["snp"][..., :order])
    assert np.array_equal(got["risky_mask"], want["risky_mask"])
    assert np.array_equal(got["conf"], want["conf"])
    gn, wn = np.isnan(got["accuracy"]), np.isnan(want["ba"])
    assert np.array_equal(gn, wn)
    assert np.array_equal(got["accuracy"][~gn], want["ba"][~wn])


SEARCH_SHAPES = [
    # nv, A, U, F, rank
    (40, 49, 98, 10, 50),
    (70, 100, 100, 10, 50),      # BW=4 balanced
    (100, 1000, 1000, 10, 50),   # c2-shaped samples (BW=4, 100 per segment)
    (90, 2000, 2000, 10, 20),    # c3-shaped samples (BW=8 single, u8)
    (50, 1500, 1100, 4, 30),     # u16 multi-block, unbalanced
    (37, 300, 200, 3, 700),      # rank > number of pairs: every pair comes back, sorted
    (60, 720, 720, 3, 30),       # BW=8 single with 240 per segment: all 8 words of a block in use (no 7-word compress)
    (40, 2500, 2500, 2, 30),     # 1250 per segment: 8-word multi-block layout, balanced pre-filter on 16-bit counters
    (64, 90, 90, 5, 40),         # odd fold count: the last byte-counter word holds one real fold and one padding fold
    (60, 360, 360, 3, 40),       # BW=4 proper (120 per segment: too long for the tri layout), balanced
    (50, 230, 460, 4, 40),       # BW=4 proper, unbalanced (57..58 / 115 per segment)
    (40, 130, 130, 13, 30),      # 28 blocks: too many for the tri layout, 4-word blocks with short segments
    (50, 985, 970, 10, 40),      # tri layout, unbalanced, tails partly filled (98..99 / 97 per segment)
    (48, 1200, 1200, 12, 40),    # tri layout at its limits: 24 blocks of exactly 100 samples
]


@pytest.mark.parametrize("nv,A,U,F,rank", SEARCH_SHAPES)
@pytest.mark.parametrize("subset", [h.SUBSET_TRAINING, h.SUBSET_TESTING])
def test_search_order2_matches_oracle(engine, oracle, nv, A, U, F, rank, subset):
    rng = np.random.default_rng(nv + A + F)
    g = synth.make_dataset(nv, A, U, seed=nv * 7 + F, missing=0.01, planted=2)
    fos = random_folds(rng, A, U, F)
    engine.load_dataset(g, A, U)
    engine.set_folds(F, fos)
    got = engine.search(2, subset, rank)
    want, _ = oracle.search(g, A, U, 2, fos, subset, rank, threads=8, num_folds=F)
    compare_models(got, want, 2)


@pytest.mark.parametrize("nv,A,U,F,rank", [(24, 49, 98, 5, 50), (30, 400, 400, 5, 40), (20, 2000, 2000, 5, 25), (18, 700, 500, 3, 1000)])
@pytest.mark.parametrize("subset", [h.SUBSET_TRAINING, h.SUBSET_TESTING])
def test_search_order3_matches_oracle(engine, oracle, nv, A, U, F, rank, subset):
    rng = np.random.default_rng(nv + A + F + 3)
    g = synth.make_dataset(nv, A, U, seed=nv * 11 + F, order=3, missing=0.01, planted=1)
    fos = random_folds(rng, A, U, F)
    engine.load_dataset(g, A, U)
    engine.set_folds(F, fos)
    got = engine.search(3, subset, rank)
    want, _ = oracle.search(g, A, U, 3, fos, subset, rank, threads=8, num_folds=F)
    compare_models(got, want, 3)


@pytest.mark.parametrize("nv,A,U,F,rank,missing", [
    (40, 2000, 2000, 5, 30, 0.005),     # c4's sample axis: two 7-word blocks per segment, 16-bit counters
    (36, 700, 700, 5, 30, 0.02),        # one 7-word block per segment, byte counters
    (30, 400, 400, 5, 40, 0.02),        # 4-word blocks (80 per segment), byte counters
    (28, 1300, 900, 3, 25, 0.01),       # unbalanced: the float32 risk rule, 16-bit counters
    (26, 2600, 2600, 2, 20, 0.03),      # 1300 per segment: full 8-word blocks, long lists of missing samples
    (24, 240, 240, 3, 3000, 0.01),      # rank > number of triples: lists in global memory, every triple comes back
])
def test_search_order3_resident_tiles(engine, oracle, monkeypatch, nv, A, U, F, rank, missing):
    """search3v2_kernel ((j, k) tiles resident, genotype 2 of SNP i derived from the pair table, missing samples fixed up
    one by one) against the oracle and against the plain order-3 kernel, whole range and a sub-range, both subsets."""
    rng = np.random.default_rng(nv + A)
    g = synth.make_dataset(nv, A, U, seed=nv * 13 + F, order=3, missing=missing, planted=1)
    g[3, rng.integers(0, A + U, (A + U) // 8)] = 255          # one SNP with many missing samples
    fos = random_folds(rng, A, U, F)
    total = h.num_combinations(nv, 3)
    for subset in (h.SUBSET_TRAINING, h.SUBSET_TESTING):
        monkeypatch.delenv("HPGV_SEARCH3_V2", raising=False)
        engine.load_dataset(g, A, U)
        engine.set_folds(F, fos)
        got = engine.search(3, subset, rank)
        part = engine.search(3, subset, rank, total // 7 + 3, total // 2 + 5)
        want, _ = oracle.search(g, A, U, 3, fos, subset, rank, threads=8, num_folds=F)
        compare_models(got, want, 3)
        wantp, _ = oracle.search(g, A, U, 3, fos, subset, rank, first=total // 7 + 3, last=total // 2 + 5, threads=8, num_folds=F)
        compare_models(part, wantp, 3)
        monkeypatch.setenv("HPGV_SEARCH3_V2", "0")
        engine.set_folds(F, fos)
        plain = engine.search(3, subset, rank)
        assert plain.tobytes() == got.tobytes()
        monkeypatch.delenv("HPGV_SEARCH3_V2", raising=False)
        monkeypatch.setenv("HPGV_SEARCH3_V3", "1")            # the one-thread-per-triple kernel (opt-in) where it applies
        engine.set_folds(F, fos)
        one = engine.search(3, subset, rank)
        monkeypatch.delenv("HPGV_SEARCH3_V3", raising=False)
        assert one.tobytes() == got.tobytes()


@pytest.mark.parametrize("nv,A,F", [(34, 1000, 10), (30, 2400, 10), (32, 200, 7)])     # (10 folds of 16-bit counters: outside what v3 was built for)
def test_search_order3_one_thread_kernel_fold_counts(engine, oracle, monkeypatch, nv, A, F):
    """search3v3_kernel (HPGV_SEARCH3_V3=1) with other counter shapes than c4's: 10 folds of byte counters (5 words per cell), 7
    folds of 4-word blocks (odd fold count); 10 folds of 16-bit counters fall back to the default kernel, which is also run
    on all three shapes."""
    g = synth.make_dataset(nv, A, A, seed=nv + F, order=3, missing=0.01, planted=1)
    fos, _ = h.k_folds(A, A, F, seed=8)
    want, _ = oracle.search(g, A, A, 3, fos, h.SUBSET_TRAINING, 30, threads=8, num_folds=F)
    for v3 in ("1", "0"):
        monkeypatch.setenv("HPGV_SEARCH3_V3", v3)
        engine.load_dataset(g, A, A)
        engine.set_folds(F, fos)
        got = engine.search(3, h.SUBSET_TRAINING, 30)
        compare_models(got, want, 3)


def test_search_order3_resident_tiles_many_units(engine, monkeypatch):
    """enough SNPs for several units per CTA and long i loops: bytes equal to the plain kernel's (c4's sample axis)."""
    nv, A, F, rank = 300, 2000, 5, 50
    g = synth.make_dataset(nv, A, A, seed=77, order=3, missing=0.005, planted=2)
    fos, _ = h.k_folds(A, A, F, seed=4)
    total = h.num_combinations(nv, 3)
    engine.load_dataset(g, A, A)
    engine.set_folds(F, fos)
    got = [engine.search(3, h.SUBSET_TRAINING, rank), engine.search(3, h.SUBSET_TRAINING, rank, total // 3, total // 3 * 2 + 11)]
    for switch, val in (("HPGV_SEARCH3_V2", "0"), ("HPGV_SEARCH3_V3", "1")):
        monkeypatch.setenv(switch, val)
        engine.set_folds(F, fos)
        other = [engine.search(3, h.SUBSET_TRAINING, rank), engine.search(3, h.SUBSET_TRAINING, rank, total // 3, total // 3 * 2 + 11)]
        monkeypatch.delenv(switch, raising=False)
        for a, b in zip(got, other):
            assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("order,nv,A,F", [(2, 80, 600, 6), (3, 22, 240, 4)])
def test_search_balanced_classes_uneven_folds(engine, oracle, order, nv, A, F):
    """A == U but folds that do not hold as many cases as controls: the balanced pre-filter must stand aside."""
    rng = np.random.default_rng(17 + order)
    g = synth.make_dataset(nv, A, A, seed=23 + order, order=order, missing=0.01, planted=1)
    fos = np.concatenate([rng.permutation(A) % F, rng.integers(0, F, A)]).astype(np.int32)
    engine.load_dataset(g, A, A)
    engine.set_folds(F, fos)
    for subset in (h.SUBSET_TRAINING, h.SUBSET_TESTING):
        got = engine.search(order, subset, 30)
        want, _ = oracle.search(g, A, A, order, fos, subset, 30, threads=8, num_folds=F)
        compare_models(got, want, order)


@pytest.mark.parametrize("nv,A,F,rank,subset", [
    (700, 100, 2, 50, h.SUBSET_TRAINING),     # tri layout, ~480 units: every CTA walks several units (score histogram, re-run of the first unit)
    (500, 200, 4, 50, h.SUBSET_TRAINING),     # tri layout, four folds
    (400, 720, 3, 30, h.SUBSET_TRAINING),     # 8-word single-block segments
    (500, 100, 2, 50, h.SUBSET_TESTING),      # the testing part (no pre-filter, no histogram)
])
def test_search_order2_many_units(engine, oracle, nv, A, F, rank, subset):
    """Searches large enough for the steady state of the persistent kernel: thresholds derived from the global score
    histogram, lists that stay short, first unit counted first and offered last."""
    g = synth.make_dataset(nv, A, A, seed=nv + F, missing=0.01, planted=3)
    fos, _ = h.k_folds(A, A, F, seed=5)
    engine.load_dataset(g, A, A)
    engine.set_folds(F, fos)
    got = engine.search(2, subset, rank)
    want, _ = oracle.search(g, A, A, 2, fos, subset, rank, threads=16, num_folds=F)
    compare_models(got, want, 2)
    # a sub-range that starts and ends inside tiles
    total = h.num_combinations(nv, 2)
    lo, hi = total // 5 + 7, total // 3 + 13
    got = engine.search(2, subset, rank, lo, hi)
    want, _ = oracle.search(g, A, A, 2, fos, subset, rank, first=lo, last=hi, threads=16, num_folds=F)
    compare_models(got, want, 2)


@pytest.mark.parametrize("nv,A,F", [(3000, 1000, 10), (1500, 2000, 10), (600, 25000, 10)])
def test_histogram_thresholds_do_not_change_results(engine, monkeypatch, nv, A, F):
    """c2-, c3- and c5-shaped samples at sizes the CPU oracle cannot reach: the histogram-bounded search returns the
    bytes of the plain one, whole range and as four sub-ranges merged on the device (size-independent properties)."""
    import torch
    rank = 50
    g = synth.make_dataset(nv, A, A, seed=4242 + nv, missing=0.005, planted=4)
    fos, _ = h.k_folds(A, A, F, seed=6)
    engine.load_dataset(g, A, A)
    engine.set_folds(F, fos)
    with_hist = engine.search(2, h.SUBSET_TRAINING, rank)
    total = h.num_combinations(nv, 2)
    cuts = [0, total // 9, total // 3 + 5, total // 2, total]
    parts = [engine.search(2, h.SUBSET_TRAINING, rank, cuts[i], cuts[i + 1]) for i in range(4)]
    lists = torch.from_numpy(np.stack(parts).view(np.uint8)).cuda()
    out = torch.zeros(F * rank * 40, dtype=torch.uint8, device="cuda")
    engine.merge_device(2, h.SUBSET_TRAINING, 4, rank, lists.data_ptr(), out.data_ptr())
    torch.cuda.synchronize()
    assert out.cpu().numpy().tobytes() == with_hist.tobytes()
    monkeypatch.setenv("HPGV_HIST", "0")
    without = engine.search(2, h.SUBSET_TRAINING, rank)
    assert with_hist.tobytes() == without.tobytes()
    # every returned model re-evaluates to itself through the per-combination hook
    ev = engine.eval(2, with_hist["snp"][0, :8, :2].copy(), h.SUBSET_TRAINING)
    assert np.array_equal(ev["risky_mask"][:, 0], with_hist["risky_mask"][0, :8])
    assert np.array_equal(ev["conf"][:, 0], with_hist["conf"][0, :8])


@pytest.mark.parametrize("switch", ["HPGV_UNIT_DESC=0", "HPGV_TRI_DERIVE=0", "HPGV_TRI_WARPS=16", "HPGV_STAGES=2", "HPGV_STAGGER=0", "HPGV_PACK_WARP=1", "HPGV_LIST_SCAN=0",
                                    "HPGV_MISS_LIST=0", "HPGV_FIRST_WAIT=0", "HPGV_FRESH_BOUND=1", "HPGV_BAND_TILES=5"])
def test_development_switches_do_not_change_results(engine, monkeypatch, switch):
    """Every A/B switch of the library selects another route to the same bytes (prefix-table walk instead of the per-unit
    descriptors, every cell counted directly, 16 warps, two stages, no stagger, the one-warp-per-word packer, heap-ordered
    instead of array lists, marginals without the lists of missing samples, no wait for the first units' histogram counts,
    other CTAs' bounds adopted every unit, units listed band by band -- the order of data sets larger than the L2)."""
    nv, A, F, rank = 1200, 2000, 10, 40           # c3-shaped samples: 8-word single-block layout with marginals
    g = synth.make_dataset(nv, A, A, seed=99, missing=0.01, planted=2)
    fos, _ = h.k_folds(A, A, F, seed=3)
    total = h.num_combinations(nv, 2)
    engine.load_dataset(g, A, A)
    engine.set_folds(F, fos)
    want = [engine.search(2, h.SUBSET_TRAINING, rank), engine.search(2, h.SUBSET_TESTING, rank, total // 3, total)]
    k, v = switch.split("=")
    monkeypatch.setenv(k, v)
    engine.set_folds(F, fos)
    got = [engine.search(2, h.SUBSET_TRAINING, rank), engine.search(2, h.SUBSET_TESTING, rank, total // 3, total)]
    for a, b in zip(want, got):
        assert a.tobytes() == b.tobytes()


def test_tri_layout_equals_four_word_layout(engine, monkeypatch):
    """The tri layout (3 words + shared 4-bit tails, 2 POPC per block) and the plain 4-word layout give the same bytes."""
    nv, A, U, F, rank = 150, 1000, 1000, 10, 50
    g = synth.make_dataset(nv, A, U, seed=77, missing=0.01, planted=2)
    fos = random_folds(np.random.default_rng(9), A, U, F)
    engine.load_dataset(g, A, U)
    engine.set_folds(F, fos)
    assert engine.layout()["block_words"] == 3
    tri = [engine.search(2, s, rank) for s in (h.SUBSET_TRAINING, h.SUBSET_TESTING)]
    o3 = engine.search(3, h.SUBSET_TRAINING, 10, 0, 20000)        # order 3 re-packs with 4-word blocks ...
    assert engine.layout()["block_words"] == 4
    again = engine.search(2, h.SUBSET_TRAINING, rank)             # ... and order 2 goes back to the tri layout
    assert engine.layout()["block_words"] == 3
    assert again.tobytes() == tri[0].tobytes()
    monkeypatch.setenv("HPGV_NO_TRI", "1")
    engine.set_folds(F, fos)
    assert engine.layout()["block_words"] == 4
    plain = [engine.search(2, s, rank) for s in (h.SUBSET_TRAINING, h.SUBSET_TESTING)]
    o3b = engine.search(3, h.SUBSET_TRAINING, 10, 0, 20000)
    for a, b in zip(tri, plain):
        assert a.tobytes() == b.tobytes()
    assert o3.tobytes() == o3b.tobytes()


def test_search_ranges_partition(engine, oracle):
    """Contiguous index ranges (the multi-GPU sharding) merged on the GPU == one full search."""
    import torch
    nv, A, U, F, rank = 120, 500, 500, 5, 40
    g = synth.make_dataset(nv, A, U, seed=5)
    fos = random_folds(np.random.default_rng(1), A, U, F)
    engine.load_dataset(g, A, U)
    engine.set_folds(F, fos)
    full = engine.search(2, h.SUBSET_TRAINING, rank)
    total = h.num_combinations(nv, 2)
    cuts = [0, total // 7, total // 3, total // 2 + 11, total]
    parts = [engine.search(2, h.SUBSET_TRAINING, rank, cuts[i], cuts[i + 1]) for i in range(4)]
    lists = torch.from_numpy(np.stack(parts).view(np.uint8)).cuda()
    out = torch.zeros(F * rank * 40, dtype=torch.uint8, device="cuda")
    engine.merge_device(2, h.SUBSET_TRAINING, 4, rank, lists.data_ptr(), out.data_ptr())
    torch.cuda.synchronize()
    merged = out.cpu().numpy().view(h.MODEL_DTYPE).reshape(F, rank)
    assert merged.tobytes() == full.tobytes()
    for p, (lo, hi) in zip(parts, zip(cuts[:-1], cuts[1:])):
        want, _ = oracle.search(g, A, U, 2, fos, h.SUBSET_TRAINING, rank, first=lo, last=hi, threads=4, num_folds=F)
        compare_models(p, want, 2)


@pytest.mark.parametrize("nv,A,U,F", [(5, 77, 130, 7), (4, 400, 395, 4), (3, 240, 230, 2), (3, 900, 2600, 3)])
def test_packer_roundtrip(engine, oracle, nv, A, U, F):
    """unpack(pack(bytes)) == set_genotypes_masks of the reference (model.c:28-74), folds shuffled.
    Layouts: tri without / with tails, 4-word blocks, 8-word multi-block."""
    g = synth.make_dataset(nv, A, U, seed=3, missing=0.05, planted=0)
    fos = random_folds(np.random.default_rng(2), A, U, F)
    engine.load_dataset(g, A, U)
    engine.set_folds(F, fos)
    a_pad = 16 * ((A + 15) // 16)
    for v in range(nv):
        m = engine.unpack_masks(v)
        for gt in range(3):
            want = np.zeros(m.shape[1], np.uint8)
            want[:A] = np.where(g[v, :A] == gt, 255, 0)
            want[a_pad:a_pad + U] = np.where(g[v, A:] == gt, 255, 0)
            assert np.array_equal(m[gt], want)


@pytest.mark.parametrize("order", [2, 3])
def test_edge_cases_empty_range_tiny_cohorts_and_errors(engine, oracle, order):
    """Empty and one-combination ranges, the smallest cohorts the API accepts (folds without cases or controls: 0/0 = NaN like
    the reference, NaN ranks last), a caller-owned device buffer that ends right after the matrix, and the argument errors."""
    import torch
    # (1) empty range and a range of one combination
    nv, A, U, F, rank = 12, 40, 44, 4, 6
    g = synth.make_dataset(nv, A, U, seed=order, order=order, missing=0.02, planted=1)
    fos = random_folds(np.random.default_rng(3), A, U, F)
    engine.load_dataset(g, A, U)
    engine.set_folds(F, fos)
    total = h.num_combinations(nv, order)
    empty = engine.search(order, h.SUBSET_TRAINING, rank, 5, 5)
    assert (empty["snp"] == -1).all() and np.isnan(empty["accuracy"]).all() and (empty["conf"] == 0).all()
    one = engine.search(order, h.SUBSET_TRAINING, rank, total - 1, total)
    want, _ = oracle.search(g, A, U, order, fos, 1, rank, first=total - 1, last=total, threads=1, num_folds=F)
    compare_models(one, want, order)
    assert (one["snp"][:, 0, :order] == np.arange(nv - order, nv)).all() and (one["snp"][:, 1:, 0] == -1).all()
    # (2) tiny cohorts: 3 cases, 2 controls, 3 folds -> fold 2 holds a case and no control (testing part: TN + FP = 0 -> NaN)
    nv, A, U, F = 5, 3, 2, 3
    g = np.array([[0, 1, 2, 0, 1], [1, 1, 0, 2, 255], [2, 0, 0, 1, 1], [0, 0, 1, 1, 2], [255, 2, 1, 0, 0]], np.uint8)
    fos = np.array([0, 1, 2, 0, 1], np.int32)
    # the bytes live in a device buffer that ends with the matrix (pack_rows_kernel must not read 16-byte groups past it)
    d = torch.from_numpy(g.reshape(-1).copy()).cuda()
    engine.load_dataset_device(d.data_ptr(), nv, A, U)
    engine.set_folds(F, fos)
    for subset in (h.SUBSET_TRAINING, h.SUBSET_TESTING):
        got = engine.search(order, subset, 4)
        want, _ = oracle.search(g, A, U, order, fos, subset, 4, threads=1, num_folds=F)
        compare_models(got, want, order)
    combs = np.array(list(__import__("itertools").combinations(range(nv), order)), np.int32)
    ev, ov = engine.eval(order, combs, h.SUBSET_TESTING), oracle.eval(g, A, U, order, fos, 0, combs)
    assert np.array_equal(ev["conf"], ov["conf"]) and np.array_equal(ev["ba"], ov["ba"], equal_nan=True) and np.isnan(ev["ba"][:, 2]).all()
    # (3) argument errors come back as codes, never as a crash or a CPU path
    for bad in (lambda: engine.search(4, h.SUBSET_TRAINING, 4), lambda: engine.search(order, 7, 4), lambda: engine.search(order, h.SUBSET_TRAINING, 0),
                lambda: engine.search(order, h.SUBSET_TRAINING, 5000), lambda: engine.eval(order, np.array([[3, 1, 0][:order]], np.int32)),
                lambda: engine.set_folds(1, np.zeros(A + U, np.int32)), lambda: engine.set_folds(F, np.full(A + U, F, np.int32)),
                lambda: engine.set_eval_function(h.EVAL_WBA), lambda: engine.set_eval_function(9)):
        with pytest.raises(h.HpgvError):
            bad()
    engine.set_folds(F, fos)                              # the context is still usable afterwards
    assert engine.search(order, h.SUBSET_TRAINING, 4).shape == (F, 4)


@pytest.mark.parametrize("order,nv,A,U,F,rank", [(2, 420, 300, 200, 3, 300), (2, 150, 500, 500, 5, 3000), (3, 60, 240, 200, 3, 300)])
def test_merge_long_lists_take_the_selection_path(engine, oracle, order, nv, A, U, F, rank):
    """More than 8192 valid list entries per fold (long rankings, every CTA's list full): merge_kernel leaves the
    shared-memory sort for the radix select over global memory.  Whole search against the oracle, and four sub-range
    rankings merged on the device (the multi-GPU merge at a long rank size) against the whole search."""
    import torch
    g = synth.make_dataset(nv, A, U, seed=nv + rank, order=order, missing=0.01, planted=2)
    fos = random_folds(np.random.default_rng(rank), A, U, F)
    engine.load_dataset(g, A, U)
    engine.set_folds(F, fos)
    total = h.num_combinations(nv, order)
    assert min(total, 148 * rank) > 8192
    full = engine.search(order, h.SUBSET_TRAINING, rank)
    want, _ = oracle.search(g, A, U, order, fos, h.SUBSET_TRAINING, rank, threads=8, num_folds=F)
    compare_models(full, want, order)
    cuts = [0, total // 5, total // 2, total - total // 7, total]
    parts = [engine.search(order, h.SUBSET_TRAINING, rank, cuts[i], cuts[i + 1]) for i in range(4)]
    lists = torch.from_numpy(np.stack(parts).view(np.uint8)).cuda()
    out = torch.zeros(F * rank * 40, dtype=torch.uint8, device="cuda")
    engine.merge_device(order, h.SUBSET_TRAINING, 4, rank, lists.data_ptr(), out.data_ptr())
    torch.cuda.synchronize()
    merged = out.cpu().numpy().view(h.MODEL_DTYPE).reshape(F, rank)
    assert merged.tobytes() == full.tobytes()
