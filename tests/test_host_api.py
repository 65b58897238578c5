"""The reference's C host API re-provided by libhpgv_epi_host.so (include/hpgv_epi_compat.h).

CPU part: symbols, option verification codes, dataset headers, enumerators, fold masks, the
CV-C / CV-A merge and the report text.  GPU part (marked): `hpg-var-gwas-b200 epi` end to end
against the oracle's per-fold rankings pushed through an independent Python restatement of
merge_rankings / epistasis_report (epistasis.c:96-153, epistasis_report.c:28-82)."""
import ctypes as C
import math
import os
import struct
import subprocess

import numpy as np
import pytest

import hpg_variant_b200 as h
from hpg_variant_b200 import synth
from hpg_variant_b200 import build as hbuild
from golden_util import load_golden

GOLD = load_golden()

PKG = os.path.dirname(os.path.abspath(h.__file__))
HOSTLIB = os.path.join(PKG, "libhpgv_epi_host.so")
CLI = os.path.join(PKG, "hpg-var-gwas-b200")

COMPAT_SYMBOLS = [
    "run_epistasis", "epistasis", "epistasis_dataset_load", "epistasis_dataset_close", "get_block_stride", "get_next_block",
    "get_first_combination_in_block", "get_next_combination_in_block", "get_genotype_combinations",
    "get_next_genotype_combination", "get_k_folds", "get_k_folds_masks", "hpgv_epi_merge_rankings", "hpgv_epi_write_report",
    "hpgv_epi_host_open_log",
    # leaf functions (model.h:91-155, mdr.h:37-39, cross_validation.h:14-23) and the producer side of the file format
    "masks_info_init", "set_genotypes_masks", "combination_counts", "combination_counts_all_folds", "mdr_high_risk_combinations2",
    "choose_high_risk_combinations2", "risky_combination_new", "risky_combination_free", "confusion_matrix", "evaluate_model",
    "test_model", "get_genotypes_of_block_coord", "epistasis_dataset_write", "epistasis_dataset_encode_genotype",
    "group_individuals_by_phenotype",
]


class ReportRow(C.Structure):
    _fields_ = [("cv_accuracy", C.c_double), ("cv_count", C.c_int), ("order", C.c_int), ("snp", C.c_int * 3),
                ("num_risky", C.c_int), ("risky_genotypes", (C.c_uint8 * 3) * 27)]


def load_host():
    if not os.path.exists(HOSTLIB):
        hbuild.build()
    lib = C.CDLL(HOSTLIB)
    lib.epistasis_dataset_load.restype = C.POINTER(C.c_uint8)
    lib.epistasis_dataset_load.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                           C.POINTER(C.c_size_t), C.c_char_p]
    lib.epistasis_dataset_close.argtypes = [C.POINTER(C.c_uint8), C.c_size_t]
    lib.get_block_stride.argtypes = [C.c_size_t, C.c_int]
    lib.get_k_folds.restype = C.POINTER(C.POINTER(C.c_int))
    lib.get_k_folds.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.POINTER(C.POINTER(C.c_uint))]
    lib.get_k_folds_masks.restype = C.POINTER(C.c_uint8)
    lib.get_k_folds_masks.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.c_uint)]
    lib.get_genotype_combinations.restype = C.POINTER(C.POINTER(C.c_uint8))
    lib.hpgv_epi_merge_rankings.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(ReportRow), C.c_int]
    lib.epistasis.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_char_p]
    return lib


@pytest.fixture(scope="module")
def host():
    return load_host()


def test_host_library_exports_the_reference_api(host):
    for name in COMPAT_SYMBOLS:
        assert hasattr(host, name), name
    hdr = open(os.path.join(os.path.dirname(PKG), "include", "hpgv_epi_compat.h")).read()
    for name in COMPAT_SYMBOLS:
        assert name + "(" in hdr, f"{name} is exported but not declared in hpgv_epi_compat.h"


def call_epistasis(host, args, config=None):
    argv = (C.c_char_p * (len(args) + 1))(b"epi", *[a.encode() for a in args])
    return host.epistasis(len(args) + 1, argv, config.encode() if config else None)


def test_option_verification_codes(host, tmp_path):
    """error codes of verify_epistasis_options (epistasis_options_parsing.c:143-182, src/error.h:44-50)"""
    conf = tmp_path / "hpg-variant.conf"
    conf.write_text('gwas:\n{\n  epistasis:\n  {\n    stride = 0 ;\n    num-folds = 0 ;\n    num-cv-repetitions = 0 ;\n'
                    '    max-ranking-size = 50 ;\n    evaluation-subset = "bogus" ;\n    evaluation-mode = "count" ;\n    num-threads = 4 ;\n  };\n};\n')
    c = str(conf)
    assert call_epistasis(host, ["--order", "2"], c) == 210
    assert call_epistasis(host, ["-d", "x.bin"], c) == 211
    assert call_epistasis(host, ["-d", "x.bin", "--order", "2"], c) == 212
    assert call_epistasis(host, ["-d", "x.bin", "--order", "2", "--num-folds", "10"], c) == 213
    assert call_epistasis(host, ["-d", "x.bin", "--order", "2", "--num-folds", "10", "--num-cv-runs", "1"], c) == 214
    assert call_epistasis(host, ["-d", "x.bin", "--order=2", "--num-folds=10", "--num-cv-runs=1", "--eval-subset=training"], c) == 216
    bad = tmp_path / "bad.conf"
    bad.write_text("gwas: { epistasis: { stride = ; ")
    assert call_epistasis(host, ["-d", "x.bin", "--order", "2"], str(bad)) == 2
    assert call_epistasis(host, ["--help"]) == 0


def test_reference_config_file_is_understood(host, tmp_path):
    """the shipped etc/hpg-variant/hpg-variant.conf layout: only the dataset and the order are missing"""
    conf = tmp_path / "hpg-variant.conf"
    conf.write_text('# comment\ngwas:\n{\n    assoc:\n    {\n        num-threads = 4 ;\n    };\n    epistasis:\n    {\n'
                    '        stride                  = 100 ;\n        num-folds               = 10 ;\n'
                    '        num-cv-repetitions      = 10 ;\n        max-ranking-size        = 50 ;\n'
                    '        evaluation-subset       = "training" ;\n        evaluation-mode         = "count" ;\n'
                    '        num-threads             = 4 ;\n    };\n};\n')
    assert call_epistasis(host, ["--num-folds", "5"], str(conf)) == 210
    assert call_epistasis(host, ["-d", "x.bin"], str(conf)) == 211


@pytest.mark.parametrize("legacy", [False, True])
def test_dataset_load_both_headers(host, tmp_path, legacy):
    g = synth.make_dataset(7, 9, 12, seed=1, planted=0)
    path = tmp_path / "d.bin"
    if legacy:   # test/random_dataset_gen.c:46-49: size_t + 2 x uint32, a few trailing bytes
        path.write_bytes(struct.pack("<QII", 7, 9, 12) + g.tobytes() + b"\0\0\0\0")
    else:
        synth.write_dataset(str(path), g, 9, 12)
    a, u, nv, flen, off = C.c_int(), C.c_int(), C.c_size_t(), C.c_size_t(), C.c_size_t()
    p = host.epistasis_dataset_load(C.byref(a), C.byref(u), C.byref(nv), C.byref(flen), C.byref(off), str(path).encode())
    assert bool(p)
    assert (nv.value, a.value, u.value, off.value) == (7, 9, 12, 16 if legacy else 12)
    got = np.ctypeslib.as_array(p, shape=(flen.value,))[off.value:off.value + 7 * 21].reshape(7, 21)
    assert np.array_equal(got, g)
    assert host.epistasis_dataset_close(p, flen.value) == 0
    assert not host.epistasis_dataset_load(C.byref(a), C.byref(u), C.byref(nv), C.byref(flen), C.byref(off), b"/nonexistent/file")


def enumerate_blocked(host, nv, order, stride):
    """the runner's enumeration (singlenode/epistasis_runner.c:114-258) on the host library's enumerators"""
    nb = math.ceil(nv / stride)
    out = []
    block = (C.c_int * order)(*([0] * order))
    while True:
        comb = (C.c_int * order)()
        host.get_first_combination_in_block(order, comb, block, stride)
        while True:
            out.append(tuple(comb))
            if not host.get_next_combination_in_block(order, comb, block, stride, nv):
                break
        if not host.get_next_block(nb, order, block):
            break
    return out


@pytest.mark.parametrize("rec", GOLD["blocked"], ids=lambda r: f'{r["nv"]}-{r["order"]}-{r["stride"]}')
def test_enumerators_match_the_reference(host, rec):
    got = enumerate_blocked(host, rec["nv"], rec["order"], rec["stride"])
    assert [list(t) for t in got] == rec["combs"]


def test_enumerator_vectors(host):
    """test/test_epistasis_dataset.c:198-209 (stride) and :211-268 (next block)"""
    assert host.get_block_stride(64, 2) == 8 and host.get_block_stride(1000, 3) == 10 and host.get_block_stride(65, 2) == 9
    b = (C.c_int * 3)(0, 1, 3)
    assert host.get_next_block(4, 3, b) == 1 and list(b) == [0, 2, 2]
    b = (C.c_int * 2)(3, 3)
    assert host.get_next_block(4, 2, b) == 0
    n = C.c_int()
    cells = host.get_genotype_combinations(2, C.byref(n))
    assert n.value == 9 and [(cells[c][0], cells[c][1]) for c in range(9)] == [(a, b) for a in range(3) for b in range(3)]
    cells = host.get_genotype_combinations(3, C.byref(n))
    assert n.value == 27 and (cells[26][0], cells[26][1], cells[26][2]) == (2, 2, 2) and (cells[5][0], cells[5][1], cells[5][2]) == (0, 1, 2)


def test_folds_and_masks(host, oracle, monkeypatch):
    monkeypatch.setenv("HPGV_EPI_SEED", "77")
    A, U, k = 49, 98, 10
    sizes = C.POINTER(C.c_uint)()
    folds = host.get_k_folds(A, U, k, C.byref(sizes))
    fos, want_sizes = h.k_folds(A, U, k, 77)
    got_fos = np.full(A + U, -1, np.int32)
    for f in range(k):
        assert [sizes[3 * f + x] for x in range(3)] == list(want_sizes[f])
        ids = [folds[f][x] for x in range(sizes[3 * f])]
        assert ids == sorted(ids)
        got_fos[ids] = f
    assert np.array_equal(got_fos, fos)
    # fold sizes of test/test_cross_validation.c:73-76: fold f holds ceil((A - f)/k) cases and ceil((U - f)/k) controls
    for f in range(k):
        assert sizes[3 * f + 1] == math.ceil((A - f) / k) and sizes[3 * f + 2] == math.ceil((U - f) / k)
    masks = host.get_k_folds_masks(A, U, k, folds, sizes)
    s_pad = 16 * math.ceil(A / 16) + 16 * math.ceil(U / 16)
    got = np.ctypeslib.as_array(masks, shape=(k * s_pad,)).reshape(k, s_pad)
    assert np.array_equal(got, oracle.fold_masks(A, U, k, fos))


def test_dataset_writer_and_producer_helpers(host, tmp_path):
    """SURVEY 8(f)2, the producer side (vcf-tools/vcf2epi/dataset_creator.c:172-223, 255-265, 302-320): the C writer's
    file is what the loaders read back; genotype bytes and the cases-first column order follow the reference."""
    rng = np.random.default_rng(3)
    nv, A, U = 7, 5, 9
    g = rng.choice(np.array([0, 1, 2, 255], np.uint8), size=(nv, A + U))
    path = tmp_path / "w.bin"
    host.epistasis_dataset_write.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
    assert host.epistasis_dataset_write(str(path).encode(), g.ctypes.data, nv, A, U) == 0
    raw = path.read_bytes()
    assert struct.unpack("<III", raw[:12]) == (nv, A, U) and raw[12:] == g.tobytes()
    g2, a2, u2 = synth.read_dataset(str(path))
    assert (a2, u2) == (A, U) and np.array_equal(g2, g)
    na, nu, nvar, flen, off = C.c_int(), C.c_int(), C.c_size_t(), C.c_size_t(), C.c_size_t()
    p = host.epistasis_dataset_load(C.byref(na), C.byref(nu), C.byref(nvar), C.byref(flen), C.byref(off), str(path).encode())
    assert (na.value, nu.value, nvar.value, off.value) == (A, U, nv, 12)
    assert bytes(p[off.value:off.value + nv * (A + U)]) == g.tobytes()
    host.epistasis_dataset_close(p, flen.value)
    assert host.epistasis_dataset_write(str(tmp_path / "no" / "dir.bin").encode(), g.ctypes.data, nv, A, U) == -1
    host.epistasis_dataset_encode_genotype.restype = C.c_uint8
    enc = host.epistasis_dataset_encode_genotype
    assert [enc(0, 0, 0), enc(0, 1, 0), enc(1, 0, 0), enc(1, 1, 0), enc(2, 2, 0), enc(1, 2, 0), enc(0, 0, 1)] == [0, 1, 1, 2, 2, 1, 255]
    host.group_individuals_by_phenotype.restype = C.POINTER(C.c_int)
    ph = np.array([0, 1, 1, 0, 0, 1, 0], np.uint8)
    dest = host.group_individuals_by_phenotype(ph.ctypes.data_as(C.POINTER(C.c_uint8)), 3, 4)
    assert [dest[x] for x in range(7)] == [3, 0, 1, 4, 5, 2, 6]


# ---- merge_rankings + report ---------------------------------------------------------------------------
def py_merge(models, order, num_folds, mode):
    """independent restatement of merge_rankings + the report order"""
    acc = {}
    for f in range(models.shape[0]):
        for m in models[f]:
            if m["snp"][0] < 0:
                continue
            key = tuple(int(x) for x in m["snp"][:order])
            if key not in acc:
                cells = [c for c in range(3 ** order) if (int(m["risky_mask"]) >> c) & 1]
                acc[key] = dict(sum=0.0, count=0, cells=cells)
            acc[key]["sum"] += float(m["accuracy"])
            acc[key]["count"] += 1
    rows = [(k, v["sum"] / num_folds, v["count"], v["cells"]) for k, v in acc.items()]
    if mode == "count":
        rows.sort(key=lambda r: (-r[2], -r[1], r[0]))
    else:
        rows.sort(key=lambda r: (-r[1], r[0]))
    return rows


def py_report(rows, order, rep, mode, subset, max_rank):
    out = [f"#CROSS VALIDATION {rep + 1}", f"#COMBINATIONS OF: {order} SNPs",
           "#EVALUATION MODE: Cross-validation consistency" if mode == "count" else "#EVALUATION MODE: Cross-validation accuracy",
           "#EVALUATION PARTITION: Training" if subset == "training" else "#EVALUATION PARTITION: Testing",
           "#POSITION\tSNPs\tGENOTYPES\tCV-C\tCV-A"]
    for pos, (key, cva, cvc, cells) in enumerate(rows[:max_rank]):
        snps = "(" + "".join(f" {s}," for s in key[:-1]) + f" {key[-1]} )"
        gts = ""
        for c in cells:
            g = [(c // 3 ** (order - 1 - p)) % 3 for p in range(order)]
            gts += f"({g[0]}-" + "".join(f"{x}, " for x in g[1:-1]) + f"{g[-1]}), "
        out.append(f"{pos + 1}\t{snps}\t{gts}{cvc}\t{cva:.3f}")
    return "\n".join(out) + "\n"


def random_models(rng, order, F, rank, nv):
    m = np.zeros((F, rank), h.MODEL_DTYPE)
    m["snp"][:] = -1
    m["accuracy"][:] = np.nan
    pool = [tuple(sorted(rng.choice(nv, order, replace=False))) for _ in range(rank + 6)]
    for f in range(F):
        picks = rng.permutation(len(pool))[: rank - (f % 3)]       # ragged: some folds return fewer models
        for r, p in enumerate(picks):
            m["snp"][f, r, :order] = pool[p]
            m["accuracy"][f, r] = rng.integers(500, 560) / 1000.0       # plenty of equal accuracies
            m["risky_mask"][f, r] = rng.integers(1, 1 << (3 ** order))
    return m


@pytest.mark.parametrize("order", [2, 3])
@pytest.mark.parametrize("mode", ["count", "accu"])
def test_merge_rankings_and_report(host, tmp_path, order, mode):
    rng = np.random.default_rng(order * 10 + len(mode))
    F, rank = 7, 12
    models = random_models(rng, order, F, rank, 40)
    rows = (ReportRow * (F * rank))()
    n = host.hpgv_epi_merge_rankings(order, F, rank, models.ctypes.data, 0 if mode == "count" else 1, rows, F * rank)
    want = py_merge(models, order, F, mode)
    assert n == len(want)
    for r, (key, cva, cvc, cells) in zip(rows, want):
        assert tuple(r.snp[:order]) == key and r.cv_count == cvc and abs(r.cv_accuracy - cva) < 1e-12
        assert r.num_risky == len(cells)
    path = tmp_path / "r.epi"
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    fd = libc.fopen(str(path).encode(), b"w")
    host.hpgv_epi_write_report.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(ReportRow), C.c_int, C.c_int, C.c_void_p]
    host.hpgv_epi_write_report(order, 2, 0 if mode == "count" else 1, 1, rows, n, 10, fd)
    libc.fclose(fd)
    assert path.read_text() == py_report(want, order, 2, mode, "training", 10)


def _write_report(host, tmp_path, order, rep, mode, subset, rows, n, max_rank):
    path = tmp_path / f"r{mode}.epi"
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    fd = libc.fopen(str(path).encode(), b"w")
    host.hpgv_epi_write_report.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(ReportRow), C.c_int, C.c_int, C.c_void_p]
    host.hpgv_epi_write_report(order, rep, mode, subset, rows, n, max_rank, fd)
    libc.fclose(fd)
    return path.read_text()


def check_against_reference_merge(host, tmp_path, case, models):
    """models [F, rank] (h.MODEL_DTYPE) -> hpgv_epi_merge_rankings / hpgv_epi_write_report == what the reference's own
    merge_rankings + epistasis_report made of the same per-fold rankings (tests/golden, `merge`)."""
    from golden_util import unb64
    order, F, rank = case["order"], case["F"], case["rank"]
    for mode in (0, 1):
        gold = case["modes"][str(mode)]
        rows = (ReportRow * (F * rank))()
        n = host.hpgv_epi_merge_rankings(order, F, rank, models.ctypes.data, mode, rows, F * rank)
        assert n == len(gold["rows"])
        for r, (snp, cvc, cva_hex, cells) in zip(rows, gold["rows"]):
            assert list(r.snp[:order]) == snp and r.cv_count == cvc
            assert r.cv_accuracy == float.fromhex(cva_hex)              # same additions in the same order, then / F
            got_cells = [sum(int(r.risky_genotypes[q][p]) * 3 ** (order - 1 - p) for p in range(order)) for q in range(r.num_risky)]
            assert got_cells == cells                                   # the risky genotypes of the fold the reference keeps
        text = _write_report(host, tmp_path, order, gold["cv_repetition"], mode, case["subset"], rows, n, gold["max_ranking_size"])
        assert text == gold["report"]


@pytest.mark.parametrize("idx", range(len(GOLD["merge"])))
def test_merge_rankings_and_report_match_the_reference(host, tmp_path, idx):
    """a16/a17: rows and .epi bytes against goldens produced by the reference's merge_rankings (epistasis.c:96-153) and
    epistasis_report (epistasis_report.c:28-82) -- tests/golden/make_golden.py, tie-free inputs."""
    from golden_util import unb64
    case = GOLD["merge"][idx]
    models = unb64(case["models"], h.MODEL_DTYPE, (case["F"], case["rank"]))
    check_against_reference_merge(host, tmp_path, case, models)


def test_report_line_format_of_the_reference():
    """a row of the reference's own output reads `1\\t( 0, 2 )\\t(1-0), (1-1), 10\\t0.546`"""
    rows = [((0, 2), 0.5461, 10, [3, 4])]
    assert py_report(rows, 2, 0, "accu", "training", 50).splitlines()[-1] == "1\t( 0, 2 )\t(1-0), (1-1), 10\t0.546"


# ---- end to end on the GPU --------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("order,nv,A,U,F,mode,subset", [(2, 60, 90, 110, 5, "count", "training"), (2, 45, 64, 64, 4, "accu", "testing"),
                                                         (3, 14, 80, 80, 3, "accu", "training")])
def test_cli_end_to_end_matches_oracle(oracle, tmp_path, order, nv, A, U, F, mode, subset):
    g = synth.make_dataset(nv, A, U, seed=nv + order, order=order, missing=0.01, planted=2)
    data = tmp_path / "d.bin"
    synth.write_dataset(str(data), g, A, U)
    out = tmp_path / "out"
    reps, rank, seed = 2, 15, 4321
    cmd = [CLI, "epi", "-d", str(data), "--order", str(order), "--num-folds", str(F), "--num-cv-runs", str(reps), "--rank-size", str(rank),
           "--eval-subset", subset, "--eval-mode", mode, "--outdir", str(out), "--seed", str(seed), "--stride", "7"]
    res = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr + res.stdout
    assert "Running cross-validation #2..." in res.stdout
    sub = h.SUBSET_TRAINING if subset == "training" else h.SUBSET_TESTING
    for r in range(reps):
        fos, _ = h.k_folds(A, U, F, seed + r)
        want, _ = oracle.search(g, A, U, order, fos, sub, rank, threads=4, num_folds=F)
        models = np.zeros((F, rank), h.MODEL_DTYPE)
        models["snp"][..., :order] = want["snp"][..., :order]
        models["snp"][..., order:] = -1
        models["accuracy"] = want["ba"]
        models["risky_mask"] = want["risky_mask"]
        text = (out / f"hpg-variant.cv{r + 1}.epi").read_text()
        assert text == py_report(py_merge(models, order, F, mode), order, r, mode, subset, rank)


@pytest.mark.gpu
def test_reference_unit_tests_pass_against_the_product():
    """The reference's own test/test_epistasis_model.c (7 tests: byte masks, count tables whole and per fold for order 2
    and 3, confusion matrices, evaluation formulas), compiled against the reference's headers and linked against
    libhpgv_epi_host.so: the leaf functions it calls are the adapters over the CUDA engine (include/hpgv_epi_compat.h).
    The binary is built where the reference's sources are (make -C oracle product-test) and travels with the snapshot."""
    exe = os.path.join(os.path.dirname(PKG), "oracle", "_ref", "test_model_product")
    if not os.path.exists(exe):
        if os.path.isdir("/root/reference"):
            subprocess.run(["make", "-s", "-C", os.path.join(os.path.dirname(PKG), "oracle"), "product-test"], check=True,
                           env={k: v for k, v in os.environ.items() if k not in ("CC", "CXX")})
        else:
            pytest.skip("oracle/_ref/test_model_product not built and /root/reference absent")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]
    assert "7 tests, 0 failed" in p.stderr


@pytest.mark.gpu
def test_cli_two_gpus_equal_one(tmp_path):
    """run_epistasis with HPGV_EPI_GPUS / --gpus 2 (one host thread per GPU, contiguous index ranges, merge on GPU 0) writes
    the bytes of the single-GPU run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    nv, A, U, F = 300, 400, 400, 5
    g = synth.make_dataset(nv, A, U, seed=11, missing=0.01, planted=2)
    data = tmp_path / "d.bin"
    synth.write_dataset(str(data), g, A, U)
    outs = []
    for gpus in (1, 2):
        out = tmp_path / f"out{gpus}"
        cmd = [CLI, "epi", "-d", str(data), "--order", "2", "--num-folds", str(F), "--num-cv-runs", "2", "--rank-size", "20",
               "--eval-subset", "training", "--eval-mode", "accu", "--outdir", str(out), "--seed", "5", "--stride", "50", "--gpus", str(gpus)]
        res = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stderr + res.stdout
        if gpus == 2:
            assert "Range finished: GPU 1" in res.stdout
        outs.append([(out / f"hpg-variant.cv{r}.epi").read_text() for r in (1, 2)])
    assert outs[0] == outs[1]


@pytest.mark.gpu
def test_cli_eval_function_and_out_options(oracle, tmp_path):
    """--eval-function and the shared --out option through the CLI: gamma-ranked report == oracle."""
    nv, A, U, F, rank = 40, 150, 170, 4, 12
    g = synth.make_dataset(nv, A, U, seed=21, missing=0.01, planted=2)
    data = tmp_path / "d.bin"
    synth.write_dataset(str(data), g, A, U)
    out = tmp_path / "out"
    cmd = [CLI, "epi", "-d", str(data), "--order", "2", "--num-folds", str(F), "--num-cv-runs", "1", "--rank-size", str(rank), "--eval-subset", "training",
           "--eval-mode", "accu", "--outdir", str(out), "--seed", "9", "--stride", "10", "--eval-function", "gamma", "--out", "gamma.epi", "-l", "warn"]
    res = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr + res.stdout
    assert "Running cross-validation" not in res.stdout           # --log-level warn silences the INFO lines
    fos, _ = h.k_folds(A, U, F, 9)
    oracle.set_eval_function(h.EVAL_GAMMA)
    try:
        want, _ = oracle.search(g, A, U, 2, fos, 1, rank, threads=4, num_folds=F)
    finally:
        oracle.set_eval_function(h.EVAL_BA)
    models = np.zeros((F, rank), h.MODEL_DTYPE)
    models["snp"][..., :2] = want["snp"][..., :2]
    models["snp"][..., 2:] = -1
    models["accuracy"] = want["ba"]
    models["risky_mask"] = want["risky_mask"]
    assert (out / "gamma.epi").read_text() == py_report(py_merge(models, 2, F, "accu"), 2, 0, "accu", "training", rank)


@pytest.mark.gpu
def test_run_epistasis_missing_dataset_is_fatal(tmp_path):
    res = subprocess.run([CLI, "epi", "-d", str(tmp_path / "nope.bin"), "--order", "2"], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert res.returncode == 1 and "does not exist!" in res.stderr
