"""CPU checks of the product boundary: the C-ABI library builds for sm_100a, loads, exports every
symbol include/hpgv_epi.h declares, and refuses to run without a GPU (no CPU fallback)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from hpg_variant_b200 import build
    build.build()
    return build


def declared_symbols(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:hpgv_epi_|hpgv_)\w+)\s*\(", text)))


def exported(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


def test_library_exports_every_declared_symbol(built):
    from hpg_variant_b200 import _lib
    syms = declared_symbols("hpgv_epi.h")
    assert sorted(_lib.SYMBOLS) == syms
    exp = exported(built.LIB)
    missing = [s for s in syms if s not in exp]
    assert not missing, missing
    lib = _lib.load()
    for s in syms:
        assert hasattr(lib, s)


def test_cuda_code_is_sm100a_with_bulk_copies(built):
    sass = subprocess.run(["cuobjdump", "-sass", built.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in sass or "SM100a" in sass.upper() or "sm_100" in sass
    assert "UBLKCP" in sass          # cp.async.bulk (TMA engine) staging
    assert "POPC" in sass and "LOP3" in sass


def test_no_cpu_fallback(built):
    import torch
    import hpg_variant_b200 as h
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(h.HpgvError) as e:
        h.EpistasisEngine(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "hpg_variant_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "__init__.py" and False, f"{f} mentions the oracle"


def test_k_folds_host_matches_golden():
    """hpgv_epi_k_folds (product, own 48-bit LCG) == the reference's get_k_folds with the clock pinned."""
    import hpg_variant_b200 as h
    from golden_util import load_golden
    for rec in load_golden()["kfolds"]:
        fos, sizes = h.k_folds(rec["A"], rec["U"], rec["k"], rec["seed"])
        assert fos.tolist() == rec["fold_of_sample"]
        assert sizes.tolist() == rec["sizes"]


def test_num_combinations():
    import hpg_variant_b200 as h
    assert h.num_combinations(10_000, 2) == 49_995_000
    assert h.num_combinations(100_000, 2) == 4_999_950_000
    assert h.num_combinations(5_000, 3) == 20_820_835_000
    assert h.num_combinations(20_000, 2) == 199_990_000


def test_synth_dataset_roundtrip(tmp_path):
    from hpg_variant_b200 import synth
    g = synth.make_dataset(50, 30, 40, seed=9)
    assert g.shape == (50, 70) and set(np.unique(g)) <= {0, 1, 2, 255}
    p = tmp_path / "d.bin"
    synth.write_dataset(p, g, 30, 40)
    g2, a, u = synth.read_dataset(p)
    assert (a, u) == (30, 40) and np.array_equal(g, g2)
    assert np.array_equal(synth.make_dataset(50, 30, 40, seed=9), g)
