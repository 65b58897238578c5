import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.Checker("oracle")


@pytest.fixture(scope="session")
def ref():
    import oracle_lib
    if not oracle_lib.available("ref"):
        if os.path.isdir("/root/reference"):
            oracle_lib.build_oracle("ref")
        else:
            pytest.skip("oracle/_ref/libhpgref.so not built and /root/reference absent")
    return oracle_lib.Checker("ref")


@pytest.fixture(scope="session")
def engine():
    import hpg_variant_b200 as h
    try:
        eng = h.EpistasisEngine(0)
    except h.HpgvError as e:
        if e.code in (-1, -4):            # HPGV_E_CUDA / HPGV_E_UNSUPPORTED: no (Blackwell) GPU on this box
            pytest.skip(f"no usable CUDA device: {e}")
        raise
    yield eng
    eng.close()
