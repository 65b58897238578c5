"""The GPU's high-risk rule (csrc/epi_device.cuh: high_risk) decides by the sign of the integer
margin ca*U - cu*A whenever the margin lies outside a 2^-19 band, and replays the six float32
operations of mdr.c:45-75 inside it.  This test re-states both in numpy (float32 arithmetic in
numpy is IEEE round-to-nearest, unfused) and checks, over dense grids, random and adversarial
near-tie inputs, that the integer shortcut never disagrees with the float32 sequence."""
import numpy as np
import pytest


def float32_rule(ca, cu, A, U):
    with np.errstate(all="ignore"):
        r = np.float32(A) / np.float32(U)
        fa, fu = ca.astype(np.float32), cu.astype(np.float32)
        total = fa + fu
        pu = fu * r
        rr = total / (pu + fa)
        nu = pu * rr
        na = total - nu
        return na >= nu


def gpu_rule(ca, cu, A, U):
    """numpy mirror of hpgv::high_risk; returns (decision, used_fast_path)."""
    ca, cu = ca.astype(np.int64), cu.astype(np.int64)
    if A == U:
        return (ca >= cu) & (ca > 0), np.ones(ca.shape, bool)
    m = cu * A
    d = ca * U - m
    band = (m >> 19) + 1
    fast = (d > band) | (d < -band)
    slow = float32_rule(ca, cu, A, U)
    return np.where(fast, d > 0, slow), fast


CLASS_SIZES = [(1900, 2100), (10, 80), (49, 98), (1000, 1000), (2000, 2000), (1234, 4321), (25000, 25000), (24999, 25001),
               (3, 7), (7, 3), (60000, 50000), (999, 1000), (1, 1), (65535, 1)]


@pytest.mark.parametrize("A,U", CLASS_SIZES)
def test_fast_path_agrees_with_float32(A, U):
    rng = np.random.default_rng(A * 7 + U)
    lim = min(A, 300) + 1, min(U, 300) + 1
    ca, cu = np.meshgrid(np.arange(lim[0]), np.arange(lim[1]), indexing="ij")
    sets = [(ca.ravel(), cu.ravel())]
    n = 400_000
    rca = rng.integers(0, A + 1, n)
    sets.append((rca, rng.integers(0, U + 1, n)))
    # adversarial: cu chosen so that ca*U ~= cu*A, +-2
    near = np.clip(np.rint(rca * (U / A)).astype(np.int64)[:, None] + np.arange(-2, 3)[None, :], 0, U)
    sets.append((np.repeat(rca, 5), near.ravel()))
    for a, u in sets:
        want = float32_rule(a, u, A, U)
        got, fast = gpu_rule(a, u, A, U)
        assert np.array_equal(got, want)
        # the shortcut itself (where taken) must equal the float32 sequence
        assert np.array_equal(got[fast], want[fast])
        if A != U and A * U > 1000:
            assert fast.mean() > 0.5


def test_balanced_rule_is_exact():
    for A in (1, 5, 1000, 25000, 60000):
        rng = np.random.default_rng(A)
        ca, cu = rng.integers(0, A + 1, 300_000), rng.integers(0, A + 1, 300_000)
        ca[:1000] = cu[:1000]                      # ties
        ca[1000:1100] = 0; cu[1000:1100] = 0       # empty cells: 0/0 = NaN -> low risk
        assert np.array_equal((ca >= cu) & (ca > 0), float32_rule(ca, cu, A, A))
