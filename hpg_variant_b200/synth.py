"""Deterministic synthetic epistasis datasets.

Per SNP a minor-allele frequency q ~ U(0.05, 0.5); genotypes drawn under
Hardy-Weinberg ((1-q)^2, 2q(1-q), q^2) as bytes 0/1/2; 0.5 % of the entries set
to 255 (missing); cases first.  A few causal tuples are planted by redrawing the
cases' genotypes at those SNPs with 3x odds on the all-heterozygous and
all-homozygous-minor cells, which makes the top models non-trivial and tie-free.
Files use the reference's current 12-byte header (dataset.c:58-63).
"""
import struct

import numpy as np

CONFIGS = {
    # name: (num_variants, num_affected, num_unaffected, order, num_folds, seed)
    "c2": (10_000, 1_000, 1_000, 2, 10, 1002),
    "c3": (100_000, 2_000, 2_000, 2, 10, 1003),
    "c4": (5_000, 2_000, 2_000, 3, 5, 1004),
    "c5": (20_000, 25_000, 25_000, 2, 10, 1005),
}


def make_dataset(num_variants, num_affected, num_unaffected, seed, order=2, missing=0.005, planted=5, out=None, planted_out=None):
    """planted_out: optional list that receives the planted causal tuples (ascending SNP indices)."""
    rng = np.random.default_rng(seed)
    S = num_affected + num_unaffected
    g = out if out is not None else np.empty((num_variants, S), np.uint8)
    q = rng.uniform(0.05, 0.5, num_variants)
    t1 = ((1 - q) ** 2 * 65536.0).astype(np.uint32)
    t2 = (((1 - q) ** 2 + 2 * q * (1 - q)) * 65536.0).astype(np.uint32)
    step = max(1, (1 << 24) // max(S, 1))
    miss_thr = int(missing * 65536.0)
    for lo in range(0, num_variants, step):
        hi = min(num_variants, lo + step)
        u = rng.integers(0, 65536, size=(hi - lo, S), dtype=np.uint16)
        blk = (u >= t1[lo:hi, None]).astype(np.uint8)
        blk += (u >= t2[lo:hi, None])
        if miss_thr > 0:
            m = rng.integers(0, 65536, size=(hi - lo, S), dtype=np.uint16) < miss_thr
            blk[m] = 255
        g[lo:hi] = blk
    # planted causal tuples (cases only)
    nplant = min(planted, num_variants // order)
    if nplant > 0:
        snps = rng.choice(num_variants, size=nplant * order, replace=False).reshape(nplant, order)
        if planted_out is not None:
            planted_out.extend(tuple(sorted(int(v) for v in tup)) for tup in snps)
        for tup in snps:
            probs = []
            for v in tup:
                probs.append(np.array([(1 - q[v]) ** 2, 2 * q[v] * (1 - q[v]), q[v] ** 2]))
            joint = probs[0]
            for p in probs[1:]:
                joint = np.multiply.outer(joint, p)
            w = np.ones_like(joint)
            w[(1,) * order] = 3.0
            w[(2,) * order] = 3.0
            joint = (joint * w).ravel()
            joint /= joint.sum()
            cells = rng.choice(joint.size, size=num_affected, p=joint)
            keep_missing = [g[v, :num_affected] == 255 for v in tup]
            for pos, v in enumerate(tup):
                gv = (cells // (3 ** (order - 1 - pos))) % 3
                row = g[v, :num_affected]
                row[:] = np.where(keep_missing[pos], 255, gv).astype(np.uint8)
    return g


def make_config(name, num_variants=None):
    nv, a, u, order, folds, seed = CONFIGS[name]
    if num_variants is not None:
        nv = num_variants
    return make_dataset(nv, a, u, seed, order=order), a, u, order, folds


def write_dataset(path, genotypes, num_affected, num_unaffected):
    with open(path, "wb") as f:
        f.write(struct.pack("<III", genotypes.shape[0], num_affected, num_unaffected))
        f.write(np.ascontiguousarray(genotypes, dtype=np.uint8).tobytes())


def read_dataset(path):
    """Reads the current 12-byte header or the legacy 16-byte one."""
    raw = np.fromfile(path, dtype=np.uint8)
    n = raw.size
    nv, a, u = struct.unpack("<III", raw[:12].tobytes())
    if nv >= 1 and a >= 1 and u >= 1 and 12 + nv * (a + u) == n:
        off = 12
    else:
        nv64, a, u = struct.unpack("<QII", raw[:16].tobytes())
        nv = nv64
        if not (nv >= 1 and 16 + nv * (a + u) <= n <= 16 + nv * (a + u) + 7):
            raise ValueError("unrecognised epistasis dataset header")
        off = 16
    return raw[off:off + nv * (a + u)].reshape(nv, a + u).copy(), a, u
