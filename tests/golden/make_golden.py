"""Generates tests/golden/epi_golden.json by RUNNING THE REFERENCE's own code
(oracle/_ref/libhpgref.so, built from /root/reference by oracle/Makefile) in this
container.  The vectors travel to the GPU box, where /root/reference does not exist.

    python tests/golden/make_golden.py

Every case stores its inputs and the reference's outputs:
  eval     per-combination, per-fold TRAINING counts, float32 risk masks, confusion
           matrices (both subsets) and BA through set_genotypes_masks ->
           combination_counts_all_folds -> choose_high_risk_combinations2(mdr_high_risk_combinations2)
           -> confusion_matrix -> evaluate_model, with injected folds
  topn     canonical top-N of an exhaustive search driven over the same leaf functions
  risk     mdr_high_risk_combinations2 flags on grids of (ca, cu), incl. the A=1900/U=2100 tie vector
  kfolds   get_k_folds with the microsecond clock pinned (gettimeofday interposed)
  blocked  the reference's blocked enumeration (get_first/next_combination_in_block), defects and all
  formulas evaluate_model on the two matrices of test/test_epistasis_model.c:513-534
  merge    the reference's own merge_rankings (epistasis.c:96-153) and epistasis_report (epistasis_report.c:28-82) on
           per-fold rankings of an exhaustive search: rows in report order (CV-C, CV-A, kept risky cells) and the .epi
           text, for CV-A and CV-C.  Inputs are chosen tie-free (distinct CV-A values; CV-A monotone in CV-C so that the
           reference's CV-C comparator, model.c:525-543, is a consistent order -- SURVEY F9), the generator asserts it.
"""
import base64
import itertools
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib  # noqa: E402
from hpg_variant_b200 import synth  # noqa: E402


def b64(a):
    return base64.b64encode(np.ascontiguousarray(a).tobytes()).decode()


def tie_free(rows):
    accs = [r[2] for r in rows]
    if len(set(accs)) != len(accs) or any(np.isnan(a) for a in accs):
        return False
    # a higher CV-C never comes with a lower CV-A: compare_risky_heap_count_max is then a consistent strict order
    return all(not (a[1] > b[1] and a[2] < b[2]) for a in rows for b in rows)


def merge_cases(ref):
    import tempfile
    cases = []
    for order, nv, A, U, F, rank, subset in [(2, 14, 90, 110, 5, 6, 1), (2, 12, 700, 640, 4, 8, 0), (3, 8, 80, 70, 3, 7, 1), (2, 10, 149, 198, 10, 3, 1)]:
        for seed in range(1000):
            g = synth.make_dataset(nv, A, U, seed=9000 + seed, order=order, missing=0.01, planted=2)
            rng = np.random.default_rng(seed)
            fos = np.concatenate([rng.permutation(A) % F, rng.permutation(U) % F]).astype(np.int32)
            top, _ = ref.search(g, A, U, order, fos, subset, rank, threads=2, num_folds=F)
            models = np.zeros((F, rank), oracle_lib.MODEL_DTYPE)
            models[:] = top
            rows_a = ref.merge_rankings(order, models, 1)
            rows_c = ref.merge_rankings(order, models, 0)
            if tie_free(rows_a) and tie_free(rows_c) and any(r[1] > 1 for r in rows_a) and any(r[1] < F for r in rows_a):
                break
        else:
            raise SystemExit(f"no tie-free merge case found for {(order, nv, A, U, F, rank, subset)}")
        rec = {"order": order, "nv": nv, "A": A, "U": U, "F": F, "rank": rank, "subset": subset, "dataset_seed": 9000 + seed,
               "genotypes": b64(g), "fold_of_sample": b64(fos), "models": b64(models), "modes": {}}
        for mode, rows in ((0, rows_c), (1, rows_a)):
            with tempfile.TemporaryDirectory() as tmp:
                text = ref.report(order, models, mode, subset, 2, rank + 3, os.path.join(tmp, "r.epi"))
            rec["modes"][str(mode)] = {"rows": [[list(r[0]), r[1], r[2].hex(), r[3]] for r in rows], "report": text,
                                       "cv_repetition": 2, "max_ranking_size": rank + 3}
        cases.append(rec)
    return cases


def main():
    oracle_lib.build_oracle("ref")
    ref = oracle_lib.Checker("ref")
    out = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref/libhpgref.so (reference sources @ /root/reference)", "eval": [], "topn": [], "risk": [], "kfolds": [], "blocked": [], "formulas": []}

    cases = []
    # the shipped fixture (legacy 16-byte header), deterministic folds i % 10 per class
    g, A, U = synth.read_dataset("/root/reference/test/epistasis_dataset.bin")
    fos = np.concatenate([np.arange(A) % 10, np.arange(U) % 10]).astype(np.int32)
    cases.append(("fixture_order2", g, A, U, 10, fos, 2))
    cases.append(("fixture_order3", g, A, U, 10, fos, 3))
    rng = np.random.default_rng(20261017)
    for name, nv, A, U, F, order, miss in [
        ("small_o2", 8, 21, 30, 3, 2, 0.03), ("small_o3", 7, 21, 30, 3, 3, 0.03),
        ("balanced_o2", 9, 40, 40, 5, 2, 0.01), ("unbalanced_ties_o2", 8, 19, 21, 2, 2, 0.0),
        ("blocks_o2", 6, 300, 280, 2, 2, 0.02), ("blocks_o3", 5, 300, 280, 2, 3, 0.02),
    ]:
        g = synth.make_dataset(nv, A, U, seed=int(rng.integers(1 << 30)), order=order, missing=miss, planted=1)
        fos = np.concatenate([rng.permutation(A) % F, rng.permutation(U) % F]).astype(np.int32)
        cases.append((name, g, A, U, F, fos, order))

    for name, g, A, U, F, fos, order in cases:
        nv = g.shape[0]
        combs = np.array(list(itertools.combinations(range(nv), order)), np.int32)
        rec = {"name": name, "nv": nv, "A": A, "U": U, "F": F, "order": order,
               "genotypes": b64(g), "fold_of_sample": b64(fos), "combs": b64(combs)}
        for subset, sname in ((1, "training"), (0, "testing")):
            r = ref.eval(g, A, U, order, fos, subset, combs)
            if subset == 1:
                rec["counts_aff"] = b64(r["counts_aff"])
                rec["counts_unaff"] = b64(r["counts_unaff"])
                rec["risky_mask"] = b64(r["risky_mask"])
            rec["conf_" + sname] = b64(r["conf"])
            rec["ba_" + sname] = b64(r["ba"])
        out["eval"].append(rec)
        for subset, sname in ((1, "training"), (0, "testing")):
            n = 5
            top, n_out = ref.search(g, A, U, order, fos, subset, n, threads=2, num_folds=F)
            out["topn"].append({"name": name, "subset": subset, "rank": n, "models": b64(top), "n_out": n_out.tolist()})

    for A, U, lim in [(1900, 2100, 64), (10, 80, 90), (49, 98, 50), (1000, 1000, 40), (1234, 4321, 70)]:
        ca, cu = np.meshgrid(np.arange(lim), np.arange(lim), indexing="ij")
        flags = ref.high_risk(ca.ravel(), cu.ravel(), A, U)
        out["risk"].append({"A": A, "U": U, "lim": lim, "flags": b64(np.packbits(flags))})
    # test/test_mdr.c:52-66 vector
    f = ref.high_risk([8, 4, 9, 8, 4], [40, 75, 20, 63, 40], 10, 80)
    out["risk_test_mdr"] = {"ca": [8, 4, 9, 8, 4], "cu": [40, 75, 20, 63, 40], "A": 10, "U": 80, "flags": f.astype(int).tolist()}

    for A, U, k, seed in [(200, 200, 10, 5), (150, 250, 7, 123456), (50, 75, 4, 31337), (49, 98, 10, 999999), (8, 12, 5, 1), (16, 4, 10, 77)]:
        fos, sizes = ref.k_folds(A, U, k, seed)
        out["kfolds"].append({"A": A, "U": U, "k": k, "seed": seed, "fold_of_sample": fos.tolist(), "sizes": sizes.tolist()})

    for nv, order, stride in [(10, 2, 4), (9, 2, 4), (12, 2, 5), (10, 3, 4), (7, 3, 7), (10, 3, 10), (4, 2, 100)]:
        combs, n = ref.enumerate_blocked(nv, order, stride)
        out["blocked"].append({"nv": nv, "order": order, "stride": stride, "n": n, "combs": combs.tolist()})

    for m in ([40, 2, 4, 10], [20, 10, 10, 20]):
        out["formulas"].append({"conf": m, "values": [ref.evaluate(m, fn) for fn in (0, 1, 3, 4)], "functions": ["CA", "BA", "GAMMA", "TAU_B"]})

    out["merge"] = merge_cases(ref)

    path = os.path.join(HERE, "epi_golden.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=0)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
