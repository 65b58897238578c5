import base64
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load_golden():
    with open(os.path.join(HERE, "golden", "epi_golden.json")) as fh:
        return json.load(fh)


def unb64(s, dtype, shape=None):
    a = np.frombuffer(base64.b64decode(s), dtype=dtype).copy()
    return a.reshape(shape) if shape is not None else a


def eval_case_arrays(rec):
    nv, A, U, F, order = rec["nv"], rec["A"], rec["U"], rec["F"], rec["order"]
    C = 3 ** order
    g = unb64(rec["genotypes"], np.uint8, (nv, A + U))
    fos = unb64(rec["fold_of_sample"], np.int32)
    combs = unb64(rec["combs"], np.int32, (-1, order))
    n = combs.shape[0]
    want = {
        "counts_aff": unb64(rec["counts_aff"], np.int32, (n, F, C)),
        "counts_unaff": unb64(rec["counts_unaff"], np.int32, (n, F, C)),
        "risky_mask": unb64(rec["risky_mask"], np.uint32, (n, F)),
        "conf_training": unb64(rec["conf_training"], np.uint32, (n, F, 4)),
        "conf_testing": unb64(rec["conf_testing"], np.uint32, (n, F, 4)),
        "ba_training": unb64(rec["ba_training"], np.float64, (n, F)),
        "ba_testing": unb64(rec["ba_testing"], np.float64, (n, F)),
    }
    return g, A, U, F, order, fos, combs, want
