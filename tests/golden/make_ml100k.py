"""MovieLens-100K convergence fixture (SURVEY.md section 4, item iii) from the reference's shipped
demo data and the UNMODIFIED reference trainer.

Run in the build container only (needs /root/reference and oracle/_ref):
    python tests/golden/make_ml100k.py

Writes next to this script:
  ml100k.npz          demo/basicMF/ua.base (90 570 ratings) in a fixed shuffled order and ua.test
                      (9 430 ratings): user, item (0-based, as demo/basicMF/mkbasicfeature.py makes
                      them), rating as uint16 / uint16 / uint8
  ml100k_curve.json   test RMSE of the compiled reference trainer (oracle/_ref/libsvdf_ref.so,
                      ISVDTrainer::update loop, demo/basicMF/basicMF.conf, seed 10) after rounds
                      0, 1, 5, 10, 20, 30, 40, and the sha256 of its model file after round 40
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference/demo/basicMF"

from _oracle import RefTrainer, build_oracle  # noqa: E402
from svdfeature_b200 import synth  # noqa: E402

PARAMS = dict(num_user=943, num_item=1682, num_factor=64, learning_rate=0.005, wd_user=0.004, wd_item=0.004,
              base_score=3.0)
ROUNDS = [0, 1, 5, 10, 20, 30, 40]


def load(name):
    a = np.loadtxt(os.path.join(REF, name), dtype=np.int64)
    return (a[:, 0] - 1).astype(np.uint16), (a[:, 1] - 1).astype(np.uint16), a[:, 2].astype(np.uint8)


def csr(u, i, r):
    ones = np.ones(len(r), np.float32)
    return synth.fixed_csr(r.astype(np.float32), uidx=u.astype(np.uint32), uval=ones, iidx=i.astype(np.uint32), ival=ones)


def main():
    build_oracle()
    bu, bi, br = load("ua.base")
    perm = np.random.default_rng(100).permutation(len(br))  # the demo shuffles the lines too (line_shuffle)
    bu, bi, br = bu[perm], bi[perm], br[perm]
    tu, ti, tr = load("ua.test")
    np.savez_compressed(os.path.join(HERE, "ml100k.npz"), base_user=bu, base_item=bi, base_rating=br, test_user=tu,
                        test_item=ti, test_rating=tr)
    train, test = csr(bu, bi, br), csr(tu, ti, tr)
    t = RefTrainer(0, 0, 0, PARAMS)
    t.init(10)
    curve = {}
    for r in range(41):
        if r in ROUNDS:
            p = t.predict_csr(test).astype(np.float64)
            curve[str(r)] = float(np.sqrt(np.mean((p - tr) ** 2)))
        if r < 40:
            t.set_round(r)
            t.update_csr(train)
    with tempfile.TemporaryDirectory() as d:
        sha = hashlib.sha256(t.model_bytes(d)).hexdigest()
    json.dump({"params": PARAMS, "seed": 10, "test_rmse_after_round": curve, "model_sha256_after_round_40": sha,
               "source": "compiled unmodified reference (oracle/_ref), demo/basicMF data, tests/golden/make_ml100k.py"},
              open(os.path.join(HERE, "ml100k_curve.json"), "w"), indent=1)
    print(curve, sha)


if __name__ == "__main__":
    main()
