"""MovieLens-100K fixture shared by the oracle and the GPU convergence tests
(tests/golden/ml100k.npz, ml100k_curve.json; made by tests/golden/make_ml100k.py)."""
import hashlib
import json
import os

import numpy as np

from svdfeature_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    z = np.load(os.path.join(HERE, "golden", "ml100k.npz"))
    gold = json.load(open(os.path.join(HERE, "golden", "ml100k_curve.json")))

    def csr(u, i, r):
        ones = np.ones(len(r), np.float32)
        return synth.fixed_csr(r.astype(np.float32), uidx=u.astype(np.uint32), uval=ones, iidx=i.astype(np.uint32), ival=ones)

    train = csr(z["base_user"], z["base_item"], z["base_rating"])
    test = csr(z["test_user"], z["test_item"], z["test_rating"])
    return train, test, z["test_rating"].astype(np.float64), gold


def run(trainer, train, test, truth, rounds, seed=10, tmp=None):
    """40 rounds like demo/basicMF/run.sh; returns ({round: test RMSE}, sha256 of the model file or None)."""
    trainer.init(seed)
    curve = {}
    for r in range(41):
        if r in rounds:
            p = trainer.predict_csr(test).astype(np.float64)
            curve[r] = float(np.sqrt(np.mean((p - truth) ** 2)))
        if r < 40:
            trainer.set_round(r)
            trainer.update_csr(train)
            if hasattr(trainer, "finish_round"):
                trainer.finish_round()
    sha = hashlib.sha256(trainer.model_bytes(tmp)).hexdigest() if tmp is not None else None
    return curve, sha
