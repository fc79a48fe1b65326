"""GPU experiment: how far does Hogwild drift from the reference's sequential order?
Trains the CPU oracle (sequential, = the reference bit for bit) and the GPU (Hogwild, default
options) on the same Netflix-shaped ratings for E epochs and reports, per epoch, the RMSE between
their predictions on a held-out set and each side's held-out RMSE.  TEST TOOLING (uses oracle/)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _oracle import COracle, build_oracle  # noqa: E402
from svdfeature_b200 import api, synth  # noqa: E402

build_oracle()
nu, ni, n, k, epochs = 120000, 18000, 3_000_000, 64, 5
if len(sys.argv) > 1:
    nu, n = int(sys.argv[1]), int(sys.argv[2])
params = dict(num_user=nu, num_item=ni, num_factor=k, learning_rate=0.005, wd_user=0.004, wd_item=0.004, base_score=3.6)
train = synth.basic_mf(n, nu, ni, seed=21, zipf_s=1.0, zipf_q=70.0, user_sigma=1.0)
test = synth.basic_mf(200000, nu, ni, seed=22, zipf_s=1.0, zipf_q=70.0, user_sigma=1.0)
o = COracle(0, 0, 0, params)
o.init(10)
out = {"users": nu, "items": ni, "ratings": n, "k": k, "epochs": []}
for name, opts in (("red/red (default)", {}), ("store/red", {"scatter_user": 0})):
    o2 = COracle(0, 0, 0, params)
    o2.init(10)
    g = api.SvdGpu(nu, ni, k)
    g.set_hparams(learning_rate=0.005, wd_user=0.004, wd_item=0.004, base_score=o2.base_score)
    g.set_mode(api.MODE_HOGWILD)
    for kk, v in opts.items():
        g.set_option(kk, v)
    g.upload(*[a.copy() for a in o2.arrays()])
    for e in range(epochs):
        o2.update_csr(train)
        g.update_csr(train)
        po, pg = o2.predict_csr(test), g.predict_csr(test)
        out["epochs"].append({"scatter": name, "epoch": e + 1,
                              "rmse_pred_vs_sequential": float(np.sqrt(np.mean((po - pg) ** 2))),
                              "heldout_rmse_sequential": float(np.sqrt(np.mean((po - test[1]) ** 2))),
                              "heldout_rmse_hogwild": float(np.sqrt(np.mean((pg - test[1]) ** 2)))})
    g.close()
print(json.dumps(out))
