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

# ---- SVD++ user blocks (configs[2] shape, scaled): Hogwild across users, a user's rows sequential
if "--svdpp" in sys.argv:
    nu2, n2, fb = 120000, 6_000_000, 100
    pp = dict(num_user=nu2, num_item=ni, num_factor=k, learning_rate=0.005, wd_user=0.004, wd_item=0.004, base_score=3.6,
              num_ufeedback=ni, wd_ufeedback=0.004, ufeedback_init_sigma=0.01)
    tr = synth.user_grouped(n2, nu2, ni, avg_fb=fb, seed=31, zipf_s=1.0, zipf_q=70.0, user_sigma=1.0)
    o3 = COracle(1, 0, 0, pp)
    o3.init(10)
    g3 = api.SvdGpu(nu2, ni, k, num_ufeedback=ni, format_type=1)
    g3.set_hparams(learning_rate=0.005, wd_user=0.004, wd_item=0.004, wd_ufeedback=0.004, base_score=o3.base_score)
    g3.set_mode(api.MODE_HOGWILD)
    for a in sys.argv:
        if a.startswith("--opt="):
            kk, v = a[6:].split("=")
            g3.set_option(kk, int(v))
    g3.upload(*[a.copy() for a in o3.arrays()])
    res = {"svdpp_users": nu2, "ratings": n2, "avg_feedback": fb, "epochs": []}
    lab = tr[6]
    import time
    for e in range(3):
        o3.update_ugroup(tr)
        t0 = time.perf_counter()
        g3.update_ugroup(tr)
        g3.sync()
        res.setdefault("gpu_epoch_s", []).append(round(time.perf_counter() - t0, 3))
        po, pg = o3.predict_ugroup(tr), g3.predict_ugroup(tr)
        res.setdefault("nan_frac", []).append(float(np.mean(~np.isfinite(pg))))
        res["epochs"].append({"epoch": e + 1, "rmse_pred_vs_sequential": float(np.sqrt(np.mean((po - pg) ** 2))),
                              "train_rmse_sequential": float(np.sqrt(np.mean((po - lab) ** 2))),
                              "train_rmse_hogwild": float(np.sqrt(np.mean((pg - lab) ** 2)))})
    print(json.dumps(res))
