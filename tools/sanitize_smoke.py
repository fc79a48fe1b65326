"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel family once."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svdfeature_b200 import api, synth  # noqa: E402

rng = np.random.default_rng(0)


def model(g, rows, k, ng):
    g.upload(np.zeros(rows, np.float32), (rng.standard_normal((rows, k)) * 0.01).astype(np.float32),
             np.zeros(max(ng, 1), np.float32))


for mode in (api.MODE_HOGWILD, api.MODE_EXACT):
    # basic MF (fast path), odd sizes so that tiles / windows are ragged
    g = api.SvdGpu(3001, 701, 64)
    g.set_hparams(learning_rate=0.005, wd_user=0.004, wd_item=0.004, base_score=3.6)
    g.set_mode(mode)
    model(g, 3001 + 701, 64, 0)
    d = synth.basic_mf(20011, 3001, 701, seed=1)
    g.update_csr(d)
    b = g.batch_create(d)
    if mode == api.MODE_HOGWILD:
        g.batch_update(b, 13, 19999)
    g.batch_predict(b, 5, 20000)
    g.predict_csr(d)
    b.close()
    g.close()
    # ragged general rows with globals (generic pass), k not a multiple of 4
    g = api.SvdGpu(300, 200, 13, num_global=40)
    g.set_hparams(learning_rate=0.01, wd_user=0.004, wd_item=0.004, wd_global=0.001, base_score=3.6)
    g.set_mode(mode)
    model(g, 500, 13, 40)
    d = synth.random_general(3000, 300, 200, 40, seed=2, max_g=20, max_u=5, max_i=40, allow_dup=True)
    g.update_csr(d)
    g.predict_csr(d)
    g.close()
    # SVD++ user blocks
    g = api.SvdGpu(400, 150, 32, num_ufeedback=150, format_type=1)
    g.set_hparams(learning_rate=0.005, wd_user=0.004, wd_item=0.004, wd_ufeedback=0.004, base_score=3.6)
    g.set_mode(mode)
    model(g, 150 + 400 + 150, 32, 0)
    ug = synth.user_grouped(5000, 400, 150, avg_fb=12, seed=3)
    g.update_ugroup(ug)
    g.predict_ugroup(ug)
    g.items_snapshot()
    g.items_pack_delta()
    g.items_apply_delta(1.0)
    g.sync()
    g.close()
# ordered mode through the item-owner kernels: resident batch + host call, the fast link (k = 64, 128),
# the generic link (k = 20), few shared-memory item slots, the split link (k_own2)
for k, opts in ((64, {}), (128, {"own_slots": 2}), (20, {}), (64, {"own_partner": 1}), (16, {"own_partner": 1, "chunk_rows": 6000})):
    g = api.SvdGpu(1501, 301, k)
    g.set_hparams(learning_rate=0.005, wd_user=0.004, wd_item=0.004, wd_user_bias=0.001, base_score=3.6)
    g.set_mode(api.MODE_EXACT)
    for name, v in opts.items():
        g.set_option(name, v)
    model(g, 1501 + 301, k, 0)
    d = synth.basic_mf(15013, 1501, 301, seed=5, zipf_q=3.0)
    b = g.batch_create(d)
    g.batch_update(b)
    g.update_csr(d)
    g.sync()
    assert g.counter("own_launches") >= 2, g.counter("own_launches")
    b.close()
    g.close()
print("sanitize_smoke done")
