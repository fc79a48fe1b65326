"""GPU experiment: what bounds the ordered (exact) kernel?  Same row count, different item
popularity (hot-row chain length) -> throughput."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svdfeature_b200 import api, synth  # noqa: E402

NU, NI, K, N = 480000, 18000, 64, 2_000_000
rng = np.random.default_rng(1)
W = (rng.standard_normal((NU + NI, K)) * 0.01).astype(np.float32)
for name, kw in [("zipf_q70", dict(zipf_q=70.0)), ("zipf_q1000", dict(zipf_q=1000.0)), ("uniform", dict(zipf_s=0.0)),
                 ("zipf_q0", dict(zipf_q=0.0))]:
    data = synth.basic_mf(N, NU, NI, seed=3, **kw)
    top = np.bincount(data[2][1::2]).max()
    g = api.SvdGpu(NU, NI, K)
    g.set_hparams(learning_rate=0.005, wd_user=0.004, wd_item=0.004, base_score=3.6)
    g.set_mode(api.MODE_EXACT)
    g.upload(np.zeros(NU + NI, np.float32), W, np.zeros(1, np.float32))
    b = g.batch_create(data)
    g.batch_update(b)
    g.sync()
    g.timer_start()
    for _ in range(3):
        g.batch_update(b)
    ms = g.timer_stop() / 3
    print(json.dumps(dict(items=name, rows=N, hottest_item_rows=int(top), ms=ms, minst_s=N / ms / 1e3,
                          us_per_hot_row=1e3 * ms / top)), flush=True)
    b.close()
    g.close()
