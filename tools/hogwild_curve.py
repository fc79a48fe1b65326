"""GPU experiment: distance from the sequential order against throughput, configs[1] shape.
The yardstick is the GPU's own ordered mode (bit-identical to the reference's loop).  For every setting:
  * rmse_one_pass: 2 M ratings trained once on a fresh model, predictions on those ratings against the
    ordered run's (what bench.py's `parity` object measures);
  * rmse_3_epochs: 10 M ratings x 3 epochs, predictions on 200 k held-out ratings against the ordered run's;
  * ginst_s: resident throughput over a 20 M-rating batch.
Settings vary how many instances Hogwild keeps in flight (CTAs per SM, gather ring depth).
One JSON line per setting -> gpurun_out/hogwild_curve.jsonl."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svdfeature_b200 import api, synth  # noqa: E402

NU, NI, K = 480000, 18000, 64
HP = dict(learning_rate=0.005, wd_user=0.004, wd_item=0.004, base_score=3.6)
rng = np.random.default_rng(10)
W0 = (rng.standard_normal((NU + NI, K)) * 0.01).astype(np.float32)
one = synth.basic_mf(2_000_000, NU, NI, seed=10)
multi = synth.basic_mf(10_000_000, NU, NI, seed=21)
held = synth.basic_mf(200_000, NU, NI, seed=22)
big = synth.basic_mf(20_000_000, NU, NI, seed=3)
out = open(os.path.join(ROOT, "gpurun_out", "hogwild_curve.jsonl"), "a")


def trainer(mode, opts):
    g = api.SvdGpu(NU, NI, K)
    g.set_hparams(**HP)
    g.set_mode(mode)
    for k, v in opts.items():
        g.set_option(k, v)
    g.upload(np.zeros(NU + NI, np.float32), W0, np.zeros(1, np.float32))
    return g


def run(mode, opts):
    g = trainer(mode, opts)
    g.update_csr(one)
    p1 = g.predict_csr(one)
    g.close()
    g = trainer(mode, opts)
    for _ in range(3):
        g.update_csr(multi)
    p3 = g.predict_csr(held)
    g.close()
    g = trainer(mode, opts)
    b = g.batch_create(big)
    g.batch_update(b)
    g.sync()
    g.timer_start()
    for _ in range(3):
        g.batch_update(b)
    ms = g.timer_stop() / 3
    b.close()
    g.close()
    return p1, p3, len(big[1]) / ms / 1e6


y1, y3, gy = run(api.MODE_EXACT, {})
rows = [("ordered (k_own)", api.MODE_EXACT, {})]
for name, opts in (("hogwild default (2 CTAs/SM, ring 4)", {}), ("hogwild ring 2", {"ring_depth": 2}),
                   ("hogwild 1 CTA/SM", {"ctas_per_sm": 1}), ("hogwild 1 CTA/SM, ring 2", {"ctas_per_sm": 1, "ring_depth": 2}),
                   ("hogwild stores on user rows", {"scatter_user": 0}), ("hogwild reference dot order", {"exact_dot": 1})):
    rows.append((name, api.MODE_HOGWILD, opts))
for name, mode, opts in rows:
    p1, p3, gi = (y1, y3, gy) if mode == api.MODE_EXACT else run(mode, opts)
    line = dict(setting=name, options=opts, ginst_s=gi,
                rmse_one_pass=float(np.sqrt(np.mean((p1.astype(np.float64) - y1) ** 2))),
                max_abs_one_pass=float(np.abs(p1 - y1).max()),
                rmse_3_epochs_heldout=float(np.sqrt(np.mean((p3.astype(np.float64) - y3) ** 2))),
                heldout_rmse_vs_labels=float(np.sqrt(np.mean((p3.astype(np.float64) - held[1]) ** 2))))
    print(json.dumps(line), flush=True)
    out.write(json.dumps(line) + "\n")
    out.flush()
