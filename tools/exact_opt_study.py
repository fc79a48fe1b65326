"""GPU experiment: hand-off variants of the ordered kernel (option "exact_opt", bit mask:
1 = no per-lane fence before the release reds, 2 = poll back to back before sleeping,
4 = stage the instance's index/value slice in shared memory before the ticket wait).
Every variant must leave the same model bytes as variant 0 (which the parity tests pin to the
oracle); the time per row of the hottest item is the hand-off latency of the chain."""
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
data = synth.basic_mf(N, NU, NI, seed=3, zipf_q=70.0)
top = int(np.bincount(data[2][1::2]).max())
out = open(os.path.join(ROOT, "gpurun_out", "exact_opt_study.jsonl"), "a")
ref = None
# "owner" = the experimental item-owner kernel (option exact_owner) on top of the default exact_opt
for arg in sys.argv[1:] or ["0", "1", "2", "4", "3", "5", "6", "7"]:
    owner = arg == "owner"
    opt = 5 if owner else int(arg)
    g = api.SvdGpu(NU, NI, K)
    g.set_hparams(learning_rate=0.005, wd_user=0.004, wd_item=0.004, base_score=3.6)
    g.set_mode(api.MODE_EXACT)
    g.set_option("exact_opt", opt)
    g.set_option("exact_owner", 1 if owner else 0)
    g.upload(np.zeros(NU + NI, np.float32), W, np.zeros(1, np.float32))
    b = g.batch_create(data)
    g.batch_update(b)
    g.sync()
    model = [a.copy() for a in g.download()]
    if ref is None:
        ref = model
    same = all(np.array_equal(a, c) for a, c in zip(model, ref))
    g.timer_start()
    for _ in range(3):
        g.batch_update(b)
    ms = g.timer_stop() / 3
    line = dict(exact_opt=opt, exact_owner=owner, rows=N, hottest_item_rows=top, ms=ms, minst_s=N / ms / 1e3,
                us_per_hot_row=1e3 * ms / top, same_model_as_first=bool(same))
    print(json.dumps(line), flush=True)
    out.write(json.dumps(line) + "\n")
    out.flush()
    b.close()
    g.close()
