"""GPU experiment: where does k_mf's time go?  Times the configs[1] batch (20M resident rows)
as train (red / store scatter) and as predict-only (gathers + dot, no scatter), with any
extra `name=value` options from the command line applied first."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from svdfeature_b200 import api  # noqa: E402

N = 20_000_000
dev = torch.device("cuda", 0)
rp, lab, idx, val = bench.gen_rows_torch(N, 10, dev)
host = [t.cpu().numpy() for t in (rp, lab, idx, val)]
del rp, lab, idx, val
g = api.SvdGpu(bench.NUM_USER, bench.NUM_ITEM, bench.K)
g.set_hparams(**bench.HP)
g.set_mode(api.MODE_HOGWILD)
opts = [a for a in sys.argv[1:] if "=" in a]
for o in opts:
    k, v = o.split("=")
    g.set_option(k, int(v))
rng = np.random.default_rng(10)
rows = bench.NUM_USER + bench.NUM_ITEM
g.upload(np.zeros(rows, np.float32), (rng.standard_normal((rows, bench.K)) * 0.01).astype(np.float32),
         np.zeros(1, np.float32))
b = g.batch_create(tuple(host))


def timed(fn, reps=5):
    fn()
    g.sync()
    g.timer_start()
    for _ in range(reps):
        fn()
    return g.timer_stop() / reps


res = {"opts": opts}
res["train_red"] = N / timed(lambda: g.batch_update(b)) / 1e6
g.set_option("scatter_user", 0)
res["train_ustore_ired"] = N / timed(lambda: g.batch_update(b)) / 1e6
g.set_option("scatter_item", 0)
res["train_store"] = N / timed(lambda: g.batch_update(b)) / 1e6
res["predict"] = N / timed(lambda: g.batch_predict(b, fetch=False)) / 1e6
print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in res.items()}), "(G inst/s)")
