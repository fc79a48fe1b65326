import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, _ml100k
from svdfeature_b200 import api
train, test, truth, gold = _ml100k.load()
want = {int(k): v for k, v in gold["test_rmse_after_round"].items()}
for rep in range(6):
    g = api.GpuTrainer(0, 0, 0, dict(gold["params"], **{"gpu:mode": "hogwild"}))
    curve, _ = _ml100k.run(g, train, test, truth, set(want), gold["seed"], None)
    g.close()
    print(rep, {r: round(curve[r] - v, 5) for r, v in want.items()}, flush=True)
