"""GPU experiment: when does Hogwild diverge?  Ordered vs Hogwild held-out RMSE over three epochs for several
(k, learning rate, labels, options); `hogwild_safety=0` switches the stability guard off (the uncapped launch).
Finding that led to the guard: k = 32 / lr = 0.01 / 3000 items -> NaN uncapped; fine at k = 64, at lr = 0.005,
or with 1 CTA per SM and ring depth 2."""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from svdfeature_b200 import api, synth
nu, ni = 60000, 3000
for k, lr, planted, opts in ((32, 0.01, True, {"hogwild_safety": 0}), (32, 0.01, True, {}), (32, 0.02, True, {}), (32, 0.05, True, {}),
                             (64, 0.01, True, {}), (32, 0.005, True, {}), (32, 0.01, False, {}), (32, 0.01, True, {"ctas_per_sm": 1, "ring_depth": 2, "hogwild_safety": 0}),
                             (32, 0.01, True, {"scatter_user": 0, "scatter_item": 0}), (32, 0.02, True, {"ctas_per_sm": 1, "ring_depth": 2})):
    gen = synth.planted_mf if planted else synth.basic_mf
    train = gen(3000000, nu, ni, seed=1)
    held = gen(100000, nu, ni, seed=2)
    rng = np.random.default_rng(10)
    W0 = (rng.standard_normal((nu + ni, k)) * 0.01).astype(np.float32)
    res = {}
    for mode in (api.MODE_EXACT, api.MODE_HOGWILD):
        g = api.SvdGpu(nu, ni, k)
        g.set_hparams(learning_rate=lr, wd_user=0.004, wd_item=0.004, base_score=3.6)
        g.set_mode(mode)
        if mode == api.MODE_HOGWILD:
            for n_, v in opts.items():
                g.set_option(n_, v)
        g.upload(np.zeros(nu + ni, np.float32), W0, np.zeros(1, np.float32))
        c = []
        for e in range(3):
            g.update_csr(train)
            p = g.predict_csr(held)
            c.append(round(float(np.sqrt(np.mean((p - held[1]) ** 2))), 4))
        res[mode] = c
        cap = g.counter("inflight_cap")
        g.close()
    print(dict(k=k, lr=lr, planted=planted, opts=opts, ordered=res[0], hogwild=res[1], inflight_cap=cap), flush=True)
