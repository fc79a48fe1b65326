"""GPU experiment: the ordered mode through the item-owner kernel (k_own, svdgpu_own.cu) on the
configs[1] shape.  Checks on a small prefix that k_own leaves the same model bytes as k_exact
(which the parity tests pin to the oracle), then times the plan build and the epochs at full size.

    python tools/own_study.py [rows] [name=value ...]      options: own_batch, own_urgent_gap, own_slots
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svdfeature_b200 import api, synth  # noqa: E402

NU, NI, K = 480000, 18000, 64
N = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 20_000_000
opts = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
NI = int(opts.pop("ni", NI))  # ni=1: every rating on one item = the bare chain of one owner
FLAGS = set(a for a in sys.argv[1:] if a in ("nocheck", "nohost", "stats"))
rng = np.random.default_rng(1)
W = (rng.standard_normal((NU + NI, K)) * 0.01).astype(np.float32)
data = synth.basic_mf(N, NU, NI, seed=3, zipf_q=70.0)
top = int(np.bincount(data[2][1::2]).max())
out = open(os.path.join(ROOT, "gpurun_out", "own_study.jsonl"), "a")


def trainer(owner):
    g = api.SvdGpu(NU, NI, K)
    g.set_hparams(learning_rate=0.005, wd_user=0.004, wd_item=0.004, base_score=3.6)
    g.set_mode(api.MODE_EXACT)
    g.set_option("exact_owner", owner)
    for k, v in opts.items():
        g.set_option(k, int(v))
    g.upload(np.zeros(NU + NI, np.float32), W, np.zeros(1, np.float32))
    return g


# ---- same bytes as k_exact on a prefix ---------------------------------------------------------
M = min(N, 1_000_000)
sub = (data[0][:3 * M + 1], data[1][:M], data[2][:2 * M], data[3][:2 * M])
models = []
for owner in (() if "nocheck" in FLAGS else (0, 1)):
    g = trainer(owner)
    b = g.batch_create(sub)
    g.timer_start()
    g.batch_update(b)
    ms = g.timer_stop()
    g.sync()
    models.append([a.copy() for a in g.download()])
    print(json.dumps(dict(check_rows=M, exact_owner=owner, ms=ms, minst_s=M / ms / 1e3,
                          own_launches=g.counter("own_launches"))), flush=True)
    b.close()
    g.close()
same = all(np.array_equal(a, c) for a, c in zip(*models)) if models else None
print(json.dumps(dict(same_model_as_k_exact=same)), flush=True)

# ---- full size ------------------------------------------------------------------------------------
g = trainer(1)
t0 = time.perf_counter()
b = g.batch_create(data)
g.sync()
t_create = time.perf_counter() - t0
times = []
for _ in range(4):
    g.timer_start()
    g.batch_update(b)
    times.append(g.timer_stop())
g.sync()
ms = float(np.median(times[1:]))
line = dict(rows=N, hottest_item_rows=top, batch_create_s=t_create, ms_epochs=times, ms=ms, minst_s=N / ms / 1e3,
            us_per_hot_row=1e3 * ms / top, same_model_as_k_exact=same, options=opts,
            own_launches=g.counter("own_launches"))
print(json.dumps(line), flush=True)
out.write(json.dumps(line) + "\n")
if "stats" in FLAGS:
    g.set_option("own_stats", 1)
    g.timer_start()
    g.batch_update(b)
    ms1 = g.timer_stop()
    st = g.own_stats()
    cnt = np.bincount(data[2][1::2], minlength=NI)
    order = np.argsort(-st[:, 0])
    rep = dict(ms_with_stats=ms1, mhz_assumed=1965)
    for name, sel in (("slowest16", order[:16]), ("owners0_15", np.arange(16)), ("median16", order[len(order) // 2 - 8:len(order) // 2 + 8])):
        rep[name] = dict(owner=sel.tolist(), us_total=(st[sel, 0] / 1965).round(0).tolist(), us_wait=(st[sel, 1] / 1965).round(0).tolist(),
                         us_flush=(st[sel, 2] / 1965).round(0).tolist(), waits=(st[sel, 3] & 0xffffffff).tolist(),
                         not_landed=(st[sel, 3] >> 32).tolist())
    rep["all"] = dict(us_total_max=float(st[:, 0].max() / 1965), us_total_mean=float(st[:, 0].mean() / 1965),
                      us_wait_mean=float(st[:, 1].mean() / 1965), us_flush_mean=float(st[:, 2].mean() / 1965),
                      waits_total=int((st[:, 3] & 0xffffffff).sum()), not_landed_total=int((st[:, 3] >> 32).sum()))
    print(json.dumps(rep), flush=True)
    out.write(json.dumps(rep) + "\n")
    g.set_option("own_stats", 0)
# host-pointer call (plan per chunk inside the call)
for chunk in (() if "nohost" in FLAGS else (1 << 20, 1 << 22, 1 << 23)):
    g.set_option("chunk_rows", chunk)
    c0 = (g.counter("own_lpt_us"), g.counter("own_cntwait_us"))
    t0 = time.perf_counter()
    g.update_csr(data)
    g.sync()
    dt = time.perf_counter() - t0
    line = dict(host_call_rows=N, chunk_rows=chunk, s=dt, minst_s=N / dt / 1e6, plan_host_lpt_ms=(g.counter("own_lpt_us") - c0[0]) / 1e3,
                plan_host_wait_counts_ms=(g.counter("own_cntwait_us") - c0[1]) / 1e3)
    print(json.dumps(line), flush=True)
    out.write(json.dumps(line) + "\n")
b.close()
g.close()
