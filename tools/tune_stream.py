"""GPU experiment: throughput of k_stream on the C2 workload for every scatter policy /
dot order / lane geometry, plus a Hogwild-vs-sequential parity study on a scaled-down
Netflix-shaped problem.  Writes JSON lines to gpurun_out/tune_stream.jsonl."""
import itertools
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from svdfeature_b200 import api, synth  # noqa: E402

out = open(os.path.join(ROOT, "gpurun_out", "tune_stream.jsonl"), "a")


def emit(**kw):
    print(json.dumps(kw), flush=True)
    out.write(json.dumps(kw) + "\n")
    out.flush()


def throughput():
    dev = torch.device("cuda", 0)
    rows, nchunk = 5_000_000, 8
    rp, lab, idx, val = bench.gen_rows_torch(rows * nchunk, 10, dev)
    host = [t.cpu().numpy() for t in (rp, lab, idx, val)]
    del rp, lab, idx, val
    rng = np.random.default_rng(10)
    W0 = (rng.standard_normal((bench.NUM_USER + bench.NUM_ITEM, 64)) * 0.01).astype(np.float32)
    g = api.SvdGpu(bench.NUM_USER, bench.NUM_ITEM, 64)
    g.set_hparams(**bench.HP)
    g.set_mode(api.MODE_HOGWILD)
    g.upload(np.zeros(len(W0), np.float32), W0, np.zeros(1, np.float32))
    b = g.batch_create(tuple(host))
    lanes_list = [int(x) for x in os.environ.get("TUNE_LANES", "8,16,4").split(",")]
    combos = [(1, 1, 0, l, 0) for l in lanes_list] + [(1, 1, 1, l, 0) for l in lanes_list]
    combos += [(0, 0, 0, lanes_list[0], 0), (0, 1, 0, lanes_list[0], 0), (1, 1, 0, lanes_list[0], 1)]
    for su, si, ed, lanes, cps in combos:
        g.set_option("scatter_user", su)
        g.set_option("scatter_item", si)
        g.set_option("exact_dot", ed)
        g.set_option("lanes", lanes)
        g.set_option("ctas_per_sm", cps)
        for s in range(2):
            g.batch_update(b, s * rows, (s + 1) * rows)
        g.sync()
        g.timer_start()
        for s in range(nchunk):
            g.batch_update(b, s * rows, (s + 1) * rows)
        ms = g.timer_stop() / nchunk
        emit(exp="throughput", scatter_user=su, scatter_item=si, exact_dot=ed, lanes=lanes, ctas_per_sm=cps,
             ms_per_5M=ms, ginst_s=rows / ms / 1e6, gbs=rows * 1072 / ms / 1e6)
    b.close()
    g.close()


def parity():
    from _oracle import COracle

    nu, ni, n, k = 120000, 18000, 3_000_000, 64
    params = dict(num_user=nu, num_item=ni, num_factor=k, learning_rate=0.005, wd_user=0.004, wd_item=0.004,
                  base_score=3.6)
    train = synth.basic_mf(n, nu, ni, seed=21)
    test = synth.basic_mf(300000, nu, ni, seed=22)
    epochs = 5
    o = COracle(0, 0, 0, params)
    o.init(10)
    init = [a.copy() for a in o.arrays()]
    seq = []
    t0 = time.perf_counter()
    for e in range(epochs):
        o.update_csr(train)
        p = o.predict_csr(test)
        seq.append(p)
        emit(exp="parity", who="oracle", epoch=e + 1, test_rmse=float(np.sqrt(np.mean((p - test[1]) ** 2))))
    emit(exp="parity", who="oracle", seconds=time.perf_counter() - t0)
    for mode, su, si in [(api.MODE_EXACT, 0, 0), (api.MODE_HOGWILD, 0, 0), (api.MODE_HOGWILD, 0, 1),
                         (api.MODE_HOGWILD, 1, 1)]:
        g = api.SvdGpu(nu, ni, k)
        g.set_hparams(learning_rate=0.005, wd_user=0.004, wd_item=0.004, base_score=o.base_score)
        g.set_mode(mode)
        g.set_option("scatter_user", su)
        g.set_option("scatter_item", si)
        g.set_option("chunk_rows", 1 << 22)
        g.upload(*init)
        for e in range(epochs):
            t0 = time.perf_counter()
            g.update_csr(train)
            g.sync()
            dt = time.perf_counter() - t0
            p = g.predict_csr(test)
            emit(exp="parity", who="gpu", mode=mode, scatter_user=su, scatter_item=si, epoch=e + 1,
                 test_rmse=float(np.sqrt(np.mean((p - test[1]) ** 2))),
                 rmse_vs_sequential=float(np.sqrt(np.mean((p - seq[e]) ** 2))),
                 max_abs_vs_sequential=float(np.abs(p - seq[e]).max()), epoch_seconds=dt)
        g.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["throughput", "parity"]
    if "throughput" in which:
        throughput()
    if "parity" in which:
        parity()
