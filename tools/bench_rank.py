"""Ranker throughput: one svdgpu_rank_csr call over an item set and many user sections,
against the CPU oracle's restatement of SVDFeatureRanker on a sample of the same sections.
    python tools/bench_rank.py [num_items] [num_users] [k] [top_k]
Prints one JSON line; scores/s = (user, candidate) pairs scored and ranked per second."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from svdfeature_b200 import api, synth  # noqa: E402


def stream_of(n_items, n_users, num_user, seed, n_pos=5, n_ban=20):
    """ITEM rows (one item feature each = the catalogue) + USER/POS/BAN/PROCESS sections, vectorised."""
    rng = np.random.default_rng(seed)
    rows_i = n_items
    per = 4  # USER, POS, BAN, PROCESS
    n = rows_i + per * n_users
    cnt = np.zeros((n, 3), np.int64)
    cnt[:rows_i, 2] = 1
    u0 = rows_i + per * np.arange(n_users)
    cnt[u0, 1] = 1
    cnt[u0 + 1, 1] = n_pos
    cnt[u0 + 2, 1] = n_ban
    row_ptr = np.concatenate([[0], np.cumsum(cnt.reshape(-1))]).astype(np.int32)
    label = np.zeros(n, np.float32)
    label[u0] = synth.RK_USER
    label[u0 + 1] = synth.RK_POS
    label[u0 + 2] = synth.RK_BAN
    label[u0 + 3] = synth.RK_PROCESS
    tagged = np.argsort(rng.random((n_users, n_items)), axis=1)[:, :n_pos + n_ban].astype(np.uint32) if n_items < 50000 \
        else rng.integers(0, n_items, (n_users, n_pos + n_ban)).astype(np.uint32)
    users = rng.integers(0, num_user, n_users).astype(np.uint32)
    sec = np.concatenate([users[:, None], tagged], axis=1).reshape(-1)
    index = np.concatenate([np.arange(n_items, dtype=np.uint32), sec]).astype(np.uint32)
    value = np.ones(len(index), np.float32)
    return row_ptr, label, index, value


def main():
    n_items = int(sys.argv[1]) if len(sys.argv) > 1 else 18000
    n_users = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    top_k = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    num_user = 480000
    rng = np.random.default_rng(1)
    rows = num_user + n_items
    W = (rng.standard_normal((rows, k)) * 0.1).astype(np.float32)
    ub = (rng.standard_normal(rows) * 0.1).astype(np.float32)
    g = api.SvdGpu(num_user, n_items, k)
    g.set_hparams(base_score=3.6)
    g.upload(ub, W, np.zeros(0, np.float32))
    stream = stream_of(n_items, n_users, num_user, seed=2)
    out = {"exp": "rank", "items": n_items, "users": n_users, "k": k, "top_k": top_k}
    for rep in range(3):
        g.rank_init(n_items, top_k)
        t0 = time.perf_counter()
        got = g.rank(stream, cap=n_users * max(top_k, 8))
        dt = time.perf_counter() - t0
    out["gpu_s"] = dt
    out["gpu_scores_per_s"] = n_items * n_users / dt
    out["gpu_users_per_s"] = n_users / dt
    try:  # CPU oracle on a sample of the sections (test infrastructure; here only as the timed baseline)
        import tempfile
        from _oracle import COracle, COracleRanker
        o = COracle(0, 0, 0, dict(num_user=num_user, num_item=n_items, num_factor=k, base_score=3.6))
        o.init(1)
        a_ub, a_W, _ = o.arrays()
        a_ub[:] = ub
        a_W[:, :k] = W
        path = os.path.join(tempfile.mkdtemp(), "m.model")
        o.save_model(path)
        n_cpu = min(n_users, 100)
        sub = stream_of(n_items, n_users, num_user, seed=2)
        nrow = n_items + 4 * n_cpu
        sub = (sub[0][:3 * nrow + 1], sub[1][:nrow], sub[2], sub[3])
        r = COracleRanker(path, n_items, {"top_k": top_k})
        t0 = time.perf_counter()
        want = r.rank(sub)
        dtc = time.perf_counter() - t0
        out["cpu_users"] = n_cpu
        out["cpu_s"] = dtc
        out["cpu_scores_per_s"] = n_items * n_cpu / dtc
        out["match_on_sample"] = bool(np.array_equal(want, got[:len(want)]))
        out["speedup"] = out["gpu_scores_per_s"] / out["cpu_scores_per_s"]
    except Exception as e:  # noqa: BLE001
        out["cpu_error"] = repr(e)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
