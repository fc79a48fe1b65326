"""CPU study for DESIGN section 8, item 1: what would an ordered kernel with item-owner warps be
bounded by?  List-schedules the Netflix-shaped stream on W owner warps: an instance starts when
(a) its owner has finished the previous instance of its queue (input order) and (b) the previous
instance touching its user row has released it (cross-warp hand-off through L2).  Compared with
today's kernel, where BOTH rows are handed over through L2 (measured 5.9 us per hand-off).

    python tools/owner_order_model.py [rows] [warps]
"""
import heapq
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from svdfeature_b200 import synth  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 16
NU, NI = 480000, 18000
T_INST = 0.4  # us: gather of the (already released) user row is prefetched; compute + issue of the stores
T_HAND = 5.9  # us: measured hand-off of a row between warps (profiles/r1_exact_opt_study.jsonl)

data = synth.basic_mf(N, NU, NI, seed=3, zipf_q=70.0)
u = data[2][0::2].astype(np.int64)
it = data[2][1::2].astype(np.int64)
cnt = np.bincount(it, minlength=NI)

# owners: items in decreasing popularity to the least loaded warp (LPT)
owner = np.empty(NI, np.int64)
heap = [(0, w) for w in range(W)]
heapq.heapify(heap)
for i in np.argsort(-cnt):
    load, w = heapq.heappop(heap)
    owner[i] = w
    heapq.heappush(heap, (load + int(cnt[i]), w))
loads = np.bincount(owner[it], minlength=W)

ul, ol = u.tolist(), owner[it].tolist()
owner_free = [0.0] * W
user_free = [0.0] * NU
user_last_owner = [-1] * NU
end = 0.0
for a, w in zip(ul, ol):
    t = owner_free[w]
    ready = user_free[a] + (T_HAND if user_last_owner[a] not in (-1, w) else 0.0)
    if ready > t:
        t = ready
    t += T_INST
    owner_free[w] = t
    user_free[a] = t
    user_last_owner[a] = w
    if t > end:
        end = t
today = float(cnt.max()) * T_HAND
print(json.dumps(dict(rows=N, warps=W, hottest_item_rows=int(cnt.max()), heaviest_owner_rows=int(loads.max()),
                      mean_owner_rows=float(loads.mean()), today_us=today, today_minst_s=N / today,
                      owner_model_us=end, owner_model_minst_s=N / end,
                      floor_us=float(loads.max()) * T_INST, t_inst_us=T_INST, t_handoff_us=T_HAND)))
