"""Host side of the experimental item-owner ordered mode (svdfeature_b200/csrc/svdgpu_owner.h):
the per-owner queues partition the rows, keep input order, give every item one owner and balance
the load; a protocol-level simulation (random warp scheduling, user tickets) reproduces the
sequential order on every touched row and never deadlocks."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from svdfeature_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = r"""
#include "svdgpu_owner.h"
extern "C" int owner_plan(int n, const int *rp, const unsigned *idx, int num_item, int num_owner, int *queue_off, int *queue) {
  svdowner::Plan p;
  if (!svdowner::build_plan(n, rp, idx, num_item, num_owner, p)) return 0;
  for (size_t i = 0; i < p.queue_off.size(); ++i) queue_off[i] = p.queue_off[i];
  for (size_t i = 0; i < p.queue.size(); ++i) queue[i] = p.queue[i];
  return 1;
}
"""


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("owner")
    src = d / "shim.cpp"
    src.write_text(SHIM)
    so = d / "libowner.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "svdfeature_b200", "csrc"),
                           "-o", str(so), str(src)])
    lib = C.CDLL(str(so))
    lib.owner_plan.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return lib


def _plan(lib, data, num_item, num_owner):
    rp, lab, idx, val = data
    n = len(lab)
    qo = np.zeros(num_owner + 1, np.int32)
    q = np.full(n, -1, np.int32)
    ok = lib.owner_plan(n, rp.ctypes.data, idx.ctypes.data, num_item, num_owner, qo.ctypes.data, q.ctypes.data)
    return bool(ok), qo, q


@pytest.mark.parametrize("num_owner", [1, 7, 64, 2368])
def test_plan_partitions_rows_in_order(lib, num_owner):
    nu, ni, n = 5000, 300, 40000
    data = synth.basic_mf(n, nu, ni, seed=4, zipf_q=5.0)
    ok, qo, q = _plan(lib, data, ni, num_owner)
    assert ok
    assert qo[0] == 0 and qo[-1] == n and np.all(np.diff(qo) >= 0)
    assert np.array_equal(np.sort(q), np.arange(n))  # every row exactly once
    items = data[2][1::2]
    owner_of_item = {}
    for w in range(num_owner):
        rows = q[qo[w]:qo[w + 1]]
        assert np.all(np.diff(rows) > 0)  # input order inside an owner
        for i in np.unique(items[rows]):
            assert owner_of_item.setdefault(int(i), w) == w  # one owner per item
    cnt = np.bincount(items, minlength=ni)
    loads = np.diff(qo)
    # LPT: the heaviest owner carries at most the hottest item or 4/3 of the mean
    assert loads.max() <= max(cnt.max(), int(np.ceil(4 / 3 * n / num_owner)) + 1)


def test_plan_refuses_other_shapes(lib):
    ok, _, _ = _plan(lib, synth.random_general(200, 50, 40, 10, seed=1), 40, 8)
    assert not ok
    d = synth.basic_mf(100, 50, 40, seed=2)
    assert _plan(lib, d, 40, 8)[0]
    assert not _plan(lib, d, 10, 8)[0]  # an item index beyond num_item


@pytest.mark.parametrize("seed", range(5))
def test_protocol_reproduces_the_sequential_order(lib, seed):
    """Warps step in random order; an instance runs when it is at the head of its owner's queue
    and its user row has seen all earlier writers (ticket).  The per-row sequences of writers must
    be those of the sequential loop, and some warp can always make progress."""
    nu, ni, n, W = 40, 25, 3000, 12
    data = synth.basic_mf(n, nu, ni, seed=seed, zipf_q=1.0)
    ok, qo, q = _plan(lib, data, ni, W)
    assert ok
    users, items = data[2][0::2].astype(int), data[2][1::2].astype(int)
    ticket = np.zeros(n, int)
    seen = np.zeros(nu, int)
    for r in range(n):
        ticket[r] = seen[users[r]]
        seen[users[r]] += 1
    rng = np.random.default_rng(seed)
    head = qo[:-1].astype(int).copy()
    ver = np.zeros(nu, int)
    user_log = [[] for _ in range(nu)]
    item_log = [[] for _ in range(ni)]
    done = 0
    while done < n:
        progressed = False
        for w in rng.permutation(W):
            if head[w] == qo[w + 1]:
                continue
            r = int(q[head[w]])
            if ver[users[r]] != ticket[r]:
                continue  # blocked on the user row
            user_log[users[r]].append(r)
            item_log[items[r]].append(r)
            ver[users[r]] += 1
            head[w] += 1
            done += 1
            progressed = True
            if rng.random() < 0.5:
                break  # another random schedule
        assert progressed, "deadlock"
    for log in user_log + item_log:
        assert log == sorted(log)  # every row saw its writers in input order
