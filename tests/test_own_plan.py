"""Host part of the ordered mode with item-owner warps (svdfeature_b200/csrc/svdgpu_ownplan.h):
the LPT deal gives every item one owner, numbers an owner's items by popularity and balances the
load; a protocol-level simulation of k_own (svdgpu_own.cu: per-owner queues in input order, ring
slots filled by loader lanes once the user's version has reached the ticket, publishes held back
in batches and flushed before an owner blocks) under random schedules reproduces the sequential
order on every row and never deadlocks."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from svdfeature_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = r"""
#include "svdgpu_ownplan.h"
extern "C" long long own_assign(const unsigned *cnt, int num_item, int num_owner, int max_batch, int *item_owner,
                                unsigned *item_slot, int *queue_off, int *item_off, unsigned *items, int *batch) {
  svdown::HostPlan p;
  svdown::assign(cnt, num_item, num_owner, max_batch, p);
  for (int i = 0; i < num_item; ++i) { item_owner[i] = p.item_owner[i]; item_slot[i] = p.item_slot[i]; }
  for (int w = 0; w <= num_owner; ++w) { queue_off[w] = p.queue_off[w]; item_off[w] = p.item_off[w]; }
  for (size_t i = 0; i < p.items.size(); ++i) items[i] = p.items[i];
  for (int w = 0; w < num_owner; ++w) batch[w] = p.batch[w];
  return p.max_load;
}
extern "C" int own_hot_owners(const unsigned *cnt, int num_item, int num_owner, int percent) {
  return svdown::hot_owners(cnt, num_item, num_owner, percent);
}
extern "C" int own_top_owners(const unsigned *cnt, int num_item, int percent) { return svdown::top_owners(cnt, num_item, percent); }
// deal on cnt_a (all items get owners), carry over to cnt_b: returns 1 if carried; outputs describe the plan for cnt_b
extern "C" int own_redeal(const unsigned *cnt_a, const unsigned *cnt_b, int num_item, int num_owner, int slack, int *item_owner,
                          int *queue_off, int *batch, long long *max_load) {
  svdown::HostPlan p;
  svdown::assign(cnt_a, num_item, num_owner, 16, p, nullptr, true);
  const int ok = svdown::redeal(cnt_b, num_item, 16, slack, p) ? 1 : 0;
  for (int i = 0; i < num_item; ++i) item_owner[i] = p.item_owner[i];
  for (int w = 0; w <= num_owner; ++w) queue_off[w] = p.queue_off[w];
  for (int w = 0; w < num_owner; ++w) batch[w] = p.batch[w];
  *max_load = p.max_load;
  return ok;
}
extern "C" long long own_assign_closed(const unsigned *cnt, int num_item, int num_owner, const char *closed, int *item_owner,
                                       int *queue_off) {
  svdown::HostPlan p;
  std::vector<char> c(closed, closed + num_owner);
  svdown::assign(cnt, num_item, num_owner, 16, p, &c);
  for (int i = 0; i < num_item; ++i) item_owner[i] = p.item_owner[i];
  for (int w = 0; w <= num_owner; ++w) queue_off[w] = p.queue_off[w];
  return p.max_load;
}
"""


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("own")
    src = d / "shim.cpp"
    src.write_text(SHIM)
    so = d / "libown.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "svdfeature_b200", "csrc"),
                           "-o", str(so), str(src)])
    lib = C.CDLL(str(so))
    lib.own_assign.restype = C.c_longlong
    lib.own_assign.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6
    lib.own_hot_owners.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.own_top_owners.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.own_assign_closed.restype = C.c_longlong
    lib.own_assign_closed.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.own_redeal.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def _assign(lib, cnt, num_owner, max_batch=16):
    cnt = np.ascontiguousarray(cnt, np.uint32)
    ni = len(cnt)
    owner = np.zeros(ni, np.int32)
    slot = np.zeros(ni, np.uint32)
    qo = np.zeros(num_owner + 1, np.int32)
    io = np.zeros(num_owner + 1, np.int32)
    items = np.zeros(max(int((cnt > 0).sum()), 1), np.uint32)
    batch = np.zeros(num_owner, np.int32)
    ml = lib.own_assign(cnt.ctypes.data, ni, num_owner, max_batch, owner.ctypes.data, slot.ctypes.data, qo.ctypes.data,
                        io.ctypes.data, items.ctypes.data, batch.ctypes.data)
    return owner, slot, qo, io, items, batch, ml


@pytest.mark.parametrize("num_owner", [1, 7, 64, 2368])
def test_assign_deals_items_out(lib, num_owner):
    nu, ni, n = 5000, 300, 40000
    data = synth.basic_mf(n, nu, ni, seed=4, zipf_q=5.0)
    cnt = np.bincount(data[2][1::2], minlength=ni)
    owner, slot, qo, io, items, batch, ml = _assign(lib, cnt, num_owner)
    assert qo[0] == 0 and qo[-1] == n and np.all(np.diff(qo) >= 0)
    assert io[0] == 0 and io[-1] == (cnt > 0).sum()
    assert np.all(owner[cnt == 0] == -1) and np.all(owner[cnt > 0] >= 0)
    loads = np.zeros(num_owner, np.int64)
    for w in range(num_owner):
        mine = items[io[w]:io[w + 1]]
        assert np.all(owner[mine] == w)
        assert np.array_equal(slot[mine], np.arange(len(mine)))  # slot = position in the owner's list
        assert np.all(np.diff(cnt[mine].astype(np.int64)) <= 0)  # most popular first
        loads[w] = cnt[mine].sum()
    assert np.array_equal(loads, np.diff(qo))
    assert ml == loads.max()
    # LPT: the heaviest owner carries at most the hottest item or 4/3 of the mean
    assert loads.max() <= max(cnt.max(), int(np.ceil(4 / 3 * n / num_owner)) + 1)
    # the hottest items sit on different owners, dealt out in owner order (-> different SMs)
    top = np.argsort(-cnt, kind="stable")[:min(num_owner, (cnt > 0).sum())]
    assert np.array_equal(owner[top], np.arange(len(top)))
    # publish batches: proportional to the load, the busiest owner gets the maximum
    assert batch.max() == 16 and batch.min() >= 1
    assert np.all(batch == np.clip(16 * loads // max(loads.max(), 1), 1, 16))


def test_assign_empty_and_single(lib):
    owner, slot, qo, io, items, batch, ml = _assign(lib, np.zeros(10, np.uint32), 4)
    assert np.all(owner == -1) and qo[-1] == 0 and io[-1] == 0 and ml == 0 and np.all(batch == 1)
    owner, slot, qo, io, items, batch, ml = _assign(lib, np.array([0, 5, 0], np.uint32), 3)
    assert owner.tolist() == [-1, 0, -1] and qo.tolist() == [0, 5, 5, 5] and ml == 5


@pytest.mark.parametrize("seed", range(6))
def test_protocol_reproduces_the_sequential_order(lib, seed):
    """k_own's protocol, step by step under a random scheduler.  Owner: takes the head of its queue
    when its ring slot is full, writes the user row, holds the publish back (batch / urgent flag)
    and flushes before it blocks.  Loader lane of slot s: fills entry j (j mod D = s) once the slot
    is vacant and the user's PUBLISHED version equals the ticket."""
    nu, ni, n, W, D = 40, 25, 3000, 12, 4
    data = synth.basic_mf(n, nu, ni, seed=seed, zipf_q=1.0)
    users, items_of = data[2][0::2].astype(int), data[2][1::2].astype(int)
    cnt = np.bincount(items_of, minlength=ni)
    owner, slot, qo, io, items, batch, ml = _assign(lib, cnt, W, max_batch=8)
    queues = [[] for _ in range(W)]
    for r in range(n):
        queues[owner[items_of[r]]].append(r)
    assert [len(q) for q in queues] == np.diff(qo).tolist()
    ticket = np.zeros(n, int)
    nxt_gap = np.full(n, 1 << 30)
    last = {}
    for r in range(n):
        u = users[r]
        ticket[r] = 0 if u not in last else ticket[last[u]] + 1
        if u in last:
            nxt_gap[last[u]] = r - last[u]
        last[u] = r
    urgent = nxt_gap < 50
    rng = np.random.default_rng(seed)
    head = [0] * W                 # next entry the owner consumes
    loaded = [0] * W               # entries [0, loaded) ... tracked per slot below
    slot_entry = [[-1] * D for _ in range(W)]   # entry index sitting in the slot (full) or -1 (vacant)
    next_fill = [[s for s in range(D)] for _ in range(W)]  # entry each loader lane loads next
    ver = np.zeros(nu, int)        # published versions
    pend = [[] for _ in range(W)]  # held-back publishes (user, version)
    user_log = [[] for _ in range(nu)]
    item_log = [[] for _ in range(ni)]
    done = 0

    def flush(w):
        for u, v in pend[w]:
            ver[u] = v
        pend[w].clear()

    idle_rounds = 0
    while done < n:
        progressed = False
        for w in rng.permutation(W):
            act = rng.integers(0, 3)
            if act == 0:  # a loader lane of this owner tries
                s = int(rng.integers(0, D))
                j = next_fill[w][s]
                if j < len(queues[w]) and slot_entry[w][s] == -1:
                    r = queues[w][j]
                    if ver[users[r]] == ticket[r]:
                        slot_entry[w][s] = j
                        next_fill[w][s] = j + D
                        progressed = True
            else:  # the owner tries
                j = head[w]
                if j >= len(queues[w]):
                    if pend[w]:
                        flush(w)
                        progressed = True
                    continue
                s = j % D
                if slot_entry[w][s] != j:
                    if pend[w]:  # about to block: publish first
                        flush(w)
                        progressed = True
                    continue
                r = queues[w][j]
                slot_entry[w][s] = -1
                user_log[users[r]].append(r)
                item_log[items_of[r]].append(r)
                pend[w].append((users[r], ticket[r] + 1))
                if urgent[r] or len(pend[w]) >= batch[w]:
                    flush(w)
                head[w] += 1
                done += 1
                progressed = True
        idle_rounds = 0 if progressed else idle_rounds + 1
        assert idle_rounds < 200, "deadlock"
    for log in user_log + item_log:
        assert log == sorted(log)  # every row saw its writers in input order


def test_closed_owners_get_nothing_and_hot_items_are_counted(lib):
    """Issue-port / SM isolation of the hot chains (svdgpu_own.cu, own_plan_build): owners marked closed get no
    items, everything else is dealt out as before; hot_owners / top_owners count the items that qualify."""
    ni, num_owner = 400, 48
    cnt = (100000.0 / (np.arange(ni) + 3.0)).astype(np.uint32)  # Zipf-like, item 0 the hottest
    total = int(cnt.sum())
    hot = lib.own_hot_owners(cnt.ctypes.data, ni, num_owner, 200)
    assert hot == int((cnt.astype(np.int64) * num_owner * 100 > total * 200).sum()) and 0 < hot < num_owner
    top = lib.own_top_owners(cnt.ctypes.data, ni, 75)
    assert top == int((cnt >= 0.75 * cnt.max()).sum()) and top >= 1
    assert lib.own_hot_owners(np.full(ni, 50, np.uint32).ctypes.data, ni, num_owner, 200) == 0  # flat popularity: nobody is hot
    closed = np.zeros(num_owner, np.int8)
    closed[[5, 17, 40]] = 1
    owner = np.zeros(ni, np.int32)
    qo = np.zeros(num_owner + 1, np.int32)
    ml = lib.own_assign_closed(cnt.ctypes.data, ni, num_owner, closed.ctypes.data, owner.ctypes.data, qo.ctypes.data)
    loads = np.diff(qo)
    assert np.all(loads[closed == 1] == 0) and not np.isin(owner, [5, 17, 40]).any()
    assert loads.sum() == total and ml == loads.max() == cnt.max()
    assert np.array_equal(owner[:5], np.arange(5)) and owner[5] == 6  # the r-th hottest item on the r-th open owner


def test_redeal_carries_a_deal_over_while_it_is_good(lib):
    """svdown::redeal: the same owners for new counts of the same distribution (queue offsets, loads and batch
    sizes follow the new counts; items the first batch never touched have owners too); refused when the
    popularity has moved."""
    ni, num_owner = 600, 32
    rng = np.random.default_rng(5)
    p = 1.0 / (np.arange(ni) + 5.0)
    p /= p.sum()
    a = rng.multinomial(200000, p).astype(np.uint32)
    b = rng.multinomial(200000, p).astype(np.uint32)
    a[-40:] = 0  # untouched by the first batch, touched by the second
    owner = np.zeros(ni, np.int32)
    qo = np.zeros(num_owner + 1, np.int32)
    batch = np.zeros(num_owner, np.int32)
    ml = C.c_longlong(0)
    ok = lib.own_redeal(a.ctypes.data, b.ctypes.data, ni, num_owner, 25, owner.ctypes.data, qo.ctypes.data, batch.ctypes.data,
                        C.byref(ml))
    assert ok == 1 and np.all(owner >= 0)
    loads = np.bincount(owner, weights=b, minlength=num_owner).astype(np.int64)
    assert np.array_equal(np.diff(qo), loads) and ml.value == loads.max()
    assert loads.max() <= max(b.max(), 1.25 * np.ceil(b.sum() / num_owner))
    assert batch.min() >= 1 and batch.max() == 16 and batch[np.argmax(loads)] == 16
    # the popularity moves to the other end of the catalogue: the old deal piles the load on a few owners
    moved = b[::-1].copy()
    ok = lib.own_redeal(a.ctypes.data, moved.ctypes.data, ni, num_owner, 25, owner.ctypes.data, qo.ctypes.data,
                        batch.ctypes.data, C.byref(ml))
    assert ok == 0
