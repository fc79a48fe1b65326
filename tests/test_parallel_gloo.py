"""CPU, world_size 2, gloo: the host logic of the multi-GPU layout (sharding by user hash and
the item-side delta exchange protocol).  The device side of the same protocol
(svdgpu_items_*) is covered by tests/test_gpu_parity.py::test_items_delta_protocol."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from svdfeature_b200 import parallel, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        csr = synth.basic_mf(5000, 300, 50, seed=7)
        mine, keep = parallel.shard_rows(csr, rank, world)
        # every row of my shard belongs to me and keeps its label / item
        uid = mine[2][mine[0][1::3][:len(mine[1])]]
        assert np.all(uid % world == rank)
        assert np.array_equal(mine[1], csr[1][keep])
        cnt = torch.tensor([len(keep)])
        dist.all_reduce(cnt)
        assert int(cnt) == 5000  # disjoint and complete
        # replicated item slab: snapshot + local "training" + one all-reduce of the deltas
        rng = np.random.default_rng(100)  # same start on every rank
        snap = rng.standard_normal(50 * 8).astype(np.float32)
        cur = snap.copy()
        np.add.at(cur, (mine[2][1::2] * 8) % len(cur), 0.01 * (rank + 1))  # rank-specific updates
        delta = torch.from_numpy(cur - snap)
        local = delta.clone()
        s = parallel.exchange_deltas(delta, dist, scale=1.0)
        new = snap + s * delta.numpy()
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        assert np.allclose(new, snap + sum(g.numpy() for g in gathered), atol=1e-6)
        check = torch.from_numpy(new.copy())
        dist.broadcast(check, 0)
        assert np.array_equal(check.numpy(), new)  # replicas agree bit for bit after the exchange
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_user_sharding_and_delta_exchange_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_rows_rejects_multi_user_rows():
    rows = synth.ragged_csr([(1.0, [], [(0, 1.0), (1, 1.0)], [(2, 1.0)])])
    with pytest.raises(ValueError):
        parallel.shard_rows(rows, 0, 2)


def test_shard_rows_keeps_globals():
    csr = synth.neighborhood(200, 40, 30, 12, ng=3, seed=2)
    parts = [parallel.shard_rows(csr, r, 3)[0] for r in range(3)]
    assert sum(len(p[1]) for p in parts) == 200
    assert sum(len(p[2]) for p in parts) == len(csr[2])
    for p in parts:
        n = len(p[1])
        assert np.all(np.diff(p[0].astype(np.int64)) >= 0) and p[0][3 * n] == len(p[2])
