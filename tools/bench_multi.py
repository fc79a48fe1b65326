"""BASELINE.json configs[3] and configs[4] on N GPUs of one box (run under torchrun): user-hash shards,
replicated item side, the library's NCCL exchange (svdgpu_allreduce_items) after every step.

  c4  pairwiseRank (BPR): 1M users x 300k items, k=128, rows (0 | 1 | 2) with values +-1, sigmoid-rank
      loss, no user bias.  Message: 300k x 129 floats = 154.8 MB.
  c5  neighbourhood model: 10M users x 1M items, 1M global (neighbour) features, 8 per row, k=256.
      The user slab is sharded N ways (1.25M users = 1.28 GB per rank at N=8), the item slab and the
      global biases are replicated.  Message: 1M x 257 + 1M floats = 1.03 GB.  Also streamed from pinned
      host memory through svdgpu_update_csr (e2e).

Hogwild mode (the ordered item-owner kernel covers the basic-MF shape only).  One JSON line per config on
rank 0, appended to gpurun_out/bench_multi_n<N>.jsonl.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_multi.py [c4 c5] [--rows R]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from svdfeature_b200 import api  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
DEV = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=DEV)


def log(*a):
    print("[rank %d]" % rank, *a, file=sys.stderr, flush=True)


def zipf_items(n, num_item, g, q=70.0):
    w = 1.0 / (torch.arange(1, num_item + 1, device=DEV, dtype=torch.float64) + q)
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    gp = torch.Generator(device=DEV)
    gp.manual_seed(1234)  # the same popularity order on every rank
    perm = torch.randperm(num_item, generator=gp, device=DEV)
    return perm[torch.searchsorted(cdf, torch.rand(n, generator=g, device=DEV, dtype=torch.float64)).clamp_(max=num_item - 1)]


def lognormal_users(n, num_user, g):
    w = torch.exp(torch.randn(num_user, generator=g, device=DEV, dtype=torch.float64))
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    return torch.searchsorted(cdf, torch.rand(n, generator=g, device=DEV, dtype=torch.float64)).clamp_(max=num_user - 1)


def fixed_csr(lab, cols_idx, cols_val, ng, nu, ni):
    n = lab.numel()
    per = ng + nu + ni
    index = torch.stack(cols_idx, 1).reshape(-1).to(torch.int32)
    value = torch.stack(cols_val, 1).reshape(-1).float()
    base = torch.arange(n, device=DEV, dtype=torch.int64) * per
    rp = torch.empty(3 * n + 1, device=DEV, dtype=torch.int32)
    rp[0:3 * n:3] = base.int()
    rp[1:3 * n:3] = (base + ng).int()
    rp[2:3 * n:3] = (base + ng + nu).int()
    rp[3 * n] = n * per
    return rp, lab, index, value


def pinned(ts):
    out = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t) for t in ts]
    torch.cuda.synchronize()
    return out


def init_model(g, rows, k, ngl, sigma=0.01):
    gen = torch.Generator(device=DEV)
    gen.manual_seed(10)  # replicas start equal (the user part differs by content only, which is fine)
    step = 1 << 20
    W = np.empty((rows, k), np.float32)
    for r0 in range(0, rows, step):
        r1 = min(rows, r0 + step)
        W[r0:r1] = (torch.randn(r1 - r0, k, generator=gen, device=DEV) * sigma).cpu().numpy()
    g.upload(np.zeros(rows, np.float32), W, np.zeros(max(ngl, 1), np.float32))


def join(g):
    if world > 1:
        ids = [api.comm_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        g.comm_init(world, rank, ids[0])
        g.allreduce_items(1.0 / world)  # snapshot


def run(g, name, host_csr, rows, nstep, bytes_per_row, shape, stream_e2e):
    """resident steps with an exchange after each; optionally the same rows streamed from pinned host memory"""
    stream = torch.cuda.Stream(device=DEV)
    g.set_stream(stream.cuda_stream)
    join(g)
    rp, lab, idx, val = host_csr
    b = g.batch_create((rp.numpy(), lab.numpy(), idx.numpy(), val.numpy()))
    per = (rp.numel() - 1) // 3 // nstep

    def step(s):
        g.batch_update(b, s * per, (s + 1) * per)

    def xchg():
        if world > 1:
            g.allreduce_items(1.0 / world)

    with torch.cuda.stream(stream):
        for s in range(2):
            step(s % nstep)
            xchg()
        g.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        c0, b0 = g.counter("collectives"), g.counter("collective_bytes")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for s in range(nstep):
            step(s)
            xchg()
        e1.record(stream)
        g.sync()
        ms = e0.elapsed_time(e1)
        ncoll, nbytes = g.counter("collectives") - c0, g.counter("collective_bytes") - b0
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(stream)
        for s in range(nstep):
            step(s)
        k1.record(stream)
        g.sync()
        kms = k0.elapsed_time(k1)
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        x0.record(stream)
        for _ in range(3):
            xchg()
        x1.record(stream)
        g.sync()
        xms = x0.elapsed_time(x1) / 3
    res = dict(ms=ms, kms=kms)
    e2e = None
    if stream_e2e:
        def e2e_step(s):
            r0, r1 = s * per, (s + 1) * per
            g.update_csr((rp[3 * r0:3 * r1 + 1], lab[r0:r1], idx, val))
            with torch.cuda.stream(stream):
                xchg()
        e2e_step(0)
        g.sync()
        if world > 1:
            dist.barrier()
        h0 = g.counter("h2d_bytes")
        t0 = time.perf_counter()
        for s in range(nstep):
            e2e_step(s)
        g.sync()
        e2e = dict(s=time.perf_counter() - t0, h2d_bytes_per_step=(g.counter("h2d_bytes") - h0) // nstep)
    if world > 1:
        tt = torch.tensor([res["ms"], res["kms"], xms, e2e["s"] if e2e else 0.0], device=DEV, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        res["ms"], res["kms"], xms = float(tt[0]), float(tt[1]), float(tt[2])
        if e2e:
            e2e["s"] = float(tt[3])
    ub, W, _ = g.download()
    finite = bool(np.isfinite(W).all())
    b.close()
    g.close()
    if rank == 0:
        total = world * per * nstep
        line = dict(config=name, n_gpus=world, shape=shape, mode="hogwild", rows_per_rank_per_step=per, steps=nstep,
                    ms_per_step=res["ms"] / nstep, ginst_s=total / res["ms"] / 1e6,
                    kernel_only_ms_per_step=res["kms"] / nstep, per_gpu_ginst_s=per * nstep / res["kms"] / 1e6,
                    algorithmic_gbs_per_gpu=per * nstep * bytes_per_row / res["kms"] / 1e6, bytes_per_row=bytes_per_row,
                    exchange=dict(ms=xms, message_bytes=(nbytes // max(ncoll, 1)), collectives_per_step=ncoll / nstep,
                                  algbw_gbs=(nbytes / max(ncoll, 1)) / max(xms, 1e-9) / 1e6,
                                  share_of_step=xms / (res["ms"] / nstep)),
                    model_finite=finite)
        if e2e:
            line["e2e"] = dict(ginst_s=total / e2e["s"] / 1e9, h2d_bytes_per_step_per_rank=e2e["h2d_bytes_per_step"],
                               what="the same rows streamed from pinned host memory through svdgpu_update_csr + the exchange")
        print(json.dumps(line), flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "bench_multi_n%d.jsonl" % world), "a") as f:
            f.write(json.dumps(line) + "\n")


def c4(rows):
    NU, NI, K = 1_000_000, 300_000, 128
    nu_local = (NU + world - 1) // world
    nstep = 3
    n = rows * nstep
    gen = torch.Generator(device=DEV)
    gen.manual_seed(40 + rank)
    u = lognormal_users(n, nu_local, gen)  # local ids of this rank's users
    pos = zipf_items(n, NI, gen)
    neg = torch.randint(0, NI, (n,), generator=gen, device=DEV)
    neg = torch.where(neg == pos, (neg + 1) % NI, neg)
    lo, hi = torch.minimum(pos, neg), torch.maximum(pos, neg)
    vlo = torch.where(pos < neg, 1.0, -1.0)
    ones = torch.ones(n, device=DEV)
    csr = pinned(fixed_csr(ones.clone(), [u, lo, hi], [ones, vlo, -vlo], 0, 1, 2))
    del u, pos, neg, lo, hi, vlo, ones
    torch.cuda.empty_cache()
    g = api.SvdGpu(nu_local, NI, K, no_user_bias=1, active_type=3, device=local)
    g.set_hparams(learning_rate=0.005, wd_user=0.004, wd_item=0.004, base_score=0.0)
    g.set_mode(api.MODE_HOGWILD)
    init_model(g, nu_local + NI, K, 0)
    run(g, "configs[3] pairwiseRank 1M x 300k, k=128", csr, rows, nstep, 3128,
        dict(num_user=NU, num_user_per_rank=nu_local, num_item=NI, k=K), stream_e2e=True)


def c5(rows):
    NU, NI, NGL, K, NG = 10_000_000, 1_000_000, 1_000_000, 256, 8
    nu_local = (NU + world - 1) // world
    nstep = 2
    n = rows * nstep
    gen = torch.Generator(device=DEV)
    gen.manual_seed(50 + rank)
    u = lognormal_users(n, nu_local, gen)
    it = zipf_items(n, NI, gen)
    gi = torch.sort(torch.randint(0, NGL, (n, NG), generator=gen, device=DEV, dtype=torch.int32), 1).values
    gv = 0.5 * torch.randn(n, NG, generator=gen, device=DEV)
    ones = torch.ones(n, device=DEV)
    lab = torch.clamp(torch.round(3.6 + 1.1 * torch.randn(n, generator=gen, device=DEV)), 1, 5).float()
    csr = pinned(fixed_csr(lab, [gi[:, j] for j in range(NG)] + [u, it], [gv[:, j] for j in range(NG)] + [ones, ones], NG, 1, 1))
    del u, it, gi, gv, ones, lab
    torch.cuda.empty_cache()
    g = api.SvdGpu(nu_local, NI, K, num_global=NGL, device=local)
    g.set_hparams(learning_rate=0.005, wd_user=0.004, wd_item=0.004, wd_global=0.001, base_score=3.6)
    g.set_mode(api.MODE_HOGWILD)
    init_model(g, nu_local + NI, K, NGL)
    run(g, "configs[4] neighbourhood 10M x 1M, 1M globals (8 per row), k=256", csr, rows, nstep, 4272,
        dict(num_user=NU, num_user_per_rank=nu_local, num_item=NI, num_global=NGL, k=K,
             user_slab_gb_per_rank=nu_local * K * 4 / 1e9, item_slab_gb=NI * K * 4 / 1e9), stream_e2e=True)


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if a in ("c4", "c5")] or ["c4", "c5"]
    rows = int(sys.argv[sys.argv.index("--rows") + 1]) if "--rows" in sys.argv else 0
    for nm in names:
        t0 = time.perf_counter()
        {"c4": c4, "c5": c5}[nm](rows or (20_000_000 if nm == "c4" else 12_500_000))
        log("%s done in %.1fs" % (nm, time.perf_counter() - t0))
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()
