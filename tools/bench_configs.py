"""GPU experiment: kernel-resident throughput of BASELINE.json configs[2..4] (SVD++ user blocks,
pairwise rows, neighbourhood rows with globals), beside configs[1].  Not the headline bench
(bench.py is); writes JSON lines to gpurun_out/bench_configs.jsonl.

    python tools/bench_configs.py [c2 c3 c4 c5] [--scale 0.2]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from svdfeature_b200 import api  # noqa: E402

DEV = torch.device("cuda", 0)
out = open(os.path.join(ROOT, "gpurun_out", "bench_configs.jsonl"), "a")


def emit(**kw):
    print(json.dumps(kw), flush=True)
    out.write(json.dumps(kw) + "\n")
    out.flush()


def zipf_items(n, num_item, g, q=70.0):
    w = 1.0 / (torch.arange(1, num_item + 1, device=DEV, dtype=torch.float64) + q)
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    perm = torch.randperm(num_item, generator=g, device=DEV)
    return perm[torch.searchsorted(cdf, torch.rand(n, generator=g, device=DEV, dtype=torch.float64)).clamp_(max=num_item - 1)]


def lognormal_users(n, num_user, g):
    w = torch.exp(torch.randn(num_user, generator=g, device=DEV, dtype=torch.float64))
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    return torch.searchsorted(cdf, torch.rand(n, generator=g, device=DEV, dtype=torch.float64)).clamp_(max=num_user - 1)


def labels(n, g):
    return torch.clamp(torch.round(3.6 + 1.1 * torch.randn(n, generator=g, device=DEV)), 1, 5).float()


def fixed_csr(lab, cols_idx, cols_val, ng, nu, ni):
    """rows with fixed feature counts -> CSR tensors (index as int32 bits)."""
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


def host(ts):
    return tuple(t.cpu().numpy() for t in ts)


def init_model(g, rows, k, sigma=0.01, seed=10):
    gen = torch.Generator(device=DEV)
    gen.manual_seed(seed)
    step = 1 << 20
    W = np.empty((rows, k), np.float32)
    for r0 in range(0, rows, step):
        r1 = min(rows, r0 + step)
        W[r0:r1] = (torch.randn(r1 - r0, k, generator=gen, device=DEV) * sigma).cpu().numpy()
    g.upload(np.zeros(rows, np.float32), W, np.zeros(max(g.shape.num_global, 1), np.float32))


def apply_opts(g):
    """--opt name=value (repeatable) -> svdgpu_set_option"""
    opts = {}
    for i, a in enumerate(sys.argv):
        if a == "--opt":
            k, v = sys.argv[i + 1].split("=")
            g.set_option(k, int(v))
            opts[k] = int(v)
    return opts


def run_steps(g, batch, nstep, rows_per_step, bytes_per_row, name, units=None, extra=None):
    extra = dict(extra or {}, opts=apply_opts(g))
    def step(s):
        if units is None:
            g.batch_update(batch, s * rows_per_step, (s + 1) * rows_per_step)
        else:
            g.batch_update(batch, units[s], units[s + 1])

    for s in range(min(2, nstep)):
        step(s)
    g.sync()
    g.timer_start()
    for s in range(nstep):
        step(s)
    ms = g.timer_stop()
    g.sync()
    rows = rows_per_step * nstep
    emit(exp="config", config=name, rows=rows, ms=ms, ginst_s=rows / ms / 1e6,
         algorithmic_gbs=rows * bytes_per_row / ms / 1e6, bytes_per_row=bytes_per_row, **(extra or {}))


def c2(scale):
    import bench

    n = int(40_000_000 * scale)
    rows = 5_000_000
    n = max(rows * 2, n // rows * rows)
    rp, lab, idx, val = bench.gen_rows_torch(n, 10, DEV)
    g = api.SvdGpu(bench.NUM_USER, bench.NUM_ITEM, 64)
    g.set_hparams(**bench.HP)
    g.set_mode(api.MODE_HOGWILD)
    init_model(g, bench.NUM_USER + bench.NUM_ITEM, 64)
    b = g.batch_create(host((rp, lab, idx, val)))
    run_steps(g, b, n // rows, rows, 1072, "c2 basicMF 480kx18k k=64")
    b.close()
    g.close()


def c4(scale):
    nu_, ni_, k = 1_000_000, 300_000, 128
    n = int(30_000_000 * scale)
    rows = 2_000_000
    n = max(rows * 2, n // rows * rows)
    gen = torch.Generator(device=DEV)
    gen.manual_seed(4)
    u = lognormal_users(n, nu_, gen)
    pos = zipf_items(n, ni_, gen)
    neg = torch.randint(0, ni_, (n,), generator=gen, device=DEV)
    neg = torch.where(neg == pos, (neg + 1) % ni_, neg)
    lo, hi = torch.minimum(pos, neg), torch.maximum(pos, neg)
    vlo = torch.where(pos < neg, 1.0, -1.0)
    ones = torch.ones(n, device=DEV)
    csr = fixed_csr(ones.clone(), [u, lo, hi], [ones, vlo, -vlo], 0, 1, 2)
    g = api.SvdGpu(nu_, ni_, k, no_user_bias=1, active_type=3)
    g.set_hparams(learning_rate=0.005, wd_user=0.004, wd_item=0.004, base_score=0.0)
    g.set_mode(api.MODE_HOGWILD)
    init_model(g, nu_ + ni_, k)
    b = g.batch_create(host(csr))
    run_steps(g, b, n // rows, rows, 3128, "c4 pairwise 1Mx300k k=128 (rows 0|1|2, +-1)")
    b.close()
    g.close()


def c4u(scale):
    """configs[3] the way the reference runs it: rated user blocks -> pairwise-rank samples (on the
    device, svdgpu_batch_sample_pairs) -> training on the pair blocks."""
    nu_, ni_, k = 1_000_000, 300_000, 128
    n = int(40_000_000 * scale)
    gen = torch.Generator(device=DEV)
    gen.manual_seed(4)
    u = torch.sort(lognormal_users(n, nu_, gen)).values
    it = zipf_items(n, ni_, gen)
    lab = (torch.rand(n, generator=gen, device=DEV) < 0.3).float()
    ones = torch.ones(n, device=DEV)
    csr = host(fixed_csr(lab, [u, it], [ones, ones], 0, 1, 1))
    uu, cnt = torch.unique_consecutive(u, return_counts=True)
    bro = np.concatenate([[0], np.cumsum(cnt.cpu().numpy())]).astype(np.int32)
    nb = len(bro) - 1
    ug = (bro, np.zeros(nb + 1, np.int32), None, np.zeros(0, np.uint32), np.zeros(0, np.float32))
    g = api.SvdGpu(nu_, ni_, k, num_ufeedback=1, no_user_bias=1, active_type=3, format_type=1)
    g.set_hparams(learning_rate=0.005, wd_user=0.004, wd_item=0.004, base_score=0.0)
    g.set_mode(api.MODE_HOGWILD)
    init_model(g, 1 + nu_ + ni_, k)
    opts = apply_opts(g)
    src = g.batch_create(csr, ugroup=ug)
    g.batch_sample_pairs(src, seed=0).close()  # first call: module loading of the sampler's kernels
    g.sync()
    t0 = time.perf_counter()
    pairs = g.batch_sample_pairs(src, seed=1)
    g.sync()
    t_sample = time.perf_counter() - t0
    for _ in range(2):
        g.batch_update(pairs)
    g.sync()
    g.timer_start()
    reps = 5
    for _ in range(reps):
        g.batch_update(pairs)
    ms = g.timer_stop() / reps
    emit(exp="config", config="c4u pairwise via user blocks + device sampler 1Mx300k k=128", rated_rows=n, blocks=nb,
         pairs=pairs.num_row, sample_ms=1e3 * t_sample, sample_gpairs_s=pairs.num_row / t_sample / 1e9, train_ms=ms,
         ginst_s=pairs.num_row / ms / 1e6, algorithmic_gbs=pairs.num_row * 3128 / ms / 1e6, bytes_per_row=3128, opts=opts)
    g.close()


def c5(scale):
    nu_, ni_, ngl, k, ng = 10_000_000, 1_000_000, 1_000_000, 256, 8
    n = int(20_000_000 * scale)
    rows = 1_000_000
    n = max(rows * 2, n // rows * rows)
    gen = torch.Generator(device=DEV)
    gen.manual_seed(5)
    u = lognormal_users(n, nu_, gen)
    it = zipf_items(n, ni_, gen)
    gi = torch.sort(torch.randint(0, ngl, (n, ng), generator=gen, device=DEV), 1).values
    gv = 0.5 * torch.randn(n, ng, generator=gen, device=DEV)
    ones = torch.ones(n, device=DEV)
    csr = fixed_csr(labels(n, gen), [gi[:, j] for j in range(ng)] + [u, it], [gv[:, j] for j in range(ng)] + [ones, ones],
                    ng, 1, 1)
    g = api.SvdGpu(nu_, ni_, k, num_global=ngl)
    g.set_hparams(learning_rate=0.005, wd_user=0.004, wd_item=0.004, wd_global=0.001, base_score=3.6)
    g.set_mode(api.MODE_HOGWILD)
    init_model(g, nu_ + ni_, k)
    b = g.batch_create(host(csr))
    run_steps(g, b, n // rows, rows, 4272, "c5 neighbourhood 10Mx1M k=256, 8 globals/row")
    b.close()
    g.close()


def c3(scale):
    """user-grouped rows + per-user feedback list (rated items + random extras, ~200 entries)."""
    nu_, ni_, k, fb_target = 480_000, 18_000, 64, 200
    n = int(100_000_000 * scale)
    gen = torch.Generator(device=DEV)
    gen.manual_seed(3)
    u = lognormal_users(n, nu_, gen)
    it = zipf_items(n, ni_, gen)
    lab = labels(n, gen)
    uperm = torch.randperm(nu_, generator=gen, device=DEV)  # user processing order (shuffled users)
    key = uperm[u] * (1 << 20) + torch.randint(0, 1 << 20, (n,), generator=gen, device=DEV)
    order = torch.argsort(key)
    u, it, lab = u[order], it[order], lab[order]
    pos = uperm[u]  # non-decreasing
    cnt = torch.bincount(pos, minlength=nu_)
    present = cnt > 0
    blk_rows = cnt[present]
    blk_row_off = torch.zeros(blk_rows.numel() + 1, dtype=torch.int64, device=DEV)
    blk_row_off[1:] = torch.cumsum(blk_rows, 0)
    # feedback: unique(rated) + extras up to ~fb_target, keyed by block position
    n_extra = torch.clamp(fb_target - cnt, min=0) * present
    extra_pos = torch.repeat_interleave(torch.arange(nu_, device=DEV), n_extra)
    extra_item = torch.randint(0, ni_, (extra_pos.numel(),), generator=gen, device=DEV)
    pair = torch.cat([pos * ni_ + it, extra_pos * ni_ + extra_item])
    pair = torch.unique(pair)  # sorted by (pos, item)
    fpos, fitem = pair // ni_, pair % ni_
    fcnt = torch.bincount(fpos, minlength=nu_)[present]
    blk_fb_off = torch.zeros(fcnt.numel() + 1, dtype=torch.int64, device=DEV)
    blk_fb_off[1:] = torch.cumsum(fcnt, 0)
    fval = (1.0 / torch.sqrt(fcnt.float()))[torch.repeat_interleave(torch.arange(fcnt.numel(), device=DEV), fcnt)]
    ones = torch.ones(n, device=DEV)
    csr = fixed_csr(lab, [u, it], [ones, ones], 0, 1, 1)
    nb = blk_rows.numel()
    g = api.SvdGpu(nu_, ni_, k, num_ufeedback=ni_, format_type=1)
    g.set_hparams(learning_rate=0.005, wd_user=0.004, wd_item=0.004, wd_ufeedback=0.004, base_score=3.6)
    g.set_mode(api.MODE_HOGWILD)
    init_model(g, ni_ + nu_ + ni_, k)
    ug = (blk_row_off.int().cpu().numpy(), blk_fb_off.int().cpu().numpy(), np.zeros(nb, np.int32),
          fitem.int().cpu().numpy().view(np.uint32), fval.cpu().numpy())
    b = g.batch_create(host(csr), ugroup=ug)
    F, R = float(fcnt.float().mean()), float(blk_rows.float().mean())
    bytes_row = 1072 + (2 * F * 4 * k + 16 * F) / R
    nstep = 4
    units = [int(round(i * nb / nstep)) for i in range(nstep + 1)]
    for s in range(2):
        g.batch_update(b, units[s], units[s + 1])
    g.sync()
    g.timer_start()
    g.batch_update(b, 0, nb)
    ms = g.timer_stop()
    g.sync()
    emit(exp="config", config="c3 SVD++ 480kx18k k=64, user blocks", rows=n, ms=ms, ginst_s=n / ms / 1e6,
         algorithmic_gbs=n * bytes_row / ms / 1e6, bytes_per_row=bytes_row, users=nb, avg_rows_per_user=R,
         avg_feedback=F, max_rows_per_user=int(blk_rows.max()))
    b.close()
    g.close()


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a in ("c2", "c3", "c4", "c4u", "c5")]
    scale = 1.0
    if "--scale" in sys.argv:
        scale = float(sys.argv[sys.argv.index("--scale") + 1])
    for name in args or ["c2", "c4", "c5", "c3"]:
        t0 = time.perf_counter()
        {"c2": c2, "c3": c3, "c4": c4, "c4u": c4u, "c5": c5}[name](scale)
        print("# %s done in %.1fs" % (name, time.perf_counter() - t0), file=sys.stderr, flush=True)
        torch.cuda.empty_cache()
