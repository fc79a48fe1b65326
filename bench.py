#!/usr/bin/env python
"""bench.py -- SGD training instances/sec of the SVDFeature hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): Netflix-shaped basicMF, 480k users x 18k items,
k=64, 100M synthetic ratings; one STEP = one pass of the hot path over one batch of
--rows-per-step ratings (default: all 100M, i.e. one step = one epoch).  Prints ONE JSON line.

  The headline is the ORDERED mode (--mode exact, the default): the execution mode whose
  predictions equal the reference's sequential loop (north star: <= 1e-4 RMSE; measured 0.0,
  `parity.ordered`).  Hogwild -- faster, but 1e-2 away from the sequential order -- is reported
  as the labelled secondary `hogwild` of the same line (or as the headline with --mode hogwild).

  value      whole-job instances/s with the batches resident in HBM (device-timed, CUDA
             events on the launch stream, max over ranks).  The ordered mode's plan (per-owner
             queues, tickets: a function of the batch, not of the model) is built when the batch
             is made resident; `plan` carries what it costs and the value with it rebuilt every step
  e2e        the same metric through the C ABI with HOST (pinned) buffers: H2D of every
             step's batch, the plan, and a D2H read of a probe prediction inside the timed region
  roofline   dominant kernel (ordered: k_own, the item-owner kernel; Hogwild: k_mf): algorithmic
             bytes (1072 B/instance) / measured launch time vs the measured HBM copy
             peak.  k_own is bound by the chain of dependent updates of the hottest item
             (`chain`), not by bytes; `traffic` = DRAM bytes per launch from the committed ncu
             capture (profiles/)
  cpu_baseline  the UNMODIFIED reference (oracle/_ref) timed on the host, 1 thread,
             on a bounded prefix of the same workload
  parity     (N=1) the first --parity-rows of that prefix trained and predicted through the
             ISVDTrainer seam on the GPU in both modes: RMSE / max-abs of the predictions
             against the reference's own (ordered mode: 0.0; Hogwild: its measured distance)

N>1 (torchrun): users are hash-partitioned (user id mod N), so user rows are private to a rank;
item rows / biases are replicated and their deltas are all-reduced over NCCL inside the library
(svdgpu_allreduce_items).  The headline is WEAK scaling (every rank trains its own 480k users /
rows-per-step ratings per step; `scaling`: "weak"); the same line carries `strong_scaling`: the
stated config (480k x 18k, 100M ratings per step) split over the ranks, a step's shard trained in
--exchanges-per-step pieces with one exchange after each (--scaling strong makes it the headline).
The `convergence` object (N>1) compares held-out RMSE on a planted-signal stream with the
single-GPU ordered run.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_USER, NUM_ITEM, K = 480000, 18000, 64
TOTAL_ROWS = 100_000_000
HP = dict(learning_rate=0.005, wd_user=0.004, wd_item=0.004, base_score=3.6)
BYTES_PER_INSTANCE = 8 * K * 2 + 8 * 2 + 16 + 8 * 2  # = 1072, SURVEY.md section 8(d)
# dram__bytes_read.sum + dram__bytes_write.sum of ONE k_mf launch over 100M ratings, from
# `ncu --set full` (profiles/r1_kmf_v5_100M_ncu_full_summary.txt): 7.549 GB + 8.126 GB
NCU_DRAM_BYTES_PER_INSTANCE = (7.548526e9 + 8.126198e9) / 100e6
# the same for ONE k_own launch (ordered mode); None until the capture is committed
NCU_OWN_DRAM_BYTES_PER_INSTANCE = (9.135374e9 + 8.447472e9) / 100e6
NCU_OWN_TRAFFIC_SOURCE = ("ncu --set full, profiles/r2_kown_100M_ncu_full_summary.txt (175.8 B/instance: user rows stream "
                          "through L2/HBM once per rating, item rows stay in shared memory; the queue entries add 32 B)")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------
# synthetic Netflix-shaped ratings
# ---------------------------------------------------------------------------
def marginals(seed=10):
    from svdfeature_b200 import synth

    rng = np.random.default_rng(seed)
    ucdf = synth.lognormal_cdf(NUM_USER, 1.0, rng)
    icdf = synth.zipf_cdf(NUM_ITEM, 1.0, 70.0)
    iperm = rng.permutation(NUM_ITEM)
    return ucdf, icdf, iperm


def gen_rows_torch(n, seed, device, rank=0, world=1, shard=False, planted=False):
    """(row_ptr, label, index, value) as torch tensors on `device`.
    shard=False (weak scaling): n rows over NUM_USER users -- with world > 1 the global population is
    NUM_USER * world users, rank r owns the users congruent to r mod world and stores them under the
    local index id // world, i.e. every rank trains the same per-GPU problem as the single-GPU run.
    shard=True (strong scaling): the GLOBAL stream of n rows is generated (same seed on every rank)
    and the rows of this rank's users (user id mod world == rank) are kept, global user ids.
    planted=True: labels = 3.6 + b_user + b_item + <p_user, q_item> + N(0, 0.5): biases and a rank-8 signal a model can learn, so
    that held-out RMSE says something (the BASELINE labels are noise independent of user and item)."""
    import torch

    ucdf, icdf, iperm = marginals()
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    ucdf_t = torch.from_numpy(ucdf).to(device)
    icdf_t = torch.from_numpy(icdf).to(device)
    iperm_t = torch.from_numpy(iperm.astype(np.int64)).to(device)
    u = torch.searchsorted(ucdf_t, torch.rand(n, generator=g, device=device, dtype=torch.float64)).clamp_(max=NUM_USER - 1)
    it = iperm_t[torch.searchsorted(icdf_t, torch.rand(n, generator=g, device=device, dtype=torch.float64)).clamp_(max=NUM_ITEM - 1)]
    if planted:
        gp = torch.Generator(device=device)
        gp.manual_seed(4242)  # the planted factors do not depend on the stream's seed
        P = torch.randn(NUM_USER, 8, generator=gp, device=device) * 0.4
        Q = torch.randn(NUM_ITEM, 8, generator=gp, device=device) * 0.4
        bu = torch.randn(NUM_USER, generator=gp, device=device) * 0.6
        bi = torch.randn(NUM_ITEM, generator=gp, device=device) * 0.6
        lab = (3.6 + bu[u] + bi[it] + (P[u] * Q[it]).sum(1) + 0.5 * torch.randn(n, generator=g, device=device)).float()
    else:
        lab = torch.clamp(torch.round(3.6 + 1.1 * torch.randn(n, generator=g, device=device)), 1, 5).float()
    if shard and world > 1:
        keep = (u % world) == rank
        u, it, lab = u[keep], it[keep], lab[keep]
        n = int(u.numel())
    index = torch.stack([u, it], 1).reshape(-1).to(torch.int32)  # ids < 2^31: same bits as uint32
    value = torch.ones(2 * n, device=device, dtype=torch.float32)
    base = torch.arange(n, device=device, dtype=torch.int64) * 2
    row_ptr = torch.empty(3 * n + 1, device=device, dtype=torch.int32)
    row_ptr[0:3 * n:3] = base.int()
    row_ptr[1:3 * n:3] = base.int()
    row_ptr[2:3 * n:3] = (base + 1).int()
    row_ptr[3 * n] = 2 * n
    return row_ptr, lab, index, value


def gen_rows_numpy(n, seed):
    from svdfeature_b200 import synth

    return synth.basic_mf(n, NUM_USER, NUM_ITEM, seed=seed, zipf_s=1.0, zipf_q=70.0, user_sigma=1.0)


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ---------------------------------------------------------------------------
# reference arm / cpu baseline (the ONLY place bench.py touches oracle/)
# ---------------------------------------------------------------------------
def ref_trainer():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _oracle import COracle, RefTrainer, have_ref

    params = dict(num_user=NUM_USER, num_item=NUM_ITEM, num_factor=K, **HP)
    if have_ref():
        t, kind = RefTrainer(0, 0, 0, params), "reference"
    else:
        t, kind = COracle(0, 0, 0, params), "port"
    t.init(10)
    return t, kind


def time_cpu(rows_total, rows_per_call, parity_rows=0):
    """instances/s of the reference's single-threaded update() loop on host cores.  With
    parity_rows > 0 the first rows of the stream are first trained and predicted on a fresh
    model (seed 10): (sample, predictions) come back as the yardstick of parity_leg()."""
    t, kind = ref_trainer()
    data = gen_rows_numpy(rows_total, seed=10)
    warm = min(rows_per_call, rows_total)
    yard = None
    if parity_rows > 0:
        warm = min(parity_rows, rows_total)
    nv = int(data[0][3 * warm])
    sample = (data[0][:3 * warm + 1], data[1][:warm], data[2][:nv], data[3][:nv])
    t.update_csr(sample)  # touches the model once before the timed loop
    if parity_rows > 0:
        yard = (sample, t.predict_csr(sample))
    t0 = time.perf_counter()
    done = 0
    for r0 in range(0, rows_total, rows_per_call):
        r1 = min(rows_total, r0 + rows_per_call)
        t.update_csr((data[0][3 * r0:3 * r1 + 1], data[1][r0:r1], data[2], data[3]))
        done += r1 - r0
    dt = time.perf_counter() - t0
    return done / dt, kind, dt, yard


def parity_leg(yard, kind, device):
    """The north star's parity statement, measured in the same run: the first rows of the CPU
    sample, trained once and predicted on a fresh model (seed 10, same order) through the
    reference's own plug-in seam (create_svd_trainer -> GpuSVDFeature : ISVDTrainer), in both
    execution modes, against what the reference's trainer predicted for them."""
    from svdfeature_b200 import api

    sample, want = yard
    n = len(sample[1])
    lab = sample[1].astype(np.float64)
    out = {"rows": n, "against": "%s ISVDTrainer, same seed and order" % kind, "north_star_tolerance_rmse": 1e-4,
           "reference_rmse_vs_labels": float(np.sqrt(np.mean((want.astype(np.float64) - lab) ** 2)))}
    for mode in ("exact", "hogwild"):
        params = dict(num_user=NUM_USER, num_item=NUM_ITEM, num_factor=K, **HP)
        params["gpu:device"] = device
        params["gpu:mode"] = mode
        t = api.GpuTrainer(0, 0, 0, params)
        t.init(10)
        t.update_csr(sample)
        got = t.predict_csr(sample)
        t.close()
        d = got.astype(np.float64) - want.astype(np.float64)
        out["ordered" if mode == "exact" else mode] = {
            "rmse_vs_reference": float(np.sqrt(np.mean(d * d))), "max_abs": float(np.max(np.abs(d))),
            "rmse_vs_labels": float(np.sqrt(np.mean((got.astype(np.float64) - lab) ** 2)))}
    # throughput of the ordered mode (the one that meets the tolerance), batch resident in HBM,
    # tickets already computed: bounded by the hand-off chain of the hottest item row
    g = api.SvdGpu(NUM_USER, NUM_ITEM, K, device=device)
    g.set_hparams(**HP)
    g.set_mode(api.MODE_EXACT)
    rng = np.random.default_rng(10)
    g.upload(np.zeros(NUM_USER + NUM_ITEM, np.float32),
             (rng.standard_normal((NUM_USER + NUM_ITEM, K)) * 0.01).astype(np.float32), np.zeros(1, np.float32))
    b = g.batch_create(sample)
    g.batch_update(b)
    g.sync()
    g.timer_start()
    for _ in range(3):
        g.batch_update(b)
    ms = g.timer_stop() / 3
    out["ordered"]["instances_per_s"] = n / (ms * 1e-3)
    out["ordered"]["hottest_item_rows"] = int(np.bincount(sample[2][1:2 * n:2]).max())
    b.close()
    g.close()
    return out


def seam_leg(device, rows):
    """SURVEY hard part 2: the reference's per-instance virtual call.  `rows` ratings are handed to the GPU
    trainer one ISVDTrainer::update(Elem) at a time (what svd_feature.cpp:231-247 does; the trainer stages
    gpu:batch rows and flushes them to svdgpu_update_csr), wall clock including the final flush."""
    from svdfeature_b200 import api

    data = gen_rows_numpy(rows, seed=11)
    out = {"rows": rows, "what": "ISVDTrainer::update(const SVDFeatureCSR::Elem&) per rating through the C++ GpuSVDFeature "
                                 "(gpu:batch = 2^20 rows staged per flush), one host thread"}
    for mode in ("exact", "hogwild"):
        params = dict(num_user=NUM_USER, num_item=NUM_ITEM, num_factor=K, **HP)
        params["gpu:device"] = device
        params["gpu:mode"] = mode
        t = api.GpuTrainer(0, 0, 0, params, bulk=False)
        t.init(10)
        warm = 1 << 20
        t.update_csr((data[0][:3 * warm + 1], data[1][:warm], data[2], data[3]))
        t.finish_round()
        t0 = time.perf_counter()
        t.update_csr(data)
        t.finish_round()
        t.predict_csr((data[0][:4], data[1][:1], data[2], data[3]))  # (synchronises)
        out["ordered" if mode == "exact" else mode] = rows / (time.perf_counter() - t0)
        t.close()
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = args.ref_rows_per_step
    t, kind = ref_trainer()
    nchunk = 4
    data = gen_rows_numpy(rows * nchunk, seed=10)

    def step(s):
        c = s % nchunk
        r0, r1 = c * rows, (c + 1) * rows
        t.update_csr((data[0][3 * r0:3 * r1 + 1], data[1][r0:r1], data[2], data[3]))

    for s in range(args.warmup):
        step(s)
    t0 = time.perf_counter()
    for s in range(args.steps):
        step(args.warmup + s)
    dt = time.perf_counter() - t0
    v = rows * args.steps / dt
    line = {
        "impl": "reference", "metric": "sgd_training_instances_per_sec", "value": v, "unit": "instances/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.rows_per_step), "mode": "sequential (the reference's own loop)",
        "cpu_baseline": {"value": v, "unit": "instances/s", "cores": 1, "kind": kind,
                         "sample": "%d steps x %d ratings of the 480kx18k k=64 stream, ISVDTrainer::update loop, 1 thread of %d host cores"
                                   % (args.steps, rows, os.cpu_count())},
        "e2e": {"value": v, "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(rows_per_step):
    """The same dictionary for both arms (the reference arm times a bounded sample of it, described
    in its cpu_baseline.sample)."""
    return {"workload": "configs[1] Netflix-shaped basicMF: 480k users x 18k items, k=64, 100M synthetic ratings "
                        "(users lognormal, items Zipf-Mandelbrot s=1 q=70, shuffled)",
            "rows_per_step": rows_per_step, "l2_policy": "inputs larger than L2 (batch+model > 126 MB per step)",
            "hparams": HP}


def convergence_leg(args, api, torch, dist, dev, stream, rank, world, local, W0, epochs=5, heldout=1_000_000):
    """Planted signal (labels = 3.6 + b_u + b_i + <p_u, q_i> + N(0, 0.5)), `--convergence-rows` ratings split by
    user hash over the ranks, ordered mode, item side all-reduced E times per epoch: held-out RMSE of the
    completed model (svdgpu_allgather_users) beside the single-GPU ordered run of the same stream on rank 0
    (= the reference's sequential loop) and the noise floor."""
    n, rows_model = args.convergence_rows, NUM_USER + NUM_ITEM
    E = max(4, args.exchanges_per_step)
    scale = args.allreduce_scale if args.allreduce_scale > 0 else 1.0 / world

    def to_np(ts):
        return tuple(t.cpu().numpy() for t in ts)

    def trainer():
        g = api.SvdGpu(NUM_USER, NUM_ITEM, K, device=local)
        g.set_hparams(**HP)
        g.set_mode(api.MODE_EXACT)
        g.set_stream(stream.cuda_stream)
        g.upload(np.zeros(rows_model, np.float32), W0, np.zeros(1, np.float32))
        return g

    test = to_np(gen_rows_torch(heldout, seed=777, device=dev, planted=True))
    out = {"rows": n, "epochs": epochs, "exchanges_per_epoch": E, "scale": scale, "noise_floor_rmse": 0.5,
           "labels": "planted user/item biases + rank-8 signal + N(0, 0.5), label std 1.08; held-out = %d fresh ratings" % heldout}
    # sharded run
    mine = to_np(gen_rows_torch(n, seed=99, device=dev, rank=rank, world=world, shard=True, planted=True))
    m = len(mine[1])
    g = trainer()
    ids = [api.comm_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    g.comm_init(world, rank, ids[0])
    g.allreduce_items(scale)
    cuts = [(m * e) // E for e in range(E + 1)]
    bs = [g.batch_create((mine[0][3 * a:3 * b + 1] - mine[0][3 * a], mine[1][a:b], mine[2][2 * a:2 * b], mine[3][2 * a:2 * b]))
          for a, b in zip(cuts[:-1], cuts[1:])]
    curve = []
    for _ in range(epochs):
        for b in bs:
            g.batch_update(b)
            g.allreduce_items(scale)
        g.allgather_users()
        sse, cnt = g.eval_csr(test)
        curve.append(float(np.sqrt(sse / max(cnt, 1))))
    for b in bs:
        b.close()
    g.close()
    out["heldout_rmse_by_epoch_sharded"] = curve
    dist.barrier()
    # the single-GPU ordered run of the same global stream (rank 0 only)
    if rank == 0:
        full = to_np(gen_rows_torch(n, seed=99, device=dev, planted=True))
        g = trainer()
        b = g.batch_create(full)
        curve1 = []
        for _ in range(epochs):
            g.batch_update(b)
            sse, cnt = g.eval_csr(test)
            curve1.append(float(np.sqrt(sse / max(cnt, 1))))
        b.close()
        g.close()
        out["heldout_rmse_by_epoch_single_gpu_ordered"] = curve1
        out["gap_last_epoch"] = curve[-1] - curve1[-1]
    dist.barrier()
    return out


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="exact", choices=["hogwild", "exact"],
                    help="exact = ordered (bit-identical to the sequential reference; the headline), hogwild = throughput mode")
    ap.add_argument("--no-secondary", action="store_true", help="skip the Hogwild secondary of an ordered run")
    ap.add_argument("--rows-per-step", type=int, default=TOTAL_ROWS)
    ap.add_argument("--ref-rows-per-step", type=int, default=2_000_000)
    ap.add_argument("--cpu-rows", type=int, default=20_000_000, help="rows of the cpu_baseline sample")
    ap.add_argument("--parity-rows", type=int, default=2_000_000,
                    help="rows of the cpu_baseline sample that are also trained on the GPU and compared (N=1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="N=1: skip the kernel-resident runs of configs[2..4] (SVD++ blocks, pairwise rows, neighbourhood rows)")
    ap.add_argument("--seam-rows", type=int, default=10_000_000,
                    help="N=1: ratings of the per-Elem seam leg (ISVDTrainer::update one rating at a time); 0 = skip")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="name=value passed to svdgpu_set_option")
    ap.add_argument("--allreduce-every", type=int, default=1)
    ap.add_argument("--allreduce-scale", type=float, default=0.0,
                    help="factor on the summed item-side deltas; 0 = mean (1/world), the default: it tracks the single-GPU "
                         "run at any exchange frequency, the plain sum only when exchanges are frequent "
                         "(profiles/r2_convergence_sweep.jsonl)")
    ap.add_argument("--no-strong", action="store_true", help="N>1, weak headline: skip the strong-scaling leg")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = every rank trains its own 480k users / rows-per-step ratings (the global problem "
                         "grows with N); strong = the stated config (480k users, rows-per-step ratings) split over the ranks")
    ap.add_argument("--exchanges-per-step", type=int, default=4,
                    help="strong scaling: a step's shard is trained in this many resident batches, the item side "
                         "all-reduced after each")
    ap.add_argument("--convergence-rows", type=int, default=20_000_000,
                    help="N>1: rows of the planted-signal convergence leg (0 = skip it)")
    ap.add_argument("--e2e-chunk-rows", type=int, default=1 << 22, help="rows per launch of the ordered e2e call")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference_arm(args)
        return

    # stdout must carry exactly ONE JSON line: libraries (NCCL prints its version banner to
    # stdout) are sent to stderr for the duration of the run
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)

    import torch
    from svdfeature_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    import types

    def make_data(strong):
        """The rating stream of this rank, staged in pinned host memory (the C ABI takes host pointers)."""
        D = types.SimpleNamespace(strong=strong)
        rows = args.rows_per_step
        D.nchunk = max(1, min(TOTAL_ROWS // rows, args.steps + args.warmup))
        D.global_rows = rows * (1 if strong else world)  # ratings all ranks train per step
        if strong:
            # the stated problem (480k users, rows-per-step ratings per step) split by user hash: every rank
            # generates the same global stream and keeps the rows of its users; one resident chunk per rank
            ts = gen_rows_torch(rows, seed=10, device=dev, rank=rank, world=world, shard=True)
            D.rows, D.nchunk = int(ts[1].numel()), 1
        else:
            ts = gen_rows_torch(rows * D.nchunk, seed=10 + rank, device=dev, rank=rank, world=world)
            D.rows = rows
        D.total = D.rows * D.nchunk
        D.h_rp, D.h_lab, D.h_idx, D.h_val = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t) for t in ts]
        torch.cuda.synchronize()
        del ts
        torch.cuda.empty_cache()
        probe_n, check_n = 1024, min(D.rows, 1_000_000)
        D.probe = (D.h_rp[:3 * probe_n + 1].numpy(), D.h_lab[:probe_n].numpy(), D.h_idx.numpy(), D.h_val.numpy())
        D.check = (D.h_rp[:3 * check_n + 1].numpy(), D.h_lab[:check_n].numpy(), D.h_idx.numpy(), D.h_val.numpy())
        return D

    t_setup = time.perf_counter()
    main_strong = world > 1 and args.scaling == "strong"
    data_main = make_data(main_strong)
    rows, global_rows, strong = data_main.rows, data_main.global_rows, main_strong
    h_idx = data_main.h_idx

    stream = torch.cuda.Stream(device=dev)
    rng = np.random.default_rng(10)
    rows_model = NUM_USER + NUM_ITEM
    W0 = (rng.standard_normal((rows_model, K)) * 0.01).astype(np.float32)  # N(0, 0.01^2) as rand_init
    use_allreduce = world > 1

    def measure(D, mode, steps, warmup, want_e2e):
        """One trainer in `mode` ("exact" = ordered, "hogwild") on the stream D: resident `value`,
        kernel-only time, e2e through host buffers, and the state of the model afterwards."""
        rows, nchunk, total, global_rows, strong = D.rows, D.nchunk, D.total, D.global_rows, D.strong
        h_rp, h_lab, h_idx, h_val, probe, check = D.h_rp, D.h_lab, D.h_idx, D.h_val, D.probe, D.check
        g = api.SvdGpu(NUM_USER, NUM_ITEM, K, device=local)
        g.set_hparams(**HP)
        g.set_mode(api.MODE_HOGWILD if mode == "hogwild" else api.MODE_EXACT)
        for o in args.opt:
            name, v = o.split("=")
            g.set_option(name, int(v))
        g.set_stream(stream.cuda_stream)
        g.upload(np.zeros(rows_model, np.float32), W0, np.zeros(1, np.float32))
        res = {"mode": mode}

        # ---- resident batches (value) ----
        t_b = time.perf_counter()
        # a step = one chunk of `rows` ratings; with --exchanges-per-step E (strong scaling) the chunk is
        # trained in E pieces and the item side is all-reduced after each
        E = max(1, args.exchanges_per_step) if strong else 1
        cuts = [[c * rows + (rows * e) // E for e in range(E + 1)] for c in range(nchunk)]
        if mode == "exact":  # the ordered mode trains whole resident batches (its plan is per batch)
            batches = []
            for c in range(nchunk):
                for e in range(E):
                    a, b = cuts[c][e], cuts[c][e + 1]
                    sl = (h_rp[3 * a:3 * b + 1] - int(h_rp[3 * a])).contiguous().numpy()
                    batches.append(g.batch_create((sl, h_lab[a:b].numpy(), h_idx[2 * a:2 * b].numpy(),
                                                   h_val[2 * a:2 * b].numpy())))

            def piece(s, e):
                g.batch_update(batches[(s % nchunk) * E + e])
        else:
            batches = [g.batch_create((h_rp, h_lab, h_idx, h_val))]

            def piece(s, e):
                c = s % nchunk
                g.batch_update(batches[0], cuts[c][e], cuts[c][e + 1])

        def step(s):
            for e in range(E):
                piece(s, e)
                if E > 1:
                    exchange()
        g.sync()
        log("[rank %d] %s: %d rows resident in %d batch(es), %.2f s" % (rank, mode, total, len(batches), time.perf_counter() - t_b))

        def exchange():
            if use_allreduce:
                # pack the item-side deltas, ONE ncclAllReduce on the launch stream, apply: all inside
                # the library (svdgpu_allreduce_items, svdgpu_comm.cu)
                g.allreduce_items(args.allreduce_scale if args.allreduce_scale > 0 else 1.0 / world)

        with torch.cuda.stream(stream):
            if use_allreduce:
                ids = [api.comm_id() if rank == 0 else None]
                dist.broadcast_object_list(ids, src=0)
                g.comm_init(world, rank, ids[0])
                exchange()  # (first call: snapshot of the replicated slabs)
            for s in range(warmup):
                step(s)
                if E == 1 and (s + 1) % args.allreduce_every == 0:
                    exchange()
            g.sync()
            clocks = ClockSampler(local)
            if rank == 0:
                clocks.start()
                time.sleep(0.2)  # let nvidia-smi attach before the timed region
            if dist:
                dist.barrier()  # (after rank 0's sleep: every rank enters the timed region together)
            torch.cuda.synchronize()
            l0, o0 = g.counter("kernel_launches"), g.counter("own_launches")
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            c0, b0 = g.counter("collectives"), g.counter("collective_bytes")
            for s in range(steps):
                step(warmup + s)
                if E == 1 and (s + 1) % args.allreduce_every == 0:
                    exchange()
            ev1.record(stream)
            g.sync()
            torch.cuda.synchronize()
            if dist:
                dist.barrier()
            res["clocks"] = clocks.stop() if rank == 0 else None
            ms = ev0.elapsed_time(ev1)
            res["launches"] = g.counter("kernel_launches") - l0
            res["own_launches"] = g.counter("own_launches") - o0
            res["collectives"] = g.counter("collectives") - c0
            res["collective_bytes"] = g.counter("collective_bytes") - b0

            # kernel-only duration of the training launches (roofline numerator), same stream
            kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for s in range(steps):
                kev[s][0].record(stream)
                step(warmup + s)
                kev[s][1].record(stream)
            g.sync()
            res["kms"] = float(np.mean([a.elapsed_time(b) for a, b in kev]))
            if use_allreduce:  # per-rank diagnostics (stderr): the exchange alone, back to back
                xe0, xe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                dist.barrier()
                xe0.record(stream)
                for _ in range(5):
                    exchange()
                xe1.record(stream)
                g.sync()
                torch.cuda.synchronize()
                res["exchange_ms"] = xe0.elapsed_time(xe1) / 5
                log("[rank %d] %s kernel-only %.3f ms/step, exchange-only %.3f ms, timed loop %.3f ms/step"
                    % (rank, mode, res["kms"], res["exchange_ms"], ms / steps))
        if dist:
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        res["ms"] = ms
        res["value"] = global_rows * steps / (ms * 1e-3)

        # ---- the ordered mode's plan: what rebuilding it every step would cost ----
        if mode == "exact" and res["own_launches"] > 0:
            ksteps = max(1, min(steps, 5))
            g.sync()
            t0 = time.perf_counter()
            for s in range(ksteps):
                g.batch_plan(batches[s % nchunk])
            g.sync()
            plan_ms = 1e3 * (time.perf_counter() - t0) / ksteps
            res["plan"] = {"ms_per_batch": plan_ms, "what": "device: shape/bound checks, item and user histograms, two radix sorts "
                           "(by user: tickets; by owner: queues), queue entries; host: LPT of the item counts",
                           "value_with_plan_rebuilt_every_step": global_rows / ((ms / steps + plan_ms * E) * 1e-3)}

        # ---- end to end through the C ABI with host buffers ----
        if want_e2e:
            def e2e_step(s):
                c = s % nchunk
                r0, r1 = c * rows, (c + 1) * rows
                g.update_csr((h_rp[3 * r0:3 * r1 + 1], h_lab[r0:r1], h_idx, h_val))
                if use_allreduce:
                    with torch.cuda.stream(stream):
                        exchange()
                return g.predict_csr(probe)  # D2H read of the step's result (probe predictions)

            g.set_option("chunk_rows", args.e2e_chunk_rows if mode == "exact" else 1 << 20)

            def measure_e2e():
                for s in range(max(1, min(warmup, 3))):
                    e2e_step(s)
                g.sync()
                if dist:
                    dist.barrier()
                h0, d0 = g.counter("h2d_bytes"), g.counter("d2h_bytes")
                p0 = [g.counter(k) for k in ("own_deals", "own_redeals", "own_lpt_us", "own_cntwait_us")]
                t0 = time.perf_counter()
                for s in range(steps):
                    e2e_step(warmup + s)
                g.sync()
                dt = time.perf_counter() - t0
                if dist:
                    tt = torch.tensor([dt], device=dev, dtype=torch.float64)
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    dt = float(tt.item())
                out = {"value": global_rows * steps / dt, "unit": "instances/s",
                       "h2d_bytes_per_step": (g.counter("h2d_bytes") - h0) // steps,
                       "d2h_bytes_per_step": (g.counter("d2h_bytes") - d0) // steps}
                if mode == "exact":
                    p1 = [g.counter(k) for k in ("own_deals", "own_redeals", "own_lpt_us", "own_cntwait_us")]
                    out["plans_per_step"] = {"dealt_anew": (p1[0] - p0[0]) / steps, "deal_carried_over": (p1[1] - p0[1]) / steps,
                                             "host_ms_dealing": (p1[2] - p0[2]) / 1e3 / steps,
                                             "host_ms_waiting_for_item_counts": (p1[3] - p0[3]) / 1e3 / steps}
                return out

            e2e = measure_e2e()
            e2e["timing"] = "wall clock around K calls of svdgpu_update_csr (pinned host buffers) + probe predict"
            if mode == "exact":
                e2e["includes"] = "H2D of the batch, the ordered mode's plan of every chunk (device sorts + host LPT), k_own, D2H of the probe"
            compact_text = ("inside the timed region host threads verify, element by element, that a chunk's row_ptr is "
                            "the progression of constant feature counts and that its values are all 1.0f; such arrays "
                            "are rebuilt on the device instead of copied (12 of the reference layout's 32 bytes per "
                            "instance cross PCIe)")
            if mode == "exact" and int(os.environ.get("LOCAL_WORLD_SIZE", "1")) < 4:
                # the ordered mode copies everything unless four or more ranks share the host (option compact_h2d = 2 asks for the compact path)
                e2e["h2d"] = "all 32 bytes per instance are copied; compact_h2d = the same call with option compact_h2d=2: " + compact_text
                side, side_opt = "compact_h2d", 2
            else:
                e2e["compact_h2d"] = compact_text + "; full_copy = the same call with the option off"
                side, side_opt = "full_copy", 0
            try:
                g.set_option("compact_h2d", side_opt)
                leg = measure_e2e()
                e2e[side] = {"value": leg["value"], "h2d_bytes_per_step": leg["h2d_bytes_per_step"]}
            except api.SvdGpuError as e:  # (a side measurement must not cost the run its line)
                e2e[side] = {"error": str(e)}
            finally:
                g.set_option("compact_h2d", 1)
            res["e2e"] = e2e

        # ---- the model after all those epochs: finite, and still fitting the labels ----
        try:
            sse, cnt = g.eval_csr(check)
            ub, Wm, _ = g.download()
            res["model_check"] = {"finite": bool(np.isfinite(Wm).all() and np.isfinite(ub).all()),
                                  "rmse_vs_labels": float(np.sqrt(sse / max(cnt, 1))), "rows": int(cnt),
                                  "note": "first rows of the training stream, after every timed and untimed epoch of this run"}
        except Exception as e:
            res["model_check"] = {"error": "%s: %s" % (type(e).__name__, e)}
        for b_ in batches:
            b_.close()
        g.close()
        return res

    main_res = measure(data_main, args.mode, args.steps, args.warmup, not args.no_e2e)
    second = None
    if args.mode == "exact" and not args.no_secondary:
        try:
            second = measure(data_main, "hogwild", max(1, min(args.steps, 10)), 3, not args.no_e2e and world == 1)
        except Exception as e:  # the secondary must not cost the run its line
            second = {"error": "%s: %s" % (type(e).__name__, e)}
    # N > 1, weak headline: the stated config (fixed total problem) split over the ranks, same run
    other = None
    if world > 1 and not main_strong and not args.no_strong:
        try:
            del data_main.h_rp, data_main.h_lab, data_main.h_val
            other = measure(make_data(True), args.mode, max(1, min(args.steps, 5)), 3, False)
        except Exception as e:
            other = {"error": "%s: %s" % (type(e).__name__, e)}
    ms, kms, value, launches, clk, e2e = (main_res["ms"], main_res["kms"], main_res["value"], main_res["launches"],
                                          main_res["clocks"], main_res.get("e2e"))

    # ---- N > 1: does the sharded run still learn what the single-GPU ordered run learns? ----
    conv = None
    if world > 1 and args.convergence_rows > 0:
        try:
            conv = convergence_leg(args, api, torch, dist, dev, stream, rank, world, local, W0)
        except Exception as e:  # a diagnostic leg must not cost the run its line
            conv = {"error": "%s: %s" % (type(e).__name__, e)}

    if dist:
        dist.barrier()
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    def roof(kms_, mode):
        achieved = rows * BYTES_PER_INSTANCE / (kms_ * 1e-3) / 1e9
        r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
             "kernel": "k_mf" if mode == "hogwild" else "k_own",
             "algorithmic_bytes_per_instance": BYTES_PER_INSTANCE, "launch_ms": kms_, "peak_source": peak_src,
             "frac_of_nominal_8000": achieved / 8000.0}
        if mode == "hogwild":
            r["traffic"] = rows * NCU_DRAM_BYTES_PER_INSTANCE
            r["traffic_source"] = ("ncu --set full, profiles/r1_kmf_v5_100M_ncu_full_summary.txt (156.7 B/instance: the "
                                   "128 MB model is L2-resident, so DRAM moves less than the algorithmic bytes)")
        else:
            hot = int(np.bincount(h_idx[1:2 * rows:2].numpy()).max())
            r["traffic"] = rows * NCU_OWN_DRAM_BYTES_PER_INSTANCE if NCU_OWN_DRAM_BYTES_PER_INSTANCE else None
            r["traffic_source"] = NCU_OWN_TRAFFIC_SOURCE
            r["chain"] = {"hottest_item_rows": hot, "us_per_link": 1e3 * kms_ / hot,
                          "note": "the ordered mode is bound by latency, not bytes: the updates of one item are a chain of "
                                  "dependent dot products (the reference's sequential semantics), so a launch cannot finish "
                                  "before (rows of the hottest item) x (time of one link); every other row overlaps with it"}
        return r

    roofline = roof(kms, args.mode)
    # the L2 read rate (a 48 MB buffer streamed 40 times): the second roofline of the Hogwild gather kernel, whose
    # 128 MB model is almost L2-resident (DRAM moves 157 of the 1072 algorithmic bytes per instance)
    l2_peak = None
    try:
        gp = api.SvdGpu(16, 16, 4, device=local)
        l2_peak = {"read_gbs": gp.microbench(0, 48 << 20, 40), "copy_gbs": gp.microbench(1, 24 << 20, 40),
                   "hbm_read_gbs": gp.microbench(0, 4 << 30, 2),
                   "how": "svdgpu_microbench: 16-byte ld.global.cg over 48 MB x 40 passes (L2-resident), copy of 24 MB "
                          "(read+write), and the same read over 4 GB (HBM)"}
        gp.close()
    except Exception as e:
        l2_peak = {"error": "%s: %s" % (type(e).__name__, e)}
    cpu = parity = None
    if not args.no_cpu_baseline:
        v, kind, dt, yard = time_cpu(args.cpu_rows, 1_000_000, 0 if world > 1 else args.parity_rows)
        if yard is not None:
            try:
                parity = parity_leg(yard, kind, local)
            except Exception as e:  # a diagnostic leg must not cost the run its line
                parity = {"error": "%s: %s" % (type(e).__name__, e)}
        cpu = {"value": v, "unit": "instances/s", "cores": 1, "kind": kind,
               "sample": "first %d ratings of the same stream, ISVDTrainer::update loop, %.1f s, 1 thread of %d host cores"
                         % (args.cpu_rows, dt, os.cpu_count())}

    seam = None
    if world == 1 and args.seam_rows > 0:
        try:
            seam = seam_leg(local, args.seam_rows)
        except Exception as e:  # a diagnostic leg must not cost the run its line
            seam = {"error": "%s: %s" % (type(e).__name__, e)}
    # the other single-GPU BASELINE configs, kernel-resident, scaled down (tools/bench_configs.py)
    others = None
    if world == 1 and not args.no_other_configs:
        try:
            p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_configs.py"), "c3", "c4", "c5", "--scale", "0.25"],
                               capture_output=True, text=True, timeout=240)
            others = []
            for l in p.stdout.splitlines():
                if l.startswith("{"):
                    d = json.loads(l)
                    others.append({"config": d.get("config"), "mode": "hogwild", "rows": d.get("rows", d.get("pairs")),
                                   "value": 1e9 * d["ginst_s"], "unit": "instances/s",
                                   "algorithmic_gbs": d.get("algorithmic_gbs"), "bytes_per_row": d.get("bytes_per_row"),
                                   "frac_of_hbm_peak": (d.get("algorithmic_gbs") or 0.0) / peak})
            if p.returncode != 0 and not others:
                others = {"error": p.stderr[-300:]}
        except Exception as e:  # a side measurement must not cost the run its line
            others = {"error": "%s: %s" % (type(e).__name__, e)}
    cfg = workload_config(rows)
    line = {
        "metric": "sgd_training_instances_per_sec", "value": value, "unit": "instances/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg, "mode": ("ordered: bit-identical to the reference's sequential loop (k_own, item-owner warps)"
                                if args.mode == "exact" else "hogwild: no ordering between the instances of a launch"),
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu, "model_check": main_res.get("model_check"),
    }
    if "plan" in main_res:
        line["plan"] = main_res["plan"]
    if seam is not None:
        line["seam"] = seam
    if others is not None:
        line["other_configs"] = others
    if second is not None:
        if "error" in second:
            line["hogwild"] = second
        else:
            line["hogwild"] = {"what": "secondary: the throughput mode on the same workload; NOT within the north star's 1e-4 "
                                       "of the sequential order (see parity.hogwild)",
                               "value": second["value"], "ms_per_step": second["ms"] / max(1, min(args.steps, 10)),
                               "e2e": second.get("e2e"), "roofline": dict(roof(second["kms"], "hogwild"), **(
                                   {"l2_peak": l2_peak, "l2_frac": roof(second["kms"], "hogwild")["achieved"] / l2_peak["read_gbs"]}
                                   if l2_peak and "read_gbs" in l2_peak else {"l2_peak": l2_peak})),
                               "gpu_launches": int(second["launches"]), "model_check": second.get("model_check")}
    if parity:
        line["parity"] = parity
    if world > 1:
        if strong:
            line["config"]["parallelism"] = (
                "strong scaling: the stated problem (%d users x %d items, %d ratings per step) split by user id mod %d; "
                "a step's shard is trained in %d piece(s), the item side all-reduced (NCCL, inside the library) after each"
                % (NUM_USER, NUM_ITEM, global_rows, world, max(1, args.exchanges_per_step)))
        else:
            line["config"]["parallelism"] = (
                "weak scaling: user-hash shards x%d (global problem %d users x %d items, %d ratings per step; each rank "
                "owns %d users), item-side delta allreduce (NCCL, inside the library) every %d step(s)"
                % (world, NUM_USER * world, NUM_ITEM, global_rows, NUM_USER, args.allreduce_every))
        ncoll = max(1, main_res.get("collectives", 0))
        line["exchange"] = {"collectives_per_step": main_res.get("collectives", 0) / args.steps,
                            "message_bytes": main_res.get("collective_bytes", 0) // ncoll,
                            "ms_per_exchange": main_res.get("exchange_ms"),
                            "scale": args.allreduce_scale if args.allreduce_scale > 0 else 1.0 / world,
                            "what": "delta = current - snapshot of W_item, i_bias, g_bias packed into one buffer, ONE ncclAllReduce(sum) "
                                    "on the launch stream, current = snapshot + scale * sum (svdgpu_allreduce_items)"}
        if conv is not None:
            line["convergence"] = conv
        if other is not None:
            if "error" in other:
                line["strong_scaling"] = other
            else:
                line["strong_scaling"] = {
                    "what": "the stated config (480k users x 18k items, %d ratings per step) split by user id mod %d, a step's "
                            "shard trained in %d pieces with the item side all-reduced after each; bounded by the chain of "
                            "the heaviest USER (its ratings stay on one rank and hand the user row from owner to owner)"
                            % (args.rows_per_step, world, max(1, args.exchanges_per_step)),
                    "value": other["value"], "ms_per_step": other["ms"] / max(1, min(args.steps, 5)),
                    "collectives_per_step": other.get("collectives", 0) / max(1, min(args.steps, 5)),
                    "model_check": other.get("model_check")}
    emit(line)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
