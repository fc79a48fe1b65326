"""CPU-only: bench.py's own arm, end to end, against stand-ins for the CUDA pieces (torch.cuda,
the ctypes binding).  Catches control-flow and naming errors in the harness itself; says nothing
about the kernels (the -m gpu suite and the real run do)."""
import contextlib
import json
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Event:
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self, stream=None):
        import time

        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max(1e-3, (other.t - self.t) * 1e3)


class _Stream:
    cuda_stream = 0

    def __init__(self, device=None):
        pass


class _Batch:
    def close(self):
        pass


class _FakeGpu:
    """What bench.py calls on api.SvdGpu; predictions are the base score."""

    def __init__(self, *a, **k):
        self.c = {"kernel_launches": 0, "h2d_bytes": 0, "d2h_bytes": 0, "own_launches": 0, "collectives": 0, "collective_bytes": 0,
                  "own_deals": 0, "own_redeals": 0, "own_lpt_us": 0, "own_cntwait_us": 0}
        self.compact = 1
        self.mode = 1

    def set_hparams(self, **kw):
        pass

    def set_mode(self, m):
        self.mode = m

    def set_option(self, name, v):
        if name == "compact_h2d":
            self.compact = v

    def set_stream(self, s):
        pass

    def upload(self, *a):
        pass

    def batch_create(self, csr):
        return _Batch()

    def batch_update(self, b, begin=0, end=None):
        self.c["kernel_launches"] += 3
        if self.mode == 0:  # ordered: the item-owner kernel
            self.c["own_launches"] += 1

    def batch_plan(self, b):
        return True

    def eval_csr(self, csr, scale=1.0):
        n = len(csr[1])
        return 1.21 * n, n

    def microbench(self, which, nbytes, iters):
        return 20000.0

    def download(self):
        return np.zeros(4, np.float32), np.zeros((4, 2), np.float32), np.zeros(1, np.float32)

    def update_csr(self, csr):
        n = len(csr[1])
        compact = self.compact >= (2 if self.mode == 0 else 1)  # (the ordered mode copies everything unless asked)
        self.c["h2d_bytes"] += n * (12 if compact else 32)

    def predict_csr(self, csr):
        n = len(csr[1])
        self.c["d2h_bytes"] += 4 * n
        return np.full(n, 3.6, np.float32)

    def sync(self):
        pass

    def counter(self, name):
        return self.c[name]

    def timer_start(self):
        pass

    def timer_stop(self):
        return 1.0

    def close(self):
        pass


class _FakeTrainer:
    def __init__(self, *a, **k):
        pass

    def init(self, seed):
        pass

    def update_csr(self, csr):
        pass

    def predict_csr(self, csr):
        return np.full(len(csr[1]), 3.6, np.float32)

    def finish_round(self):
        pass

    def close(self):
        pass


def test_bench_main_runs_against_stand_ins(monkeypatch, capfd):
    import torch

    sys.path.insert(0, ROOT)
    import bench
    from svdfeature_b200 import api

    cpu = torch.device("cpu")
    real_empty, real_device, real_gen = torch.empty, torch.device, torch.Generator
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(torch.cuda, "Stream", _Stream)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch, "device", lambda *a: cpu)
    monkeypatch.setattr(torch, "Generator", lambda device=None: real_gen())
    monkeypatch.setattr(torch, "empty", lambda *a, pin_memory=False, **k: real_empty(*a, **k))
    monkeypatch.setattr(api, "SvdGpu", _FakeGpu)
    monkeypatch.setattr(api, "GpuTrainer", _FakeTrainer)
    monkeypatch.setattr(bench, "TOTAL_ROWS", 40000)
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "2", "--warmup", "1", "--rows-per-step", "20000",
                                      "--cpu-rows", "30000", "--parity-rows", "5000", "--seam-rows", "3000",
                                      "--no-other-configs"])
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    capfd.readouterr()
    bench.main()
    sys.stdout.flush()
    lines = [l for l in capfd.readouterr().out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline",
                "parity"):
        assert key in d, key
    assert d["steps"] == 2 and d["warmup"] == 1 and d["gpu_launches"] == 6
    # the headline is the ordered mode; Hogwild is the labelled secondary of the same line
    assert d["mode"].startswith("ordered") and d["roofline"]["kernel"] == "k_own" and "chain" in d["roofline"]
    assert d["plan"]["value_with_plan_rebuilt_every_step"] > 0
    assert d["model_check"]["finite"] is True and abs(d["model_check"]["rmse_vs_labels"] - 1.1) < 1e-6
    assert d["hogwild"]["value"] > 0 and d["hogwild"]["roofline"]["kernel"] == "k_mf"
    assert "mode" not in d["config"] and d["config"]["rows_per_step"] == 20000
    assert d["seam"]["ordered"] > 0 and d["seam"]["hogwild"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 20000 * 32 and d["e2e"]["compact_h2d"]["h2d_bytes_per_step"] == 20000 * 12
    assert d["hogwild"]["e2e"]["h2d_bytes_per_step"] == 20000 * 12 and d["hogwild"]["e2e"]["full_copy"]["h2d_bytes_per_step"] == 20000 * 32
    assert d["e2e"]["plans_per_step"]["dealt_anew"] == 0
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1
    assert "error" not in d["parity"]
    assert d["parity"]["ordered"]["instances_per_s"] > 0 and "rmse_vs_reference" in d["parity"]["hogwild"]
