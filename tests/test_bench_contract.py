"""CPU-only: bench.py's reference arm prints the contract's JSON line; our arm refuses to run
without a CUDA device (no CPU fallback behind the headline number)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                          env=e, timeout=600)


def test_reference_arm_line():
    p = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-rows-per-step", "50000")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sgd_training_instances_per_sec" and d["unit"] == "instances/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "configs[1]" in d["config"]["workload"] and "model" not in d["config"]
    # the same config dictionary as our arm prints (the bounded sample is described in cpu_baseline.sample)
    sys.path.insert(0, ROOT)
    import bench

    assert d["config"] == bench.workload_config(bench.TOTAL_ROWS)


def test_reference_arm_other_ranks_stay_silent():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_our_arm_needs_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = _run("--steps", "1", "--warmup", "1")
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)
    assert not any(l.startswith("{") for l in p.stdout.splitlines())
