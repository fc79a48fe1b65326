"""GPU: the C++ GPU trainer against the committed reference vectors (no oracle involved)."""
import hashlib
import json
import os

import numpy as np
import pytest

from test_oracle_golden import BASICMF_CONF, G, HOTPATH, _cli_example_run, load_case, train

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", HOTPATH)
def test_gpu_trainer_reproduces_reference_vectors(native, name, tmp_path):
    fmt, act, params, data, kind, model, pred = load_case(name)
    g = native.GpuTrainer(fmt, act, 0, dict(params, **{"gpu:mode": "exact"}))
    p = train(g, data, kind)
    blob = g.model_bytes(tmp_path)
    if name in ("sigmoid", "pairwise"):  # expf: 1 ulp
        a = np.frombuffer(blob, np.uint8)
        assert len(blob) == len(model) and blob[:1060] == model[:1060]
        fa, fb = np.frombuffer(blob[1060:], np.float32), np.frombuffer(model[1060:], np.float32)
        ok = np.isfinite(fb) & (np.abs(fb) < 1e3)  # skip the int32 tensor headers reinterpreted as floats
        assert np.abs(fa[ok] - fb[ok]).max() <= 2e-6
        assert np.abs(p - pred).max() <= 2e-6
    else:
        assert blob == model
        assert np.array_equal(p, pred)


def test_gpu_trainer_reproduces_reference_cli_run(native, tmp_path):
    """demo/basicMF/run.sh on the GPU: same predictions and the same 683588-byte model file
    as the reference CLI (0040.model)."""
    g = native.GpuTrainer(0, 0, 0, dict(BASICMF_CONF, **{"gpu:mode": "exact"}))
    p = _cli_example_run(g)
    ref = np.loadtxt(os.path.join(G, "ref_cli_pred.txt"))
    assert ["%f" % x for x in p] == ["%f" % x for x in ref]
    meta = json.load(open(os.path.join(G, "ref_cli_0040.model.json")))
    blob = g.model_bytes(tmp_path)
    assert len(blob) == meta["size"] and hashlib.sha256(blob).hexdigest() == meta["sha256"]
