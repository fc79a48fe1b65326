"""CPU-only: the oracle and the host-side format code against committed golden fixtures
(tests/golden/, generated from the unmodified reference by tests/golden/make_golden.py)
and against the reference's own shipped golden files."""
import hashlib
import json
import os

import numpy as np
import pytest

from _oracle import COracle, parse_model
from svdfeature_b200 import buffer_io

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HOTPATH = sorted(f[len("hotpath_"):-4] for f in os.listdir(G) if f.startswith("hotpath_"))


def load_case(name):
    z = np.load(os.path.join(G, "hotpath_%s.npz" % name))
    params = {k: v for k, v in z["params"]}
    n = len([k for k in z.files if k.startswith("d")])
    data = tuple(z["d%d" % i] for i in range(n))
    return int(z["fmt"]), int(z["act"]), params, data, str(z["kind"]), z["model"].tobytes(), z["pred"]


def train(t, data, kind, rounds=2):
    t.init(10)
    for r in range(rounds):
        t.set_round(r)
        (t.update_csr if kind == "csr" else t.update_ugroup)(data)
    return (t.predict_csr if kind == "csr" else t.predict_ugroup)(data)


@pytest.mark.parametrize("name", HOTPATH)
def test_oracle_reproduces_reference_vectors(name, tmp_path):
    fmt, act, params, data, kind, model, pred = load_case(name)
    o = COracle(fmt, act, 0, params)
    p = train(o, data, kind)
    assert o.model_bytes(tmp_path) == model
    assert np.array_equal(p, pred)


def test_feature_buffer_format_matches_reference_golden(tmp_path):
    """demo/basicMF/ua.base.buffer (152 B) is the reference's byte-exact fixture of BINARY_BUFFER."""
    golden = open(os.path.join(G, "ua.base.buffer"), "rb").read()
    assert len(golden) == 152
    csr = buffer_io.parse_feature_text(os.path.join(G, "ua.base.example.txt"))
    out = str(tmp_path / "b.buffer")
    buffer_io.write_feature_buffer(out, csr)
    assert open(out, "rb").read() == golden
    back, hdr = buffer_io.read_feature_buffer(os.path.join(G, "ua.base.buffer"))
    assert hdr == dict(num_batch=1, batch_size=1000, max_batch_num=8)
    for a, b in zip(back, csr):
        assert np.array_equal(a, b)
    assert back[2].tolist() == [1, 282, 2, 270, 4, 221, 5, 258] and back[1].tolist() == [5, 3, 4, 1]


def test_neighborhood_buffer_with_float_globals(tmp_path):
    golden = open(os.path.join(G, "neighborhood.buffer"), "rb").read()
    csr = buffer_io.parse_feature_text(os.path.join(G, "neighborhood.example.txt"))
    out = str(tmp_path / "n.buffer")
    buffer_io.write_feature_buffer(out, csr)
    assert open(out, "rb").read() == golden
    assert (np.diff(csr[0])[0::3] > 0).any()  # rows do carry global features


def test_ugroup_buffer_round_trip(tmp_path):
    path = os.path.join(G, "implicit.buffer.svdpp")
    ug, hdr = buffer_io.read_ugroup_buffer(path)
    assert hdr["num_batch"] == len(ug[2]) == 3 and len(ug[6]) == 9  # 3 users / 9 rows
    out = str(tmp_path / "u.buffer")
    buffer_io.write_ugroup_buffer(out, ug)
    assert open(out, "rb").read() == open(path, "rb").read()


def test_batched_buffer_round_trip(tmp_path):
    from svdfeature_b200 import synth

    csr = synth.random_general(2500, 50, 40, 10, seed=3)
    out = str(tmp_path / "r.buffer")
    buffer_io.write_feature_buffer(out, csr, batch_size=1000)
    back, hdr = buffer_io.read_feature_buffer(out)
    assert hdr["num_batch"] == 3
    for a, b in zip(back, csr):
        assert np.array_equal(a, b)


def _cli_example_run(trainer):
    """What demo/basicMF/run.sh does: 40 rounds over the 4-row example, then predict the test rows."""
    train_rows = buffer_io.parse_feature_text(os.path.join(G, "ua.base.example.txt"))
    test_rows = buffer_io.parse_feature_text(os.path.join(G, "ua.test.example.txt"))
    trainer.init(10)  # svd_feature.cpp:293 seeds with 10
    for r in range(40):
        trainer.set_round(r)
        trainer.update_csr(train_rows)
    return trainer.predict_csr(test_rows)


BASICMF_CONF = dict(base_score=3, learning_rate=0.005, wd_item=0.004, wd_user=0.004, num_item=1682, num_user=943,
                    num_global=0, num_factor=64)


def test_oracle_reproduces_reference_cli_run(tmp_path):
    o = COracle(0, 0, 0, BASICMF_CONF)
    p = _cli_example_run(o)
    ref = np.loadtxt(os.path.join(G, "ref_cli_pred.txt"), dtype=np.float64)
    assert ["%f" % x for x in p] == ["%f" % x for x in ref]
    # the reference's shipped eg.pred.txt came from another libc's rand(): tolerance fixture
    shipped = np.loadtxt(os.path.join(G, "eg.pred.txt"))
    assert np.abs(p - shipped).max() < 5e-3
    meta = json.load(open(os.path.join(G, "ref_cli_0040.model.json")))
    blob = o.model_bytes(tmp_path)
    assert len(blob) == meta["size"] == 683588
    assert hashlib.sha256(blob).hexdigest() == meta["sha256"]
    m = parse_model(blob)
    assert m["W_user"].shape == (943, 64) and m["W_item"].shape == (1682, 64)


def test_oracle_reproduces_the_ml100k_convergence_curve(tmp_path):
    """demo/basicMF on MovieLens-100K, 40 rounds, k = 64.  The plain-C oracle
    lands on the compiled reference's test RMSE after every recorded round (0.9327 after round 40) and
    on its model file, bit for bit."""
    import _ml100k

    train, test, truth, gold = _ml100k.load()
    want = {int(k): v for k, v in gold["test_rmse_after_round"].items()}
    curve, sha = _ml100k.run(COracle(0, 0, 0, gold["params"]), train, test, truth, set(want), gold["seed"], tmp_path)
    assert abs(want[40] - 0.9327) < 1e-3 and want[0] > 1.2  # the band SURVEY quotes
    for r, v in want.items():
        assert abs(curve[r] - v) < 1e-9, (r, curve[r], v)
    assert sha == gold["model_sha256_after_round_40"]


def test_planted_signal_is_learnable_by_the_sequential_loop():
    """synth.planted_mf: the oracle's held-out RMSE falls epoch by epoch towards the 0.5 noise floor (on the
    BASELINE noise labels it cannot fall at all), so convergence tests built on it can discriminate."""
    from svdfeature_b200 import synth

    nu, ni = 2000, 300
    params = dict(num_user=nu, num_item=ni, num_factor=16, learning_rate=0.01, wd_user=0.004, wd_item=0.004, base_score=3.6)
    train = synth.planted_mf(200000, nu, ni, seed=1)
    held = synth.planted_mf(20000, nu, ni, seed=2)
    assert abs(float(held[1].std()) - 1.08) < 0.08
    o = COracle(0, 0, 0, params)
    o.init(10)
    curve = []
    for _ in range(4):
        o.update_csr(train)
        curve.append(float(np.sqrt(np.mean((o.predict_csr(held) - held[1]) ** 2))))
    assert curve[0] < 0.95 and curve[3] < curve[2] < curve[1] < curve[0] and curve[3] < 0.72, curve
