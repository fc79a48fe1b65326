"""Pin the C restatement (oracle/liboracle.so) bit-for-bit against the UNMODIFIED
reference compiled into oracle/_ref (built from /root/reference by oracle/Makefile).
CPU only."""
import numpy as np
import pytest

import _cases
from _oracle import COracle, COracleRanker, RefRanker, RefTrainer, have_ref, build_oracle
from svdfeature_b200 import synth

build_oracle()
CASES = _cases.cases()


def _train(cls, fmt, act, params, data, kind, tmp, rounds=2):
    t = cls(fmt, act, 0, params)
    t.init(10)
    for r in range(rounds):
        t.set_round(r)
        (t.update_csr if kind == "csr" else t.update_ugroup)(data)
    pred = (t.predict_csr if kind == "csr" else t.predict_ugroup)(data)
    return t.model_bytes(tmp), pred


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_compiled_reference(name, tmp_path):
    fmt, act, params, data, kind = CASES[name]
    mo, po = _train(COracle, fmt, act, params, data, kind, tmp_path)
    mr, pr = _train(RefTrainer, fmt, act, params, data, kind, tmp_path)
    assert mo == mr, "model bytes differ"
    assert np.array_equal(po, pr), "predictions differ"


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("reg_method,reg_global", [(1, 1), (2, 0), (3, 0)])
def test_oracle_other_regularisers(reg_method, reg_global, tmp_path):
    fmt, act, params, data, kind = CASES["general_k13_dups"]
    params = dict(params, reg_method=reg_method, reg_global=reg_global)
    mo, po = _train(COracle, fmt, act, params, data, kind, tmp_path)
    mr, pr = _train(RefTrainer, fmt, act, params, data, kind, tmp_path)
    assert mo == mr and np.array_equal(po, pr)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["general_k13_dups", "general_k40", "basic_k16", "active_2", "svdpp_k16", "reg_l1"])
def test_oracle_side_features(name, tmp_path):
    """feature_user / feature_item expansion (base.h:298-308,330-349,365-379,399-422) against the
    compiled reference: same model bytes and predictions."""
    fmt, act, params, data, kind = CASES[name]
    fu, fi = str(tmp_path / "user.side"), str(tmp_path / "item.side")
    _cases.write_side_features(fu, params["num_user"], params["num_user"], seed=1)
    _cases.write_side_features(fi, params["num_item"], params["num_item"], seed=2)
    params = dict(params, feature_user=fu, feature_item=fi)
    mo, po = _train(COracle, fmt, act, params, data, kind, tmp_path)
    mr, pr = _train(RefTrainer, fmt, act, params, data, kind, tmp_path)
    assert mo == mr and np.array_equal(po, pr)
    # and the side features do change the result
    m0, p0 = _train(COracle, fmt, act, CASES[name][2], data, kind, tmp_path)
    assert not np.array_equal(p0, po)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_oracle_loads_reference_model_file(tmp_path):
    fmt, act, params, data, kind = CASES["svdpp_k16"]
    r = RefTrainer(fmt, act, 0, params)
    r.init(10)
    r.update_ugroup(data)
    path = str(tmp_path / "ref.model")
    r.save_model(path)
    o = COracle(fmt, act, 0, params)
    o.load_model(path)
    o.init_trainer = None
    o.lib.svdo_init_trainer(o.h)
    assert np.array_equal(o.predict_ugroup(data), r.predict_ugroup(data))
    assert o.model_bytes(tmp_path) == open(path, "rb").read()


RANK_CASES = _cases.rank_cases()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", sorted(RANK_CASES))
def test_oracle_ranker_matches_compiled_reference(name, tmp_path):
    """SVDFeatureRanker (base.h:597-813): the restatement returns the same item indices / rank positions
    as the compiled reference's create_svd_ranker on the same model file and tagged stream."""
    fmt, params, skw, rparams = RANK_CASES[name]
    path = _cases.rank_model(fmt, params, tmp_path)
    stream = synth.rank_stream(num_user=params["num_user"], num_item=params["num_item"], **skw)
    kind = "ug" if skw.get("ugroup") else "csr"
    ro = COracleRanker(path, skw["num_item_set"], rparams).rank(stream, kind)
    rr = RefRanker(path, skw["num_item_set"], rparams).rank(stream, kind)
    assert len(ro) > 0 and np.array_equal(ro, rr)
    if rparams.get("top_k"):
        assert len(ro) == rparams["top_k"] * skw["num_sections"]


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_oracle_ranker_side_features(tmp_path):
    fmt, params, skw, rparams = RANK_CASES["rank_pos_k20"]
    path = _cases.rank_model(fmt, params, tmp_path)
    fu, fi = str(tmp_path / "user.side"), str(tmp_path / "item.side")
    _cases.write_side_features(fu, params["num_user"], params["num_user"], seed=1)
    _cases.write_side_features(fi, params["num_item"], params["num_item"], seed=2)
    stream = synth.rank_stream(num_user=params["num_user"], num_item=params["num_item"], **skw)
    rp = dict(rparams, feature_user=fu, feature_item=fi)
    ro = COracleRanker(path, skw["num_item_set"], rp).rank(stream)
    rr = RefRanker(path, skw["num_item_set"], rp).rank(stream)
    assert np.array_equal(ro, rr)
    assert not np.array_equal(ro, COracleRanker(path, skw["num_item_set"], rparams).rank(stream))


def _random_config(seed):
    """A random point of the hot path's configuration space: factor count (any residue mod 4),
    loss, regulariser, decays (zero, tiny = the "scalar is one" shortcut, ordinary), bias switches,
    learning-rate decay, ragged rows with duplicates; user-grouped (SVD++) for odd seeds."""
    rng = np.random.default_rng(1000 + seed)
    act = int(rng.choice([0, 0, 1, 2, 3, 5, 6, 7]))
    wd = lambda: float(rng.choice([0.0, 1e-5, 0.004, 0.05]))  # noqa: E731
    params = dict(num_user=NU_R, num_item=NI_R, num_global=NG_R, num_factor=int(rng.integers(1, 41)),
                  learning_rate=float(rng.choice([0.001, 0.01, 0.05])), wd_user=wd(), wd_item=wd(),
                  wd_user_bias=wd(), wd_item_bias=wd(), wd_global=wd(),
                  base_score=0.6 if act in (1, 2, 3, 7) else float(rng.choice([0.0, 3.6])),
                  no_user_bias=int(rng.integers(0, 2)), num_regfree_global=int(rng.choice([0, 5])),
                  reg_method=int(rng.choice([0, 0, 1, 2, 3])), reg_global=int(rng.choice([0, 1])),
                  user_nonnegative=int(rng.choice([0, 0, 1])),
                  u_init_sigma=float(rng.choice([0.01, 0.1])), i_init_sigma=float(rng.choice([0.01, 0.1])))
    if rng.random() < 0.3:
        params.update(decay_learning_rate=1, decay_rate=0.8)
    ug = seed % 2 == 1
    if ug:
        params.update(num_ufeedback=NI_R, wd_ufeedback=wd(), wd_ufeedback_bias=wd(),
                      scale_lr_ufeedback=float(rng.choice([1.0, 0.3])), ufeedback_init_sigma=0.01)
        data = synth.user_grouped(600, NU_R, NI_R, avg_fb=int(rng.integers(1, 12)), seed=seed)
        if rng.random() < 0.5:
            data = _cases.split_tags(data, every=2)
        if act in (1, 2, 3, 7):  # labels in [0,1] for the sigmoid losses
            data = data[:6] + ((data[6] > 3).astype(np.float32),) + data[7:]
        return 1, act, params, data, "ug"
    data = synth.random_general(500, NU_R, NI_R, NG_R, seed=seed, max_g=int(rng.integers(0, 6)),
                                max_u=int(rng.integers(1, 4)), max_i=int(rng.integers(1, 5)),
                                allow_dup=bool(rng.integers(0, 2)))
    if act in (1, 2, 3, 7):
        data = _cases._binary(data)
    return 0, act, params, data, "csr"


NU_R, NI_R, NG_R = 60, 40, 12


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", range(48))
def test_oracle_matches_reference_on_random_configurations(seed, tmp_path):
    fmt, act, params, data, kind = _random_config(seed)
    mo, po = _train(COracle, fmt, act, params, data, kind, tmp_path, rounds=3)
    mr, pr = _train(RefTrainer, fmt, act, params, data, kind, tmp_path, rounds=3)
    assert mo == mr, ("model bytes differ", params, act)
    assert np.array_equal(po, pr, equal_nan=True), ("predictions differ", params, act)
