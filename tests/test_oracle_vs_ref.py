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
