"""GPU: SVDFeatureRanker (base.h:597-813) -- the C ABI's svdgpu_rank_* and the C++
GpuSVDRanker against the CPU oracle's restatement (itself pinned to the compiled reference by
tests/test_oracle_vs_ref.py).  Results are integers (item positions / rank positions): the bar is
bit-exact."""
import numpy as np
import pytest

import _cases
from _oracle import COracle, COracleRanker
from svdfeature_b200 import synth

pytestmark = pytest.mark.gpu
RANK_CASES = _cases.rank_cases()


def _gpu_with_model(native, path, fmt, params):
    o = COracle(fmt, 0, 0)
    o.load_model(path)
    g = native.SvdGpu(**_cases.shape_of(params, fmt, 0))
    g.set_hparams(base_score=o.base_score)
    g.upload(*[a.copy() for a in o.arrays()])
    return g


def _case(name, tmp_path):
    fmt, params, skw, rparams = RANK_CASES[name]
    path = _cases.rank_model(fmt, params, tmp_path)
    stream = synth.rank_stream(num_user=params["num_user"], num_item=params["num_item"], **skw)
    kind = "ug" if skw.get("ugroup") else "csr"
    return fmt, params, skw, rparams, path, stream, kind


@pytest.mark.parametrize("name", sorted(RANK_CASES))
def test_rank_abi_matches_oracle(native, name, tmp_path):
    fmt, params, skw, rparams, path, stream, kind = _case(name, tmp_path)
    want = COracleRanker(path, skw["num_item_set"], rparams).rank(stream, kind)
    g = _gpu_with_model(native, path, fmt, params)
    g.rank_init(skw["num_item_set"], rparams.get("top_k", 0))
    got = g.rank(stream, kind)
    assert len(want) > 0 and np.array_equal(got, want)
    assert g.counter("kernel_launches") > 0


@pytest.mark.parametrize("name", ["rank_pos_k20", "rank_top5_k20"])
@pytest.mark.parametrize("piece", [1, 7, 50])
def test_rank_stream_can_be_cut_anywhere(native, name, piece, tmp_path):
    """The stream state machine keeps open sections across calls (the per-row ISVDRanker::process
    is the piece = 1 case)."""
    fmt, params, skw, rparams, path, stream, kind = _case(name, tmp_path)
    want = COracleRanker(path, skw["num_item_set"], rparams).rank(stream, kind)
    g = _gpu_with_model(native, path, fmt, params)
    g.rank_init(skw["num_item_set"], rparams.get("top_k", 0))
    rp, lab, idx, val = stream
    got = []
    for r0 in range(0, len(lab), piece):
        r1 = min(r0 + piece, len(lab))
        got.append(g.rank((rp[3 * r0:3 * r1 + 1], lab[r0:r1], idx, val)))
    assert np.array_equal(np.concatenate(got), want)


@pytest.mark.parametrize("name", sorted(RANK_CASES))
def test_gpu_ranker_class_matches_oracle(native, name, tmp_path):
    """create_svd_ranker -> GpuSVDRanker through the reference's own ISVDRanker virtuals."""
    fmt, params, skw, rparams, path, stream, kind = _case(name, tmp_path)
    want = COracleRanker(path, skw["num_item_set"], rparams).rank(stream, kind)
    r = native.GpuRanker(path, skw["num_item_set"], rparams)
    assert np.array_equal(r.rank(stream, kind), want)


@pytest.mark.parametrize("top_k", [0, 6])
def test_rank_side_features(native, top_k, tmp_path):
    fmt, params, skw, rparams, path, stream, kind = _case("rank_pos_k20", tmp_path)
    fu, fi = str(tmp_path / "user.side"), str(tmp_path / "item.side")
    su = _cases.write_side_features(fu, params["num_user"], params["num_user"], seed=1)
    si = _cases.write_side_features(fi, params["num_item"], params["num_item"], seed=2)
    rp = {"feature_user": fu, "feature_item": fi, "top_k": top_k}
    want = COracleRanker(path, skw["num_item_set"], rp).rank(stream)
    g = _gpu_with_model(native, path, fmt, params)
    g.set_side_features(0, su)
    g.set_side_features(1, si)
    g.rank_init(skw["num_item_set"], top_k)
    assert np.array_equal(g.rank(stream), want)
    assert np.array_equal(native.GpuRanker(path, skw["num_item_set"], rp).rank(stream), want)


@pytest.mark.parametrize("top_k", [0, 10, 40, -10])
def test_rank_larger_set(native, top_k, tmp_path):
    """2 500 candidates x 400 users, k = 64 (one million scores): still bit-exact, and every top_k
    answer is a set of distinct, non-banned candidates."""
    params = dict(_cases.BASE, num_user=3000, num_item=2000, num_factor=64)
    path = _cases.rank_model(0, params, tmp_path)
    skw = dict(num_item_set=2500, num_sections=400, seed=9, max_pos=12, max_ban=40)
    stream = synth.rank_stream(num_user=3000, num_item=2000, **skw)
    force_sort = top_k < 0  # small top_k normally takes the selection kernel: cover the sort path too
    top_k = abs(top_k)
    want = COracleRanker(path, 2500, {"top_k": top_k}).rank(stream)
    g = _gpu_with_model(native, path, 0, params)
    g.set_option("rank_force_sort", int(force_sort))
    g.rank_init(2500, top_k)
    got = g.rank(stream)
    assert np.array_equal(got, want)
    if top_k:
        top = got.reshape(400, top_k)
        assert all(len(set(row)) == top_k for row in top)
    else:
        assert got.min() >= 0 and got.max() < 2500


@pytest.mark.parametrize("top_k", [0, 40])
def test_gpu_ranker_class_large_results(native, top_k, tmp_path):
    """GpuSVDRanker's result buffer grows with the stream: top_k = 40 and users with up to 60 POS
    items (a PROCESS row then returns more than any earlier one) through the ISVDRanker virtuals."""
    params = dict(_cases.BASE, num_user=3000, num_item=2000, num_factor=16)
    path = _cases.rank_model(0, params, tmp_path)
    skw = dict(num_item_set=300, num_sections=40, seed=11, max_pos=60, max_ban=10)
    stream = synth.rank_stream(num_user=3000, num_item=2000, **skw)
    want = COracleRanker(path, 300, {"top_k": top_k}).rank(stream)
    got = native.GpuRanker(path, 300, {"top_k": top_k}).rank(stream)
    assert len(want) > 40 * 16 and np.array_equal(got, want)


def test_rank_errors(native, tmp_path):
    """The reference's assert messages (base.h:720,751-752,758,760,775)."""
    fmt, params, skw, rparams, path, stream, kind = _case("rank_pos_k20", tmp_path)
    g = _gpu_with_model(native, path, fmt, params)
    with pytest.raises(native.SvdGpuError, match="rank_init"):
        g.rank(stream)
    one = np.ones(1, np.float32)
    item = lambda i: (synth.RK_ITEM, [], [], [(i, 0.5)])
    g.rank_init(2, 0)
    with pytest.raises(native.SvdGpuError, match="item instance exceed specified item set size"):
        g.rank(synth.ragged_csr([item(1), item(2), item(3)]))
    g.rank_init(3, 0)
    g.rank(synth.ragged_csr([item(1), item(2), item(3), (synth.RK_USER, [], [(5, 1.0)], [])]))
    with pytest.raises(native.SvdGpuError, match="sample item index exceed bound"):
        g.rank(synth.ragged_csr([(synth.RK_POS, [], [(3, 1.0)], [])]))
    g.rank(synth.ragged_csr([(synth.RK_POS, [], [(1, 1.0)], [])]))
    with pytest.raises(native.SvdGpuError, match="can not occur in baned sample list"):
        g.rank(synth.ragged_csr([(synth.RK_BAN, [], [(1, 1.0)], [])]))
    with pytest.raises(native.SvdGpuError, match="must specify item index"):
        g.rank(synth.ragged_csr([(synth.RK_SPEC, [], [(0, 1.0), (2, 1.0)], [])]))
    got = g.rank(synth.ragged_csr([(synth.RK_PROCESS, [], [], [])]))
    assert len(got) == 1 and 0 <= got[0] < 3
    g.rank_init(3, 3)
    g.rank(synth.ragged_csr([item(1), item(2), item(3), (synth.RK_USER, [], [(5, 1.0)], []),
                             (synth.RK_BAN, [], [(0, 1.0)], [])]))
    with pytest.raises(native.SvdGpuError, match="k can not exceed candidate size"):
        g.rank(synth.ragged_csr([(synth.RK_PROCESS, [], [], [])]))
    del one
