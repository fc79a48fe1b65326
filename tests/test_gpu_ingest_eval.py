"""GPU: bulk ingest of the reference's buffer files and the device-side RMSE
evaluation (f4), both through the C ABI, against the oracle fed with the same rows."""
import os

import numpy as np
import pytest

import _cases
from _oracle import COracle
from svdfeature_b200 import buffer_io

pytestmark = pytest.mark.gpu
CASES = _cases.cases()
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _pair(native, name, mode, chunk_rows=None):
    fmt, act, params, data, kind = CASES[name]
    o = COracle(fmt, act, 0, params)
    o.init(10)
    g = native.SvdGpu(**_cases.shape_of(params, fmt, act))
    g.set_hparams(**_cases.hparams_of(params, o.base_score))
    g.set_mode(mode)
    if chunk_rows:
        g.set_option("chunk_rows", chunk_rows)
    g.upload(*[a.copy() for a in o.arrays()])
    return o, g, data, kind


def _sse(pred, label, scale=1.0):
    d = ((pred - label) * np.float32(scale)).astype(np.float64)
    return float(np.sum(d * d))


@pytest.mark.parametrize("name,batch_size,chunk_rows", [("basic_k64", 1000, None), ("general_k13_dups", 37, 500),
                                                        ("neighborhood_k32", 256, 300), ("basic_k16", 5000, 1 << 20)])
def test_feature_buffer_file_pass_equals_rowwise_oracle(native, name, batch_size, chunk_rows, tmp_path):
    """A training pass + a prediction pass + an evaluation pass over a BINARY_BUFFER file (written in
    the byte layout tools/make_feature_buffer produces) equal the oracle's row-by-row loop."""
    o, g, data, kind = _pair(native, name, native.MODE_EXACT, chunk_rows)
    path = str(tmp_path / "train.buffer")
    buffer_io.write_feature_buffer(path, data, batch_size=batch_size)
    n = g.update_buffer_file(path)
    assert n == len(data[1])
    o.update_csr(data)
    ub, W, gb = o.arrays()
    gub, gW, ggb = g.download()
    k = o.info(3)
    assert np.array_equal(ub, gub) and np.array_equal(W[:, :k], gW[:, :k]) and np.array_equal(gb, ggb)
    po = o.predict_csr(data)
    pg = g.predict_buffer_file(path, len(data[1]))
    assert np.array_equal(po, pg)
    sse, cnt = g.eval_buffer_file(path, scale=0.5)
    assert cnt == len(data[1])
    assert abs(sse - _sse(po, data[1], 0.5)) <= 1e-9 * max(1.0, sse)
    with pytest.raises(native.SvdGpuError, match="output holds"):
        g.predict_buffer_file(path, len(data[1]) - 1)
    with pytest.raises(native.SvdGpuError, match="can't open"):
        g.update_buffer_file(str(tmp_path / "missing.buffer"))


def test_reference_shipped_buffer_is_ingested(native):
    """demo/basicMF/ua.base.buffer as shipped by the reference (tests/golden/ua.base.buffer)."""
    path = os.path.join(GOLDEN, "ua.base.buffer")
    csr, hdr = buffer_io.read_feature_buffer(path)
    params = dict(num_user=943, num_item=1682, num_factor=8, learning_rate=0.005, wd_user=0.004, wd_item=0.004,
                  base_score=3.6)
    o = COracle(0, 0, 0, params)
    o.init(10)
    g = native.SvdGpu(943, 1682, 8)
    g.set_hparams(learning_rate=0.005, wd_user=0.004, wd_item=0.004, base_score=o.base_score)
    g.set_mode(native.MODE_EXACT)
    g.upload(*[a.copy() for a in o.arrays()])
    assert g.update_buffer_file(path) == len(csr[1])
    o.update_csr(csr)
    assert np.array_equal(o.predict_csr(csr), g.predict_buffer_file(path, len(csr[1])))


@pytest.mark.parametrize("name,chunk_rows", [("svdpp_k16", None), ("svdpp_k64_tags", 200), ("pairwise_ugroup", 64)])
def test_ugroup_buffer_file_pass_equals_oracle(native, name, chunk_rows, tmp_path):
    o, g, data, kind = _pair(native, name, native.MODE_EXACT, chunk_rows)
    path = str(tmp_path / "train.ugroup.buffer")
    buffer_io.write_ugroup_buffer(path, data)
    assert g.update_buffer_file(path) == len(data[6])
    o.update_ugroup(data)
    po = o.predict_ugroup(data)
    pg = g.predict_buffer_file(path, len(data[6]))
    tol = 2e-6 if name in _cases.SIGMOID_CASES else 0.0
    assert np.abs(po - pg).max() <= tol
    sse, cnt = g.eval_buffer_file(path)
    assert cnt == len(data[6])
    assert abs(sse - _sse(pg, data[6])) <= 1e-9 * max(1.0, sse)


def test_reference_shipped_ugroup_buffer_is_ingested(native):
    path = os.path.join(GOLDEN, "implicit.buffer.svdpp")
    ug, hdr = buffer_io.read_ugroup_buffer(path)
    nfb = int(ug[3].max()) + 1 if len(ug[3]) else 1
    nu = int(ug[7][ug[5][1:-1:3]].max()) + 1
    ni = int(ug[7].max()) + 1
    params = dict(num_user=nu, num_item=ni, num_ufeedback=nfb, num_factor=8, learning_rate=0.005, wd_user=0.004,
                  wd_item=0.004, wd_ufeedback=0.004, base_score=3.6, ufeedback_init_sigma=0.01)
    o = COracle(1, 0, 0, params)
    o.init(10)
    g = native.SvdGpu(**_cases.shape_of(params, 1, 0))
    g.set_hparams(**_cases.hparams_of(params, o.base_score))
    g.set_mode(native.MODE_EXACT)
    g.upload(*[a.copy() for a in o.arrays()])
    assert g.update_buffer_file(path) == len(ug[6])
    o.update_ugroup(ug)
    assert np.array_equal(o.predict_ugroup(ug), g.predict_buffer_file(path, len(ug[6])))


@pytest.mark.parametrize("mode", ["exact", "hogwild"])
def test_device_eval_matches_host_rmse(native, mode):
    """svdgpu_eval_csr / svdgpu_batch_eval = the reference's RMSEEvaluator over the same predictions
    (svd_feature_infer.cpp:45-52), with only two doubles crossing PCIe."""
    o, g, data, kind = _pair(native, "general_k40", native.MODE_EXACT if mode == "exact" else native.MODE_HOGWILD, 700)
    g.update_csr(data)
    pred = g.predict_csr(data)
    d0 = g.counter("d2h_bytes")
    sse, cnt = g.eval_csr(data, scale=2.0)
    assert g.counter("d2h_bytes") - d0 == 16
    assert cnt == len(data[1])
    assert abs(sse - _sse(pred, data[1], 2.0)) <= 1e-9 * sse
    b = g.batch_create(data)
    sse_b, cnt_b = g.batch_eval(b, 100, 1100, scale=2.0)
    assert cnt_b == 1000
    assert abs(sse_b - _sse(pred[100:1100], data[1][100:1100], 2.0)) <= 1e-9 * sse_b
    # empty input
    empty = (np.zeros(1, np.int32), np.zeros(0, np.float32), np.zeros(0, np.uint32), np.zeros(0, np.float32))
    assert g.eval_csr(empty) == (0.0, 0)
    b.close()


def test_device_eval_ugroup(native):
    o, g, data, kind = _pair(native, "svdpp_k16", native.MODE_EXACT)
    g.update_ugroup(data)
    pred = g.predict_ugroup(data)
    sse, cnt = g.eval_ugroup(data)
    assert cnt == len(data[6])
    assert abs(sse - _sse(pred, data[6])) <= 1e-9 * sse
