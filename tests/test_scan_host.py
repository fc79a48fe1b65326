"""Host-side checks of the compact H2D path (svdfeature_b200/csrc/svdgpu_scan.h): plain C++,
compiled here with g++ and compared with numpy on regular, nearly regular and ragged inputs."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = r"""
#include "svdgpu_scan.h"
extern "C" int scan_rp(const int *p, long long n, int *abc) { return svdscan::rp_regular(p, n, abc[0], abc[1], abc[2]) ? 1 : 0; }
extern "C" int scan_ones(const float *v, long long n) { return svdscan::all_ones(v, n) ? 1 : 0; }
// what run_csr_host does with the pool: start, take the verdicts in chunk order
extern "C" void scan_pool(const int *rp, const float *v, int num_row, int chunk_rows, int threads, int *out) {
  const int nchunk = (num_row + chunk_rows - 1) / chunk_rows;
  svdscan::ScanPool pool(nchunk);
  pool.start(num_row, chunk_rows, rp, v, threads);
  for (int c = 0; c < nchunk; ++c) {
    const svdscan::ChunkScan &s = pool.wait(c);
    out[5 * c] = s.rp_regular;
    out[5 * c + 1] = s.val_ones;
    out[5 * c + 2] = s.a;
    out[5 * c + 3] = s.b;
    out[5 * c + 4] = s.c;
  }
}
// leaving early (an error return of the caller) must join the workers
extern "C" void scan_pool_abandon(const int *rp, const float *v, int num_row, int chunk_rows, int threads) {
  const int nchunk = (num_row + chunk_rows - 1) / chunk_rows;
  svdscan::ScanPool pool(nchunk);
  pool.start(num_row, chunk_rows, rp, v, threads);
  if (nchunk > 0) pool.wait(0);
}
"""


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("scan")
    src = d / "shim.cpp"
    src.write_text(SHIM)
    so = d / "libscan.so"
    subprocess.check_call(["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I", os.path.join(ROOT, "svdfeature_b200", "csrc"),
                           "-o", str(so), str(src)])
    lib = C.CDLL(str(so))
    lib.scan_rp.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p]
    lib.scan_ones.argtypes = [C.c_void_p, C.c_longlong]
    return lib


def _rp(n, a, b, c, v0=0):
    r = np.arange(n + 1, dtype=np.int64) * (a + b + c) + v0
    rp = np.empty(3 * n + 1, np.int32)
    rp[0::3] = r
    rp[1::3] = r[:-1] + a
    rp[2::3] = r[:-1] + a + b
    return rp


def _scan(lib, rp, n):
    abc = np.zeros(3, np.int32)
    ok = lib.scan_rp(rp.ctypes.data, n, abc.ctypes.data)
    return bool(ok), tuple(int(x) for x in abc)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 1023, 1024, 1025, 4099, 100003])
@pytest.mark.parametrize("shape", [(0, 1, 1), (0, 1, 2), (8, 1, 1), (0, 0, 0), (3, 0, 5)])
def test_rp_regular_accepts_constant_shapes(lib, n, shape):
    for v0 in (0, 12345):
        rp = _rp(n, *shape, v0=v0)
        ok, abc = _scan(lib, rp, n)
        assert ok and abc == shape


@pytest.mark.parametrize("n", [2, 4, 5, 1024, 1030, 50001])
def test_rp_regular_rejects_any_single_deviation(lib, n):
    rng = np.random.default_rng(n)
    base = _rp(n, 0, 1, 1, v0=7)
    # every position class: first rows, block borders, the tail, the final entry
    spots = {0, 1, 2, 3, 3 * n, 3 * n - 1, 3 * (n // 2), 3 * (n // 2) + 1} | set(rng.integers(0, 3 * n + 1, 40).tolist())
    for j in sorted(s for s in spots if 0 <= s <= 3 * n):
        for delta in (1, -1):
            rp = base.copy()
            rp[j] += delta
            ok, _ = _scan(lib, rp, n)
            assert not ok, (n, j, delta)


def test_rp_regular_rejects_ragged_and_degenerate(lib):
    # same total as a regular batch, different split
    rp = _rp(10, 0, 1, 1)
    rp[3 * 4 + 2] += 1  # row 4: (0|2|0) instead of (0|1|1)
    assert not _scan(lib, rp, 10)[0]
    assert not _scan(lib, np.zeros(1, np.int32), 0)[0]  # no rows
    neg = _rp(5, 0, 1, 1)
    neg -= 3  # negative base offset
    assert not _scan(lib, neg, 5)[0]
    dec = np.array([4, 4, 3, 5], np.int32)  # decreasing inside a row
    assert not _scan(lib, dec, 1)[0]
    assert _scan(lib, np.array([7, 8, 8, 9], np.int32), 1) == (True, (1, 0, 1))  # one row is its own shape


def test_rp_regular_near_int_max(lib):
    n, w = 1000, 2
    v0 = 2**31 - 1 - n * w
    rp = _rp(n, 0, 1, 1, v0=v0)
    assert rp[-1] == 2**31 - 1
    assert _scan(lib, rp, n) == (True, (0, 1, 1))


@pytest.mark.parametrize("n", [0, 1, 5, 8191, 8192, 8193, 300001])
def test_all_ones(lib, n):
    v = np.ones(max(n, 1), np.float32)
    assert lib.scan_ones(v.ctypes.data, n) == 1
    if n == 0:
        return
    rng = np.random.default_rng(n)
    for j in {0, n - 1, n // 2} | set(rng.integers(0, n, 10).tolist()):
        for bad in (0.99999994, 1.0000001, -1.0, 0.0, np.nan):
            w = v.copy()
            w[j] = bad
            assert lib.scan_ones(w.ctypes.data, n) == 0, (n, j, bad)


@pytest.mark.parametrize("threads", [0, 1, 3, 16])
@pytest.mark.parametrize("chunk_rows", [1000, 777, 100000])
def test_scan_pool_verdicts_per_chunk(lib, threads, chunk_rows):
    """The batch of test_compact_h2d_is_invisible: unit-valued basic rows, basic rows with other
    values, a ragged tail -- per chunk the pool must say exactly what numpy says."""
    n = 6000
    nf = np.array([2] * 4800 + [2 if r % 3 else 1 for r in range(4800, n)], np.int64)
    nu = np.array([1] * 4800 + [1 if r % 3 else 0 for r in range(4800, n)], np.int64)
    start = np.concatenate([[0], np.cumsum(nf)]) + 40  # (the slice does not start at entry 0)
    rp = np.empty(3 * n + 1, np.int32)
    rp[0::3] = start
    rp[1::3] = start[:-1]
    rp[2::3] = start[:-1] + nu
    val = np.ones(int(start[-1]), np.float32)
    val[start[3500]:start[4800]] = 0.5
    nchunk = (n + chunk_rows - 1) // chunk_rows
    out = np.full(5 * nchunk, -1, np.int32)
    lib.scan_pool.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.scan_pool(rp.ctypes.data, val.ctypes.data, n, chunk_rows, threads, out.ctypes.data)
    for c in range(nchunk):
        r0, r1 = c * chunk_rows, min(n, (c + 1) * chunk_rows)
        regular = bool(np.all(nf[r0:r1] == nf[r0]) and np.all(nu[r0:r1] == nu[r0]))
        ones = bool(np.all(val[start[r0]:start[r1]] == 1.0))
        assert bool(out[5 * c]) == regular, (c, "row_ptr")
        assert bool(out[5 * c + 1]) == ones, (c, "values")
        if regular:
            assert tuple(out[5 * c + 2:5 * c + 5]) == (0, int(nu[r0]), int(nf[r0] - nu[r0]))


@pytest.mark.parametrize("threads", [0, 2, 8])
def test_scan_pool_can_be_abandoned(lib, threads):
    n = 200000
    rp = _rp(n, 0, 1, 1)
    val = np.ones(2 * n, np.float32)
    lib.scan_pool_abandon.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    for _ in range(20):
        lib.scan_pool_abandon(rp.ctypes.data, val.ctypes.data, n, 1000, threads)
