"""ctypes bindings for the CPU oracles (TEST INFRASTRUCTURE ONLY).

Two interchangeable checkers, both driven with the same (name, value) strings
the reference's ``set_param`` chain takes:

* ``COracle``   -- oracle/liboracle.so, the plain-C restatement (svdf_oracle.c)
* ``RefTrainer``-- oracle/_ref/libsvdf_ref.so, the UNMODIFIED reference compiled
  from /root/reference and driven through ISVDTrainer (csrc/trainer_cabi.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this.
"""
import ctypes as C
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libsvdf_ref.so")
REFERENCE_ROOT = "/root/reference"

_i32p = C.POINTER(C.c_int)
_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def build_oracle():
    """Compile the C restatement (and, when /root/reference exists, oracle/_ref)."""
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
        os.path.join(ORACLE_DIR, "svdf_oracle.c")
    ):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
    if os.path.isdir(REFERENCE_ROOT) and not os.path.exists(REF_SO):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL)


def have_ref():
    return os.path.exists(REF_SO)


def _csr_args(csr):
    row_ptr, label, index, value = csr
    assert row_ptr.dtype == np.int32 and label.dtype == np.float32
    assert index.dtype == np.uint32 and value.dtype == np.float32
    return (len(label), _p(row_ptr, _i32p), _p(label, _f32p), _p(index, _u32p), _p(value, _f32p))


def _ug_args(ug):
    blk_row_off, blk_fb_off, blk_tag, fb_index, fb_value = ug[:5]
    csr = ug[5:]
    n, rp, lb, ix, vl = _csr_args(csr)
    return (
        len(blk_row_off) - 1,
        _p(blk_row_off, _i32p),
        _p(blk_fb_off, _i32p),
        _p(blk_tag, _i32p),
        _p(fb_index, _u32p),
        _p(fb_value, _f32p),
        rp,
        lb,
        ix,
        vl,
    ), n


class _Base:
    prefix = None
    lib = None

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def set_params(self, params):
        for k, v in params.items():
            if k == "wd_ranges":  # ordered (key, value) pairs: up:wd / up:bound / ip:... / gp:...
                for kk, vv in v:
                    self._f("set_param")(self.h, str(kk).encode(), str(vv).encode())
                continue
            self._f("set_param")(self.h, str(k).encode(), str(v).encode())

    def init(self, seed=10):
        """seed -> init_model -> init_trainer (svd_feature.cpp:194-218, 292-296)."""
        self._f("seed")(C.c_uint(seed))
        self._f("init_model")(self.h)
        self._f("init_trainer")(self.h)

    def set_round(self, r):
        self._f("set_round")(self.h, int(r))

    def update_csr(self, csr):
        self._f("update_csr")(self.h, *_csr_args(csr))

    def predict_csr(self, csr):
        out = np.empty(len(csr[1]), np.float32)
        self._f("predict_csr")(self.h, *_csr_args(csr), _p(out, _f32p))
        return out

    def update_ugroup(self, ug):
        args, _ = _ug_args(ug)
        self._f("update_ugroup")(self.h, *args)

    def predict_ugroup(self, ug):
        args, n = _ug_args(ug)
        out = np.empty(n, np.float32)
        self._f("predict_ugroup")(self.h, *args, _p(out, _f32p))
        return out

    def save_model(self, path):
        assert self._f("save_model")(self.h, path.encode()) == 0

    def load_model(self, path):
        assert self._f("load_model")(self.h, path.encode()) == 0

    def model_bytes(self, tmpdir):
        path = os.path.join(str(tmpdir), "m_%s_%d.model" % (self.prefix, id(self)))
        self.save_model(path)
        with open(path, "rb") as f:
            return f.read()

    def close(self):
        if self.h:
            self._f("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _declare(lib, prefix):
    vp = C.c_void_p
    sig = {
        "create": (vp, [C.c_int, C.c_int, C.c_int]),
        "destroy": (None, [vp]),
        "seed": (None, [C.c_uint]),
        "set_param": (None, [vp, C.c_char_p, C.c_char_p]),
        "init_model": (None, [vp]),
        "init_trainer": (None, [vp]),
        "set_round": (None, [vp, C.c_int]),
        "save_model": (C.c_int, [vp, C.c_char_p]),
        "load_model": (C.c_int, [vp, C.c_char_p]),
        "update_csr": (None, [vp, C.c_int, _i32p, _f32p, _u32p, _f32p]),
        "predict_csr": (None, [vp, C.c_int, _i32p, _f32p, _u32p, _f32p, _f32p]),
        "update_ugroup": (None, [vp, C.c_int, _i32p, _i32p, _i32p, _u32p, _f32p, _i32p, _f32p, _u32p, _f32p]),
        "predict_ugroup": (None, [vp, C.c_int, _i32p, _i32p, _i32p, _u32p, _f32p, _i32p, _f32p, _u32p, _f32p, _f32p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, prefix + name)
        fn.restype = res
        fn.argtypes = args


class COracle(_Base):
    """oracle/liboracle.so -- plain-C restatement."""

    prefix = "svdo_"

    def __init__(self, format_type=2, active_type=0, extend_type=0, params=None):
        if COracle.lib is None:
            build_oracle()
            COracle.lib = C.CDLL(ORACLE_SO)
            _declare(COracle.lib, self.prefix)
            COracle.lib.svdo_data.restype = _f32p
            COracle.lib.svdo_data.argtypes = [C.c_void_p, C.c_int]
            COracle.lib.svdo_info.restype = C.c_long
            COracle.lib.svdo_info.argtypes = [C.c_void_p, C.c_int]
            COracle.lib.svdo_base_score.restype = C.c_float
            COracle.lib.svdo_base_score.argtypes = [C.c_void_p]
            COracle.lib.svdo_learning_rate.restype = C.c_float
            COracle.lib.svdo_learning_rate.argtypes = [C.c_void_p]
        self.h = self.lib.svdo_create(format_type, active_type, extend_type)
        if params:
            self.set_params(params)

    def info(self, what):
        return int(self.lib.svdo_info(self.h, what))

    def arrays(self):
        """(ui_bias[rows], W_uiset[rows, pitch], g_bias[num_global]) as numpy VIEWS."""
        rows, pitch, ng = self.info(0), self.info(1), self.info(6)
        ub = np.ctypeslib.as_array(self.lib.svdo_data(self.h, 0), shape=(max(rows, 1),))[:rows]
        W = np.ctypeslib.as_array(self.lib.svdo_data(self.h, 1), shape=(max(rows, 1), max(pitch, 1)))[:rows]
        gb = np.ctypeslib.as_array(self.lib.svdo_data(self.h, 2), shape=(max(ng, 1),))[:ng]
        return ub, W, gb

    @property
    def base_score(self):
        return float(self.lib.svdo_base_score(self.h))

    @property
    def learning_rate(self):
        return float(self.lib.svdo_learning_rate(self.h))


class RefTrainer(_Base):
    """oracle/_ref/libsvdf_ref.so -- the compiled, unmodified reference."""

    prefix = "svdtr_"

    def __init__(self, format_type=2, active_type=0, extend_type=0, params=None):
        if RefTrainer.lib is None:
            build_oracle()
            RefTrainer.lib = C.CDLL(REF_SO)
            _declare(RefTrainer.lib, self.prefix)
        self.h = self.lib.svdtr_create(format_type, active_type, extend_type)
        if params:
            self.set_params(params)

    def load_model(self, path):
        # The fork's SVDModel::load_from_file (apex_svd_model.h:586-621) dumps
        # u_bias.txt / w_user.txt / ... into the CWD: run it from a scratch dir.
        cwd = os.getcwd()
        os.chdir(os.path.dirname(os.path.abspath(path)))
        try:
            assert self.lib.svdtr_load_model(self.h, os.path.abspath(path).encode()) == 0
        finally:
            os.chdir(cwd)


# ---------------------------------------------------------------------------
# SVDFeatureRanker (base.h:597-813) behind the same two checkers
# ---------------------------------------------------------------------------
def _rank_call(fn, h, stream, kind, cap):
    out = np.full(cap, -12345, np.int32)
    if kind == "csr":
        n = fn(h, *_csr_args(stream), _p(out, _i32p), cap)
    else:
        args, _ = _ug_args(stream)
        n = fn(h, *args, _p(out, _i32p), cap)
    assert 0 <= n <= cap, n
    return out[:n].copy()


_RANK_CSR = [C.c_void_p, C.c_int, _i32p, _f32p, _u32p, _f32p, _i32p, C.c_long]
_RANK_UG = [C.c_void_p, C.c_int, _i32p, _i32p, _i32p, _u32p, _f32p, _i32p, _f32p, _u32p, _f32p, _i32p, C.c_long]


class COracleRanker:
    """The restated ranker: load_model -> set_param -> init_ranker -> tagged stream."""

    def __init__(self, model_path, num_item_set, params=None):
        with open(model_path, "rb") as f:
            fmt, act, ext, _ = struct.unpack("<4B", f.read(4))
        self.o = COracle(fmt, act, ext)
        self.o.load_model(model_path)
        self.o.set_params(params or {})
        lib = self.o.lib
        lib.svdo_init_ranker.argtypes = [C.c_void_p, C.c_int]
        lib.svdo_rank_csr.restype = C.c_long
        lib.svdo_rank_csr.argtypes = _RANK_CSR
        lib.svdo_rank_ugroup.restype = C.c_long
        lib.svdo_rank_ugroup.argtypes = _RANK_UG
        lib.svdo_init_ranker(self.o.h, num_item_set)

    def rank(self, stream, kind="csr", cap=1 << 20):
        fn = self.o.lib.svdo_rank_csr if kind == "csr" else self.o.lib.svdo_rank_ugroup
        return _rank_call(fn, self.o.h, stream, kind, cap)


class RefRanker:
    """The compiled reference's SVDFeatureRanker through create_svd_ranker (apex_svd.cpp:45-47)."""

    def __init__(self, model_path, num_item_set, params=None):
        build_oracle()
        if RefTrainer.lib is None:
            RefTrainer.lib = C.CDLL(REF_SO)
            _declare(RefTrainer.lib, RefTrainer.prefix)
        lib = self.lib = RefTrainer.lib
        lib.svdrk_create_from_model.restype = C.c_void_p
        lib.svdrk_create_from_model.argtypes = [C.c_char_p]
        lib.svdrk_destroy.argtypes = [C.c_void_p]
        lib.svdrk_set_param.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        lib.svdrk_init_ranker.argtypes = [C.c_void_p, C.c_int]
        lib.svdrk_rank_csr.restype = C.c_long
        lib.svdrk_rank_csr.argtypes = _RANK_CSR
        lib.svdrk_rank_ugroup.restype = C.c_long
        lib.svdrk_rank_ugroup.argtypes = _RANK_UG
        # the fork's load_from_file dumps *.txt into the CWD (apex_svd_model.h:586-621)
        cwd = os.getcwd()
        os.chdir(os.path.dirname(os.path.abspath(model_path)))
        try:
            self.h = lib.svdrk_create_from_model(os.path.abspath(model_path).encode())
        finally:
            os.chdir(cwd)
        assert self.h
        for k, v in (params or {}).items():
            lib.svdrk_set_param(self.h, str(k).encode(), str(v).encode())
        lib.svdrk_init_ranker(self.h, num_item_set)

    def rank(self, stream, kind="csr", cap=1 << 20):
        fn = self.lib.svdrk_rank_csr if kind == "csr" else self.lib.svdrk_rank_ugroup
        return _rank_call(fn, self.h, stream, kind, cap)

    def __del__(self):
        try:
            if self.h:
                self.lib.svdrk_destroy(self.h)
                self.h = None
        except Exception:
            pass


# ---------------------------------------------------------------------------
# model-file reader (layout: SURVEY.md section 8b; apex_svd_model.h:638-660)
# ---------------------------------------------------------------------------
def parse_model(buf):
    """Decode a model file into a dict of numpy arrays (independent of both oracles)."""
    off = 0
    fmt, act, ext, var = struct.unpack_from("<4B", buf, off)
    off += 4
    ints = struct.unpack_from("<264i", buf, off)
    flts = struct.unpack_from("<264f", buf, off)
    off += 1056
    p = dict(
        num_user=ints[0], num_item=ints[1], num_factor=ints[2], num_global=ints[3],
        u_init_sigma=flts[4], i_init_sigma=flts[5], base_score=flts[6], no_user_bias=ints[7],
        num_ufeedback=ints[8], ufeedback_init_sigma=flts[9],
    )

    def t1(off):
        (n,) = struct.unpack_from("<i", buf, off)
        a = np.frombuffer(buf, np.float32, n, off + 4)
        return a, off + 4 + 4 * n

    def t2(off):
        x, y = struct.unpack_from("<2i", buf, off)
        a = np.frombuffer(buf, np.float32, x * y, off + 8).reshape(y, x)
        return a, off + 8 + 4 * x * y

    out = dict(format_type=fmt, active_type=act, extend_type=ext, param=p)
    out["u_bias"], off = t1(off)
    out["W_user"], off = t2(off)
    out["i_bias"], off = t1(off)
    out["W_item"], off = t2(off)
    out["g_bias"], off = t1(off)
    if fmt == 1:
        out["ufeedback_bias"], off = t1(off)
        out["W_ufeedback"], off = t2(off)
    assert off == len(buf), (off, len(buf))
    return out
