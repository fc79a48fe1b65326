"""ctypes bindings of the native libraries (no compute happens in Python).

``SvdGpu``     -- thin wrapper over the C ABI of include/svdgpu.h (libsvdgpu.so)
``GpuTrainer`` -- the C++ ``GpuSVDFeature : ISVDTrainer`` (libsvdf_gpu.so) driven
                  through the same ``svdtr_*`` shim the reference is wrapped with,
                  so a test can run the reference and the GPU trainer with
                  identical calls.

Both fail loudly when the CUDA library is missing or no GPU is present: there
is no CPU fallback in this package.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_GPU = os.environ.get("SVDGPU_LIB") or os.path.join(HERE, "libsvdgpu.so")  # (override: kernel A/B experiments)
LIB_TRAINER = os.path.join(HERE, "libsvdf_gpu.so")

MODE_EXACT, MODE_HOGWILD = 0, 1

_i32p = C.POINTER(C.c_int)
_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint)
_vp = C.c_void_p


class Shape(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "num_user", "num_item", "num_global", "num_ufeedback", "num_factor", "no_user_bias",
        "active_type", "format_type")]


class HParams(C.Structure):
    _fields_ = [
        ("learning_rate", C.c_float), ("wd_user", C.c_float), ("wd_item", C.c_float),
        ("wd_user_bias", C.c_float), ("wd_item_bias", C.c_float), ("wd_global", C.c_float),
        ("reg_method", C.c_int), ("reg_global", C.c_int), ("num_regfree_global", C.c_uint),
        ("scale_lr_ufeedback", C.c_float), ("wd_ufeedback", C.c_float),
        ("wd_ufeedback_bias", C.c_float), ("base_score", C.c_float), ("user_nonnegative", C.c_int)]


class PairParams(C.Structure):
    _fields_ = [("rank_sample_method", C.c_int), ("rank_sample_num", C.c_int), ("rank_sample_max", C.c_int),
                ("rank_sample_pointwise", C.c_int), ("pos_sample_lowerb", C.c_float),
                ("neg_sample_upperb", C.c_float), ("rank_sample_gap", C.c_float), ("seed", C.c_ulonglong)]


# every symbol include/svdgpu.h declares: (restype, argtypes)
SVDGPU_SYMBOLS = {
    "svdgpu_create": (C.c_int, [C.POINTER(_vp), C.POINTER(Shape), C.c_int]),
    "svdgpu_destroy": (None, [_vp]),
    "svdgpu_last_error": (C.c_char_p, [_vp]),
    "svdgpu_set_hparams": (C.c_int, [_vp, C.POINTER(HParams)]),
    "svdgpu_set_mode": (C.c_int, [_vp, C.c_int]),
    "svdgpu_set_option": (C.c_int, [_vp, C.c_char_p, C.c_longlong]),
    "svdgpu_set_stream": (C.c_int, [_vp, _vp]),
    "svdgpu_set_side_features": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "svdgpu_set_wd_ranges": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp]),
    "svdgpu_upload_model": (C.c_int, [_vp, _f32p, _f32p, C.c_size_t, _f32p]),
    "svdgpu_download_model": (C.c_int, [_vp, _f32p, _f32p, C.c_size_t, _f32p]),
    "svdgpu_update_csr": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp]),
    "svdgpu_predict_csr": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp]),
    "svdgpu_update_ugroup": (C.c_int, [_vp, C.c_int] + [_vp] * 9),
    "svdgpu_predict_ugroup": (C.c_int, [_vp, C.c_int] + [_vp] * 10),
    "svdgpu_eval_csr": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, C.c_float, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "svdgpu_eval_ugroup": (C.c_int, [_vp, C.c_int] + [_vp] * 9 + [C.c_float, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "svdgpu_update_buffer_file": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_longlong)]),
    "svdgpu_predict_buffer_file": (C.c_int, [_vp, C.c_char_p, _vp, C.c_longlong, C.POINTER(C.c_longlong)]),
    "svdgpu_eval_buffer_file": (C.c_int, [_vp, C.c_char_p, C.c_float, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "svdgpu_batch_eval": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "svdgpu_batch_sample_pairs": (C.c_int, [_vp, _vp, C.POINTER(PairParams), C.POINTER(_vp)]),
    "svdgpu_batch_download": (C.c_int, [_vp, _vp, C.POINTER(C.c_int), C.POINTER(C.c_longlong)] + [_vp] * 5),
    "svdgpu_batch_create": (C.c_int, [_vp, C.POINTER(_vp), C.c_int, _vp, _vp, _vp, _vp]),
    "svdgpu_batch_set_ugroup": (C.c_int, [_vp, _vp, C.c_int] + [_vp] * 5),
    "svdgpu_batch_update": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "svdgpu_batch_plan": (C.c_int, [_vp, _vp, C.POINTER(C.c_int)]),
    "svdgpu_batch_predict": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp]),
    "svdgpu_batch_destroy": (None, [_vp, _vp]),
    "svdgpu_rank_init": (C.c_int, [_vp, C.c_int, C.c_int]),
    "svdgpu_rank_csr": (C.c_int, [_vp, C.c_int] + [_vp] * 5 + [C.c_longlong, C.POINTER(C.c_longlong)]),
    "svdgpu_rank_ugroup": (C.c_int, [_vp, C.c_int] + [_vp] * 10 + [C.c_longlong, C.POINTER(C.c_longlong)]),
    "svdgpu_sync": (C.c_int, [_vp]),
    "svdgpu_timer_start": (C.c_int, [_vp]),
    "svdgpu_timer_stop": (C.c_int, [_vp, _f32p]),
    "svdgpu_get_counter": (C.c_longlong, [_vp, C.c_char_p]),
    "svdgpu_own_stats": (C.c_int, [_vp, _vp, C.c_int, C.POINTER(C.c_int)]),
    "svdgpu_microbench": (C.c_int, [_vp, C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_double)]),
    "svdgpu_device_ptr": (_vp, [_vp, C.c_int, C.POINTER(C.c_size_t)]),
    "svdgpu_items_snapshot": (C.c_int, [_vp]),
    "svdgpu_items_pack_delta": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_size_t)]),
    "svdgpu_items_apply_delta": (C.c_int, [_vp, C.c_float]),
    "svdgpu_comm_id": (C.c_int, [_vp]),
    "svdgpu_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "svdgpu_comm_destroy": (C.c_int, [_vp]),
    "svdgpu_comm_rank": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "svdgpu_allreduce_items": (C.c_int, [_vp, C.c_float]),
    "svdgpu_allgather_users": (C.c_int, [_vp]),
    "svdgpu_allreduce_items_group": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_float]),
}

_lib = None
_tlib = None


class SvdGpuError(RuntimeError):
    pass


def load_library():
    """dlopen libsvdgpu.so and declare every symbol of include/svdgpu.h."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_GPU):
            raise SvdGpuError(
                "native library %s is missing: run `python -m svdfeature_b200.build` "
                "(there is no CPU fallback)" % LIB_GPU)
        lib = C.CDLL(LIB_GPU, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SVDGPU_SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _ptr(a):
    """Address of a numpy array / torch tensor (host memory) or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    if hasattr(a, "data_ptr"):  # torch tensor (e.g. pinned host memory)
        assert a.is_contiguous()
        return a.data_ptr()
    return a


class Batch:
    def __init__(self, owner, handle, num_row):
        self.owner, self.h, self.num_row = owner, handle, num_row
        self.num_unit = None

    def close(self):
        if self.h:
            self.owner.lib.svdgpu_batch_destroy(self.owner.h, self.h)
            self.h = None


class SvdGpu:
    """One trainer on one GPU: the C ABI, one method per entry point."""

    def __init__(self, num_user, num_item, num_factor, num_global=0, num_ufeedback=0, no_user_bias=0,
                 active_type=0, format_type=0, device=0):
        self.lib = load_library()
        self.shape = Shape(num_user, num_item, num_global, num_ufeedback, num_factor, no_user_bias,
                           active_type, format_type)
        h = _vp()
        if self.lib.svdgpu_create(C.byref(h), C.byref(self.shape), device) != 0:
            raise SvdGpuError(self.lib.svdgpu_last_error(None).decode())
        self.h = h
        self.ustart = num_ufeedback if format_type == 1 else 0
        self.rows = self.ustart + num_user + num_item
        self.pitch = (num_factor + 3) // 4 * 4

    def _ck(self, rc):
        if rc != 0:
            raise SvdGpuError(self.lib.svdgpu_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.svdgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # configuration
    def set_hparams(self, **kw):
        hp = HParams(learning_rate=0.01, scale_lr_ufeedback=1.0)
        for k, v in kw.items():
            setattr(hp, k, v)
        self.hp = hp
        self._ck(self.lib.svdgpu_set_hparams(self.h, C.byref(hp)))

    def set_mode(self, mode):
        self._ck(self.lib.svdgpu_set_mode(self.h, mode))

    def set_option(self, name, value):
        self._ck(self.lib.svdgpu_set_option(self.h, name.encode(), int(value)))

    def set_stream(self, cuda_stream):
        self._ck(self.lib.svdgpu_set_stream(self.h, cuda_stream))

    def set_side_features(self, which, side):
        """side: list of (index array, value array) per feature index (see svdgpu_set_side_features)."""
        rp = np.zeros(len(side) + 1, np.uint32)
        rp[1:] = np.cumsum([len(i) for i, _ in side])
        idx = np.concatenate([i for i, _ in side]).astype(np.uint32) if side else np.zeros(0, np.uint32)
        val = np.concatenate([v for _, v in side]).astype(np.float32) if side else np.zeros(0, np.float32)
        self._ck(self.lib.svdgpu_set_side_features(self.h, which, len(side), _ptr(rp), _ptr(idx), _ptr(val)))

    def set_wd_ranges(self, which, bounds, wds):
        """Ranged weight decay of the user (0) / item (1) / global (2) space: indices below bounds[0] take
        wds[0], those in [bounds[j-1], bounds[j]) take wds[j] (the reference's up:/ip:/gp: keys)."""
        b = np.asarray(bounds, np.uint32)
        w = np.asarray(wds, np.float32)[: len(b)]
        self._ck(self.lib.svdgpu_set_wd_ranges(self.h, which, len(b), _ptr(b), _ptr(w)))

    def set_wd_range_params(self, pairs):
        """pairs: the ordered (key, value) configuration pairs with the prefixes up: / ip: / uip: / gp:."""
        sets = {0: ("up:", "uip:"), 1: ("ip:", "uip:"), 2: ("gp:", "gp:")}
        for which, prefixes in sets.items():
            bounds, wds = [], []
            for k, v in pairs:
                for pre in prefixes:
                    if k.startswith(pre):
                        (bounds if k[len(pre):] == "bound" else wds).append(v)
                        break
            self.set_wd_ranges(which, bounds, wds)

    # model
    def upload(self, ui_bias, W, g_bias):
        assert W.dtype == np.float32 and ui_bias.dtype == np.float32 and g_bias.dtype == np.float32
        W = np.ascontiguousarray(W)
        pitch = W.shape[1] if W.ndim == 2 else self.pitch
        self._ck(self.lib.svdgpu_upload_model(
            self.h, ui_bias.ctypes.data_as(_f32p), W.ctypes.data_as(_f32p), pitch,
            g_bias.ctypes.data_as(_f32p)))

    def download(self):
        ub = np.zeros(max(self.rows, 1), np.float32)
        W = np.zeros((max(self.rows, 1), self.pitch), np.float32)
        gb = np.zeros(max(self.shape.num_global, 1), np.float32)
        self._ck(self.lib.svdgpu_download_model(
            self.h, ub.ctypes.data_as(_f32p), W.ctypes.data_as(_f32p), self.pitch,
            gb.ctypes.data_as(_f32p)))
        return ub[:self.rows], W[:self.rows], gb[:self.shape.num_global]

    # hot path, host buffers
    def update_csr(self, csr):
        rp, lb, ix, vl = csr
        self._ck(self.lib.svdgpu_update_csr(self.h, len(lb), _ptr(rp), _ptr(lb), _ptr(ix), _ptr(vl)))

    def predict_csr(self, csr):
        rp, lb, ix, vl = csr
        out = np.empty(len(lb), np.float32)
        self._ck(self.lib.svdgpu_predict_csr(self.h, len(lb), _ptr(rp), _ptr(lb), _ptr(ix), _ptr(vl),
                                             _ptr(out)))
        return out

    def update_ugroup(self, ug):
        bro, bfo, tag, fi, fv, rp, lb, ix, vl = ug
        self._ck(self.lib.svdgpu_update_ugroup(self.h, len(bro) - 1, _ptr(bro), _ptr(bfo), _ptr(tag),
                                               _ptr(fi), _ptr(fv), _ptr(rp), _ptr(lb), _ptr(ix), _ptr(vl)))

    def predict_ugroup(self, ug):
        bro, bfo, tag, fi, fv, rp, lb, ix, vl = ug
        out = np.empty(len(lb), np.float32)
        self._ck(self.lib.svdgpu_predict_ugroup(self.h, len(bro) - 1, _ptr(bro), _ptr(bfo), _ptr(tag),
                                                _ptr(fi), _ptr(fv), _ptr(rp), _ptr(lb), _ptr(ix),
                                                _ptr(vl), _ptr(out)))
        return out

    # evaluation on the device: (sum of squared errors, rows)
    def eval_csr(self, csr, scale=1.0):
        rp, lb, ix, vl = csr
        s, n = C.c_double(), C.c_longlong()
        self._ck(self.lib.svdgpu_eval_csr(self.h, len(lb), _ptr(rp), _ptr(lb), _ptr(ix), _ptr(vl), scale,
                                          C.byref(s), C.byref(n)))
        return s.value, n.value

    def eval_ugroup(self, ug, scale=1.0):
        bro, bfo, tag, fi, fv, rp, lb, ix, vl = ug
        s, n = C.c_double(), C.c_longlong()
        self._ck(self.lib.svdgpu_eval_ugroup(self.h, len(bro) - 1, _ptr(bro), _ptr(bfo), _ptr(tag), _ptr(fi),
                                             _ptr(fv), _ptr(rp), _ptr(lb), _ptr(ix), _ptr(vl), scale,
                                             C.byref(s), C.byref(n)))
        return s.value, n.value

    # bulk ingest of the reference's binary buffer files
    def update_buffer_file(self, path):
        n = C.c_longlong()
        self._ck(self.lib.svdgpu_update_buffer_file(self.h, str(path).encode(), C.byref(n)))
        return n.value

    def predict_buffer_file(self, path, max_rows):
        out = np.empty(max_rows, np.float32)
        n = C.c_longlong()
        self._ck(self.lib.svdgpu_predict_buffer_file(self.h, str(path).encode(), _ptr(out), max_rows, C.byref(n)))
        return out[:n.value]

    def eval_buffer_file(self, path, scale=1.0):
        s, n = C.c_double(), C.c_longlong()
        self._ck(self.lib.svdgpu_eval_buffer_file(self.h, str(path).encode(), scale, C.byref(s), C.byref(n)))
        return s.value, n.value

    def batch_eval(self, batch, begin=0, end=None, scale=1.0):
        if end is None:
            end = batch.num_unit if batch.num_unit is not None else batch.num_row
        s, n = C.c_double(), C.c_longlong()
        self._ck(self.lib.svdgpu_batch_eval(self.h, batch.h, begin, end, scale, C.byref(s), C.byref(n)))
        return s.value, n.value

    # hot path, resident batches
    def batch_create(self, csr, ugroup=None):
        rp, lb, ix, vl = csr
        b = _vp()
        self._ck(self.lib.svdgpu_batch_create(self.h, C.byref(b), len(lb), _ptr(rp), _ptr(lb), _ptr(ix),
                                              _ptr(vl)))
        batch = Batch(self, b, len(lb))
        if ugroup is not None:
            bro, bfo, tag, fi, fv = ugroup
            self._ck(self.lib.svdgpu_batch_set_ugroup(self.h, b, len(bro) - 1, _ptr(bro), _ptr(bfo),
                                                      _ptr(tag), _ptr(fi), _ptr(fv)))
            batch.num_unit = len(bro) - 1 if tag is None or not np.any(tag) else None
        return batch

    def batch_update(self, batch, begin=0, end=None):
        if end is None:
            end = batch.num_unit if batch.num_unit is not None else batch.num_row
        self._ck(self.lib.svdgpu_batch_update(self.h, batch.h, begin, end))

    def batch_plan(self, batch):
        """(Re)build the ordered mode's owner plan of a resident batch; True if the batch qualifies."""
        ok = C.c_int()
        self._ck(self.lib.svdgpu_batch_plan(self.h, batch.h, C.byref(ok)))
        return bool(ok.value)

    def batch_predict(self, batch, begin=0, end=None, fetch=True):
        if end is None:
            end = batch.num_unit if batch.num_unit is not None else batch.num_row
        n = batch.num_row if batch.num_unit is not None else end - begin
        out = np.empty(n, np.float32) if fetch else None
        self._ck(self.lib.svdgpu_batch_predict(self.h, batch.h, begin, end, _ptr(out)))
        return out

    # pairwise-rank samples on the device
    def batch_sample_pairs(self, batch, seed=0, method=0, num=-1, maxn=-1, pointwise=0, pos_lowerb=0.8,
                           neg_upperb=1e-6, gap=1e-4):
        pp = PairParams(method, num, maxn, pointwise, pos_lowerb, neg_upperb, gap, seed)
        out = _vp()
        self._ck(self.lib.svdgpu_batch_sample_pairs(self.h, batch.h, C.byref(pp), C.byref(out)))
        n, nv = C.c_int(), C.c_longlong()
        self._ck(self.lib.svdgpu_batch_download(self.h, out, C.byref(n), C.byref(nv), None, None, None, None, None))
        b = Batch(self, out, n.value)
        b.num_unit = batch.num_unit
        b.num_val = nv.value
        return b

    def batch_download(self, batch):
        """(blk_row_off or None, row_ptr, label, index, value) of a resident batch."""
        n, nv = C.c_int(), C.c_longlong()
        self._ck(self.lib.svdgpu_batch_download(self.h, batch.h, C.byref(n), C.byref(nv), None, None, None, None, None))
        rp = np.zeros(3 * n.value + 1, np.int32)
        lab = np.zeros(n.value, np.float32)
        idx = np.zeros(nv.value, np.uint32)
        val = np.zeros(nv.value, np.float32)
        bro = np.zeros(batch.num_unit + 1, np.int32) if batch.num_unit is not None else None
        self._ck(self.lib.svdgpu_batch_download(self.h, batch.h, None, None, _ptr(rp), _ptr(lab), _ptr(idx), _ptr(val),
                                                _ptr(bro)))
        return bro, rp, lab, idx, val

    # sync / timing / introspection
    # ranking (SVDFeatureRanker)
    def rank_init(self, num_item_set, top_k=0):
        self._ck(self.lib.svdgpu_rank_init(self.h, int(num_item_set), int(top_k)))

    def rank(self, stream, kind="csr", cap=1 << 20):
        """Feed a tagged ranker stream (CSR arrays, or the user-grouped tuple with kind="ug"); returns the
        results of the sections closed inside it."""
        out = np.empty(cap, np.int32)
        n = C.c_longlong(0)
        if kind == "csr":
            row_ptr, label, index, value = stream
            self._ck(self.lib.svdgpu_rank_csr(self.h, len(label), _ptr(row_ptr), _ptr(label), _ptr(index), _ptr(value),
                                              _ptr(out), cap, C.byref(n)))
        else:
            bro, bfo, tag, fbi, fbv, row_ptr, label, index, value = stream
            self._ck(self.lib.svdgpu_rank_ugroup(self.h, len(bro) - 1, _ptr(bro), _ptr(bfo), _ptr(tag), _ptr(fbi),
                                                 _ptr(fbv), _ptr(row_ptr), _ptr(label), _ptr(index), _ptr(value),
                                                 _ptr(out), cap, C.byref(n)))
        return out[: n.value].copy()

    def sync(self):
        self._ck(self.lib.svdgpu_sync(self.h))

    def timer_start(self):
        self._ck(self.lib.svdgpu_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self._ck(self.lib.svdgpu_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def counter(self, name):
        return int(self.lib.svdgpu_get_counter(self.h, name.encode()))

    def own_stats(self):
        """Per-owner counters of the last item-owner launch (option own_stats=1 before it):
        int64[num_owner, 4] = cycles in the queue loop, waiting for a slot, in publish fences, waits."""
        n = C.c_int()
        self._ck(self.lib.svdgpu_own_stats(self.h, None, 0, C.byref(n)))
        out = np.zeros((n.value, 4), np.int64)
        self._ck(self.lib.svdgpu_own_stats(self.h, out.ctypes.data, n.value, C.byref(n)))
        return out

    def microbench(self, which, nbytes, iters):
        """GB/s of a streaming read (which=0) or copy (1) over `nbytes` of device memory, `iters` passes."""
        out = C.c_double()
        self._ck(self.lib.svdgpu_microbench(self.h, which, nbytes, iters, C.byref(out)))
        return out.value

    def device_ptr(self, which):
        pitch = C.c_size_t()
        p = self.lib.svdgpu_device_ptr(self.h, which, C.byref(pitch))
        return p, pitch.value

    # multi-GPU exchange
    def items_snapshot(self):
        self._ck(self.lib.svdgpu_items_snapshot(self.h))

    def items_pack_delta(self):
        p, n = _vp(), C.c_size_t()
        self._ck(self.lib.svdgpu_items_pack_delta(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    # the exchange inside the library (NCCL)
    def comm_init(self, world, rank, id_bytes):
        """Join the communicator: id_bytes = the 128 bytes of comm_id() made on rank 0."""
        buf = C.create_string_buffer(bytes(id_bytes), 128) if id_bytes is not None else None
        self._ck(self.lib.svdgpu_comm_init(self.h, world, rank, buf))

    def allreduce_items(self, scale=1.0):
        self._ck(self.lib.svdgpu_allreduce_items(self.h, scale))

    def allgather_users(self):
        self._ck(self.lib.svdgpu_allgather_users(self.h))

    def items_apply_delta(self, scale=1.0):
        self._ck(self.lib.svdgpu_items_apply_delta(self.h, scale))


# ---------------------------------------------------------------------------
# the C++ ISVDTrainer implementation behind the svdtr_* shim
# ---------------------------------------------------------------------------
def comm_id():
    """The 128-byte NCCL rendezvous id (call on rank 0, hand the bytes to the other ranks)."""
    lib = load_library()
    buf = C.create_string_buffer(128)
    if lib.svdgpu_comm_id(buf) != 0:
        raise SvdGpuError((lib.svdgpu_last_error(None) or b"").decode())
    return buf.raw


def load_trainer_library():
    global _tlib
    if _tlib is None:
        load_library()
        if not os.path.exists(LIB_TRAINER):
            raise SvdGpuError("native library %s is missing: run `python -m svdfeature_b200.build`" % LIB_TRAINER)
        lib = C.CDLL(LIB_TRAINER)
        vp = _vp
        sig = {
            "svdtr_create": (vp, [C.c_int, C.c_int, C.c_int]),
            "svdtr_destroy": (None, [vp]),
            "svdtr_seed": (None, [C.c_uint]),
            "svdtr_set_param": (None, [vp, C.c_char_p, C.c_char_p]),
            "svdtr_init_model": (None, [vp]),
            "svdtr_init_trainer": (None, [vp]),
            "svdtr_set_round": (None, [vp, C.c_int]),
            "svdtr_finish_round": (None, [vp]),
            "svdtr_save_model": (C.c_int, [vp, C.c_char_p]),
            "svdtr_load_model": (C.c_int, [vp, C.c_char_p]),
            "svdtr_sync": (None, [vp]),
            "svdtr_gpu_handle": (vp, [vp]),
            # ISVDRanker behind the same shim (trainer_cabi.cpp)
            "svdrk_create_from_model": (vp, [C.c_char_p]),
            "svdrk_destroy": (None, [vp]),
            "svdrk_set_param": (None, [vp, C.c_char_p, C.c_char_p]),
            "svdrk_init_ranker": (None, [vp, C.c_int]),
            "svdrk_rank_csr": (C.c_long, [vp, C.c_int] + [vp] * 5 + [C.c_long]),
            "svdrk_rank_ugroup": (C.c_long, [vp, C.c_int] + [vp] * 10 + [C.c_long]),
        }
        for base in ("svdtr_update_csr", "svdtr_update_csr_bulk"):
            sig[base] = (None, [vp, C.c_int] + [vp] * 4)
        for base in ("svdtr_predict_csr", "svdtr_predict_csr_bulk"):
            sig[base] = (None, [vp, C.c_int] + [vp] * 5)
        for base in ("svdtr_update_ugroup", "svdtr_update_ugroup_bulk"):
            sig[base] = (None, [vp, C.c_int] + [vp] * 9)
        for base in ("svdtr_predict_ugroup", "svdtr_predict_ugroup_bulk"):
            sig[base] = (None, [vp, C.c_int] + [vp] * 10)
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _tlib = lib
    return _tlib


class GpuTrainer:
    """ISVDTrainer on the GPU.  ``bulk=False`` drives the per-row virtuals
    (``update(Elem)`` / ``update(SVDPlusBlock)``) exactly like svd_feature.cpp:231-247;
    ``bulk=True`` hands whole batches over in one call."""

    def __init__(self, format_type=2, active_type=0, extend_type=0, params=None, bulk=True):
        self.lib = load_trainer_library()
        self.h = self.lib.svdtr_create(format_type, active_type, extend_type)
        self.sfx = "_bulk" if bulk else ""
        if params:
            self.set_params(params)

    def set_params(self, params):
        for k, v in params.items():
            if k == "wd_ranges":  # ordered (key, value) pairs: up:wd / up:bound / ip:... / gp:...
                for kk, vv in v:
                    self.lib.svdtr_set_param(self.h, str(kk).encode(), str(vv).encode())
                continue
            self.lib.svdtr_set_param(self.h, str(k).encode(), str(v).encode())

    def init(self, seed=10):
        self.lib.svdtr_seed(seed)
        self.lib.svdtr_init_model(self.h)
        self.lib.svdtr_init_trainer(self.h)

    def set_round(self, r):
        self.lib.svdtr_set_round(self.h, int(r))

    def finish_round(self):
        self.lib.svdtr_finish_round(self.h)

    def update_csr(self, csr):
        rp, lb, ix, vl = csr
        getattr(self.lib, "svdtr_update_csr" + self.sfx)(self.h, len(lb), _ptr(rp), _ptr(lb), _ptr(ix), _ptr(vl))

    def predict_csr(self, csr):
        rp, lb, ix, vl = csr
        out = np.empty(len(lb), np.float32)
        getattr(self.lib, "svdtr_predict_csr" + self.sfx)(self.h, len(lb), _ptr(rp), _ptr(lb), _ptr(ix),
                                                          _ptr(vl), _ptr(out))
        return out

    def update_ugroup(self, ug):
        bro, bfo, tag, fi, fv, rp, lb, ix, vl = ug
        getattr(self.lib, "svdtr_update_ugroup" + self.sfx)(
            self.h, len(bro) - 1, _ptr(bro), _ptr(bfo), _ptr(tag), _ptr(fi), _ptr(fv), _ptr(rp),
            _ptr(lb), _ptr(ix), _ptr(vl))

    def predict_ugroup(self, ug):
        bro, bfo, tag, fi, fv, rp, lb, ix, vl = ug
        out = np.empty(len(lb), np.float32)
        getattr(self.lib, "svdtr_predict_ugroup" + self.sfx)(
            self.h, len(bro) - 1, _ptr(bro), _ptr(bfo), _ptr(tag), _ptr(fi), _ptr(fv), _ptr(rp),
            _ptr(lb), _ptr(ix), _ptr(vl), _ptr(out))
        return out

    def save_model(self, path):
        assert self.lib.svdtr_save_model(self.h, path.encode()) == 0

    def load_model(self, path):
        assert self.lib.svdtr_load_model(self.h, path.encode()) == 0

    def model_bytes(self, tmpdir):
        path = os.path.join(str(tmpdir), "m_gpu_%d.model" % id(self))
        self.save_model(path)
        with open(path, "rb") as f:
            return f.read()

    def close(self):
        if getattr(self, "h", None):
            self.lib.svdtr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GpuRanker:
    """ISVDRanker on the GPU (GpuSVDRanker), driven row by row like svd_feature_infer.cpp:349-371:
    create_svd_ranker(type of the model file) -> load_model -> set_param -> init_ranker -> process."""

    def __init__(self, model_path, num_item_set, params=None):
        self.lib = load_trainer_library()
        self.h = self.lib.svdrk_create_from_model(model_path.encode())
        if not self.h:
            raise SvdGpuError("can not open model file %s" % model_path)
        for k, v in (params or {}).items():
            self.lib.svdrk_set_param(self.h, str(k).encode(), str(v).encode())
        self.lib.svdrk_init_ranker(self.h, int(num_item_set))

    def rank(self, stream, kind="csr", cap=1 << 20):
        out = np.empty(cap, np.int32)
        if kind == "csr":
            row_ptr, label, index, value = stream
            n = self.lib.svdrk_rank_csr(self.h, len(label), _ptr(row_ptr), _ptr(label), _ptr(index), _ptr(value),
                                        _ptr(out), cap)
        else:
            bro, bfo, tag, fbi, fbv, row_ptr, label, index, value = stream
            n = self.lib.svdrk_rank_ugroup(self.h, len(bro) - 1, _ptr(bro), _ptr(bfo), _ptr(tag), _ptr(fbi), _ptr(fbv),
                                           _ptr(row_ptr), _ptr(label), _ptr(index), _ptr(value), _ptr(out), cap)
        return out[:n].copy()

    def close(self):
        if self.h:
            self.lib.svdrk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
