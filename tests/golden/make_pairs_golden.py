"""Golden vectors of the reference's PairwiseRankGenerator from the
UNMODIFIED reference, driven through oracle/ref_pairs_shim.cpp.

Run in the build container only (needs /root/reference and oracle/_ref):
    python tests/golden/make_pairs_golden.py

Writes tests/golden/pairs_ref.npz:
  general_*   rated user blocks (labels 1 / 0 / 0.5, optional global features, zero-valued user
              features) and, for every parameter set in PARAMS, the number of rows the reference's
              generator emits per block (`cnt_<name>`) -- the rand()-driven choice of WHICH rows pair
              cannot be reproduced on a GPU, how many pairs a block yields can
  duo_*       blocks of exactly one positive and one negative row (several features per segment,
              shared indices, zero-valued user features): here the generator's output does not
              depend on rand() at all, so the emitted rows (`duo_out_<name>_*`) pin genpair / merge /
              the label rule / the pointwise variant bit for bit
"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from _oracle import build_oracle  # noqa: E402
from svdfeature_b200 import buffer_io, synth  # noqa: E402

# name -> (reference parameter pairs, device sampler keywords)
PARAMS = {
    "default": ({}, {}),
    "num5": ({"rank_sample_num": 5}, {"num": 5}),
    "num5_max3": ({"rank_sample_num": 5, "rank_sample_max": 3}, {"num": 5, "maxn": 3}),
    "pointwise": ({"rank_sample_num": 2, "rank_sample_pointwise": 1}, {"num": 2, "pointwise": 1}),
    "bounds": ({"pos_sample_lowerb": 0.4, "neg_sample_upperb": 0.2}, {"pos_lowerb": 0.4, "neg_upperb": 0.2}),
    "cmp": ({"rank_sample_method": 1}, {"method": 1}),
    "cmp_gap": ({"rank_sample_method": 1, "rank_sample_gap": 0.6}, {"method": 1, "gap": 0.6}),
    "cmp_pointwise": ({"rank_sample_method": 1, "rank_sample_pointwise": 1}, {"method": 1, "pointwise": 1}),
}


def general_blocks(n_user, n_item, seed):
    rng = np.random.default_rng(seed)
    rows, bro = [], [0]
    for u in range(n_user):
        nr = int(rng.integers(0, 12))
        for it in rng.permutation(n_item)[:nr]:
            lab = float(rng.choice([1.0, 0.0, 0.5], p=[0.3, 0.5, 0.2]))
            g = [(int(x), float(np.round(rng.uniform(-1, 1), 2))) for x in np.sort(rng.permutation(6)[:rng.integers(0, 3)])]
            uf = [(u, 1.0)] + ([(int((u + 1) % n_user), 0.0)] if rng.random() < 0.2 else [])
            rows.append((lab, g, uf, [(int(it), 1.0)]))
        bro.append(len(rows))
    return _ug(rows, bro)


def duo_blocks(n_block, seed):
    """one positive + one negative row per block; rich, overlapping feature lists"""
    rng = np.random.default_rng(seed)
    rows, bro = [], [0]

    def seg(hi, nmax):
        idx = np.sort(rng.permutation(hi)[:int(rng.integers(0, nmax + 1))])
        return [(int(i), float(np.float32(np.round(rng.uniform(-2, 2), 3)))) for i in idx]

    for b in range(n_block):
        for lab in ((1.0, 0.0) if b % 2 == 0 else (0.0, 1.0)):  # either order inside the block
            uf = [(b % 50, 1.0)] + ([((b + 3) % 50 + 50, 0.0)] if rng.random() < 0.4 else []) + \
                 ([((b + 7) % 50 + 100, float(np.float32(rng.uniform(-1, 1))))] if rng.random() < 0.4 else [])
            rows.append((lab, seg(8, 4), uf, seg(12, 5) or [(int(rng.integers(0, 12)), 1.0)]))
        bro.append(len(rows))
    return _ug(rows, bro)


def _ug(rows, bro):
    nb = len(bro) - 1
    return (np.asarray(bro, np.int32), np.zeros(nb + 1, np.int32), np.zeros(nb, np.int32), np.zeros(0, np.uint32),
            np.zeros(0, np.float32)) + synth.ragged_csr(rows)


def reference_pairs(lib, ug, params, seed=10):
    """(blk_row_off, row_ptr, label, index, value) of what the reference's generator emits"""
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "blocks.buffer")
        buffer_io.write_ugroup_buffer(path, ug)
        keys = [k.encode() for k in params]
        vals = [str(v).encode() for v in params.values()]
        h = lib.refpairs_open(path.encode(), len(keys), (C.c_char_p * len(keys))(*keys), (C.c_char_p * len(vals))(*vals), seed)
        cap_r, cap_v = 1 << 14, 1 << 18
        rp = np.zeros(3 * cap_r + 1, np.int32)
        lab, idx, val = np.zeros(cap_r, np.float32), np.zeros(cap_v, np.uint32), np.zeros(cap_v, np.float32)
        nr, nv = C.c_int(), C.c_int()
        bro, rps, labs, idxs, vals_, off = [0], [np.zeros(1, np.int64)], [], [], [], 0
        while True:
            rc = lib.refpairs_next(h, cap_r, cap_v, C.byref(nr), C.byref(nv), rp.ctypes.data, lab.ctypes.data, idx.ctypes.data,
                                   val.ctypes.data)
            assert rc >= 0
            if rc == 0:
                break
            n, m = nr.value, nv.value
            rps.append(rp[1:3 * n + 1].astype(np.int64) + off)
            off += m
            labs.append(lab[:n].copy())
            idxs.append(idx[:m].copy())
            vals_.append(val[:m].copy())
            bro.append(bro[-1] + n)
        lib.refpairs_close(h)
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    return np.asarray(bro, np.int32), cat(rps, np.int32), cat(labs, np.float32), cat(idxs, np.uint32), cat(vals_, np.float32)


def main():
    build_oracle()
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsvdf_refpairs.so"))
    lib.refpairs_open.restype = C.c_void_p
    lib.refpairs_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_uint]
    lib.refpairs_next.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)] + [C.c_void_p] * 4
    lib.refpairs_close.argtypes = [C.c_void_p]
    out = {}
    names = ("bro", "bfo", "tag", "fi", "fv", "rp", "lab", "idx", "val")
    gen, duo = general_blocks(150, 200, 5), duo_blocks(120, 6)
    for pre, ug in (("general", gen), ("duo", duo)):
        for n, a in zip(names, ug):
            out["%s_%s" % (pre, n)] = a
    for name, (ref_kv, _) in PARAMS.items():
        bro = reference_pairs(lib, gen, ref_kv)[0]
        assert len(bro) == len(gen[0])
        out["cnt_" + name] = np.diff(bro).astype(np.int32)
        res = reference_pairs(lib, duo, ref_kv)
        for n, a in zip(("bro", "rp", "lab", "idx", "val"), res):
            out["duo_out_%s_%s" % (name, n)] = a
    np.savez_compressed(os.path.join(HERE, "pairs_ref.npz"), **out)
    print({k: int(v.sum()) for k, v in out.items() if k.startswith("cnt_")})


if __name__ == "__main__":
    main()
