"""Readers/writers for the reference's on-disk instance formats (host plumbing, numpy).

BINARY_BUFFER (random order; writer apex_svd_data.cpp:131-195, reader :240-248 +
apex_svd_data.h:220-230):
    header  int32 num_batch, batch_size, max_batch_num
    batch   int32 num_row, num_val; int32 row_ptr[3*num_row+1] (first = 0);
            float32 label[num_row]; uint32 index[num_val]; float32 value[num_val]
It already is the SoA the kernels consume: a batch maps 1:1 onto svdgpu_update_csr.

User-group buffer (apex_svd_data.cpp:556-640, apex_svd_data.h:419-450):
    header  int32 num_batch, max_num_ufeedback, max_num_row, max_num_val
    block   int32 num_ufeedback (bit 31 set => next int32 is extend_tag);
            uint32 fb_index[]; float32 fb_value[]; then one CSR batch as above

Text feature line (apex_svd_data.cpp:60-108): ``label ng nu ni gid:gval.. uid:uval.. iid:ival..``
"""
import struct

import numpy as np


def parse_feature_text(path):
    """Random-order text feature file -> CSR arrays."""
    row_ptr, label, index, value = [0], [], [], []
    toks = open(path).read().split()
    p = 0
    while p < len(toks):
        lab, ng, nu, ni = float(toks[p]), int(toks[p + 1]), int(toks[p + 2]), int(toks[p + 3])
        p += 4
        label.append(lab)
        for cnt in (ng, nu, ni):
            for _ in range(cnt):
                i, v = toks[p].split(":")
                index.append(int(i))
                value.append(float(v))
                p += 1
            row_ptr.append(len(index))
    return (np.asarray(row_ptr, np.int32), np.asarray(label, np.float32), np.asarray(index, np.uint32),
            np.asarray(value, np.float32))


def _csr_batch_bytes(row_ptr, label, index, value):
    n = len(label)
    base = int(row_ptr[0])
    nv = int(row_ptr[3 * n]) - base
    return b"".join([
        struct.pack("<2i", n, nv), (row_ptr[:3 * n + 1] - base).astype("<i4").tobytes(),
        label.astype("<f4").tobytes(), index[base:base + nv].astype("<u4").tobytes(),
        value[base:base + nv].astype("<f4").tobytes()])


def _read_csr_batch(buf, off):
    n, nv = struct.unpack_from("<2i", buf, off)
    off += 8
    rp = np.frombuffer(buf, "<i4", 3 * n + 1, off)
    off += 4 * (3 * n + 1)
    lab = np.frombuffer(buf, "<f4", n, off)
    off += 4 * n
    idx = np.frombuffer(buf, "<u4", nv, off)
    off += 4 * nv
    val = np.frombuffer(buf, "<f4", nv, off)
    off += 4 * nv
    return (rp, lab, idx, val), off


def _concat_csr(batches):
    if not batches:
        return (np.zeros(1, np.int32), np.zeros(0, np.float32), np.zeros(0, np.uint32), np.zeros(0, np.float32))
    rps, base = [np.zeros(1, np.int64)], 0
    for rp, _, _, _ in batches:
        rps.append(rp[1:].astype(np.int64) + base)
        base += int(rp[-1])
    assert base < 2 ** 31
    return (np.concatenate(rps).astype(np.int32), np.concatenate([b[1] for b in batches]).astype(np.float32),
            np.concatenate([b[2] for b in batches]).astype(np.uint32),
            np.concatenate([b[3] for b in batches]).astype(np.float32))


def write_feature_buffer(path, csr, batch_size=1000):
    """CSR arrays -> BINARY_BUFFER file (what tools/make_feature_buffer writes)."""
    row_ptr, label, index, value = csr
    n = len(label)
    chunks, max_num, nb = [], 0, 0
    for r0 in range(0, n, batch_size):
        r1 = min(n, r0 + batch_size)
        chunks.append(_csr_batch_bytes(row_ptr[3 * r0:3 * r1 + 1], label[r0:r1], index, value))
        max_num = max(max_num, int(row_ptr[3 * r1]) - int(row_ptr[3 * r0]))
        nb += 1
    with open(path, "wb") as f:
        f.write(struct.pack("<3i", nb, batch_size, max_num))
        for c in chunks:
            f.write(c)


def read_feature_buffer(path):
    """BINARY_BUFFER file -> (CSR arrays of all batches concatenated, header dict)."""
    buf = open(path, "rb").read()
    nb, bs, mx = struct.unpack_from("<3i", buf, 0)
    off, batches = 12, []
    for _ in range(nb):
        b, off = _read_csr_batch(buf, off)
        batches.append(b)
    assert off == len(buf), "trailing bytes in buffer file"
    return _concat_csr(batches), dict(num_batch=nb, batch_size=bs, max_batch_num=mx)


def write_ugroup_buffer(path, ug):
    bro, bfo, tag, fi, fv, rp, lab, idx, val = ug
    nb = len(bro) - 1
    blocks, mfb, mrow, mval = [], 0, 0, 0
    for b in range(nb):
        r0, r1, f0, f1 = int(bro[b]), int(bro[b + 1]), int(bfo[b]), int(bfo[b + 1])
        t = int(tag[b]) if tag is not None else 0
        head = struct.pack("<i", f1 - f0) if t == 0 else struct.pack("<Ii", (f1 - f0) | (1 << 31), t)
        blocks.append(head + fi[f0:f1].astype("<u4").tobytes() + fv[f0:f1].astype("<f4").tobytes() +
                      _csr_batch_bytes(rp[3 * r0:3 * r1 + 1], lab[r0:r1], idx, val))
        mfb, mrow = max(mfb, f1 - f0), max(mrow, r1 - r0)
        mval = max(mval, int(rp[3 * r1]) - int(rp[3 * r0]))
    with open(path, "wb") as f:
        f.write(struct.pack("<4i", nb, mfb, mrow, mval))
        for blk in blocks:
            f.write(blk)


def read_ugroup_buffer(path):
    """User-group buffer -> (blk_row_off, blk_fb_off, blk_tag, fb_index, fb_value, row_ptr, label,
    index, value), header dict."""
    buf = open(path, "rb").read()
    nb, mfb, mrow, mval = struct.unpack_from("<4i", buf, 0)
    off = 16
    bro, bfo, tags, fis, fvs, batches = [0], [0], [], [], [], []
    for _ in range(nb):
        (nfb,) = struct.unpack_from("<i", buf, off)
        off += 4
        t = 0
        if nfb < 0:
            nfb &= 0x7FFFFFFF
            (t,) = struct.unpack_from("<i", buf, off)
            off += 4
        fis.append(np.frombuffer(buf, "<u4", nfb, off))
        off += 4 * nfb
        fvs.append(np.frombuffer(buf, "<f4", nfb, off))
        off += 4 * nfb
        b, off = _read_csr_batch(buf, off)
        batches.append(b)
        tags.append(t)
        bro.append(bro[-1] + len(b[1]))
        bfo.append(bfo[-1] + nfb)
    assert off == len(buf), "trailing bytes in buffer file"
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    return ((np.asarray(bro, np.int32), np.asarray(bfo, np.int32), np.asarray(tags, np.int32),
             cat(fis, np.uint32), cat(fvs, np.float32)) + _concat_csr(batches),
            dict(num_batch=nb, max_num_ufeedback=mfb, max_num_row=mrow, max_num_val=mval))
