"""GPU: pairwise-rank sample generation on the device against PairwiseRankGenerator
(apex_svd_data.cpp:812-1025).  The reference's rand() shuffles cannot be reproduced on a GPU, so
parity has two parts:
  * PINNED to the compiled reference (tests/golden/pairs_ref.npz, recorded from the unmodified
    generator by tests/golden/make_pairs_golden.py): the number of rows every block yields under
    every parameter set, and -- on blocks of one positive and one negative row, where rand() has no
    say -- the emitted rows bit for bit (genpair, merge, dropped zero-valued user features, the label
    rule, the pointwise variant, both sampling methods);
  * structural, against the same semantics restated in numpy below: which rows may pair, that every
    negative is used once per cycle, the merged feature lists of every emitted pair."""
import numpy as np
import pytest

from _oracle import COracle
import _cases
from svdfeature_b200 import synth

pytestmark = pytest.mark.gpu


def _blocks(n_user, n_item, seed, with_globals=False):
    """user blocks of rated rows: label 1 (positive) / 0 (negative) / 0.5 (neither)"""
    rng = np.random.default_rng(seed)
    rows, bro = [], [0]
    for u in range(n_user):
        nr = int(rng.integers(0, 12))
        items = rng.permutation(n_item)[:nr]
        for it in items:
            lab = float(rng.choice([1.0, 0.0, 0.5], p=[0.3, 0.5, 0.2]))
            g = [(int(x), float(np.round(rng.uniform(-1, 1), 2))) for x in np.sort(rng.permutation(6)[:rng.integers(0, 3)])] if with_globals else []
            uf = [(u, 1.0)] + ([(int((u + 1) % n_user), 0.0)] if rng.random() < 0.2 else [])  # a zero-valued user feature is dropped
            rows.append((lab, g, uf, [(int(it), 1.0)]))
        bro.append(len(rows))
    csr = synth.ragged_csr(rows)
    nb = len(bro) - 1
    ug = (np.asarray(bro, np.int32), np.zeros(nb + 1, np.int32), np.zeros(nb, np.int32), np.zeros(0, np.uint32),
          np.zeros(0, np.float32))
    return csr, ug


def _merge(a, b):
    """apex_svd_data.cpp:828-860 on two index-sorted (idx, val) lists"""
    out, i, j = [], 0, 0
    while i < len(a) and j < len(b):
        if a[i][0] < b[j][0]:
            out.append(a[i]); i += 1
        elif b[j][0] < a[i][0]:
            out.append((b[j][0], -b[j][1])); j += 1
        else:
            out.append((a[i][0], np.float32(a[i][1]) - np.float32(b[j][1]))); i += 1; j += 1
    out += a[i:]
    out += [(x, -v) for x, v in b[j:]]
    return out


def _rows_of(csr):
    rp, lab, idx, val = csr
    out = []
    for r in range(len(lab)):
        seg = [list(zip(idx[rp[3 * r + s]:rp[3 * r + s + 1]].tolist(), val[rp[3 * r + s]:rp[3 * r + s + 1]].tolist())) for s in range(3)]
        out.append((float(lab[r]), seg[0], seg[1], seg[2]))
    return out


@pytest.mark.parametrize("with_globals", [False, True])
def test_pairs_have_the_reference_structure(native, with_globals):
    nu, ni = 300, 500
    csr, ug = _blocks(nu, ni, 5, with_globals)
    g = native.SvdGpu(nu, ni, 16, num_global=6, num_ufeedback=1, no_user_bias=1, active_type=3, format_type=1)
    g.set_hparams(learning_rate=0.01, base_score=0.0)
    g.set_mode(native.MODE_HOGWILD)
    src = g.batch_create(csr, ugroup=ug)
    pairs = g.batch_sample_pairs(src, seed=7)
    bro, rp, lab, idx, val = g.batch_download(pairs)
    src_rows = _rows_of(csr)
    out_rows = _rows_of((rp, lab, idx, val))
    assert len(bro) == len(ug[0]) and bro[0] == 0 and bro[-1] == len(out_rows)
    for b in range(len(ug[0]) - 1):
        rows = src_rows[ug[0][b]:ug[0][b + 1]]
        pos = [r for r in rows if r[0] - 0.8 > -1e-6]
        neg = [r for r in rows if r[0] - 1e-6 < 1e-6]
        mine = out_rows[bro[b]:bro[b + 1]]
        assert len(mine) == (len(neg) if pos and neg else 0)  # sample_posneg: one pair per negative row
        used_neg = []
        for lab_o, og, ou, oi in mine:
            assert lab_o == 1.0
            # the pair is genpair(p, n) for SOME positive p and negative n of this block
            cands = [(p, n) for p in pos for n in neg
                     if [(x, np.float32(v)) for x, v in _merge(p[3], n[3])] == [(x, np.float32(v)) for x, v in oi]
                     and [(x, np.float32(v)) for x, v in _merge(p[1], n[1])] == [(x, np.float32(v)) for x, v in og]]
            assert cands, (b, oi)
            assert ou == [f for f in cands[0][0][2] if abs(f[1]) > 1e-6]  # zero-valued user features dropped
            used_neg.append(tuple(cands[0][1][3]))
        assert len(set(used_neg)) == len(used_neg)  # a cycle uses every negative exactly once
    # another seed gives other pairs; the same seed the same pairs
    again = g.batch_download(g.batch_sample_pairs(src, seed=7))
    other = g.batch_download(g.batch_sample_pairs(src, seed=8))
    assert all(np.array_equal(a, b) for a, b in zip(again, (bro, rp, lab, idx, val)))
    assert not np.array_equal(other[3], idx)


def test_pair_count_options_and_pointwise(native):
    nu, ni = 120, 300
    csr, ug = _blocks(nu, ni, 9)
    g = native.SvdGpu(nu, ni, 8, num_ufeedback=1, no_user_bias=1, active_type=3, format_type=1)
    g.set_hparams(learning_rate=0.01, base_score=0.0)
    g.set_mode(native.MODE_HOGWILD)
    src = g.batch_create(csr, ugroup=ug)
    labs = csr[1]
    has = []
    for b in range(nu):
        l = labs[ug[0][b]:ug[0][b + 1]]
        has.append(bool((l > 0.8 - 1e-6).any() and (l < 2e-6).any()))
    bro = g.batch_download(g.batch_sample_pairs(src, seed=1, num=5))[0]
    assert np.array_equal(np.diff(bro), np.where(has, 5, 0))  # rank_sample_num
    bro = g.batch_download(g.batch_sample_pairs(src, seed=1, num=5, maxn=3))[0]
    assert np.array_equal(np.diff(bro), np.where(has, 3, 0))  # rank_sample_max
    bro, rp, lab, idx, val = g.batch_download(g.batch_sample_pairs(src, seed=1, num=2, pointwise=1))
    assert np.array_equal(np.diff(bro), np.where(has, 4, 0))
    assert np.array_equal(lab, np.tile([1.0, 0.0], len(lab) // 2).astype(np.float32))  # (p, 1), (n, 0)
    bro, rp, lab, idx, val = g.batch_download(g.batch_sample_pairs(src, seed=1, method=10))
    assert np.all(lab == 1.0)  # p.label - n.label = 1 - 0
    with pytest.raises(native.SvdGpuError, match="unkown rank sample method"):  # apex_svd_data.cpp:1010
        g.batch_sample_pairs(src, method=2)
    with pytest.raises(native.SvdGpuError, match="rank_sample_gap"):  # apex_svd_data.cpp:994
        g.batch_sample_pairs(src, method=1, gap=0.0)


@pytest.mark.parametrize("with_globals", [False, True])
def test_label_gap_pairs_have_the_reference_structure(native, with_globals):
    """rank_sample_method = 1 (sample_cmp, apex_svd_data.cpp:920-944): every row of a block is paired with
    one row of the same block whose label differs by more than rank_sample_gap; the higher-labelled
    row is the positive.  A block yields one pair per row that has such a partner."""
    nu, ni = 300, 500
    csr, ug = _blocks(nu, ni, 11, with_globals)
    g = native.SvdGpu(nu, ni, 16, num_global=6, num_ufeedback=1, no_user_bias=1, active_type=3, format_type=1)
    g.set_hparams(learning_rate=0.01, base_score=0.0)
    g.set_mode(native.MODE_HOGWILD)
    src = g.batch_create(csr, ugroup=ug)
    src_rows = _rows_of(csr)
    f32 = lambda seg: [(x, np.float32(v)) for x, v in seg]
    for method, gap in ((1, 1e-4), (11, 0.6)):
        bro, rp, lab, idx, val = g.batch_download(g.batch_sample_pairs(src, seed=3, method=method, gap=gap))
        out_rows = _rows_of((rp, lab, idx, val))
        assert bro[0] == 0 and bro[-1] == len(out_rows)
        seen_partner_sets = 0
        for b in range(len(ug[0]) - 1):
            rows = src_rows[ug[0][b]:ug[0][b + 1]]
            labs = np.array([r[0] for r in rows], np.float32)
            lo = labs - np.float32(gap)
            hi = lo + np.float32(gap) * np.float32(2)
            n_anchor = sum(int(((labs < lo[i]) | (labs >= hi[i])).any()) for i in range(len(rows)))
            mine = out_rows[bro[b]:bro[b + 1]]
            assert len(mine) == n_anchor, (b, len(mine), n_anchor)
            for lab_o, og, ou, oi in mine:
                cands = [(p, n) for p in rows for n in rows
                         if p[0] - n[0] > gap * 0.999 and f32(_merge(p[3], n[3])) == f32(oi) and f32(_merge(p[1], n[1])) == f32(og)]
                assert cands, (b, oi)
                assert ou == [f for f in cands[0][0][2] if abs(f[1]) > 1e-6]
                want = {np.float32(1.0)} if method == 1 else {np.float32(p[0]) - np.float32(n[0]) for p, n in cands}
                assert np.float32(lab_o) in want
                seen_partner_sets += 1
        assert seen_partner_sets > 100
    a = g.batch_download(g.batch_sample_pairs(src, seed=3, method=1))
    b_ = g.batch_download(g.batch_sample_pairs(src, seed=3, method=1))
    c = g.batch_download(g.batch_sample_pairs(src, seed=4, method=1))
    assert all(np.array_equal(x, y) for x, y in zip(a, b_)) and not np.array_equal(a[3], c[3])
    pw = g.batch_download(g.batch_sample_pairs(src, seed=3, method=1, pointwise=1))
    assert len(pw[2]) == 2 * len(a[2]) and np.array_equal(pw[2], np.tile([1.0, 0.0], len(a[2])).astype(np.float32))


def test_training_on_device_pairs_learns_the_ranking(native):
    """End to end on the device: rated blocks -> pairs -> Hogwild BPR steps; the fraction of
    correctly ordered (positive, negative) pairs rises well above chance."""
    nu, ni, k = 4000, 800, 16
    rng = np.random.default_rng(3)
    taste = rng.standard_normal((nu, 4)).astype(np.float32)
    attr = rng.standard_normal((ni, 4)).astype(np.float32)
    rows, bro = [], [0]
    for u in range(nu):
        items = rng.permutation(ni)[:20]
        score = attr[items] @ taste[u]
        for it, s in zip(items, score):
            rows.append((1.0 if s > 0 else 0.0, [], [(u, 1.0)], [(int(it), 1.0)]))
        bro.append(len(rows))
    csr = synth.ragged_csr(rows)
    ug = (np.asarray(bro, np.int32), np.zeros(nu + 1, np.int32), np.zeros(nu, np.int32), np.zeros(0, np.uint32),
          np.zeros(0, np.float32))
    params = dict(num_user=nu, num_item=ni, num_factor=k, num_ufeedback=1, no_user_bias=1, learning_rate=0.05,
                  wd_user=0.001, wd_item=0.001, base_score=0.5, u_init_sigma=0.1, i_init_sigma=0.1)
    o = COracle(1, 3, 0, params)
    o.init(10)
    g = native.SvdGpu(**_cases.shape_of(params, 1, 3))
    g.set_hparams(**_cases.hparams_of(params, o.base_score))
    g.set_mode(native.MODE_HOGWILD)
    g.upload(*[a.copy() for a in o.arrays()])
    src = g.batch_create(csr, ugroup=ug)

    def auc():
        pairs = g.batch_sample_pairs(src, seed=999)
        p = g.batch_predict(pairs)  # score of (positive - negative): > 0 when ordered correctly
        pairs.close()
        return float((p > 0).mean())

    before = auc()
    for epoch in range(30):
        pairs = g.batch_sample_pairs(src, seed=epoch)
        g.batch_update(pairs)
        g.sync()
        pairs.close()
    after = auc()
    assert abs(before - 0.5) < 0.05 and after > 0.75, (before, after)


PAIR_PARAMS = {  # name -> device sampler keywords (the reference's parameters: tests/golden/make_pairs_golden.py)
    "default": {}, "num5": {"num": 5}, "num5_max3": {"num": 5, "maxn": 3}, "pointwise": {"num": 2, "pointwise": 1},
    "bounds": {"pos_lowerb": 0.4, "neg_upperb": 0.2}, "cmp": {"method": 1}, "cmp_gap": {"method": 1, "gap": 0.6},
    "cmp_pointwise": {"method": 1, "pointwise": 1},
}


def _golden_ug(z, pre):
    return tuple(z["%s_%s" % (pre, n)] for n in ("bro", "bfo", "tag", "fi", "fv", "rp", "lab", "idx", "val"))


@pytest.mark.parametrize("name", sorted(PAIR_PARAMS))
def test_device_sampler_is_pinned_to_the_compiled_reference(native, name):
    import os

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pairs_ref.npz"))
    g = native.SvdGpu(200, 200, 8, num_global=8, num_ufeedback=1, no_user_bias=1, active_type=3, format_type=1)
    g.set_hparams(learning_rate=0.01, base_score=0.0)
    g.set_mode(native.MODE_HOGWILD)
    kw = PAIR_PARAMS[name]
    # (1) rows per block on the general blocks: exactly the reference's
    gen = _golden_ug(z, "general")
    src = g.batch_create(gen[5:], ugroup=gen[:5])
    for seed in (1, 2):
        bro = g.batch_download(g.batch_sample_pairs(src, seed=seed, **kw))[0]
        assert np.array_equal(np.diff(bro), z["cnt_" + name]), name
    assert z["cnt_" + name].sum() > 250
    src.close()
    # (2) one positive + one negative per block: the reference's rows, bit for bit
    duo = _golden_ug(z, "duo")
    src = g.batch_create(duo[5:], ugroup=duo[:5])
    bro, rp, lab, idx, val = g.batch_download(g.batch_sample_pairs(src, seed=3, **kw))
    want = [z["duo_out_%s_%s" % (name, n)] for n in ("bro", "rp", "lab", "idx", "val")]
    assert np.array_equal(bro, want[0])
    assert np.array_equal(rp, want[1])
    assert np.array_equal(lab, want[2])
    assert np.array_equal(idx, want[3])
    assert np.array_equal(val.view(np.uint32), want[4].view(np.uint32))  # the merged values, to the bit
    assert len(lab) >= 120
    src.close()
    g.close()
