"""Seeded synthetic SVDFeatureCSR batches shaped like BASELINE.json's configs.

All generators return the reference's CSR batch layout (apex_svd_data.h:34-231):
``row_ptr`` int32[3n+1] (global|user|item segment bounds per row), ``label``
float32[n], ``index`` uint32[nnz], ``value`` float32[nnz].  User-grouped
generators additionally return ``blk_row_off``, ``blk_fb_off``, ``blk_tag``,
``fb_index``, ``fb_value`` (SVDPlusBlock, apex_svd_data.h:376-466).

Distributions: user activity log-normal, item popularity
Zipf-Mandelbrot p(r) ~ 1/(r+q)^s (q shifts the head so that the hottest item
holds ~0.25% of the ratings, as in Netflix; q=0 is plain Zipf), labels
clip(round(3.6 + N(0,1.1)), 1, 5).
"""
import numpy as np


def zipf_cdf(n_item, s=1.0, q=70.0):
    w = 1.0 / np.power(np.arange(1, n_item + 1, dtype=np.float64) + q, s)
    c = np.cumsum(w)
    return c / c[-1]


def lognormal_cdf(n_user, sigma=1.0, rng=None):
    rng = rng or np.random.default_rng(0)
    w = rng.lognormal(0.0, sigma, n_user)
    c = np.cumsum(w)
    return c / c[-1]


def fixed_csr(label, gidx=None, gval=None, uidx=None, uval=None, iidx=None, ival=None):
    """Pack rows with a FIXED number of global/user/item features into CSR arrays."""
    n = len(label)

    def norm(a, dt):
        if a is None:
            return np.zeros((n, 0), dt)
        a = np.asarray(a, dt)
        return a.reshape(n, -1)

    gidx, uidx, iidx = norm(gidx, np.uint32), norm(uidx, np.uint32), norm(iidx, np.uint32)
    gval, uval, ival = norm(gval, np.float32), norm(uval, np.float32), norm(ival, np.float32)
    ng, nu, ni = gidx.shape[1], uidx.shape[1], iidx.shape[1]
    per = ng + nu + ni
    base = np.arange(n, dtype=np.int64) * per
    row_ptr = np.empty(3 * n + 1, np.int64)
    row_ptr[0:3 * n:3] = base
    row_ptr[1:3 * n:3] = base + ng
    row_ptr[2:3 * n:3] = base + ng + nu
    row_ptr[3 * n] = n * per
    assert n * per < 2 ** 31, "CSR offsets are int32 (apex_svd_data.h:110-119)"
    index = np.concatenate([gidx, uidx, iidx], axis=1).reshape(-1)
    value = np.concatenate([gval, uval, ival], axis=1).reshape(-1)
    return (row_ptr.astype(np.int32), np.asarray(label, np.float32), np.ascontiguousarray(index),
            np.ascontiguousarray(value))


def ragged_csr(rows):
    """rows: list of (label, [(gid,gval)..], [(uid,uval)..], [(iid,ival)..]) -> CSR arrays."""
    row_ptr, label, index, value = [0], [], [], []
    for lab, g, u, i in rows:
        label.append(lab)
        for seg in (g, u, i):
            for idx, val in seg:
                index.append(idx)
                value.append(val)
            row_ptr.append(len(index))
    return (np.asarray(row_ptr, np.int32), np.asarray(label, np.float32),
            np.asarray(index, np.uint32), np.asarray(value, np.float32))


def ratings(n, num_user, num_item, seed=10, zipf_s=1.0, zipf_q=70.0, user_sigma=1.0):
    """(user, item, label) triples: Netflix-shaped marginals, iid order (= shuffled)."""
    rng = np.random.default_rng(seed)
    ucdf = lognormal_cdf(num_user, user_sigma, rng)
    icdf = zipf_cdf(num_item, zipf_s, zipf_q)
    iperm = rng.permutation(num_item)  # popularity rank -> item id
    u = np.minimum(np.searchsorted(ucdf, rng.random(n)), num_user - 1).astype(np.uint32)
    i = iperm[np.minimum(np.searchsorted(icdf, rng.random(n)), num_item - 1)].astype(np.uint32)
    lab = np.clip(np.rint(3.6 + rng.normal(0.0, 1.1, n)), 1, 5).astype(np.float32)
    return u, i, lab


def planted_mf(n, num_user, num_item, seed=10, rank=8, factor_scale=0.4, bias_scale=0.6, noise=0.5, **kw):
    """basic-MF rows whose labels carry a signal a model can learn: label = 3.6 + b_user + b_item +
    <p_user, q_item> + N(0, noise).  The planted parameters depend on (num_user, num_item, rank) only, not on
    `seed`, so a training stream (seed a) and a held-out stream (seed b) share them.  With the defaults the
    label std is 1.08 and the noise floor of the held-out RMSE is 0.5.  (BASELINE's own labels are noise
    independent of user and item: nothing to learn, held-out RMSE cannot tell a good run from a bad one.)"""
    u, i, _ = ratings(n, num_user, num_item, seed, **kw)
    pr = np.random.default_rng(4242)
    P = pr.standard_normal((num_user, rank)).astype(np.float32) * factor_scale
    Q = pr.standard_normal((num_item, rank)).astype(np.float32) * factor_scale
    bu = pr.standard_normal(num_user).astype(np.float32) * bias_scale
    bi = pr.standard_normal(num_item).astype(np.float32) * bias_scale
    rng = np.random.default_rng(seed + 7)
    lab = (3.6 + bu[u] + bi[i] + np.einsum("nk,nk->n", P[u], Q[i]) + noise * rng.standard_normal(n)).astype(np.float32)
    ones = np.ones(n, np.float32)
    return fixed_csr(lab, uidx=u, uval=ones, iidx=i, ival=ones)


def basic_mf(n, num_user, num_item, seed=10, **kw):
    """configs[0]/[1]: one user feature, one item feature, values 1 (demo/basicMF)."""
    u, i, lab = ratings(n, num_user, num_item, seed, **kw)
    ones = np.ones(n, np.float32)
    return fixed_csr(lab, uidx=u, uval=ones, iidx=i, ival=ones)


def pairwise(n, num_user, num_item, seed=10, **kw):
    """configs[3]: rows as PairwiseRankGenerator emits them (apex_svd_data.cpp:886-915):
    label 1, one user feature, two item features +1 (positive) / -1 (negative),
    merged in index order."""
    u, pos, _ = ratings(n, num_user, num_item, seed, **kw)
    rng = np.random.default_rng(seed + 1)
    neg = rng.integers(0, num_item, n).astype(np.uint32)
    neg = np.where(neg == pos, (neg + 1) % num_item, neg).astype(np.uint32)
    lo = np.minimum(pos, neg)
    hi = np.maximum(pos, neg)
    vlo = np.where(pos < neg, 1.0, -1.0).astype(np.float32)
    iidx = np.stack([lo, hi], 1)
    ival = np.stack([vlo, -vlo], 1)
    return fixed_csr(np.ones(n, np.float32), uidx=u, uval=np.ones(n, np.float32), iidx=iidx, ival=ival)


def neighborhood(n, num_user, num_item, num_global, ng=8, seed=10, **kw):
    """configs[4]: basicMF rows + ng real-valued global (neighbourhood) features
    (demo/neighborhoodModel/ua.base.example)."""
    u, i, lab = ratings(n, num_user, num_item, seed, **kw)
    rng = np.random.default_rng(seed + 2)
    gidx = rng.integers(0, num_global, (n, ng)).astype(np.uint32)
    gidx.sort(axis=1)
    gval = rng.normal(0.0, 0.5, (n, ng)).astype(np.float32)
    ones = np.ones(n, np.float32)
    return fixed_csr(lab, gidx=gidx, gval=gval, uidx=u, uval=ones, iidx=i, ival=ones)


def user_grouped(n, num_user, num_item, avg_fb=20, seed=10, **kw):
    """configs[2]: the same ratings grouped by user (shuffled inside a user and
    across users, tools/svdpp_randorder.cpp:67-73) + an implicit-feedback list per
    user = the items this user rated plus random extras, value 1/sqrt(count)
    (demo/implicitFeedback/mkimplicitfeedbackfeature.py:53)."""
    u, i, lab = ratings(n, num_user, num_item, seed, **kw)
    rng = np.random.default_rng(seed + 3)
    order = np.lexsort((rng.random(n), u))
    u, i, lab = u[order], i[order], lab[order]
    users, start = np.unique(u, return_index=True)
    bounds = np.append(start, n)
    perm = rng.permutation(len(users))
    rows_u, rows_i, rows_l = [], [], []
    blk_row_off, blk_fb_off, fb_index, fb_value = [0], [0], [], []
    for b in perm:
        s, e = bounds[b], bounds[b + 1]
        rows_u.append(u[s:e]); rows_i.append(i[s:e]); rows_l.append(lab[s:e])
        extra = rng.integers(0, num_item, max(0, int(rng.poisson(max(avg_fb - (e - s), 0)))))
        fb = np.unique(np.concatenate([i[s:e], extra.astype(np.uint32)]))
        fb_index.append(fb.astype(np.uint32))
        fb_value.append(np.full(len(fb), 1.0 / np.sqrt(len(fb)), np.float32))
        blk_row_off.append(blk_row_off[-1] + (e - s))
        blk_fb_off.append(blk_fb_off[-1] + len(fb))
    u, i, lab = np.concatenate(rows_u), np.concatenate(rows_i), np.concatenate(rows_l)
    ones = np.ones(n, np.float32)
    csr = fixed_csr(lab, uidx=u, uval=ones, iidx=i, ival=ones)
    return (np.asarray(blk_row_off, np.int32), np.asarray(blk_fb_off, np.int32),
            np.zeros(len(perm), np.int32), np.concatenate(fb_index), np.concatenate(fb_value)) + csr


def random_general(n, num_user, num_item, num_global, seed=10, max_g=3, max_u=2, max_i=3,
                   allow_empty=True, allow_dup=False):
    """Ragged rows with random feature counts/values -- edge-case fodder for parity tests."""
    rng = np.random.default_rng(seed)
    rows = []
    for _ in range(n):
        ng = int(rng.integers(0 if allow_empty else 1, max_g + 1)) if num_global > 0 else 0
        nu = int(rng.integers(0 if allow_empty else 1, max_u + 1))
        ni = int(rng.integers(0 if allow_empty else 1, max_i + 1))

        def pick(cnt, hi):
            if cnt == 0:
                return []
            idx = rng.integers(0, hi, cnt) if allow_dup else rng.choice(hi, cnt, replace=False)
            idx = np.sort(idx)
            val = rng.normal(0.0, 1.0, cnt).astype(np.float32)
            val[rng.random(cnt) < 0.3] = 1.0
            return list(zip(idx.tolist(), val.tolist()))

        rows.append((float(rng.integers(1, 6)), pick(ng, max(num_global, 1)), pick(nu, num_user),
                     pick(ni, num_item)))
    return ragged_csr(rows)


# svdranker_tag (apex_svd.h:115-152): the tag of a ranker input row sits in its label field
RK_ITEM, RK_POS, RK_USER, RK_SPEC, RK_PROCESS, RK_BAN = 0, 1, 2, 3, 4, -1


def rank_stream(num_item_set, num_sections, num_user, num_item, num_global=0, seed=10, max_pos=4, max_ban=3,
                max_spec=2, ugroup=False, num_ufeedback=0, avg_fb=6, split_every=0):
    """A tagged ranker input stream (SVDFeatureRanker, base.h:597-813): ``num_item_set`` ITEM rows (item
    and global features of each candidate), then ``num_sections`` user sections USER / POS* / BAN* /
    SPEC* / PROCESS.  Item indices of POS/BAN/SPEC rows are positions in the item set and sit in the
    user-feature field.  Returns CSR arrays; with ``ugroup`` the SVDPlusBlock arrays in front (block 0
    = the item set, then one block per section carrying that user's feedback list; every
    ``split_every``-th section is cut into START / END blocks)."""
    rng = np.random.default_rng(seed)
    rows = []
    pick = lambda hi, n: [int(x) for x in rng.choice(hi, size=n, replace=False)]
    fval = lambda: float(np.float32(rng.uniform(-1.0, 1.0)))
    for _ in range(num_item_set):
        ni = int(rng.integers(1, 4))
        ng = int(rng.integers(0, 3)) if num_global else 0
        rows.append((RK_ITEM, [(g, fval()) for g in pick(num_global, ng)] if ng else [], [],
                     # (a lone item feature never has value 1: two such rows on one item would tie, and
                     # the order of equal scores is std::sort's business, base.h:773)
                     [(i, 1.0 if j == 0 and ni > 1 else fval()) for j, i in enumerate(pick(num_item, ni))]))
    item_rows = len(rows)
    sec_start, sec_cut = [], []
    for s_ in range(num_sections):
        sec_start.append(len(rows))
        nu = int(rng.integers(1, 3))
        rows.append((RK_USER, [], [(u, 1.0 if j == 0 else fval()) for j, u in enumerate(pick(num_user, nu))], []))
        npos, nban = int(rng.integers(0, max_pos + 1)), int(rng.integers(0, max_ban + 1))
        tagged = pick(num_item_set, min(npos + nban, num_item_set))
        pos, ban = tagged[:npos], tagged[npos:]
        if pos:
            half = len(pos) // 2
            for part in ([pos[:half], pos[half:]] if half else [pos]):  # several indices per row, several rows
                rows.append((RK_POS, [], [(i, 1.0) for i in part], []))
        sec_cut.append(len(rows))
        if ban:
            rows.append((RK_BAN, [], [(i, 1.0) for i in ban], []))
        for _ in range(int(rng.integers(0, max_spec + 1))):
            ng = int(rng.integers(0, 3)) if num_global else 0
            ni = int(rng.integers(0, 3))
            rows.append((RK_SPEC, [(g, fval()) for g in pick(num_global, ng)] if ng else [],
                         [(int(rng.integers(0, num_item_set)), 1.0)], [(i, fval()) for i in pick(num_item, ni)]))
        rows.append((RK_PROCESS, [], [], []))
    csr = ragged_csr(rows)
    if not ugroup:
        return csr
    bro, bfo, tag, fbi, fbv = [0, item_rows], [0, 0], [0], [], []
    ends = sec_start[1:] + [len(rows)]
    for s_ in range(num_sections):
        nfb = int(rng.integers(0, 2 * avg_fb + 1)) if num_ufeedback else 0
        fi = pick(num_ufeedback, min(nfb, num_ufeedback)) if nfb else []
        fv = [float(np.float32(1.0 / np.sqrt(max(len(fi), 1))))] * len(fi)
        pieces = [ends[s_]]
        if split_every and s_ % split_every == 0 and sec_cut[s_] < ends[s_]:
            pieces = [sec_cut[s_], ends[s_]]
        for pi, end in enumerate(pieces):
            bro.append(end)
            fbi += fi
            fbv += fv
            bfo.append(len(fbi))
            tag.append(0 if len(pieces) == 1 else (1 if pi == 0 else 2))
    return (np.asarray(bro, np.int32), np.asarray(bfo, np.int32), np.asarray(tag, np.int32),
            np.asarray(fbi, np.uint32), np.asarray(fbv, np.float32)) + csr
