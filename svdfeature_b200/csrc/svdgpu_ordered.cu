// svdgpu_ordered.cu -- the kernels that are not the Hogwild stream:
//
//   k_exact   : ordered data-flow training.  One instance per warp, fetched in
//               input order; each feature waits until its row's version counter
//               reaches the feature's ticket, so the result equals the reference's
//               sequential loop (base.h:456-462) while independent instances overlap.
//   k_ugroup  : user-grouped (SVD++) input, one lane group (Hogwild) or one warp
//               (ordered) per user unit: feedback gather, the unit's rows in order,
//               feedback scatter (base.h:523-582).
//   k_delta   : replicated-slab snapshot / delta / apply for the multi-GPU exchange.
#include "svdgpu_internal.h"
#include "svdgpu_fb.cuh"

namespace svdk {

// ---------------------------------------------------------------------------
// k_exact
// ---------------------------------------------------------------------------
constexpr int EX_WARPS = 4;
constexpr int EXOPT_NOFENCE = 1, EXOPT_SPIN = 2, EXOPT_STAGE = 4;  // option "exact_opt" (bit mask)
constexpr int EX_STAGE = 32;  // features of an instance staged in shared memory

__device__ __forceinline__ unsigned *version_of(const DevModel &m, int f, int rp1, int rp2, unsigned id) {
  if (f < rp1) return id < (unsigned)m.num_global ? m.ver_g + id : nullptr;
  if (f < rp2) return id < (unsigned)m.num_user ? m.ver_ui + m.user_off + id : nullptr;
  return id < (unsigned)m.num_item ? m.ver_ui + m.item_off + id : nullptr;
}

template <int LANES, int VEC>
__device__ __forceinline__ void wait_tickets(const Group<LANES, VEC> &g, const DevModel &m, int rp0,
                                             int rp1, int rp2, int rp3, const unsigned *idx,
                                             const unsigned *tk, int opt = 0) {
  for (int f = rp0 + g.gl; f < rp3; f += LANES) {
    const unsigned *ver = version_of(m, f, rp1, rp2, idx[f]);
    if (ver) {
      const unsigned want = tk[f];
      unsigned ns = 20, spins = (opt & EXOPT_SPIN) ? 64u : 0u;
      while (ld_acquire_u32(ver) != want) {
        if (spins) {  // the hand-off of a hot row is the critical path: poll back to back first
          --spins;
          continue;
        }
        __nanosleep(ns);
        if (ns < 200) ns += 20;
      }
    }
  }
  g.gsync();
}
template <int LANES, int VEC>
__device__ __forceinline__ void release_tickets(const Group<LANES, VEC> &g, const DevModel &m,
                                                int rp0, int rp1, int rp2, int rp3,
                                                const unsigned *idx, int opt = 0) {
  // Every lane's row stores must be visible before any version moves.  Default: each lane fences
  // its own stores, then the group meets.  EXOPT_NOFENCE: the group meets (the stores of all
  // lanes then happen-before every lane's release), and red.release.gpu -- cumulative over what
  // happens-before it -- publishes them: one fence less on the hand-off path.
  if (!(opt & EXOPT_NOFENCE)) __threadfence();
  g.gsync();
  for (int f = rp0 + g.gl; f < rp3; f += LANES) {
    unsigned *ver = version_of(m, f, rp1, rp2, idx[f]);
    if (ver) red_release_add_u32(ver, 1u);
  }
}

template <int LANES, int VEC>
__global__ void __launch_bounds__(EX_WARPS * 32)
k_exact(DevModel m, DevHP hp, DevCsr csr, int row_begin, int row_end, unsigned *counter,
        int *err_flag, int opt) {
  __shared__ float dot_s[EX_WARPS][Group<LANES, VEC>::DOT_FLOATS];
  __shared__ unsigned st_idx[EX_WARPS][EX_STAGE];
  __shared__ float st_val[EX_WARPS][EX_STAGE], st_val2[EX_WARPS][EX_STAGE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane >= LANES) return;  // one group per warp
  Group<LANES, VEC> g;
  g.gl = lane;
  g.gmask = (LANES == 32) ? 0xffffffffu : ((1u << LANES) - 1u);
  g.dot_s = dot_s[warp];
  const unsigned *idx = csr.index - csr.val_base;
  const float *val = csr.value - csr.val_base;
  const float *val2 = csr.value2 ? csr.value2 - csr.val_base : nullptr;
  const unsigned *tk = csr.ticket - csr.val_base;
  for (;;) {
    unsigned n = 0;
    if (lane == 0) n = atomicAdd(counter, 1u);
    n = g.bcast(n, 0);
    const long long r = (long long)row_begin + n;
    if (r >= row_end) break;
    const int rp0 = csr.row_ptr[3 * r], rp1 = csr.row_ptr[3 * r + 1];
    const int rp2 = csr.row_ptr[3 * r + 2], rp3 = csr.row_ptr[3 * r + 3];
    // (row_ptr was validated on the host together with the tickets)
    const float lab = csr.label[r];
    // EXOPT_STAGE: the acquire below invalidates L1, so index/value loads issued after it are a
    // full L2 round trip that the row gathers then depend on.  Copy the instance's slice to
    // shared memory while it waits: after the hand-off the row addresses are already on chip.
    const unsigned *pidx = idx;
    const float *pval = val, *pval2 = val2;
    if ((opt & EXOPT_STAGE) && rp3 - rp0 <= EX_STAGE) {
      g.gsync();  // the previous instance of this warp has been read out of the stage
      for (int f = rp0 + g.gl; f < rp3; f += LANES) {
        st_idx[warp][f - rp0] = idx[f];
        st_val[warp][f - rp0] = val[f];
        if (val2) st_val2[warp][f - rp0] = val2[f];
      }
      g.gsync();
      pidx = st_idx[warp] - rp0;
      pval = st_val[warp] - rp0;
      if (val2) pval2 = st_val2[warp] - rp0;
    }
    wait_tickets(g, m, rp0, rp1, rp2, rp3, idx, tk, opt);
    process_instance<LANES, VEC, true, true, false>(g, m, hp, rp0, rp1, rp2, rp3, lab, pidx, pval,
                                                    SCATTER_STORE, SCATTER_STORE, nullptr, err_flag,
                                                    pval2);
    release_tickets(g, m, rp0, rp1, rp2, rp3, pidx, opt);
  }
}

// ---------------------------------------------------------------------------
// k_ugroup
// ---------------------------------------------------------------------------
template <int LANES, int VEC>
__device__ __forceinline__ bool strictly_increasing(const Group<LANES, VEC> &g, const unsigned *idx,
                                                    int n) {
  bool bad = false;
  for (int i = g.gl; i + 1 < n; i += LANES) bad |= !(idx[i] < idx[i + 1]);
  return !__any_sync(g.gmask, bad);
}

// base.h:523-538
template <int LANES, int VEC>
__device__ __forceinline__ bool prepare_ufeedback(const Group<LANES, VEC> &g, const DevModel &m,
                                                  const unsigned *fi, const float *fv, int nfb,
                                                  FbState<VEC> &s, int *err_flag) {
  bool bad = false;
  for (int i = g.gl; i < nfb; i += LANES) bad |= fi[i] >= (unsigned)m.num_ufeedback;
  if (__any_sync(g.gmask, bad)) {
    if (bad) atomicCAS(err_flag, 0, ERR_FB_INDEX);
    return false;
  }
  s.norm = 0.0f;
  s.fb_bias = 0.0f;
#pragma unroll
  for (int v = 0; v < VEC; ++v) s.fb[v] = f4_zero();
  {
#pragma unroll 4
    for (int i = 0; i < nfb; ++i) {
      float4 w[VEC];
      g.load_row(m, (size_t)fi[i], w);
      const float x = fv[i];
      const bool one = scalar_is_one(x);
#pragma unroll
      for (int v = 0; v < VEC; ++v) s.fb[v] = f4_add_scaled(s.fb[v], w[v], x, one);
      s.norm = __fadd_rn(s.norm, __fmul_rn(x, x));
    }
  }
  if (!m.no_user_bias) {
    for (int base = 0; base < nfb; base += LANES) {
      const int i = base + g.gl;
      float p = 0.0f;
      if (i < nfb) p = __fmul_rn(__ldcg(m.bias + fi[i]), fv[i]);
      const int cnt = min(LANES, nfb - base);
      for (int j = 0; j < cnt; ++j) s.fb_bias = __fadd_rn(s.fb_bias, g.bcast(p, j));
    }
  }
  return true;
}

// base.h:539-554
template <int LANES, int VEC>
__device__ __forceinline__ void update_ufeedback(const Group<LANES, VEC> &g, const DevModel &m,
                                                 const unsigned *fi, const float *fv, int nfb,
                                                 FbState<VEC> &s, const float4 (&old)[VEC],
                                                 float old_bias, int scatter) {
  if (nfb == 0) return;
  const float inv = __fdiv_rn(1.0f, s.norm);
  const bool inv_one = scalar_is_one(inv);
  float4 d[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    d[v] = f4_sub(s.fb[v], old[v]);
    if (!inv_one) d[v] = f4_scale(d[v], inv);
  }
  const float dbias = __fmul_rn(__fsub_rn(s.fb_bias, old_bias), inv);
  const bool unique = strictly_increasing(g, fi, nfb);
  const bool use_red = unique && scatter == SCATTER_RED;
#pragma unroll 4
  for (int i = 0; i < nfb; ++i) {
    const float x = fv[i];
    const bool one = scalar_is_one(x);
    float *p = m.W + (size_t)fi[i] * (size_t)m.pitch;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int c = g.gl + v * LANES;
      if (4 * c < m.pitch) {
        if (use_red) {
          red4(p + 4 * c, one ? d[v] : f4_scale(d[v], x));
        } else {
          const float4 w = ldcg4(p + 4 * c);
          stcg4(p + 4 * c, f4_add_scaled(w, d[v], x, one));
        }
      }
    }
  }
  if (!m.no_user_bias) {
    if (unique) {
      for (int i = g.gl; i < nfb; i += LANES) {
        float *p = m.bias + fi[i];
        const float add = __fmul_rn(dbias, fv[i]);
        if (use_red) red1(p, add);
        else __stcg(p, __fadd_rn(__ldcg(p), add));
      }
    } else {
      if (g.gl == 0)
        for (int i = 0; i < nfb; ++i) {
          float *p = m.bias + fi[i];
          __stcg(p, __fadd_rn(__ldcg(p), __fmul_rn(dbias, fv[i])));
        }
      g.gsync();
    }
  }
}

// ORDERED: exact data-flow (one unit per warp, tickets); else Hogwild (one unit per group).
// TRAIN=false: prediction (base.h:583-591), pred_out indexed by row - row_base.
template <int LANES, int VEC, bool EXACT_DOT, bool ORDERED, bool TRAIN>
__global__ void __launch_bounds__(EX_WARPS * 32)
k_ugroup(DevModel m, DevHP hp, DevCsr csr, DevUgroup ug, int unit_begin, int unit_end,
         int scatter_user, int scatter_item, unsigned *counter, float *pred_out, int *err_flag,
         const unsigned char *skip_kind /* units k_svdpp has done (kind 1), or null */) {
  constexpr int GPW = ORDERED ? 1 : 32 / LANES;
  __shared__ float dot_s[EX_WARPS][GPW * Group<LANES, VEC>::DOT_FLOATS];
  // k-vector hand-off between a lane group and the whole warp (cooperative feedback phases)
  __shared__ __align__(16) float xch_s[ORDERED ? 1 : EX_WARPS][ORDERED ? 4 : LANES * VEC * 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (ORDERED && lane >= LANES) return;
  Group<LANES, VEC> g;
  g.gl = lane % LANES;
  const int gw = ORDERED ? 0 : lane / LANES;
  g.gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (gw * LANES));
  g.dot_s = &dot_s[warp][gw * Group<LANES, VEC>::DOT_FLOATS];
  const unsigned *idx = csr.index - csr.val_base;
  const float *val = csr.value - csr.val_base;
  const float *val2 = csr.value2 ? csr.value2 - csr.val_base : nullptr;
  const unsigned *tk = ORDERED ? csr.ticket - csr.val_base : nullptr;
  const int *row_ptr = csr.row_ptr - 3 * (long long)ug.row_base;
  const float *label = csr.label - ug.row_base;
  if (ORDERED) scatter_user = scatter_item = SCATTER_STORE;  // ordered results must not depend on atomics

  // ---- per-unit state of this group ------------------------------------------------------
  FbState<VEC> s;
  float4 old[VEC];
  float old_bias = 0.0f;
  s.norm = s.fb_bias = 0.0f;
#pragma unroll
  for (int v = 0; v < VEC; ++v) s.fb[v] = old[v] = f4_zero();
  int f0 = 0, nf0 = 0, f1 = 0, nf1 = 0, r_cur = 0, r_end = 0;
  // Hogwild fast path: a unit is one user, so the user row (and bias) of consecutive
  // basic-MF-shaped rows is the same -- it stays in registers from the first row that needs
  // it to the last and is written back once (the reference reads and writes it per row,
  // base.h:354-427; nobody else touches it meanwhile, so the result is the same).
  unsigned c_uid = 0xffffffffu;
  float4 c_wu[VEC];
  float c_ub = 0.0f;
#pragma unroll
  for (int v = 0; v < VEC; ++v) c_wu[v] = f4_zero();
  auto flush_user = [&]() {
    if (c_uid != 0xffffffffu && TRAIN) {
      g.store_row(m, (size_t)m.user_off + c_uid, c_wu);
      if (!m.no_user_bias && g.gl == 0) __stcg(m.bias + m.user_off + c_uid, c_ub);
    }
    c_uid = 0xffffffffu;
  };
  // cooperative feedback phases need the row to split evenly over 32 lanes
  const int cpl = ORDERED ? 0 : ((m.pitch & 31) == 0 ? m.pitch >> 5 : 0);
  const bool coop = !ORDERED && (cpl == 1 || cpl == 2 || cpl == 4);
  // take the next unit: false when none is left.  `ok` = its feedback list is usable.
  auto next_unit = [&](bool &ok) -> bool {
    int u;
    for (;;) {
      unsigned n = 0;
      if (g.gl == 0) n = atomicAdd(counter, 1u);
      n = g.bcast(n, 0);
      if ((long long)unit_begin + n >= unit_end) return false;
      u = unit_begin + (int)n;
      if (!ORDERED && ug.order) u = ug.order[u];
      if (!(skip_kind && skip_kind[u - unit_begin])) break;
    }
    const int b0 = ug.unit_off[u], b1 = ug.unit_off[u + 1];
    // feedback list of the first block feeds the gather, of the last block the scatter
    f0 = ug.blk_fb_off[b0] - ug.fb_base;
    nf0 = ug.blk_fb_off[b0 + 1] - ug.blk_fb_off[b0];
    f1 = ug.blk_fb_off[b1 - 1] - ug.fb_base;
    nf1 = ug.blk_fb_off[b1] - ug.blk_fb_off[b1 - 1];
    r_cur = ug.blk_row_off[b0];
    r_end = ug.blk_row_off[b1];
    if (ORDERED) {
      // hold the feedback rows of this unit from gather to scatter
      for (int i = g.gl; i < nf0; i += LANES) {
        const unsigned id = ug.fb_index[f0 + i];
        if (id < (unsigned)m.num_ufeedback) {
          const unsigned want = ug.fb_ticket[f0 + i];
          unsigned ns = 20;
          while (ld_acquire_u32(m.ver_ui + id) != want) {
            __nanosleep(ns);
            if (ns < 200) ns += 20;
          }
        }
      }
      g.gsync();
    }
    if (coop) return true;  // the warp gathers the feedback rows together (below)
    ok = prepare_ufeedback(g, m, ug.fb_index + f0, ug.fb_value + f0, nf0, s, err_flag);
    old_bias = s.fb_bias;
#pragma unroll
    for (int v = 0; v < VEC; ++v) old[v] = s.fb[v];
    return true;
  };
  auto do_row = [&](int r) {
    const int rp0 = row_ptr[3 * (long long)r], rp1 = row_ptr[3 * (long long)r + 1];
    const int rp2 = row_ptr[3 * (long long)r + 2], rp3 = row_ptr[3 * (long long)r + 3];
    if (!ORDERED && !row_ok(rp0, rp1, rp2, rp3, csr.val_base, csr.val_end)) {
      if (g.gl == 0) atomicCAS(err_flag, 0, ERR_ROW_PTR);
      return;
    }
    if (!ORDERED && hp.plain && !val2 && rp1 == rp0 && rp2 == rp1 + 1 && rp3 == rp2 + 1 &&
        idx[rp1] < (unsigned)m.num_user && idx[rp2] < (unsigned)m.num_item) {
      const unsigned uid = idx[rp1], iid = idx[rp2];
      const size_t irow = (size_t)m.item_off + iid;
      float4 wi[VEC];
      g.load_row(m, irow, wi);
      const float ib = __ldcg(m.bias + irow);
      if (uid != c_uid) {
        flush_user();
        c_uid = uid;
        g.load_row(m, (size_t)m.user_off + uid, c_wu);
        c_ub = m.no_user_bias ? 0.0f : __ldcg(m.bias + m.user_off + uid);
      }
      const float uval = val[rp1], ival = val[rp2];
      const float um = scalar_is_one(uval) ? 1.0f : uval, im = scalar_is_one(ival) ? 1.0f : ival;
      // prepare_tmp (base.h:354-381) on top of tmp_ufeedback (prepare_svdpp, :506-508)
      float4 tu[VEC], ti[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        tu[v] = f4_add_scaled(s.fb[v], c_wu[v], um, false);
        ti[v] = f4_add_scaled(f4_zero(), wi[v], im, false);
      }
      // calc_bias (base.h:313-353) with get_bias_svdpp (:509-511), pred (:445-454)
      double bsum = 0.0;
      if (!m.no_user_bias) {
        bsum = __dadd_rn(bsum, (double)__fmul_rn(uval, c_ub));
        bsum = __dadd_rn(bsum, (double)s.fb_bias);
      }
      bsum = __dadd_rn(bsum, (double)__fmul_rn(ival, ib));
      const float d = g.template dot<EXACT_DOT>(m, tu, ti);
      double sum = __dadd_rn((double)hp.base_score, bsum);
      sum = __dadd_rn(sum, (double)d);
      const float p = map_active((float)sum, m.active_type);
      if (!TRAIN) {
        if (g.gl == 0) pred_out[r - ug.row_base] = p;
        return;
      }
      // update_no_decay + regularize(after) (base.h:383-427, 211-283)
      const float err = cal_grad(label[r], p, m.active_type);
      const float lrerr = __fmul_rn(hp.lr, err);
      const float su = __fmul_rn(lrerr, uval), si = __fmul_rn(lrerr, ival);
      const float su_m = scalar_is_one(su) ? 1.0f : su, si_m = scalar_is_one(si) ? 1.0f : si;
      float4 ni[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        c_wu[v] = f4_add_scaled(c_wu[v], ti[v], su_m, false);
        ni[v] = f4_add_scaled(wi[v], tu[v], si_m, false);
        if (!hp.du_skip) c_wu[v] = f4_scale(c_wu[v], hp.du);
        if (!hp.di_skip) ni[v] = f4_scale(ni[v], hp.di);
      }
      if (!m.no_user_bias) c_ub = __fmul_rn(__fadd_rn(c_ub, su), hp.dub);
      if (scatter_item == SCATTER_RED) g.red_row(m, irow, ni, wi);
      else g.store_row(m, irow, ni);
      if (g.gl == LANES - 1) {
        const float nib = __fmul_rn(__fadd_rn(ib, si), hp.dib);
        if (scatter_item == SCATTER_RED) red1(m.bias + irow, __fsub_rn(nib, ib));
        else __stcg(m.bias + irow, nib);
      }
      // update_svdpp (base.h:512-520)
      const float sf = __fmul_rn(__fmul_rn(hp.lr_fb, err), s.norm);
      const float sf_m = scalar_is_one(sf) ? 1.0f : sf;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        s.fb[v] = f4_add_scaled(s.fb[v], ti[v], sf_m, false);
        if (!hp.dfb_skip) s.fb[v] = f4_scale(s.fb[v], hp.dfb);
      }
      if (!m.no_user_bias) {
        s.fb_bias = __fadd_rn(s.fb_bias, sf);
        s.fb_bias = __fmul_rn(s.fb_bias, hp.dfbb);
      }
      return;
    }
    flush_user();  // the generic routine reads the user row from memory
    if (ORDERED && TRAIN) wait_tickets(g, m, rp0, rp1, rp2, rp3, idx, tk);
    const float p = process_instance<LANES, VEC, EXACT_DOT, TRAIN, true>(
        g, m, hp, rp0, rp1, rp2, rp3, label[r], idx, val, scatter_user, scatter_item, &s,
        err_flag, val2);
    if (!TRAIN && g.gl == 0) pred_out[r - ug.row_base] = p;
    if (ORDERED && TRAIN) release_tickets(g, m, rp0, rp1, rp2, rp3, idx);
  };
  auto finish_unit = [&](bool ok) {
    flush_user();
    if (coop) return;  // the warp scatters the feedback rows together (below)
    if (ok && TRAIN) update_ufeedback(g, m, ug.fb_index + f1, ug.fb_value + f1, nf1, s, old, old_bias, scatter_item);
    if (ORDERED && TRAIN) {
      __threadfence();
      g.gsync();
      for (int i = g.gl; i < nf0; i += LANES) {
        const unsigned id = ug.fb_index[f0 + i];
        if (id < (unsigned)m.num_ufeedback) red_release_add_u32(m.ver_ui + id, 1u);
      }
    }
  };

  if (ORDERED) {  // one unit per warp, units in input order
    for (;;) {
      bool ok = false;
      if (!next_unit(ok)) break;
      if (ok)
        for (; r_cur < r_end; ++r_cur) do_row(r_cur);
      finish_unit(ok);
    }
    return;
  }
  // Hogwild: one unit per lane group.  The groups of a warp are brought back together before
  // every row (ncu of the free-running version: 11 of 32 lanes active on average, ~600
  // warp-instructions per row): the row step -- the bulk of the work -- then runs as one SIMD
  // instruction stream for all groups that have a row, and only the per-unit gather / scatter
  // of the feedback rows runs divergent.
  bool have = false, done = false, ok = false;
  float *xch = xch_s[warp];
  const unsigned my_group = g.gmask;
  for (;;) {
    // ---- unit begin: every group without a unit takes one; the feedback gather of each is done
    //      by the whole warp, one group after the other
    const bool want = !have && !done;
    if (want) {
      have = next_unit(ok);
      done = !have;
      if (have && !coop && !ok) r_cur = r_end;  // unusable feedback list (error flagged): skip the rows
    }
    if (coop) {
      unsigned todo = __ballot_sync(0xffffffffu, want && have);
      while (todo) {
        const int src = __ffs(todo) - 1;  // first lane of the group served now
        todo &= ~__shfl_sync(0xffffffffu, my_group, src);
        const int cf0 = __shfl_sync(0xffffffffu, f0, src), cnf0 = __shfl_sync(0xffffffffu, nf0, src);
        float norm = 0.0f, fbb = 0.0f;
        bool good = false;
        if (cpl == 4) good = coop_prepare_ufeedback<4>(m, ug.fb_index + cf0, ug.fb_value + cf0, cnf0, lane, xch, norm, fbb, err_flag);
        else if (cpl == 2) good = coop_prepare_ufeedback<2>(m, ug.fb_index + cf0, ug.fb_value + cf0, cnf0, lane, xch, norm, fbb, err_flag);
        else good = coop_prepare_ufeedback<1>(m, ug.fb_index + cf0, ug.fb_value + cf0, cnf0, lane, xch, norm, fbb, err_flag);
        if (lane / LANES == src / LANES) {  // the owning group takes the result in its chunk layout
          ok = good;
          if (!good) r_cur = r_end;
          s.norm = norm;
          s.fb_bias = old_bias = fbb;
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const int ch = g.gl + v * LANES;
            s.fb[v] = (good && 4 * ch < m.pitch) ? *reinterpret_cast<const float4 *>(xch + 4 * ch) : f4_zero();
            old[v] = s.fb[v];
          }
        }
        __syncwarp();
      }
    }
    if (__all_sync(0xffffffffu, !have)) break;
    // ---- one rating row per group, all groups in step
    bool fin = false;
    if (have) {
      if (r_cur < r_end) {
        do_row(r_cur);
        ++r_cur;
      }
      if (r_cur >= r_end) {
        finish_unit(ok);
        have = false;
        fin = true;
      }
    }
    __syncwarp();
    // ---- unit end: the feedback scatter of each finished group, again by the whole warp
    if (coop) {
      unsigned todo = __ballot_sync(0xffffffffu, fin && ok && TRAIN);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= ~__shfl_sync(0xffffffffu, my_group, src);
        const int cf1 = __shfl_sync(0xffffffffu, f1, src), cnf1 = __shfl_sync(0xffffffffu, nf1, src);
        float dbias = 0.0f;
        if (lane / LANES == src / LANES && cnf1 > 0) {  // d = (tmp_ufeedback - old) / norm, base.h:541-546
          const float inv = __fdiv_rn(1.0f, s.norm);
          const bool inv_one = scalar_is_one(inv);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            float4 d = f4_sub(s.fb[v], old[v]);
            if (!inv_one) d = f4_scale(d, inv);
            const int ch = g.gl + v * LANES;
            if (4 * ch < m.pitch) *reinterpret_cast<float4 *>(xch + 4 * ch) = d;
          }
          dbias = __fmul_rn(__fsub_rn(s.fb_bias, old_bias), inv);
        }
        dbias = __shfl_sync(0xffffffffu, dbias, src);
        __syncwarp();
        if (cpl == 4) coop_update_ufeedback<4>(m, ug.fb_index + cf1, ug.fb_value + cf1, cnf1, lane, xch, dbias, scatter_item);
        else if (cpl == 2) coop_update_ufeedback<2>(m, ug.fb_index + cf1, ug.fb_value + cf1, cnf1, lane, xch, dbias, scatter_item);
        else coop_update_ufeedback<1>(m, ug.fb_index + cf1, ug.fb_value + cf1, cnf1, lane, xch, dbias, scatter_item);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// replicated-slab delta exchange (multi-GPU)
// ---------------------------------------------------------------------------
// mode 0: snap = cur ; 1: delta = cur - snap ; 2: cur = snap + scale*delta, snap = cur
__global__ void k_delta(DeltaPlan plan, float *snap, float *delta, int mode, float scale) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (int s = 0; s < plan.nseg; ++s) {
    const DeltaSeg sg = plan.seg[s];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < sg.n; i += stride) {
      const long long j = sg.off + i;
      if (mode == 0) snap[j] = sg.cur[i];
      else if (mode == 1) delta[j] = __fsub_rn(sg.cur[i], snap[j]);
      else {
        const float x = __fadd_rn(snap[j], __fmul_rn(scale, delta[j]));
        sg.cur[i] = x;
        snap[j] = x;
      }
    }
  }
}


template <int L, int V>
static int exact_geo(svdgpu *h, const DevCsr &csr, int r0, int r1) {
  auto k = k_exact<L, V>;
  int grid = 1;
  if (grid_for(h, k, EX_WARPS * 32, ((long long)(r1 - r0) + EX_WARPS - 1) / EX_WARPS, &grid)) return 1;
  CU(h, cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned), h->stream));
  CU(h, cudaMemsetAsync(h->dm.ver_ui, 0, sizeof(unsigned) * std::max<size_t>(h->rows, 1), h->stream));
  CU(h, cudaMemsetAsync(h->dm.ver_g, 0, sizeof(unsigned) * std::max(h->shape.num_global, 1), h->stream));
  k<<<grid, EX_WARPS * 32, 0, h->stream>>>(h->dm, h->dhp, csr, r0, r1, h->d_counter, h->d_err, h->exact_opt);
  CU(h, cudaGetLastError());
  h->n_launch++;
  return 0;
}

template <int L, int V>
static int ugroup_geo(svdgpu *h, const DevCsr &csr, const DevUgroup &ug, int u0, int u1, bool train,
                      bool ordered, float *pred) {
  int grid = 1;
  // Hogwild training with the default dot order: units made of basic-MF rows of one user go to
  // k_svdpp (one warp per unit, svdgpu_svdpp.cu); k_ugroup then skips them
  const unsigned char *kind = nullptr;
  if (train && !ordered && !h->exact_dot && h->dhp.plain && !csr.value2 && h->svdpp_fast) {
    const int units = h->ugroup_units > 0 ? h->ugroup_units : (ug.has_fb ? 64 : h->num_sm * 16);
    if (launch_svdpp(h, csr, ug, u0, u1, std::min(units, std::max(1, (u1 - u0) / 32)), &kind)) return 1;
  }
  CU(h, cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned), h->stream));
#define GO(ED, ORD, TR)                                                                         \
  {                                                                                             \
    auto k = k_ugroup<L, V, ED, ORD, TR>;                                                       \
    const int gpw = ORD ? 1 : 32 / L;                                                           \
    if (grid_for(h, k, EX_WARPS * 32, ((long long)(u1 - u0) + EX_WARPS * gpw - 1) / (EX_WARPS * gpw), &grid)) \
      return 1;                                                                                 \
    /* Hogwild across users is only SGD-like while the users in flight are a small   */        \
    /* fraction of the launch: cap them at 1/32 of the units.                          */        \
    if (!ORD && TR) grid = std::max(1, std::min(grid, (u1 - u0) / (32 * EX_WARPS * gpw)));      \
    /* With feedback lists (SVD++) the W_ufeedback rows of popular items are shared by a   */   \
    /* large share of all users: measured (tools/hogwild_parity.py --svdpp, 120k users,    */   \
    /* 100 feedback entries each) 32..64 users in flight track the sequential order to     */   \
    /* 1e-2, 128 wobble, 256 drift away and >= 600 end in NaN.  Default 64.                */   \
    {                                                                                           \
      const int units = h->ugroup_units > 0 ? h->ugroup_units : (ug.has_fb ? 64 : 0);           \
      if (!ORD && TR && units > 0) grid = std::max(1, std::min(grid, units / (EX_WARPS * gpw))); \
    }                                                                                           \
    k<<<grid, EX_WARPS * 32, 0, h->stream>>>(h->dm, h->dhp, csr, ug, u0, u1, h->scatter_user,   \
                                             h->scatter_item, h->d_counter, pred, h->d_err,     \
                                             (!ORD && TR) ? kind : nullptr);                    \
  }
  if (!train) {
    GO(true, false, false)
  } else if (ordered) {
    CU(h, cudaMemsetAsync(h->dm.ver_ui, 0, sizeof(unsigned) * std::max<size_t>(h->rows, 1), h->stream));
    CU(h, cudaMemsetAsync(h->dm.ver_g, 0, sizeof(unsigned) * std::max(h->shape.num_global, 1), h->stream));
    GO(true, true, true)
  } else {
    if (h->exact_dot) GO(true, false, true) else GO(false, false, true)
  }
#undef GO
  CU(h, cudaGetLastError());
  h->n_launch++;
  return 0;
}

// the ordered / user-group kernels are instantiated for the automatic geometries only
#ifdef SVDGPU_TUNE_BUILD
#define ORDERED_GEOS(X) X(4, 1) X(8, 2)
#else
#define ORDERED_GEOS(X) X(4, 1) X(4, 2) X(8, 2) X(16, 2) X(32, 2) X(32, 4) X(16, 1) X(32, 1)
#endif

int launch_exact(svdgpu *h, const Geometry &g, const DevCsr &csr, int r0, int r1) {
#define GEO(L, V) \
  if (g.lanes == L && g.vec == V) return exact_geo<L, V>(h, csr, r0, r1);
  ORDERED_GEOS(GEO)
#undef GEO
  return fail(h, "ordered mode: no kernel for lanes=%d vec=%d (unset option lanes)", g.lanes, g.vec);
}

int launch_ugroup(svdgpu *h, const Geometry &g, const DevCsr &csr, const DevUgroup &ug, int u0, int u1,
                  bool train, bool ordered, float *pred) {
#define GEO(L, V) \
  if (g.lanes == L && g.vec == V) return ugroup_geo<L, V>(h, csr, ug, u0, u1, train, ordered, pred);
  ORDERED_GEOS(GEO)
#undef GEO
  return fail(h, "user-group input: no kernel for lanes=%d vec=%d (unset option lanes)", g.lanes, g.vec);
}

int launch_delta(svdgpu *h, int mode, float scale) {
  k_delta<<<h->num_sm * 4, 256, 0, h->stream>>>(h->plan, h->d_snap, h->d_delta, mode, scale);
  CU(h, cudaGetLastError());
  h->n_launch++;
  return 0;
}

}  // namespace svdk
