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

namespace svdk {

// ---------------------------------------------------------------------------
// k_exact
// ---------------------------------------------------------------------------
constexpr int EX_WARPS = 4;

__device__ __forceinline__ unsigned *version_of(const DevModel &m, int f, int rp1, int rp2, unsigned id) {
  if (f < rp1) return id < (unsigned)m.num_global ? m.ver_g + id : nullptr;
  if (f < rp2) return id < (unsigned)m.num_user ? m.ver_ui + m.user_off + id : nullptr;
  return id < (unsigned)m.num_item ? m.ver_ui + m.item_off + id : nullptr;
}

template <int LANES, int VEC>
__device__ __forceinline__ void wait_tickets(const Group<LANES, VEC> &g, const DevModel &m, int rp0,
                                             int rp1, int rp2, int rp3, const unsigned *idx,
                                             const unsigned *tk) {
  for (int f = rp0 + g.gl; f < rp3; f += LANES) {
    const unsigned *ver = version_of(m, f, rp1, rp2, idx[f]);
    if (ver) {
      const unsigned want = tk[f];
      unsigned ns = 20;
      while (ld_acquire_u32(ver) != want) {
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
                                                const unsigned *idx) {
  __threadfence();
  g.gsync();
  for (int f = rp0 + g.gl; f < rp3; f += LANES) {
    unsigned *ver = version_of(m, f, rp1, rp2, idx[f]);
    if (ver) red_release_add_u32(ver, 1u);
  }
}

template <int LANES, int VEC>
__global__ void __launch_bounds__(EX_WARPS * 32)
k_exact(DevModel m, DevHP hp, DevCsr csr, int row_begin, int row_end, unsigned *counter,
        int *err_flag) {
  __shared__ float dot_s[EX_WARPS][Group<LANES, VEC>::DOT_FLOATS];
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
    wait_tickets(g, m, rp0, rp1, rp2, rp3, idx, tk);
    process_instance<LANES, VEC, true, true, false>(g, m, hp, rp0, rp1, rp2, rp3, csr.label[r], idx,
                                                    val, SCATTER_STORE, SCATTER_STORE, nullptr,
                                                    err_flag, val2);
    release_tickets(g, m, rp0, rp1, rp2, rp3, idx);
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
         int scatter_user, int scatter_item, unsigned *counter, float *pred_out, int *err_flag) {
  constexpr int GPW = ORDERED ? 1 : 32 / LANES;
  __shared__ float dot_s[EX_WARPS][GPW * Group<LANES, VEC>::DOT_FLOATS];
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

  for (;;) {
    unsigned n = 0;
    if (g.gl == 0) n = atomicAdd(counter, 1u);
    n = g.bcast(n, 0);
    if ((long long)unit_begin + n >= unit_end) break;
    int u = unit_begin + (int)n;
    if (!ORDERED && ug.order) u = ug.order[u];
    const int b0 = ug.unit_off[u], b1 = ug.unit_off[u + 1];
    // feedback list of the first block feeds the gather, of the last block the scatter
    const int f0 = ug.blk_fb_off[b0] - ug.fb_base, nf0 = ug.blk_fb_off[b0 + 1] - ug.blk_fb_off[b0];
    const int f1 = ug.blk_fb_off[b1 - 1] - ug.fb_base, nf1 = ug.blk_fb_off[b1] - ug.blk_fb_off[b1 - 1];
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
    FbState<VEC> s;
    float4 old[VEC];
    const bool ok = prepare_ufeedback(g, m, ug.fb_index + f0, ug.fb_value + f0, nf0, s, err_flag);
    const float old_bias = s.fb_bias;
#pragma unroll
    for (int v = 0; v < VEC; ++v) old[v] = s.fb[v];
    if (ok) {
      for (int r = ug.blk_row_off[b0]; r < ug.blk_row_off[b1]; ++r) {
        const int rp0 = row_ptr[3 * (long long)r], rp1 = row_ptr[3 * (long long)r + 1];
        const int rp2 = row_ptr[3 * (long long)r + 2], rp3 = row_ptr[3 * (long long)r + 3];
        if (!ORDERED && !row_ok(rp0, rp1, rp2, rp3, csr.val_base, csr.val_end)) {
          if (g.gl == 0) atomicCAS(err_flag, 0, ERR_ROW_PTR);
          continue;
        }
        if (ORDERED && TRAIN) wait_tickets(g, m, rp0, rp1, rp2, rp3, idx, tk);
        const float p = process_instance<LANES, VEC, EXACT_DOT, TRAIN, true>(
            g, m, hp, rp0, rp1, rp2, rp3, label[r], idx, val, scatter_user, scatter_item, &s,
            err_flag, val2);
        if (!TRAIN && g.gl == 0) pred_out[r - ug.row_base] = p;
        if (ORDERED && TRAIN) release_tickets(g, m, rp0, rp1, rp2, rp3, idx);
      }
      if (TRAIN)
        update_ufeedback(g, m, ug.fb_index + f1, ug.fb_value + f1, nf1, s, old, old_bias,
                         scatter_item);
    }
    if (ORDERED && TRAIN) {
      __threadfence();
      g.gsync();
      for (int i = g.gl; i < nf0; i += LANES) {
        const unsigned id = ug.fb_index[f0 + i];
        if (id < (unsigned)m.num_ufeedback) red_release_add_u32(m.ver_ui + id, 1u);
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
  k<<<grid, EX_WARPS * 32, 0, h->stream>>>(h->dm, h->dhp, csr, r0, r1, h->d_counter, h->d_err);
  CU(h, cudaGetLastError());
  h->n_launch++;
  return 0;
}

template <int L, int V>
static int ugroup_geo(svdgpu *h, const DevCsr &csr, const DevUgroup &ug, int u0, int u1, bool train,
                      bool ordered, float *pred) {
  int grid = 1;
  CU(h, cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned), h->stream));
#define GO(ED, ORD, TR)                                                                         \
  {                                                                                             \
    auto k = k_ugroup<L, V, ED, ORD, TR>;                                                       \
    const int gpw = ORD ? 1 : 32 / L;                                                           \
    if (grid_for(h, k, EX_WARPS * 32, ((long long)(u1 - u0) + EX_WARPS * gpw - 1) / (EX_WARPS * gpw), &grid)) \
      return 1;                                                                                 \
    /* Hogwild across users is only SGD-like while the users in flight are a small   */        \
    /* fraction of the launch: cap them at 1/32 of the units (no effect at C3 scale). */        \
    if (!ORD && TR) grid = std::max(1, std::min(grid, (u1 - u0) / (32 * EX_WARPS * gpw)));      \
    k<<<grid, EX_WARPS * 32, 0, h->stream>>>(h->dm, h->dhp, csr, ug, u0, u1, h->scatter_user,   \
                                             h->scatter_item, h->d_counter, pred, h->d_err);    \
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
