// svdgpu_kernels.cuh -- the __global__ kernels built on svdgpu_device.cuh.
//
//   k_stream  : Hogwild training / prediction over a CSR batch.  Persistent CTAs;
//               a producer warp stages each tile's row_ptr/label and index/value
//               slices into shared memory with 1-D bulk async copies (TMA,
//               mbarrier-tracked, double buffered); 8 consumer warps run one lane
//               group per instance.
//   k_exact   : ordered data-flow training.  One instance per warp, fetched in
//               input order; each feature waits until its row's version counter
//               reaches the feature's ticket, so the result equals the reference's
//               sequential loop while independent instances run in parallel.
//   k_ugroup  : user-grouped (SVD++) blocks, one lane group (Hogwild) or one warp
//               (exact) per user unit: feedback gather, the unit's rows in order,
//               feedback scatter (base.h:523-582).
//   k_delta_* : replicated-slab delta pack/apply for the multi-GPU exchange.
#pragma once
#include "svdgpu_device.cuh"

namespace svdk {

// ---------------------------------------------------------------------------
// k_stream
// ---------------------------------------------------------------------------
constexpr int HW_TILE = 128;           // instances per tile
constexpr int HW_STAGES = 2;           // tiles in flight per CTA
constexpr int HW_CAP = 8 * HW_TILE;    // staged index/value entries per tile
constexpr int HW_CWARPS = 8;           // consumer warps per CTA
constexpr int HW_THREADS = (HW_CWARPS + 1) * 32;

struct __align__(16) HwStage {
  int rp[3 * HW_TILE + 8];
  float label[HW_TILE + 4];
  unsigned idx[HW_CAP + 8];
  float val[HW_CAP + 8];
};
struct HwMeta {
  int a_off;    // rp[a_off] is row_ptr[3*r0]
  int l_off;    // label[l_off] is label[r0]
  int sm_base;  // absolute feature position held by idx[0]/val[0]
  int staged;   // 0: the tile's features did not fit, read them from global
  int nrow;
  int r0;
};

template <int LANES, int VEC>
struct HwSmem {
  HwStage st[HW_STAGES];
  uint64_t barA[HW_STAGES], full[HW_STAGES], empty[HW_STAGES];
  HwMeta meta[HW_STAGES];
  float dot[HW_CWARPS][(32 / LANES) * Group<LANES, VEC>::DOT_FLOATS];
};

template <int LANES, int VEC, bool EXACT_DOT, bool TRAIN>
__global__ void __launch_bounds__(HW_THREADS)
k_stream(DevModel m, DevHP hp, DevCsr csr, int row_begin, int row_end, int scatter_user,
         int scatter_item, float *pred_out, int *err_flag) {
  __shared__ HwSmem<LANES, VEC> sm;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile = (row_end - row_begin + HW_TILE - 1) / HW_TILE;

  if (threadIdx.x == 0) {
    for (int s = 0; s < HW_STAGES; ++s) {
      mbar_init(&sm.barA[s], 1);
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], HW_CWARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == HW_CWARPS) {
    // ===== producer: one lane drives the bulk copies =====
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < ntile; t += gridDim.x, ++it) {
        const int s = it % HW_STAGES;
        const unsigned ph = (it / HW_STAGES) & 1;
        mbar_wait(&sm.empty[s], ph ^ 1);
        HwStage &st = sm.st[s];
        const int r0 = row_begin + t * HW_TILE;
        const int nrow = min(HW_TILE, row_end - r0);
        // phase A: row_ptr[3*r0 .. 3*(r0+nrow)] and label[r0 .. r0+nrow), 16-byte aligned windows
        const int a_off = (3 * r0) & 3, l_off = r0 & 3;
        const unsigned bytesA = (unsigned)((a_off + 3 * nrow + 1 + 3) & ~3) * 4u;
        const unsigned bytesL = (unsigned)((l_off + nrow + 3) & ~3) * 4u;
        mbar_arrive_expect_tx(&sm.barA[s], bytesA + bytesL);
        bulk_g2s(st.rp, csr.row_ptr + (3 * r0 - a_off), bytesA, &sm.barA[s]);
        bulk_g2s(st.label, csr.label + (r0 - l_off), bytesL, &sm.barA[s]);
        mbar_wait(&sm.barA[s], ph);
        // phase B: the tile's feature slice
        const int v0 = st.rp[a_off] - csr.val_base;
        const int v1 = st.rp[a_off + 3 * nrow] - csr.val_base;
        const int v_off = v0 & 3;
        const int nel = (v_off + (v1 - v0) + 3) & ~3;
        HwMeta mt;
        mt.a_off = a_off; mt.l_off = l_off; mt.nrow = nrow; mt.r0 = r0;
        mt.sm_base = v0 - v_off + csr.val_base;
        mt.staged = (nel <= HW_CAP + 8) ? 1 : 0;
        sm.meta[s] = mt;
        if (mt.staged && nel > 0) {
          mbar_arrive_expect_tx(&sm.full[s], 2u * (unsigned)nel * 4u);
          bulk_g2s(st.idx, csr.index + (v0 - v_off), (unsigned)nel * 4u, &sm.full[s]);
          bulk_g2s(st.val, csr.value + (v0 - v_off), (unsigned)nel * 4u, &sm.full[s]);
        } else {
          mbar_arrive(&sm.full[s]);
        }
      }
    }
    return;
  }

  // ===== consumers =====
  constexpr int GPW = 32 / LANES;  // groups per warp
  Group<LANES, VEC> g;
  g.gl = lane % LANES;
  const int gw = lane / LANES;
  g.gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (gw * LANES));
  g.dot_s = &sm.dot[warp][gw * Group<LANES, VEC>::DOT_FLOATS];
  const int gid = warp * GPW + gw;
  constexpr int NGROUP = HW_CWARPS * GPW;

  int it = 0;
  for (int t = blockIdx.x; t < ntile; t += gridDim.x, ++it) {
    const int s = it % HW_STAGES;
    const unsigned ph = (it / HW_STAGES) & 1;
    mbar_wait(&sm.barA[s], ph);
    mbar_wait(&sm.full[s], ph);
    const HwMeta mt = sm.meta[s];
    const HwStage &st = sm.st[s];
    const int *rp = st.rp + mt.a_off;
    const float *lab = st.label + mt.l_off;
    const unsigned *idx = mt.staged ? (st.idx - mt.sm_base) : (csr.index - csr.val_base);
    const float *val = mt.staged ? (st.val - mt.sm_base) : (csr.value - csr.val_base);
    for (int q = gid; q < mt.nrow; q += NGROUP) {
      const int rp0 = rp[3 * q], rp1 = rp[3 * q + 1], rp2 = rp[3 * q + 2], rp3 = rp[3 * q + 3];
      const float p = process_instance<LANES, VEC, EXACT_DOT, TRAIN, false>(
          g, m, hp, rp0, rp1, rp2, rp3, lab[q], idx, val, scatter_user, scatter_item, nullptr,
          err_flag);
      if (!TRAIN && g.gl == 0) pred_out[mt.r0 + q - row_begin] = p;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.empty[s]);
  }
}

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
  const unsigned *tk = csr.ticket - csr.val_base;
  for (;;) {
    unsigned n = 0;
    if (lane == 0) n = atomicAdd(counter, 1u);
    n = g.bcast(n, 0);
    const long long r = (long long)row_begin + n;
    if (r >= row_end) break;
    const int rp0 = csr.row_ptr[3 * r], rp1 = csr.row_ptr[3 * r + 1];
    const int rp2 = csr.row_ptr[3 * r + 2], rp3 = csr.row_ptr[3 * r + 3];
    wait_tickets(g, m, rp0, rp1, rp2, rp3, idx, tk);
    process_instance<LANES, VEC, true, true, false>(g, m, hp, rp0, rp1, rp2, rp3, csr.label[r], idx,
                                                    val, SCATTER_STORE, SCATTER_STORE, nullptr,
                                                    err_flag);
    release_tickets(g, m, rp0, rp1, rp2, rp3, idx);
  }
}

// ---------------------------------------------------------------------------
// k_ugroup
// ---------------------------------------------------------------------------
// A "unit" is one user's consecutive blocks: a DEFAULT block, or START..END
// (apex_svd_data.h:353-371).  unit_off[u]..unit_off[u+1] are its blocks.
struct DevUgroup {
  const int *unit_off;     // [num_unit+1] block ranges
  const int *blk_row_off;  // [num_block+1]
  const int *blk_fb_off;   // [num_block+1]
  const unsigned *fb_index;
  const float *fb_value;
  const unsigned *fb_ticket;  // exact mode
  const int *order;           // hogwild: unit processing order (longest first), may be null
  int row_base;               // blk_row_off values are absolute rows; csr arrays start at row_base
  int fb_base;                // blk_fb_off values are absolute; fb arrays start at fb_base
};

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
        if (ORDERED && TRAIN) wait_tickets(g, m, rp0, rp1, rp2, rp3, idx, tk);
        const float p = process_instance<LANES, VEC, EXACT_DOT, TRAIN, true>(
            g, m, hp, rp0, rp1, rp2, rp3, label[r], idx, val, scatter_user, scatter_item, &s,
            err_flag);
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
struct DeltaSeg {
  float *cur;       // live slab segment
  long long n;      // floats
  long long off;    // offset in the packed buffers
};
struct DeltaPlan {
  DeltaSeg seg[5];
  int nseg;
  long long total;
};
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

}  // namespace svdk
