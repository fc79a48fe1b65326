// svdgpu_stream.cu -- k_stream: Hogwild training / prediction over a CSR batch.
//
// Persistent CTAs (256 threads, 2 per SM).  A producer warp stages each
// 128-instance tile into shared memory with 1-D bulk asynchronous copies
// (cp.async.bulk = TMA unit, SASS UBLKCP; completion on mbarriers): phase A brings
// the tile's row_ptr/label window, phase B -- once A has landed and the feature
// range is known -- its index/value window.  Four tiles are in flight and the two
// phases of consecutive tiles overlap, so consumers never chase
// row_ptr -> index -> row through DRAM.  Seven consumer warps run one lane GROUP
// per instance (svdgpu_device.cuh).
//
// Two passes over a batch (Hogwild has no order to keep):
//   pass 1 (SIMPLE)  rows of the basic-MF shape -- no global feature, one user
//           feature, one item feature (configs[0..1]) -- take a straight-line path
//           whose gathers for the NEXT instance are issued before the current
//           instance is computed (register double buffering).  A tile that holds
//           any other row shape (or did not fit the staging window) is flagged.
//   pass 2 (GENERIC) visits flagged tiles only and runs the remaining rows through
//           the generic process_instance().  For configs[1] it finds nothing.
// Both paths perform bit-identical arithmetic.
#include "svdgpu_internal.h"

namespace svdk {

constexpr int HW_TILE = 128;           // instances per tile
constexpr int HW_STAGES = 4;           // tiles in flight per CTA
constexpr int HW_CAP = 4 * HW_TILE;    // staged index/value entries per tile
constexpr int HW_CWARPS = 7;           // consumer warps per CTA (+1 producer = 256 threads)
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
  int tile;
  int skip;     // GENERIC pass: tile not flagged, nothing to do
  int v_hi;     // one past the last staged feature position (absolute)
};

template <int LANES, int VEC, bool EXACT_DOT>
struct HwSmem {
  HwStage st[HW_STAGES];
  uint64_t barA[HW_STAGES], full[HW_STAGES], empty[HW_STAGES];
  HwMeta meta[HW_STAGES];
  float dot[EXACT_DOT ? HW_CWARPS * (32 / LANES) * Group<LANES, VEC>::DOT_FLOATS : 4];
};

// gathers of one instance, issued ahead of its compute
template <int VEC>
struct Pre {
  float4 wu[VEC], wi[VEC];
  float ub, ib, uval, ival, label;
  unsigned uid, iid;
  int q;      // row inside the tile, -1: nothing loaded
};

__device__ __forceinline__ bool is_simple(const int *rp, int q) {
  const int rp0 = rp[3 * q], rp1 = rp[3 * q + 1], rp2 = rp[3 * q + 2], rp3 = rp[3 * q + 3];
  return rp1 == rp0 && rp2 == rp1 + 1 && rp3 == rp2 + 1;
}

template <int LANES, int VEC, bool EXACT_DOT, bool TRAIN, bool GENERIC>
__global__ void __launch_bounds__(HW_THREADS, 2)
k_stream(DevModel m, DevHP hp, DevCsr csr, int row_begin, int row_end, int scatter_user,
         int scatter_item, float *pred_out, int *tile_flag, int *err_flag) {
  __shared__ HwSmem<LANES, VEC, EXACT_DOT> sm;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile = (row_end - row_begin + HW_TILE - 1) / HW_TILE;
  const int nlocal = ntile > (int)blockIdx.x ? (ntile - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

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
    // ===== producer: one lane drives the bulk copies; phase A of tile i overlaps
    // phase B of tile i-1 =====
    if (lane == 0) {
      for (int it = 0; it <= nlocal; ++it) {
        if (it < nlocal) {  // phase A of local tile `it`
          const int s = it % HW_STAGES;
          const unsigned ph = (it / HW_STAGES) & 1;
          mbar_wait_sleepy(&sm.empty[s], ph ^ 1);
          HwStage &st = sm.st[s];
          const int t = blockIdx.x + it * gridDim.x;
          const int r0 = row_begin + t * HW_TILE;
          const int nrow = min(HW_TILE, row_end - r0);
          HwMeta mt;
          mt.tile = t; mt.r0 = r0; mt.nrow = nrow;
          mt.a_off = (3 * r0) & 3; mt.l_off = r0 & 3;
          mt.sm_base = 0; mt.staged = 0; mt.v_hi = 0;
          mt.skip = (GENERIC && tile_flag[t] == 0) ? 1 : 0;
          sm.meta[s] = mt;
          if (mt.skip) {
            mbar_arrive(&sm.barA[s]);
          } else {
            const unsigned bytesA = (unsigned)((mt.a_off + 3 * nrow + 1 + 3) & ~3) * 4u;
            const unsigned bytesL = (unsigned)((mt.l_off + nrow + 3) & ~3) * 4u;
            mbar_arrive_expect_tx(&sm.barA[s], bytesA + bytesL);
            bulk_g2s(st.rp, csr.row_ptr + (3 * r0 - mt.a_off), bytesA, &sm.barA[s]);
            bulk_g2s(st.label, csr.label + (r0 - mt.l_off), bytesL, &sm.barA[s]);
          }
        }
        const int j = it - 1;  // phase B of local tile `j`
        if (j >= 0) {
          const int s = j % HW_STAGES;
          const unsigned ph = (j / HW_STAGES) & 1;
          mbar_wait_sleepy(&sm.barA[s], ph);
          HwStage &st = sm.st[s];
          HwMeta &mt = sm.meta[s];
          if (mt.skip) {
            mbar_arrive(&sm.full[s]);
          } else {
            const int v0 = st.rp[mt.a_off] - csr.val_base;
            const int v1 = st.rp[mt.a_off + 3 * mt.nrow] - csr.val_base;
            const int v_off = v0 & 3;
            const int nel = (v_off + (v1 - v0) + 3) & ~3;
            mt.sm_base = v0 - v_off + csr.val_base;
            mt.v_hi = v1 + csr.val_base;
            mt.staged = (v0 >= 0 && v1 >= v0 && v1 + csr.val_base <= csr.val_end && nel <= HW_CAP + 8) ? 1 : 0;
            if (mt.staged && nel > 0) {
              mbar_arrive_expect_tx(&sm.full[s], 2u * (unsigned)nel * 4u);
              bulk_g2s(st.idx, csr.index + (v0 - v_off), (unsigned)nel * 4u, &sm.full[s]);
              bulk_g2s(st.val, csr.value + (v0 - v_off), (unsigned)nel * 4u, &sm.full[s]);
            } else {
              mbar_arrive(&sm.full[s]);
            }
          }
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
  g.dot_s = EXACT_DOT ? &sm.dot[(warp * GPW + gw) * Group<LANES, VEC>::DOT_FLOATS] : nullptr;
  const int gid = warp * GPW + gw;
  constexpr int NGROUP = HW_CWARPS * GPW;

  for (int it = 0; it < nlocal; ++it) {
    const int s = it % HW_STAGES;
    const unsigned ph = (it / HW_STAGES) & 1;
    mbar_wait(&sm.full[s], ph);  // the producer arrives on `full` only after phase A has landed
    const HwMeta mt = sm.meta[s];
    const HwStage &st = sm.st[s];
    const int *rp = st.rp + mt.a_off;
    const float *lab = st.label + mt.l_off;

    if (GENERIC) {
      // ---- pass 2: whatever pass 1 left in this (flagged) tile ----
      if (!mt.skip) {
        const unsigned *idx = mt.staged ? (st.idx - mt.sm_base) : (csr.index - csr.val_base);
        const float *val = mt.staged ? (st.val - mt.sm_base) : (csr.value - csr.val_base);
        for (int q = gid; q < mt.nrow; q += NGROUP) {
          if (mt.staged && is_simple(rp, q) && rp[3 * q] >= mt.sm_base && rp[3 * q + 3] <= mt.v_hi) continue;  // done by pass 1
          if (!row_ok(rp[3 * q], rp[3 * q + 1], rp[3 * q + 2], rp[3 * q + 3], csr.val_base, csr.val_end)) {
            if (g.gl == 0) atomicCAS(err_flag, 0, ERR_ROW_PTR);
            continue;
          }
          const float pr = process_instance<LANES, VEC, EXACT_DOT, TRAIN, false>(
              g, m, hp, rp[3 * q], rp[3 * q + 1], rp[3 * q + 2], rp[3 * q + 3], lab[q], idx, val,
              scatter_user, scatter_item, nullptr, err_flag);
          if (!TRAIN && g.gl == 0) pred_out[mt.r0 + q - row_begin] = pr;
        }
      }
    } else if (!mt.staged) {
      // window did not fit: the whole tile goes to pass 2
      if (threadIdx.x == 0) tile_flag[mt.tile] = 1;
    } else {
      // ---- pass 1: straight-line basic-MF rows, gathers one instance ahead ----
      const unsigned *sidx = st.idx - mt.sm_base;
      const float *sval = st.val - mt.sm_base;
      bool other = false;  // this group met a row of another shape

      auto pre_load = [&](Pre<VEC> &p, int q) {
        p.q = -1;
        if (q >= mt.nrow) return;
        const int f = rp[3 * q + 1];
        if (!is_simple(rp, q) || f < mt.sm_base || f + 2 > mt.v_hi) {
          other = true;
          return;
        }
        p.uid = sidx[f];
        p.iid = sidx[f + 1];
        if (p.uid >= (unsigned)m.num_user || p.iid >= (unsigned)m.num_item) {
          if (g.gl == 0) atomicCAS(err_flag, 0, p.uid >= (unsigned)m.num_user ? ERR_USER_INDEX : ERR_ITEM_INDEX);
          return;
        }
        p.q = q;
        p.uval = sval[f];
        p.ival = sval[f + 1];
        p.label = lab[q];
        g.load_row(m, (size_t)m.user_off + p.uid, p.wu);
        g.load_row(m, (size_t)m.item_off + p.iid, p.wi);
        p.ub = m.no_user_bias ? 0.0f : __ldcg(m.bias + m.user_off + p.uid);
        p.ib = __ldcg(m.bias + m.item_off + p.iid);
      };

      auto compute = [&](const Pre<VEC> &p) {
        if (p.q < 0) return;
        // prepare_tmp (base.h:354-381): tmp = 0 + w*val
        float4 tu[VEC], ti[VEC];
        const bool one_u = scalar_is_one(p.uval), one_i = scalar_is_one(p.ival);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          tu[v] = f4_add_scaled(f4_zero(), p.wu[v], p.uval, one_u);
          ti[v] = f4_add_scaled(f4_zero(), p.wi[v], p.ival, one_i);
        }
        // calc_bias (base.h:313-353) + pred (base.h:445-454)
        double bsum = 0.0;
        if (!m.no_user_bias) bsum = __dadd_rn(bsum, (double)__fmul_rn(p.uval, p.ub));
        bsum = __dadd_rn(bsum, (double)__fmul_rn(p.ival, p.ib));
        const float d = g.template dot<EXACT_DOT>(m, tu, ti);
        double sum = __dadd_rn((double)hp.base_score, bsum);
        sum = __dadd_rn(sum, (double)d);
        const float pred = map_active((float)sum, m.active_type);
        if (!TRAIN) {
          if (g.gl == 0) pred_out[mt.r0 + p.q - row_begin] = pred;
          return;
        }
        // update_no_decay + regularize(after), fused (base.h:383-427, 211-283)
        const float err = cal_grad(p.label, pred, m.active_type);
        const float lrerr = __fmul_rn(hp.lr, err);
        const float su = __fmul_rn(lrerr, p.uval), si = __fmul_rn(lrerr, p.ival);
        const bool one_su = scalar_is_one(su), one_si = scalar_is_one(si);
        float4 nw[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          nw[v] = f4_add_scaled(p.wu[v], ti[v], su, one_su);
          if (!hp.du_skip) nw[v] = f4_scale(nw[v], hp.du);
        }
        if (scatter_user == SCATTER_RED) g.red_row(m, (size_t)m.user_off + p.uid, nw, p.wu);
        else g.store_row(m, (size_t)m.user_off + p.uid, nw);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          nw[v] = f4_add_scaled(p.wi[v], tu[v], si, one_si);
          if (!hp.di_skip) nw[v] = f4_scale(nw[v], hp.di);
        }
        if (scatter_item == SCATTER_RED) g.red_row(m, (size_t)m.item_off + p.iid, nw, p.wi);
        else g.store_row(m, (size_t)m.item_off + p.iid, nw);
        if (g.gl == 0 && !m.no_user_bias) {
          float *bp = m.bias + m.user_off + p.uid;
          const float nb = __fmul_rn(__fadd_rn(p.ub, su), hp.dub);
          if (scatter_user == SCATTER_RED) red1(bp, __fsub_rn(nb, p.ub));
          else __stcg(bp, nb);
        }
        if (g.gl == 1) {
          float *bp = m.bias + m.item_off + p.iid;
          const float nb = __fmul_rn(__fadd_rn(p.ib, si), hp.dib);
          if (scatter_item == SCATTER_RED) red1(bp, __fsub_rn(nb, p.ib));
          else __stcg(bp, nb);
        }
      };

      Pre<VEC> A, B;
      int q = gid;
      pre_load(A, q);
      while (q < mt.nrow) {
        pre_load(B, q + NGROUP);
        compute(A);
        q += NGROUP;
        if (q >= mt.nrow) break;
        pre_load(A, q + NGROUP);
        compute(B);
        q += NGROUP;
      }
      if (__any_sync(g.gmask, other) && g.gl == 0) tile_flag[mt.tile] = 1;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.empty[s]);
  }
}

template <int L, int V>
static int launch_geo(svdgpu *h, const DevCsr &csr, int r0, int r1, bool train, float *pred) {
  const long long ntile = ((long long)(r1 - r0) + HW_TILE - 1) / HW_TILE;
  // per-tile flags: which tiles pass 2 must visit
  if ((size_t)ntile * sizeof(int) > h->tile_flag_cap) {
    if (h->d_tile_flag) CU(h, cudaFree(h->d_tile_flag));
    h->d_tile_flag = nullptr;
    h->tile_flag_cap = 0;
    const size_t cap = (size_t)ntile * sizeof(int) * 2;
    CU(h, cudaMalloc(&h->d_tile_flag, cap));
    h->tile_flag_cap = cap;
  }
  CU(h, cudaMemsetAsync(h->d_tile_flag, 0, (size_t)ntile * sizeof(int), h->stream));
  int grid = 1;
#define GO(ED, TR, GEN)                                                                          \
  {                                                                                              \
    auto k = k_stream<L, V, ED, TR, GEN>;                                                        \
    if (grid_for(h, k, HW_THREADS, ntile, &grid)) return 1;                                      \
    k<<<grid, HW_THREADS, 0, h->stream>>>(h->dm, h->dhp, csr, r0, r1, h->scatter_user,           \
                                          h->scatter_item, pred, h->d_tile_flag, h->d_err);      \
    h->n_launch++;                                                                               \
  }
  if (train) {
    if (h->exact_dot) { GO(true, true, false) GO(true, true, true) }
    else { GO(false, true, false) GO(false, true, true) }
  } else {
    GO(true, false, false) GO(true, false, true)
  }
#undef GO
  CU(h, cudaGetLastError());
  return 0;
}

int launch_stream(svdgpu *h, const Geometry &g, const DevCsr &csr, int r0, int r1, bool train,
                  float *pred) {
#define GEO(L, V) \
  if (g.lanes == L && g.vec == V) return launch_geo<L, V>(h, csr, r0, r1, train, pred);
#ifdef SVDGPU_TUNE_BUILD
  GEO(4, 1) GEO(4, 4) GEO(8, 2) GEO(16, 1)
#else
  GEO(4, 1) GEO(4, 2) GEO(4, 4) GEO(8, 1) GEO(8, 2) GEO(8, 4) GEO(16, 1) GEO(16, 2) GEO(16, 4)
  GEO(32, 1) GEO(32, 2) GEO(32, 4)
#endif
#undef GEO
  return fail(h, "k_stream: no kernel instantiated for lanes=%d vec=%d", g.lanes, g.vec);
}

}  // namespace svdk
