// svdgpu_stream.cu -- k_stream: Hogwild training / prediction over a CSR batch.
//
// Persistent CTAs.  A producer warp stages each tile's row_ptr/label window and
// its index/value window into shared memory with 1-D bulk asynchronous copies
// (cp.async.bulk = TMA unit, SASS UBLKCP; completion on mbarriers; two tiles in
// flight), so the consumers never chase row_ptr -> index -> row through DRAM.
// Eight consumer warps run one lane GROUP per instance (svdgpu_device.cuh).
//
// Rows of the common shape -- no global feature, one user feature, one item
// feature (configs[0..1]: basicMF) -- take a straight-line path whose gathers for
// the NEXT instance are issued before the current instance is computed (register
// double buffering), which doubles the bytes each warp keeps in flight.  Every
// other row shape goes through the generic process_instance().  Both paths
// perform bit-identical arithmetic.
#include "svdgpu_internal.h"

namespace svdk {

constexpr int HW_TILE = 256;           // instances per tile
constexpr int HW_STAGES = 2;           // tiles in flight per CTA
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
  int q;
  int kind;  // 0 none, 1 simple (0|1|1 features), 2 generic, 3 index out of bound
};
enum { PRE_NONE = 0, PRE_SIMPLE = 1, PRE_GENERIC = 2, PRE_BAD = 3 };

// the generic row path, kept out of line so that it does not inflate the register
// budget of the straight-line path
template <int LANES, int VEC, bool EXACT_DOT, bool TRAIN>
__device__ __noinline__ float generic_instance(const Group<LANES, VEC> &g, const DevModel &m, const DevHP &hp,
                                               int rp0, int rp1, int rp2, int rp3, float label,
                                               const unsigned *idx, const float *val, int scatter_user,
                                               int scatter_item, int *err_flag) {
  return process_instance<LANES, VEC, EXACT_DOT, TRAIN, false>(g, m, hp, rp0, rp1, rp2, rp3, label, idx, val,
                                                               scatter_user, scatter_item, nullptr, err_flag);
}

template <int LANES, int VEC, bool EXACT_DOT, bool TRAIN>
__global__ void __launch_bounds__(HW_THREADS, 2)
k_stream(DevModel m, DevHP hp, DevCsr csr, int row_begin, int row_end, int scatter_user,
         int scatter_item, float *pred_out, int *err_flag) {
  __shared__ HwSmem<LANES, VEC, EXACT_DOT> sm;
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
        // phase B: the tile's feature window
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
  g.dot_s = EXACT_DOT ? &sm.dot[(warp * GPW + gw) * Group<LANES, VEC>::DOT_FLOATS] : nullptr;
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
    const bool staged = mt.staged != 0;
    const int gbase = csr.val_base;

    auto IDX = [&](int f) -> unsigned { return staged ? st.idx[f - mt.sm_base] : csr.index[f - gbase]; };
    auto VAL = [&](int f) -> float { return staged ? st.val[f - mt.sm_base] : csr.value[f - gbase]; };

    // ---- issue the gathers of instance q ----
    auto pre_load = [&](Pre<VEC> &p, int q) {
      p.q = q;
      const int rp0 = rp[3 * q], rp1 = rp[3 * q + 1], rp2 = rp[3 * q + 2], rp3 = rp[3 * q + 3];
      if (rp1 == rp0 && rp2 == rp1 + 1 && rp3 == rp2 + 1) {
        p.uid = IDX(rp1);
        p.iid = IDX(rp2);
        p.uval = VAL(rp1);
        p.ival = VAL(rp2);
        p.label = lab[q];
        if (p.uid >= (unsigned)m.num_user || p.iid >= (unsigned)m.num_item) {
          p.kind = PRE_BAD;
          return;
        }
        p.kind = PRE_SIMPLE;
        g.load_row(m, (size_t)m.user_off + p.uid, p.wu);
        g.load_row(m, (size_t)m.item_off + p.iid, p.wi);
        p.ub = m.no_user_bias ? 0.0f : __ldcg(m.bias + m.user_off + p.uid);
        p.ib = __ldcg(m.bias + m.item_off + p.iid);
      } else {
        p.kind = PRE_GENERIC;
      }
    };

    // ---- compute instance p (gathers already in registers when simple) ----
    auto compute = [&](const Pre<VEC> &p) {
      if (p.kind == PRE_SIMPLE) {
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
      } else if (p.kind == PRE_GENERIC) {
        const int q = p.q;
        const unsigned *idx = staged ? (st.idx - mt.sm_base) : (csr.index - gbase);
        const float *val = staged ? (st.val - mt.sm_base) : (csr.value - gbase);
        const float pr = generic_instance<LANES, VEC, EXACT_DOT, TRAIN>(
            g, m, hp, rp[3 * q], rp[3 * q + 1], rp[3 * q + 2], rp[3 * q + 3], lab[q], idx, val,
            scatter_user, scatter_item, err_flag);
        if (!TRAIN && g.gl == 0) pred_out[mt.r0 + q - row_begin] = pr;
      } else if (p.kind == PRE_BAD) {
        if (g.gl == 0)
          atomicCAS(err_flag, 0, p.uid >= (unsigned)m.num_user ? ERR_USER_INDEX : ERR_ITEM_INDEX);
      }
    };

    Pre<VEC> A, B;
    int q = gid;
    A.kind = PRE_NONE;
    if (q < mt.nrow) pre_load(A, q);
    while (q < mt.nrow) {
      int qn = q + NGROUP;
      B.kind = PRE_NONE;
      if (qn < mt.nrow) pre_load(B, qn);
      compute(A);
      q = qn;
      if (q >= mt.nrow) break;
      qn = q + NGROUP;
      A.kind = PRE_NONE;
      if (qn < mt.nrow) pre_load(A, qn);
      compute(B);
      q = qn;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.empty[s]);
  }
}

template <int L, int V>
static int launch_geo(svdgpu *h, const DevCsr &csr, int r0, int r1, bool train, float *pred) {
  const long long ntile = ((long long)(r1 - r0) + HW_TILE - 1) / HW_TILE;
  int grid = 1;
#define GO(ED, TR)                                                                           \
  {                                                                                          \
    auto k = k_stream<L, V, ED, TR>;                                                         \
    if (grid_for(h, k, HW_THREADS, ntile, &grid)) return 1;                                  \
    k<<<grid, HW_THREADS, 0, h->stream>>>(h->dm, h->dhp, csr, r0, r1, h->scatter_user,       \
                                          h->scatter_item, pred, h->d_err);                  \
  }
  if (train) {
    if (h->exact_dot) GO(true, true) else GO(false, true)
  } else {
    GO(true, false)
  }
#undef GO
  CU(h, cudaGetLastError());
  h->n_launch++;
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
