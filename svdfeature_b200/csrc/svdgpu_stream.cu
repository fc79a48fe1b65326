// svdgpu_stream.cu -- k_stream: Hogwild training / prediction over a CSR batch.
//
// Persistent CTAs (256 threads = 8 warps, 2 CTAs per SM).  EVERY warp owns a private
// staging pipeline: it takes whole 64-instance tiles (tile w, w+W, w+2W, ... for warp slot w
// of W), and its lane 0 stages them into the warp's slice of shared memory with 1-D bulk
// asynchronous copies (cp.async.bulk = TMA unit, SASS UBLKCP; completion on the warp's own
// mbarriers): phase A brings the tile's row_ptr/label window, phase B -- once A has landed
// and the feature range is known -- its index/value window.  Four tiles per warp are in
// flight (A of tile j+3 and B of tile j+2 are issued before tile j is computed; tile j+1 has
// landed and its user rows are being pulled into L2 with prefetch.global.L2), so nobody
// chases row_ptr -> index -> row through DRAM and no warp ever waits for another warp
// (the first version shared tiles between warps through a producer warp; ncu showed the
// consumers stalling on the shared tile barriers).
//
// Inside a warp one lane GROUP (svdgpu_device.cuh) runs one instance.
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

constexpr int HW_TILE = 64;            // instances per tile (one tile belongs to ONE warp)
constexpr int HW_STAGES = 4;           // tiles in flight per warp
constexpr int HW_CAP = 4 * HW_TILE;    // staged index/value entries per tile
constexpr int HW_WARPS = 8;            // warps per CTA, every one a consumer with its own pipeline
constexpr int HW_THREADS = HW_WARPS * 32;

struct __align__(16) HwStage {
  int rp[3 * HW_TILE + 8];
  float label[HW_TILE + 4];
  unsigned idx[HW_CAP + 8];
  float val[HW_CAP + 8];
};
struct __align__(16) HwWarp {
  HwStage st[HW_STAGES];
  uint64_t barA[HW_STAGES], barB[HW_STAGES];
};

template <int LANES, int VEC, bool EXACT_DOT>
constexpr size_t hw_smem_bytes() {
  return sizeof(HwWarp) * HW_WARPS +
         (EXACT_DOT ? sizeof(float) * HW_WARPS * (32 / LANES) * Group<LANES, VEC>::DOT_FLOATS : 0);
}

// gathers of one instance, issued ahead of its compute
template <int VEC>
struct Pre {
  float4 wu[VEC], wi[VEC];
  float ub, ib;
  int q;  // row inside the tile, -1: nothing loaded
};

__device__ __forceinline__ bool is_simple(const int *rp, int q) {
  const int rp0 = rp[3 * q], rp1 = rp[3 * q + 1], rp2 = rp[3 * q + 2], rp3 = rp[3 * q + 3];
  return rp1 == rp0 && rp2 == rp1 + 1 && rp3 == rp2 + 1;
}

template <int LANES, int VEC, bool EXACT_DOT, bool TRAIN, bool GENERIC>
__global__ void __launch_bounds__(HW_THREADS, 2)
k_stream(DevModel m, DevHP hp, DevCsr csr, int row_begin, int row_end, int scatter_user,
         int scatter_item, float *pred_out, int *tile_flag, int *err_flag) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  HwWarp &sw = reinterpret_cast<HwWarp *>(smem_raw)[warp];
  float *dot_base = reinterpret_cast<float *>(smem_raw + sizeof(HwWarp) * HW_WARPS);

  const int ntile = (row_end - row_begin + HW_TILE - 1) / HW_TILE;
  const int wslot = blockIdx.x * HW_WARPS + warp;  // this warp's first tile
  const int wstride = gridDim.x * HW_WARPS;
  const int nlocal = ntile > wslot ? (ntile - wslot + wstride - 1) / wstride : 0;

  if (lane == 0) {
    for (int s = 0; s < HW_STAGES; ++s) {
      mbar_init(&sw.barA[s], 1);
      mbar_init(&sw.barB[s], 1);
    }
    mbar_fence_init();
  }
  __syncwarp();

  constexpr int GPW = 32 / LANES;  // groups per warp
  Group<LANES, VEC> g;
  g.gl = lane % LANES;
  const int gw = lane / LANES;
  g.gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (gw * LANES));
  g.dot_s = EXACT_DOT ? dot_base + (warp * GPW + gw) * Group<LANES, VEC>::DOT_FLOATS : nullptr;

  // ---- the warp's private staging pipeline ---------------------------------
  // phase parity of every stage barrier (bit s = stage s); a barrier is armed once per
  // non-skipped tile and awaited exactly once, so the bits follow the barrier phases
  unsigned ph_a = 0, ph_b = 0;
  auto tile_of = [&](int j) { return wslot + j * wstride; };
  auto skip_tile = [&](int j) -> bool { return GENERIC && tile_flag[tile_of(j)] == 0; };
  // phase A: row_ptr[3*r0 .. 3*(r0+nrow)] and label[r0 .. r0+nrow), 16-byte aligned windows
  auto issue_a = [&](int j) {
    if (j >= nlocal || skip_tile(j)) return;
    __syncwarp();  // every lane is done reading the stage being refilled
    if (lane == 0) {
      HwStage &st = sw.st[j % HW_STAGES];
      const int r0 = row_begin + tile_of(j) * HW_TILE;
      const int nrow = min(HW_TILE, row_end - r0);
      const int a_off = (3 * r0) & 3, l_off = r0 & 3;
      const unsigned bytesA = (unsigned)((a_off + 3 * nrow + 1 + 3) & ~3) * 4u;
      const unsigned bytesL = (unsigned)((l_off + nrow + 3) & ~3) * 4u;
      mbar_arrive_expect_tx(&sw.barA[j % HW_STAGES], bytesA + bytesL);
      bulk_g2s(st.rp, csr.row_ptr + (3 * r0 - a_off), bytesA, &sw.barA[j % HW_STAGES]);
      bulk_g2s(st.label, csr.label + (r0 - l_off), bytesL, &sw.barA[j % HW_STAGES]);
    }
  };
  // the tile's feature window [v0, v1) (positions relative to csr.val_base), known once A landed
  struct Win {
    int v0, v1, v_off, nel, staged;
  };
  auto window = [&](int j) -> Win {
    const HwStage &st = sw.st[j % HW_STAGES];
    const int r0 = row_begin + tile_of(j) * HW_TILE;
    const int nrow = min(HW_TILE, row_end - r0);
    const int a_off = (3 * r0) & 3;
    Win w;
    w.v0 = st.rp[a_off] - csr.val_base;
    w.v1 = st.rp[a_off + 3 * nrow] - csr.val_base;
    w.v_off = w.v0 & 3;
    w.nel = (w.v_off + (w.v1 - w.v0) + 3) & ~3;
    w.staged = (w.v0 >= 0 && w.v1 >= w.v0 && w.v1 + csr.val_base <= csr.val_end && w.nel <= HW_CAP + 8) ? 1 : 0;
    return w;
  };
  // phase B: the tile's index/value window
  auto issue_b = [&](int j) {
    if (j >= nlocal || skip_tile(j)) return;
    mbar_wait(&sw.barA[j % HW_STAGES], (ph_a >> (j % HW_STAGES)) & 1u);
    ph_a ^= 1u << (j % HW_STAGES);
    const Win w = window(j);
    if (lane == 0 && w.staged && w.nel > 0) {
      HwStage &st = sw.st[j % HW_STAGES];
      mbar_arrive_expect_tx(&sw.barB[j % HW_STAGES], 2u * (unsigned)w.nel * 4u);
      bulk_g2s(st.idx, csr.index + (w.v0 - w.v_off), (unsigned)w.nel * 4u, &sw.barB[j % HW_STAGES]);
      bulk_g2s(st.val, csr.value + (w.v0 - w.v_off), (unsigned)w.nel * 4u, &sw.barB[j % HW_STAGES]);
    }
  };

  // wait for phase B of tile j (once per tile) ...
  auto wait_b = [&](int j) {
    if (j >= nlocal || skip_tile(j)) return;
    const Win w = window(j);
    if (w.staged && w.nel > 0) {
      mbar_wait(&sw.barB[j % HW_STAGES], (ph_b >> (j % HW_STAGES)) & 1u);
      ph_b ^= 1u << (j % HW_STAGES);
    }
  };
  // ... and pull the user rows (and user biases) of that tile into L2 a whole tile ahead of
  // their use: ~1/3 of the user-row gathers of configs[1] miss L2, and a DRAM miss is longer
  // than the one-instance register prefetch covers.  Item rows are L2-resident already.
  auto l2_prefetch_tile = [&](int j) {
    if (GENERIC || j >= nlocal) return;
    const Win w = window(j);
    if (!w.staged) return;
    const HwStage &st = sw.st[j % HW_STAGES];
    const int r0 = row_begin + tile_of(j) * HW_TILE;
    const int nrow = min(HW_TILE, row_end - r0);
    const int *rp = st.rp + ((3 * r0) & 3);
    const int sm_base = w.v0 - w.v_off + csr.val_base, v_hi = w.v1 + csr.val_base;
    const int row_bytes = m.pitch * 4;
    for (int q = lane; q < nrow; q += 32) {
      const int f = rp[3 * q + 1];
      if (f < sm_base || f >= v_hi) continue;
      const unsigned uid = st.idx[f - sm_base];
      if (uid >= (unsigned)m.num_user) continue;
      const char *row = reinterpret_cast<const char *>(m.W + ((size_t)m.user_off + uid) * (size_t)m.pitch);
      for (int b = 0; b < row_bytes; b += 128) prefetch_l2(row + b);
      if (!m.no_user_bias) prefetch_l2(m.bias + m.user_off + uid);
    }
  };

  issue_a(0);
  issue_a(1);
  issue_a(2);
  issue_b(0);
  issue_b(1);
  wait_b(0);
  l2_prefetch_tile(0);
  for (int j = 0; j < nlocal; ++j) {
    issue_a(j + 3);
    issue_b(j + 2);
    wait_b(j + 1);
    l2_prefetch_tile(j + 1);
    if (skip_tile(j)) continue;
    const HwStage &st = sw.st[j % HW_STAGES];
    const int t = tile_of(j);
    const int r0 = row_begin + t * HW_TILE;
    const int nrow = min(HW_TILE, row_end - r0);
    const int *rp = st.rp + ((3 * r0) & 3);
    const float *lab = st.label + (r0 & 3);
    const Win w = window(j);  // A(j) and B(j) were awaited one tile ago
    const int sm_base = w.v0 - w.v_off + csr.val_base;  // absolute feature position held by idx[0]
    const int v_hi = w.v1 + csr.val_base;

    if (GENERIC) {
      // ---- pass 2: whatever pass 1 left in this (flagged) tile ----
      const unsigned *idx = w.staged ? (st.idx - sm_base) : (csr.index - csr.val_base);
      const float *val = w.staged ? (st.val - sm_base) : (csr.value - csr.val_base);
      for (int q = gw; q < nrow; q += GPW) {
        // done by pass 1 (which runs only for the plain L2-decay regulariser)
        if (hp.plain && w.staged && is_simple(rp, q) && rp[3 * q] >= sm_base && rp[3 * q + 3] <= v_hi) continue;
        if (!row_ok(rp[3 * q], rp[3 * q + 1], rp[3 * q + 2], rp[3 * q + 3], csr.val_base, csr.val_end)) {
          if (g.gl == 0) atomicCAS(err_flag, 0, ERR_ROW_PTR);
          continue;
        }
        const float pr = process_instance<LANES, VEC, EXACT_DOT, TRAIN, false>(
            g, m, hp, rp[3 * q], rp[3 * q + 1], rp[3 * q + 2], rp[3 * q + 3], lab[q], idx, val,
            scatter_user, scatter_item, nullptr, err_flag);
        if (!TRAIN && g.gl == 0) pred_out[r0 + q - row_begin] = pr;
      }
    } else if (!w.staged) {
      // window did not fit (or is malformed): the whole tile goes to pass 2
      if (lane == 0) tile_flag[t] = 1;
    } else {
      // ---- pass 1: straight-line basic-MF rows, gathers one instance ahead ----
      const unsigned *sidx = st.idx - sm_base;
      const float *sval = st.val - sm_base;
      bool other = false;  // this group met a row of another shape

      auto pre_load = [&](Pre<VEC> &p, int q) {
        p.q = -1;
        if (q >= nrow) return;
        const int f = rp[3 * q + 1];
        if (!is_simple(rp, q) || f < sm_base || f + 2 > v_hi) {
          other = true;
          return;
        }
        const unsigned uid = sidx[f], iid = sidx[f + 1];
        if (uid >= (unsigned)m.num_user || iid >= (unsigned)m.num_item) {
          if (g.gl == 0) atomicCAS(err_flag, 0, uid >= (unsigned)m.num_user ? ERR_USER_INDEX : ERR_ITEM_INDEX);
          return;
        }
        p.q = q;
        g.load_row(m, (size_t)m.user_off + uid, p.wu);
        g.load_row(m, (size_t)m.item_off + iid, p.wi);
        p.ub = m.no_user_bias ? 0.0f : __ldcg(m.bias + m.user_off + uid);
        p.ib = __ldcg(m.bias + m.item_off + iid);
      };

      auto compute = [&](const Pre<VEC> &p) {
        if (p.q < 0) return;
        // the row's scalars come back from the staged tile (cheaper than carrying them)
        const int f = rp[3 * p.q + 1];
        const unsigned uid = sidx[f], iid = sidx[f + 1];
        const float uval = sval[f], ival = sval[f + 1];
        // prepare_tmp (base.h:354-381): tmp = 0 + w*val
        float4 tu[VEC], ti[VEC];
        const bool one_u = scalar_is_one(uval), one_i = scalar_is_one(ival);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          tu[v] = f4_add_scaled(f4_zero(), p.wu[v], uval, one_u);
          ti[v] = f4_add_scaled(f4_zero(), p.wi[v], ival, one_i);
        }
        // calc_bias (base.h:313-353) + pred (base.h:445-454)
        double bsum = 0.0;
        if (!m.no_user_bias) bsum = __dadd_rn(bsum, (double)__fmul_rn(uval, p.ub));
        bsum = __dadd_rn(bsum, (double)__fmul_rn(ival, p.ib));
        const float d = g.template dot<EXACT_DOT>(m, tu, ti);
        double sum = __dadd_rn((double)hp.base_score, bsum);
        sum = __dadd_rn(sum, (double)d);
        const float pred = map_active((float)sum, m.active_type);
        if (!TRAIN) {
          if (g.gl == 0) pred_out[r0 + p.q - row_begin] = pred;
          return;
        }
        // update_no_decay + regularize(after), fused (base.h:383-427, 211-283)
        const float err = cal_grad(lab[p.q], pred, m.active_type);
        const float lrerr = __fmul_rn(hp.lr, err);
        const float su = __fmul_rn(lrerr, uval), si = __fmul_rn(lrerr, ival);
        const bool one_su = scalar_is_one(su), one_si = scalar_is_one(si);
        float4 nw[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          nw[v] = f4_add_scaled(p.wu[v], ti[v], su, one_su);
          if (!hp.du_skip) nw[v] = f4_scale(nw[v], hp.du);
        }
        if (scatter_user == SCATTER_RED) g.red_row(m, (size_t)m.user_off + uid, nw, p.wu);
        else g.store_row(m, (size_t)m.user_off + uid, nw);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          nw[v] = f4_add_scaled(p.wi[v], tu[v], si, one_si);
          if (!hp.di_skip) nw[v] = f4_scale(nw[v], hp.di);
        }
        if (scatter_item == SCATTER_RED) g.red_row(m, (size_t)m.item_off + iid, nw, p.wi);
        else g.store_row(m, (size_t)m.item_off + iid, nw);
        if (g.gl == 0 && !m.no_user_bias) {
          float *bp = m.bias + m.user_off + uid;
          const float nb = __fmul_rn(__fadd_rn(p.ub, su), hp.dub);
          if (scatter_user == SCATTER_RED) red1(bp, __fsub_rn(nb, p.ub));
          else __stcg(bp, nb);
        }
        if (g.gl == 1) {
          float *bp = m.bias + m.item_off + iid;
          const float nb = __fmul_rn(__fadd_rn(p.ib, si), hp.dib);
          if (scatter_item == SCATTER_RED) red1(bp, __fsub_rn(nb, p.ib));
          else __stcg(bp, nb);
        }
      };

      Pre<VEC> A, B;
      int q = gw;
      pre_load(A, q);
      while (q < nrow) {
        pre_load(B, q + GPW);
        compute(A);
        q += GPW;
        if (q >= nrow) break;
        pre_load(A, q + GPW);
        compute(B);
        q += GPW;
      }
      if (__any_sync(0xffffffffu, other) && lane == 0) tile_flag[t] = 1;
    }
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
  // the straight-line pass implements the plain L2-decay regulariser only: any other
  // reg_method / reg_global / user_nonnegative sends every tile to the generic pass
  const bool pass1 = h->dhp.plain != 0;
  CU(h, cudaMemsetAsync(h->d_tile_flag, pass1 ? 0 : 1, (size_t)ntile * sizeof(int), h->stream));
  int grid = 1;
#define GO(ED, TR, GEN)                                                                          \
  {                                                                                              \
    auto k = k_stream<L, V, ED, TR, GEN>;                                                        \
    const size_t smem = hw_smem_bytes<L, V, ED>();                                               \
    CU(h, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
    if (grid_for(h, k, HW_THREADS, (ntile + HW_WARPS - 1) / HW_WARPS, &grid, smem)) return 1;    \
    k<<<grid, HW_THREADS, smem, h->stream>>>(h->dm, h->dhp, csr, r0, r1, h->scatter_user,        \
                                             h->scatter_item, pred, h->d_tile_flag, h->d_err);   \
    h->n_launch++;                                                                               \
  }
  if (train) {
    if (h->exact_dot) { if (pass1) GO(true, true, false) GO(true, true, true) }
    else { if (pass1) GO(false, true, false) GO(false, true, true) }
  } else {
    if (pass1) GO(true, false, false) GO(true, false, true)
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
