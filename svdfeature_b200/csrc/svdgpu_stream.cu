// svdgpu_stream.cu -- Hogwild training / prediction over a CSR batch: the launcher of the two
// passes and k_stream, the generic pass.
//
// Two passes over a batch (Hogwild has no order to keep):
//   pass 1  k_mf (svdgpu_mf.cu): rows of the basic-MF shape -- no global feature, one user
//           feature, one item feature (configs[0..1]) -- on a straight-line path fed by a
//           shared-memory ring of asynchronously gathered rows.  Every row it does not take
//           is recorded in a bit mask (one word per 32 rows).
//   pass 2  k_stream: visits the 64-row tiles whose masks are not zero and runs the marked rows
//           through the generic process_instance().  For configs[1] it finds nothing.
// Both passes perform the same per-instance arithmetic (svdgpu_device.cuh).
//
// k_stream: persistent CTAs (256 threads = 8 warps, 2 CTAs per SM).  EVERY warp owns a private
// staging pipeline: it takes whole 64-instance tiles (tile w, w+W, w+2W, ... for warp slot w
// of W), and its lane 0 stages them into the warp's slice of shared memory with 1-D bulk
// asynchronous copies (cp.async.bulk = TMA unit, SASS UBLKCP; completion on the warp's own
// mbarriers): phase A brings the tile's row_ptr/label window, phase B -- once A has landed
// and the feature range is known -- its index/value window.  Four tiles per warp are in
// flight, so nobody chases row_ptr -> index -> row through DRAM and no warp ever waits for
// another warp.  Inside a warp one lane GROUP (svdgpu_device.cuh) runs one instance.
#include "svdgpu_internal.h"

namespace svdk {

constexpr int HW_TILE = 64;            // most instances per tile (one tile belongs to ONE warp); the launch picks
                                       // 64, 32, 16 or 8 rows per tile so that a tile's index/value window fits
                                       // HW_CAP entries (rows with ~10 features, configs[4]: 16 rows per tile --
                                       // an unstaged tile reads its features from global memory on the critical
                                       // path of every instance: 0.39 G inst/s there instead of ...)
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

template <int LANES, int VEC, bool EXACT_DOT, bool TRAIN>
__global__ void __launch_bounds__(HW_THREADS, 2)
k_stream(DevModel m, DevHP hp, DevCsr csr, int row_begin, int row_end, int scatter_user,
         int scatter_item, float *pred_out, const unsigned *row_mask, const unsigned *any_left,
         int *err_flag, int l2_ahead, int tile_rows) {
  if (*any_left == 0u) return;  // pass 1 took every row
  const int T = tile_rows;  // rows per tile: 64, 32, 16 or 8
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  HwWarp &sw = reinterpret_cast<HwWarp *>(smem_raw)[warp];
  float *dot_base = reinterpret_cast<float *>(smem_raw + sizeof(HwWarp) * HW_WARPS);

  const int ntile = (row_end - row_begin + T - 1) / T;
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
  // rows left by pass 1: one mask bit per row (bit r - row_begin), T bits per tile
  auto tile_mask = [&](int t) -> unsigned long long {
    const long long bit0 = (long long)t * T;
    const unsigned *w = row_mask + (bit0 >> 5);
    if (T == 64) return (unsigned long long)w[0] | ((unsigned long long)w[1] << 32);
    return (unsigned long long)((w[0] >> (bit0 & 31)) & (T == 32 ? 0xffffffffu : ((1u << T) - 1u)));
  };
  auto skip_tile = [&](int j) -> bool { return tile_mask(tile_of(j)) == 0ULL; };
  // phase A: row_ptr[3*r0 .. 3*(r0+nrow)] and label[r0 .. r0+nrow), 16-byte aligned windows
  auto issue_a = [&](int j) {
    if (j >= nlocal || skip_tile(j)) return;
    __syncwarp();  // every lane is done reading the stage being refilled
    if (lane == 0) {
      HwStage &st = sw.st[j % HW_STAGES];
      const int r0 = row_begin + tile_of(j) * T;
      const int nrow = min(T, row_end - r0);
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
    const int r0 = row_begin + tile_of(j) * T;
    const int nrow = min(T, row_end - r0);
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

  // wait for phase B of tile j (once per tile)
  auto wait_b = [&](int j) {
    if (j >= nlocal || skip_tile(j)) return;
    const Win w = window(j);
    if (w.staged && w.nel > 0) {
      mbar_wait(&sw.barB[j % HW_STAGES], (ph_b >> (j % HW_STAGES)) & 1u);
      ph_b ^= 1u << (j % HW_STAGES);
    }
  };
  issue_a(0);
  issue_a(1);
  issue_a(2);
  issue_b(0);
  issue_b(1);
  wait_b(0);
  for (int j = 0; j < nlocal; ++j) {
    issue_a(j + 3);
    issue_b(j + 2);
    wait_b(j + 1);
    if (skip_tile(j)) continue;
    const HwStage &st = sw.st[j % HW_STAGES];
    const int t = tile_of(j);
    const int r0 = row_begin + t * T;
    const int nrow = min(T, row_end - r0);
    const int *rp = st.rp + ((3 * r0) & 3);
    const float *lab = st.label + (r0 & 3);
    const Win w = window(j);  // A(j) and B(j) were awaited one tile ago
    const int sm_base = w.v0 - w.v_off + csr.val_base;  // absolute feature position held by idx[0]
    const int v_hi = w.v1 + csr.val_base;

    const unsigned long long mk = tile_mask(t);
    const unsigned *idx = w.staged ? (st.idx - sm_base) : (csr.index - csr.val_base);
    const float *val = w.staged ? (st.val - sm_base) : (csr.value - csr.val_base);
    const float *val2 = csr.value2 ? csr.value2 - csr.val_base : nullptr;  // (not staged: side features only)
    // Pull the model rows (and biases) of the instance this group runs L2_AHEAD turns from now into
    // L2.  For models far larger than the 126 MB L2 (configs[3..4]: 0.7-11 GB) every gather of the
    // generic routine is otherwise a DRAM round trip on the instance's critical path.  (A whole
    // tile ahead is too far: 16 warps x 64 rows x 1.5 KB per SM do not survive in L2.)
    constexpr int L2_AHEAD = 2;
    auto l2_prefetch_row = [&](int q) {
      if (!l2_ahead || q >= nrow || !w.staged) return;
      if (((mk >> q) & 1ULL) == 0ULL) return;
      const int rp0 = rp[3 * q], rp1 = rp[3 * q + 1], rp2 = rp[3 * q + 2], rp3 = rp[3 * q + 3];
      if (!(rp0 >= sm_base && rp0 <= rp1 && rp1 <= rp2 && rp2 <= rp3 && rp3 <= v_hi)) return;
      const int row_bytes = m.pitch * 4;
      for (int f = rp0 + g.gl; f < rp3; f += LANES) {
        const unsigned id = st.idx[f - sm_base];
        if (f < rp1) {
          if (id < (unsigned)m.num_global) prefetch_l2(m.g_bias + id);
          continue;
        }
        const bool is_user = f < rp2;
        if (id >= (unsigned)(is_user ? m.num_user : m.num_item)) continue;
        const size_t row = (size_t)(is_user ? m.user_off : m.item_off) + id;
        const char *p = reinterpret_cast<const char *>(m.W + row * (size_t)m.pitch);
        for (int b = 0; b < row_bytes; b += 128) prefetch_l2(p + b);
        prefetch_l2(m.bias + row);
      }
    };
    for (int a = 0; a < L2_AHEAD; ++a) l2_prefetch_row(gw + a * GPW);
    for (int q = gw; q < nrow; q += GPW) {
      l2_prefetch_row(q + L2_AHEAD * GPW);
      if (((mk >> q) & 1ULL) == 0ULL) continue;  // done by pass 1
      if (!row_ok(rp[3 * q], rp[3 * q + 1], rp[3 * q + 2], rp[3 * q + 3], csr.val_base, csr.val_end)) {
        if (g.gl == 0) atomicCAS(err_flag, 0, ERR_ROW_PTR);
        continue;
      }
      // a staged window holds the whole tile or nothing, so every row of it is addressable
      const float pr = process_instance<LANES, VEC, EXACT_DOT, TRAIN, false>(
          g, m, hp, rp[3 * q], rp[3 * q + 1], rp[3 * q + 2], rp[3 * q + 3], lab[q], idx, val,
          scatter_user, scatter_item, nullptr, err_flag, val2);
      if (!TRAIN && g.gl == 0) pred_out[r0 + q - row_begin] = pr;
    }
  }
}

template <int L, int V>
static int launch_geo(svdgpu *h, const DevCsr &csr, int r0, int r1, bool train, float *pred) {
  // rows per tile: the largest of 64 / 32 / 16 / 8 whose average feature window (+25 %) fits the stage
  int tile_rows = h->stream_tile;
  if (tile_rows != 64 && tile_rows != 32 && tile_rows != 16 && tile_rows != 8) {
    const double avg = csr.avg_nnz > 0.0f ? (double)csr.avg_nnz
                                          : (double)(csr.val_end - csr.val_base) / (double)std::max(1, r1 - r0);
    tile_rows = 64;
    while (tile_rows > 8 && tile_rows * avg * 1.25 > HW_CAP) tile_rows >>= 1;
  }
  const long long ntile = ((long long)(r1 - r0) + tile_rows - 1) / tile_rows;
  int grid = 1;
  // tile-ahead L2 prefetch of the gathered rows when the model cannot live in L2 (option "l2_ahead")
  const size_t model_bytes = h->rows * (size_t)h->dm.pitch * sizeof(float);
  const int l2_ahead = h->l2_ahead >= 0 ? h->l2_ahead : (model_bytes > (size_t)(96u << 20) ? 1 : 0);
#define GO(ED, TR)                                                                               \
  {                                                                                              \
    auto k = k_stream<L, V, ED, TR>;                                                             \
    const size_t smem = hw_smem_bytes<L, V, ED>();                                               \
    CU(h, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
    if (grid_for(h, k, HW_THREADS, (ntile + HW_WARPS - 1) / HW_WARPS, &grid, smem)) return 1;    \
    k<<<grid, HW_THREADS, smem, h->stream>>>(h->dm, h->dhp, csr, r0, r1, h->scatter_user,        \
                                             h->scatter_item, pred, h->d_row_mask,               \
                                             h->d_row_mask + h->flag_for_generic, h->d_err,      \
                                             l2_ahead, tile_rows);                               \
    h->n_launch++;                                                                               \
  }
  if (train) {
    if (h->exact_dot) GO(true, true) else GO(false, true)
  } else {
    GO(true, false)
  }
#undef GO
  CU(h, cudaGetLastError());
  return 0;
}

int launch_stream(svdgpu *h, const Geometry &g, const DevCsr &csr, int r0, int r1, bool train,
                  float *pred) {
  // row masks: one word per 32 rows (+ padding so that the generic pass may read them in pairs); the
  // last three words are the "left something" flags of the three fast passes
  const long long nmask = (((long long)(r1 - r0) + 31) / 32 + 5) & ~1LL;
  h->any_left_at = (size_t)nmask - 1;
  if ((size_t)nmask * sizeof(unsigned) > h->row_mask_cap) {
    if (h->d_row_mask) CU(h, cudaFree(h->d_row_mask));
    h->d_row_mask = nullptr;
    h->row_mask_cap = 0;
    const size_t cap = (size_t)nmask * sizeof(unsigned) * 2;
    CU(h, cudaMalloc(&h->d_row_mask, cap));
    h->row_mask_cap = cap;
  }
  // the straight-line pass implements the plain L2-decay regulariser only: any other
  // reg_method / reg_global / user_nonnegative sends every row to the generic pass
  const bool pass1 = h->dhp.plain != 0 && h->pass1 != 0;
  CU(h, cudaMemsetAsync(h->d_row_mask, pass1 ? 0 : 0xff, (size_t)nmask * sizeof(unsigned), h->stream));
  unsigned *flags = h->d_row_mask + h->any_left_at - 2;  // [0] first, [1] second, [2] third pass
  size_t last = h->any_left_at - 2;
  if (pass1 && launch_mf(h, g, csr, r0, r1, train, pred, 0, flags, flags)) return 1;
  // second fast pass: rows with two item features (pairwise-rank rows) among those left
  const bool pass1b = pass1 && h->pass1 >= 2;
  if (pass1b) {
    if (launch_mf(h, g, csr, r0, r1, train, pred, 1, flags + 1, flags)) return 1;
    last = h->any_left_at - 1;
  }
  // third fast pass: basic rows with a few global features (neighbourhood rows) among those left
  const bool pass1c = pass1 && h->pass1 >= 3 && h->shape.num_global > 0;
  if (pass1c) {
    if (launch_mf(h, g, csr, r0, r1, train, pred, 2, flags + 2, h->d_row_mask + last)) return 1;
    last = h->any_left_at;
  }
  h->flag_for_generic = last;
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
