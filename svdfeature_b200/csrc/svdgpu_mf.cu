// svdgpu_mf.cu -- k_mf: the Hogwild fast pass over rows of the basic-MF shape.
//
// A row of this shape has no global feature, one user feature and one item feature
// (BASELINE configs[0..1]; SVDFeature::update_inner with ng=0, nu=1, ni=1, base.h:456-462).
// Rows of any other shape, rows whose tile did not fit the staging window and rows with an
// out-of-range index are left to the generic pass (k_stream): the fast pass records them in
// a per-tile bit mask, one bit per row.
//
// Persistent CTAs, every warp works alone (no block-wide barrier in the loop):
//   * CSR staging: the warp takes whole 32-row tiles (tile w, w+W, ...) and its lane 0 stages
//     them with 1-D bulk copies (cp.async.bulk = TMA unit, SASS UBLKCP; completion on the
//     warp's own mbarriers): phase A = the row_ptr/label window, phase B = the index/value
//     window once A has landed; four tiles in flight.
//   * row gathers: a ring of DEPTH iterations per warp in shared memory.  One iteration =
//     one instance per lane group; its user row, item row and the two 16-byte windows
//     holding its biases are fetched with cp.async.cg (SASS LDGSTS, L2-coherent, no register
//     destination) DEPTH iterations before they are used, so every warp keeps
//     DEPTH x (32/LANES) x 2 rows in flight without holding them in registers.  (The first
//     version held ONE prefetched instance in registers: ncu showed the warps waiting for
//     L2 -- long_scoreboard -- with DRAM 11 % busy.)
//   * compute: rows come back from shared memory (LDS.128), the update is formed in
//     registers and leaves as red.global.add.v4.f32 of (new - old) (or plain stores).
//
// Arithmetic is that of process_instance() (svdgpu_device.cuh), specialised: with
// EXACT_DOT the result is bit-identical to the generic routine; without it the dot uses the
// shuffle-tree order and the "0 +" of prepare_tmp (base.h:357,372) is dropped (it can only
// change the sign of a zero).
//
// NG > 0 (the third fast pass): rows (1..NG | 1 | 1) -- basic rows with a few global features, the
// neighbourhood rows of configs[4] -- among the rows the earlier passes left.  Three staged tiles
// per warp instead of two, so that the index/value window of the tile being computed is still
// in shared memory: lane j of a group reads global feature j from it, fetches the 16-byte window
// of g_bias holding its bias into the ring slot with the rows, and later updates that bias.
#include "svdgpu_internal.h"

namespace svdk {

constexpr int MF_TILE = 32;            // rows per tile = bits of one row mask
constexpr int MF_STAGES = 2;           // staged tiles per warp (a decoded tile lives in registers)
#ifndef MF_CAP_PER_ROW
#define MF_CAP_PER_ROW 3
#endif
constexpr int MF_CAP = MF_CAP_PER_ROW * MF_TILE;    // staged index/value entries per tile (up to 3 per row)
#ifndef MF_WARPS_PER_CTA
#define MF_WARPS_PER_CTA 8
#endif
constexpr int MF_WARPS = MF_WARPS_PER_CTA;
constexpr int MF_THREADS = MF_WARPS * 32;

// staged entries per tile / staged tiles per warp of a variant (NG = most global features per row it takes)
// LATEB (NG > 0 only): two stages; the index/value window of tile j+2 is requested only after tile j
// has been computed (it lands in the stage tile j occupied), which leaves its latency exposed once
// per tile but costs a third less shared memory -- two CTAs per SM instead of one.
template <int NG, bool LATEB> struct MfCfg {
  static constexpr int CAP = NG > 0 ? (LATEB ? 14 * MF_TILE : (NG + 2) * MF_TILE) : MF_CAP;
  static constexpr int STAGES = (NG > 0 && !LATEB) ? 3 : MF_STAGES;
};

template <int CAP>
struct __align__(16) MfStageT {
  int rp[3 * MF_TILE + 8];
  float label[MF_TILE + 4];
  unsigned idx[CAP + 8];
  float val[CAP + 8];
};
template <int CAP, int STAGES>
struct __align__(16) MfWarpT {
  MfStageT<CAP> st[STAGES];
  uint64_t barA[STAGES], barB[STAGES];
};

template <int LANES, int VEC, int DEPTH, int NI, int NG>
struct MfRing {
  static constexpr int GPW = 32 / LANES;
  static constexpr int NCH = LANES * VEC;
  // per (iteration slot, group): user row | NI item rows | user-bias window | NI item-bias windows |
  // NG global-bias windows
  static constexpr int SLOT_F4 = (1 + NI) * NCH + (1 + NI) + NG;
  static constexpr size_t BYTES = (size_t)DEPTH * GPW * SLOT_F4 * sizeof(float4);
};

template <int LANES, int VEC, bool EXACT_DOT, int DEPTH, int NI, int NG, bool LATEB>
constexpr size_t mf_smem_bytes() {
  return (sizeof(MfWarpT<MfCfg<NG, LATEB>::CAP, MfCfg<NG, LATEB>::STAGES>) + MfRing<LANES, VEC, DEPTH, NI, NG>::BYTES) * MF_WARPS +
         (EXACT_DOT ? sizeof(float) * MF_WARPS * (32 / LANES) * Group<LANES, VEC>::DOT_FLOATS : 0);
}

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// One decoded tile: lane q of the warp holds row q (a tile is 32 rows).  Decoding once per
// tile and handing the fields to the lane groups with shuffles keeps the per-instance loop
// free of row_ptr / index loads and shape checks.
struct MfDec {
  unsigned urow, irow;  // slab rows of the user / item feature (valid if the row is taken)
  unsigned irow2;       // NI = 2: slab row of the second item feature, 0xffffffff if the row has one
  float uval, ival, ival2, lab;
  unsigned gp;          // NG > 0: (position of the row's first global feature in the staged window << 8) | count
  int stage;            // warp-uniform: the stage that holds this tile's index/value window
  unsigned take;        // warp-uniform: bit q = the fast pass takes row q
  unsigned left;        // warp-uniform: bit q = row q exists and is left to the generic pass
  bool all_one;         // warp-uniform: every taken row has uval = ival = "one"
  int r0;               // first row of the tile (absolute)
};

// NI = 1: rows (0 | 1 | 1), every row is a candidate.  NI = 2 (the second fast pass): rows
// (0 | 1 | 1..2) -- the pairwise-rank shape of configs[3], two distinct item features -- among the
// rows the first pass left (`gate` = its "something left" flag: nothing left, nothing to do).
// NG > 0 (the third fast pass): rows (1..NG | 1 | 1), strictly ascending global indices, among the
// rows the earlier passes left.
template <int LANES, int VEC, bool EXACT_DOT, bool TRAIN, int DEPTH, int MINB, int NI, int NG, bool LATEB>
__global__ void __launch_bounds__(MF_THREADS, MINB)
k_mf(DevModel m, DevHP hp, DevCsr csr, int row_begin, int row_end, int scatter_user, int scatter_item,
     float *pred_out, unsigned *row_mask, unsigned *any_left, const unsigned *gate) {
  constexpr bool FOLLOW = NI == 2 || NG > 0;  // a pass over what earlier passes left
  static_assert(NG <= LANES, "lane j of a group handles global feature j");
  static_assert(NG == 0 || NI == 1, "the global-feature pass takes one item feature");
  if (FOLLOW && *gate == 0u) return;
  using Ring = MfRing<LANES, VEC, DEPTH, NI, NG>;
  static_assert(!LATEB || NG > 0, "late B requests only exist for the global-feature pass");
  constexpr int MF_STAGES = MfCfg<NG, LATEB>::STAGES, MF_CAP = MfCfg<NG, LATEB>::CAP;
  using MfStage = MfStageT<MF_CAP>;
  using MfWarp = MfWarpT<MF_CAP, MF_STAGES>;
  constexpr int GPW = Ring::GPW, NCH = Ring::NCH;
  constexpr int ITER = MF_TILE / GPW;  // iterations per tile
  static_assert(ITER % DEPTH == 0 && DEPTH <= ITER, "ring depth must divide the iterations of a tile");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  MfWarp &sw = reinterpret_cast<MfWarp *>(smem_raw)[warp];
  float4 *ring = reinterpret_cast<float4 *>(smem_raw + sizeof(MfWarp) * MF_WARPS + Ring::BYTES * warp);
  float *dot_base = reinterpret_cast<float *>(smem_raw + (sizeof(MfWarp) + Ring::BYTES) * MF_WARPS);

  const int ntile = (row_end - row_begin + MF_TILE - 1) / MF_TILE;
  const int wslot = blockIdx.x * MF_WARPS + warp;
  const int wstride = gridDim.x * MF_WARPS;
  const int nlocal = ntile > wslot ? (ntile - wslot + wstride - 1) / wstride : 0;

  if (lane == 0) {
    for (int s = 0; s < MF_STAGES; ++s) {
      mbar_init(&sw.barA[s], 1);
      mbar_init(&sw.barB[s], 1);
    }
    mbar_fence_init();
  }
  __syncwarp();

  Group<LANES, VEC> g;
  g.gl = lane % LANES;
  const int gw = lane / LANES;
  g.gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (gw * LANES));
  g.dot_s = EXACT_DOT ? dot_base + (warp * GPW + gw) * Group<LANES, VEC>::DOT_FLOATS : nullptr;

  // ---- CSR staging (same scheme as k_stream, 32-row tiles) -------------------------------
  unsigned ph_a = 0, ph_b = 0;
  auto tile_of = [&](int j) { return wslot + j * wstride; };
  auto issue_a = [&](int j) {
    if (j >= nlocal) return;
    __syncwarp();  // every lane is done reading the stage being refilled
    if (lane == 0) {
      MfStage &st = sw.st[j % MF_STAGES];
      const int r0 = row_begin + tile_of(j) * MF_TILE;
      const int nrow = min(MF_TILE, row_end - r0);
      const int a_off = (3 * r0) & 3, l_off = r0 & 3;
      const unsigned bytesA = (unsigned)((a_off + 3 * nrow + 1 + 3) & ~3) * 4u;
      const unsigned bytesL = (unsigned)((l_off + nrow + 3) & ~3) * 4u;
      mbar_arrive_expect_tx(&sw.barA[j % MF_STAGES], bytesA + bytesL);
      bulk_g2s(st.rp, csr.row_ptr + (3 * r0 - a_off), bytesA, &sw.barA[j % MF_STAGES]);
      bulk_g2s(st.label, csr.label + (r0 - l_off), bytesL, &sw.barA[j % MF_STAGES]);
    }
  };
  struct Win {
    int v0, v1, v_off, nel, staged;
  };
  auto window = [&](int j) -> Win {
    const MfStage &st = sw.st[j % MF_STAGES];
    const int r0 = row_begin + tile_of(j) * MF_TILE;
    const int nrow = min(MF_TILE, row_end - r0);
    const int a_off = (3 * r0) & 3;
    Win w;
    w.v0 = st.rp[a_off] - csr.val_base;
    w.v1 = st.rp[a_off + 3 * nrow] - csr.val_base;
    w.v_off = w.v0 & 3;
    w.nel = (w.v_off + (w.v1 - w.v0) + 3) & ~3;
    w.staged = (w.v0 >= 0 && w.v1 >= w.v0 && w.v1 + csr.val_base <= csr.val_end && w.nel <= MF_CAP + 8) ? 1 : 0;
    return w;
  };
  auto issue_b = [&](int j) {
    if (j >= nlocal) return;
    mbar_wait(&sw.barA[j % MF_STAGES], (ph_a >> (j % MF_STAGES)) & 1u);
    ph_a ^= 1u << (j % MF_STAGES);
    const Win w = window(j);
    if (lane == 0 && w.staged && w.nel > 0) {
      MfStage &st = sw.st[j % MF_STAGES];
      mbar_arrive_expect_tx(&sw.barB[j % MF_STAGES], 2u * (unsigned)w.nel * 4u);
      bulk_g2s(st.idx, csr.index + (w.v0 - w.v_off), (unsigned)w.nel * 4u, &sw.barB[j % MF_STAGES]);
      bulk_g2s(st.val, csr.value + (w.v0 - w.v_off), (unsigned)w.nel * 4u, &sw.barB[j % MF_STAGES]);
    }
  };
  auto wait_b = [&](int j) {
    if (j >= nlocal) return;
    const Win w = window(j);
    if (w.staged && w.nel > 0) {
      mbar_wait(&sw.barB[j % MF_STAGES], (ph_b >> (j % MF_STAGES)) & 1u);
      ph_b ^= 1u << (j % MF_STAGES);
    }
  };
  // decode tile j (its A and B phases have landed): lane q looks at row q
  auto decode = [&](int j) -> MfDec {
    MfDec d;
    d.urow = d.irow = 0u;
    d.irow2 = 0xffffffffu;
    d.uval = d.ival = d.ival2 = d.lab = 0.0f;
    d.gp = 0u;
    d.stage = j % MF_STAGES;
    d.take = d.left = 0u;
    d.all_one = true;
    d.r0 = 0;
    if (j >= nlocal) return d;
    const MfStage &st = sw.st[j % MF_STAGES];
    d.r0 = row_begin + tile_of(j) * MF_TILE;
    const int nrow = min(MF_TILE, row_end - d.r0);
    const Win w = window(j);
    const int sm_base = w.v0 - w.v_off + csr.val_base;  // absolute feature position held by idx[0]
    const int v_hi = w.v1 + csr.val_base;
    const int *rp = st.rp + ((3 * d.r0) & 3) + 3 * lane;
    // candidates: every row of the tile (first pass) / the rows the first pass left (second pass)
    unsigned cand = nrow >= 32 ? 0xffffffffu : ((1u << nrow) - 1u);
    if (FOLLOW) cand &= row_mask[tile_of(j)];
    bool ok = w.staged && ((cand >> lane) & 1u);
    bool one = true;
    if (ok) {
      const int rp0 = rp[0], rp1 = rp[1], rp2 = rp[2], rp3 = rp[3];
      // the basic-MF shape (0 | 1 | 1 features; 0 | 1 | 2 too when NI = 2), inside the staged window
      const bool two = NI == 2 && rp3 == rp2 + 2;
      const int ngl = rp1 - rp0;
      ok = (NG > 0 ? (ngl >= 1 && ngl <= NG) : ngl == 0) && rp2 == rp1 + 1 && (rp3 == rp2 + 1 || two) &&
           rp0 >= sm_base && rp3 <= v_hi;
      if (NG > 0 && ok) {  // global indices in range and strictly ascending (no repeats: the generic pass)
        const int gpos = rp0 - sm_base;
        unsigned prev = st.idx[gpos];
        ok = prev < (unsigned)m.num_global;
        for (int q = 1; q < ngl; ++q) {
          const unsigned cur_g = st.idx[gpos + q];
          ok = ok && cur_g > prev && cur_g < (unsigned)m.num_global;
          prev = cur_g;
        }
        d.gp = ((unsigned)gpos << 8) | (unsigned)ngl;
      }
      if (ok) {
        const int f = rp1 - sm_base;
        const unsigned uid = st.idx[f], iid = st.idx[f + 1];
        ok = uid < (unsigned)m.num_user && iid < (unsigned)m.num_item;  // else: generic pass reports it
        d.urow = (unsigned)m.user_off + uid;
        d.irow = (unsigned)m.item_off + iid;
        d.uval = st.val[f];
        d.ival = st.val[f + 1];
        d.lab = st.label[(d.r0 & 3) + lane];
        one = scalar_is_one(d.uval) && scalar_is_one(d.ival);
        if (two) {
          const unsigned iid2 = st.idx[f + 2];
          ok = ok && iid2 < (unsigned)m.num_item && iid2 != iid;  // a repeated index: the generic pass
          d.irow2 = (unsigned)m.item_off + iid2;
          d.ival2 = st.val[f + 2];
          one = false;
        }
      }
    }
    d.take = __ballot_sync(0xffffffffu, ok);
    d.left = ~d.take & cand;
    d.all_one = __all_sync(0xffffffffu, !ok || one);
    return d;
  };

  // ---- the row ring -------------------------------------------------------------------------
  const int row_f4 = m.pitch >> 2;  // float4 chunks a row really has (<= NCH)
  float4 *const my_ring = ring + (size_t)gw * Ring::SLOT_F4;  // this group's part of slot 0
  constexpr int SLOT_STRIDE = GPW * Ring::SLOT_F4;             // float4 between consecutive slots
  // issue the gathers of iteration i of tile d into ring slot `slot` (always one commit group)
  auto prefetch = [&](const MfDec &d, int i, int slot) {
    const int src = i * GPW + gw;
    const unsigned urow = __shfl_sync(0xffffffffu, d.urow, src);
    const unsigned irow = __shfl_sync(0xffffffffu, d.irow, src);
    const unsigned irow2 = NI == 2 ? __shfl_sync(0xffffffffu, d.irow2, src) : 0xffffffffu;
    const unsigned gp = NG > 0 ? __shfl_sync(0xffffffffu, d.gp, src) : 0u;
    if ((d.take >> src) & 1u) {
      float4 *dst = my_ring + slot * SLOT_STRIDE;
      const float *pu = m.W + (size_t)urow * (size_t)m.pitch, *pi = m.W + (size_t)irow * (size_t)m.pitch;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int ch = g.gl + v * LANES;
        if (ch < row_f4) {
          cp_async16(dst + ch, pu + 4 * ch);
          cp_async16(dst + NCH + ch, pi + 4 * ch);
          if (NI == 2 && irow2 != 0xffffffffu) cp_async16(dst + 2 * NCH + ch, m.W + (size_t)irow2 * (size_t)m.pitch + 4 * ch);
        }
      }
      // biases: the aligned 16-byte window that holds the element (cp.async.cg moves 16 bytes)
      constexpr int WIN = (1 + NI) * NCH;
      if (g.gl == 0 && !m.no_user_bias) cp_async16(dst + WIN, m.bias + (urow & ~3u));
      if (g.gl == LANES - 1) cp_async16(dst + WIN + 1, m.bias + (irow & ~3u));
      if (NI == 2 && g.gl == LANES - 2 && irow2 != 0xffffffffu) cp_async16(dst + WIN + 2, m.bias + (irow2 & ~3u));
      if (NG > 0 && g.gl < (int)(gp & 0xffu)) {  // lane j: the g_bias window of global feature j
        const unsigned gid = sw.st[d.stage].idx[(gp >> 8) + g.gl];
        cp_async16(dst + WIN + 1 + NI + g.gl, m.g_bias + (gid & ~3u));
      }
    }
    cp_async_commit();
  };

    // one instance per lane group: rows from ring slot `slot`, scalars from the decoded tile.
  // Groups whose row is not taken run the arithmetic on whatever the slot holds and skip the
  // memory operations (no divergent control flow in the common path).
  struct Rows {
    float4 wu[VEC], wi[VEC], wi2[VEC];
    float ub, ib, ib2;
    unsigned urow, irow, irow2;
    // NG > 0: lane j of the group holds global feature j of the instance
    int ng;
    unsigned gid;
    float gval, gb;
  };
  auto load_slot = [&](const MfDec &d, int i, int slot, Rows &r) {
    const int src = i * GPW + gw;
    r.urow = __shfl_sync(0xffffffffu, d.urow, src);
    r.irow = __shfl_sync(0xffffffffu, d.irow, src);
    const float4 *rs = my_ring + slot * SLOT_STRIDE;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int ch = g.gl + v * LANES;
      r.wu[v] = rs[ch];
      r.wi[v] = rs[NCH + ch];
      if (ch >= row_f4) r.wu[v] = r.wi[v] = f4_zero();  // chunks the row does not have
    }
    const float *bw = reinterpret_cast<const float *>(rs + (1 + NI) * NCH);
    r.ub = m.no_user_bias ? 0.0f : bw[r.urow & 3u];
    r.ib = bw[4 + (r.irow & 3u)];
    r.irow2 = 0xffffffffu;
    r.ib2 = 0.0f;
    if (NI == 2) {
      r.irow2 = __shfl_sync(0xffffffffu, d.irow2, src);
      const bool two = r.irow2 != 0xffffffffu;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int ch = g.gl + v * LANES;
        r.wi2[v] = (two && ch < row_f4) ? rs[2 * NCH + ch] : f4_zero();  // (never garbage: it enters tmp_ifactor)
      }
      r.ib2 = two ? bw[8 + (r.irow2 & 3u)] : 0.0f;
    }
    r.ng = 0;
    r.gid = 0u;
    r.gval = r.gb = 0.0f;
    if (NG > 0) {
      const unsigned gp = __shfl_sync(0xffffffffu, d.gp, src);
      r.ng = ((d.take >> src) & 1u) ? (int)(gp & 0xffu) : 0;
      if (g.gl < r.ng) {
        const MfStage &st = sw.st[d.stage];
        r.gid = st.idx[(gp >> 8) + g.gl];
        r.gval = st.val[(gp >> 8) + g.gl];
        r.gb = bw[4 * (1 + NI + g.gl) + (r.gid & 3u)];
      }
    }
  };
  auto compute = [&](const MfDec &d, int i, const Rows &r) {
    const int src = i * GPW + gw;
    const bool take = (d.take >> src) & 1u;
    const unsigned urow = r.urow, irow = r.irow;
    const float4(&wu)[VEC] = r.wu;
    const float4(&wi)[VEC] = r.wi;
    const float ub = r.ub, ib = r.ib;

    // prepare_tmp (base.h:354-381): tmp = 0 + w*val
    float4 tu[VEC], ti[VEC];
    float uval = 1.0f, ival = 1.0f;
    if (d.all_one && !EXACT_DOT) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        tu[v] = wu[v];
        ti[v] = wi[v];
      }
    } else {
      uval = __shfl_sync(0xffffffffu, d.uval, src);
      ival = __shfl_sync(0xffffffffu, d.ival, src);
      // w*1.0f is w exactly, so the "scalar is one" shortcut needs no second code path
      const float um = scalar_is_one(uval) ? 1.0f : uval, im = scalar_is_one(ival) ? 1.0f : ival;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        tu[v] = f4_add_scaled(f4_zero(), wu[v], um, false);
        ti[v] = f4_add_scaled(f4_zero(), wi[v], im, false);
      }
    }
    // second item feature (NI = 2; a row without one contributes exact zeros)
    const bool two = NI == 2 && r.irow2 != 0xffffffffu;
    float ival2 = 0.0f;
    if (NI == 2) {
      const float iv2 = __shfl_sync(0xffffffffu, d.ival2, src);  // (every lane takes part in the shuffle)
      ival2 = two ? iv2 : 0.0f;
      const float im2 = scalar_is_one(ival2) ? 1.0f : ival2;
#pragma unroll
      for (int v = 0; v < VEC; ++v) ti[v] = f4_add_scaled(ti[v], r.wi2[v], im2, false);
    }
    // calc_bias (base.h:313-353) + pred (base.h:445-454)
    double bsum = 0.0;
    if (NG > 0) {  // globals first, in feature order (base.h:318-322)
      const float p = g.gl < r.ng ? __fmul_rn(r.gval, r.gb) : 0.0f;
      for (int q = 0; q < r.ng; ++q) bsum = __dadd_rn(bsum, (double)g.bcast(p, q));
    }
    if (!m.no_user_bias) bsum = __dadd_rn(bsum, (double)__fmul_rn(uval, ub));
    bsum = __dadd_rn(bsum, (double)__fmul_rn(ival, ib));
    if (NI == 2) bsum = __dadd_rn(bsum, two ? (double)__fmul_rn(ival2, r.ib2) : 0.0);
    // (measured: an out-of-line sigmoid path or full-mask shuffles here cost 4 % each)
    const float dt = g.template dot<EXACT_DOT>(m, tu, ti);
    double sum = __dadd_rn((double)hp.base_score, bsum);
    sum = __dadd_rn(sum, (double)dt);
    const float lab = TRAIN ? __shfl_sync(0xffffffffu, d.lab, src) : 0.0f;
    const float pred = map_active((float)sum, m.active_type);
    const float err = cal_grad(lab, pred, m.active_type);
    if (!TRAIN) {
      if (take && g.gl == 0) pred_out[d.r0 + src - row_begin] = pred;
      return;
    }
    // update_no_decay + regularize(after), fused (base.h:383-427, 211-283)
    const float lrerr = __fmul_rn(hp.lr, err);
    const float su = __fmul_rn(lrerr, uval), si = __fmul_rn(lrerr, ival);
    const float su_m = scalar_is_one(su) ? 1.0f : su, si_m = scalar_is_one(si) ? 1.0f : si;
    float4 nu[VEC], ni[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      nu[v] = f4_add_scaled(wu[v], ti[v], su_m, false);
      ni[v] = f4_add_scaled(wi[v], tu[v], si_m, false);
      if (!hp.du_skip) nu[v] = f4_scale(nu[v], hp.du);
      if (!hp.di_skip) ni[v] = f4_scale(ni[v], hp.di);
    }
    const float nub = __fmul_rn(__fadd_rn(ub, su), hp.dub), nib = __fmul_rn(__fadd_rn(ib, si), hp.dib);
    if (NG > 0 && g.gl < r.ng) {  // g_bias[gid] += lr*err*gval, then *= 1 - lr*wd_global (base.h:388-390,188-192)
      float x = __fadd_rn(r.gb, __fmul_rn(lrerr, r.gval));
      if (r.gid >= hp.regfree) x = __fmul_rn(x, hp.dg);
      if (scatter_item == SCATTER_RED) red1(m.g_bias + r.gid, __fsub_rn(x, r.gb));
      else __stcg(m.g_bias + r.gid, x);
    }
    if (take) {
      if (scatter_user == SCATTER_RED) g.red_row(m, urow, nu, wu);
      else g.store_row(m, urow, nu);
      if (scatter_item == SCATTER_RED) g.red_row(m, irow, ni, wi);
      else g.store_row(m, irow, ni);
      if (g.gl == 0 && !m.no_user_bias) {
        if (scatter_user == SCATTER_RED) red1(m.bias + urow, __fsub_rn(nub, ub));
        else __stcg(m.bias + urow, nub);
      }
      if (g.gl == LANES - 1) {
        if (scatter_item == SCATTER_RED) red1(m.bias + irow, __fsub_rn(nib, ib));
        else __stcg(m.bias + irow, nib);
      }
    }
    if (NI == 2 && take && two) {  // second item feature: same update with its own value
      const float si2 = __fmul_rn(lrerr, ival2);
      const float si2_m = scalar_is_one(si2) ? 1.0f : si2;
      float4 n2[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        n2[v] = f4_add_scaled(r.wi2[v], tu[v], si2_m, false);
        if (!hp.di_skip) n2[v] = f4_scale(n2[v], hp.di);
      }
      if (scatter_item == SCATTER_RED) g.red_row(m, r.irow2, n2, r.wi2);
      else g.store_row(m, r.irow2, n2);
      if (g.gl == LANES - 2) {
        const float nib2 = __fmul_rn(__fadd_rn(r.ib2, si2), hp.dib);
        if (scatter_item == SCATTER_RED) red1(m.bias + r.irow2, __fsub_rn(nib2, r.ib2));
        else __stcg(m.bias + r.irow2, nib2);
      }
    }
  };

  // ---- main loop ------------------------------------------------------------------------------
  // Staging schedule (two stages; tile t uses stage t % 2): a tile's stage is free again as soon
  // as the tile is decoded into registers, so while tile j is computed, tile j+1 is decoded,
  // B(j+2) is in flight in the stage tile j left and A(j+3) in the stage tile j+1 left.
  issue_a(0);
  issue_a(1);
  issue_b(0);
  wait_b(0);
  MfDec cur = decode(0);
  issue_a(2);
  issue_b(1);
#pragma unroll
  for (int i = 0; i < DEPTH; ++i) prefetch(cur, i, i);  // fill the ring: DEPTH groups pending
  for (int j = 0; j < nlocal; ++j) {
    wait_b(j + 1);
    const MfDec nxt = decode(j + 1);
    issue_a(j + 3);
    if (!LATEB) issue_b(j + 2);
    if (lane == 0 && (cur.left != 0u || (FOLLOW && cur.take != 0u))) {
      row_mask[tile_of(j)] = cur.left;   // (later passes: also clear the rows they take)
      if (cur.left != 0u) *any_left = 1u;  // the generic pass has something to do
    }
    if (cur.take == 0u) {  // (warp-uniform) nothing of this tile is ours: its DEPTH pending groups are
#pragma unroll             // empty; queue the next tile's first iterations behind them and move on
      for (int i = 0; i < DEPTH; ++i) prefetch(nxt, i, i);
      if (LATEB) issue_b(j + 2);
      cur = nxt;
      continue;
    }
#pragma unroll 1
    for (int i = 0; i < ITER; ++i) {  // iteration i lives in ring slot i % DEPTH (one copy of the code:
      const int slot = i % DEPTH;     // the unrolled loop did not fit the instruction cache)
      cp_async_wait<DEPTH - 1>();     // its gathers have landed (this thread's part)
      __syncwarp();                   // ... and every lane's part
      Rows r;
      load_slot(cur, i, slot, r);
      __syncwarp();                   // the slot has been read by every lane: refill it
      if (i + DEPTH < ITER) prefetch(cur, i + DEPTH, slot);
      else prefetch(nxt, i + DEPTH - ITER, slot);
      compute(cur, i, r);
    }
    if (LATEB) issue_b(j + 2);  // tile j's window is no longer needed: its stage takes tile j+2's
    cur = nxt;
  }
  cp_async_wait<0>();
}

template <int L, int V, int DEPTH, int MINB, int NI, int NG, bool LATEB = false>
static int launch_mf_geo(svdgpu *h, const DevCsr &csr, int r0, int r1, bool train, float *pred, unsigned *flag_out,
                         const unsigned *flag_gate) {
  const long long ntile = ((long long)(r1 - r0) + MF_TILE - 1) / MF_TILE;
  int grid = 1;
#define GO(ED, TR)                                                                               \
  {                                                                                              \
    auto k = k_mf<L, V, ED, TR, DEPTH, MINB, NI, NG, LATEB>;                                     \
    const size_t smem = mf_smem_bytes<L, V, ED, DEPTH, NI, NG, LATEB>();                         \
    CU(h, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
    if (grid_for(h, k, MF_THREADS, (ntile + MF_WARPS - 1) / MF_WARPS, &grid, smem)) return 1;    \
    /* Hogwild stability: at most h->inflight_cap instances in flight (svdgpu_api.cu, hot_cap) */ \
    if (TR && h->inflight_cap > 0) {                                                             \
      const long long per_cta = (long long)MF_WARPS * (32 / L) * DEPTH;                          \
      if ((long long)grid * per_cta > h->inflight_cap) grid = (int)std::max<long long>(1, h->inflight_cap / per_cta); \
    }                                                                                            \
    k<<<grid, MF_THREADS, smem, h->stream>>>(h->dm, h->dhp, csr, r0, r1, h->scatter_user,        \
                                             h->scatter_item, pred, h->d_row_mask,               \
                                             flag_out, flag_gate);                               \
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

// The fast passes over rows [r0, r1); h->d_row_mask (one word per 32-row tile, zeroed by the
// caller) receives the rows left for the generic pass.  First pass: rows (0|1|1).  Second pass
// (`second`): rows (0|1|1..2) among those the first pass left; exits at once if it left none.
// `which`: 0 first pass, 1 second pass (two item features), 2 third pass (global features).
// flag_out receives "this pass left something"; a later pass exits at once when *flag_gate == 0.
int launch_mf(svdgpu *h, const Geometry &g, const DevCsr &csr, int r0, int r1, bool train, float *pred,
              int which, unsigned *flag_out, const unsigned *flag_gate) {
  // tuning knobs (options "ring_depth", "mf_ctas"): ring depth 2 or 4, 2 or 3 CTAs per SM
  int depth = h->ring_depth == 2 ? 2 : 4;
  // a capped launch first halves the ring (half the instances in flight at nearly the same speed), then
  // shrinks the grid
  if (train && h->inflight_cap > 0 && h->inflight_cap < (long long)h->num_sm * 2 * MF_WARPS * (32 / g.lanes) * 4) depth = 2;
  const int minb = h->mf_ctas == 3 ? 3 : 2;
#define GEO(L, V)                                                                                      \
  if (g.lanes == L && g.vec == V) {                                                                    \
    constexpr int NGV = L < 16 ? L : 16;                                                               \
    if (which == 2 && h->mfg != 0) return launch_mf_geo<L, V, 2, 2, 1, NGV, true>(h, csr, r0, r1, train, pred, flag_out, flag_gate); \
    if (which == 2) return launch_mf_geo<L, V, 4, 1, 1, NGV>(h, csr, r0, r1, train, pred, flag_out, flag_gate); \
    if (which == 1) return launch_mf_geo<L, V, 2, 2, 2, 0>(h, csr, r0, r1, train, pred, flag_out, flag_gate);   \
    if (depth == 2 && minb == 3) return launch_mf_geo<L, V, 2, 3, 1, 0>(h, csr, r0, r1, train, pred, flag_out, flag_gate); \
    if (depth == 2) return launch_mf_geo<L, V, 2, 2, 1, 0>(h, csr, r0, r1, train, pred, flag_out, flag_gate);   \
    if (minb == 3) return launch_mf_geo<L, V, 4, 3, 1, 0>(h, csr, r0, r1, train, pred, flag_out, flag_gate);    \
    return launch_mf_geo<L, V, 4, 2, 1, 0>(h, csr, r0, r1, train, pred, flag_out, flag_gate);                   \
  }
#ifdef SVDGPU_TUNE_BUILD
  GEO(4, 4) GEO(8, 2) GEO(16, 2)
#else
  GEO(4, 1) GEO(4, 2) GEO(4, 4) GEO(8, 1) GEO(8, 2) GEO(8, 4) GEO(16, 1) GEO(16, 2) GEO(16, 4)
  GEO(32, 1) GEO(32, 2) GEO(32, 4)
#endif
#undef GEO
  return fail(h, "k_mf: no kernel instantiated for lanes=%d vec=%d", g.lanes, g.vec);
}

}  // namespace svdk
