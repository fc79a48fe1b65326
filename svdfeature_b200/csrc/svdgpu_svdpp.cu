// svdgpu_svdpp.cu -- k_svdpp: Hogwild SVD++ training, one WARP per user unit.
//
// Why a warp per user: SVD++ (SVDPPFeature::update(SVDPlusBlock), base.h:568-582) is sequential
// inside a user -- every rating row reads and updates tmp_ufeedback and the user's factor -- and
// the W_ufeedback rows of popular items are shared by a large share of all users, so only a few
// dozen users may be in flight at once before Hogwild stops tracking the sequential order
// (tools/hogwild_parity.py --svdpp: 64 users in flight agree to 1e-2, 256 drift away, 600 end in
// NaN).  With the number of users in flight bounded, throughput is the speed of ONE user; this
// kernel makes a user fast instead of running many slow ones:
//   * the 32 lanes split every k-vector (lane l owns CPL = pitch/32 consecutive components);
//     the user's factor, its bias and tmp_ufeedback live in registers for the whole unit;
//   * the feedback gather / scatter run on all 32 lanes (svdgpu_fb.cuh);
//   * the unit's CSR rows are decoded 32 at a time, lane per row, one tile ahead;
//   * the item rows (and the 16-byte windows holding their biases) stream through a ring in
//     shared memory, fetched with cp.async.cg RING rows ahead, so the per-row critical path is
//     the arithmetic alone.
// Units this kernel cannot take (rows of another shape, several users in a block, side features,
// k not a multiple of 32, the reference dot order requested) stay with k_ugroup: k_unit_kind
// classifies every unit first.  Arithmetic per component is that of process_instance<SVDPP>; the
// dot product uses the shuffle-tree order (Hogwild default, exact_dot = 0).
#include "svdgpu_internal.h"
#include "svdgpu_fb.cuh"

namespace svdk {

constexpr int SP_RING = 8;  // item rows in flight

// kind[u - unit_begin] = 1 when every row of unit u is (0 | 1 | 1) with one and the same user
// index and all indices in range (one warp per unit)
__global__ void k_unit_kind(DevModel m, DevCsr csr, DevUgroup ug, int unit_begin, int unit_end, unsigned char *kind) {
  const int u = unit_begin + (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (u >= unit_end) return;
  const unsigned *idx = csr.index - csr.val_base;
  const int *row_ptr = csr.row_ptr - 3 * (long long)ug.row_base;
  const int r0 = ug.blk_row_off[ug.unit_off[u]], r1 = ug.blk_row_off[ug.unit_off[u + 1]];
  bool ok = true;
  unsigned uid0 = 0xffffffffu;
  if (r1 > r0) {
    const int f0 = row_ptr[3 * (long long)r0 + 1];
    if (f0 >= csr.val_base && f0 < csr.val_end) uid0 = idx[f0];
    else ok = false;
  }
  for (int r = r0 + lane; r < r1 && ok; r += 32) {
    const int rp0 = row_ptr[3 * (long long)r], rp1 = row_ptr[3 * (long long)r + 1];
    const int rp2 = row_ptr[3 * (long long)r + 2], rp3 = row_ptr[3 * (long long)r + 3];
    ok = rp0 >= csr.val_base && rp3 <= csr.val_end && rp1 == rp0 && rp2 == rp1 + 1 && rp3 == rp2 + 1;
    if (ok) ok = idx[rp1] == uid0 && uid0 < (unsigned)m.num_user && idx[rp2] < (unsigned)m.num_item;
  }
  ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) kind[u - unit_begin] = ok ? 1 : 0;
}

// one decoded tile of 32 rows: lane j holds row (tile start + j)
struct SpDec {
  unsigned irow;  // slab row of the item
  float uval, ival, lab;
};

template <int CPL>
__global__ void __launch_bounds__(32)
k_svdpp(DevModel m, DevHP hp, DevCsr csr, DevUgroup ug, int unit_begin, int unit_end, const unsigned char *kind,
        int scatter_item, unsigned *counter, int *err_flag) {
  constexpr int ROWF = 32 * CPL;        // floats of a row
  constexpr int SLOTF = ROWF + 4;       // + the 16-byte window holding the item's bias
  __shared__ __align__(16) float ring[SP_RING * SLOTF];
  const int lane = threadIdx.x;
  const unsigned *idx = csr.index - csr.val_base;
  const float *val = csr.value - csr.val_base;
  const int *row_ptr = csr.row_ptr - 3 * (long long)ug.row_base;
  const float *label = csr.label - ug.row_base;

  for (;;) {
    unsigned n = 0;
    if (lane == 0) n = atomicAdd(counter, 1u);
    n = __shfl_sync(0xffffffffu, n, 0);
    if ((long long)unit_begin + n >= unit_end) break;
    int u = unit_begin + (int)n;
    if (ug.order) u = ug.order[u];
    if (!kind[u - unit_begin]) continue;  // k_ugroup's
    const int b0 = ug.unit_off[u], b1 = ug.unit_off[u + 1];
    const int f0 = ug.blk_fb_off[b0] - ug.fb_base, nf0 = ug.blk_fb_off[b0 + 1] - ug.blk_fb_off[b0];
    const int f1 = ug.blk_fb_off[b1 - 1] - ug.fb_base, nf1 = ug.blk_fb_off[b1] - ug.blk_fb_off[b1 - 1];
    const int r0 = ug.blk_row_off[b0], r1 = ug.blk_row_off[b1];

    // ---- prepare_ufeedback (base.h:523-538) ----
    float fb[CPL], old[CPL], norm, fb_bias;
    if (!coop_prepare_ufeedback_regs<CPL>(m, ug.fb_index + f0, ug.fb_value + f0, nf0, lane, fb, norm, fb_bias, err_flag))
      continue;
    const float old_bias = fb_bias;
#pragma unroll
    for (int c = 0; c < CPL; ++c) old[c] = fb[c];
    if (r1 == r0) continue;  // no rating rows: tmp_ufeedback unchanged, the scatter would add zeros

    // ---- the user's factor and bias: registers for the whole unit ----
    const unsigned uid = idx[row_ptr[3 * (long long)r0 + 1]];
    const size_t urow = (size_t)m.user_off + uid;
    float wu[CPL], ub = 0.0f;
    ld_cpl<CPL>(m.W + urow * (size_t)m.pitch + lane * CPL, wu);
    if (!m.no_user_bias) ub = __ldcg(m.bias + urow);

    auto decode = [&](int t0) -> SpDec {  // rows t0 .. t0+31 of the unit, lane per row
      SpDec d;
      d.irow = 0u;
      d.uval = d.ival = d.lab = 0.0f;
      const int r = t0 + lane;
      if (r < r1) {
        const int f = row_ptr[3 * (long long)r + 1];
        d.irow = (unsigned)m.item_off + idx[f + 1];
        d.uval = val[f];
        d.ival = val[f + 1];
        d.lab = label[r];
      }
      return d;
    };
    // gathers of row j of the tile starting at t0 into ring slot `slot` (always one commit group)
    auto prefetch = [&](const SpDec &d, int t0, int j, int slot) {
      const unsigned irow = __shfl_sync(0xffffffffu, d.irow, j & 31);
      if (t0 + j < r1) {
        float *dst = ring + slot * SLOTF;
        if (lane < ROWF / 4)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + 4 * lane)),
                       "l"(m.W + (size_t)irow * (size_t)m.pitch + 4 * lane)
                       : "memory");
        if (lane == 31)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + ROWF)),
                       "l"(m.bias + (irow & ~3u))
                       : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };

    SpDec cur = decode(r0);
#pragma unroll
    for (int j = 0; j < SP_RING; ++j) prefetch(cur, r0, j, j);
    for (int t0 = r0; t0 < r1; t0 += 32) {
      const SpDec nxt = decode(t0 + 32);
      const int nrow = min(32, r1 - t0);
#pragma unroll 1
      for (int j = 0; j < nrow; ++j) {
        const int slot = j % SP_RING;  // (32 is a multiple of SP_RING: slots line up across tiles)
        asm volatile("cp.async.wait_group %0;" ::"n"(SP_RING - 1) : "memory");
        __syncwarp();
        const float *src = ring + slot * SLOTF;
        const unsigned irow = __shfl_sync(0xffffffffu, cur.irow, j);
        float wi[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) wi[c] = src[lane * CPL + c];
        const float ib = src[ROWF + (irow & 3u)];
        __syncwarp();  // the slot has been read by every lane: refill it
        if (j + SP_RING < 32) prefetch(cur, t0, j + SP_RING, slot);
        else prefetch(nxt, t0 + 32, j + SP_RING - 32, slot);

        const float uval = __shfl_sync(0xffffffffu, cur.uval, j), ival = __shfl_sync(0xffffffffu, cur.ival, j);
        const float lab = __shfl_sync(0xffffffffu, cur.lab, j);
        const float um = scalar_is_one(uval) ? 1.0f : uval, im = scalar_is_one(ival) ? 1.0f : ival;
        // prepare_tmp (base.h:354-381) on top of tmp_ufeedback (prepare_svdpp, :506-508)
        float tu[CPL], ti[CPL], dt = 0.0f;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          tu[c] = __fadd_rn(fb[c], __fmul_rn(wu[c], um));
          ti[c] = __fadd_rn(0.0f, __fmul_rn(wi[c], im));
          dt = __fadd_rn(dt, __fmul_rn(tu[c], ti[c]));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dt = __fadd_rn(dt, __shfl_xor_sync(0xffffffffu, dt, o));
        // calc_bias (base.h:313-353) with get_bias_svdpp (:509-511), pred (:445-454)
        double bsum = 0.0;
        if (!m.no_user_bias) {
          bsum = __dadd_rn(bsum, (double)__fmul_rn(uval, ub));
          bsum = __dadd_rn(bsum, (double)fb_bias);
        }
        bsum = __dadd_rn(bsum, (double)__fmul_rn(ival, ib));
        double sum = __dadd_rn((double)hp.base_score, bsum);
        sum = __dadd_rn(sum, (double)dt);
        const float p = map_active((float)sum, m.active_type);
        // update_no_decay + regularize(after) (base.h:383-427, 211-283)
        const float err = cal_grad(lab, p, m.active_type);
        const float lrerr = __fmul_rn(hp.lr, err);
        const float su = __fmul_rn(lrerr, uval), si = __fmul_rn(lrerr, ival);
        const float su_m = scalar_is_one(su) ? 1.0f : su, si_m = scalar_is_one(si) ? 1.0f : si;
        float ni[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          wu[c] = __fadd_rn(wu[c], __fmul_rn(ti[c], su_m));
          ni[c] = __fadd_rn(wi[c], __fmul_rn(tu[c], si_m));
          if (!hp.du_skip) wu[c] = __fmul_rn(wu[c], hp.du);
          if (!hp.di_skip) ni[c] = __fmul_rn(ni[c], hp.di);
        }
        if (!m.no_user_bias) ub = __fmul_rn(__fadd_rn(ub, su), hp.dub);
        float *ip = m.W + (size_t)irow * (size_t)m.pitch + lane * CPL;
        if (scatter_item == SCATTER_RED) {
#pragma unroll
          for (int c = 0; c < CPL; ++c) ni[c] = __fsub_rn(ni[c], wi[c]);
          red_cpl<CPL>(ip, ni);
        } else {
          st_cpl<CPL>(ip, ni);
        }
        if (lane == 0) {
          const float nib = __fmul_rn(__fadd_rn(ib, si), hp.dib);
          if (scatter_item == SCATTER_RED) red1(m.bias + irow, __fsub_rn(nib, ib));
          else __stcg(m.bias + irow, nib);
        }
        // update_svdpp (base.h:512-520)
        const float sf = __fmul_rn(__fmul_rn(hp.lr_fb, err), norm);
        const float sf_m = scalar_is_one(sf) ? 1.0f : sf;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          fb[c] = __fadd_rn(fb[c], __fmul_rn(ti[c], sf_m));
          if (!hp.dfb_skip) fb[c] = __fmul_rn(fb[c], hp.dfb);
        }
        if (!m.no_user_bias) {
          fb_bias = __fadd_rn(fb_bias, sf);
          fb_bias = __fmul_rn(fb_bias, hp.dfbb);
        }
      }
      cur = nxt;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    // ---- write the user back; update_ufeedback (base.h:539-554) ----
    st_cpl<CPL>(m.W + urow * (size_t)m.pitch + lane * CPL, wu);
    if (!m.no_user_bias && lane == 0) __stcg(m.bias + urow, ub);
    if (nf1 > 0) {
      const float inv = __fdiv_rn(1.0f, norm);
      const bool inv_one = scalar_is_one(inv);
      float d[CPL];
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        d[c] = __fsub_rn(fb[c], old[c]);
        if (!inv_one) d[c] = __fmul_rn(d[c], inv);
      }
      const float dbias = __fmul_rn(__fsub_rn(fb_bias, old_bias), inv);
      coop_update_ufeedback_regs<CPL>(m, ug.fb_index + f1, ug.fb_value + f1, nf1, lane, d, dbias, scatter_item);
    }
  }
}

// Units [u0,u1): classify, then run k_svdpp over the units of kind 1.  Returns through *kind_out
// the device array k_ugroup must honour (kind 1 = done here).  `warps` = units in flight.
int launch_svdpp(svdgpu *h, const DevCsr &csr, const DevUgroup &ug, int u0, int u1, int warps,
                 const unsigned char **kind_out) {
  *kind_out = nullptr;
  const int cpl = (h->dm.pitch & 31) == 0 ? h->dm.pitch >> 5 : 0;
  if (cpl != 1 && cpl != 2 && cpl != 4) return 0;
  const size_t nu = (size_t)(u1 - u0);
  if (nu > h->unit_kind_cap) {
    if (h->d_unit_kind) CU(h, cudaFree(h->d_unit_kind));
    h->d_unit_kind = nullptr;
    h->unit_kind_cap = 0;
    CU(h, cudaMalloc(&h->d_unit_kind, nu * 2 + 64));
    h->unit_kind_cap = nu * 2;
  }
  k_unit_kind<<<(int)((nu * 32 + 255) / 256), 256, 0, h->stream>>>(h->dm, csr, ug, u0, u1, h->d_unit_kind);
  CU(h, cudaMemsetAsync(h->d_counter, 0, sizeof(unsigned), h->stream));
  const int grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(warps, 1), nu));
  if (cpl == 4) k_svdpp<4><<<grid, 32, 0, h->stream>>>(h->dm, h->dhp, csr, ug, u0, u1, h->d_unit_kind, h->scatter_item, h->d_counter, h->d_err);
  else if (cpl == 2) k_svdpp<2><<<grid, 32, 0, h->stream>>>(h->dm, h->dhp, csr, ug, u0, u1, h->d_unit_kind, h->scatter_item, h->d_counter, h->d_err);
  else k_svdpp<1><<<grid, 32, 0, h->stream>>>(h->dm, h->dhp, csr, ug, u0, u1, h->d_unit_kind, h->scatter_item, h->d_counter, h->d_err);
  CU(h, cudaGetLastError());
  h->n_launch += 2;
  *kind_out = h->d_unit_kind;
  return 0;
}

}  // namespace svdk
