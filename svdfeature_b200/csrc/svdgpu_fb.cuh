// svdgpu_fb.cuh -- whole-warp gather / scatter of a user's implicit-feedback rows (SVD++),
// shared by k_ugroup (svdgpu_ordered.cu) and k_svdpp (svdgpu_svdpp.cu).
#pragma once
#include "svdgpu_device.cuh"

namespace svdk {

// ---------------------------------------------------------------------------
// Whole-warp versions of the feedback gather / scatter (Hogwild).  A unit's feedback list is
// as long as its rating rows (configs[2]: ~200 each) and one lane group alone handles it at a
// quarter of the warp's width, so the warp does it together for one group at a time: lane l owns
// the CPL = pitch/32 components [l*CPL, (l+1)*CPL) of every row.  Per component the arithmetic
// and its order are those of prepare_ufeedback / update_ufeedback, so the result is bit-identical;
// the owning group receives / provides the k-vector through `xch` (shared memory, pitch floats).
// ---------------------------------------------------------------------------
template <int CPL>
__device__ __forceinline__ void ld_cpl(const float *p, float (&w)[CPL]) {
  if (CPL == 4) {
    const float4 q = __ldcg(reinterpret_cast<const float4 *>(p));
    w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3 % CPL] = q.w;
  } else if (CPL == 2) {
    const float2 q = __ldcg(reinterpret_cast<const float2 *>(p));
    w[0] = q.x; w[1 % CPL] = q.y;
  } else {
    w[0] = __ldcg(p);
  }
}
template <int CPL>
__device__ __forceinline__ void st_cpl(float *p, const float (&w)[CPL]) {
  if (CPL == 4) __stcg(reinterpret_cast<float4 *>(p), make_float4(w[0], w[1 % CPL], w[2 % CPL], w[3 % CPL]));
  else if (CPL == 2) __stcg(reinterpret_cast<float2 *>(p), make_float2(w[0], w[1 % CPL]));
  else __stcg(p, w[0]);
}
template <int CPL>
__device__ __forceinline__ void red_cpl(float *p, const float (&w)[CPL]) {
  if (CPL == 4) red4(p, make_float4(w[0], w[1 % CPL], w[2 % CPL], w[3 % CPL]));
  else if (CPL == 2)
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(w[0]), "f"(w[1 % CPL]) : "memory");
  else red1(p, w[0]);
}

// base.h:523-538, all 32 lanes.  Returns false (error flagged) on a bad feedback id.
// register form: acc[] = this lane's CPL components of tmp_ufeedback; B = rows gathered before any
// is consumed (measured: 32 instead of 8 changes nothing in k_svdpp)
template <int CPL, int B = 8>
__device__ __forceinline__ bool coop_prepare_ufeedback_regs(const DevModel &m, const unsigned *fi, const float *fv,
                                                            int nfb, int lane, float (&acc)[CPL], float &norm,
                                                            float &fb_bias, int *err_flag) {
  bool bad = false;
  for (int i = lane; i < nfb; i += 32) bad |= fi[i] >= (unsigned)m.num_ufeedback;
  if (__any_sync(0xffffffffu, bad)) {
    if (bad) atomicCAS(err_flag, 0, ERR_FB_INDEX);
    return false;
  }
#pragma unroll
  for (int c = 0; c < CPL; ++c) acc[c] = 0.0f;
  norm = 0.0f;
  fb_bias = 0.0f;
  for (int i0 = 0; i0 < nfb; i0 += B) {
    float w[B][CPL], x[B];
#pragma unroll
    for (int j = 0; j < B; ++j) {
      const int i = min(i0 + j, nfb - 1);
      x[j] = fv[i];
      ld_cpl<CPL>(m.W + (size_t)fi[i] * (size_t)m.pitch + lane * CPL, w[j]);
    }
#pragma unroll
    for (int j = 0; j < B; ++j) {
      if (i0 + j < nfb) {
        const float xm = scalar_is_one(x[j]) ? 1.0f : x[j];  // w*1.0f is w: the "scalar is one" shortcut
#pragma unroll
        for (int c = 0; c < CPL; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(w[j][c], xm));
        norm = __fadd_rn(norm, __fmul_rn(x[j], x[j]));
      }
    }
  }
  if (!m.no_user_bias) {
    for (int base = 0; base < nfb; base += 32) {
      const int i = base + lane;
      float p = 0.0f;
      if (i < nfb) p = __fmul_rn(__ldcg(m.bias + fi[i]), fv[i]);
      const int cnt = min(32, nfb - base);
      for (int j = 0; j < cnt; ++j) fb_bias = __fadd_rn(fb_bias, __shfl_sync(0xffffffffu, p, j));
    }
  }
  return true;
}
// shared-memory form: the k-vector goes to xch[0..pitch) for a lane group to pick up
template <int CPL>
__device__ __forceinline__ bool coop_prepare_ufeedback(const DevModel &m, const unsigned *fi, const float *fv,
                                                       int nfb, int lane, float *xch, float &norm, float &fb_bias,
                                                       int *err_flag) {
  float acc[CPL];
  if (!coop_prepare_ufeedback_regs<CPL>(m, fi, fv, nfb, lane, acc, norm, fb_bias, err_flag)) return false;
#pragma unroll
  for (int c = 0; c < CPL; ++c) xch[lane * CPL + c] = acc[c];
  __syncwarp();
  return true;
}

// base.h:539-554, all 32 lanes; xch holds d = (tmp_ufeedback - old) / norm, dbias its bias part
template <int CPL>
__device__ __forceinline__ void coop_update_ufeedback_regs(const DevModel &m, const unsigned *fi, const float *fv,
                                                           int nfb, int lane, const float (&d)[CPL], float dbias,
                                                           int scatter) {
  if (nfb == 0) return;
  bool bad = false;
  for (int i = lane; i + 1 < nfb; i += 32) bad |= !(fi[i] < fi[i + 1]);
  const bool unique = !__any_sync(0xffffffffu, bad);
  const bool use_red = unique && scatter == SCATTER_RED;
#pragma unroll 4
  for (int i = 0; i < nfb; ++i) {
    const float x = fv[i];
    const float xm = scalar_is_one(x) ? 1.0f : x;
    float *p = m.W + (size_t)fi[i] * (size_t)m.pitch + lane * CPL;
    float v[CPL];
    if (use_red) {
#pragma unroll
      for (int c = 0; c < CPL; ++c) v[c] = __fmul_rn(d[c], xm);
      red_cpl<CPL>(p, v);
    } else {
      ld_cpl<CPL>(p, v);
#pragma unroll
      for (int c = 0; c < CPL; ++c) v[c] = __fadd_rn(v[c], __fmul_rn(d[c], xm));
      st_cpl<CPL>(p, v);
    }
  }
  if (!m.no_user_bias) {
    if (unique) {
      for (int i = lane; i < nfb; i += 32) {
        float *p = m.bias + fi[i];
        const float add = __fmul_rn(dbias, fv[i]);
        if (use_red) red1(p, add);
        else __stcg(p, __fadd_rn(__ldcg(p), add));
      }
    } else {
      if (lane == 0)
        for (int i = 0; i < nfb; ++i) {
          float *p = m.bias + fi[i];
          __stcg(p, __fadd_rn(__ldcg(p), __fmul_rn(dbias, fv[i])));
        }
    }
  }
  __syncwarp();
}
template <int CPL>
__device__ __forceinline__ void coop_update_ufeedback(const DevModel &m, const unsigned *fi, const float *fv,
                                                      int nfb, int lane, const float *xch, float dbias,
                                                      int scatter) {
  if (nfb == 0) return;
  float d[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) d[c] = xch[lane * CPL + c];
  coop_update_ufeedback_regs<CPL>(m, fi, fv, nfb, lane, d, dbias, scatter);
}

}  // namespace svdk
