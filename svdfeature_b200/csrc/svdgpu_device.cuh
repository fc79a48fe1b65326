// svdgpu_device.cuh -- device code of the SGD hot path (sm_100a).
//
// One lane GROUP (LANES = pitch/4 lanes, <= 32, each lane owning VEC float4
// chunks of a factor row) executes one training instance: the whole of
// SVDFeature::update_inner (base.h:456-462) -- bias gather, row gather, dot,
// loss gradient, scatter update, L2 decay -- with the rows held in registers
// (the reference makes ~7 passes over k floats through memory).
//
// Arithmetic contract (identical to oracle/svdf_oracle.c and the reference):
//   fp32 rows, separate multiply and add (compiled with --fmad=false), fp64
//   accumulation of base+bias+dot in the reference's order, the |s-1|<=1e-6
//   "scalar is one" shortcut, and -- when EXACT_DOT -- the reference's 4-lane
//   strided dot order (sse.h:289-317).  File:line citations use the short names
//   base.h = solvers/base-solver/apex_svd_base.h, model.h = apex_svd_model.h,
//   sse.h = apex-tensor/apex_tensor_sse.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace svdk {

// ---------------------------------------------------------------------------
// parameter blocks passed by value to every kernel
// ---------------------------------------------------------------------------
struct DevModel {
  float *W;           // [rows][pitch]   feedback rows | user rows | item rows
  float *bias;        // [rows]
  float *g_bias;      // [num_global]
  unsigned *ver_ui;   // [rows]          row version counters (exact mode)
  unsigned *ver_g;    // [num_global]
  int pitch;          // floats per row, multiple of 4
  int k;              // num_factor
  int num_user, num_item, num_global, num_ufeedback;
  int user_off, item_off;  // first user / item row in the slab
  int no_user_bias;
  int active_type;
};

// ParameterSet (base.h:33-75): per-index-range weight decay ("up:" / "ip:" / "uip:" / "gp:" keys).
// bound[j] is the INCLUSIVE last index of range j (the reference stores bound-1, base.h:59).
enum { WD_MAX_RANGES = 8 };
struct WdRanges {
  int n;  // 0: no ranges, the default weight decay applies to every index
  unsigned bound[WD_MAX_RANGES];
  float wd[WD_MAX_RANGES];
};

struct DevHP {
  float lr;
  float du, di;       // 1 - lr*wd_user, 1 - lr*wd_item              (base.h:216,257)
  float dub, dib;     // 1 - lr*wd_user_bias, 1 - lr*wd_item_bias    (base.h:248,282)
  float dg;           // 1 - lr*wd_global                            (base.h:192)
  int du_skip, di_skip;  // row decay skipped: |d-1| <= 1e-6         (sse.h:231-242)
  unsigned regfree;   // num_regfree_global                          (base.h:190)
  float base_score;
  float lr_fb;        // lr * scale_lr_ufeedback                     (base.h:513)
  float dfb;          // 1 - lr_fb*wd_ufeedback                      (base.h:515)
  int dfb_skip;
  float dfbb;         // 1 - lr_fb*wd_ufeedback_bias                 (base.h:518)
  // regularisers other than L2 decay (base.h:188-283)
  int reg_user;       // row regulariser of user rows: 0 L2 decay, 1 L1 soft threshold, 2 projection
  int reg_item;       //   ... of item rows (reg_method 3 = L1 on users, L2 decay on items)
  int reg_global;     // 0 L2 decay, 1 L1                            (base.h:192-193)
  float l1_u, l1_i;   // lr*wd_user, lr*wd_item: the L1 thresholds   (base.h:214,254)
  float pb_u, pb_i;   // wd_user, wd_item: the projection bounds B   (base.h:181-186,226,266)
  float l1_g;         // lr*wd_global                                (base.h:189)
  int user_nonneg;    // model.param.user_nonnegative                (base.h:242-245)
  int plain;          // reg_user == reg_item == reg_global == 0 and !user_nonneg and no ranged weight decay
  WdRanges ru, ri, rg;  // ranged weight decay of user rows / item rows / global biases (base.h:189,213,253)
};

// A CSR batch in HBM.  index/value/ticket hold the elements [val_base, ...) of
// the caller's arrays, so a row_ptr value p addresses index[p - val_base].
struct DevCsr {
  const int *row_ptr;
  const float *label;
  const unsigned *index;
  const float *value;
  const float *value2;     // side-feature expansion only: second factor of an item entry's value, else null
  const unsigned *ticket;  // exact mode: row version each feature waits for
  int val_base;
  int val_end;  // one past the last feature position the arrays hold (absolute)
  float avg_nnz = 0.0f;  // features per row over the whole arrays, if the host knows it (tile sizing)
};

enum { SCATTER_STORE = 0, SCATTER_RED = 1 };
enum { ERR_NONE = 0, ERR_GLOBAL_INDEX = 1, ERR_USER_INDEX = 2, ERR_ITEM_INDEX = 3, ERR_FB_INDEX = 4, ERR_ROW_PTR = 5,
       ERR_WD_BOUND = 6, ERR_TIMEOUT = 7 };

// a row's four segment bounds must be ordered and inside the batch (checked on the
// device so that the host never walks row_ptr)
__device__ __forceinline__ bool row_ok(int rp0, int rp1, int rp2, int rp3, int lo, int hi) {
  return rp0 >= lo && rp0 <= rp1 && rp1 <= rp2 && rp2 <= rp3 && rp3 <= hi;
}

// ---------------------------------------------------------------------------
// small PTX helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ float4 ldcg4(const float *p) {
  return __ldcg(reinterpret_cast<const float4 *>(p));
}
__device__ __forceinline__ void stcg4(float *p, float4 v) {
  __stcg(reinterpret_cast<float4 *>(p), v);
}
// vector reduction into global memory (sm_90+): one 16-byte atomic add per lane
__device__ __forceinline__ void red4(float *p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void red1(float *p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned *p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// pull one line into L2 without occupying a register or the scoreboard
__device__ __forceinline__ void prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// same, with a suspend-time hint: the (single) producer lane mostly waits for consumers
// to free a stage; a long hardware sleep keeps it out of the issue slots
__device__ __forceinline__ void mbar_wait_sleepy(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAITS_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONES_%=;\n"
      "bra WAITS_%=;\n"
      "DONES_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
// 1-D bulk asynchronous copy global -> shared (TMA unit, SASS UBLKCP), completion
// signalled on an mbarrier.  dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---------------------------------------------------------------------------
// scalar pieces of the reference
// ---------------------------------------------------------------------------
// sse.h:231-242: a row scaled by s with |s-1| <= 1e-6 is not multiplied
__device__ __forceinline__ bool scalar_is_one(float s) {
  // fabs((double)(s-1.0f)) > 1e-6  <=>  fabsf(s-1.0f) > 1e-6f, because (float)1e-6 < 1e-6
  return !(fabsf(__fsub_rn(s, 1.0f)) > 1e-6f);
}
// ParameterSet::get_wd (base.h:69-74): lower_bound over the inclusive range ends; an index beyond
// the last range is the reference's "bound set err".  Branch-free selects: hp lives in the
// kernel parameter bank and must not be indexed dynamically.
__device__ __forceinline__ bool wd_lookup(const WdRanges &r, unsigned id, float &wd) {
  int k = 0;
#pragma unroll
  for (int j = 0; j < WD_MAX_RANGES; ++j) k += (j < r.n && r.bound[j] < id) ? 1 : 0;
  float w = 0.0f;
#pragma unroll
  for (int j = 0; j < WD_MAX_RANGES; ++j) w = (j == k) ? r.wd[j] : w;
  wd = w;
  return k < r.n;
}
// the four constants reg_user / reg_item derive from one weight decay (base.h:213-226,253-266)
struct RowReg {
  float d;   // 1 - lr*wd
  int skip;  // d is "one": no multiply
  float l1;  // lr*wd
  float pb;  // wd
};
__device__ __forceinline__ RowReg row_reg(const WdRanges &r, unsigned id, float lr, float d, int skip, float l1,
                                          float pb, int *err_flag) {
  RowReg o = {d, skip, l1, pb};
  if (r.n) {
    float wd;
    if (!wd_lookup(r, id, wd)) {
      atomicCAS(err_flag, 0, ERR_WD_BOUND);
      return o;
    }
    o.pb = wd;
    o.l1 = __fmul_rn(lr, wd);
    o.d = __fsub_rn(1.0f, o.l1);
    o.skip = !(fabsf(__fsub_rn(o.d, 1.0f)) > 1e-6f);
  }
  return o;
}
// decay constant of one global bias (base.h:189-193): 1 - lr*wd, or lr*wd for the L1 variant
__device__ __forceinline__ float global_dec(const WdRanges &r, unsigned gid, float lr, bool l1, float dflt,
                                            int *err_flag) {
  if (!r.n) return dflt;
  float wd;
  if (!wd_lookup(r, gid, wd)) {
    atomicCAS(err_flag, 0, ERR_WD_BOUND);
    return dflt;
  }
  const float lambda = __fmul_rn(lr, wd);
  return l1 ? lambda : __fsub_rn(1.0f, lambda);
}
// expf: the reference calls glibc expf (<= 0.502 ulp).  exp() in fp64 rounded to
// fp32 is correctly rounded in all but ~1e-8 of cases, i.e. equal to glibc's in
// all but the ~0.2% of arguments where glibc itself is 1 ulp off.
__device__ __forceinline__ float ref_expf(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float sigmoidf_ref(float x) {
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, ref_expf(-x)));
}
// model.h:112-123
__device__ __forceinline__ float map_active(float sum, int type) {
  return (type == 1 || type == 2) ? sigmoidf_ref(sum) : sum;
}
__device__ __forceinline__ float smooth_hinge_grad(float z) {  // model.h:90-94
  if (z > 1.0f) return 0.0f;
  if (z < 0.0f) return 1.0f;
  return __fsub_rn(1.0f, z);
}
// model.h:132-156
__device__ __forceinline__ float cal_grad(float r, float pred, int type) {
  switch (type) {
    case 1: return __fmul_rn(__fmul_rn(__fsub_rn(r, pred), pred), __fsub_rn(1.0f, pred));
    case 3:
    case 7: return __fsub_rn(r, sigmoidf_ref(pred));
    case 5:
      if (r > 0.5f) return smooth_hinge_grad(__fsub_rn(pred, 0.5f));
      return -smooth_hinge_grad(__fsub_rn(0.5f, pred));
    case 6:
      if (r > 0.5f) return pred > 1.0f ? 0.0f : __fsub_rn(r, pred);
      return pred < 0.0f ? 0.0f : __fsub_rn(r, pred);
    default: return __fsub_rn(r, pred);  // 0 and 2
  }
}

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
// t + w*s with the reference's shortcut (row_add_scaled in the oracle)
__device__ __forceinline__ float4 f4_add_scaled(float4 t, float4 w, float s, bool one) {
  if (one) {
    t.x = __fadd_rn(t.x, w.x); t.y = __fadd_rn(t.y, w.y);
    t.z = __fadd_rn(t.z, w.z); t.w = __fadd_rn(t.w, w.w);
  } else {
    t.x = __fadd_rn(t.x, __fmul_rn(w.x, s)); t.y = __fadd_rn(t.y, __fmul_rn(w.y, s));
    t.z = __fadd_rn(t.z, __fmul_rn(w.z, s)); t.w = __fadd_rn(t.w, __fmul_rn(w.w, s));
  }
  return t;
}
__device__ __forceinline__ float4 f4_scale(float4 t, float s) {
  t.x = __fmul_rn(t.x, s); t.y = __fmul_rn(t.y, s);
  t.z = __fmul_rn(t.z, s); t.w = __fmul_rn(t.w, s);
  return t;
}
// base.h:175-180 / apex_tensor_cpu_inline_common.h:168-175: L1 soft threshold
__device__ __forceinline__ float reg_l1(float w, float eps) {
  if (w > eps) return __fsub_rn(w, eps);
  if (w < -eps) return __fadd_rn(w, eps);
  return 0.0f;
}
__device__ __forceinline__ float4 f4_reg_l1(float4 t, float eps) {
  return make_float4(reg_l1(t.x, eps), reg_l1(t.y, eps), reg_l1(t.z, eps), reg_l1(t.w, eps));
}
// tensor::smaller_then_fill(w, 0): w <= 0 -> 0 (base.h:242-245)
__device__ __forceinline__ float4 f4_nonneg(float4 t) {
  return make_float4(t.x <= 0.0f ? 0.0f : t.x, t.y <= 0.0f ? 0.0f : t.y, t.z <= 0.0f ? 0.0f : t.z,
                     t.w <= 0.0f ? 0.0f : t.w);
}
__device__ __forceinline__ float4 f4_sub(float4 a, float4 b) {
  return make_float4(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z),
                     __fsub_rn(a.w, b.w));
}

// ---------------------------------------------------------------------------
// the lane group
// ---------------------------------------------------------------------------
template <int LANES, int VEC>
struct Group {
  static constexpr int NCH = LANES * VEC;     // float4 chunks per row covered
  static constexpr int DOT_STRIDE = NCH + 4;  // padded component stride in smem
  static constexpr int DOT_FLOATS = 4 * DOT_STRIDE;

  int gl;          // lane within the group
  unsigned gmask;  // lanes of this group
  float *dot_s;    // DOT_FLOATS floats of shared scratch private to the group

  __device__ __forceinline__ void gsync() const { __syncwarp(gmask); }
  template <typename T>
  __device__ __forceinline__ T bcast(T v, int src) const {
    return __shfl_sync(gmask, v, src, LANES);
  }

  // ---- row access ------------------------------------------------------
  __device__ __forceinline__ void load_row(const DevModel &m, size_t row, float4 (&w)[VEC]) const {
    const float *p = m.W + row * (size_t)m.pitch;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int c = gl + v * LANES;
      w[v] = (4 * c < m.pitch) ? ldcg4(p + 4 * c) : f4_zero();
    }
  }
  __device__ __forceinline__ void store_row(const DevModel &m, size_t row, const float4 (&w)[VEC]) const {
    float *p = m.W + row * (size_t)m.pitch;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int c = gl + v * LANES;
      if (4 * c < m.pitch) stcg4(p + 4 * c, w[v]);
    }
  }
  __device__ __forceinline__ void red_row(const DevModel &m, size_t row, const float4 (&nw)[VEC],
                                          const float4 (&old)[VEC]) const {
    float *p = m.W + row * (size_t)m.pitch;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int c = gl + v * LANES;
      if (4 * c < m.pitch) red4(p + 4 * c, f4_sub(nw[v], old[v]));
    }
  }

  // ---- dot product -------------------------------------------------------
  // Reference order (sse.h:289-317): the k/4 SSE vectors a[4i..4i+3]*b[4i..4i+3]
  // are accumulated in order i = 0,1,2,... into one 4-wide accumulator, which is
  // then folded (l0+l2)+(l1+l3); k%4 tail elements are added serially after.
  // Chunk (v,lane) of this group IS SSE vector i = v*LANES + lane, so component c
  // of every chunk goes to dot_s[c*DOT_STRIDE + i] and lane c adds its NCH values
  // in order.
  template <bool EXACT>
  __device__ __forceinline__ float dot(const DevModel &m, const float4 (&a)[VEC],
                                       const float4 (&b)[VEC]) const {
    if (EXACT) {
      const int nfull = m.k >> 2;  // full SSE vectors
      gsync();
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int i = gl + v * LANES;
        dot_s[0 * DOT_STRIDE + i] = __fmul_rn(a[v].x, b[v].x);
        dot_s[1 * DOT_STRIDE + i] = __fmul_rn(a[v].y, b[v].y);
        dot_s[2 * DOT_STRIDE + i] = __fmul_rn(a[v].z, b[v].z);
        dot_s[3 * DOT_STRIDE + i] = __fmul_rn(a[v].w, b[v].w);
      }
      gsync();
      float acc = 0.0f;
      if (gl < 4) {
        const float *src = dot_s + gl * DOT_STRIDE;
        int i = 0;
        for (; i + 4 <= nfull; i += 4) {
          const float4 q = *reinterpret_cast<const float4 *>(src + i);
          acc = __fadd_rn(acc, q.x); acc = __fadd_rn(acc, q.y);
          acc = __fadd_rn(acc, q.z); acc = __fadd_rn(acc, q.w);
        }
        for (; i < nfull; ++i) acc = __fadd_rn(acc, src[i]);
      }
      const float l0 = bcast(acc, 0), l1 = bcast(acc, 1), l2 = bcast(acc, 2), l3 = bcast(acc, 3);
      float sum = __fadd_rn(__fadd_rn(l0, l2), __fadd_rn(l1, l3));
      const int tail = m.k & 3;  // elements nfull*4 .. k-1 live in chunk nfull
      for (int t = 0; t < tail; ++t) sum = __fadd_rn(sum, dot_s[t * DOT_STRIDE + nfull]);
      return sum;
    } else {
      float acc = 0.0f;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float s = __fadd_rn(__fadd_rn(__fmul_rn(a[v].x, b[v].x), __fmul_rn(a[v].z, b[v].z)),
                            __fadd_rn(__fmul_rn(a[v].y, b[v].y), __fmul_rn(a[v].w, b[v].w)));
        acc = __fadd_rn(acc, s);
      }
#pragma unroll
      for (int o = LANES / 2; o > 0; o >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(gmask, acc, o, LANES));
      return acc;
    }
  }

  // ---- row regulariser (the factor part of reg_user / reg_item, base.h:211-283) ----
  // method 0: w *= d unless d is "one"; 1: L1 soft threshold eps; 2: projection onto
  // |w|^2 <= B (base.h:181-186: sum = dot(w,w); if sum > B, w *= sqrtf(B/sum)).
  // Pad floats beyond k are zero and stay zero under all three.
  template <bool EXACT>
  __device__ __forceinline__ void reg_row(const DevModel &m, float4 (&w)[VEC], int method, float d,
                                          int d_skip, float eps, float B, bool nonneg) const {
    if (method == 0) {
      if (!d_skip) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) w[v] = f4_scale(w[v], d);
      }
    } else if (method == 1) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) w[v] = f4_reg_l1(w[v], eps);
    } else {
      const float sum = dot<EXACT>(m, w, w);
      if (sum > B) {
        const float sc = __fsqrt_rn(__fdiv_rn(B, sum));
        if (!scalar_is_one(sc)) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) w[v] = f4_scale(w[v], sc);
        }
      }
    }
    if (nonneg) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) w[v] = f4_nonneg(w[v]);
    }
  }

  // ---- ordered fp64 sum of val[f]*table[off+idx[f]] over f in [beg,end) -----
  // base.h:318-322,325-334,340-350: each product is fp32, the running sum fp64,
  // added in feature order.  Lanes fetch LANES features at a time; the adds are
  // replayed in order through shuffles so every lane ends with the same sum.
  __device__ __forceinline__ double bias_sum(double sum, const float *table, int off,
                                            const unsigned *idx, const float *val, int beg,
                                            int end, const float *val2 = nullptr) const {
    for (int base = beg; base < end; base += LANES) {
      const int f = base + gl;
      float p = 0.0f;
      if (f < end) {
        p = __fmul_rn(val[f], __ldcg(table + off + idx[f]));
        if (val2) p = __fmul_rn(p, val2[f]);  // side feature of an item: (i_bias*value)*ival, base.h:348
      }
      const int cnt = min(LANES, end - base);
      for (int j = 0; j < cnt; ++j) sum = __dadd_rn(sum, (double)bcast(p, j));
    }
    return sum;
  }

  // any duplicate index inside [beg,end)?  (sequential semantics matter then)
  __device__ __forceinline__ bool has_dup(const unsigned *idx, int beg, int end) const {
    const int n = end - beg;
    if (n <= 1) return false;
    if (n > LANES) return true;  // conservative: take the unfused path
    const unsigned mine = (gl < n) ? idx[beg + gl] : 0xffffffffu;
    bool dup = false;
    for (int j = 0; j < n - 1; ++j) {
      const unsigned other = bcast(mine, j);
      dup |= (gl > j && gl < n && other == mine);
    }
    return __any_sync(gmask, dup);
  }

  // scalar table RMW over a feature segment: x += lrerr*val (if upd), x *= decay (if dec)
  // parallel over lanes when the segment has no duplicate index, else serial on lane 0.
  __device__ __forceinline__ void scalar_seg(float *table, int off, const unsigned *idx,
                                             const float *val, int beg, int end, float lrerr,
                                             float decay, bool upd, bool dec, unsigned regfree,
                                             bool parallel, int scatter, bool l1 = false,
                                             const float *val2 = nullptr, const WdRanges *rng = nullptr,
                                             float lr = 0.0f, int *err_flag = nullptr) const {
    // rng (globals only): the decay constant comes from the index's weight-decay range
    // l1: the decay step is reg_L1(x, decay) instead of x *= decay (reg_global 1, base.h:193)
    if (parallel) {
      for (int f = beg + gl; f < end; f += LANES) {
        float *p = table + off + idx[f];
        const float x0 = __ldcg(p);
        float x = x0;
        if (upd) x = __fadd_rn(x, val2 ? __fmul_rn(__fmul_rn(lrerr, val[f]), val2[f]) : __fmul_rn(lrerr, val[f]));
        if (dec && rng) decay = global_dec(*rng, idx[f], lr, l1, decay, err_flag);
        if (dec && idx[f] >= regfree) x = l1 ? reg_l1(x, decay) : __fmul_rn(x, decay);
        if (scatter == SCATTER_RED) red1(p, __fsub_rn(x, x0));
        else __stcg(p, x);
      }
    } else {
      if (gl == 0) {
        for (int f = beg; f < end; ++f) {
          float *p = table + off + idx[f];
          float x = __ldcg(p);
          if (upd) x = __fadd_rn(x, val2 ? __fmul_rn(__fmul_rn(lrerr, val[f]), val2[f]) : __fmul_rn(lrerr, val[f]));
          if (dec && rng) decay = global_dec(*rng, idx[f], lr, l1, decay, err_flag);
          if (dec && idx[f] >= regfree) x = l1 ? reg_l1(x, decay) : __fmul_rn(x, decay);
          __stcg(p, x);
        }
      }
      gsync();
    }
  }
};

// SVD++ per-user state carried across the rows of one block (base.h:486-488)
template <int VEC>
struct FbState {
  float4 fb[VEC];  // tmp_ufeedback (this lane's chunks)
  float fb_bias;   // tmp_ufeedback_bias
  float norm;      // norm_ufeedback
};

// ---------------------------------------------------------------------------
// one training / prediction instance
// ---------------------------------------------------------------------------
// idx/val are addressed with ABSOLUTE row_ptr values (the caller pre-offsets the
// pointers); rp0..rp3 are the four segment bounds of the row.
// Returns the prediction (base.h:445-454); when TRAIN also applies the update.
template <int LANES, int VEC, bool EXACT_DOT, bool TRAIN, bool SVDPP>
__device__ __forceinline__ float process_instance(const Group<LANES, VEC> &g, const DevModel &m,
                                                  const DevHP &hp, int rp0, int rp1, int rp2,
                                                  int rp3, float label, const unsigned *idx,
                                                  const float *val, int scatter_user,
                                                  int scatter_item, FbState<VEC> *fbs,
                                                  int *err_flag, const float *val2 = nullptr) {
  // val2 (side-feature expansion, base.h:375-379,417-422): an item entry's value is the pair
  // (val, val2) -- prepare_tmp uses val*val2, the update (lr*err*val)*val2 like the reference
  // ---- index bounds (base.h:320,327,343: assert_true -> error) -------------
  {
    int bad = 0;
    for (int f = rp0 + g.gl; f < rp3; f += LANES) {
      const unsigned id = idx[f];
      if (f < rp1) { if (id >= (unsigned)m.num_global) bad = ERR_GLOBAL_INDEX; }
      else if (f < rp2) { if (id >= (unsigned)m.num_user) bad = ERR_USER_INDEX; }
      else { if (id >= (unsigned)m.num_item) bad = ERR_ITEM_INDEX; }
    }
    if (__any_sync(g.gmask, bad != 0)) {
      if (bad) atomicCAS(err_flag, 0, bad);
      return 0.0f;
    }
  }

  // ---- gathers of the common small shapes, all issued before anything is consumed ----------
  // (first user row, first two item rows, their biases: one memory round trip instead of one
  // per feature; the generic routine used to load a row, use it, load the next)
  const int nu_ = rp2 - rp1, ni_ = rp3 - rp2;
  float4 tu[VEC], ti[VEC], wu0[VEC], wi0[VEC], wi1[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) wu0[v] = wi0[v] = wi1[v] = f4_zero();
  if (nu_ > 0) g.load_row(m, (size_t)m.user_off + idx[rp1], wu0);
  if (ni_ > 0) g.load_row(m, (size_t)m.item_off + idx[rp2], wi0);
  if (ni_ > 1) g.load_row(m, (size_t)m.item_off + idx[rp2 + 1], wi1);
  float pre_ub = 0.0f, pre_ib0 = 0.0f, pre_ib1 = 0.0f;
  if (nu_ > 0 && !m.no_user_bias) pre_ub = __ldcg(m.bias + m.user_off + idx[rp1]);
  if (ni_ > 0) pre_ib0 = __ldcg(m.bias + m.item_off + idx[rp2]);
  if (ni_ > 1) pre_ib1 = __ldcg(m.bias + m.item_off + idx[rp2 + 1]);

  // ---- calc_bias, global part (base.h:318-322): its gathers overlap the row gathers --------
  // Up to LANES global features (the neighbourhood rows of configs[4] carry 8): lane gl keeps the
  // bias of feature rp0+gl for the update below, so g_bias is read once per instance.
  const int ng_ = rp1 - rp0;
  const bool g_small = ng_ > 0 && ng_ <= LANES;
  float pre_gb = 0.0f, pre_gv = 0.0f;
  unsigned pre_gid = 0u;
  double bsum = 0.0;
  if (g_small) {
    float p = 0.0f;
    if (g.gl < ng_) {
      pre_gid = idx[rp0 + g.gl];
      pre_gv = val[rp0 + g.gl];
      pre_gb = __ldcg(m.g_bias + pre_gid);
      p = __fmul_rn(pre_gv, pre_gb);
    }
    for (int j = 0; j < ng_; ++j) bsum = __dadd_rn(bsum, (double)g.bcast(p, j));
  } else {
    bsum = g.bias_sum(bsum, m.g_bias, 0, idx, val, rp0, rp1);
  }

  // ---- prepare_tmp (base.h:354-381) ----------------------------------------
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    tu[v] = SVDPP ? fbs->fb[v] : f4_zero();  // prepare_svdpp, base.h:430-432 / 506-508
    ti[v] = f4_zero();
  }
  for (int f = rp1; f < rp2; ++f) {
    float4 w[VEC];
    if (f == rp1) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) w[v] = wu0[v];
    } else {
      g.load_row(m, (size_t)m.user_off + idx[f], w);
    }
    const float s = val[f];
    const bool one = scalar_is_one(s);
#pragma unroll
    for (int v = 0; v < VEC; ++v) tu[v] = f4_add_scaled(tu[v], w[v], s, one);
  }
  for (int f = rp2; f < rp3; ++f) {
    float4 w[VEC];
    if (f == rp2) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) w[v] = wi0[v];
    } else if (f == rp2 + 1) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) w[v] = wi1[v];
    } else {
      g.load_row(m, (size_t)m.item_off + idx[f], w);
    }
    const float s = val2 ? __fmul_rn(val[f], val2[f]) : val[f];
    const bool one = scalar_is_one(s);
#pragma unroll
    for (int v = 0; v < VEC; ++v) ti[v] = f4_add_scaled(ti[v], w[v], s, one);
  }

  // ---- calc_bias, user and item part (base.h:324-350), first entries from the early gathers --
  if (!m.no_user_bias) {
    if (nu_ > 0) bsum = __dadd_rn(bsum, (double)__fmul_rn(val[rp1], pre_ub));
    bsum = g.bias_sum(bsum, m.bias, m.user_off, idx, val, min(rp1 + 1, rp2), rp2);
    if (SVDPP) bsum = __dadd_rn(bsum, (double)fbs->fb_bias);  // get_bias_svdpp, base.h:509-511
  }
  if (ni_ > 0) {
    float p = __fmul_rn(val[rp2], pre_ib0);
    if (val2) p = __fmul_rn(p, val2[rp2]);
    bsum = __dadd_rn(bsum, (double)p);
  }
  if (ni_ > 1) {
    float p = __fmul_rn(val[rp2 + 1], pre_ib1);
    if (val2) p = __fmul_rn(p, val2[rp2 + 1]);
    bsum = __dadd_rn(bsum, (double)p);
  }
  bsum = g.bias_sum(bsum, m.bias, m.item_off, idx, val, min(rp2 + 2, rp3), rp3, val2);

  // ---- pred (base.h:445-454) ------------------------------------------------
  const float d = g.template dot<EXACT_DOT>(m, tu, ti);
  double sum = __dadd_rn((double)hp.base_score, bsum);
  sum = __dadd_rn(sum, (double)d);
  const float pred = map_active((float)sum, m.active_type);
  if (!TRAIN) return pred;

  // ---- update_no_decay + regularize(after) (base.h:383-427, 286-311) --------
  const float err = cal_grad(label, pred, m.active_type);
  const float lrerr = __fmul_rn(hp.lr, err);

  const bool dup_g = g.has_dup(idx, rp0, rp1);
  const bool dup_u = g.has_dup(idx, rp1, rp2);
  const bool dup_i = g.has_dup(idx, rp2, rp3);
  const bool fused = !(dup_u || dup_i);

  // globals: g_bias[gid] += lr*err*gval ; later g_bias[gid] *= 1-lr*wd_global
  const bool g_l1 = hp.reg_global == 1;
  const float g_dec = g_l1 ? hp.l1_g : hp.dg;
  if (g_small && !dup_g) {  // every global index once: update and decay from the value gathered above
    if (g.gl < ng_) {
      float x = __fadd_rn(pre_gb, __fmul_rn(lrerr, pre_gv));
      const float dec = global_dec(hp.rg, pre_gid, hp.lr, g_l1, g_dec, err_flag);
      if (pre_gid >= hp.regfree) x = g_l1 ? reg_l1(x, dec) : __fmul_rn(x, dec);
      if (scatter_item == SCATTER_RED) red1(m.g_bias + pre_gid, __fsub_rn(x, pre_gb));
      else __stcg(m.g_bias + pre_gid, x);
    }
  } else if (rp1 > rp0) {
    g.scalar_seg(m.g_bias, 0, idx, val, rp0, rp1, lrerr, g_dec, true, !dup_g, hp.regfree, !dup_g,
                 dup_g ? SCATTER_STORE : scatter_item, g_l1, nullptr, &hp.rg, hp.lr, err_flag);
  }

  // bias read-modify-write of one feature whose old value was gathered early (lane `who` does it)
  auto bias_rmw = [&](float *p, float x0, float v, float v2, bool has2, float decay, int who, int scatter) {
    if (g.gl != who) return;
    float add = __fmul_rn(lrerr, v);
    if (has2) add = __fmul_rn(add, v2);
    const float x = __fmul_rn(__fadd_rn(x0, add), decay);
    if (scatter == SCATTER_RED) red1(p, __fsub_rn(x, x0));
    else __stcg(p, x);
  };
  if (fused) {
    // every touched row appears once: update and decay in one register pass
    for (int f = rp1; f < rp2; ++f) {
      const size_t row = (size_t)m.user_off + idx[f];
      float4 w[VEC], nw[VEC];
      if (f == rp1) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) w[v] = wu0[v];
      } else {
        g.load_row(m, row, w);
      }
      const float sc = __fmul_rn(lrerr, val[f]);
      const bool one = scalar_is_one(sc);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        nw[v] = f4_add_scaled(w[v], ti[v], sc, one);
        if (hp.plain && !hp.du_skip) nw[v] = f4_scale(nw[v], hp.du);
      }
      if (!hp.plain) {
        const RowReg rr = row_reg(hp.ru, idx[f], hp.lr, hp.du, hp.du_skip, hp.l1_u, hp.pb_u, err_flag);
        g.template reg_row<EXACT_DOT>(m, nw, hp.reg_user, rr.d, rr.skip, rr.l1, rr.pb, hp.user_nonneg != 0);
      }
      if (scatter_user == SCATTER_RED) g.red_row(m, row, nw, w);
      else g.store_row(m, row, nw);
    }
    if (!m.no_user_bias) {
      if (nu_ > 0) bias_rmw(m.bias + m.user_off + idx[rp1], pre_ub, val[rp1], 1.0f, false, hp.dub, 0, scatter_user);
      g.scalar_seg(m.bias, m.user_off, idx, val, min(rp1 + 1, rp2), rp2, lrerr, hp.dub, true, true, 0u, true,
                   scatter_user);
    }
    for (int f = rp2; f < rp3; ++f) {
      const size_t row = (size_t)m.item_off + idx[f];
      float4 w[VEC], nw[VEC];
      if (f == rp2) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) w[v] = wi0[v];
      } else if (f == rp2 + 1) {  // (fused path: no row repeats, so the early gather is still current)
#pragma unroll
        for (int v = 0; v < VEC; ++v) w[v] = wi1[v];
      } else {
        g.load_row(m, row, w);
      }
      const float sc = val2 ? __fmul_rn(__fmul_rn(lrerr, val[f]), val2[f]) : __fmul_rn(lrerr, val[f]);
      const bool one = scalar_is_one(sc);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        nw[v] = f4_add_scaled(w[v], tu[v], sc, one);
        if (hp.plain && !hp.di_skip) nw[v] = f4_scale(nw[v], hp.di);
      }
      if (!hp.plain) {
        const RowReg rr = row_reg(hp.ri, idx[f], hp.lr, hp.di, hp.di_skip, hp.l1_i, hp.pb_i, err_flag);
        g.template reg_row<EXACT_DOT>(m, nw, hp.reg_item, rr.d, rr.skip, rr.l1, rr.pb, false);
      }
      if (scatter_item == SCATTER_RED) g.red_row(m, row, nw, w);
      else g.store_row(m, row, nw);
    }
    if (ni_ > 0)
      bias_rmw(m.bias + m.item_off + idx[rp2], pre_ib0, val[rp2], val2 ? val2[rp2] : 1.0f, val2 != nullptr, hp.dib,
               LANES - 1, scatter_item);
    if (ni_ > 1)
      bias_rmw(m.bias + m.item_off + idx[rp2 + 1], pre_ib1, val[rp2 + 1], val2 ? val2[rp2 + 1] : 1.0f,
               val2 != nullptr, hp.dib, LANES - 2, scatter_item);
    g.scalar_seg(m.bias, m.item_off, idx, val, min(rp2 + 2, rp3), rp3, lrerr, hp.dib, true, true, 0u, true,
                 scatter_item, false, val2);
  } else {
    // a row index repeats inside this instance: replay the reference's passes
    // through memory in its order (all updates, then all decays).
    for (int f = rp1; f < rp2; ++f) {
      const size_t row = (size_t)m.user_off + idx[f];
      float4 w[VEC];
      g.load_row(m, row, w);
      const float sc = __fmul_rn(lrerr, val[f]);
      const bool one = scalar_is_one(sc);
#pragma unroll
      for (int v = 0; v < VEC; ++v) w[v] = f4_add_scaled(w[v], ti[v], sc, one);
      g.store_row(m, row, w);
    }
    if (!m.no_user_bias)
      g.scalar_seg(m.bias, m.user_off, idx, val, rp1, rp2, lrerr, 0.f, true, false, 0u, false,
                   SCATTER_STORE);
    for (int f = rp2; f < rp3; ++f) {
      const size_t row = (size_t)m.item_off + idx[f];
      float4 w[VEC];
      g.load_row(m, row, w);
      const float sc = val2 ? __fmul_rn(__fmul_rn(lrerr, val[f]), val2[f]) : __fmul_rn(lrerr, val[f]);
      const bool one = scalar_is_one(sc);
#pragma unroll
      for (int v = 0; v < VEC; ++v) w[v] = f4_add_scaled(w[v], tu[v], sc, one);
      g.store_row(m, row, w);
    }
    g.scalar_seg(m.bias, m.item_off, idx, val, rp2, rp3, lrerr, 0.f, true, false, 0u, false,
                 SCATTER_STORE, false, val2);
  }

  // ---- update_svdpp (base.h:512-520) -----------------------------------------
  if (SVDPP) {
    const float s = __fmul_rn(__fmul_rn(hp.lr_fb, err), fbs->norm);
    const bool one = scalar_is_one(s);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      fbs->fb[v] = f4_add_scaled(fbs->fb[v], ti[v], s, one);
      if (!hp.dfb_skip) fbs->fb[v] = f4_scale(fbs->fb[v], hp.dfb);
    }
    if (!m.no_user_bias) {
      fbs->fb_bias = __fadd_rn(fbs->fb_bias, s);
      fbs->fb_bias = __fmul_rn(fbs->fb_bias, hp.dfbb);
    }
  }

  // ---- regularize(after) for the unfused cases --------------------------------
  if (dup_g) g.scalar_seg(m.g_bias, 0, idx, val, rp0, rp1, 0.f, g_dec, false, true, hp.regfree, false,
                          SCATTER_STORE, g_l1, nullptr, &hp.rg, hp.lr, err_flag);
  if (!fused) {
    for (int f = rp1; f < rp2; ++f) {
      const size_t row = (size_t)m.user_off + idx[f];
      if (!hp.plain || !hp.du_skip) {
        float4 w[VEC];
        g.load_row(m, row, w);
        const RowReg rr = row_reg(hp.ru, idx[f], hp.lr, hp.du, hp.du_skip, hp.l1_u, hp.pb_u, err_flag);
        g.template reg_row<EXACT_DOT>(m, w, hp.reg_user, rr.d, rr.skip, rr.l1, rr.pb, hp.user_nonneg != 0);
        g.store_row(m, row, w);
      }
      if (!m.no_user_bias)
        g.scalar_seg(m.bias, m.user_off, idx, val, f, f + 1, 0.f, hp.dub, false, true, 0u, false,
                     SCATTER_STORE);
    }
    for (int f = rp2; f < rp3; ++f) {
      const size_t row = (size_t)m.item_off + idx[f];
      if (!hp.plain || !hp.di_skip) {
        float4 w[VEC];
        g.load_row(m, row, w);
        const RowReg rr = row_reg(hp.ri, idx[f], hp.lr, hp.di, hp.di_skip, hp.l1_i, hp.pb_i, err_flag);
        g.template reg_row<EXACT_DOT>(m, w, hp.reg_item, rr.d, rr.skip, rr.l1, rr.pb, false);
        g.store_row(m, row, w);
      }
      g.scalar_seg(m.bias, m.item_off, idx, val, f, f + 1, 0.f, hp.dib, false, true, 0u, false,
                   SCATTER_STORE);
    }
  }
  return pred;
}

}  // namespace svdk
