// svdgpu_own.cu -- ordered ("exact") training of basic-MF rows with ITEM-OWNER warps.
//
// The reference's loop is strictly sequential (base.h:456-462): instance n+1 sees what instance n
// wrote.  Instances that share no row commute exactly, so the result of the sequential loop is
// reproduced bit for bit by any schedule that keeps, per row, the input order of the instances
// touching it.  k_exact (svdgpu_ordered.cu) hands every row from warp to warp through L2 and is
// bound by the hand-off chain of the hottest item (5.9 us per hand-off, 67 M instances/s on
// configs[1]).  Here the hot rows never move:
//
//   * every ITEM belongs to one persistent warp (its owner; LPT by popularity, svdgpu_ownplan.h)
//     which takes the instances of its items in input order from its own queue and keeps the item
//     rows in shared memory / registers for the whole launch: the chain of the hottest item runs
//     at register speed, one dot + update per link;
//   * USER rows travel between owners through L2 under a version counter per user: an instance
//     carries a ticket (= number of earlier instances of its user) and its user row may only be
//     fetched once the counter has reached it.  Checking, waiting and fetching is done by LOADER
//     lanes, never by the owner: loader lane (c, s) feeds ring slot s of owner c -- it reads the
//     queue entry, polls the version (ld.relaxed.gpu), and when the row is final issues one TMA
//     bulk copy (cp.async.bulk, 256 B at k = 64) that lands in the slot and completes its
//     mbarrier.  The owner only ever waits on shared memory;
//   * an owner publishes the user rows it has written with one release per BATCH of instances
//     (st.release.gpu of ticket+1; the fence is what costs), at once when the plan says the user
//     comes back soon, and always before it blocks -- so the oldest unfinished instance of the
//     whole launch can always proceed (all CTAs are co-resident: cooperative launch);
//   * the launch lasts as long as the chain of the hottest item, and a link of that chain is slower in
//     company: the owners of hot items get their issue port, the hottest ones their SM, to themselves (the
//     warps that would share it are dealt no items: own_plan_build below).
//
// Arithmetic = process_instance (svdgpu_device.cuh) specialised to the shape (0 | 1 | 1) with
// plain L2 decay: the same operations in the same order, the reference's dot order included.
//
// The plan (queues, tickets, item lists) is built on the device from the CSR batch: two radix
// sorts (by user for the tickets, by owner for the queues) around a host LPT of the item counts -- or,
// for the chunks of a host-pointer call, around the deal of the chunk before while it stays balanced.
#include "svdgpu_internal.h"
#include "svdgpu_ownplan.h"

#include <cub/cub.cuh>

#include <chrono>
#include <cstring>

namespace svdk {

// ---------------------------------------------------------------------------
// plan kernels
// ---------------------------------------------------------------------------
enum { OWN_BAD_ROWPTR = 1, OWN_NOT_BASIC = 2, OWN_BAD_USER = 4, OWN_BAD_ITEM = 8 };

// shape / bound checks (the reference's asserts, base.h:327,343), per-item and per-user counts
__global__ void k_own_scan(DevCsr csr, int r0, int n, int num_user, int num_item, unsigned *cnt_item,
                           unsigned *cnt_user, unsigned *key_user, unsigned *key_item, unsigned *row_id,
                           int *flag) {
  const unsigned *idx = csr.index - csr.val_base;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = r0 + i;
    const int rp0 = csr.row_ptr[3 * r], rp1 = csr.row_ptr[3 * r + 1], rp2 = csr.row_ptr[3 * r + 2],
              rp3 = csr.row_ptr[3 * r + 3];
    unsigned u = 0, it = 0;
    int bad = 0;
    if (!row_ok(rp0, rp1, rp2, rp3, csr.val_base, csr.val_end)) bad = OWN_BAD_ROWPTR;
    else if (!(rp1 == rp0 && rp2 == rp1 + 1 && rp3 == rp2 + 1)) bad = OWN_NOT_BASIC;
    else {
      u = idx[rp1];
      it = idx[rp2];
      if (u >= (unsigned)num_user) bad = OWN_BAD_USER;
      else if (it >= (unsigned)num_item) bad = OWN_BAD_ITEM;
    }
    if (bad) {
      atomicOr(flag, bad);
      u = it = 0;
    } else {
      atomicAdd(cnt_item + it, 1u);
      atomicAdd(cnt_user + u, 1u);
    }
    key_user[i] = u;
    key_item[i] = it;
    row_id[i] = (unsigned)i;
  }
}

// rows sorted by user (stable: input order inside a user): ticket = rank inside the user's run;
// bit 31 = the user's next instance follows within `gap` rows (publish at once)
__global__ void k_own_ticket(const unsigned *ku, const unsigned *rows, const unsigned *start_user, int n,
                             unsigned gap, unsigned *tick) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned u = ku[i], r = rows[i];
    unsigned t = (unsigned)i - start_user[u];
    if (i + 1 < n && ku[i + 1] == u && rows[i + 1] - r < gap) t |= 0x80000000u;
    tick[r] = t;
  }
}

__global__ void k_own_keyowner(const unsigned *key_item, const int *item_owner, int n, unsigned *key, unsigned *row_id) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    key[i] = (unsigned)item_owner[key_item[i]];
    row_id[i] = (unsigned)i;
  }
}

// one queue entry per instance, 32 bytes, in owner order (= the stable sort's output order)
struct OwnEntry {
  unsigned user, ticket;
  float label;
  unsigned item;
  unsigned slot;  // position of the item in its owner's list
  float uval, ival;
  unsigned flags;  // bit 0: publish the user row at once
};
static_assert(sizeof(OwnEntry) == 32, "queue entries are two 16-byte words");

__global__ void k_own_entries(DevCsr csr, int r0, int n, const unsigned *rows, const unsigned *tick,
                              const unsigned *item_slot, OwnEntry *out) {
  const unsigned *idx = csr.index - csr.val_base;
  const float *val = csr.value - csr.val_base;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
    const long long r = r0 + (long long)rows[q];
    const int rp1 = csr.row_ptr[3 * r + 1];
    const unsigned t = tick[rows[q]];
    OwnEntry e;
    e.user = idx[rp1];
    e.ticket = t & 0x7fffffffu;
    e.label = csr.label[r];
    e.item = idx[rp1 + 1];
    e.slot = item_slot[e.item];
    e.uval = val[rp1];
    e.ival = val[rp1 + 1];
    e.flags = t >> 31;
    uint4 *o = reinterpret_cast<uint4 *>(out + q);
    o[0] = make_uint4(e.user, e.ticket, __float_as_uint(e.label), e.item);
    o[1] = make_uint4(e.slot, __float_as_uint(e.uval), __float_as_uint(e.ival), e.flags);
  }
}

// ---------------------------------------------------------------------------
// k_own
// ---------------------------------------------------------------------------
constexpr int OWN_C = 12;                      // owner (compute) warps per CTA (12 + 3 loader warps: 136 registers each)
constexpr int OWN_SLOT_EXTRA = 48;             // [row | 16-byte bias window | 32-byte entry]
constexpr long long OWN_TIMEOUT = 6000000000LL;  // cycles (~3 s): a wait this long is a bug, not load
constexpr int own_threads(int D) { return (OWN_C + OWN_C * D / 32) * 32; }  // + one loader lane per ring slot

struct OwnArgs {
  DevModel m;
  DevHP hp;
  const uint4 *entries;  // OwnEntry[], two 16-byte words each
  const int *queue_off;
  const int *item_off;
  const unsigned *items;
  const int *batch;
  int S;                  // item rows an owner keeps in shared memory (the rest stay in L2)
  unsigned region_bytes;  // shared memory per owner
  unsigned off_items, off_ibias, off_dot, off_bars;
  unsigned off_hand;      // k_own2: the owner -> partner hand-off slots
  int *err_flag;
  unsigned *abort_flag;
  long long *stats;  // option "own_stats": per owner {cycles, cycles waiting for a slot, cycles in flushes, waits} or null
  int flags;         // OWN_F_*
  unsigned poll_ns;  // loader warps: sleep between polls when no lane made progress
};
enum { OWN_F_ACQUIRE = 1,   // loaders poll with ld.acquire.gpu (LDG + CCTL.IVALL) instead of ld.relaxed.gpu
       OWN_F_REVERSE = 2 }; // busiest owners on the highest warp ids of a CTA

__device__ __forceinline__ bool mbar_test(uint64_t *bar, unsigned parity) {  // non-blocking
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, unsigned parity) {  // may suspend briefly
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(100000u)  // suspend-time hint (ns): a waiting owner stays out of the issue slots
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned *p, unsigned v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// the version was read through the generic proxy; the row is read by the TMA unit (async proxy)
__device__ __forceinline__ void fence_proxy_async_global() {
  asm volatile("fence.proxy.async.global;" ::: "memory");
}

// Shared memory through 32-bit addresses (the generic pointers of the rest of this file make the
// compiler rebuild the shared window base inside the loop), and loop constants pinned in registers
// (kernel parameters are otherwise re-read from the constant bank every iteration).
__device__ __forceinline__ float4 lds128f(unsigned a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint4 lds128u(unsigned a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float lds32f(unsigned a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts32f(unsigned a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
// (a warp shuffle: neither the compiler nor ptxas can rematerialise its result from the constant bank)
__device__ __forceinline__ unsigned pin(unsigned x) { return __shfl_sync(0xffffffffu, x, threadIdx.x & 31); }
__device__ __forceinline__ float pin(float x) { return __uint_as_float(pin(__float_as_uint(x))); }
__device__ __forceinline__ double pin(double x) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return __longlong_as_double((long long)(((unsigned long long)pin((unsigned)(b >> 32)) << 32) | pin((unsigned)b)));
}
template <typename T>
__device__ __forceinline__ T *pin(T *p) {
  const unsigned long long b = (unsigned long long)p;
  return reinterpret_cast<T *>(((unsigned long long)pin((unsigned)(b >> 32)) << 32) | pin((unsigned)b));
}
__device__ __forceinline__ bool mbar_test_s(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive_s(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// One instance of the shape (0 | 1 | 1) with plain L2 decay: process_instance (svdgpu_device.cuh)
// with everything that cannot happen removed; same operations, same order (base.h:313-462).
template <int VEC>
__device__ __forceinline__ void own_step(const Group<32, VEC> &g, const DevModel &m, const DevHP &hp,
                                         float4 (&wu)[VEC], float &ub, float4 (&wi)[VEC], float &ib,
                                         float uval, float ival, float label) {
  const bool uone = scalar_is_one(uval), ione = scalar_is_one(ival);
  float4 tu[VEC], ti[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {  // prepare_tmp, base.h:354-381
    tu[v] = f4_add_scaled(f4_zero(), wu[v], uval, uone);
    ti[v] = f4_add_scaled(f4_zero(), wi[v], ival, ione);
  }
  double bsum = 0.0;  // calc_bias, base.h:313-353
  if (!m.no_user_bias) bsum = __dadd_rn(bsum, (double)__fmul_rn(uval, ub));
  bsum = __dadd_rn(bsum, (double)__fmul_rn(ival, ib));
  const float d = g.template dot<true>(m, tu, ti);
  double sum = __dadd_rn((double)hp.base_score, bsum);  // pred, base.h:445-454
  sum = __dadd_rn(sum, (double)d);
  const float p = map_active((float)sum, m.active_type);
  const float err = cal_grad(label, p, m.active_type);
  const float lrerr = __fmul_rn(hp.lr, err);
  const float su = __fmul_rn(lrerr, uval), si = __fmul_rn(lrerr, ival);  // base.h:391,412
  const bool su1 = scalar_is_one(su), si1 = scalar_is_one(si);
#pragma unroll
  for (int v = 0; v < VEC; ++v) {  // update_no_decay + regularize(after), base.h:383-427,211-283
    float4 nu = f4_add_scaled(wu[v], ti[v], su, su1);
    float4 ni = f4_add_scaled(wi[v], tu[v], si, si1);
    if (!hp.du_skip) nu = f4_scale(nu, hp.du);
    if (!hp.di_skip) ni = f4_scale(ni, hp.di);
    wu[v] = nu;
    wi[v] = ni;
  }
  if (!m.no_user_bias) ub = __fmul_rn(__fadd_rn(ub, su), hp.dub);
  ib = __fmul_rn(__fadd_rn(ib, si), hp.dib);
}

// Loader lane (c, s) of a CTA with C owners: feeds ring slot s of owner c.  `t` = index of the thread
// among the CTA's loader threads.
template <int D>
__device__ __forceinline__ void own_loader(const OwnArgs &a, unsigned char *own_smem, int t, int C) {
  const DevModel &m = a.m;
  const unsigned row_bytes = (unsigned)m.pitch * 4u, slot_bytes = row_bytes + OWN_SLOT_EXTRA;
  unsigned *const ver = m.ver_ui + m.user_off;
    // =========================== loader lane: ring slot s of owner c ===========================
    // Per iteration ONE round trip to L2 for the whole warp: the poll of the version the current
    // entry waits for and the fetch of the entry after it are issued together, consumed together.
    const int c = t / D, s = t % D;
    const int w = ((a.flags & OWN_F_REVERSE) ? C - 1 - c : c) * (int)gridDim.x + (int)blockIdx.x;
    unsigned char *reg = own_smem + (size_t)c * a.region_bytes;
    unsigned char *slot = reg + (size_t)s * slot_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(reg + a.off_bars) + s;
    uint64_t *empty = full + D;
    const int q0 = a.queue_off[w], n = a.queue_off[w + 1] - q0;
    const uint4 *q = a.entries + 2 * (size_t)q0;
    int j = s;  // this lane loads entries s, s+D, s+2D, ... of the owner's queue
    unsigned fill = 0, idle = 0;
    bool ready = false, vacant = false, have_nxt = false;
    uint4 e0 = make_uint4(0, 0, 0, 0), e1 = e0, n0 = e0, n1 = e0;
    if (j < n) {
      e0 = __ldg(q + 2 * (size_t)j);
      e1 = __ldg(q + 2 * (size_t)j + 1);
    }
    for (unsigned it = 0;; ++it) {
      const bool active = j < n;
      if (!__any_sync(0xffffffffu, active)) break;
      bool prog = false;
      // ---- issue
      const bool want_nxt = active && !have_nxt && j + D < n;
      if (want_nxt) {
        n0 = __ldg(q + 2 * (size_t)(j + D));
        n1 = __ldg(q + 2 * (size_t)(j + D) + 1);
      }
      // the row is final once every earlier instance of this user has been published; nobody
      // writes it again before this very instance does
      unsigned v = 0;
      const bool poll = active && !ready;
      // (relaxed poll: the row itself is then read by the TMA unit straight from L2, after this load
      // has returned the awaited value -- an acquire would only add an L1 invalidation, CCTL.IVALL,
      // per poll, which stalls the shared-memory pipe of the owners on this SM)
      if (poll) v = (a.flags & OWN_F_ACQUIRE) ? ld_acquire_u32(ver + e0.x) : ld_relaxed_u32(ver + e0.x);
      // the slot is vacant once the owner has read fill-1 out of it (a fresh barrier passes parity 1)
      if (active && !vacant) vacant = mbar_test(empty, (fill & 1u) ^ 1u);
      // ---- consume
      if (want_nxt) have_nxt = true;
      if (poll) ready = v == e0.y;
      if (active && vacant && ready) {
        *reinterpret_cast<uint4 *>(slot + row_bytes + 16) = e0;
        *reinterpret_cast<uint4 *>(slot + row_bytes + 32) = e1;
        fence_proxy_async_global();
        const size_t row = (size_t)m.user_off + e0.x;
        mbar_arrive_expect_tx(full, row_bytes + (m.no_user_bias ? 0u : 16u));
        bulk_g2s(slot, m.W + row * (size_t)m.pitch, row_bytes, full);
        if (!m.no_user_bias) bulk_g2s(slot + row_bytes, m.bias + (row & ~(size_t)3), 16u, full);
        j += D;
        ++fill;
        e0 = n0;
        e1 = n1;
        ready = vacant = false;
        if (!have_nxt && j < n) {  // (the entry after was not fetched yet: cold start)
          e0 = __ldg(q + 2 * (size_t)j);
          e1 = __ldg(q + 2 * (size_t)j + 1);
        }
        have_nxt = false;
        prog = true;
      }
      if (__any_sync(0xffffffffu, prog)) {
        idle = 0;
      } else {
        ++idle;
        if (idle > 2) __nanosleep(a.poll_ns);
      }
      if ((it & 63u) == 63u && ld_relaxed_u32(a.abort_flag)) break;
    }
}

// CH > 0: rows of exactly CH chunks, linear loss (fast link); CH = 0: any row width up to
// 32*VEC chunks, any loss.  D = ring slots per owner (a power of two).
template <int CH, int VEC, int D>
__global__ void __launch_bounds__(own_threads(D), 1) k_own(const OwnArgs a) {
  extern __shared__ __align__(128) unsigned char own_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const DevModel &m = a.m;
  const unsigned row_bytes = (unsigned)m.pitch * 4u, slot_bytes = row_bytes + OWN_SLOT_EXTRA;
  const int S = a.S;
  // every owner's barriers: full[D] then empty[D], one arrival each
  if (warp < OWN_C) {
    uint64_t *bars = reinterpret_cast<uint64_t *>(own_smem + (size_t)warp * a.region_bytes + a.off_bars);
    for (int i = lane; i < 2 * D; i += 32) mbar_init(bars + i, 1);
  }
  mbar_fence_init();
  __syncthreads();
  unsigned *const ver = m.ver_ui + m.user_off;

  if (warp >= OWN_C) {
    own_loader<D>(a, own_smem, (int)threadIdx.x - OWN_C * 32, OWN_C);
    return;
  }

  // ================================ owner warp ================================================
  const int w = ((a.flags & OWN_F_REVERSE) ? OWN_C - 1 - warp : warp) * (int)gridDim.x + (int)blockIdx.x;
  unsigned char *reg = own_smem + (size_t)warp * a.region_bytes;
  float *items_s = reinterpret_cast<float *>(reg + a.off_items);
  float *ibias_s = reinterpret_cast<float *>(reg + a.off_ibias);
  uint64_t *full = reinterpret_cast<uint64_t *>(reg + a.off_bars);
  uint64_t *empty = full + D;
  Group<32, VEC> g;
  g.gl = lane;
  g.gmask = 0xffffffffu;
  g.dot_s = reinterpret_cast<float *>(reg + a.off_dot);

  // the owner's item rows: the first S (its most popular) live in shared memory for the launch
  const int it0 = a.item_off[w], nit = a.item_off[w + 1] - it0, nres = min(nit, S);
  for (int sl = 0; sl < nres; ++sl) {
    const size_t row = (size_t)m.item_off + a.items[it0 + sl];
    const float *src = m.W + row * (size_t)m.pitch;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int ch = lane + v * 32;
      if (4 * ch < m.pitch) *reinterpret_cast<float4 *>(items_s + (size_t)sl * m.pitch + 4 * ch) = ldcg4(src + 4 * ch);
    }
    if (lane == 0) ibias_s[sl] = __ldcg(m.bias + row);
  }
  __syncwarp();

  float4 wi[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) wi[v] = f4_zero();
  float ib = 0.0f;
  unsigned cur_item = 0xffffffffu, cur_slot = 0;
  auto put_item = [&]() {  // the current item row leaves the registers
    if (cur_item == 0xffffffffu) return;
    if ((int)cur_slot < S) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int ch = lane + v * 32;
        if (4 * ch < m.pitch) *reinterpret_cast<float4 *>(items_s + (size_t)cur_slot * m.pitch + 4 * ch) = wi[v];
      }
      if (lane == 0) ibias_s[cur_slot] = ib;
    } else {
      g.store_row(m, (size_t)m.item_off + cur_item, wi);
      if (lane == 0) __stcg(m.bias + m.item_off + cur_item, ib);
    }
    __syncwarp();
  };
  auto get_item = [&](unsigned item, unsigned sl) {
    if ((int)sl < S) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int ch = lane + v * 32;
        wi[v] = (4 * ch < m.pitch) ? *reinterpret_cast<const float4 *>(items_s + (size_t)sl * m.pitch + 4 * ch) : f4_zero();
      }
      ib = ibias_s[sl];
    } else {
      g.load_row(m, (size_t)m.item_off + item, wi);
      ib = __ldcg(m.bias + m.item_off + item);
    }
    cur_item = item;
    cur_slot = sl;
  };

  const int q0 = a.queue_off[w], n = a.queue_off[w + 1] - q0;
  const int B = a.batch[w];
  int pend = 0;
  unsigned my_u = 0, my_t = 0;
  long long st_wait = 0, st_flush = 0, st_nwait = 0;
  const long long st_begin = clock64();
  unsigned *const pver = pin(ver);
  const bool want_stats = pin((unsigned)(a.stats != nullptr)) != 0u;
  auto flush = [&]() {  // publish the user rows written since the last flush: one release fence
    if (pend) {
      const long long tf = want_stats ? clock64() : 0;
      __syncwarp();  // the row stores of all lanes happen-before the releases (cumulative)
      if (lane < pend) st_release_u32(pver + my_u, my_t);
      pend = 0;
      if (want_stats) st_flush += clock64() - tf;
    }
  };
  // wait until ring slot s holds fill number `par`; false: the launch is being aborted
  auto wait_full = [&](int s, unsigned par) -> bool {
    if (mbar_try(full + s, par)) return true;
    flush();  // never block with unpublished rows: somebody may be waiting for them
    const long long t0 = clock64();
    for (unsigned it = 0; !mbar_try(full + s, par); ++it) {
      if ((it & 63u) == 63u) {
        if (ld_relaxed_u32(a.abort_flag)) return false;
        if (clock64() - t0 > OWN_TIMEOUT) {
          if (lane == 0) {
            atomicCAS(a.err_flag, 0, ERR_TIMEOUT);
            st_relaxed_u32(a.abort_flag, 1u);
          }
          return false;
        }
      }
    }
    st_wait += clock64() - t0;
    ++st_nwait;
    return true;
  };
  // what an owner keeps of a ring slot: the entry, its lane's chunk(s) of the user row, the bias
  struct Inst {
    uint4 e0, e1;
    float4 wu[VEC];
    float ub;
  };
  auto read_slot = [&](int s, Inst &x) {
    const unsigned char *slot = reg + (size_t)s * slot_bytes;
    x.e0 = *reinterpret_cast<const uint4 *>(slot + row_bytes + 16);
    x.e1 = *reinterpret_cast<const uint4 *>(slot + row_bytes + 32);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int ch = lane + v * 32;
      x.wu[v] = (CH ? ch < CH : 4 * ch < m.pitch) ? *reinterpret_cast<const float4 *>(slot + 16 * ch) : f4_zero();
    }
    const float4 bw = *reinterpret_cast<const float4 *>(slot + row_bytes);
    const unsigned k = (unsigned)(m.user_off + x.e0.x) & 3u;
    x.ub = m.no_user_bias ? 0.0f : (k == 0 ? bw.x : k == 1 ? bw.y : k == 2 ? bw.z : bw.w);
  };
  const float du = a.hp.du_skip ? 1.0f : a.hp.du, di = a.hp.di_skip ? 1.0f : a.hp.di;

  if (CH) {
    // ---------------------------------------------------------------------------------------
    // The fast link: rows of exactly CH chunks (lane c < CH holds chunk c), linear loss.  This
    // loop is one dependent chain per instance of the owner's item -- what bounds the whole
    // launch -- so it is kept lean: branch-free arithmetic, constants in registers, the next entry
    // read out of its slot while this one is computed, two entries per trip (no register moves).
    //   * "s is one => no multiply" (sse.h:231-242) becomes a multiply by exactly 1.0f
    //     (x * 1.0f == x bit for bit); a skipped decay likewise;
    //   * the dot keeps the reference's order (sse.h:289-317): the products go to shared memory
    //     transposed, every lane adds the CH products of component (lane & 3) in order (lanes with
    //     the same component read the same words: broadcast), then (l0+l2)+(l1+l3);
    //   * lanes >= CH run along on whatever their registers hold; they neither store nor are read.
    // ---------------------------------------------------------------------------------------
    constexpr int LOG_D = D == 16 ? 4 : 3;
    constexpr unsigned ROW = CH * 16u, SLOT = ROW + OWN_SLOT_EXTRA, DOTROW = 36u * 4u, DOTBUF = 4u * DOTROW;
    const unsigned reg_s = pin(smem_u32(reg));
    const unsigned full_s = pin(reg_s + a.off_bars), empty_s = pin(reg_s + a.off_bars + 8u * D);
    const unsigned dot_w = pin(reg_s + a.off_dot + 4u * (unsigned)lane);          // this lane's column of the products
    const unsigned dot_r = pin(reg_s + a.off_dot + DOTROW * ((unsigned)lane & 3u));  // the row this lane adds up
    const unsigned lane16 = pin(16u * (unsigned)lane);
    char *const wbase = pin(reinterpret_cast<char *>(m.W + (size_t)m.user_off * (size_t)m.pitch) + 16 * lane);
    float *const bbase = pin(m.bias + m.user_off);
    const unsigned koff = pin((unsigned)m.user_off & 3u);
    const float lr = pin(a.hp.lr), dub = pin(a.hp.dub), dib = pin(a.hp.dib), pdu = pin(du), pdi = pin(di);
    const double base = pin((double)a.hp.base_score);
    const bool has_ub = pin((unsigned)(m.no_user_bias == 0)) != 0u;
    const bool row_lane = lane < CH, lane0 = lane == 0;
    const bool st_ub = has_ub && lane0;
    struct Link {
      uint4 e0, e1;  // {user, ticket, label, item}, {slot, uval, ival, flags}
      float4 wu, tu; // the user row (this lane's chunk) and uval * row (prepare_tmp, base.h:354-381)
      float ub, um, im;  // the user bias; uval, ival with "is one" folded in (below)
    };
    // What is known of an entry before its link starts -- everything but the item row -- is worked out when
    // the entry is read, one link ahead and off the chain.
    auto finish_link = [&](Link &x) {
      const float uval = __uint_as_float(x.e1.y), ival = __uint_as_float(x.e1.z);
      x.um = scalar_is_one(uval) ? 1.0f : uval;
      x.im = scalar_is_one(ival) ? 1.0f : ival;
      x.tu = f4_add_scaled(f4_zero(), x.wu, x.um, false);
    };
    auto read_link = [&](int s, Link &x) {
      const unsigned sa = reg_s + (unsigned)s * SLOT;
      x.e0 = lds128u(sa + ROW + 16u);
      x.e1 = lds128u(sa + ROW + 32u);
      x.wu = lds128f(sa + lane16);
      x.ub = lds32f(sa + ROW + 4u * ((x.e0.x + koff) & 3u));
      finish_link(x);
    };
    float4 &wi0 = wi[0];
    // What a link can ask for besides its arithmetic; anything set ends the inner loop below.
    // (a change of item is reported beside these, in `evx`)
    enum : unsigned { EV_PUBLISH = 1u, EV_LAST = 2u, EV_WAIT = 4u };
    // Has the entry after the current one landed?  The barrier test takes ~150 cycles to answer
    // (B300_MICROARCH.md: test_wait 149), so it is made one link ahead: link j asks for entry j+2 when it
    // starts, and reads -- where it is about to read entry j+1 out of its slot -- what link j-1 asked for.
    // "Not yet" sends the owner to the blocking wait (which looks again), so a stale "no" costs a detour,
    // never a wrong read.  Asking and reading the answer are two statements around a predicate register of
    // this function, so that the links' work stands between them (as one statement the answer is consumed at
    // once: a 150-cycle stall); two registers, one per link of a loop trip, since two questions are open.
    unsigned evx = 0u;
    const unsigned lane0i = lane0 ? 1u : 0u;
    asm volatile(".reg .pred own_landed0, own_landed1;");
    auto ask_entry = [&](auto which, int e) {  // (past the end of the queue: some barrier of the ring, answer unused)
      const unsigned bar = full_s + 8u * (unsigned)(e & (D - 1)), par = (unsigned)(e >> LOG_D) & 1u;
      if (decltype(which)::value == 0)
        asm volatile("mbarrier.test_wait.parity.shared::cta.b64 own_landed0, [%0], %1;" ::"r"(bar), "r"(par) : "memory");
      else
        asm volatile("mbarrier.test_wait.parity.shared::cta.b64 own_landed1, [%0], %1;" ::"r"(bar), "r"(par) : "memory");
    };
    auto answer = [&](auto which) -> unsigned {
      unsigned ok;
      if (decltype(which)::value == 0)
        asm volatile("selp.u32 %0, 1, 0, own_landed0;" : "=r"(ok)::"memory");
      else
        asm volatile("selp.u32 %0, 1, 0, own_landed1;" : "=r"(ok)::"memory");
      return ok;
    };
    typedef std::integral_constant<int, 0> P0;
    typedef std::integral_constant<int, 1> P1;
    // One link: the arithmetic of `cur` (whose item row is in registers), its user row stored, the next
    // entry read out of its ring slot into `nxt` if it has landed.  Straight-line code; the warp issues in
    // order, so the statements stand in the order the chain wants them:
    //   1. the barrier test for the entry after the next is asked for; the products of the dot go to shared
    //      memory -- only the item row is new to this link;
    //   2. the off-chain sums;
    //   3. the transposed products come back, the next entry is read behind them;
    //   4. the adds of the dot, the error, the two new rows.
    auto body = [&](auto mine, auto other, Link &cur, Link &nxt, int j) -> unsigned {
      ask_entry(mine, j + 2);
      const float4 ti = f4_add_scaled(f4_zero(), wi0, cur.im, false);  // prepare_tmp, base.h:354-381
      const unsigned dw = dot_w + ((unsigned)j & 1u) * DOTBUF, dr = dot_r + ((unsigned)j & 1u) * DOTBUF;
      sts32f(dw, __fmul_rn(cur.tu.x, ti.x));
      sts32f(dw + DOTROW, __fmul_rn(cur.tu.y, ti.y));
      sts32f(dw + 2u * DOTROW, __fmul_rn(cur.tu.z, ti.z));
      sts32f(dw + 3u * DOTROW, __fmul_rn(cur.tu.w, ti.w));
      // (flags as integers, not predicates: the few predicate registers are left to the barrier test, whose
      // answer must stay in one from here to the end of the link)
      const int j1 = j + 1, s1 = j1 & (D - 1);
      const unsigned more = (unsigned)(j1 - n) >> 31;  // 1: there is a next entry
      const unsigned nready = more & answer(other);    // 1: ... and it has landed (asked by the link before)
      const float uval = __uint_as_float(cur.e1.y), ival = __uint_as_float(cur.e1.z), label = __uint_as_float(cur.e0.z);
      double bsum = 0.0;  // calc_bias, base.h:313-353 (needs no dot: off the chain)
      if (has_ub) bsum = __dadd_rn(bsum, (double)__fmul_rn(uval, cur.ub));
      bsum = __dadd_rn(bsum, (double)__fmul_rn(ival, ib));
      const double s0 = __dadd_rn(base, bsum);
      __syncwarp();
      constexpr int NQ = CH ? CH / 4 : 1;
      float4 q[NQ];
#pragma unroll
      for (int i = 0; i < NQ; ++i) q[i] = lds128f(dr + 16u * i);  // (all in flight before the first add)
      // the answer of the test enters the address: these reads cannot be issued before it (not landed: the
      // owner's first slot is read and the result discarded)
      const unsigned sa = reg_s + nready * ((unsigned)s1 * SLOT);
      nxt.e0 = lds128u(sa + ROW + 16u);
      nxt.e1 = lds128u(sa + ROW + 32u);
      nxt.wu = lds128f(sa + lane16);
      float acc = 0.0f;
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        acc = __fadd_rn(acc, q[i].x);
        acc = __fadd_rn(acc, q[i].y);
        acc = __fadd_rn(acc, q[i].z);
        acc = __fadd_rn(acc, q[i].w);
      }
      const float l0 = __shfl_sync(0xffffffffu, acc, 0), l1 = __shfl_sync(0xffffffffu, acc, 1),
                  l2 = __shfl_sync(0xffffffffu, acc, 2), l3 = __shfl_sync(0xffffffffu, acc, 3);
      nxt.ub = lds32f(sa + ROW + 4u * ((nxt.e0.x + koff) & 3u));
      finish_link(nxt);
      __syncwarp();
      // every lane has read the next entry out of its ring slot: the slot may be refilled
      if (nready + lane0i == 2u) mbar_arrive_s(empty_s + 8u * (unsigned)s1);
      const float d = __fadd_rn(__fadd_rn(l0, l2), __fadd_rn(l1, l3));
      const float p = (float)__dadd_rn(s0, (double)d);  // pred, base.h:445-454 (linear: map_active is the identity)
      const float err = __fsub_rn(label, p);           // cal_grad, model.h:132-156
      const float lrerr = __fmul_rn(lr, err);
      const float su = __fmul_rn(lrerr, uval), si = __fmul_rn(lrerr, ival);  // base.h:391,412
      const float sum_ = scalar_is_one(su) ? 1.0f : su, sim = scalar_is_one(si) ? 1.0f : si;
      wi0 = f4_scale(f4_add_scaled(wi0, cur.tu, sim, false), pdi);  // update_no_decay + regularize(after)
      ib = __fmul_rn(__fadd_rn(ib, si), dib);
      const float4 nwu = f4_scale(f4_add_scaled(cur.wu, ti, sum_, false), pdu);
      const unsigned user = cur.e0.x;
      if (row_lane) stcg4(reinterpret_cast<float *>(wbase + (size_t)user * ROW), nwu);
      if (st_ub) __stcg(bbase + user, __fmul_rn(__fadd_rn(cur.ub, su), dub));
      if (lane == pend) {
        my_u = user;
        my_t = cur.e0.y + 1u;
      }
      ++pend;
      evx = (0u - nready) & (nxt.e0.w ^ cur_item);  // non-zero: the next entry is of another item
      const unsigned ev = (cur.e1.w & 1u) | ((unsigned)(B - 1 - pend) >> 31)   // EV_PUBLISH: asked for, or the batch is full
                          | ((more ^ 1u) << 1)                                  // EV_LAST
                          | ((more & (nready ^ 1u)) << 2);                      // EV_WAIT
      return ev;
    };
    Link la, lb;
    if (n > 0 && wait_full(0, 0u)) {
      read_link(0, la);
      __syncwarp();
      if (lane0) mbar_arrive_s(empty_s);
      get_item(la.e0.w, la.e1.x);
      ask_entry(P1(), 1);
      // The inner loop is nothing but links (two per trip: `la` and `lb` swap roles, no register moves);
      // whatever else has to happen -- the publish fence, a next entry that has not landed, a change of
      // item, the end of the queue -- leaves it, is dealt with here, and re-enters with the current entry in `la`.
      for (int j = 0;;) {
        unsigned ev;
        for (;;) {
          ev = body(P0(), P1(), la, lb, j);
          if (ev | evx) {
            la = lb;
            break;
          }
          ++j;
          ev = body(P1(), P0(), lb, la, j);
          if (ev | evx) break;
          ++j;
        }
        if (ev & EV_PUBLISH) flush();
        if (ev & EV_LAST) break;
        ++j;
        if (ev & EV_WAIT) {
          st_nwait += 1LL << 32;  // (own_stats: upper half = how often the next entry had not landed one link ahead)
          const int s = j & (D - 1);
          if (!wait_full(s, (unsigned)(j >> LOG_D) & 1u)) break;
          read_link(s, la);
          __syncwarp();
          if (lane0) mbar_arrive_s(empty_s + 8u * (unsigned)s);
        }
        ask_entry(P1(), j + 1);  // (the first link of a trip reads this register)
        if (la.e0.w != cur_item) {
          put_item();
          get_item(la.e0.w, la.e1.x);
        }
      }
    }
  } else {
    Inst cur, nxt;
    bool alive = n > 0;
    if (alive) {
      alive = wait_full(0, 0u);
      read_slot(0, cur);
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + 0);  // the slot may be refilled
    }
    for (int j = 0; j < n && alive; ++j) {
      // the next entry is read out of its slot (if it has landed) while this one is computed
      const int s1 = (j + 1) & (D - 1);
      const unsigned par1 = (unsigned)((j + 1) / D) & 1u;
      const bool more = j + 1 < n;
      const bool nready = more && mbar_test(full + s1, par1);
      if (nready) read_slot(s1, nxt);  // (behind the branch: never before the test has answered)
      if (cur.e0.w != cur_item) {
        put_item();
        get_item(cur.e0.w, cur.e1.x);
      }
      own_step<VEC>(g, m, a.hp, cur.wu, cur.ub, wi, ib, __uint_as_float(cur.e1.y), __uint_as_float(cur.e1.z),
                    __uint_as_float(cur.e0.z));
      const size_t urow = (size_t)m.user_off + cur.e0.x;
      g.store_row(m, urow, cur.wu);
      if (!m.no_user_bias && lane == 0) __stcg(m.bias + urow, cur.ub);
      if (nready) {
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s1);
      }
      if (lane == pend) {
        my_u = cur.e0.x;
        my_t = cur.e0.y + 1u;
      }
      ++pend;
      if ((cur.e1.w & 1u) || pend >= B) flush();
      if (more && !nready) {
        alive = wait_full(s1, par1);
        read_slot(s1, nxt);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s1);
      }
      cur = nxt;
    }
  }
  flush();
  put_item();
  if (a.stats && lane == 0) {
    a.stats[4 * w + 0] = clock64() - st_begin;
    a.stats[4 * w + 1] = st_wait;
    a.stats[4 * w + 2] = st_flush;
    a.stats[4 * w + 3] = st_nwait;
  }
  // resident item rows go home
  for (int sl = 0; sl < nres; ++sl) {
    const size_t row = (size_t)m.item_off + a.items[it0 + sl];
    float *dst = m.W + row * (size_t)m.pitch;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int ch = lane + v * 32;
      if (4 * ch < m.pitch) stcg4(dst + 4 * ch, *reinterpret_cast<const float4 *>(items_s + (size_t)sl * m.pitch + 4 * ch));
    }
    if (lane == 0) __stcg(m.bias + row, ibias_s[sl]);
  }
}

// ---------------------------------------------------------------------------
// k_own2 -- the same scheme with the link split over TWO warps (rows of CH chunks, linear loss).
//
// What makes the launch long is the chain of one item's updates, and only half of a link is on
// that chain: the dot, the error and the new ITEM row.  The new USER row, its stores and the
// publish fences are not -- nothing of the next link needs them.  Here the owner warp does the
// first half and hands {ti = the item row the user row needs, lr*err} to its PARTNER warp through a
// small shared-memory ring; the partner reads the user row and the entry out of the same ring
// slot the owner used, forms and stores the new user row and bias, and publishes versions: in
// batches while it is behind the owner, at once whenever it has caught up (it would block
// anyway) -- so a user row is published as soon as there is nothing better to do, and no "returns
// soon" flag is needed.  The ring slot is released by the partner, the last one to read it.
// 8 owners + 8 partners + 2 loader warps per CTA.
// ---------------------------------------------------------------------------
constexpr int OWN2_C = 8;   // owners per CTA
constexpr int OWN2_H = 4;   // hand-off slots per owner
constexpr int own2_threads(int D) { return (2 * OWN2_C + OWN2_C * D / 32) * 32; }

template <int CH, int D>
__global__ void __launch_bounds__(own2_threads(D), 1) k_own2(const OwnArgs a) {
  extern __shared__ __align__(128) unsigned char own_smem[];
  constexpr int C = OWN2_C, H = OWN2_H;
  constexpr int LOG_D = D == 16 ? 4 : 3;
  constexpr unsigned ROW = CH * 16u, SLOT = ROW + OWN_SLOT_EXTRA, DOTROW = 36u * 4u, DOTBUF = 4u * DOTROW;
  constexpr unsigned HSLOT = 32u * 16u + 16u;  // every lane's float4 (lanes >= CH: never read) + {lr*err}
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const DevModel &m = a.m;
  const int S = a.S;
  // barriers of an owner: full[D], empty[D] (ring; one arrival each), hfull[H] (32 arrivals: every lane of
  // the owner after its own store), hempty[H] (one arrival: the partner)
  if (warp < C) {
    uint64_t *bars = reinterpret_cast<uint64_t *>(own_smem + (size_t)warp * a.region_bytes + a.off_bars);
    for (int i = lane; i < 2 * D + 2 * H; i += 32) mbar_init(bars + i, (i >= 2 * D && i < 2 * D + H) ? 32 : 1);
  }
  mbar_fence_init();
  __syncthreads();
  if (warp >= 2 * C) {
    own_loader<D>(a, own_smem, (int)threadIdx.x - 2 * C * 32, C);
    return;
  }
  // owner c and its partner sit on different schedulers (warp ids c and C + (c + C - 2) % C)
  const bool is_owner = warp < C;
  const int c = is_owner ? warp : (warp - C + 2) % C;
  const int w = ((a.flags & OWN_F_REVERSE) ? C - 1 - c : c) * (int)gridDim.x + (int)blockIdx.x;
  unsigned char *reg = own_smem + (size_t)c * a.region_bytes;
  const unsigned reg_s = pin(smem_u32(reg));
  const unsigned full_s = pin(reg_s + a.off_bars), empty_s = pin(reg_s + a.off_bars + 8u * D);
  const unsigned hfull_s = pin(reg_s + a.off_bars + 16u * D), hempty_s = pin(reg_s + a.off_bars + 16u * D + 8u * H);
  const unsigned hand_s = pin(reg_s + a.off_hand);
  const unsigned lane16 = pin(16u * (unsigned)lane);
  const unsigned koff = pin((unsigned)m.user_off & 3u);
  const bool lane0 = lane == 0;
  const int q0 = a.queue_off[w], n = a.queue_off[w + 1] - q0;
  (void)q0;
  // wait until barrier `bar` has completed the phase of parity `par`; false: the launch is being aborted
  auto wait_bar = [&](unsigned bar, unsigned par) -> bool {
    const long long t0 = clock64();
    for (unsigned it = 0;; ++it) {
      unsigned ok;
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
          "selp.u32 %0, 1, 0, p;\n"
          "}\n"
          : "=r"(ok)
          : "r"(bar), "r"(par), "r"(100000u)
          : "memory");
      if (ok) return true;
      if ((it & 63u) == 63u) {
        if (ld_relaxed_u32(a.abort_flag)) return false;
        if (clock64() - t0 > OWN_TIMEOUT) {
          if (lane0) {
            atomicCAS(a.err_flag, 0, ERR_TIMEOUT);
            st_relaxed_u32(a.abort_flag, 1u);
          }
          return false;
        }
      }
    }
  };

  if (!is_owner) {
    // ================================ partner warp ==========================================
    char *const wbase = pin(reinterpret_cast<char *>(m.W + (size_t)m.user_off * (size_t)m.pitch) + 16 * lane);
    float *const bbase = pin(m.bias + m.user_off);
    unsigned *const pver = pin(m.ver_ui + m.user_off);
    const float du = pin(a.hp.du_skip ? 1.0f : a.hp.du), dub = pin(a.hp.dub);
    const bool has_ub = pin((unsigned)(m.no_user_bias == 0)) != 0u;
    const bool row_lane = lane < CH, st_ub = has_ub && lane0;
    const int B = a.batch[w];
    int pend = 0;
    unsigned my_u = 0, my_t = 0;
    auto flush = [&]() {  // publish the user rows written since the last flush: one release fence
      if (pend) {
        __syncwarp();  // the row stores of all lanes happen-before the releases (cumulative)
        if (lane < pend) st_release_u32(pver + my_u, my_t);
        pend = 0;
      }
    };
    for (int j = 0; j < n; ++j) {
      const int hs = j & (H - 1), s = j & (D - 1);
      const unsigned hb = hfull_s + 8u * (unsigned)hs, hpar = (unsigned)(j / H) & 1u;
      if (!mbar_test_s(hb, hpar)) {
        flush();  // caught up with the owner: publish now, then wait for the next hand-off
        if (!wait_bar(hb, hpar)) break;
      }
      // (the owner has consumed ring slot s, so its bulk copy has landed; test it all the same: this
      // warp's own acquire on the barrier the TMA unit completed)
      if (!mbar_test_s(full_s + 8u * (unsigned)s, (unsigned)(j >> LOG_D) & 1u) &&
          !wait_bar(full_s + 8u * (unsigned)s, (unsigned)(j >> LOG_D) & 1u))
        break;
      const unsigned ha = hand_s + (unsigned)hs * HSLOT, sa = reg_s + (unsigned)s * SLOT;
      const float4 ti = lds128f(ha + lane16);
      const float lrerr = lds32f(ha + 32u * 16u);
      const uint4 e0 = lds128u(sa + ROW + 16u), e1 = lds128u(sa + ROW + 32u);
      const float4 wu = lds128f(sa + lane16);
      const float ub = lds32f(sa + ROW + 4u * ((e0.x + koff) & 3u));
      __syncwarp();
      if (lane0) {  // both slots may be refilled
        mbar_arrive_s(hempty_s + 8u * (unsigned)hs);
        mbar_arrive_s(empty_s + 8u * (unsigned)s);
      }
      const float uval = __uint_as_float(e1.y);
      const float su = __fmul_rn(lrerr, uval);  // base.h:391
      const float sum_ = scalar_is_one(su) ? 1.0f : su;
      const float4 nwu = f4_scale(f4_add_scaled(wu, ti, sum_, false), du);  // update_no_decay + regularize(after)
      const unsigned user = e0.x;
      if (row_lane) stcg4(reinterpret_cast<float *>(wbase + (size_t)user * ROW), nwu);
      if (st_ub) __stcg(bbase + user, __fmul_rn(__fadd_rn(ub, su), dub));
      if (lane == pend) {
        my_u = user;
        my_t = e0.y + 1u;
      }
      if (++pend >= B) flush();
    }
    flush();
    return;
  }

  // ================================ owner warp ================================================
  float *items_s = reinterpret_cast<float *>(reg + a.off_items);
  float *ibias_s = reinterpret_cast<float *>(reg + a.off_ibias);
  const int it0 = a.item_off[w], nit = a.item_off[w + 1] - it0, nres = min(nit, S);
  const bool row_lane = lane < CH;
  for (int sl = 0; sl < nres; ++sl) {  // the owner's most popular item rows live in shared memory for the launch
    const size_t row = (size_t)m.item_off + a.items[it0 + sl];
    if (row_lane) *reinterpret_cast<float4 *>(items_s + (size_t)sl * m.pitch + 4 * lane) = ldcg4(m.W + row * (size_t)m.pitch + 4 * lane);
    if (lane0) ibias_s[sl] = __ldcg(m.bias + row);
  }
  __syncwarp();
  float4 wi = f4_zero();
  float ib = 0.0f;
  unsigned cur_item = 0xffffffffu, cur_slot = 0;
  auto put_item = [&]() {  // the current item row leaves the registers
    if (cur_item == 0xffffffffu) return;
    if ((int)cur_slot < S) {
      if (row_lane) *reinterpret_cast<float4 *>(items_s + (size_t)cur_slot * m.pitch + 4 * lane) = wi;
      if (lane0) ibias_s[cur_slot] = ib;
    } else {
      if (row_lane) stcg4(m.W + ((size_t)m.item_off + cur_item) * (size_t)m.pitch + 4 * lane, wi);
      if (lane0) __stcg(m.bias + m.item_off + cur_item, ib);
    }
    __syncwarp();
  };
  auto get_item = [&](unsigned item, unsigned sl) {
    if ((int)sl < S) {
      wi = row_lane ? *reinterpret_cast<const float4 *>(items_s + (size_t)sl * m.pitch + 4 * lane) : f4_zero();
      ib = ibias_s[sl];
    } else {
      wi = row_lane ? ldcg4(m.W + ((size_t)m.item_off + item) * (size_t)m.pitch + 4 * lane) : f4_zero();
      ib = __ldcg(m.bias + m.item_off + item);
    }
    cur_item = item;
    cur_slot = sl;
  };
  const unsigned dot_w = pin(reg_s + a.off_dot + 4u * (unsigned)lane);             // this lane's column of the products
  const unsigned dot_r = pin(reg_s + a.off_dot + DOTROW * ((unsigned)lane & 3u));  // the row this lane adds up
  const float lr = pin(a.hp.lr), dib = pin(a.hp.dib), pdi = pin(a.hp.di_skip ? 1.0f : a.hp.di);
  const double base = pin((double)a.hp.base_score);
  const bool has_ub = pin((unsigned)(m.no_user_bias == 0)) != 0u;
  long long st_wait = 0, st_nwait = 0;
  const long long st_begin = clock64();
  struct Link {
    uint4 e0, e1;  // {user, ticket, label, item}, {slot, uval, ival, flags}
    float4 wu;
    float ub;
  };
  auto read_link = [&](int s, Link &x, unsigned landed = 1u) {  // (`landed` enters the address: see k_own)
    const unsigned sa = reg_s + landed * ((unsigned)s * SLOT);
    x.e0 = lds128u(sa + ROW + 16u);
    x.e1 = lds128u(sa + ROW + 32u);
    x.wu = lds128f(sa + lane16);
    x.ub = lds32f(sa + ROW + 4u * ((x.e0.x + koff) & 3u));
  };
  auto link = [&](Link &cur, Link &nxt, int j) -> bool {
    const int j1 = j + 1, s1 = j1 & (D - 1), hs = j & (H - 1);
    const unsigned par1 = (unsigned)(j1 >> LOG_D) & 1u;
    const bool more = j1 < n;
    const bool nready = more && mbar_test_s(full_s + 8u * (unsigned)s1, par1);
    // hand-off slot hs is free once the partner has read link j - H out of it (a fresh barrier passes parity 1)
    const bool hfree = mbar_test_s(hempty_s + 8u * (unsigned)hs, ((unsigned)(j / H) & 1u) ^ 1u);
    read_link(s1, nxt, nready ? 1u : 0u);  // (not landed yet: read again below)
    const float uval = __uint_as_float(cur.e1.y), ival = __uint_as_float(cur.e1.z), label = __uint_as_float(cur.e0.z);
    const float um = scalar_is_one(uval) ? 1.0f : uval, im = scalar_is_one(ival) ? 1.0f : ival;
    const float4 tu = f4_add_scaled(f4_zero(), cur.wu, um, false);  // prepare_tmp, base.h:354-381
    const float4 ti = f4_add_scaled(f4_zero(), wi, im, false);
    const unsigned dw = dot_w + ((unsigned)j & 1u) * DOTBUF, dr = dot_r + ((unsigned)j & 1u) * DOTBUF;
    sts32f(dw, __fmul_rn(tu.x, ti.x));
    sts32f(dw + DOTROW, __fmul_rn(tu.y, ti.y));
    sts32f(dw + 2u * DOTROW, __fmul_rn(tu.z, ti.z));
    sts32f(dw + 3u * DOTROW, __fmul_rn(tu.w, ti.w));
    double bsum = 0.0;  // calc_bias, base.h:313-353 (needs no dot: off the chain)
    if (has_ub) bsum = __dadd_rn(bsum, (double)__fmul_rn(uval, cur.ub));
    bsum = __dadd_rn(bsum, (double)__fmul_rn(ival, ib));
    const double s0 = __dadd_rn(base, bsum);
    __syncwarp();
    float acc = 0.0f;
    float4 q[CH / 4];
#pragma unroll
    for (int i = 0; i < CH / 4; ++i) q[i] = lds128f(dr + 16u * i);  // (all in flight before the first add)
#pragma unroll
    for (int i = 0; i < CH / 4; ++i) {
      acc = __fadd_rn(acc, q[i].x);
      acc = __fadd_rn(acc, q[i].y);
      acc = __fadd_rn(acc, q[i].z);
      acc = __fadd_rn(acc, q[i].w);
    }
    const float l0 = __shfl_sync(0xffffffffu, acc, 0), l1 = __shfl_sync(0xffffffffu, acc, 1),
                l2 = __shfl_sync(0xffffffffu, acc, 2), l3 = __shfl_sync(0xffffffffu, acc, 3);
    const float d = __fadd_rn(__fadd_rn(l0, l2), __fadd_rn(l1, l3));
    const float p = (float)__dadd_rn(s0, (double)d);  // pred, base.h:445-454 (linear: map_active is the identity)
    const float err = __fsub_rn(label, p);           // cal_grad, model.h:132-156
    const float lrerr = __fmul_rn(lr, err);
    const float si = __fmul_rn(lrerr, ival);  // base.h:412
    const float sim = scalar_is_one(si) ? 1.0f : si;
    wi = f4_scale(f4_add_scaled(wi, tu, sim, false), pdi);  // update_no_decay + regularize(after), item side
    ib = __fmul_rn(__fadd_rn(ib, si), dib);
    // The user-side half of the link goes to the partner: ti and lr*err.  This comes BEFORE any wait
    // for the next entry: that entry may be the same user again, whose version only the partner
    // can publish -- once it has been handed this link.
    if (!hfree) {  // (the partner is H links behind: rare)
      const long long t0 = clock64();
      if (!wait_bar(hempty_s + 8u * (unsigned)hs, ((unsigned)(j / H) & 1u) ^ 1u)) return false;
      st_wait += clock64() - t0;
    }
    const unsigned ha = hand_s + (unsigned)hs * HSLOT;
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ha + lane16), "f"(ti.x), "f"(ti.y), "f"(ti.z), "f"(ti.w) : "memory");
    if (lane0) sts32f(ha + 32u * 16u, lrerr);
    mbar_arrive_s(hfull_s + 8u * (unsigned)hs);  // (all 32 lanes arrive, each after its own store)
    // everything else that is rare behind one branch: next entry not landed, change of item
    if (more && (!nready || nxt.e0.w != cur_item)) {
      if (!nready) {
        const long long t0 = clock64();
        if (!wait_bar(full_s + 8u * (unsigned)s1, par1)) return false;
        st_wait += clock64() - t0;
        ++st_nwait;
        read_link(s1, nxt);
      }
      if (nxt.e0.w != cur_item) {
        put_item();
        get_item(nxt.e0.w, nxt.e1.x);
      }
    }
    return true;
  };
  Link la, lb;
  if (n > 0 && wait_bar(full_s, 0u)) {
    read_link(0, la);
    get_item(la.e0.w, la.e1.x);
    for (int j = 0; j < n; j += 2) {
      if (!link(la, lb, j)) break;
      if (j + 1 < n && !link(lb, la, j + 1)) break;
    }
  }
  put_item();
  if (a.stats && lane0) {
    a.stats[4 * w + 0] = clock64() - st_begin;
    a.stats[4 * w + 1] = st_wait;
    a.stats[4 * w + 2] = 0;
    a.stats[4 * w + 3] = st_nwait;
  }
  for (int sl = 0; sl < nres; ++sl) {  // resident item rows go home
    const size_t row = (size_t)m.item_off + a.items[it0 + sl];
    if (row_lane) stcg4(m.W + row * (size_t)m.pitch + 4 * lane, *reinterpret_cast<const float4 *>(items_s + (size_t)sl * m.pitch + 4 * lane));
    if (lane0) __stcg(m.bias + row, ibias_s[sl]);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int own_reserve(svdgpu *h, DevBuf &b, size_t bytes) {
  bytes += 64;
  if (bytes <= b.cap) return 0;
  if (b.p) CU(h, cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  CU(h, cudaMalloc(&b.p, bytes + bytes / 8));
  b.cap = bytes + bytes / 8;
  return 0;
}

static int bits_for(unsigned n) {  // key bits needed for values < n
  int b = 1;
  while (b < 32 && (1ull << b) < (unsigned long long)n) ++b;
  return b;
}

// the split link (k_own2) covers linear loss and rows of exactly 4 / 8 / 16 / 32 chunks
static bool own_fast_shape(const svdgpu *h) {
  const int chunks = h->dm.pitch / 4;
  return h->own_fast && h->dm.active_type == 0 && h->dm.k == h->dm.pitch && (chunks == 4 || chunks == 8 || chunks == 16 || chunks == 32);
}
int own_owners_per_cta(const svdgpu *h) { return (h->own_partner && own_fast_shape(h)) ? OWN2_C : OWN_C; }

bool own_supported(const svdgpu *h) {
  // plain L2 decay only (the other regularisers keep k_exact), rows of at most 256 floats
  return h->dhp.plain && h->dm.pitch <= 256 && h->shape.format_type == 0 && h->dm.num_user > 0 && h->dm.num_item > 0;
}

void own_plan_free(OwnPlan &p) {
  DevBuf *db[] = {&p.entries, &p.queue_off, &p.item_off, &p.items, &p.batch};
  for (DevBuf *d : db) {
    if (d->p) cudaFree(d->p);
    d->p = nullptr;
    d->cap = 0;
  }
  if (p.stage.p) cudaFreeHost(p.stage.p);
  p.stage.p = nullptr;
  p.stage.cap = 0;
  p.valid = false;
}
void own_scratch_free(OwnScratch &s) {
  DevBuf *db[] = {&s.cnt_item, &s.cnt_user, &s.start_user, &s.flag, &s.keyA, &s.keyB, &s.valA, &s.valB,
                  &s.key_item, &s.tick, &s.tmp, &s.item_owner, &s.item_slot, &s.stats};
  for (DevBuf *d : db) {
    if (d->p) cudaFree(d->p);
    d->p = nullptr;
    d->cap = 0;
  }
  if (s.h_cnt) cudaFreeHost(s.h_cnt);
  s.h_cnt = nullptr;
  s.h_cnt_cap = 0;
  if (s.ev) cudaEventDestroy(s.ev);
  s.ev = nullptr;
  delete s.deal;
  s.deal = nullptr;
}

// Build the plan of rows [r0, r0+n) of a device CSR on `st`.  Returns non-zero on failure (message
// set).  p.valid tells whether the rows can take k_own; *bad receives the OWN_* bits otherwise.
int own_plan_build(svdgpu *h, const DevCsr &csr, int r0, int n, OwnPlan &p, cudaStream_t st, int *bad, int ctas) {
  p.valid = false;
  if (bad) *bad = 0;
  if (n <= 0) return 0;
  OwnScratch &s = h->own;
  const DevModel &m = h->dm;
  if (ctas <= 0 || ctas > h->num_sm) ctas = h->num_sm;
  const int per_cta = own_owners_per_cta(h);
  const int W = ctas * per_cta;
  const size_t nn = (size_t)n;
  if (own_reserve(h, s.cnt_item, (size_t)m.num_item * 4) || own_reserve(h, s.cnt_user, (size_t)m.num_user * 4) ||
      own_reserve(h, s.start_user, (size_t)m.num_user * 4) || own_reserve(h, s.flag, 4) ||
      own_reserve(h, s.keyA, nn * 4) || own_reserve(h, s.keyB, nn * 4) || own_reserve(h, s.valA, nn * 4) ||
      own_reserve(h, s.valB, nn * 4) || own_reserve(h, s.key_item, nn * 4) || own_reserve(h, s.tick, nn * 4) ||
      own_reserve(h, s.item_owner, (size_t)m.num_item * 4) || own_reserve(h, s.item_slot, (size_t)m.num_item * 4))
    return 1;
  if (!s.ev) CU(h, cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
  if (s.h_cnt_cap < (size_t)m.num_item * 4 + 4) {
    if (s.h_cnt) cudaFreeHost(s.h_cnt);
    s.h_cnt = nullptr;
    s.h_cnt_cap = 0;
    CU(h, cudaMallocHost(&s.h_cnt, (size_t)m.num_item * 4 + 64));
    s.h_cnt_cap = (size_t)m.num_item * 4 + 4;
  }
  const int blocks = (int)std::min<long long>((long long)h->num_sm * 16, ((long long)n + 255) / 256);
  CU(h, cudaMemsetAsync(s.cnt_item.p, 0, (size_t)m.num_item * 4, st));
  CU(h, cudaMemsetAsync(s.cnt_user.p, 0, (size_t)m.num_user * 4, st));
  CU(h, cudaMemsetAsync(s.flag.p, 0, 4, st));
  unsigned *keyA = (unsigned *)s.keyA.p, *keyB = (unsigned *)s.keyB.p, *valA = (unsigned *)s.valA.p,
           *valB = (unsigned *)s.valB.p;
  k_own_scan<<<blocks, 256, 0, st>>>(csr, r0, n, m.num_user, m.num_item, (unsigned *)s.cnt_item.p,
                                     (unsigned *)s.cnt_user.p, keyA, (unsigned *)s.key_item.p, valA, (int *)s.flag.p);
  CU(h, cudaGetLastError());
  h->n_launch++;
  // item counts (+ the flag word behind them) to the host: the LPT runs there while the device sorts by user
  CU(h, cudaMemcpyAsync(s.h_cnt, s.cnt_item.p, (size_t)m.num_item * 4, cudaMemcpyDeviceToHost, st));
  CU(h, cudaMemcpyAsync((char *)s.h_cnt + (size_t)m.num_item * 4, s.flag.p, 4, cudaMemcpyDeviceToHost, st));
  CU(h, cudaEventRecord(s.ev, st));
  h->n_d2h += (long long)m.num_item * 4 + 4;

  // tickets: stable sort by user, rank inside the user's run
  size_t tmp_scan = 0, tmp_sort = 0;
  CU(h, cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, (unsigned *)s.cnt_user.p, (unsigned *)s.start_user.p, m.num_user, st));
  {
    cub::DoubleBuffer<unsigned> dk(keyA, keyB), dv(valA, valB);
    CU(h, cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, dk, dv, n, 0, 32, st));
  }
  if (own_reserve(h, s.tmp, std::max(tmp_scan, tmp_sort))) return 1;
  size_t tmp_bytes = s.tmp.cap;
  CU(h, cub::DeviceScan::ExclusiveSum(s.tmp.p, tmp_bytes, (unsigned *)s.cnt_user.p, (unsigned *)s.start_user.p, m.num_user, st));
  {
    cub::DoubleBuffer<unsigned> dk(keyA, keyB), dv(valA, valB);
    tmp_bytes = s.tmp.cap;
    CU(h, cub::DeviceRadixSort::SortPairs(s.tmp.p, tmp_bytes, dk, dv, n, 0, bits_for((unsigned)m.num_user), st));
    k_own_ticket<<<blocks, 256, 0, st>>>(dk.Current(), dv.Current(), (const unsigned *)s.start_user.p, n,
                                         (unsigned)h->own_urgent_gap, (unsigned *)s.tick.p);
    CU(h, cudaGetLastError());
    h->n_launch += 3;
  }

  // host: deal the items out
  const auto t_wait0 = std::chrono::steady_clock::now();
  CU(h, cudaEventSynchronize(s.ev));
  const auto t_lpt0 = std::chrono::steady_clock::now();
  h->own_cntwait_us += std::chrono::duration_cast<std::chrono::microseconds>(t_lpt0 - t_wait0).count();
  const unsigned *hc = (const unsigned *)s.h_cnt;
  const int fl = (int)hc[m.num_item];
  if (fl) {
    if (bad) *bad = fl;
    return 0;  // not for k_own (the caller reports bound errors / falls back to k_exact)
  }
  // The deal of items to owners is carried over from the plan before while it stays a good one for the new
  // counts (svdown::redeal; the options that shape a deal are part of its key): a host-pointer call plans every
  // chunk, and dealing anew costs the host 2.3 ms of a 3.6 ms chunk.
  if (!s.deal) s.deal = new svdown::HostPlan();
  svdown::HostPlan &hp = *s.deal;
  const long long deal_key = (long long)h->own_batch | ((long long)h->own_isolate << 8) | ((long long)h->own_isolate_full << 24) |
                             ((long long)per_cta << 40) | ((long long)m.num_item << 44);
  const bool carried = h->own_redeal > 0 && s.deal_ctas == ctas && s.deal_key == deal_key && hp.age < 64 &&
                       svdown::redeal(hc, m.num_item, h->own_batch, h->own_redeal, hp);
  if (carried) ++h->n_redeal;
  else {
  // The chain of the hottest item is what the whole launch waits for, and a link of it is slower in company:
  // measured per link (20 M rows, hottest item 51 080), 0.27 us with the SM to itself, 0.32 us with the other
  // eleven owner warps at work but nobody on its issue port (warps w, w+4, w+8 share one: warp id mod 4),
  // 0.36-0.37 us with two busy neighbours there.  So the owners of hot items (more than own_isolate % of the
  // mean owner load) get the port to themselves, and those within reach of the longest chain (own_isolate_full %
  // of it, at most an eighth of the SMs) a whole SM; the warps so closed get no items.  (k_own's mapping: owner
  // o is warp OWN_C-1 - o/ctas, or o/ctas without own_reverse, of block o % ctas; LPT gives the r-th most
  // popular item to owner r.)
  std::vector<char> closed;
  if (h->own_isolate > 0 && per_cta == OWN_C) {
    const int hot = std::min(svdown::hot_owners(hc, m.num_item, W, h->own_isolate), ctas);
    const int full = h->own_isolate_full > 0
                         ? std::min(std::min(svdown::top_owners(hc, m.num_item, h->own_isolate_full), hot), std::max(ctas / 8, 1))
                         : 0;
    if (hot > 0) {
      closed.assign((size_t)W, 0);
      for (int o = 0; o < hot; ++o)
        for (int c = 1; c < OWN_C; ++c)
          if (c % 4 == 0 || o < full) closed[(size_t)c * ctas + o] = 1;  // (c % 4 == 0: same port under either mapping)
    }
  }
  svdown::assign(hc, m.num_item, W, h->own_batch, hp, closed.empty() ? nullptr : &closed, h->own_redeal > 0);
  s.deal_ctas = ctas;
  s.deal_key = deal_key;
  ++h->n_deal;
  }
  h->own_lpt_us += std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t_lpt0).count();
  if (own_reserve(h, p.queue_off, (size_t)(W + 1) * 4) || own_reserve(h, p.item_off, (size_t)(W + 1) * 4) ||
      own_reserve(h, p.items, std::max<size_t>(hp.items.size(), 1) * 4) || own_reserve(h, p.batch, (size_t)W * 4) ||
      own_reserve(h, p.entries, nn * sizeof(OwnEntry)))
    return 1;
  // the host-made arrays go through the plan's own pinned buffer (it stays untouched until the
  // plan is rebuilt, which is stream-ordered after everything that reads it)
  {
    const size_t ni = (size_t)m.num_item, nw = (size_t)W + 1, nit = hp.items.size();
    const size_t words = 2 * ni + 2 * nw + nit + (size_t)W;
    if (p.stage.cap < words * 4) {
      if (p.stage.p) CU(h, cudaFreeHost(p.stage.p));
      p.stage.p = nullptr;
      p.stage.cap = 0;
      CU(h, cudaMallocHost(&p.stage.p, words * 4 + words));
      p.stage.cap = words * 4 + words;
    }
    unsigned *sp = (unsigned *)p.stage.p;
    unsigned *s_owner = sp, *s_slot = s_owner + ni, *s_qoff = s_slot + ni, *s_ioff = s_qoff + nw,
             *s_items = s_ioff + nw, *s_batch = s_items + nit;
    memcpy(s_owner, hp.item_owner.data(), ni * 4);
    memcpy(s_slot, hp.item_slot.data(), ni * 4);
    memcpy(s_qoff, hp.queue_off.data(), nw * 4);
    memcpy(s_ioff, hp.item_off.data(), nw * 4);
    if (nit) memcpy(s_items, hp.items.data(), nit * 4);
    memcpy(s_batch, hp.batch.data(), (size_t)W * 4);
    CU(h, cudaMemcpyAsync(s.item_owner.p, s_owner, ni * 4, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(s.item_slot.p, s_slot, ni * 4, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(p.queue_off.p, s_qoff, nw * 4, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(p.item_off.p, s_ioff, nw * 4, cudaMemcpyHostToDevice, st));
    if (nit) CU(h, cudaMemcpyAsync(p.items.p, s_items, nit * 4, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(p.batch.p, s_batch, (size_t)W * 4, cudaMemcpyHostToDevice, st));
    h->n_h2d += (long long)words * 4;
  }

  // queues: stable sort by owner; the sorted position is the queue position
  k_own_keyowner<<<blocks, 256, 0, st>>>((const unsigned *)s.key_item.p, (const int *)s.item_owner.p, n, keyA, valA);
  CU(h, cudaGetLastError());
  {
    cub::DoubleBuffer<unsigned> dk(keyA, keyB), dv(valA, valB);
    tmp_bytes = s.tmp.cap;
    CU(h, cub::DeviceRadixSort::SortPairs(s.tmp.p, tmp_bytes, dk, dv, n, 0, bits_for((unsigned)W), st));
    k_own_entries<<<blocks, 256, 0, st>>>(csr, r0, n, dv.Current(), (const unsigned *)s.tick.p,
                                          (const unsigned *)s.item_slot.p, (OwnEntry *)p.entries.p);
    CU(h, cudaGetLastError());
    h->n_launch += 3;
  }
  p.num_owner = W;
  p.per_cta = per_cta;
  p.rows = n;
  p.max_load = hp.max_load;
  p.valid = true;
  return 0;
}

template <int CH, int VEC, int D>
static int own_launch_as(svdgpu *h, const OwnPlan &p, cudaStream_t st) {
  const DevModel &m = h->dm;
  const unsigned row_bytes = (unsigned)m.pitch * 4u, slot_bytes = row_bytes + OWN_SLOT_EXTRA;
  const unsigned ring = D * slot_bytes;
  // dot scratch: the generic routine's, or two transposed product buffers of the fast link
  const unsigned dot = std::max<unsigned>((unsigned)Group<32, VEC>::DOT_FLOATS * 4u, 2u * 4u * 36u * 4u);
  const unsigned bars = 2u * D * 8u;
  int max_smem = 0;
  CU(h, cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
  const long long fixed = (long long)ring + dot + bars + 128;
  long long S = ((long long)max_smem / OWN_C - fixed) / (long long)(row_bytes + 4);
  S = std::min<long long>(S, h->own_slots > 0 ? h->own_slots : 32);
  if (S < 0) return fail(h, "ordered mode: shared memory too small for k_own at num_factor %d", m.k);
  OwnArgs a;
  a.m = m;
  a.hp = h->dhp;
  a.entries = (const uint4 *)p.entries.p;
  a.queue_off = (const int *)p.queue_off.p;
  a.item_off = (const int *)p.item_off.p;
  a.items = (const unsigned *)p.items.p;
  a.batch = (const int *)p.batch.p;
  a.S = (int)S;
  a.off_items = ring;
  a.off_ibias = a.off_items + (unsigned)S * row_bytes;
  a.off_dot = (a.off_ibias + (unsigned)S * 4u + 15u) & ~15u;
  a.off_bars = a.off_dot + dot;
  a.region_bytes = (a.off_bars + bars + 127u) & ~127u;
  a.err_flag = h->d_err;
  a.abort_flag = h->d_abort;
  a.flags = (h->own_acquire ? OWN_F_ACQUIRE : 0) | (h->own_reverse ? OWN_F_REVERSE : 0);
  a.poll_ns = (unsigned)h->own_poll_ns;
  a.stats = nullptr;
  if (h->own_stats) {
    if (own_reserve(h, h->own.stats, (size_t)p.num_owner * 4 * sizeof(long long))) return 1;
    CU(h, cudaMemsetAsync(h->own.stats.p, 0, (size_t)p.num_owner * 4 * sizeof(long long), st));
    a.stats = (long long *)h->own.stats.p;
    h->own.stats_owners = p.num_owner;
  }
  const size_t smem = (size_t)a.region_bytes * OWN_C;
  auto k = k_own<CH, VEC, D>;
  CU(h, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CU(h, cudaMemsetAsync(m.ver_ui + m.user_off, 0, sizeof(unsigned) * (size_t)m.num_user, st));
  CU(h, cudaMemsetAsync(h->d_abort, 0, sizeof(unsigned), st));
  void *args[] = {&a};
  // cooperative: the launch fails instead of deadlocking if the CTAs cannot all be resident
  CU(h, cudaLaunchCooperativeKernel((void *)k, dim3(p.num_owner / OWN_C), dim3(own_threads(D)), args, smem, st));
  h->n_launch++;
  h->n_own++;
  h->n_own_rows += p.rows;
  return 0;
}

template <int CH, int D>
static int own_launch2d(svdgpu *h, const OwnPlan &p, cudaStream_t st) {
  const DevModel &m = h->dm;
  const unsigned row_bytes = (unsigned)m.pitch * 4u, slot_bytes = row_bytes + OWN_SLOT_EXTRA;
  const unsigned ring = D * slot_bytes, dot = 2u * 4u * 36u * 4u, hand = OWN2_H * (32u * 16u + 16u);
  const unsigned bars = (2u * D + 2u * OWN2_H) * 8u;
  int max_smem = 0;
  CU(h, cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
  const long long fixed = (long long)ring + dot + hand + bars + 128;
  long long S = ((long long)max_smem / OWN2_C - fixed) / (long long)(row_bytes + 4);
  S = std::min<long long>(S, h->own_slots > 0 ? h->own_slots : 32);
  if (S < 0) return fail(h, "ordered mode: shared memory too small for k_own2 at num_factor %d", m.k);
  OwnArgs a;
  a.m = m;
  a.hp = h->dhp;
  a.entries = (const uint4 *)p.entries.p;
  a.queue_off = (const int *)p.queue_off.p;
  a.item_off = (const int *)p.item_off.p;
  a.items = (const unsigned *)p.items.p;
  a.batch = (const int *)p.batch.p;
  a.S = (int)S;
  a.off_items = ring;
  a.off_ibias = a.off_items + (unsigned)S * row_bytes;
  a.off_dot = (a.off_ibias + (unsigned)S * 4u + 15u) & ~15u;
  a.off_hand = a.off_dot + dot;
  a.off_bars = a.off_hand + hand;
  a.region_bytes = (a.off_bars + bars + 127u) & ~127u;
  a.err_flag = h->d_err;
  a.abort_flag = h->d_abort;
  a.flags = (h->own_acquire ? OWN_F_ACQUIRE : 0) | (h->own_reverse ? OWN_F_REVERSE : 0);
  a.poll_ns = (unsigned)h->own_poll_ns;
  a.stats = nullptr;
  if (h->own_stats) {
    if (own_reserve(h, h->own.stats, (size_t)p.num_owner * 4 * sizeof(long long))) return 1;
    CU(h, cudaMemsetAsync(h->own.stats.p, 0, (size_t)p.num_owner * 4 * sizeof(long long), st));
    a.stats = (long long *)h->own.stats.p;
    h->own.stats_owners = p.num_owner;
  }
  const size_t smem = (size_t)a.region_bytes * OWN2_C;
  auto k = k_own2<CH, D>;
  CU(h, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CU(h, cudaMemsetAsync(m.ver_ui + m.user_off, 0, sizeof(unsigned) * (size_t)m.num_user, st));
  CU(h, cudaMemsetAsync(h->d_abort, 0, sizeof(unsigned), st));
  void *args[] = {&a};
  CU(h, cudaLaunchCooperativeKernel((void *)k, dim3(p.num_owner / OWN2_C), dim3(own2_threads(D)), args, smem, st));
  h->n_launch++;
  h->n_own++;
  h->n_own_rows += p.rows;
  return 0;
}

template <int CH>
static int own_launch2(svdgpu *h, const OwnPlan &p, cudaStream_t st) {
  return own_launch2d<CH, 8>(h, p, st);  // (ring depth 16 measured slower here too: 26.3 vs 22.4 ms per 20 M rows)
}

template <int CH, int VEC>
static int own_launch_d(svdgpu *h, const OwnPlan &p, cudaStream_t st) {
  if (h->own_depth == 16 && h->dm.pitch <= 128) return own_launch_as<CH, VEC, 16>(h, p, st);
  return own_launch_as<CH, VEC, 8>(h, p, st);
}

int launch_own(svdgpu *h, const OwnPlan &p, cudaStream_t st) {
  if (!p.valid) return fail(h, "ordered mode: no owner plan");
  if (p.num_owner <= 0 || p.per_cta <= 0 || p.num_owner % p.per_cta || p.num_owner > h->num_sm * p.per_cta)
    return fail(h, "ordered mode: plan was built for another device");
  const DevModel &m = h->dm;
  const int chunks = m.pitch / 4;
  if (p.per_cta == OWN2_C) {  // planned for the split link (owner + partner warps)
    if (!own_fast_shape(h)) return fail(h, "ordered mode: the plan was built for k_own2; re-create the batch after changing own_fast");
    if (chunks == 16) return own_launch2<16>(h, p, st);
    if (chunks == 32) return own_launch2<32>(h, p, st);
    if (chunks == 8) return own_launch2<8>(h, p, st);
    return own_launch2<4>(h, p, st);
  }
  if (p.per_cta != OWN_C) return fail(h, "ordered mode: plan was built by another version of the library");
  // the fast link: linear loss, rows of exactly 4 / 8 / 16 / 32 chunks (num_factor 16, 32, 64, 128)
  if (h->own_fast && m.active_type == 0 && m.k == m.pitch) {
    if (chunks == 16) return own_launch_d<16, 1>(h, p, st);
    if (chunks == 32) return own_launch_d<32, 1>(h, p, st);
    if (chunks == 8) return own_launch_d<8, 1>(h, p, st);
    if (chunks == 4) return own_launch_d<4, 1>(h, p, st);
  }
  if (chunks <= 32) return own_launch_d<0, 1>(h, p, st);
  if (chunks <= 64) return own_launch_d<0, 2>(h, p, st);
  return fail(h, "ordered mode: k_own supports num_factor <= 256");
}

}  // namespace svdk
