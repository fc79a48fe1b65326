// svdgpu_api.cu -- implementation of include/svdgpu.h: handle, HBM layout,
// host<->device staging, ticket computation for the ordered mode and kernel
// dispatch.  No CPU compute path exists here: every update/predict call ends in
// a kernel launch or fails.
#include "svdgpu_internal.h"
#include "svdgpu_scan.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <system_error>
#include <thread>
#include <vector>

using namespace svdk;

namespace {
thread_local std::string g_create_error;
}  // namespace


namespace {

}  // namespace

namespace svdk {
int fail(svdgpu *h, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  else g_create_error = buf;
  return 1;
}
}  // namespace svdk

// RMSEEvaluator::add_eval (svd_feature_infer.cpp:45-49): diff = (pred - label) * scale in fp32,
// diff*diff accumulated in (long) double.  acc[0] += sum of squares, acc[1] += count.
__global__ void k_sse(const float *__restrict__ pred, const float *__restrict__ label, int n, float scale,
                      double *acc) {
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double d = (double)__fmul_rn(__fsub_rn(pred[i], label[i]), scale);
    s += d * d;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(acc, s);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(acc + 1, (double)n);
}

// L2 / HBM bandwidth probes (svdgpu_microbench): every thread streams float4s of a buffer `iters`
// times; a buffer that fits the 126 MB L2 measures the L2 read rate the gather kernels live on
// (the 128 MB model of configs[1] is almost L2-resident), a larger one the HBM read rate.
__global__ void k_probe_read(const float4 *__restrict__ p, long long n4, int iters, float *sink) {
  float acc = 0.0f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; ++it)
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
      const float4 v = __ldcg(p + i);
      acc += v.x + v.y + v.z + v.w;
    }
  if (acc == 12345.678f) *sink = acc;  // (keeps the loads alive)
}
__global__ void k_probe_copy(const float4 *__restrict__ src, float4 *__restrict__ dst, long long n4, int iters) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; ++it)
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) __stcg(dst + i, __ldcg(src + i));
}

// Hogwild stability guard: how often does the most frequent item occur among rows [r0, r0+n)?
__global__ void k_item_hist(DevCsr csr, int r0, int n, unsigned num_item, unsigned *cnt) {
  const unsigned *idx = csr.index - csr.val_base;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = r0 + i;
    const int a = csr.row_ptr[3 * r + 2], b = csr.row_ptr[3 * r + 3];
    if (a < csr.val_base || b > csr.val_end || b - a > 64) continue;  // (bad rows are the kernels' business)
    for (int f = a; f < b; ++f)
      if (idx[f] < num_item) atomicAdd(cnt + idx[f], 1u);
  }
}
__global__ void k_max_u32(const unsigned *cnt, unsigned n, unsigned *out) {
  unsigned m = 0;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, cnt[i]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

namespace {

// share of the hottest item among the rows: the measurement is enqueued on the launch stream and lands in
// pinned memory {count of the hottest item, rows}; `wait` makes the caller wait for it (one small sync)
int hot_measure(svdgpu *h, const DevCsr &csr, int r0, int n, bool wait, double *frac) {
  const size_t ni = (size_t)h->shape.num_item;
  if (n <= 0 || ni == 0) {
    if (frac) *frac = 0.0;
    return 0;
  }
  if (h->hist_cap < (ni + 1) * 4) {
    if (h->d_hist) CU(h, cudaFree(h->d_hist));
    h->d_hist = nullptr;
    h->hist_cap = 0;
    CU(h, cudaMalloc(&h->d_hist, (ni + 1) * 4 + 64));
    h->hist_cap = (ni + 1) * 4;
  }
  if (!h->h_hot) {
    CU(h, cudaMallocHost(&h->h_hot, 2 * sizeof(unsigned)));
    CU(h, cudaEventCreateWithFlags(&h->ev_hot, cudaEventDisableTiming));
  }
  CU(h, cudaMemsetAsync(h->d_hist, 0, (ni + 1) * 4, h->stream));
  k_item_hist<<<(int)std::min<long long>(h->num_sm * 16, ((long long)n + 255) / 256), 256, 0, h->stream>>>(
      csr, r0, n, (unsigned)ni, h->d_hist);
  k_max_u32<<<std::min<int>(h->num_sm * 4, (int)((ni + 255) / 256)), 256, 0, h->stream>>>(h->d_hist, (unsigned)ni, h->d_hist + ni);
  CU(h, cudaGetLastError());
  h->n_launch += 2;
  CU(h, cudaMemcpyAsync(h->h_hot, h->d_hist + ni, 4, cudaMemcpyDeviceToHost, h->stream));
  h->h_hot[1] = (unsigned)n;
  h->n_d2h += 4;
  if (!wait) {
    CU(h, cudaEventRecord(h->ev_hot, h->stream));
    h->hot_pending = 1;
    return 0;
  }
  CU(h, cudaStreamSynchronize(h->stream));
  h->hot_pending = 0;
  if (frac) *frac = (double)h->h_hot[0] / (double)n;
  return 0;
}
// Host-pointer Hogwild training calls: the first call measures its first chunk and waits for the answer; every
// later call uses the share the call before it measured and enqueues a measurement of its own (no sync: a
// stream of small calls -- the per-instance seam flushing 2^20 rows at a time -- would otherwise drain the
// device once per call).
int hot_for_call(svdgpu *h, const DevCsr &csr, int n) {
  if (h->hot_pending && cudaEventQuery(h->ev_hot) == cudaSuccess) {
    h->hot_pending = 0;
    if (h->h_hot[1]) h->hot_call = (double)h->h_hot[0] / (double)h->h_hot[1];
  }
  if (h->hot_call < 0.0) return hot_measure(h, csr, 0, n, true, &h->hot_call);
  if (!h->hot_pending) return hot_measure(h, csr, 0, n, false, nullptr);
  return 0;
}
// cap on the instances a Hogwild training launch keeps in flight (0 = none)
void set_inflight_cap(svdgpu *h, double hot_frac) {
  h->hot_frac = hot_frac;
  h->inflight_cap = 0;
  const double lr = (double)h->hp.learning_rate;
  if (h->hog_safety_permille <= 0 || !(hot_frac > 0.0) || !(lr > 0.0)) return;
  const double cap = 1e-3 * h->hog_safety_permille / (lr * hot_frac);
  h->inflight_cap = cap > 4e18 ? 0 : std::max<long long>(64, (long long)cap);
}

// squared-error accumulation of one predicted chunk (only while an svdgpu_eval_* call is active)
int eval_chunk(svdgpu *h, const float *pred, const float *label, int n) {
  if (!h->eval_on || n <= 0) return 0;
  const int grid = std::min(h->num_sm * 8, (n + 255) / 256);
  k_sse<<<grid, 256, 0, h->stream>>>(pred, label, n, h->eval_scale, h->d_eval);
  CU(h, cudaGetLastError());
  h->n_launch++;
  return 0;
}

int dev_reserve(svdgpu *h, DevBuf &b, size_t bytes) {
  bytes += 64;  // bulk copies read 16-byte windows past the last element
  if (bytes <= b.cap) return 0;
  if (b.p) CU(h, cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  b.fill_n = 0;
  size_t cap = bytes + bytes / 4;
  CU(h, cudaMalloc(&b.p, cap));
  b.cap = cap;
  return 0;
}
int host_reserve(svdgpu *h, HostBuf &b, size_t bytes) {
  if (bytes <= b.cap) return 0;
  if (b.p) CU(h, cudaFreeHost(b.p));
  b.p = nullptr;
  b.cap = 0;
  size_t cap = bytes + bytes / 4 + 64;
  CU(h, cudaMallocHost(&b.p, cap));
  b.cap = cap;
  return 0;
}
void dev_free(DevBuf &b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}
void host_free(HostBuf &b) {
  if (b.p) cudaFreeHost(b.p);
  b.p = nullptr;
  b.cap = 0;
}

bool is_pinned(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// H2D of a host array: straight from the caller's memory when it is pinned,
// else through the slot's pinned mirror.
int h2d(svdgpu *h, DevBuf &d, HostBuf &stage, const void *src, size_t bytes) {
  if (dev_reserve(h, d, bytes)) return 1;
  if (bytes == 0) return 0;
  d.fill_n = 0;
  const void *from = src;
  if (!is_pinned(src)) {
    if (host_reserve(h, stage, bytes)) return 1;
    memcpy(stage.p, src, bytes);
    from = stage.p;
  }
  CU(h, cudaMemcpyAsync(d.p, from, bytes, cudaMemcpyHostToDevice, h->copy_stream));
  h->n_h2d += (long long)bytes;
  return 0;
}
// H2D of library-owned pinned memory into a slot buffer (tickets, block tables)
int h2d_pinned(svdgpu *h, void *dst, const void *src, size_t bytes) {
  if (bytes == 0) return 0;
  CU(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->copy_stream));
  h->n_h2d += (long long)bytes;
  return 0;
}

// ---- hyper-parameters -> device constants (all fp32, reference expressions) ----
bool is_one_host(float s) { return !(std::fabs((double)(s - 1.0f)) > 1e-6); }

void refresh_hp(svdgpu *h) {
  const svdgpu_hparams &p = h->hp;
  DevHP &d = h->dhp;
  d.lr = p.learning_rate;
  volatile float lu = p.learning_rate * p.wd_user;  // base.h:213-214
  volatile float li = p.learning_rate * p.wd_item;  // base.h:253-254
  d.du = 1.0f - lu;
  d.di = 1.0f - li;
  d.du_skip = is_one_host(d.du);
  d.di_skip = is_one_host(d.di);
  volatile float lub = p.learning_rate * p.wd_user_bias;  // base.h:248
  volatile float lib = p.learning_rate * p.wd_item_bias;  // base.h:282
  d.dub = 1.0f - lub;
  d.dib = 1.0f - lib;
  volatile float lg = p.learning_rate * p.wd_global;  // base.h:189,192
  d.dg = 1.0f - lg;
  d.regfree = p.num_regfree_global;
  d.base_score = p.base_score;
  volatile float lrfb = p.learning_rate * p.scale_lr_ufeedback;  // base.h:513
  d.lr_fb = lrfb;
  volatile float lf = lrfb * p.wd_ufeedback;  // base.h:515
  d.dfb = 1.0f - lf;
  d.dfb_skip = is_one_host(d.dfb);
  volatile float lfb = lrfb * p.wd_ufeedback_bias;  // base.h:518
  d.dfbb = 1.0f - lfb;
  // reg_method 3: L1 on user rows, plain decay on item rows (base.h:217,256)
  d.reg_user = p.reg_method == 3 ? 1 : p.reg_method;
  d.reg_item = p.reg_method == 3 ? 0 : p.reg_method;
  d.reg_global = p.reg_global;
  d.l1_u = lu;
  d.l1_i = li;
  d.pb_u = p.wd_user;
  d.pb_i = p.wd_item;
  d.l1_g = lg;
  d.user_nonneg = p.user_nonnegative != 0;
  // ranged weight decay (d.ru / d.ri / d.rg, set by svdgpu_set_wd_ranges) is looked up per index
  // by the generic routine only
  d.plain = (d.reg_user == 0 && d.reg_item == 0 && d.reg_global == 0 && !d.user_nonneg && d.ru.n == 0 &&
             d.ri.n == 0 && d.rg.n == 0) ? 1 : 0;
}

// ---- lane geometry -----------------------------------------------------------
// chunks = pitch/4 float4 per row.  Automatic choice: two chunks per lane (measured
// fastest on B200: more instances per warp, same bytes in flight), at least 4 lanes.
int pick_geometry(svdgpu *h, Geometry &g) {
  const int chunks = h->dm.pitch / 4;
  int lanes = h->lanes_opt;
  if (lanes == 0) {
    lanes = 4;
    while (lanes * 2 < chunks && lanes < 32) lanes <<= 1;
  }
  if (lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32)
    return fail(h, "option lanes must be 0, 4, 8, 16 or 32");
  int vec = (chunks + lanes - 1) / lanes;
  if (vec <= 1) vec = 1;
  else if (vec <= 2) vec = 2;
  else if (vec <= 4) vec = 4;
  else return fail(h, "num_factor %d needs more than 4 float4 chunks per lane with %d lanes (max k = %d)",
                   h->shape.num_factor, lanes, lanes * 16);
  g.lanes = lanes;
  g.vec = vec;
  return 0;
}

// ---- input validation + tickets (ordered mode) --------------------------------
// ticket[f] = how many earlier feature occurrences (in input order, earlier
// instances only) touched the same row; the kernel waits for the row's version
// counter to reach it.  Also the reference's index asserts (base.h:320,327,343).
int validate_csr(svdgpu *h, int num_row, const int *row_ptr) {
  if (num_row < 0) return fail(h, "num_row < 0");
  for (long long i = 0; i < 3LL * num_row; ++i)
    if (row_ptr[i + 1] < row_ptr[i]) return fail(h, "row_ptr must be non-decreasing");
  return 0;
}

int make_tickets(svdgpu *h, int r0, int r1, const int *row_ptr, const unsigned *index,
                 unsigned *ticket /* indexed from row_ptr[3*r0] */) {
  const int base = row_ptr[3LL * r0];
  const unsigned nu = (unsigned)h->shape.num_user, ni = (unsigned)h->shape.num_item,
                 ng = (unsigned)h->shape.num_global;
  unsigned *cu = h->cnt_ui.data() + h->dm.user_off, *ci = h->cnt_ui.data() + h->dm.item_off,
           *cg = h->cnt_g.data();
  for (int r = r0; r < r1; ++r) {
    const int *p = row_ptr + 3LL * r;
    for (int f = p[0]; f < p[1]; ++f) {
      if (index[f] >= ng) return fail(h, "global feature index exceed setting");
      ticket[f - base] = cg[index[f]];
    }
    for (int f = p[1]; f < p[2]; ++f) {
      if (index[f] >= nu) return fail(h, "user feature index exceed bound");
      ticket[f - base] = cu[index[f]];
    }
    for (int f = p[2]; f < p[3]; ++f) {
      if (index[f] >= ni) return fail(h, "item feature index exceed bound");
      ticket[f - base] = ci[index[f]];
    }
    for (int f = p[0]; f < p[1]; ++f) cg[index[f]]++;
    for (int f = p[1]; f < p[2]; ++f) cu[index[f]]++;
    for (int f = p[2]; f < p[3]; ++f) ci[index[f]]++;
  }
  return 0;
}

void reset_ticket_counters(svdgpu *h) {
  h->cnt_ui.assign(std::max<size_t>(h->rows, 1), 0u);
  h->cnt_g.assign((size_t)std::max(h->shape.num_global, 1), 0u);
}

// ---- side-feature expansion (feature_user / feature_item) ---------------------------------
// The reference walks, after every user feature uid, the extra pairs feat_user[uid] (value used
// as is), and after every item feature (iid, ival) the pairs feat_item[iid] whose value enters
// as value*ival (base.h:298-308,330-349,365-379,399-422).  A chunk is rewritten on the host into
// a plain CSR with those extra entries in the reference's order; item entries carry the second
// factor in val2 (1 for the base entry, ival for its extras) so that the device reproduces the
// reference's products exactly.  Host work is O(nnz): this path is for completeness, not speed.
struct Expanded {
  std::vector<int> rp;
  std::vector<unsigned> idx;
  std::vector<float> val, val2;
};
bool sides_on(const svdgpu *h) { return h->side_u.on() || h->side_i.on(); }
int expand_rows(svdgpu *h, int r0, int r1, const int *row_ptr, const unsigned *index, const float *value,
                Expanded &e) {
  e.rp.assign(1, 0);
  e.idx.clear();
  e.val.clear();
  e.val2.clear();
  auto push = [&](unsigned i, float v, float v2) {
    e.idx.push_back(i);
    e.val.push_back(v);
    e.val2.push_back(v2);
  };
  for (int r = r0; r < r1; ++r) {
    const int *p = row_ptr + 3LL * r;
    if (p[0] < 0 || p[1] < p[0] || p[2] < p[1] || p[3] < p[2]) return fail(h, "row_ptr must be non-decreasing");
    for (int f = p[0]; f < p[1]; ++f) push(index[f], value[f], 1.0f);
    e.rp.push_back((int)e.idx.size());
    for (int f = p[1]; f < p[2]; ++f) {
      push(index[f], value[f], 1.0f);
      const svdgpu::Side &sd = h->side_u;
      if (sd.on() && index[f] + 1 < sd.rp.size())
        for (unsigned j = sd.rp[index[f]]; j < sd.rp[index[f] + 1]; ++j) push(sd.idx[j], sd.val[j], 1.0f);
    }
    e.rp.push_back((int)e.idx.size());
    for (int f = p[2]; f < p[3]; ++f) {
      push(index[f], value[f], 1.0f);
      const svdgpu::Side &sd = h->side_i;
      if (sd.on() && index[f] + 1 < sd.rp.size())
        for (unsigned j = sd.rp[index[f]]; j < sd.rp[index[f] + 1]; ++j) push(sd.idx[j], sd.val[j], value[f]);
    }
    e.rp.push_back((int)e.idx.size());
    if (e.idx.size() > 0x7fffffffULL) return fail(h, "expanded batch exceeds 2^31 feature entries: lower chunk_rows");
  }
  return 0;
}

int check_ready(svdgpu *h) {
  if (!h) return 1;
  if (!h->hp_set) return fail(h, "svdgpu_set_hparams has not been called");
  if (h->hp.reg_method < 0 || h->hp.reg_method > 3)
    return fail(h, "reg_method=%d is not supported on the GPU path (0 L2 decay, 1 L1, 2 projection, 3 L1 user/L2 item)",
                h->hp.reg_method);
  if (h->hp.reg_global != 0 && h->hp.reg_global != 1)
    return fail(h, "reg_global=%d is not supported on the GPU path (0 L2 decay, 1 L1)", h->hp.reg_global);
  return 0;
}

// Staging pipeline of the host-pointer calls: the H2D copies of a chunk go to the copy
// stream into one of NSLOT slots, the kernels of that chunk wait for them on the launch
// stream (slot_copied), and the slot is reused once its kernels are done (next_slot), so
// the copy of chunk c+1 overlaps the kernels of chunk c.
Slot &next_slot(svdgpu *h) {
  Slot &s = h->slot[h->cur_slot];
  h->cur_slot = (h->cur_slot + 1) % svdgpu::NSLOT;
  if (s.used) cudaEventSynchronize(s.done);
  return s;
}
int slot_copied(svdgpu *h, Slot &s) {
  CU(h, cudaEventRecord(s.copied, h->copy_stream));
  CU(h, cudaStreamWaitEvent(h->stream, s.copied, 0));
  return 0;
}
int slot_done(svdgpu *h, Slot &s) {
  CU(h, cudaEventRecord(s.done, h->stream));
  s.used = true;
  return 0;
}
// host arrays are borrowed for the call only: wait until the copies have read them
int copies_drained(svdgpu *h) {
  CU(h, cudaEventRecord(h->ev_copy, h->copy_stream));
  CU(h, cudaEventSynchronize(h->ev_copy));
  return 0;
}

// Hogwild unit order: input order (it is the SGD order the caller shuffled), except that
// the rare very long users (> 8x the mean, >= 512 rows) start first so that none of them
// becomes the tail of the launch.
template <typename SizeOf>
void unit_order(int nu, SizeOf size_of, int *order) {
  long long total = 0;
  for (int u = 0; u < nu; ++u) total += size_of(u);
  const long long thr = std::max<long long>(512, nu > 0 ? 8 * total / nu : 0);
  int n = 0;
  for (int u = 0; u < nu; ++u)
    if (size_of(u) > thr) order[n++] = u;
  for (int u = 0; u < nu; ++u)
    if (!(size_of(u) > thr)) order[n++] = u;
}

// units = maximal runs DEFAULT | START (MIDDLE)* END
int build_units(svdgpu *h, int num_block, const int *blk_tag, std::vector<int> &unit_off) {
  unit_off.clear();
  unit_off.push_back(0);
  bool open = false;
  for (int b = 0; b < num_block; ++b) {
    const int tag = blk_tag ? blk_tag[b] : 0;
    if (tag == 0) {
      if (open) return fail(h, "DEFAULT block inside a START..END run");
      unit_off.push_back(b + 1);
    } else if (tag == 1) {
      if (open) return fail(h, "START block inside a START..END run");
      open = true;
    } else if (tag == 2) {
      if (!open) return fail(h, "END block without START");
      open = false;
      unit_off.push_back(b + 1);
    } else if (tag == 3) {
      if (!open) return fail(h, "MIDDLE block without START");
    } else {
      return fail(h, "unknown extend_tag %d", tag);
    }
  }
  if (open) return fail(h, "START..END run not closed inside the call");
  return 0;
}

}  // namespace

// =============================================================================
extern "C" {

const char *svdgpu_last_error(const svdgpu_t *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int svdgpu_create(svdgpu_t **out, const svdgpu_shape *shape, int device) {
  if (!out || !shape) return fail(nullptr, "svdgpu_create: null argument");
  *out = nullptr;
  if (shape->num_user < 0 || shape->num_item < 0 || shape->num_global < 0 || shape->num_ufeedback < 0 ||
      shape->num_factor <= 0)
    return fail(nullptr, "svdgpu_create: bad shape");
  switch (shape->active_type) {
    case 0: case 1: case 2: case 3: case 5: case 6: case 7: break;
    default: return fail(nullptr, "unkown active type");
  }
  if (shape->format_type != 0 && shape->format_type != 1)
    return fail(nullptr, "svdgpu_create: format_type must be 0 (random order) or 1 (user group)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, "svdgpu_create: no CUDA device (this library has no CPU path)");
  }
  if (device < 0 || device >= ndev) return fail(nullptr, "svdgpu_create: device %d out of range", device);
  svdgpu *h = new svdgpu();
  h->shape = *shape;
  h->device = device;
#define CUC(call)                                                                        \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      fail(nullptr, "%s failed: %s", #call, cudaGetErrorString(e_));                     \
      svdgpu_destroy(h);                                                                 \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)
  CUC(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUC(cudaGetDeviceProperties(&prop, device));
  h->num_sm = prop.multiProcessorCount;
  // The launch stream outranks the plan stream: an ordered host-pointer call plans chunk c+1 beside the k_own
  // launch of chunk c, and the cooperative launch of chunk c+1 needs its SMs all at once -- the blocks of a plan
  // kernel already queued for chunk c+2 must not be handed them first.
  int prio_least = 0, prio_greatest = 0;
  CUC(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
  CUC(cudaStreamCreateWithPriority(&h->own_stream, cudaStreamNonBlocking, prio_greatest));
  h->stream = h->own_stream;
  CUC(cudaEventCreate(&h->ev0));
  CUC(cudaEventCreate(&h->ev1));
  CUC(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
  CUC(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  CUC(cudaStreamCreateWithPriority(&h->plan_stream, cudaStreamNonBlocking, prio_least));
  CUC(cudaEventCreateWithFlags(&h->ev_plan, cudaEventDisableTiming));
  for (int i = 0; i < svdgpu::NSLOT; ++i) {
    CUC(cudaEventCreateWithFlags(&h->slot[i].done, cudaEventDisableTiming));
    CUC(cudaEventCreateWithFlags(&h->slot[i].copied, cudaEventDisableTiming));
    CUC(cudaEventCreateWithFlags(&h->slot[i].planned, cudaEventDisableTiming));
  }

  DevModel &m = h->dm;
  memset(&m, 0, sizeof(m));
  m.k = shape->num_factor;
  m.pitch = ((shape->num_factor + 3) >> 2) << 2;  // sse.h:26-27: 16-byte row pitch
  m.num_user = shape->num_user;
  m.num_item = shape->num_item;
  m.num_global = shape->num_global;
  m.num_ufeedback = shape->format_type == 1 ? shape->num_ufeedback : 0;
  m.user_off = m.num_ufeedback;                  // model.h:513
  m.item_off = m.user_off + shape->num_user;     // model.h:533-534
  m.no_user_bias = shape->no_user_bias;
  m.active_type = shape->active_type;
  h->rows = (size_t)m.item_off + (size_t)shape->num_item;
  const size_t wbytes = std::max<size_t>(h->rows, 1) * m.pitch * sizeof(float);
  CUC(cudaMalloc(&m.W, wbytes));
  CUC(cudaMemset(m.W, 0, wbytes));
  // (+4 floats: the fast pass fetches biases as aligned 16-byte windows)
  CUC(cudaMalloc(&m.bias, (std::max<size_t>(h->rows, 1) + 4) * sizeof(float)));
  CUC(cudaMemset(m.bias, 0, (std::max<size_t>(h->rows, 1) + 4) * sizeof(float)));
  CUC(cudaMalloc(&m.g_bias, (size_t)std::max(shape->num_global, 1) * sizeof(float)));
  CUC(cudaMemset(m.g_bias, 0, (size_t)std::max(shape->num_global, 1) * sizeof(float)));
  CUC(cudaMalloc(&m.ver_ui, std::max<size_t>(h->rows, 1) * sizeof(unsigned)));
  CUC(cudaMalloc(&m.ver_g, (size_t)std::max(shape->num_global, 1) * sizeof(unsigned)));
  CUC(cudaMalloc(&h->d_err, sizeof(int)));
  CUC(cudaMemset(h->d_err, 0, sizeof(int)));
  CUC(cudaMalloc(&h->d_counter, sizeof(unsigned)));
  CUC(cudaMalloc(&h->d_abort, sizeof(unsigned)));
  CUC(cudaMalloc(&h->d_eval, 2 * sizeof(double)));
#undef CUC
  Geometry g;
  if (pick_geometry(h, g)) {
    g_create_error = h->err;
    svdgpu_destroy(h);
    return 1;
  }
  *out = h;
  return 0;
}

void svdgpu_destroy(svdgpu_t *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  svdgpu_comm_destroy(h);
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  if (h->plan_stream) cudaStreamSynchronize(h->plan_stream);
  if (h->stream) cudaStreamSynchronize(h->stream);
  rank_free(h);
  cudaFree(h->dm.W);
  cudaFree(h->dm.bias);
  cudaFree(h->dm.g_bias);
  cudaFree(h->dm.ver_ui);
  cudaFree(h->dm.ver_g);
  cudaFree(h->d_err);
  cudaFree(h->d_counter);
  cudaFree(h->d_abort);
  cudaFree(h->d_hist);
  if (h->h_hot) cudaFreeHost(h->h_hot);
  if (h->ev_hot) cudaEventDestroy(h->ev_hot);
  own_scratch_free(h->own);
  for (int i = 0; i < svdgpu::NSLOT; ++i) own_plan_free(h->slot[i].own);
  cudaFree(h->d_eval);
  cudaFree(h->d_row_mask);
  cudaFree(h->d_unit_kind);
  cudaFree(h->d_snap);
  cudaFree(h->d_delta);
  {
    HostBuf *ib[] = {&h->ing_rp, &h->ing_label, &h->ing_index, &h->ing_value};
    for (HostBuf *b : ib) host_free(*b);
  }
  for (int i = 0; i < svdgpu::NSLOT; ++i) {
    Slot &s = h->slot[i];
    HostBuf *hb[] = {&s.h_rp, &s.h_label, &s.h_index, &s.h_value, &s.h_value2, &s.h_ticket, &s.h_misc, &s.h_fbi, &s.h_fbv, &s.h_fbt};
    for (HostBuf *b : hb) host_free(*b);
    DevBuf *db[] = {&s.d_rp, &s.d_label, &s.d_index, &s.d_value, &s.d_value2, &s.d_ticket, &s.d_misc, &s.d_fbi, &s.d_fbv, &s.d_fbt, &s.d_pred};
    for (DevBuf *b : db) dev_free(*b);
    if (s.done) cudaEventDestroy(s.done);
    if (s.copied) cudaEventDestroy(s.copied);
    if (s.planned) cudaEventDestroy(s.planned);
  }
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->ev_copy) cudaEventDestroy(h->ev_copy);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->plan_stream) cudaStreamDestroy(h->plan_stream);
  if (h->ev_plan) cudaEventDestroy(h->ev_plan);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
}

int svdgpu_set_hparams(svdgpu_t *h, const svdgpu_hparams *hp) {
  if (!h || !hp) return 1;
  h->hp = *hp;
  h->hp_set = true;
  refresh_hp(h);
  return 0;
}

int svdgpu_set_mode(svdgpu_t *h, int mode) {
  if (!h) return 1;
  if (mode != SVDGPU_MODE_EXACT && mode != SVDGPU_MODE_HOGWILD) return fail(h, "unknown mode %d", mode);
  h->mode = mode;
  return 0;
}

int svdgpu_set_option(svdgpu_t *h, const char *name, long long v) {
  if (!h || !name) return 1;
  if (!strcmp(name, "scatter_user")) h->scatter_user = v ? SCATTER_RED : SCATTER_STORE;
  else if (!strcmp(name, "scatter_item")) h->scatter_item = v ? SCATTER_RED : SCATTER_STORE;
  else if (!strcmp(name, "exact_dot")) h->exact_dot = v ? 1 : 0;
  else if (!strcmp(name, "stream_tile")) h->stream_tile = (int)v;
  else if (!strcmp(name, "mfg")) h->mfg = (int)v;
  else if (!strcmp(name, "rank_force_sort")) h->rank_force_sort = v ? 1 : 0;
  else if (!strcmp(name, "lanes")) {
    const int old = h->lanes_opt;
    h->lanes_opt = (int)v;
    Geometry g;
    if (pick_geometry(h, g)) {
      h->lanes_opt = old;
      return 1;
    }
  } else if (!strcmp(name, "chunk_rows")) {
    if (v < 1) return fail(h, "chunk_rows must be >= 1");
    h->chunk_rows = (int)std::min<long long>(v, 1LL << 28);
  } else if (!strcmp(name, "ctas_per_sm")) h->ctas_per_sm = (int)v;
  else if (!strcmp(name, "pass1")) h->pass1 = (int)std::max<long long>(0, std::min<long long>(v, 3));
  else if (!strcmp(name, "ring_depth")) h->ring_depth = (int)v;
  else if (!strcmp(name, "l2_ahead")) h->l2_ahead = (int)v;
  else if (!strcmp(name, "ugroup_units")) h->ugroup_units = (int)v;
  else if (!strcmp(name, "svdpp_fast")) h->svdpp_fast = v ? 1 : 0;
  else if (!strcmp(name, "mf_ctas")) h->mf_ctas = (int)v;
  else if (!strcmp(name, "exact_opt")) h->exact_opt = (int)v;
  else if (!strcmp(name, "exact_owner")) h->exact_owner = v ? 1 : 0;
  else if (!strcmp(name, "own_min_rows")) h->own_min_rows = (int)std::max<long long>(1, std::min<long long>(v, 1LL << 30));
  else if (!strcmp(name, "own_urgent_gap")) h->own_urgent_gap = (int)std::max<long long>(0, std::min<long long>(v, 1LL << 30));
  else if (!strcmp(name, "own_batch")) h->own_batch = (int)std::max<long long>(1, std::min<long long>(v, 32));
  else if (!strcmp(name, "hogwild_safety")) h->hog_safety_permille = (int)std::max<long long>(0, std::min<long long>(v, 1000000));
  else if (!strcmp(name, "own_stats")) h->own_stats = v ? 1 : 0;
  else if (!strcmp(name, "own_partner")) h->own_partner = v ? 1 : 0;
  else if (!strcmp(name, "own_poll_ns")) h->own_poll_ns = (int)std::max<long long>(0, std::min<long long>(v, 100000));
  else if (!strcmp(name, "own_spare_sms")) h->own_spare_sms = (int)std::max<long long>(0, std::min<long long>(v, 64));
  else if (!strcmp(name, "own_depth")) h->own_depth = v >= 16 ? 16 : 8;
  else if (!strcmp(name, "own_fast")) h->own_fast = v ? 1 : 0;
  else if (!strcmp(name, "own_acquire")) h->own_acquire = v ? 1 : 0;
  else if (!strcmp(name, "own_reverse")) h->own_reverse = v ? 1 : 0;
  else if (!strcmp(name, "own_isolate")) h->own_isolate = v < 0 ? 0 : v;
  else if (!strcmp(name, "own_redeal")) h->own_redeal = v < 0 ? 0 : v;
  else if (!strcmp(name, "own_plan_beside")) h->own_plan_beside = v ? 1 : 0;
  else if (!strcmp(name, "own_isolate_full")) h->own_isolate_full = v < 0 ? 0 : v;
  else if (!strcmp(name, "own_slots")) h->own_slots = (int)std::max<long long>(0, std::min<long long>(v, 32));
  else if (!strcmp(name, "compact_h2d")) h->compact_h2d = v < 0 ? 0 : v > 2 ? 2 : (int)v;
  else if (!strcmp(name, "scan_threads")) h->scan_threads = (int)std::max<long long>(0, std::min<long long>(v, 256));
  else if (!strcmp(name, "compact_min_rows")) h->compact_min_rows = (int)std::max<long long>(1, std::min<long long>(v, 1LL << 30));
  else return fail(h, "unknown option '%s'", name);
  return 0;
}

int svdgpu_set_stream(svdgpu_t *h, void *s) {
  if (!h) return 1;
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaStreamSynchronize(h->stream));
  h->stream = s ? (cudaStream_t)s : h->own_stream;
  return 0;
}

int svdgpu_set_side_features(svdgpu_t *h, int which, int num_row, const unsigned *row_ptr, const unsigned *index,
                             const float *value) {
  if (!h) return 1;
  if (which != 0 && which != 1) return fail(h, "set_side_features: which must be 0 (user) or 1 (item)");
  svdgpu::Side &sd = which == 0 ? h->side_u : h->side_i;
  sd.rp.clear();
  sd.idx.clear();
  sd.val.clear();
  if (num_row <= 0) return 0;
  if (!row_ptr || (row_ptr[num_row] > 0 && (!index || !value))) return fail(h, "set_side_features: null array");
  const unsigned bound = (unsigned)(which == 0 ? h->shape.num_user : h->shape.num_item);
  for (int r = 0; r < num_row; ++r)
    if (row_ptr[r + 1] < row_ptr[r]) return fail(h, "set_side_features: row_ptr must be non-decreasing");
  for (unsigned j = 0; j < row_ptr[num_row]; ++j)
    if (index[j] >= bound) return fail(h, which == 0 ? "user feature index exceed bound" : "item feature index exceed bound");
  sd.rp.assign(row_ptr, row_ptr + num_row + 1);
  sd.idx.assign(index, index + row_ptr[num_row]);
  sd.val.assign(value, value + row_ptr[num_row]);
  return 0;
}

int svdgpu_set_wd_ranges(svdgpu_t *h, int which, int n, const unsigned *bound, const float *wd) {
  if (!h) return 1;
  if (which < 0 || which > 2) return fail(h, "set_wd_ranges: which must be 0 (user), 1 (item) or 2 (global)");
  if (n < 0 || n > WD_MAX_RANGES) return fail(h, "set_wd_ranges: at most %d ranges", (int)WD_MAX_RANGES);
  if (n > 0 && (!bound || !wd)) return fail(h, "set_wd_ranges: null array");
  WdRanges r{};
  for (int j = 0; j < n; ++j) {
    // ParameterSet::set_param (base.h:56-58): "can't give 0 as bound", "bound must be given in order"
    if (bound[j] == 0) return fail(h, "can't give 0 as bound");
    if (j > 0 && !(bound[j - 1] < bound[j])) return fail(h, "bound must be given in order");
    r.bound[j] = bound[j] - 1;  // inclusive last index of the range (base.h:59)
    r.wd[j] = wd[j];
  }
  r.n = n;
  (which == 0 ? h->dhp.ru : which == 1 ? h->dhp.ri : h->dhp.rg) = r;
  if (h->hp_set) refresh_hp(h);
  return 0;
}

int svdgpu_upload_model(svdgpu_t *h, const float *ui_bias, const float *W, size_t pitch_floats,
                        const float *g_bias) {
  if (!h) return 1;
  CU(h, cudaSetDevice(h->device));
  const DevModel &m = h->dm;
  if (pitch_floats < (size_t)m.k) return fail(h, "upload_model: pitch %zu < num_factor %d", pitch_floats, m.k);
  if (h->rows) {
    CU(h, cudaMemsetAsync(m.W, 0, h->rows * m.pitch * sizeof(float), h->stream));
    CU(h, cudaMemcpy2DAsync(m.W, m.pitch * sizeof(float), W, pitch_floats * sizeof(float),
                            m.k * sizeof(float), h->rows, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(m.bias, ui_bias, h->rows * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  }
  if (m.num_global)
    CU(h, cudaMemcpyAsync(m.g_bias, g_bias, (size_t)m.num_global * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int svdgpu_download_model(svdgpu_t *h, float *ui_bias, float *W, size_t pitch_floats, float *g_bias) {
  if (!h) return 1;
  CU(h, cudaSetDevice(h->device));
  const DevModel &m = h->dm;
  if (pitch_floats < (size_t)m.k) return fail(h, "download_model: pitch %zu < num_factor %d", pitch_floats, m.k);
  if (svdgpu_sync(h)) return 1;
  if (h->rows) {
    CU(h, cudaMemcpy2DAsync(W, pitch_floats * sizeof(float), m.W, m.pitch * sizeof(float),
                            m.k * sizeof(float), h->rows, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(ui_bias, m.bias, h->rows * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  }
  if (m.num_global)
    CU(h, cudaMemcpyAsync(g_bias, m.g_bias, (size_t)m.num_global * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int svdgpu_sync(svdgpu_t *h) {
  if (!h) return 1;
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaStreamSynchronize(h->stream));
  int e = 0;
  CU(h, cudaMemcpy(&e, h->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (e != 0) {
    CU(h, cudaMemset(h->d_err, 0, sizeof(int)));
    switch (e) {  // the reference's assert messages, base.h:320,327,343,530
      case ERR_GLOBAL_INDEX: return fail(h, "global feature index exceed setting");
      case ERR_USER_INDEX: return fail(h, "user feature index exceed bound");
      case ERR_ITEM_INDEX: return fail(h, "item feature index exceed bound");
      case ERR_FB_INDEX: return fail(h, "ufeedback id exceed bound");
      case ERR_ROW_PTR: return fail(h, "row_ptr must be non-decreasing and inside the batch");
      case ERR_WD_BOUND: return fail(h, "bound set err");  // base.h:72
      case ERR_TIMEOUT: return fail(h, "ordered mode: a wait inside k_own timed out (internal error; the model is unusable)");
      default: return fail(h, "device error %d", e);
    }
  }
  return 0;
}

int svdgpu_timer_start(svdgpu_t *h) {
  if (!h) return 1;
  CU(h, cudaEventRecord(h->ev0, h->stream));
  return 0;
}
int svdgpu_timer_stop(svdgpu_t *h, float *ms) {
  if (!h || !ms) return 1;
  CU(h, cudaEventRecord(h->ev1, h->stream));
  CU(h, cudaEventSynchronize(h->ev1));
  CU(h, cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return 0;
}

long long svdgpu_get_counter(const svdgpu_t *h, const char *name) {
  if (!h || !name) return -1;
  if (!strcmp(name, "kernel_launches")) return h->n_launch;
  if (!strcmp(name, "instances")) return h->n_inst;
  if (!strcmp(name, "h2d_bytes")) return h->n_h2d;
  if (!strcmp(name, "d2h_bytes")) return h->n_d2h;
  if (!strcmp(name, "num_sm")) return h->num_sm;
  if (!strcmp(name, "inflight_cap")) return h->inflight_cap;  // Hogwild guard of the last training launch (0 = none)
  if (!strcmp(name, "hot_item_ppm")) return (long long)(h->hot_frac * 1e6);
  if (!strcmp(name, "collectives")) return h->n_coll;  // NCCL all-reduces issued by svdgpu_allreduce_items / _allgather_users
  if (!strcmp(name, "collective_bytes")) return h->n_coll_bytes;
  if (!strcmp(name, "own_launches")) return h->n_own;  // ordered mode: launches of the item-owner kernel
  if (!strcmp(name, "own_rows")) return h->n_own_rows;
  if (!strcmp(name, "own_deals")) return h->n_deal;      // plans that dealt the items out to owners (host LPT)
  if (!strcmp(name, "own_redeals")) return h->n_redeal;  // plans that carried the previous deal over
  if (!strcmp(name, "own_lpt_us")) return h->own_lpt_us;          // plan: host time dealing the items out
  if (!strcmp(name, "own_cntwait_us")) return h->own_cntwait_us;  // plan: host time waiting for the item counts
  if (!strcmp(name, "ingest_read_us")) return (long long)(h->ingest_read_s * 1e6);  // file -> pinned chunk
  if (!strcmp(name, "ingest_call_us")) return (long long)(h->ingest_call_s * 1e6);  // hot-path calls
  if (!strcmp(name, "lanes")) {
    Geometry g;
    return pick_geometry(const_cast<svdgpu *>(h), g) ? -1 : g.lanes;
  }
  return -1;
}

int svdgpu_own_stats(svdgpu_t *h, long long *out, int cap_owners, int *num_owner) {
  if (!h || !num_owner) return 1;
  CU(h, cudaSetDevice(h->device));
  const int W = h->own.stats_owners;
  *num_owner = W;
  if (!h->own.stats.p) return fail(h, "own_stats: set option own_stats before the ordered launch");
  if (out && cap_owners >= W) {
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, cudaMemcpy(out, h->own.stats.p, (size_t)W * 4 * sizeof(long long), cudaMemcpyDeviceToHost));
  }
  return 0;
}

int svdgpu_microbench(svdgpu_t *h, int which, size_t bytes, int iters, double *gbytes_per_s) {
  if (!h || !gbytes_per_s) return 1;
  if (which != 0 && which != 1) return fail(h, "microbench: which must be 0 (read) or 1 (copy)");
  if (bytes < 1024 || iters < 1) return fail(h, "microbench: bad size");
  CU(h, cudaSetDevice(h->device));
  const long long n4 = (long long)(bytes / 16);
  float4 *a = nullptr, *b = nullptr;
  float *sink = nullptr;
  CU(h, cudaMalloc(&a, (size_t)n4 * 16));
  CU(h, cudaMalloc(&sink, 4));
  if (which == 1) CU(h, cudaMalloc(&b, (size_t)n4 * 16));
  CU(h, cudaMemsetAsync(a, 0, (size_t)n4 * 16, h->stream));
  const int grid = h->num_sm * 8, block = 512;
  for (int rep = 0; rep < 2; ++rep) {  // the first pass warms L2 and the clocks
    CU(h, cudaEventRecord(h->ev0, h->stream));
    if (which == 0) k_probe_read<<<grid, block, 0, h->stream>>>(a, n4, iters, sink);
    else k_probe_copy<<<grid, block, 0, h->stream>>>(a, b, n4, iters);
    CU(h, cudaGetLastError());
    CU(h, cudaEventRecord(h->ev1, h->stream));
    CU(h, cudaEventSynchronize(h->ev1));
  }
  float ms = 0.0f;
  CU(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->n_launch += 2;
  *gbytes_per_s = (double)n4 * 16.0 * iters * (which == 1 ? 2.0 : 1.0) / (ms * 1e-3) / 1e9;
  cudaFree(a);
  cudaFree(b);
  cudaFree(sink);
  return 0;
}

void *svdgpu_device_ptr(svdgpu_t *h, int which, size_t *pitch_floats) {
  if (!h) return nullptr;
  if (pitch_floats) *pitch_floats = (size_t)h->dm.pitch;
  switch (which) {
    case 0: return h->dm.bias;
    case 1: return h->dm.W;
    case 2: return h->dm.g_bias;
    default: return nullptr;
  }
}

// ---------------------------------------------------------------------------
// Compact H2D for the Hogwild / predict host-pointer calls.  The reference's CSR costs 32 bytes
// per basic-MF instance on the bus (12 row_ptr + 4 label + 8 index + 8 value) and the call is
// PCIe-bound, yet two of the four arrays usually carry no information: row_ptr is an arithmetic
// progression when every row of a chunk has the same feature counts, and the values of
// indicator features are all 1.0f.  Host threads check this per chunk (exactly: every element is
// compared) while earlier chunks are being copied; an array that passes is not copied but
// rebuilt in the slot by a fill kernel.  What the kernels read is bit-identical either way.
// ---------------------------------------------------------------------------

__global__ void k_fill_row_ptr(int *rp, int n, int a, int b, int c) {
  const int w = a + b + c;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r <= n; r += (long long)gridDim.x * blockDim.x) {
    const int base = (int)(r * w);
    rp[3 * r] = base;
    if (r < n) {
      rp[3 * r + 1] = base + a;
      rp[3 * r + 2] = base + a + b;
    }
  }
}
__global__ void k_fill_ones(float *v, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    v[i] = 1.0f;
}

// Build an owner plan on the plan stream: it starts once `after` has happened (the chunk's copies, or
// whatever the launch stream had queued), and the launch stream waits for it.  The plan scratch of
// the handle is only ever touched on the plan stream.
static int plan_on_side(svdgpu *h, const DevCsr &csr, int n, OwnPlan &p, int *bad, int ctas, cudaEvent_t after,
                        cudaEvent_t done) {
  if (!after) {
    CU(h, cudaEventRecord(h->ev_plan, h->stream));
    after = h->ev_plan;
  }
  CU(h, cudaStreamWaitEvent(h->plan_stream, after, 0));
  if (!h->own_plan_beside) {  // (experiment: the plan only starts when the launch stream has drained)
    CU(h, cudaEventRecord(h->ev_plan, h->stream));
    CU(h, cudaStreamWaitEvent(h->plan_stream, h->ev_plan, 0));
  }
  if (own_plan_build(h, csr, 0, n, p, h->plan_stream, bad, ctas)) return 1;
  CU(h, cudaEventRecord(done, h->plan_stream));
  CU(h, cudaStreamWaitEvent(h->stream, done, 0));
  return 0;
}

// ---------------------------------------------------------------------------
// random-order CSR, host buffers
// ---------------------------------------------------------------------------
static int run_csr_host(svdgpu_t *h, int num_row, const int *row_ptr, const float *label,
                        const unsigned *index, const float *value, float *out, bool train) {
  if (check_ready(h)) return 1;
  CU(h, cudaSetDevice(h->device));
  if (num_row == 0) return 0;
  if (!row_ptr || !label || (row_ptr[3LL * num_row] > row_ptr[0] && (!index || !value)))
    return fail(h, "null input array");
  Geometry geo;
  if (pick_geometry(h, geo)) return 1;
  const bool exact = train && h->mode == SVDGPU_MODE_EXACT;
  // Ordered mode: chunks of basic-MF rows under plain L2 decay take the item-owner kernel, whose
  // plan (checks, tickets, queues) is built on the device; everything else takes k_exact, whose
  // tickets the host computes walking every row.  Hogwild leaves the row_ptr checks to the
  // kernels so that the host never touches the batch.
  const bool own_ok = exact && h->exact_owner && own_supported(h) && !sides_on(h);
  bool validated = false;
  if (exact && !own_ok) {
    if (validate_csr(h, num_row, row_ptr)) return 1;
    validated = true;
  }
  const int nchunk = (int)(((long long)num_row + h->chunk_rows - 1) / h->chunk_rows);
  // Host threads for the scan: what the option says, else this process's share of the cores (the
  // ranks of one box split them: torchrun's LOCAL_WORLD_SIZE) minus the calling thread, at most 16.
  // A thread verifies ~7 GB/s, i.e. 20 bytes per row against the 0.37 ns per row the bus saves:
  // with fewer than 5 threads the scan would be slower than copying everything, so it is left out.
  int scan_threads = h->scan_threads;
  if (scan_threads <= 0) {
    int hw = (int)std::thread::hardware_concurrency(), lw = 1;
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) lw = std::max(1, atoi(e));
    scan_threads = std::min(16, std::max(1, hw) / lw - 1);
    // (the ordered mode spends ~1 ns per row in its kernel: there even two threads keep ahead of it,
    // and what they save is host memory bandwidth the ranks of one box share)
    if (scan_threads < (exact ? 2 : 5)) scan_threads = 0;
  }
  // (The ordered mode: on request, or when four or more ranks share the host.  Alone its kernel, not the bus, sets
  // the pace -- 32 B per row cross in 0.6 ns, a row trains in 0.8 ns -- and measured, the chunks' plans interleave
  // better with the copies taking their time: 1.01 G inst/s with everything copied against 0.82 G compact at N=1,
  // 2.01 against 1.63 G at N=2; eight ranks on one host are short of host bandwidth: 5.25 against 5.68 G.)
  int local_world = 1;
  if (const char *e = getenv("LOCAL_WORLD_SIZE")) local_world = std::max(1, atoi(e));
  const bool compact = (exact ? own_ok && (h->compact_h2d >= 2 || (h->compact_h2d == 1 && local_world >= 4)) : h->compact_h2d >= 1) &&
                       scan_threads > 0 && !sides_on(h) && num_row >= h->compact_min_rows;
  if (h->hog_safety_permille <= 0) h->inflight_cap = 0;
  svdscan::ScanPool pool(compact ? nchunk : 0);  // (joined on every way out of this function)
  if (compact) pool.start(num_row, h->chunk_rows, row_ptr, value, scan_threads);
  for (int r0 = 0, ci = 0; r0 < num_row; r0 += h->chunk_rows, ++ci) {
    const int r1 = (int)std::min<long long>(num_row, (long long)r0 + h->chunk_rows);
    const int n = r1 - r0;
    // the chunk's arrays: the caller's, or their side-feature expansion (0-based)
    Expanded ex;
    const bool side = sides_on(h);
    if (side && expand_rows(h, r0, r1, row_ptr, index, value, ex)) return 1;
    const int *c_rp = side ? ex.rp.data() : row_ptr + 3LL * r0;
    const int v0 = c_rp[0], v1 = c_rp[3LL * n];
    if (v0 < 0 || v1 < v0) return fail(h, "row_ptr must be non-decreasing");
    const size_t nv = (size_t)(v1 - v0);
    const unsigned *c_idx = side ? ex.idx.data() : index + v0;
    const float *c_val = side ? ex.val.data() : value + v0;
    svdscan::ChunkScan sc;
    if (compact) sc = pool.wait(ci);
    Slot &s = next_slot(h);  // (the kernels that last read this slot are done)
    // ordered mode through k_own: the chunk's fills and plan run on the plan stream, beside the
    // k_own launch of the chunk before (which leaves them own_spare_sms SMs)
    const bool try_own = exact && own_ok && n >= h->own_min_rows;
    cudaStream_t aux = try_own ? h->plan_stream : h->stream;
    if (sc.rp_regular) {
      // not copied: rebuilt 0-based on the launch stream unless the slot still holds this very fill
      if (dev_reserve(h, s.d_rp, (3 * (size_t)n + 1) * 4)) return 1;
      if (s.d_rp.fill_n != n || s.d_rp.fill_a != sc.a || s.d_rp.fill_b != sc.b || s.d_rp.fill_c != sc.c) {
        k_fill_row_ptr<<<std::min(h->num_sm * 8, (n + 256) / 256), 256, 0, aux>>>((int *)s.d_rp.p, n, sc.a, sc.b, sc.c);
        CU(h, cudaGetLastError());
        h->n_launch++;
        s.d_rp.fill_n = n;
        s.d_rp.fill_a = sc.a;
        s.d_rp.fill_b = sc.b;
        s.d_rp.fill_c = sc.c;
      }
    } else if (h2d(h, s.d_rp, s.h_rp, c_rp, (3 * (size_t)n + 1) * 4)) {
      return 1;
    }
    if (h2d(h, s.d_label, s.h_label, label + r0, (size_t)n * 4)) return 1;
    if (h2d(h, s.d_index, s.h_index, c_idx, nv * 4)) return 1;
    if (sc.val_ones) {
      if (dev_reserve(h, s.d_value, nv * 4)) return 1;
      if (s.d_value.fill_n < (long long)nv) {
        k_fill_ones<<<(int)std::min<long long>(h->num_sm * 8, ((long long)nv + 255) / 256), 256, 0, aux>>>(
            (float *)s.d_value.p, (long long)nv);
        CU(h, cudaGetLastError());
        h->n_launch++;
        s.d_value.fill_n = (long long)nv;
      }
    } else if (h2d(h, s.d_value, s.h_value, c_val, nv * 4)) {
      return 1;
    }
    if (side && h2d(h, s.d_value2, s.h_value2, ex.val2.data(), nv * 4)) return 1;
    DevCsr csr;
    csr.row_ptr = (const int *)s.d_rp.p;
    csr.label = (const float *)s.d_label.p;
    csr.index = (const unsigned *)s.d_index.p;
    csr.value = (const float *)s.d_value.p;
    csr.value2 = side ? (const float *)s.d_value2.p : nullptr;
    csr.ticket = nullptr;
    csr.val_base = sc.rp_regular ? 0 : v0;  // a rebuilt row_ptr counts from 0
    csr.val_end = sc.rp_regular ? (int)nv : v1;
    bool owned = false;
    if (try_own) {
      CU(h, cudaEventRecord(s.copied, h->copy_stream));
      int bad = 0;
      const int ctas = nchunk > 1 ? std::max(1, h->num_sm - h->own_spare_sms) : h->num_sm;
      if (plan_on_side(h, csr, n, s.own, &bad, ctas, s.copied, s.planned)) return 1;
      if (s.own.valid) {
        if (launch_own(h, s.own, h->stream)) return 1;
        owned = true;
      } else if (!(bad & 3)) {  // every row has the basic shape, but an index is out of range (base.h:327,343)
        return fail(h, (bad & 4) ? "user feature index exceed bound" : "item feature index exceed bound");
      }
    }
    if (exact && !owned) {
      if (!validated) {
        if (validate_csr(h, num_row, row_ptr)) return 1;
        validated = true;
      }
      if (host_reserve(h, s.h_ticket, nv * 4 + 4)) return 1;
      reset_ticket_counters(h);
      if (side ? make_tickets(h, 0, n, ex.rp.data(), ex.idx.data(), (unsigned *)s.h_ticket.p)
               : make_tickets(h, r0, r1, row_ptr, index, (unsigned *)s.h_ticket.p))
        return 1;
      if (dev_reserve(h, s.d_ticket, nv * 4)) return 1;
      if (h2d_pinned(h, s.d_ticket.p, s.h_ticket.p, nv * 4)) return 1;
      csr.ticket = (const unsigned *)s.d_ticket.p;
      if (slot_copied(h, s)) return 1;
      // row_ptr on the device is the slice [3*r0, 3*r1]: rows are 0..n there
      if (launch_exact(h, geo, csr, 0, n)) return 1;
    } else if (!exact) {
      float *pred = nullptr;
      if (!train) {
        if (dev_reserve(h, s.d_pred, (size_t)n * 4)) return 1;
        pred = (float *)s.d_pred.p;
      }
      if (slot_copied(h, s)) return 1;
      if (train && ci == 0 && h->hog_safety_permille > 0) {  // Hogwild guard: the first chunk speaks for the call
        if (hot_for_call(h, csr, n)) return 1;
        set_inflight_cap(h, h->hot_call);
      }
      if (launch_stream(h, geo, csr, 0, n, train, pred)) return 1;
      if (!train && eval_chunk(h, pred, csr.label, n)) return 1;
      if (!train && out) {
        CU(h, cudaMemcpyAsync(out + r0, pred, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
        h->n_d2h += (long long)n * 4;
      }
    }
    if (slot_done(h, s)) return 1;
    h->n_inst += n;
  }
  return copies_drained(h);
}

int svdgpu_update_csr(svdgpu_t *h, int num_row, const int *row_ptr, const float *label,
                      const unsigned *index, const float *value) {
  return run_csr_host(h, num_row, row_ptr, label, index, value, nullptr, true);
}

int svdgpu_predict_csr(svdgpu_t *h, int num_row, const int *row_ptr, const float *label,
                       const unsigned *index, const float *value, float *out) {
  if (h && num_row > 0 && !out) return fail(h, "predict: null output");
  if (run_csr_host(h, num_row, row_ptr, label, index, value, out, false)) return 1;
  return svdgpu_sync(h);
}

// ---------------------------------------------------------------------------
// user-grouped input, host buffers
// ---------------------------------------------------------------------------
static int fb_tickets(svdgpu *h, const std::vector<int> &unit_off, int u0, int u1, const int *blk_row_off,
                      const int *blk_fb_off, const unsigned *fb_index, const int *row_ptr,
                      const unsigned *index, unsigned *fb_ticket, int fb_base, unsigned *ticket,
                      int row0, int row_shift = 0 /* row_ptr/index hold rows (absolute - row_shift) */) {
  // sequential order: unit gather (holds its feedback rows) -> rows -> scatter
  unsigned *cf = h->cnt_ui.data();  // feedback rows are rows [0,num_ufeedback) of the slab
  for (int u = u0; u < u1; ++u) {
    const int b0 = unit_off[u], b1 = unit_off[u + 1];
    const int f0 = blk_fb_off[b0], f1 = blk_fb_off[b0 + 1];
    // The unit holds the rows of its FIRST block's list from gather to scatter; the scatter goes to
    // the LAST block's list (base.h:539-554).  The reference's loader repeats a split user's list in
    // every piece; a unit whose END list differs would scatter to rows it does not hold.
    if (b1 - 1 != b0) {
      const int g0 = blk_fb_off[b1 - 1], g1 = blk_fb_off[b1];
      if (g1 - g0 != f1 - f0 || (f1 > f0 && memcmp(fb_index + g0, fb_index + f0, (size_t)(f1 - f0) * sizeof(unsigned)) != 0))
        return fail(h, "ordered mode: the END block of a split user must carry the START block's feedback list");
    }
    for (int f = f0; f < f1; ++f) {
      if (fb_index[f] >= (unsigned)h->dm.num_ufeedback) return fail(h, "ufeedback id exceed bound");
      fb_ticket[f - fb_base] = cf[fb_index[f]];
    }
    if (make_tickets(h, blk_row_off[b0] - row_shift, blk_row_off[b1] - row_shift, row_ptr, index,
                     ticket + (row_ptr[3LL * (blk_row_off[b0] - row_shift)] - row_ptr[3LL * row0])))
      return 1;
    for (int f = f0; f < f1; ++f) cf[fb_index[f]]++;
  }
  return 0;
}

static int run_ugroup_host(svdgpu_t *h, int num_block, const int *blk_row_off, const int *blk_fb_off,
                           const int *blk_tag, const unsigned *fb_index, const float *fb_value,
                           const int *row_ptr, const float *label, const unsigned *index,
                           const float *value, float *out, bool train) {
  if (check_ready(h)) return 1;
  CU(h, cudaSetDevice(h->device));
  if (h->shape.format_type != 1) return fail(h, "user-grouped input needs format_type=1 (trainer was created for random order)");
  if (num_block == 0) return 0;
  if (!blk_row_off || !blk_fb_off || !row_ptr || !label) return fail(h, "null input array");
  const int num_row = blk_row_off[num_block];
  if (blk_row_off[0] != 0 || blk_fb_off[0] != 0) return fail(h, "block offsets must start at 0");
  for (int b = 0; b < num_block; ++b)
    if (blk_row_off[b + 1] < blk_row_off[b] || blk_fb_off[b + 1] < blk_fb_off[b])
      return fail(h, "block offsets must be non-decreasing");
  if (validate_csr(h, num_row, row_ptr)) return 1;
  std::vector<int> unit_off;
  if (build_units(h, num_block, blk_tag, unit_off)) return 1;
  const int num_unit = (int)unit_off.size() - 1;
  Geometry geo;
  if (pick_geometry(h, geo)) return 1;
  const bool exact = train && h->mode == SVDGPU_MODE_EXACT;

  int u0 = 0;
  while (u0 < num_unit) {
    // chunk = as many whole units as fit in chunk_rows rows (at least one)
    int u1 = u0 + 1;
    while (u1 < num_unit && blk_row_off[unit_off[u1 + 1]] - blk_row_off[unit_off[u0]] <= h->chunk_rows) ++u1;
    const int b0 = unit_off[u0], b1 = unit_off[u1];
    const int r0 = blk_row_off[b0], r1 = blk_row_off[b1];
    const int n = r1 - r0;
    Expanded ex;
    const bool side = sides_on(h);
    if (side && expand_rows(h, r0, r1, row_ptr, index, value, ex)) return 1;
    const int *c_rp = side ? ex.rp.data() : row_ptr + 3LL * r0;
    const int v0 = c_rp[0], v1 = c_rp[3LL * n];
    const size_t nv = (size_t)(v1 - v0);
    const int fb0 = blk_fb_off[b0], fb1 = blk_fb_off[b1];
    const size_t nfb = (size_t)(fb1 - fb0);
    Slot &s = next_slot(h);
    if (h2d(h, s.d_rp, s.h_rp, c_rp, (3 * (size_t)n + 1) * 4)) return 1;
    if (h2d(h, s.d_label, s.h_label, label + r0, (size_t)n * 4)) return 1;
    if (h2d(h, s.d_index, s.h_index, side ? ex.idx.data() : index + v0, nv * 4)) return 1;
    if (h2d(h, s.d_value, s.h_value, side ? ex.val.data() : value + v0, nv * 4)) return 1;
    if (side && h2d(h, s.d_value2, s.h_value2, ex.val2.data(), nv * 4)) return 1;
    if (h2d(h, s.d_fbi, s.h_fbi, fb_index + fb0, nfb * 4)) return 1;
    if (h2d(h, s.d_fbv, s.h_fbv, fb_value + fb0, nfb * 4)) return 1;
    // misc = unit_off (relative blocks) | blk_row_off | blk_fb_off | order
    const int nu = u1 - u0, nb = b1 - b0;
    const size_t misc_ints = (size_t)(nu + 1) + 2 * (size_t)(nb + 1) + (size_t)nu;
    if (host_reserve(h, s.h_misc, misc_ints * 4)) return 1;
    int *mi = (int *)s.h_misc.p;
    int *m_unit = mi, *m_row = mi + (nu + 1), *m_fb = m_row + (nb + 1), *m_order = m_fb + (nb + 1);
    for (int u = 0; u <= nu; ++u) m_unit[u] = unit_off[u0 + u] - b0;
    for (int b = 0; b <= nb; ++b) {
      m_row[b] = blk_row_off[b0 + b];
      m_fb[b] = blk_fb_off[b0 + b];
    }
    unit_order(nu, [&](int u) { return (long long)(m_row[m_unit[u + 1]] - m_row[m_unit[u]]); }, m_order);
    if (dev_reserve(h, s.d_misc, misc_ints * 4)) return 1;
    if (h2d_pinned(h, s.d_misc.p, mi, misc_ints * 4)) return 1;

    DevCsr csr;
    csr.row_ptr = (const int *)s.d_rp.p;
    csr.label = (const float *)s.d_label.p;
    csr.index = (const unsigned *)s.d_index.p;
    csr.value = (const float *)s.d_value.p;
    csr.value2 = side ? (const float *)s.d_value2.p : nullptr;
    csr.ticket = nullptr;
    csr.val_base = v0;
    csr.val_end = v1;
    DevUgroup ug;
    const int *dmi = (const int *)s.d_misc.p;
    ug.unit_off = dmi;
    ug.blk_row_off = dmi + (nu + 1);
    ug.blk_fb_off = ug.blk_row_off + (nb + 1);
    ug.order = ug.blk_fb_off + (nb + 1);
    ug.fb_index = (const unsigned *)s.d_fbi.p;
    ug.fb_value = (const float *)s.d_fbv.p;
    ug.fb_ticket = nullptr;
    ug.row_base = r0;
    ug.fb_base = fb0;
    ug.has_fb = nfb > 0;
    if (exact) {
      if (host_reserve(h, s.h_ticket, nv * 4 + 4)) return 1;
      if (host_reserve(h, s.h_fbt, nfb * 4 + 4)) return 1;
      reset_ticket_counters(h);
      if (fb_tickets(h, unit_off, u0, u1, blk_row_off, blk_fb_off, fb_index, side ? ex.rp.data() : row_ptr,
                     side ? ex.idx.data() : index, (unsigned *)s.h_fbt.p, fb0, (unsigned *)s.h_ticket.p,
                     side ? 0 : r0, side ? r0 : 0))
        return 1;
      if (dev_reserve(h, s.d_ticket, nv * 4)) return 1;
      if (dev_reserve(h, s.d_fbt, nfb * 4)) return 1;
      if (h2d_pinned(h, s.d_ticket.p, s.h_ticket.p, nv * 4)) return 1;
      if (h2d_pinned(h, s.d_fbt.p, s.h_fbt.p, nfb * 4)) return 1;
      csr.ticket = (const unsigned *)s.d_ticket.p;
      ug.fb_ticket = (const unsigned *)s.d_fbt.p;
    }
    float *pred = nullptr;
    if (!train) {
      if (dev_reserve(h, s.d_pred, (size_t)n * 4)) return 1;
      pred = (float *)s.d_pred.p;
    }
    if (slot_copied(h, s)) return 1;
    // Units without feedback entries (pairwise-rank blocks, apex_svd_data.cpp:1001-1018) are plain
    // rows: tmp_ufeedback stays zero (base.h:506-520 with norm 0), so outside the ordered mode they
    // take the stream kernels and their fast passes.
    if (!exact && nfb == 0) {
      if (launch_stream(h, geo, csr, 0, n, train, pred)) return 1;
    } else if (launch_ugroup(h, geo, csr, ug, 0, nu, train, exact, pred)) {
      return 1;
    }
    if (!train && eval_chunk(h, pred, csr.label, n)) return 1;
    if (!train && n && out) {
      CU(h, cudaMemcpyAsync(out + r0, pred, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
      h->n_d2h += (long long)n * 4;
    }
    if (slot_done(h, s)) return 1;
    h->n_inst += n;
    u0 = u1;
  }
  return copies_drained(h);
}

int svdgpu_update_ugroup(svdgpu_t *h, int num_block, const int *blk_row_off, const int *blk_fb_off,
                         const int *blk_tag, const unsigned *fb_index, const float *fb_value,
                         const int *row_ptr, const float *label, const unsigned *index,
                         const float *value) {
  return run_ugroup_host(h, num_block, blk_row_off, blk_fb_off, blk_tag, fb_index, fb_value, row_ptr,
                         label, index, value, nullptr, true);
}
int svdgpu_predict_ugroup(svdgpu_t *h, int num_block, const int *blk_row_off, const int *blk_fb_off,
                          const int *blk_tag, const unsigned *fb_index, const float *fb_value,
                          const int *row_ptr, const float *label, const unsigned *index,
                          const float *value, float *out) {
  if (h && num_block > 0 && !out) return fail(h, "predict: null output");
  if (run_ugroup_host(h, num_block, blk_row_off, blk_fb_off, blk_tag, fb_index, fb_value, row_ptr,
                      label, index, value, out, false))
    return 1;
  return svdgpu_sync(h);
}

// ---------------------------------------------------------------------------
// evaluation on the device (svd_feature_infer.cpp:38-56, 243-277: task_eval)
// ---------------------------------------------------------------------------
static int eval_begin(svdgpu *h, float scale) {
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaMemsetAsync(h->d_eval, 0, 2 * sizeof(double), h->stream));
  h->eval_on = true;
  h->eval_scale = scale;
  return 0;
}
static int eval_end(svdgpu *h, int rc, double *sum_sq, long long *count) {
  h->eval_on = false;
  if (rc) return rc;
  double acc[2] = {0.0, 0.0};
  if (svdgpu_sync(h)) return 1;
  CU(h, cudaMemcpy(acc, h->d_eval, sizeof(acc), cudaMemcpyDeviceToHost));
  h->n_d2h += (long long)sizeof(acc);
  if (sum_sq) *sum_sq = acc[0];
  if (count) *count = (long long)(acc[1] + 0.5);
  return 0;
}

int svdgpu_eval_csr(svdgpu_t *h, int num_row, const int *row_ptr, const float *label, const unsigned *index,
                    const float *value, float scale, double *sum_sq, long long *count) {
  if (!h) return 1;
  if (eval_begin(h, scale)) return 1;
  return eval_end(h, run_csr_host(h, num_row, row_ptr, label, index, value, nullptr, false), sum_sq, count);
}

int svdgpu_eval_ugroup(svdgpu_t *h, int num_block, const int *blk_row_off, const int *blk_fb_off,
                       const int *blk_tag, const unsigned *fb_index, const float *fb_value, const int *row_ptr,
                       const float *label, const unsigned *index, const float *value, float scale,
                       double *sum_sq, long long *count) {
  if (!h) return 1;
  if (eval_begin(h, scale)) return 1;
  return eval_end(h,
                  run_ugroup_host(h, num_block, blk_row_off, blk_fb_off, blk_tag, fb_index, fb_value, row_ptr,
                                  label, index, value, nullptr, false),
                  sum_sq, count);
}

// ---------------------------------------------------------------------------
// resident batches
// ---------------------------------------------------------------------------
static DevCsr batch_csr(const svdgpu_batch *b);
static int upload_plain(svdgpu *h, DevBuf &d, const void *src, size_t bytes) {
  if (dev_reserve(h, d, bytes)) return 1;
  if (bytes) {
    CU(h, cudaMemcpyAsync(d.p, src, bytes, cudaMemcpyHostToDevice, h->stream));
    h->n_h2d += (long long)bytes;
  }
  return 0;
}

int svdgpu_batch_create(svdgpu_t *h, svdgpu_batch_t **out, int num_row, const int *row_ptr,
                        const float *label, const unsigned *index, const float *value) {
  if (!h || !out) return 1;
  *out = nullptr;
  CU(h, cudaSetDevice(h->device));
  if (num_row < 0 || !row_ptr || (num_row > 0 && !label)) return fail(h, "batch_create: bad arguments");
  if (validate_csr(h, num_row, row_ptr)) return 1;
  if (row_ptr[0] != 0) return fail(h, "batch_create: row_ptr[0] must be 0");
  Expanded ex;  // side features: the resident batch holds the expanded rows
  const bool side = sides_on(h);
  if (side) {
    if (expand_rows(h, 0, num_row, row_ptr, index, value, ex)) return 1;
    row_ptr = ex.rp.data();
    index = ex.idx.data();
    value = ex.val.data();
  }
  svdgpu_batch *b = new svdgpu_batch();
  b->num_row = num_row;
  b->num_val = row_ptr[3LL * num_row];
  const size_t nv = (size_t)b->num_val;
  int rc = 0;
  rc |= upload_plain(h, b->d_rp, row_ptr, (3 * (size_t)num_row + 1) * 4);
  rc |= upload_plain(h, b->d_label, label, (size_t)num_row * 4);
  rc |= upload_plain(h, b->d_index, index, nv * 4);
  rc |= upload_plain(h, b->d_value, value, nv * 4);
  if (side) rc |= upload_plain(h, b->d_value2, ex.val2.data(), nv * 4);
  b->has_value2 = side;
  if (!rc && h->mode == SVDGPU_MODE_EXACT && h->shape.format_type == 0) {
    // ordered mode: the owner plan when the rows qualify (built on the device), else the
    // per-feature tickets of k_exact (host)
    bool need_tickets = true;
    if (h->exact_owner && !side && num_row >= h->own_min_rows && h->hp_set && own_supported(h)) {
      int bad = 0;
      rc |= plan_on_side(h, batch_csr(b), num_row, b->own, &bad, h->num_sm, nullptr, h->ev_plan);
      if (!rc && b->own.valid) need_tickets = false;
      else if (!rc && !(bad & 3) && bad) {
        rc = fail(h, (bad & 4) ? "user feature index exceed bound" : "item feature index exceed bound");
      }
    }
    if (!rc && need_tickets) {
      std::vector<unsigned> tk(nv + 1);
      reset_ticket_counters(h);
      rc |= make_tickets(h, 0, num_row, row_ptr, index, tk.data());
      if (!rc) rc |= upload_plain(h, b->d_ticket, tk.data(), nv * 4);
      if (!rc) rc |= (cudaStreamSynchronize(h->stream) != cudaSuccess);
      b->has_ticket = !rc;
    }
  }
  if (!rc) rc |= (cudaStreamSynchronize(h->stream) != cudaSuccess);
  if (rc) {
    svdgpu_batch_destroy(h, b);
    return 1;
  }
  *out = b;
  return 0;
}

int svdgpu_batch_set_ugroup(svdgpu_t *h, svdgpu_batch_t *b, int num_block, const int *blk_row_off,
                            const int *blk_fb_off, const int *blk_tag, const unsigned *fb_index,
                            const float *fb_value) {
  if (!h || !b) return 1;
  CU(h, cudaSetDevice(h->device));
  if (h->shape.format_type != 1) return fail(h, "user-grouped input needs format_type=1");
  if (num_block < 0 || !blk_row_off || !blk_fb_off) return fail(h, "batch_set_ugroup: bad arguments");
  if (blk_row_off[0] != 0 || blk_fb_off[0] != 0 || blk_row_off[num_block] != b->num_row)
    return fail(h, "batch_set_ugroup: block offsets must cover the batch rows");
  if (build_units(h, num_block, blk_tag, b->unit_off)) return 1;
  b->num_block = num_block;
  b->num_unit = (int)b->unit_off.size() - 1;
  const size_t nfb = (size_t)blk_fb_off[num_block];
  std::vector<int> order((size_t)std::max(b->num_unit, 1));
  unit_order(b->num_unit, [&](int u) {
    return (long long)(blk_row_off[b->unit_off[u + 1]] - blk_row_off[b->unit_off[u]]);
  }, order.data());
  int rc = 0;
  rc |= upload_plain(h, b->d_unit_off, b->unit_off.data(), b->unit_off.size() * 4);
  rc |= upload_plain(h, b->d_blk_row_off, blk_row_off, ((size_t)num_block + 1) * 4);
  rc |= upload_plain(h, b->d_blk_fb_off, blk_fb_off, ((size_t)num_block + 1) * 4);
  b->has_fb = nfb > 0;
  rc |= upload_plain(h, b->d_fbi, fb_index, nfb * 4);
  rc |= upload_plain(h, b->d_fbv, fb_value, nfb * 4);
  rc |= upload_plain(h, b->d_order, order.data(), order.size() * 4);
  if (!rc) rc |= (cudaStreamSynchronize(h->stream) != cudaSuccess);
  if (rc) return fail(h, "batch_set_ugroup: upload failed");
  b->ugroup = true;
  return 0;
}

static DevCsr batch_csr(const svdgpu_batch *b) {
  DevCsr c;
  c.row_ptr = (const int *)b->d_rp.p;
  c.label = (const float *)b->d_label.p;
  c.index = (const unsigned *)b->d_index.p;
  c.value = (const float *)b->d_value.p;
  c.value2 = b->has_value2 ? (const float *)b->d_value2.p : nullptr;
  c.ticket = b->has_ticket ? (const unsigned *)b->d_ticket.p : nullptr;
  c.val_base = 0;
  c.val_end = (int)b->num_val;
  c.avg_nnz = b->num_row > 0 ? (float)((double)b->num_val / (double)b->num_row) : 0.0f;
  return c;
}
static DevUgroup batch_ug(const svdgpu_batch *b) {
  DevUgroup u;
  u.unit_off = (const int *)b->d_unit_off.p;
  u.blk_row_off = (const int *)b->d_blk_row_off.p;
  u.blk_fb_off = (const int *)b->d_blk_fb_off.p;
  u.fb_index = (const unsigned *)b->d_fbi.p;
  u.fb_value = (const float *)b->d_fbv.p;
  u.fb_ticket = nullptr;
  u.order = (const int *)b->d_order.p;
  u.row_base = 0;
  u.fb_base = 0;
  u.has_fb = b->has_fb;
  return u;
}

int svdgpu_batch_update(svdgpu_t *h, svdgpu_batch_t *b, int begin, int end) {
  if (check_ready(h)) return 1;
  if (!b) return fail(h, "null batch");
  CU(h, cudaSetDevice(h->device));
  Geometry geo;
  if (pick_geometry(h, geo)) return 1;
  if (b->ugroup) {
    if (begin < 0 || end > b->num_unit || begin > end) return fail(h, "batch_update: unit range out of bounds");
    if (begin == end) return 0;
    if (h->mode == SVDGPU_MODE_EXACT)
      return fail(h, "ordered mode on a resident user-grouped batch is not supported; use svdgpu_update_ugroup");
    // the LPT order array covers the whole batch: partial ranges run in input order
    DevUgroup ug = batch_ug(b);
    if (begin != 0 || end != b->num_unit) ug.order = nullptr;
    if (!b->has_fb && begin == 0 && end == b->num_unit) {
      // no feedback entries anywhere: plain rows (see run_ugroup_host)
      if (b->num_row > 0 && launch_stream(h, geo, batch_csr(b), 0, b->num_row, true, (float *)nullptr)) return 1;
    } else if (launch_ugroup(h, geo, batch_csr(b), ug, begin, end, true, false, (float *)nullptr)) {
      return 1;
    }
    h->n_inst += b->num_row;
    return 0;
  }
  if (begin < 0 || end > b->num_row || begin > end) return fail(h, "batch_update: row range out of bounds");
  if (begin == end) return 0;
  if (h->mode == SVDGPU_MODE_EXACT) {
    if (begin != 0 || end != b->num_row)
      return fail(h, "ordered mode needs the whole resident batch (tickets are per batch)");
    if (h->exact_owner && b->own.valid && own_supported(h)) {
      if (launch_own(h, b->own, h->stream)) return 1;
    } else {
      if (!b->has_ticket)
        return fail(h, b->own.valid ? "batch was created for the item-owner kernel (plain L2 decay): re-create it after "
                                      "changing the regulariser or the exact_owner option"
                                    : "batch was created in hogwild mode: no tickets for the ordered mode");
      if (launch_exact(h, geo, batch_csr(b), begin, end)) return 1;
    }
  } else {
    if (h->hog_safety_permille > 0) {  // Hogwild guard: the batch's hottest item, measured once
      if (b->hot_frac < 0.0 && hot_measure(h, batch_csr(b), 0, b->num_row, true, &b->hot_frac)) return 1;
      set_inflight_cap(h, b->hot_frac);
    } else {
      h->inflight_cap = 0;
    }
    if (launch_stream(h, geo, batch_csr(b), begin, end, true, (float *)nullptr)) return 1;
  }
  h->n_inst += end - begin;
  return 0;
}

int svdgpu_batch_plan(svdgpu_t *h, svdgpu_batch_t *b, int *planned) {
  if (check_ready(h)) return 1;
  if (!b) return fail(h, "null batch");
  CU(h, cudaSetDevice(h->device));
  if (planned) *planned = 0;
  if (b->ugroup || b->has_value2 || !h->exact_owner || !own_supported(h) || b->num_row <= 0) return 0;
  int bad = 0;
  if (plan_on_side(h, batch_csr(b), b->num_row, b->own, &bad, h->num_sm, nullptr, h->ev_plan)) return 1;
  if (planned) *planned = b->own.valid ? 1 : 0;
  return 0;
}

int svdgpu_batch_predict(svdgpu_t *h, svdgpu_batch_t *b, int begin, int end, float *out_host) {
  if (check_ready(h)) return 1;
  if (!b) return fail(h, "null batch");
  CU(h, cudaSetDevice(h->device));
  Geometry geo;
  if (pick_geometry(h, geo)) return 1;
  if (dev_reserve(h, b->d_pred, (size_t)std::max(b->num_row, 1) * 4)) return 1;
  float *pred = (float *)b->d_pred.p;
  int r0 = begin, r1 = end;
  if (b->ugroup) {
    if (begin < 0 || end > b->num_unit || begin > end) return fail(h, "batch_predict: unit range out of bounds");
    if (begin == end) return 0;
    DevUgroup ug = batch_ug(b);
    ug.order = nullptr;
    if (launch_ugroup(h, geo, batch_csr(b), ug, begin, end, false, false, pred)) return 1;
    // rows covered by the unit range (host copy of the offsets is not kept: copy all rows)
    r0 = 0;
    r1 = b->num_row;
  } else {
    if (begin < 0 || end > b->num_row || begin > end) return fail(h, "batch_predict: row range out of bounds");
    if (begin == end) return 0;
    if (launch_stream(h, geo, batch_csr(b), begin, end, false, pred + begin)) return 1;
  }
  if (eval_chunk(h, pred + r0, (const float *)b->d_label.p + r0, r1 - r0)) return 1;
  if (out_host) {
    CU(h, cudaMemcpyAsync(out_host, pred + r0, (size_t)(r1 - r0) * 4, cudaMemcpyDeviceToHost, h->stream));
    h->n_d2h += (long long)(r1 - r0) * 4;
    return svdgpu_sync(h);
  }
  return 0;
}

int svdgpu_batch_eval(svdgpu_t *h, svdgpu_batch_t *b, int begin, int end, float scale, double *sum_sq,
                      long long *count) {
  if (!h) return 1;
  if (eval_begin(h, scale)) return 1;
  return eval_end(h, svdgpu_batch_predict(h, b, begin, end, nullptr), sum_sq, count);
}

void svdgpu_batch_destroy(svdgpu_t *h, svdgpu_batch_t *b) {
  if (!b) return;
  if (h) {
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
  }
  DevBuf *db[] = {&b->d_rp, &b->d_label, &b->d_index, &b->d_value, &b->d_value2, &b->d_ticket, &b->d_pred,
                  &b->d_unit_off, &b->d_blk_row_off, &b->d_blk_fb_off, &b->d_fbi, &b->d_fbv, &b->d_fbt, &b->d_order};
  for (DevBuf *d : db) dev_free(*d);
  own_plan_free(b->own);
  delete b;
}

// ---------------------------------------------------------------------------
// multi-GPU exchange of the replicated slabs
// ---------------------------------------------------------------------------
static int ensure_plan(svdgpu *h) {
  if (h->d_snap) return 0;
  const DevModel &m = h->dm;
  DeltaPlan &p = h->plan;
  p.nseg = 0;
  long long off = 0;
  auto add = [&](float *cur, long long n) {
    if (n <= 0) return;
    p.seg[p.nseg].cur = cur;
    p.seg[p.nseg].n = n;
    p.seg[p.nseg].off = off;
    off += (n + 3) & ~3LL;
    p.nseg++;
  };
  add(m.W + (size_t)m.item_off * m.pitch, (long long)m.num_item * m.pitch);  // W_item
  add(m.bias + m.item_off, m.num_item);                                      // i_bias
  add(m.g_bias, m.num_global);                                               // g_bias
  add(m.W, (long long)m.num_ufeedback * m.pitch);                            // W_ufeedback
  add(m.bias, m.num_ufeedback);                                              // ufeedback_bias
  p.total = off;
  CU(h, cudaMalloc(&h->d_snap, std::max<long long>(off, 4) * sizeof(float)));
  CU(h, cudaMalloc(&h->d_delta, std::max<long long>(off, 4) * sizeof(float)));
  CU(h, cudaMemsetAsync(h->d_delta, 0, std::max<long long>(off, 4) * sizeof(float), h->stream));
  return 0;
}
static int run_delta(svdgpu *h, int mode, float scale) {
  if (ensure_plan(h)) return 1;
  return launch_delta(h, mode, scale);
}
int svdgpu_items_snapshot(svdgpu_t *h) {
  if (!h) return 1;
  CU(h, cudaSetDevice(h->device));
  return run_delta(h, 0, 0.f);
}
int svdgpu_items_pack_delta(svdgpu_t *h, void **dev_ptr, size_t *num_floats) {
  if (!h || !dev_ptr || !num_floats) return 1;
  CU(h, cudaSetDevice(h->device));
  if (!h->d_snap) return fail(h, "items_pack_delta: call svdgpu_items_snapshot first");
  if (run_delta(h, 1, 0.f)) return 1;
  *dev_ptr = h->d_delta;
  *num_floats = (size_t)h->plan.total;
  return 0;
}
int svdgpu_items_apply_delta(svdgpu_t *h, float scale) {
  if (!h) return 1;
  CU(h, cudaSetDevice(h->device));
  if (!h->d_snap) return fail(h, "items_apply_delta: call svdgpu_items_snapshot first");
  return run_delta(h, 2, scale);
}

}  // extern "C"
