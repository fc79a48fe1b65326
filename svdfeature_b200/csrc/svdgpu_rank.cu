// svdgpu_rank.cu -- SVDFeatureRanker on the device.
//
// The reference's ranker (base.h:597-813) consumes a TAGGED instance stream (svdranker_tag,
// apex_svd.h:115-152; the tag sits in the label field): ITEM rows define the candidate set
// (prepare_ifactor: tmp_ifactors[idx] = sum ival*W_item[iid], bias_ifactors[idx] = sum
// ival*i_bias[iid] + sum gval*g_bias[gid], base.h:687-717), then every user section
// USER / POS* / BAN* / SPEC* / PROCESS ranks the whole set for one user: score[i] = spec[i] +
// (bias_ifactors[i] + dot(tmp_ufactor, tmp_ifactors[i])) for every item that is not banned,
// sorted by score (base.h:759-782); the answer is the top_k item indices or, with top_k = 0, the
// rank position of every POS item.  On the CPU that is one k-length dot per (user, item) pair.
//
// Here the host only walks the stream (a state machine over tiny rows, with the reference's
// checks and messages); everything O(k) runs on the GPU, batched over all user sections that
// were closed by a PROCESS row in the call:
//
//   k_rank_items   one warp per new ITEM row: gather + axpy of its item rows -> the candidate
//                  matrix, stored TRANSPOSED in float4 chunks (IF[c][i]) so that k_rank_score's
//                  thread-per-item reads are coalesced; lane 0 folds the bias
//   k_rank_users   one warp per section: tmp_ufactor (SVD++: starts from sum val*W_ufeedback[fid])
//   k_rank_spec    one warp per SPEC row: its own item vector, dot with the section's user vector
//   k_rank_mark    BAN entries -> a sentinel in the score matrix
//   k_rank_score   grid (items/128, sections): thread = one (section, item) pair, user vector in
//                  shared memory; the dot is evaluated in the reference's order (four lane-strided
//                  partial sums folded (l0+l2)+(l1+l3), then the k%4 tail, sse.h:88-97,289-317),
//                  so scores -- and therefore ranks -- are bit-identical to the CPU's
//   k_rank_pos     top_k = 0: one block per POS entry counts the candidates ranked before it
//   k_rank_select  0 < top_k <= 32: one block per section picks the best candidates one by one
//   k_rank_keys + cub::DeviceRadixSort   larger top_k: one stable sort of (section, ~score) keys
//
// Equal scores: the reference's std::sort leaves their order unspecified; here (and in the
// oracle) the lower item index comes first.  -0 and +0 compare equal, as they do for operator<.
// Arithmetic is fp32 with separate multiply and add (--fmad=false), like everything else.
#include "svdgpu_internal.h"

#include <cub/device/device_radix_sort.cuh>

#include <cstring>

using namespace svdk;

namespace {

enum { TAG_ITEM = 0, TAG_POS = 1, TAG_USER = 2, TAG_SPEC = 3, TAG_PROCESS = 4, TAG_BAN = -1 };
constexpr unsigned BANNED_BITS = 0x7fc0dead;  // a quiet NaN no arithmetic produces

// one expanded feature entry: factor scale sf, bias factors b1, b2 (bias += (b*b1)*b2).
// Plain item entry: sf = ival, b1 = ival, b2 = 1.  Side-feature entry of item feature (iid, ival)
// (base.h:698-702): sf = float(double(v)*double(ival)) (the two scalars fold in double,
// apex_exp_template.h:500-503), b1 = v, b2 = ival.
struct Entry {
  unsigned idx;
  float sf, b1, b2;
};
// a feature row: entries [e0, e1) of the item space, entries [g0, g1) of (gid, gval) pairs
struct FRow {
  int e0, e1, g0, g1;
};
struct GEntry {
  unsigned gid;
  float gval;
};
struct Sec {
  int u0, u1;    // user entries (idx, sf)
  int f0, f1;    // feedback entries (idx, sf)
  int n_cand;    // items in the set when PROCESS arrived
};
struct Spec {
  int sec, item;
  FRow row;
};
struct Mark {
  int sec, item;
};

template <typename T>
struct DVec {  // a device array that only grows
  T *p = nullptr;
  size_t cap = 0;
  int reserve(svdgpu *h, size_t n) {
    if (n <= cap) return 0;
    if (p) CU(h, cudaFree(p));
    p = nullptr;
    cap = 0;
    const size_t want = n + n / 4 + 16;
    CU(h, cudaMalloc(&p, want * sizeof(T)));
    cap = want;
    return 0;
  }
  int upload(svdgpu *h, const std::vector<T> &v) {
    if (reserve(h, v.size())) return 1;
    if (!v.empty())
      CU(h, cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    h->n_h2d += (long long)(v.size() * sizeof(T));
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace

struct svdgpu_rank_state {
  int cap_items = 0, n_items = 0, n_items_dev = 0, top_k = 0;
  // candidate set on the device: IF[c * cap_items + i] (float4 chunk c of item i), ibias[i]
  DVec<float4> d_IF;
  DVec<float> d_ibias;
  // ITEM rows not yet on the device
  std::vector<FRow> it_rows;
  std::vector<Entry> it_ent;
  std::vector<GEntry> it_g;
  // stream state
  bool open = false;  // inside a user section (USER seen, PROCESS not yet)
  std::vector<unsigned> cur_fbi;
  std::vector<float> cur_fbv;
  std::vector<int> stamp;  // per item: id of the last section that tagged it
  int sec_id = 0;
  // closed (and the open) sections
  std::vector<Sec> secs;
  std::vector<Entry> u_ent, f_ent;  // only idx, sf used
  std::vector<Mark> pos, ban;
  std::vector<Spec> spec;
  std::vector<Entry> sp_ent;
  std::vector<GEntry> sp_g;
  int n_closed = 0;
  // device scratch
  DVec<FRow> d_rows;
  DVec<Entry> d_ent, d_uent, d_fent, d_spent;
  DVec<GEntry> d_g, d_spg;
  DVec<Sec> d_secs;
  DVec<Mark> d_pos, d_ban;
  DVec<Spec> d_spec;
  DVec<float> d_U, d_score;
  DVec<unsigned long long> d_key, d_key2;
  DVec<int> d_val, d_val2, d_out;
  DVec<unsigned char> d_tmp;
  std::vector<int> h_out;

  void release() {
    d_IF.release(); d_ibias.release(); d_rows.release(); d_ent.release(); d_uent.release(); d_fent.release();
    d_spent.release(); d_g.release(); d_spg.release(); d_secs.release(); d_pos.release(); d_ban.release();
    d_spec.release(); d_U.release(); d_score.release(); d_key.release(); d_key2.release(); d_val.release();
    d_val2.release(); d_out.release(); d_tmp.release();
  }
};

namespace {

// ---- device code -----------------------------------------------------------------------------
// the reference's dot order over two k-vectors held as float4 chunks; b chunks come from a
// strided array (stride in float4)
template <typename LoadB>
__device__ __forceinline__ float ref_dot(const float4 *a, LoadB loadb, int k) {
  float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
  const int full = k >> 2;
  for (int c = 0; c < full; ++c) {
    const float4 x = a[c], y = loadb(c);
    l0 = __fadd_rn(l0, __fmul_rn(x.x, y.x));
    l1 = __fadd_rn(l1, __fmul_rn(x.y, y.y));
    l2 = __fadd_rn(l2, __fmul_rn(x.z, y.z));
    l3 = __fadd_rn(l3, __fmul_rn(x.w, y.w));
  }
  float sum = __fadd_rn(__fadd_rn(l0, l2), __fadd_rn(l1, l3));
  const int tail = k & 3;
  if (tail) {
    const float4 x = a[full], y = loadb(full);
    sum = __fadd_rn(sum, __fmul_rn(x.x, y.x));
    if (tail > 1) sum = __fadd_rn(sum, __fmul_rn(x.y, y.y));
    if (tail > 2) sum = __fadd_rn(sum, __fmul_rn(x.z, y.z));
  }
  return sum;
}

// chunk c of sum_e sf_e * W[row_off + idx_e]  (prepare_ifactor / proc_user accumulate, in entry order)
__device__ __forceinline__ float4 gather_chunk(const DevModel &m, int row_off, const Entry *ent, int e0, int e1,
                                               int c, float4 acc) {
  for (int e = e0; e < e1; ++e) {
    const float4 w = ldcg4(m.W + ((size_t)row_off + ent[e].idx) * m.pitch + 4 * c);
    const float s = ent[e].sf;
    acc = f4_add_scaled(acc, w, s, scalar_is_one(s));
  }
  return acc;
}

// bias of a feature row (base.h:691-709): item entries first, then globals
__device__ __forceinline__ float row_bias(const DevModel &m, const Entry *ent, const GEntry *g, const FRow &r) {
  float bias = 0.f;
  for (int e = r.e0; e < r.e1; ++e) {
    float p = __fmul_rn(__ldcg(m.bias + m.item_off + ent[e].idx), ent[e].b1);
    if (ent[e].b2 != 1.0f) p = __fmul_rn(p, ent[e].b2);  // b2 == 1 marks a plain entry: no second factor
    bias = __fadd_rn(bias, p);
  }
  for (int j = r.g0; j < r.g1; ++j) bias = __fadd_rn(bias, __fmul_rn(g[j].gval, __ldcg(m.g_bias + g[j].gid)));
  return bias;
}

__global__ void k_rank_items(DevModel m, const FRow *rows, const Entry *ent, const GEntry *g, int n_new, int first,
                             float4 *IF, float *ibias, int cap) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_new) return;
  const FRow r = rows[warp];
  const int chunks = m.pitch >> 2;
  for (int c = lane; c < chunks; c += 32)
    IF[(size_t)c * cap + first + warp] = gather_chunk(m, m.item_off, ent, r.e0, r.e1, c, f4_zero());
  if (lane == 0) ibias[first + warp] = row_bias(m, ent, g, r);
}

__global__ void k_rank_users(DevModel m, const Sec *secs, const Entry *uent, const Entry *fent, int n_sec,
                             float4 *U) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_sec) return;
  const Sec s = secs[warp];
  const int chunks = m.pitch >> 2;
  for (int c = lane; c < chunks; c += 32) {
    float4 acc = gather_chunk(m, 0, fent, s.f0, s.f1, c, f4_zero());  // tmp_ufeedback (base.h:800-808)
    acc = gather_chunk(m, m.user_off, uent, s.u0, s.u1, c, acc);      // proc_user (base.h:719-735)
    U[(size_t)warp * chunks + c] = acc;
  }
}

// SPEC rows (base.h:750-758): item_score[idx] = bias + dot(tmp_ufactor, ifactor of this row)
__global__ void k_rank_spec(DevModel m, const Spec *spec, const Entry *ent, const GEntry *g, int n_spec,
                            const float4 *U, float *score, int width) {
  extern __shared__ float4 sh[];  // one ifactor per warp
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + wib;
  const int chunks = m.pitch >> 2;
  float4 *mine = sh + (size_t)wib * chunks;
  if (warp < n_spec) {
    const Spec sp = spec[warp];
    for (int c = lane; c < chunks; c += 32) mine[c] = gather_chunk(m, m.item_off, ent, sp.row.e0, sp.row.e1, c, f4_zero());
    __syncwarp();
    if (lane == 0) {
      const float4 *u = U + (size_t)sp.sec * chunks;
      const float d = ref_dot(u, [&](int c) { return mine[c]; }, m.k);
      score[(size_t)sp.sec * width + sp.item] = __fadd_rn(row_bias(m, ent, g, sp.row), d);
    }
  }
}

__global__ void k_rank_mark(const Mark *ban, int n_ban, float *score, int width) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_ban) score[(size_t)ban[i].sec * width + ban[i].item] = __uint_as_float(BANNED_BITS);
}

__device__ __forceinline__ bool is_banned(float s) { return __float_as_uint(s) == BANNED_BITS; }

// proc_rank's scoring loop (base.h:761-765) for every (section, item) pair
__global__ void k_rank_score(DevModel m, const Sec *secs, const float4 *U, const float4 *IF, const float *ibias,
                             int cap, float *score, int width) {
  extern __shared__ float4 su[];
  const int sec = blockIdx.y, chunks = m.pitch >> 2;
  for (int c = threadIdx.x; c < chunks; c += blockDim.x) su[c] = U[(size_t)sec * chunks + c];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= width) return;
  float *out = score + (size_t)sec * width + i;
  if (i >= secs[sec].n_cand) {
    *out = __uint_as_float(BANNED_BITS);
    return;
  }
  const float base = *out;
  if (is_banned(base)) return;
  const float d = ref_dot(su, [&](int c) { return IF[(size_t)c * cap + i]; }, m.k);
  *out = __fadd_rn(base, __fadd_rn(ibias[i], d));
}

// rank position of one POS item = number of candidates the sort puts before it (base.h:774-780)
__global__ void k_rank_pos(const Mark *pos, const float *score, int width, int *out) {
  __shared__ int part[32];
  const Mark p = pos[blockIdx.x];
  const float *row = score + (size_t)p.sec * width;
  const float mine = row[p.item];
  int cnt = 0;
  for (int i = threadIdx.x; i < width; i += blockDim.x) {
    const float s = row[i];
    if (is_banned(s)) continue;
    cnt += (s > mine || (s == mine && i < p.item)) ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += part[w];
    out[blockIdx.x] = t;
  }
}

// top_k <= RANK_SELECT_MAX: one block per section picks its best candidates one after the other.
// Pass t takes the best candidate that comes after pass t-1's pick in the order (score descending,
// position ascending), so nothing is marked and the candidate row is only read (L2-resident).
constexpr int RANK_SELECT_MAX = 32;
__global__ void k_rank_select(const float *score, int width, int top_k, int *out) {
  __shared__ float ws[32];
  __shared__ int wi[32];
  __shared__ float prev_s;
  __shared__ int prev_i;
  const float *row = score + (size_t)blockIdx.x * width;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int t = 0; t < top_k; ++t) {
    const bool first = t == 0;
    const float ps = first ? 0.0f : prev_s;
    const int pi = first ? -1 : prev_i;
    float bs = 0.0f;
    int bi = -1;  // -1: nothing yet
    for (int i = threadIdx.x; i < width; i += blockDim.x) {
      const float sc = row[i];
      if (is_banned(sc)) continue;
      if (!first && !(sc < ps || (sc == ps && i > pi))) continue;
      if (bi < 0 || sc > bs) {  // (i ascends within a thread: an equal score never replaces an earlier one)
        bs = sc;
        bi = i;
      }
    }
    auto better = [](float s1, int i1, float s2, int i2) {  // is (s1, i1) ranked before (s2, i2)?
      if (i1 < 0) return false;
      if (i2 < 0) return true;
      return s1 > s2 || (s1 == s2 && i1 < i2);
    };
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(os, oi, bs, bi)) {
        bs = os;
        bi = oi;
      }
    }
    if (lane == 0) {
      ws[warp] = bs;
      wi[warp] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < nwarp; ++w)
        if (better(ws[w], wi[w], bs, bi)) {
          bs = ws[w];
          bi = wi[w];
        }
      prev_s = bs;
      prev_i = bi;
      out[(size_t)blockIdx.x * top_k + t] = bi;
    }
    __syncthreads();
  }
}

// sort keys: section-major, score descending, banned last; equal scores keep the index order
__global__ void k_rank_keys(const float *score, long long n, int width, unsigned long long *key, int *val) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const float s = score[t];
  unsigned lo;
  if (is_banned(s)) {
    lo = 0xffffffffu;
  } else {
    const unsigned b = __float_as_uint(__fadd_rn(s, 0.0f));               // -0 -> +0
    const unsigned ord = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // ascending with the float order
    lo = ~ord;
  }
  key[t] = ((unsigned long long)(t / width) << 32) | lo;
  val[t] = (int)(t % width);
}

__global__ void k_rank_take(const int *val, int width, int top_k, int n_sec, int *out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_sec * top_k) out[t] = val[(size_t)(t / top_k) * width + t % top_k];
}

// ---- host: the stream state machine ---------------------------------------------------------
struct Row {
  int tag;
  int ng, nu, ni;
  const unsigned *gi, *ui, *ii;
  const float *gv, *uv, *iv;
};

int expand_item_row(svdgpu *h, const Row &r, std::vector<Entry> &ent, std::vector<GEntry> &g, FRow &out) {
  out.e0 = (int)ent.size();
  for (int i = 0; i < r.ni; ++i) {
    const unsigned iid = r.ii[i];
    const float ival = r.iv[i];
    if (!(iid < (unsigned)h->shape.num_item)) return fail(h, "item feature index exceed setting");
    ent.push_back(Entry{iid, ival, ival, 1.0f});
    const svdgpu::Side &sd = h->side_i;
    if (sd.on() && iid + 1 < sd.rp.size())
      for (unsigned j = sd.rp[iid]; j < sd.rp[iid + 1]; ++j) {
        // b2 == 1 means "plain entry" on the device; a side entry with ival == 1 multiplies by 1: same value
        ent.push_back(Entry{sd.idx[j], (float)((double)sd.val[j] * (double)ival), sd.val[j], ival});
      }
  }
  out.e1 = (int)ent.size();
  out.g0 = (int)g.size();
  for (int i = 0; i < r.ng; ++i) {
    if (!(r.gi[i] < (unsigned)h->shape.num_global)) return fail(h, "global feature index exceed setting");
    g.push_back(GEntry{r.gi[i], r.gv[i]});
  }
  out.g1 = (int)g.size();
  return 0;
}

int feed_row(svdgpu *h, const Row &r) {
  svdgpu_rank_state &s = *h->rank;
  switch (r.tag) {
    case TAG_ITEM: {  // proc_item, base.h:712-717
      if (s.open) return fail(h, "ranker: ITEM rows inside a user section are not supported");
      if (!(s.n_items + 1 <= s.cap_items)) return fail(h, "item instance exceed specified item set size");
      FRow fr;
      if (expand_item_row(h, r, s.it_ent, s.it_g, fr)) return 1;
      s.it_rows.push_back(fr);
      s.n_items++;
      s.stamp.push_back(-1);
      break;
    }
    case TAG_USER: {  // proc_user, base.h:719-740
      if (s.open) {  // a section abandoned without PROCESS: drop what it collected
        s.secs.pop_back();
        while (!s.pos.empty() && s.pos.back().sec == s.n_closed) s.pos.pop_back();
        while (!s.ban.empty() && s.ban.back().sec == s.n_closed) s.ban.pop_back();
        while (!s.spec.empty() && s.spec.back().sec == s.n_closed) s.spec.pop_back();
      }
      Sec sec;
      sec.u0 = (int)s.u_ent.size();
      for (int i = 0; i < r.nu; ++i) {
        const unsigned uid = r.ui[i];
        if (!(uid < (unsigned)h->shape.num_user)) return fail(h, "user feature index exceed bound");
        s.u_ent.push_back(Entry{uid, r.uv[i], 0.f, 0.f});
        const svdgpu::Side &sd = h->side_u;
        if (sd.on() && uid + 1 < sd.rp.size())
          for (unsigned j = sd.rp[uid]; j < sd.rp[uid + 1]; ++j) s.u_ent.push_back(Entry{sd.idx[j], sd.val[j], 0.f, 0.f});
      }
      sec.u1 = (int)s.u_ent.size();
      sec.f0 = (int)s.f_ent.size();
      if (h->shape.format_type == 1)
        for (size_t i = 0; i < s.cur_fbi.size(); ++i) s.f_ent.push_back(Entry{s.cur_fbi[i], s.cur_fbv[i], 0.f, 0.f});
      sec.f1 = (int)s.f_ent.size();
      sec.n_cand = 0;
      s.secs.push_back(sec);
      s.open = true;
      s.sec_id++;
      break;
    }
    case TAG_POS:
    case TAG_BAN: {  // proc_tag, base.h:741-749
      if (!s.open) return fail(h, "ranker: POS/BAN row outside a user section");
      for (int i = 0; i < r.nu; ++i) {
        const int idx = (int)r.ui[i];
        if (!(idx >= 0 && idx < s.n_items)) return fail(h, "sample item index exceed bound");
        if (s.stamp[idx] == s.sec_id) return fail(h, "each pos sample item can not occur in baned sample list");
        s.stamp[idx] = s.sec_id;
        (r.tag == TAG_POS ? s.pos : s.ban).push_back(Mark{s.n_closed, idx});
      }
      break;
    }
    case TAG_SPEC: {  // proc_spec, base.h:750-758
      if (!s.open) return fail(h, "ranker: SPEC row outside a user section");
      if (!(r.nu == 1)) return fail(h, "must specify item index of sample in user feature field\n");
      const int idx = (int)r.ui[0];
      if (!(idx >= 0 && idx < s.n_items)) return fail(h, "sample item index exceed bound");
      Spec sp;
      sp.sec = s.n_closed;
      sp.item = idx;
      if (expand_item_row(h, r, s.sp_ent, s.sp_g, sp.row)) return 1;
      // a later SPEC row of the same item replaces the earlier one (plain assignment, base.h:757)
      for (size_t j = s.spec.size(); j-- > 0 && s.spec[j].sec == sp.sec;)
        if (s.spec[j].item == idx) s.spec.erase(s.spec.begin() + (long)j);
      s.spec.push_back(sp);
      break;
    }
    case TAG_PROCESS: {  // proc_rank, base.h:759-782
      if (!s.open) return fail(h, "ranker: PROCESS row outside a user section");
      s.secs.back().n_cand = s.n_items;
      if (s.top_k > 0) {
        int nban = 0;
        for (size_t j = s.ban.size(); j-- > 0 && s.ban[j].sec == s.n_closed;) nban++;
        if (!(s.n_items - nban >= s.top_k)) return fail(h, "k can not exceed candidate size");
      }
      s.open = false;
      s.n_closed++;
      break;
    }
    default: break;  // unknown tags are ignored (base.h:785-792)
  }
  return 0;
}

int flush_items(svdgpu *h) {
  svdgpu_rank_state &s = *h->rank;
  const int n_new = s.n_items - s.n_items_dev;
  if (n_new <= 0) return 0;
  if (s.d_rows.upload(h, s.it_rows) || s.d_ent.upload(h, s.it_ent) || s.d_g.upload(h, s.it_g)) return 1;
  const int warps_per_block = 8;
  k_rank_items<<<(n_new + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, h->stream>>>(
      h->dm, s.d_rows.p, s.d_ent.p, s.d_g.p, n_new, s.n_items_dev, s.d_IF.p, s.d_ibias.p, s.cap_items);
  CU(h, cudaGetLastError());
  h->n_launch++;
  CU(h, cudaStreamSynchronize(h->stream));  // the host vectors are reused
  s.n_items_dev = s.n_items;
  s.it_rows.clear();
  s.it_ent.clear();
  s.it_g.clear();
  return 0;
}

// rank sections [first, first + n_sec) of the closed ones; their results go to s.h_out
int run_sections(svdgpu *h, int first, int n_sec, size_t pos0, size_t pos1, size_t ban0, size_t ban1, size_t sp0,
                 size_t sp1) {
  svdgpu_rank_state &s = *h->rank;
  const DevModel &m = h->dm;
  const int chunks = m.pitch >> 2, width = s.n_items_dev;
  std::vector<Sec> secs(s.secs.begin() + first, s.secs.begin() + first + n_sec);
  auto rebase = [&](std::vector<Mark> v) {
    for (auto &x : v) x.sec -= first;
    return v;
  };
  std::vector<Mark> pos = rebase(std::vector<Mark>(s.pos.begin() + (long)pos0, s.pos.begin() + (long)pos1));
  std::vector<Mark> ban = rebase(std::vector<Mark>(s.ban.begin() + (long)ban0, s.ban.begin() + (long)ban1));
  std::vector<Spec> spec(s.spec.begin() + (long)sp0, s.spec.begin() + (long)sp1);
  for (auto &x : spec) x.sec -= first;
  if (s.d_secs.upload(h, secs) || s.d_uent.upload(h, s.u_ent) || s.d_fent.upload(h, s.f_ent)) return 1;
  if (s.d_U.reserve(h, (size_t)n_sec * m.pitch)) return 1;
  k_rank_users<<<(n_sec + 7) / 8, 256, 0, h->stream>>>(m, s.d_secs.p, s.d_uent.p, s.d_fent.p, n_sec, (float4 *)s.d_U.p);
  CU(h, cudaGetLastError());
  h->n_launch++;
  const size_t cells = (size_t)n_sec * width;
  if (s.d_score.reserve(h, cells)) return 1;
  CU(h, cudaMemsetAsync(s.d_score.p, 0, cells * sizeof(float), h->stream));  // item_score = 0 (base.h:738)
  if (!spec.empty()) {
    if (s.d_spec.upload(h, spec) || s.d_spent.upload(h, s.sp_ent) || s.d_spg.upload(h, s.sp_g)) return 1;
    const int wpb = 4;
    const size_t smem = (size_t)wpb * chunks * sizeof(float4);
    k_rank_spec<<<((int)spec.size() + wpb - 1) / wpb, wpb * 32, smem, h->stream>>>(
        m, s.d_spec.p, s.d_spent.p, s.d_spg.p, (int)spec.size(), (const float4 *)s.d_U.p, s.d_score.p, width);
    CU(h, cudaGetLastError());
    h->n_launch++;
  }
  if (!ban.empty()) {
    if (s.d_ban.upload(h, ban)) return 1;
    k_rank_mark<<<((int)ban.size() + 255) / 256, 256, 0, h->stream>>>(s.d_ban.p, (int)ban.size(), s.d_score.p, width);
    CU(h, cudaGetLastError());
    h->n_launch++;
  }
  {
    dim3 grid((unsigned)((width + 127) / 128), (unsigned)n_sec);
    k_rank_score<<<grid, 128, (size_t)chunks * sizeof(float4), h->stream>>>(
        m, s.d_secs.p, (const float4 *)s.d_U.p, s.d_IF.p, s.d_ibias.p, s.cap_items, s.d_score.p, width);
    CU(h, cudaGetLastError());
    h->n_launch++;
  }
  size_t n_out = 0;
  if (s.top_k > 0 && s.top_k <= RANK_SELECT_MAX && !h->rank_force_sort) {
    n_out = (size_t)n_sec * s.top_k;
    if (s.d_out.reserve(h, n_out)) return 1;
    k_rank_select<<<(unsigned)n_sec, 256, 0, h->stream>>>(s.d_score.p, width, s.top_k, s.d_out.p);
    CU(h, cudaGetLastError());
    h->n_launch++;
  } else if (s.top_k > 0) {
    n_out = (size_t)n_sec * s.top_k;
    if (s.d_key.reserve(h, cells) || s.d_key2.reserve(h, cells) || s.d_val.reserve(h, cells) ||
        s.d_val2.reserve(h, cells) || s.d_out.reserve(h, n_out))
      return 1;
    k_rank_keys<<<(unsigned)((cells + 255) / 256), 256, 0, h->stream>>>(s.d_score.p, (long long)cells, width, s.d_key.p,
                                                                       s.d_val.p);
    CU(h, cudaGetLastError());
    h->n_launch++;
    int sec_bits = 1;
    while ((1LL << sec_bits) < n_sec) sec_bits++;
    size_t tmp_bytes = 0;
    CU(h, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, s.d_key.p, s.d_key2.p, s.d_val.p, s.d_val2.p,
                                          (long long)cells, 0, 32 + sec_bits, h->stream));
    if (s.d_tmp.reserve(h, tmp_bytes)) return 1;
    CU(h, cub::DeviceRadixSort::SortPairs(s.d_tmp.p, tmp_bytes, s.d_key.p, s.d_key2.p, s.d_val.p, s.d_val2.p,
                                          (long long)cells, 0, 32 + sec_bits, h->stream));
    h->n_launch++;
    k_rank_take<<<(unsigned)((n_out + 255) / 256), 256, 0, h->stream>>>(s.d_val2.p, width, s.top_k, n_sec, s.d_out.p);
    CU(h, cudaGetLastError());
    h->n_launch++;
  } else {
    n_out = pos.size();
    if (n_out) {
      if (s.d_pos.upload(h, pos) || s.d_out.reserve(h, n_out)) return 1;
      k_rank_pos<<<(unsigned)n_out, 256, 0, h->stream>>>(s.d_pos.p, s.d_score.p, width, s.d_out.p);
      CU(h, cudaGetLastError());
      h->n_launch++;
    }
  }
  const size_t at = s.h_out.size();
  s.h_out.resize(at + n_out);
  if (n_out) CU(h, cudaMemcpyAsync(s.h_out.data() + at, s.d_out.p, n_out * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  h->n_d2h += (long long)(n_out * sizeof(int));
  return 0;
}

// execute every closed section, in batches bounded by the score matrix (<= 2^25 cells)
int run_closed(svdgpu *h) {
  svdgpu_rank_state &s = *h->rank;
  if (s.n_closed == 0) return 0;
  if (flush_items(h)) return 1;
  const int width = std::max(s.n_items_dev, 1);
  const int per = (int)std::max<long long>(1, std::min<long long>((1LL << 25) / width, 65535));
  size_t p0 = 0, b0 = 0, x0 = 0;
  for (int first = 0; first < s.n_closed; first += per) {
    const int n_sec = std::min(per, s.n_closed - first);
    size_t p1 = p0, b1 = b0, x1 = x0;
    while (p1 < s.pos.size() && s.pos[p1].sec < first + n_sec) p1++;
    while (b1 < s.ban.size() && s.ban[b1].sec < first + n_sec) b1++;
    while (x1 < s.spec.size() && s.spec[x1].sec < first + n_sec) x1++;
    if (s.n_items_dev > 0) {
      if (run_sections(h, first, n_sec, p0, p1, b0, b1, x0, x1)) return 1;
    }
    p0 = p1;
    b0 = b1;
    x0 = x1;
  }
  // keep only what belongs to a still-open section
  auto keep_open = [&](auto &v) {
    size_t w = 0;
    for (size_t j = 0; j < v.size(); ++j)
      if (v[j].sec >= s.n_closed) {
        v[w] = v[j];
        v[w].sec = 0;
        w++;
      }
    v.resize(w);
  };
  keep_open(s.pos);
  keep_open(s.ban);
  keep_open(s.spec);
  if (s.open) {
    // the open section's feature ranges stay valid only if the arrays are kept: re-base them
    Sec o = s.secs.back();
    std::vector<Entry> u(s.u_ent.begin() + o.u0, s.u_ent.begin() + o.u1), f(s.f_ent.begin() + o.f0, s.f_ent.begin() + o.f1);
    o.u1 -= o.u0; o.u0 = 0; o.f1 -= o.f0; o.f0 = 0;
    s.u_ent.swap(u);
    s.f_ent.swap(f);
    s.secs.assign(1, o);
    // SPEC rows of the open section keep their entries (sp_ent / sp_g are not compacted)
  } else {
    s.secs.clear();
    s.u_ent.clear();
    s.f_ent.clear();
    s.sp_ent.clear();
    s.sp_g.clear();
  }
  s.n_closed = 0;
  return 0;
}

int deliver(svdgpu *h, int *result, long long cap, long long *num_result) {
  svdgpu_rank_state &s = *h->rank;
  const long long n = (long long)s.h_out.size();
  if (num_result) *num_result = n;
  if (n > cap) return fail(h, "rank: %lld results do not fit the buffer of %lld (call again with a larger one)", n, cap);
  if (n) memcpy(result, s.h_out.data(), (size_t)n * sizeof(int));
  s.h_out.clear();
  return 0;
}

Row row_of(int r, const int *row_ptr, const float *label, const unsigned *index, const float *value) {
  const int *p = row_ptr + 3 * (size_t)r;
  Row x;
  x.tag = (int)label[r];
  x.ng = p[1] - p[0];
  x.nu = p[2] - p[1];
  x.ni = p[3] - p[2];
  x.gi = index + p[0]; x.gv = value + p[0];
  x.ui = index + p[1]; x.uv = value + p[1];
  x.ii = index + p[2]; x.iv = value + p[2];
  return x;
}

int check_ready(svdgpu *h) {
  if (!h->rank) return fail(h, "rank: call svdgpu_rank_init first");
  CU(h, cudaSetDevice(h->device));
  return 0;
}

}  // namespace

namespace svdk {
void rank_free(svdgpu *h) {
  if (!h->rank) return;
  h->rank->release();
  delete h->rank;
  h->rank = nullptr;
}
}  // namespace svdk

int svdgpu_rank_init(svdgpu_t *h, int num_item_set, int top_k) {
  if (!h) return 1;
  if (num_item_set < 0) return fail(h, "rank_init: num_item_set < 0");
  CU(h, cudaSetDevice(h->device));
  rank_free(h);
  h->rank = new svdgpu_rank_state();
  svdgpu_rank_state &s = *h->rank;
  s.cap_items = num_item_set;
  s.top_k = top_k;
  const size_t chunks = (size_t)(h->dm.pitch >> 2);
  if (s.d_IF.reserve(h, std::max<size_t>(1, chunks * (size_t)num_item_set))) return 1;
  if (s.d_ibias.reserve(h, std::max<size_t>(1, (size_t)num_item_set))) return 1;
  return 0;
}

int svdgpu_rank_csr(svdgpu_t *h, int num_row, const int *row_ptr, const float *label, const unsigned *index,
                    const float *value, int *result, long long result_cap, long long *num_result) {
  if (!h) return 1;
  if (check_ready(h)) return 1;
  if (num_row > 0 && (!row_ptr || !label)) return fail(h, "rank_csr: null array");
  for (int r = 0; r < num_row; ++r)
    if (feed_row(h, row_of(r, row_ptr, label, index, value))) return 1;
  if (run_closed(h)) return 1;
  return deliver(h, result, result_cap, num_result);
}

int svdgpu_rank_ugroup(svdgpu_t *h, int num_block, const int *blk_row_off, const int *blk_fb_off, const int *blk_tag,
                       const unsigned *fb_index, const float *fb_value, const int *row_ptr, const float *label,
                       const unsigned *index, const float *value, int *result, long long result_cap,
                       long long *num_result) {
  if (!h) return 1;
  if (check_ready(h)) return 1;
  if (num_block > 0 && (!blk_row_off || !blk_fb_off)) return fail(h, "rank_ugroup: null array");
  svdgpu_rank_state &s = *h->rank;
  for (int b = 0; b < num_block; ++b) {
    const int tag = blk_tag ? blk_tag[b] : 0;
    if (tag == 0 || tag == 1) {  // DEFAULT / START_TAG: a new feedback list (base.h:799-808)
      s.cur_fbi.clear();
      s.cur_fbv.clear();
      for (int i = blk_fb_off[b]; i < blk_fb_off[b + 1]; ++i) {
        if (!(fb_index[i] < (unsigned)h->shape.num_ufeedback)) return fail(h, "ufeedback id exceed bound");
        s.cur_fbi.push_back(fb_index[i]);
        s.cur_fbv.push_back(fb_value[i]);
      }
    }
    for (int r = blk_row_off[b]; r < blk_row_off[b + 1]; ++r)
      if (feed_row(h, row_of(r, row_ptr, label, index, value))) return 1;
  }
  if (run_closed(h)) return 1;
  return deliver(h, result, result_cap, num_result);
}
