// svdgpu_pairs.cu -- pairwise-rank sample generation on the device.
//
// The reference turns every user block of rated rows into a block of PAIR rows on the host
// (PairwiseRankGenerator, apex_svd_data.cpp:812-1025): rows with label >= pos_sample_lowerb are
// positives, rows with label <= neg_sample_upperb negatives; both lists are shuffled and pair i
// is (pos[i % |pos|], neg[i % |neg|]) for i < rank_sample_num (default |neg|), its features the
// index-sorted merge of the two rows with the negative's values negated (:828-860, :886-915).
// Once the SGD kernels run at GB/s speed that host loop is the bottleneck, so here the pairs are
// generated from a user-grouped batch resident in HBM into another resident batch:
//
//   rank_sample_method 1 (sample_cmp, :920-944: every row pairs with one random row of the same
//   block whose label differs by more than rank_sample_gap):
//   (segmented sort) the rows of every block by label
//   k_pair_cmp     one warp per block: per row two binary searches in the sorted labels give the
//                  eligible partners, one is drawn; a scan compacts the pairs
//   rank_sample_method 0 (sample_posneg):
//   k_pair_split   one warp per block: stable compaction of the block's positive / negative rows
//   (scan)         pair rows per block -> first pair row of every block
//   k_pair_shape   one thread per pair row: pick (p, n), count the merged features
//   (scan)         feature counts -> row_ptr of the pair batch
//   k_pair_fill    one thread per pair row: write label, merged index/value
//
// The reference's glibc rand() shuffle cannot be reproduced on a GPU; the shuffles are replaced
// by per-block affine permutations x -> (a*x + c) mod n keyed by (seed, block), which keep the
// property the cycling relies on: every row of a list is used once before any is reused.
// Parity is therefore structural (tests/test_gpu_pairs.py), not bit-level; the host sampler of
// the reference remains the path for bit-exact runs.
#include "svdgpu_internal.h"

#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_sort.cuh>

using namespace svdk;

namespace {

struct PairCfg {
  int pointwise, label_diff, sample_num, sample_max;
  int explicit_pairs;  // method 1: pair row s is (pos_list[s], neg_list[s]) itself
  float lowerb, upperb, gap;
  unsigned long long seed;
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {  // splitmix64 finaliser
  x += 0x9e3779b97f4a7c15ULL;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
  return x ^ (x >> 31);
}
__device__ __forceinline__ unsigned gcd_u(unsigned a, unsigned b) {
  while (b) {
    const unsigned t = a % b;
    a = b;
    b = t;
  }
  return a;
}
// position x of a pseudo-random permutation of [0, n) keyed by `key`
__device__ __forceinline__ unsigned perm_at(unsigned x, unsigned n, unsigned long long key) {
  if (n <= 1) return 0;
  const unsigned long long h = mix64(key);
  unsigned a = (unsigned)(h % n);
  if (a == 0) a = 1;
  while (gcd_u(a, n) != 1) a = a + 1 >= n ? 1 : a + 1;
  const unsigned c = (unsigned)((h >> 32) % n);
  return (unsigned)(((unsigned long long)a * x + c) % n);
}

// ---- positives / negatives of every block, in row order (sample_posneg, :946-952) ----------
__global__ void k_pair_split(const float *label, const int *blk_row_off, int num_block, PairCfg cfg, int *pos_list,
                             int *neg_list, int *npos, int *nneg, int *out_rows) {
  const int b = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (b >= num_block) return;
  const int r0 = blk_row_off[b], r1 = blk_row_off[b + 1];
  int np = 0, nn = 0;
  for (int base = r0; base < r1; base += 32) {
    const int r = base + lane;
    const float lab = r < r1 ? label[r] : 0.0f;
    const bool isp = r < r1 && __fsub_rn(lab, cfg.lowerb) > -1e-6f;
    const bool isn = r < r1 && __fsub_rn(lab, cfg.upperb) < 1e-6f;
    const unsigned mp = __ballot_sync(0xffffffffu, isp), mn = __ballot_sync(0xffffffffu, isn);
    const unsigned below = (1u << lane) - 1u;
    if (isp) pos_list[r0 + np + __popc(mp & below)] = r;
    if (isn) neg_list[r0 + nn + __popc(mn & below)] = r;
    np += __popc(mp);
    nn += __popc(mn);
  }
  if (lane == 0) {
    long long snum = 0;
    if (np > 0 && nn > 0) {  // :953-964
      snum = cfg.sample_num > 0 ? cfg.sample_num : nn;
      if (cfg.sample_max > 0 && snum > cfg.sample_max) snum = cfg.sample_max;
    }
    npos[b] = np;
    nneg[b] = nn;
    out_rows[b] = (int)(cfg.pointwise ? 2 * snum : snum);
  }
}

// ---- sample_cmp (:920-944): for every row (visited in a keyed pseudo-random order, the
// reference shuffles) count the rows of the block with label < label - gap ("left") and with
// label >= (label - gap) + 2*gap ("right" .. end) in the label-sorted block, draw one of them;
// a lower-labelled partner makes this row the positive of the pair, a higher-labelled one the
// negative.  has[slot] = a partner exists; pp / nn = the pair.
__global__ void k_pair_cmp(const float *label, const int *blk_row_off, int num_block, const float *sorted_label,
                           const int *sorted_row, PairCfg cfg, int *has, int *pp, int *nn) {
  const int b = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (b >= num_block) return;
  const int r0 = blk_row_off[b], n = blk_row_off[b + 1] - r0;
  const float *sl = sorted_label + r0;
  const unsigned long long key = cfg.seed * 0x100000001b3ULL + (unsigned long long)b;
  auto lower_bound = [&](float x) {  // first position whose label is not < x
    int lo = 0, hi = n;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (sl[mid] < x) lo = mid + 1;
      else hi = mid;
    }
    return lo;
  };
  for (int t = lane; t < n; t += 32) {
    const int r = r0 + (int)perm_at((unsigned)t, (unsigned)n, key);  // the row visited t-th
    const float lo_lab = __fsub_rn(label[r], cfg.gap);
    const int left = lower_bound(lo_lab);
    const int right = lower_bound(__fadd_rn(lo_lab, __fmul_rn(cfg.gap, 2.0f)));
    const unsigned rng = (unsigned)(left + n - right);
    int ok = 0, p = 0, q = 0;
    if (rng > 0) {
      const unsigned idx = (unsigned)(mix64(key ^ (0x9e3779b97f4a7c15ULL * (unsigned long long)(t + 1))) % rng);
      ok = 1;
      if (idx < (unsigned)left) {
        p = r;
        q = sorted_row[r0 + idx];
      } else {
        p = sorted_row[r0 + right + (idx - left)];
        q = r;
      }
    }
    has[r0 + t] = ok;
    pp[r0 + t] = p;
    nn[r0 + t] = q;
  }
}
// compact the pairs (slot -> pair row at[slot]) and derive every block's first pair row
__global__ void k_pair_compact(const int *has, const int *at, const int *pp, const int *nn, int nrow, int *out_p,
                               int *out_n, const int *blk_row_off, int num_block, int mult, int *out_off) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nrow && has[t]) {
    out_p[at[t]] = pp[t];
    out_n[at[t]] = nn[t];
  }
  if (t <= num_block) out_off[t] = mult * at[blk_row_off[t]];  // at[] has nrow + 1 entries
}
__global__ void k_iota(int *v, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) v[t] = t;
}

struct RowView {
  int g0, u0, i0, i1;  // segment bounds of a source row
};
__device__ __forceinline__ RowView view(const int *rp, int r) {
  RowView v;
  v.g0 = rp[3 * (long long)r];
  v.u0 = rp[3 * (long long)r + 1];
  v.i0 = rp[3 * (long long)r + 2];
  v.i1 = rp[3 * (long long)r + 3];
  return v;
}
// merge of two index-sorted segments (:828-860): equal indices collapse into one entry
template <bool WRITE>
__device__ __forceinline__ int merge_seg(const unsigned *idx, const float *val, int a0, int a1, int b0, int b1,
                                         unsigned *oidx, float *oval) {
  int i = a0, j = b0, n = 0;
  while (i < a1 && j < b1) {
    if (idx[i] < idx[j]) {
      if (WRITE) { oidx[n] = idx[i]; oval[n] = val[i]; }
      ++i;
    } else if (idx[j] < idx[i]) {
      if (WRITE) { oidx[n] = idx[j]; oval[n] = -val[j]; }
      ++j;
    } else {
      if (WRITE) { oidx[n] = idx[i]; oval[n] = __fsub_rn(val[i], val[j]); }
      ++i;
      ++j;
    }
    ++n;
  }
  for (; i < a1; ++i, ++n)
    if (WRITE) { oidx[n] = idx[i]; oval[n] = val[i]; }
  for (; j < b1; ++j, ++n)
    if (WRITE) { oidx[n] = idx[j]; oval[n] = -val[j]; }
  return n;
}
// user features with |value| > 1e-6 (:897-903)
template <bool WRITE>
__device__ __forceinline__ int user_seg(const unsigned *idx, const float *val, int a0, int a1, unsigned *oidx,
                                        float *oval) {
  int n = 0;
  for (int i = a0; i < a1; ++i)
    if (val[i] > 1e-6f || val[i] < -1e-6f) {
      if (WRITE) { oidx[n] = idx[i]; oval[n] = val[i]; }
      ++n;
    }
  return n;
}

struct Pick {
  int b, p, n, half;  // block, positive row, negative row, (pointwise) 0 = the positive's row, 1 = the negative's
};
__device__ __forceinline__ Pick pick(long long s, const int *out_off, int num_block, const int *blk_row_off,
                                     const int *pos_list, const int *neg_list, const int *npos, const int *nneg,
                                     const PairCfg &cfg) {
  Pick k;
  if (cfg.explicit_pairs) {  // method 1: the pairs were drawn by k_pair_cmp
    const long long pi = cfg.pointwise ? (s >> 1) : s;
    k.b = 0;
    k.half = cfg.pointwise ? (int)(s & 1) : 0;
    k.p = pos_list[pi];
    k.n = neg_list[pi];
    return k;
  }
  int lo = 0, hi = num_block;  // last block whose first pair row is <= s
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (out_off[mid] <= s) lo = mid;
    else hi = mid;
  }
  k.b = lo;
  const long long i = s - out_off[lo];
  const long long pi = cfg.pointwise ? (i >> 1) : i;
  k.half = cfg.pointwise ? (int)(i & 1) : 0;
  const unsigned np = (unsigned)npos[lo], nn = (unsigned)nneg[lo];
  const unsigned long long key = cfg.seed * 0x100000001b3ULL + (unsigned long long)lo;
  const int r0 = blk_row_off[lo];
  k.p = pos_list[r0 + perm_at((unsigned)(pi % np), np, key)];
  k.n = neg_list[r0 + perm_at((unsigned)(pi % nn), nn, key ^ 0x5bd1e995ULL)];
  return k;
}

template <bool FILL>
__global__ void k_pair_rows(const int *rp, const float *label, const unsigned *idx, const float *val,
                            const int *blk_row_off, int num_block, const int *out_off, long long total,
                            const int *pos_list, const int *neg_list, const int *npos, const int *nneg, PairCfg cfg,
                            int *cnt /* shape pass: 3 counts per row */, const int *out_rp, float *out_label,
                            unsigned *out_idx, float *out_val) {
  const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (s >= total) return;
  const Pick k = pick(s, out_off, num_block, blk_row_off, pos_list, neg_list, npos, nneg, cfg);
  const RowView p = view(rp, k.p), n = view(rp, k.n);
  unsigned *oi = FILL ? out_idx + out_rp[3 * s] : nullptr;
  float *ov = FILL ? out_val + out_rp[3 * s] : nullptr;
  int ng, nu, ni;
  if (cfg.pointwise) {  // genpair_pointwise, :862-884
    const RowView r = k.half ? n : p;
    ng = r.u0 - r.g0;
    if (FILL)
      for (int i = 0; i < ng; ++i) { oi[i] = idx[r.g0 + i]; ov[i] = val[r.g0 + i]; }
    nu = user_seg<FILL>(idx, val, r.u0, r.i0, FILL ? oi + ng : nullptr, FILL ? ov + ng : nullptr);
    ni = r.i1 - r.i0;
    if (FILL)
      for (int i = 0; i < ni; ++i) { oi[ng + nu + i] = idx[r.i0 + i]; ov[ng + nu + i] = val[r.i0 + i]; }
    if (FILL) out_label[s] = k.half ? 0.0f : 1.0f;
  } else {  // genpair, :886-915
    ng = merge_seg<FILL>(idx, val, p.g0, p.u0, n.g0, n.u0, oi, ov);
    nu = user_seg<FILL>(idx, val, p.u0, p.i0, FILL ? oi + ng : nullptr, FILL ? ov + ng : nullptr);
    ni = merge_seg<FILL>(idx, val, p.i0, p.i1, n.i0, n.i1, FILL ? oi + ng + nu : nullptr, FILL ? ov + ng + nu : nullptr);
    if (FILL) out_label[s] = cfg.label_diff ? __fsub_rn(label[k.p], label[k.n]) : 1.0f;
  }
  if (!FILL) {
    cnt[3 * s] = ng;
    cnt[3 * s + 1] = nu;
    cnt[3 * s + 2] = ni;
  }
}

int dev_alloc(svdgpu *h, DevBuf &b, size_t bytes) {
  bytes += 64;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  CU(h, cudaMalloc(&b.p, bytes));
  b.cap = bytes;
  return 0;
}
int d2d(svdgpu *h, DevBuf &dst, const DevBuf &src, size_t bytes) {
  if (dev_alloc(h, dst, bytes)) return 1;
  if (bytes) CU(h, cudaMemcpyAsync(dst.p, src.p, bytes, cudaMemcpyDeviceToDevice, h->stream));
  return 0;
}
template <bool INCLUSIVE>
int scan_int(svdgpu *h, const int *in, int *out, long long n) {
  if (n <= 0) return 0;
  if (n > 0x7fffffffLL) return fail(h, "sample_pairs: more than 2^31 items");
  size_t tmp_bytes = 0;
  if (INCLUSIVE) CU(h, cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, in, out, (int)n, h->stream));
  else CU(h, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, (int)n, h->stream));
  void *tmp = nullptr;
  CU(h, cudaMalloc(&tmp, tmp_bytes + 16));
  cudaError_t e = INCLUSIVE ? cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, in, out, (int)n, h->stream)
                            : cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, (int)n, h->stream);
  cudaStreamSynchronize(h->stream);
  cudaFree(tmp);
  if (e != cudaSuccess) return fail(h, "cub scan failed: %s", cudaGetErrorString(e));
  return 0;
}

}  // namespace

extern "C" int svdgpu_batch_sample_pairs(svdgpu_t *h, svdgpu_batch_t *src, const svdgpu_pair_params *pp,
                                         svdgpu_batch_t **out) {
  if (!h || !out) return 1;
  *out = nullptr;
  if (!src || !pp) return fail(h, "sample_pairs: null argument");
  CU(h, cudaSetDevice(h->device));
  if (!src->ugroup) return fail(h, "sample_pairs: the source batch must be user grouped (svdgpu_batch_set_ugroup)");
  if (src->num_unit != src->num_block) return fail(h, "sample_pairs: START/MIDDLE/END blocks are not supported");
  if (src->has_value2) return fail(h, "sample_pairs: side features are not supported");
  const int method = pp->rank_sample_method % 10;
  if ((method != 0 && method != 1) || pp->rank_sample_method < 0 || pp->rank_sample_method > 11)
    return fail(h, "unkown rank sample method\n");  // apex_svd_data.cpp:1010
  if (method == 1 && !(pp->rank_sample_gap > 0.0f))
    return fail(h, "must set rank_sample_gap to a value bigger than 0");  // apex_svd_data.cpp:994
  PairCfg cfg;
  cfg.explicit_pairs = method == 1;
  cfg.gap = pp->rank_sample_gap;
  cfg.pointwise = pp->rank_sample_pointwise != 0;
  cfg.label_diff = pp->rank_sample_method / 10 != 0;
  cfg.sample_num = pp->rank_sample_num;
  cfg.sample_max = pp->rank_sample_max;
  cfg.lowerb = pp->pos_sample_lowerb;
  cfg.upperb = pp->neg_sample_upperb;
  cfg.seed = pp->seed;
  const int nb = src->num_block, nrow = src->num_row;
  const int *rp = (const int *)src->d_rp.p, *bro = (const int *)src->d_blk_row_off.p;
  const float *label = (const float *)src->d_label.p, *val = (const float *)src->d_value.p;
  const unsigned *idx = (const unsigned *)src->d_index.p;

  DevBuf pos, neg, npos, nneg, orow, ooff, cnt, slab, srow, rowid, has, at, cp, cn, sort_tmp;
  int rc = 0;
  svdgpu_batch *b = nullptr;
  long long total = 0, nnz = 0;
  do {
    if ((rc = dev_alloc(h, pos, (size_t)std::max(nrow, 1) * 4) | dev_alloc(h, neg, (size_t)std::max(nrow, 1) * 4) |
              dev_alloc(h, npos, (size_t)(nb + 1) * 4) | dev_alloc(h, nneg, (size_t)(nb + 1) * 4) |
              dev_alloc(h, orow, (size_t)(nb + 1) * 4) | dev_alloc(h, ooff, (size_t)(nb + 2) * 4)))
      break;
    cudaMemsetAsync(orow.p, 0, (size_t)(nb + 1) * 4, h->stream);
    if (cfg.explicit_pairs) {
      // sample_cmp: sort every block's rows by label, draw one partner per row, compact
      const size_t rbytes = (size_t)std::max(nrow, 1) * 4;
      if ((rc = dev_alloc(h, slab, rbytes) | dev_alloc(h, srow, rbytes) | dev_alloc(h, rowid, rbytes) |
                dev_alloc(h, has, rbytes + 4) | dev_alloc(h, at, rbytes + 4) | dev_alloc(h, cp, rbytes) |
                dev_alloc(h, cn, rbytes)))
        break;
      cudaMemsetAsync(has.p, 0, rbytes + 4, h->stream);
      if (nrow > 0 && nb > 0) {
        k_iota<<<(nrow + 255) / 256, 256, 0, h->stream>>>((int *)rowid.p, nrow);
        size_t tb = 0;
        cudaError_t e = cub::DeviceSegmentedSort::SortPairs(nullptr, tb, label, (float *)slab.p, (const int *)rowid.p,
                                                            (int *)srow.p, nrow, nb, bro, bro + 1, h->stream);
        if (e == cudaSuccess && !(rc = dev_alloc(h, sort_tmp, tb + 16)))
          e = cub::DeviceSegmentedSort::SortPairs(sort_tmp.p, tb, label, (float *)slab.p, (const int *)rowid.p,
                                                  (int *)srow.p, nrow, nb, bro, bro + 1, h->stream);
        if (rc) break;
        if (e != cudaSuccess) { rc = fail(h, "sample_pairs: segmented sort failed: %s", cudaGetErrorString(e)); break; }
        k_pair_cmp<<<(int)(((long long)nb * 32 + 255) / 256), 256, 0, h->stream>>>(
            label, bro, nb, (const float *)slab.p, (const int *)srow.p, cfg, (int *)has.p, (int *)pos.p, (int *)neg.p);
        h->n_launch += 3;
      }
      if ((rc = scan_int<false>(h, (const int *)has.p, (int *)at.p, (long long)nrow + 1))) break;
      const int work = std::max(nrow, nb + 1);
      k_pair_compact<<<(work + 255) / 256, 256, 0, h->stream>>>((const int *)has.p, (const int *)at.p, (const int *)pos.p,
                                                               (const int *)neg.p, nrow, (int *)cp.p, (int *)cn.p, bro, nb,
                                                               cfg.pointwise ? 2 : 1, (int *)ooff.p);
      h->n_launch++;
      std::swap(pos, cp);  // k_pair_rows reads the compacted pairs through pos / neg
      std::swap(neg, cn);
    } else {
      if (nb > 0) {
        k_pair_split<<<(int)(((long long)nb * 32 + 255) / 256), 256, 0, h->stream>>>(
            label, bro, nb, cfg, (int *)pos.p, (int *)neg.p, (int *)npos.p, (int *)nneg.p, (int *)orow.p);
        h->n_launch++;
      }
      // out_off[b] = first pair row of block b; out_off[nb] = total (orow[nb] is 0)
      if ((rc = scan_int<false>(h, (const int *)orow.p, (int *)ooff.p, nb + 1))) break;
    }
    int tot32 = 0;
    if (cudaMemcpy(&tot32, (const int *)ooff.p + nb, 4, cudaMemcpyDeviceToHost) != cudaSuccess) { rc = fail(h, "sample_pairs: D2H failed"); break; }
    total = tot32;
    if (total < 0 || total > 0x7fffffffLL / 4) { rc = fail(h, "sample_pairs: too many pair rows"); break; }
    b = new svdgpu_batch();
    b->num_row = (int)total;
    if ((rc = dev_alloc(h, cnt, (size_t)(3 * total + 1) * 4) | dev_alloc(h, b->d_rp, (size_t)(3 * total + 1) * 4) |
              dev_alloc(h, b->d_label, (size_t)std::max<long long>(total, 1) * 4)))
      break;
    cudaMemsetAsync(b->d_rp.p, 0, 4, h->stream);
    const int grid = (int)((total + 255) / 256);
    if (total > 0) {
      k_pair_rows<false><<<grid, 256, 0, h->stream>>>(rp, label, idx, val, bro, nb, (const int *)ooff.p, total,
                                                      (const int *)pos.p, (const int *)neg.p, (const int *)npos.p,
                                                      (const int *)nneg.p, cfg, (int *)cnt.p, nullptr, nullptr, nullptr,
                                                      nullptr);
      h->n_launch++;
      if ((rc = scan_int<true>(h, (const int *)cnt.p, (int *)b->d_rp.p + 1, 3 * total))) break;
      int nnz32 = 0;
      if (cudaMemcpy(&nnz32, (const int *)b->d_rp.p + 3 * total, 4, cudaMemcpyDeviceToHost) != cudaSuccess) { rc = fail(h, "sample_pairs: D2H failed"); break; }
      nnz = nnz32;
      if (nnz < 0) { rc = fail(h, "sample_pairs: more than 2^31 feature entries"); break; }
    }
    b->num_val = nnz;
    if ((rc = dev_alloc(h, b->d_index, (size_t)std::max<long long>(nnz, 1) * 4) |
              dev_alloc(h, b->d_value, (size_t)std::max<long long>(nnz, 1) * 4)))
      break;
    if (total > 0) {
      k_pair_rows<true><<<grid, 256, 0, h->stream>>>(rp, label, idx, val, bro, nb, (const int *)ooff.p, total,
                                                     (const int *)pos.p, (const int *)neg.p, (const int *)npos.p,
                                                     (const int *)nneg.p, cfg, nullptr, (const int *)b->d_rp.p,
                                                     (float *)b->d_label.p, (unsigned *)b->d_index.p, (float *)b->d_value.p);
      h->n_launch++;
    }
    // the pair batch keeps the source's blocks and feedback lists (PairwiseRankGenerator::next
    // replaces e.data only, :1001-1018)
    b->ugroup = true;
    b->has_fb = src->has_fb;
    b->num_block = nb;
    b->num_unit = nb;
    b->unit_off = src->unit_off;
    rc = d2d(h, b->d_blk_row_off, ooff, (size_t)(nb + 1) * 4) |
         d2d(h, b->d_unit_off, src->d_unit_off, (size_t)(nb + 1) * 4) |
         d2d(h, b->d_blk_fb_off, src->d_blk_fb_off, (size_t)(nb + 1) * 4) |
         d2d(h, b->d_order, src->d_order, (size_t)std::max(nb, 1) * 4);
    if (rc) break;
    int nfb = 0;
    if (nb > 0 && cudaMemcpy(&nfb, (const int *)src->d_blk_fb_off.p + nb, 4, cudaMemcpyDeviceToHost) != cudaSuccess) { rc = fail(h, "sample_pairs: D2H failed"); break; }
    rc = d2d(h, b->d_fbi, src->d_fbi, (size_t)nfb * 4) | d2d(h, b->d_fbv, src->d_fbv, (size_t)nfb * 4);
    if (rc) break;
    if (cudaStreamSynchronize(h->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = fail(h, "sample_pairs: kernel failed");
  } while (0);
  DevBuf *tmp[] = {&pos, &neg, &npos, &nneg, &orow, &ooff, &cnt, &slab, &srow, &rowid, &has, &at, &cp, &cn, &sort_tmp};
  for (DevBuf *t : tmp)
    if (t->p) cudaFree(t->p);
  if (rc) {
    if (b) svdgpu_batch_destroy(h, b);
    return 1;
  }
  *out = b;
  return 0;
}

// copy a resident batch's CSR arrays back to the host (inspection / tests); any pointer may be NULL
extern "C" int svdgpu_batch_download(svdgpu_t *h, svdgpu_batch_t *b, int *num_row, long long *num_val, int *row_ptr,
                                     float *label, unsigned *index, float *value, int *blk_row_off) {
  if (!h || !b) return 1;
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaStreamSynchronize(h->stream));
  if (num_row) *num_row = b->num_row;
  if (num_val) *num_val = b->num_val;
  if (row_ptr) CU(h, cudaMemcpy(row_ptr, b->d_rp.p, (3 * (size_t)b->num_row + 1) * 4, cudaMemcpyDeviceToHost));
  if (label && b->num_row) CU(h, cudaMemcpy(label, b->d_label.p, (size_t)b->num_row * 4, cudaMemcpyDeviceToHost));
  if (index && b->num_val) CU(h, cudaMemcpy(index, b->d_index.p, (size_t)b->num_val * 4, cudaMemcpyDeviceToHost));
  if (value && b->num_val) CU(h, cudaMemcpy(value, b->d_value.p, (size_t)b->num_val * 4, cudaMemcpyDeviceToHost));
  if (blk_row_off && b->ugroup)
    CU(h, cudaMemcpy(blk_row_off, b->d_blk_row_off.p, ((size_t)b->num_block + 1) * 4, cudaMemcpyDeviceToHost));
  return 0;
}
