// svdgpu_ingest.cu -- bulk ingest of the reference's binary instance buffers (host code).
//
// The reference trains from buffer files written by tools/make_feature_buffer and
// tools/make_ugroup_buffer: a loader thread reads one small batch (1000 rows by default) at a
// time and the trainer is called once per row (svd_feature.cpp:231-247).  Here a whole pass over
// such a file is ONE call: batches are read straight into pinned memory, concatenated into
// device-sized chunks (row_ptr rebased) and handed to the hot path.
//
// BINARY_BUFFER (writer apex_svd_data.cpp:131-195; reader :218-248 + apex_svd_data.h:220-230):
//   header  int num_batch, batch_size, max_batch_num
//   batch   int num_row, num_val; int row_ptr[3*num_row+1] (first = 0); float label[num_row];
//           unsigned index[num_val]; float value[num_val]
// user-group buffer (writer apex_svd_data.cpp:556-640; block apex_svd_data.h:419-450):
//   header  int num_batch, max_num_ufeedback, max_num_row, max_num_val
//   block   int num_ufeedback (bit 31 set => next int = extend_tag); unsigned fb_index[];
//           float fb_value[]; then one CSR batch as above
#include "svdgpu_internal.h"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <vector>

using namespace svdk;

namespace {

struct File {
  FILE *f = nullptr;
  ~File() {
    if (f) fclose(f);
  }
};
bool rd(FILE *f, void *dst, size_t bytes) { return bytes == 0 || fread(dst, 1, bytes, f) == bytes; }

// grow a pinned buffer of the handle (kept across calls), preserving its first `keep` bytes
bool reserve(HostBuf &b, size_t bytes, size_t keep) {
  if (bytes <= b.cap) return true;
  const size_t ncap = bytes + bytes / 2 + 4096;
  void *np = nullptr;
  if (cudaMallocHost(&np, ncap) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (b.p) {
    memcpy(np, b.p, keep);
    cudaFreeHost(b.p);
  }
  b.p = np;
  b.cap = ncap;
  return true;
}
double now_s() {
  timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

enum Task { TRAIN, PREDICT, EVAL };

// One pass over a BINARY_BUFFER file.
int pass_csr(svdgpu *h, const char *path, Task task, float *out, long long out_cap, float scale,
             double *sum_sq, long long *rows_out) {
  File fi;
  fi.f = fopen(path, "rb");
  if (!fi.f) return fail(h, "can't open buffer %s", path);
  int hdr[3];  // num_batch, batch_size, max_batch_num
  if (!rd(fi.f, hdr, sizeof(hdr))) return fail(h, "%s: truncated header", path);
  if (hdr[0] < 0) return fail(h, "%s: bad header", path);
  HostBuf &rp = h->ing_rp, &lab = h->ing_label, &idx = h->ing_index, &val = h->ing_value;
  {  // size the chunk buffers once from the header (batch_size rows / max_batch_num values per batch)
    const size_t nb = (size_t)std::max(1, h->chunk_rows / std::max(1, hdr[1]) + 1);
    const size_t rows_cap = nb * (size_t)std::max(1, hdr[1]), vals_cap = nb * (size_t)std::max(1, hdr[2]);
    if (rows_cap < (1u << 28) && vals_cap < (1u << 29) &&
        (!reserve(rp, (3 * rows_cap + 1) * 4, 0) || !reserve(lab, rows_cap * 4, 0) || !reserve(idx, vals_cap * 4, 0) ||
         !reserve(val, vals_cap * 4, 0)))
      return fail(h, "cudaMallocHost failed");
  }
  long long rows_done = 0;
  double sse = 0.0;
  long long cnt = 0;
  int b = 0;
  while (b < hdr[0]) {
    // ---- concatenate batches until the chunk holds chunk_rows rows ----
    long long n = 0, nv = 0;
    const double t_read = now_s();
    while (b < hdr[0] && n < h->chunk_rows) {
      int nn[2];
      if (!rd(fi.f, nn, sizeof(nn))) return fail(h, "%s: truncated batch %d", path, b);
      const int br = nn[0], bv = nn[1];
      if (br < 0 || bv < 0 || n + br > 0x7fffffffLL / 4 || nv + bv > 0x7fffffffLL) return fail(h, "%s: bad batch %d", path, b);
      if (!reserve(rp, (3 * (size_t)(n + br) + 1) * 4, (3 * (size_t)n + 1) * 4) || !reserve(lab, (size_t)(n + br) * 4, (size_t)n * 4) ||
          !reserve(idx, (size_t)(nv + bv) * 4, (size_t)nv * 4) || !reserve(val, (size_t)(nv + bv) * 4, (size_t)nv * 4))
        return fail(h, "cudaMallocHost failed");
      int *prp = (int *)rp.p + 3 * n;  // the batch's row_ptr[0] (= 0) lands on the running total
      if (!rd(fi.f, prp, (3 * (size_t)br + 1) * 4)) return fail(h, "%s: truncated batch %d", path, b);
      if (prp[0] != 0 || prp[3 * (size_t)br] != bv) return fail(h, "%s: batch %d: row_ptr does not match num_val", path, b);
      for (size_t i = 0; i <= 3 * (size_t)br; ++i) prp[i] += (int)nv;  // rebase
      if (!rd(fi.f, (float *)lab.p + n, (size_t)br * 4) || !rd(fi.f, (unsigned *)idx.p + nv, (size_t)bv * 4) ||
          !rd(fi.f, (float *)val.p + nv, (size_t)bv * 4))
        return fail(h, "%s: truncated batch %d", path, b);
      n += br;
      nv += bv;
      ++b;
    }
    h->ingest_read_s += now_s() - t_read;
    if (n == 0) continue;
    // ---- one call of the hot path per chunk (the arrays are pinned: DMA straight from them) ----
    const double t_call = now_s();
    int rc = 0;
    if (task == TRAIN) {
      rc = svdgpu_update_csr(h, (int)n, (const int *)rp.p, (const float *)lab.p, (const unsigned *)idx.p, (const float *)val.p);
    } else if (task == PREDICT) {
      if (rows_done + n > out_cap) return fail(h, "predict_buffer_file: output holds %lld rows, the file has more", out_cap);
      rc = svdgpu_predict_csr(h, (int)n, (const int *)rp.p, (const float *)lab.p, (const unsigned *)idx.p,
                              (const float *)val.p, out + rows_done);
    } else {
      double s = 0.0;
      long long c = 0;
      rc = svdgpu_eval_csr(h, (int)n, (const int *)rp.p, (const float *)lab.p, (const unsigned *)idx.p,
                           (const float *)val.p, scale, &s, &c);
      sse += s;
      cnt += c;
    }
    if (rc) return rc;
    h->ingest_call_s += now_s() - t_call;
    rows_done += n;
  }
  if (rows_out) *rows_out = rows_done;
  if (sum_sq) *sum_sq = sse;
  (void)cnt;
  return 0;
}

// One pass over a user-group buffer file.
int pass_ugroup(svdgpu *h, const char *path, Task task, float *out, long long out_cap, float scale,
                double *sum_sq, long long *rows_out) {
  File fi;
  fi.f = fopen(path, "rb");
  if (!fi.f) return fail(h, "can't open buffer %s", path);
  int hdr[4];  // num_batch, max_num_ufeedback, max_num_row, max_num_val
  if (!rd(fi.f, hdr, sizeof(hdr))) return fail(h, "%s: truncated header", path);
  std::vector<int> bro, bfo, tag, rp;
  std::vector<unsigned> fbi, idx;
  std::vector<float> fbv, lab, val;
  long long rows_done = 0;
  double sse = 0.0;
  int b = 0;
  bool open_run = false;
  while (b < hdr[0]) {
    bro.assign(1, 0);
    bfo.assign(1, 0);
    tag.clear();
    rp.assign(1, 0);
    fbi.clear(); fbv.clear(); lab.clear(); idx.clear(); val.clear();
    // whole units only: a START..END run is never split between two calls
    while (b < hdr[0] && (open_run || (long long)lab.size() < h->chunk_rows)) {
      int nfb = 0, t = 0;
      if (!rd(fi.f, &nfb, 4)) return fail(h, "%s: truncated block %d", path, b);
      if (nfb < 0) {
        nfb &= 0x7fffffff;
        if (!rd(fi.f, &t, 4)) return fail(h, "%s: truncated block %d", path, b);
      }
      const size_t f0 = fbi.size();
      fbi.resize(f0 + nfb);
      fbv.resize(f0 + nfb);
      int nn[2];
      if (!rd(fi.f, fbi.data() + f0, (size_t)nfb * 4) || !rd(fi.f, fbv.data() + f0, (size_t)nfb * 4) || !rd(fi.f, nn, 8))
        return fail(h, "%s: truncated block %d", path, b);
      const int br = nn[0], bv = nn[1];
      if (br < 0 || bv < 0) return fail(h, "%s: bad block %d", path, b);
      const size_t r0 = lab.size(), v0 = idx.size();
      rp.resize(3 * (r0 + br) + 1);
      lab.resize(r0 + br);
      idx.resize(v0 + bv);
      val.resize(v0 + bv);
      int first = 0;
      if (!rd(fi.f, &first, 4) || !rd(fi.f, rp.data() + 3 * r0 + 1, 3 * (size_t)br * 4) ||
          !rd(fi.f, lab.data() + r0, (size_t)br * 4) || !rd(fi.f, idx.data() + v0, (size_t)bv * 4) ||
          !rd(fi.f, val.data() + v0, (size_t)bv * 4))
        return fail(h, "%s: truncated block %d", path, b);
      for (size_t i = 3 * r0 + 1; i < rp.size(); ++i) rp[i] += (int)v0;
      bro.push_back((int)lab.size());
      bfo.push_back((int)fbi.size());
      tag.push_back(t);
      if (t == 1) open_run = true;       // START
      else if (t == 2) open_run = false; // END
      ++b;
    }
    const int nb = (int)tag.size();
    if (nb == 0) continue;
    const long long n = (long long)lab.size();
    int rc = 0;
    if (task == TRAIN) {
      rc = svdgpu_update_ugroup(h, nb, bro.data(), bfo.data(), tag.data(), fbi.data(), fbv.data(), rp.data(),
                                lab.data(), idx.data(), val.data());
    } else if (task == PREDICT) {
      if (rows_done + n > out_cap) return fail(h, "predict_buffer_file: output holds %lld rows, the file has more", out_cap);
      rc = svdgpu_predict_ugroup(h, nb, bro.data(), bfo.data(), tag.data(), fbi.data(), fbv.data(), rp.data(),
                                 lab.data(), idx.data(), val.data(), out + rows_done);
    } else {
      double s = 0.0;
      rc = svdgpu_eval_ugroup(h, nb, bro.data(), bfo.data(), tag.data(), fbi.data(), fbv.data(), rp.data(),
                              lab.data(), idx.data(), val.data(), scale, &s, nullptr);
      sse += s;
    }
    if (rc) return rc;
    rows_done += n;
  }
  if (rows_out) *rows_out = rows_done;
  if (sum_sq) *sum_sq = sse;
  return 0;
}

int pass(svdgpu *h, const char *path, Task task, float *out, long long out_cap, float scale, double *sum_sq,
         long long *rows_out) {
  if (!h || !path) return 1;
  return h->shape.format_type == 1 ? pass_ugroup(h, path, task, out, out_cap, scale, sum_sq, rows_out)
                                   : pass_csr(h, path, task, out, out_cap, scale, sum_sq, rows_out);
}

}  // namespace

extern "C" {

int svdgpu_update_buffer_file(svdgpu_t *h, const char *path, long long *num_row) {
  return pass(h, path, TRAIN, nullptr, 0, 1.0f, nullptr, num_row);
}
int svdgpu_predict_buffer_file(svdgpu_t *h, const char *path, float *out, long long out_cap, long long *num_row) {
  if (h && !out) return fail(h, "predict: null output");
  return pass(h, path, PREDICT, out, out_cap, 1.0f, nullptr, num_row);
}
int svdgpu_eval_buffer_file(svdgpu_t *h, const char *path, float scale, double *sum_sq, long long *num_row) {
  return pass(h, path, EVAL, nullptr, 0, scale, sum_sq, num_row);
}

}  // extern "C"
