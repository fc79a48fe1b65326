// svdgpu_scan.h -- host-side checks behind the compact H2D of the host-pointer calls
// (svdgpu_api.cu): is a chunk's row_ptr the arithmetic progression of rows with constant
// feature counts, are its values all 1.0f?  Exact (every element is compared), written so that
// the host compiler vectorises the loops; plain C++, no CUDA, unit-tested on the CPU
// (tests/test_scan_host.py).
#pragma once
#include <algorithm>
#include <atomic>
#include <cstring>
#include <system_error>
#include <thread>
#include <vector>

namespace svdscan {

// rows [0,n) of p (entries p[0..3n], SVDFeatureCSR::row_ptr, apex_svd_data.h:129-142) all have the
// feature counts (a | b | c) of row 0
inline bool rp_regular(const int *p, long long n, int &a, int &b, int &c) {
  if (n <= 0) return false;
  const long long v0 = p[0];
  a = p[1] - p[0];
  b = p[2] - p[1];
  c = p[3] - p[2];
  if (v0 < 0 || a < 0 || b < 0 || c < 0) return false;
  const long long w = (long long)a + b + c;
  if (v0 + n * w != (long long)p[3 * n]) return false;  // (so base + offsets below never overflow int)
  // four rows = twelve entries per step against a running expectation
  int expect[12];
  for (int r = 0; r < 4; ++r) {
    const int base = (int)(v0 + r * w);
    expect[3 * r] = base;
    expect[3 * r + 1] = base + a;
    expect[3 * r + 2] = base + a + b;
  }
  const int inc = (int)(4 * w);
  const long long n4 = n & ~3LL;
  constexpr long long BLK = 1024;  // rows between two early-exit tests
  long long r = 0;
  while (r < n4) {
    const long long r1 = std::min(n4, r + BLK);
    unsigned diff[12] = {0};
    for (; r < r1; r += 4) {
      const int *q = p + 3 * r;
      for (int j = 0; j < 12; ++j) {
        diff[j] |= (unsigned)(q[j] ^ expect[j]);
        expect[j] = (int)((unsigned)expect[j] + (unsigned)inc);  // (wraps harmlessly past the last block)
      }
    }
    unsigned any = 0;
    for (int j = 0; j < 12; ++j) any |= diff[j];
    if (any) return false;
  }
  for (; r < n; ++r) {
    const int base = (int)(v0 + r * w);
    const int *q = p + 3 * r;
    if (q[0] != base || q[1] != base + a || q[2] != base + a + b) return false;
  }
  return true;
}

// every one of the nv values is the bit pattern of 1.0f
inline bool all_ones(const float *v, long long nv) {
  constexpr long long BLK = 8192;
  for (long long i0 = 0; i0 < nv; i0 += BLK) {
    const long long i1 = std::min(nv, i0 + BLK);
    unsigned diff = 0;
    for (long long i = i0; i < i1; ++i) {
      unsigned u;
      std::memcpy(&u, v + i, 4);
      diff |= u ^ 0x3f800000u;
    }
    if (diff) return false;
  }
  return true;
}

// verdict of the two checks for one chunk of a call
struct ChunkScan {
  bool rp_regular = false, val_ones = false;
  int a = 0, b = 0, c = 0;  // global / user / item features per row when rp_regular
};

// the scanning threads of one call: chunk c's verdict is in scan[c] once ready[c] != 0.  Chunks are
// claimed in input order, so the verdicts the caller waits for come first.
struct ScanPool {
  std::vector<ChunkScan> scan;
  std::vector<std::atomic<int>> ready;
  std::atomic<int> next{0};
  std::atomic<bool> stop{false};
  std::vector<std::thread> threads;
  int num_row = 0, chunk_rows = 1;
  const int *row_ptr = nullptr;
  const float *value = nullptr;
  explicit ScanPool(int nchunk) : scan((size_t)nchunk), ready((size_t)nchunk) {
    for (auto &r : ready) r.store(0, std::memory_order_relaxed);
  }
  // claim the next unclaimed chunk and scan it; false when none is left
  bool scan_one() {
    const int c = next.fetch_add(1);
    if (c >= (int)scan.size()) return false;
    const long long r0 = (long long)c * chunk_rows, r1 = std::min<long long>(num_row, r0 + chunk_rows);
    ChunkScan &s = scan[(size_t)c];
    const long long v0 = row_ptr[3 * r0], v1 = row_ptr[3 * r1];
    s.rp_regular = rp_regular(row_ptr + 3 * r0, r1 - r0, s.a, s.b, s.c);
    s.val_ones = value && v0 >= 0 && v1 > v0 && all_ones(value + v0, v1 - v0);
    ready[(size_t)c].store(1, std::memory_order_release);
    return true;
  }
  void start(int num_row_, int chunk_rows_, const int *row_ptr_, const float *value_, int max_threads) {
    num_row = num_row_;
    chunk_rows = chunk_rows_;
    row_ptr = row_ptr_;
    value = value_;
    const int nt = std::max(0, std::min(max_threads, (int)scan.size()));  // 0: the caller scans in wait()
    try {
      for (int t = 0; t < nt; ++t)
        threads.emplace_back([this]() {
          while (!stop.load(std::memory_order_relaxed) && scan_one()) {
          }
        });
    } catch (const std::system_error &) {
      // out of threads: the ones that started (or, with none, the calling thread in wait()) do the work
    }
  }
  const ChunkScan &wait(int c) {
    while (!ready[(size_t)c].load(std::memory_order_acquire)) {
      if (threads.empty()) scan_one();  // (claims are in order: this reaches chunk c)
      else std::this_thread::yield();
    }
    return scan[(size_t)c];
  }
  ~ScanPool() {
    stop.store(true);
    for (auto &t : threads) t.join();
  }
};

}  // namespace svdscan
