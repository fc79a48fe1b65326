// svdgpu_scan.h -- host-side checks behind the compact H2D of the host-pointer calls
// (svdgpu_api.cu): is a chunk's row_ptr the arithmetic progression of rows with constant
// feature counts, are its values all 1.0f?  Exact (every element is compared), written so that
// the host compiler vectorises the loops; plain C++, no CUDA, unit-tested on the CPU
// (tests/test_scan_host.py).
#pragma once
#include <algorithm>
#include <cstring>

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

}  // namespace svdscan
