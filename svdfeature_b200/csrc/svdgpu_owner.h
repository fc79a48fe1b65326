// svdgpu_owner.h -- host side of the ordered mode with item-owner warps (EXPERIMENTAL, option
// "exact_owner", off by default).  Plain C++, unit-tested on the CPU
// (tests/test_owner_plan.py).
//
// The ordered kernel k_exact hands every touched row from one warp to the next through L2; the
// chain of the hottest item row bounds it.  Here every item belongs to ONE persistent warp, which
// takes the instances of its items in input order; only the user rows still travel between
// warps (tickets, as before).  This header deals the items out and builds the per-owner queues.
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <queue>
#include <utility>
#include <vector>

namespace svdowner {

struct Plan {
  std::vector<int> queue_off;  // [num_owner + 1]
  std::vector<int> queue;      // [num_row] row ids, input order inside an owner
};

// Rows [0, num_row) must all have the basic-MF shape (0 | 1 | 1) with item index < num_item:
// returns false otherwise (the caller keeps k_exact).  Items are dealt out in decreasing
// popularity to the least loaded owner (LPT), so the heaviest owner carries little more than the
// hottest item.
inline bool build_plan(int num_row, const int *row_ptr, const unsigned *index, int num_item, int num_owner,
                       Plan &plan) {
  if (num_row <= 0 || num_owner <= 0 || num_item <= 0) return false;
  std::vector<int64_t> cnt((size_t)num_item, 0);
  for (int r = 0; r < num_row; ++r) {
    const int *p = row_ptr + 3LL * r;
    if (p[1] != p[0] || p[2] != p[1] + 1 || p[3] != p[2] + 1) return false;
    const unsigned iid = index[p[2]];
    if (iid >= (unsigned)num_item) return false;
    cnt[iid]++;
  }
  std::vector<int> items;
  for (int i = 0; i < num_item; ++i)
    if (cnt[(size_t)i] > 0) items.push_back(i);
  std::stable_sort(items.begin(), items.end(), [&](int a, int b) { return cnt[(size_t)a] > cnt[(size_t)b]; });
  typedef std::pair<int64_t, int> Load;  // (rows, owner): smallest load first, then smallest owner id
  std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
  for (int w = 0; w < num_owner; ++w) heap.push(Load(0, w));
  std::vector<int> owner((size_t)num_item, -1);
  std::vector<int64_t> load((size_t)num_owner, 0);
  for (int i : items) {
    Load l = heap.top();
    heap.pop();
    owner[(size_t)i] = l.second;
    l.first += cnt[(size_t)i];
    load[(size_t)l.second] = l.first;
    heap.push(l);
  }
  plan.queue_off.assign((size_t)num_owner + 1, 0);
  for (int w = 0; w < num_owner; ++w) plan.queue_off[(size_t)w + 1] = plan.queue_off[(size_t)w] + (int)load[(size_t)w];
  std::vector<int> cur(plan.queue_off.begin(), plan.queue_off.end() - 1);
  plan.queue.resize((size_t)num_row);
  for (int r = 0; r < num_row; ++r) {
    const int w = owner[(size_t)index[row_ptr[3LL * r + 2]]];
    plan.queue[(size_t)cur[(size_t)w]++] = r;
  }
  return true;
}

}  // namespace svdowner
