// svdgpu_ownplan.h -- host part of the plan for the ordered mode with item-owner warps
// (k_own, svdgpu_own.cu).  Plain C++: unit-tested on the CPU
// (tests/test_own_plan.py) and compiled into libsvdgpu.so.
//
// The reference's SGD loop is strictly sequential (base.h:456-462).  Two instances that share
// neither a user row nor an item row commute exactly, so the ordered mode only has to keep, for
// every row, the order of the instances that touch it.  k_own does this by giving every ITEM to
// one persistent warp (its "owner"), which takes the instances of its items in input order and
// keeps the item rows on chip; user rows travel between owners through version counters.
//
// The device counts the ratings of every item; this header deals the items out: in decreasing
// popularity to the least loaded owner (LPT), so the heaviest owner carries little more than the
// hottest item -- whose chain of dependent updates is the critical path of the whole launch.
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <queue>
#include <utility>
#include <vector>

namespace svdown {

struct HostPlan {
  int num_owner = 0;
  std::vector<int> item_owner;       // [num_item] owner of an item, -1 if the batch never touches it
  std::vector<unsigned> item_slot;   // [num_item] position of the item in its owner's list
  std::vector<int> queue_off;        // [num_owner + 1] first queue entry of an owner
  std::vector<int> item_off;         // [num_owner + 1] first entry of an owner in `items`
  std::vector<unsigned> items;       // items by owner, most popular first (slot order)
  std::vector<int> batch;            // [num_owner] user-row publishes an owner may hold back
  int64_t max_load = 0;
  int num_open = 0;                  // owners that take items (num_owner - closed ones)
  int age = 0;                       // redeal() calls since the deal was made
};

// cnt[i] = ratings of item i in the batch.  Owner w of the launch is warp (w / grid) of block
// (w % grid): consecutive owners sit on different SMs, so the hottest items never share one.
// max_batch: how many user-row version publishes the most loaded owner may collect before one
// release fence; slower owners hold back proportionally fewer, so the delay is about the same
// stretch of the input order for everybody.
// closed (optional, [num_owner]): owners that get no items -- the warps that would share an issue port with
// the owner of a very hot item (hot_owners below; the mapping owner -> warp is the kernel's, so the caller
// marks them).
// deal_all: items the batch never touches get an owner too (round robin over the lighter owners, behind the
// owner's popular items), so that redeal() can carry the deal over to a batch that does touch them.
inline void assign(const unsigned *cnt, int num_item, int num_owner, int max_batch, HostPlan &p,
                   const std::vector<char> *closed = nullptr, bool deal_all = false) {
  p.num_owner = num_owner;
  p.age = 0;
  p.item_owner.assign((size_t)num_item, -1);
  p.item_slot.assign((size_t)num_item, 0u);
  // items by decreasing count, ties by increasing index: a stable LSD radix sort (3 passes of 11 bits) of the
  // touched items on the key ~count
  std::vector<int> order, tmp_order;
  order.reserve((size_t)num_item);
  for (int i = 0; i < num_item; ++i)
    if (cnt[i] > 0) order.push_back(i);
  {
    tmp_order.resize(order.size());
    std::vector<unsigned> hist(2048);
    for (int pass = 0; pass < 3; ++pass) {
      const int shift = 11 * pass;
      std::fill(hist.begin(), hist.end(), 0u);
      for (int i : order) ++hist[(~cnt[i] >> shift) & 2047u];
      unsigned run = 0;
      for (unsigned &hv : hist) {
        const unsigned c = hv;
        hv = run;
        run += c;
      }
      for (int i : order) tmp_order[hist[(~cnt[i] >> shift) & 2047u]++] = i;
      order.swap(tmp_order);
    }
  }
  // binary min-heap of (rows, owner) packed into one word (rows << 32 | owner: smallest load first, then
  // smallest owner id; a plan covers fewer than 2^31 rows); an item goes to the top, whose load grows, and the
  // top sinks back into place (one sift instead of a pop and a push).
  std::vector<int> open;
  for (int w = 0; w < num_owner; ++w)
    if (!closed || !(*closed)[(size_t)w]) open.push_back(w);
  p.num_open = (int)open.size();
  std::vector<int64_t> load((size_t)num_owner, 0);
  std::vector<int> nitem((size_t)num_owner, 0);
  std::vector<uint64_t> heap;
  heap.reserve(open.size());
  for (int w : open) heap.push_back((uint64_t)(unsigned)w);  // (ascending owner ids with equal loads: already a heap)
  const size_t hn = heap.size();
  for (int i : order) {
    if (!hn) break;
    const int w = (int)(heap[0] & 0xffffffffu);
    p.item_owner[(size_t)i] = w;
    p.item_slot[(size_t)i] = (unsigned)nitem[(size_t)w]++;
    load[(size_t)w] += cnt[i];
    const uint64_t l = ((uint64_t)load[(size_t)w] << 32) | (uint64_t)(unsigned)w;
    size_t at = 0;
    for (;;) {
      size_t c = 2 * at + 1;
      if (c >= hn) break;
      if (c + 1 < hn && heap[c + 1] < heap[c]) ++c;
      if (heap[c] >= l) break;
      heap[at] = heap[c];
      at = c;
    }
    heap[at] = l;
  }
  if (deal_all && !open.empty()) {
    // (to owners that carry no more than the mean: the owner of a hot item keeps its chain to itself)
    int64_t total = 0;
    for (int w : open) total += load[(size_t)w];
    std::vector<int> light;
    for (int w : open)
      if (load[(size_t)w] * (int64_t)open.size() <= total) light.push_back(w);
    if (!light.empty()) open.swap(light);
    size_t rr = 0;
    for (int i = 0; i < num_item; ++i)
      if (cnt[i] == 0) {
        const int w = open[rr++ % open.size()];
        p.item_owner[(size_t)i] = w;
        p.item_slot[(size_t)i] = (unsigned)nitem[(size_t)w]++;
        order.push_back(i);
      }
  }
  p.queue_off.assign((size_t)num_owner + 1, 0);
  p.item_off.assign((size_t)num_owner + 1, 0);
  p.max_load = 0;
  for (int w = 0; w < num_owner; ++w) {
    p.queue_off[(size_t)w + 1] = p.queue_off[(size_t)w] + (int)load[(size_t)w];
    p.item_off[(size_t)w + 1] = p.item_off[(size_t)w] + nitem[(size_t)w];
    p.max_load = std::max(p.max_load, load[(size_t)w]);
  }
  p.items.assign(order.size(), 0u);
  for (int i : order)
    p.items[(size_t)p.item_off[(size_t)p.item_owner[(size_t)i]] + p.item_slot[(size_t)i]] = (unsigned)i;
  p.batch.assign((size_t)num_owner, 1);
  for (int w = 0; w < num_owner; ++w) {
    const int64_t b = p.max_load > 0 ? (int64_t)max_batch * load[(size_t)w] / p.max_load : 1;
    p.batch[(size_t)w] = (int)std::max<int64_t>(1, std::min<int64_t>(b, max_batch));
  }
}

// Carry the deal of an earlier batch over to new counts: same owners, same item lists; only the queue
// offsets, the loads and the batch sizes are recomputed (microseconds instead of the milliseconds of a new
// deal -- a host-pointer call plans every chunk).  Any deal is a CORRECT one; this one is kept as long as it is
// a good one: every touched item has an owner, and the heaviest owner carries no more than the hottest item
// alone, or slack_percent % above the mean of the open owners.  false: deal again.
inline bool redeal(const unsigned *cnt, int num_item, int max_batch, int slack_percent, HostPlan &p) {
  const int num_owner = p.num_owner;
  if (num_owner <= 0 || p.num_open <= 0 || (int)p.item_owner.size() != num_item) return false;
  std::vector<int64_t> load((size_t)num_owner, 0);
  int64_t total = 0;
  unsigned mx = 0;
  for (int i = 0; i < num_item; ++i) {
    const unsigned c = cnt[i];
    if (!c) continue;
    const int w = p.item_owner[(size_t)i];
    if (w < 0) return false;
    load[(size_t)w] += c;
    total += c;
    mx = std::max(mx, c);
  }
  const int64_t max_load = *std::max_element(load.begin(), load.end());
  const int64_t mean = (total + p.num_open - 1) / p.num_open;
  if (max_load > std::max<int64_t>(mx, mean * (100 + slack_percent) / 100)) return false;
  p.max_load = max_load;
  for (int w = 0; w < num_owner; ++w) {
    p.queue_off[(size_t)w + 1] = p.queue_off[(size_t)w] + (int)load[(size_t)w];
    const int64_t b = max_load > 0 ? (int64_t)max_batch * load[(size_t)w] / max_load : 1;
    p.batch[(size_t)w] = (int)std::max<int64_t>(1, std::min<int64_t>(b, max_batch));
  }
  ++p.age;
  return true;
}

// How many of the most popular items carry more than `percent` % of an owner's mean load each: LPT gives
// the r-th most popular item to owner r, so these are owners 0 .. hot_owners-1, with one item each, and
// their chains are what the launch waits for.
inline int hot_owners(const unsigned *cnt, int num_item, int num_owner, int percent) {
  int64_t total = 0;
  for (int i = 0; i < num_item; ++i) total += cnt[i];
  if (total == 0 || num_owner <= 0) return 0;
  int hot = 0;
  for (int i = 0; i < num_item; ++i)
    if ((int64_t)cnt[i] * num_owner * 100 > total * percent) ++hot;
  return std::min(hot, num_owner);
}

// ... and how many items carry at least `percent` % of the most popular item's count: the chains within
// reach of the longest one.
inline int top_owners(const unsigned *cnt, int num_item, int percent) {
  unsigned mx = 0;
  for (int i = 0; i < num_item; ++i) mx = std::max(mx, cnt[i]);
  int top = 0;
  for (int i = 0; i < num_item; ++i)
    if (cnt[i] > 0 && (uint64_t)cnt[i] * 100u >= (uint64_t)mx * (unsigned)percent) ++top;
  return top;
}

}  // namespace svdown
