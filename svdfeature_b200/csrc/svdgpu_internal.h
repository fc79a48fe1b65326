// svdgpu_internal.h -- handle layout and launcher entry points shared by the
// translation units of libsvdgpu.so (not part of the public ABI).
#pragma once
#include "../../include/svdgpu.h"
#include "svdgpu_device.cuh"

#include <algorithm>
#include <cstdio>
#include <string>
#include <vector>

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  // what a device-side fill left in the buffer (compact H2D of regular chunks, svdgpu_api.cu):
  // fill_n = 0 when the content came from a copy
  long long fill_n = 0;
  int fill_a = 0, fill_b = 0, fill_c = 0;
};
struct HostBuf {
  void *p = nullptr;
  size_t cap = 0;
};
// Ordered mode with item-owner warps (svdgpu_own.cu).  A plan = the per-owner queues of one set
// of rows; the scratch = what building a plan needs (shared by all plans of a handle: builds are
// stream-ordered on one stream).
struct OwnPlan {
  DevBuf entries, queue_off, item_off, items, batch;
  HostBuf stage;  // pinned source of the small host-made arrays
  int num_owner = 0;
  int per_cta = 0;  // owners per CTA the plan was dealt for (12: k_own, 8: k_own2)
  long long rows = 0, max_load = 0;
  bool valid = false;
};
namespace svdown {
struct HostPlan;
}
struct OwnScratch {
  DevBuf cnt_item, cnt_user, start_user, flag, keyA, keyB, valA, valB, key_item, tick, tmp, item_owner, item_slot, stats;
  int stats_owners = 0;   // owners of the launch the stats buffer describes
  void *h_cnt = nullptr;  // pinned: item counts + flag word
  size_t h_cnt_cap = 0;
  cudaEvent_t ev = nullptr;
  svdown::HostPlan *deal = nullptr;  // the last deal of items to owners (svdgpu_ownplan.h), carried over between plans
  int deal_ctas = 0;                 // ... and what it was made for
  long long deal_key = 0;
};

// One staging slot for host-pointer calls: pinned mirrors + device arrays.
struct Slot {
  HostBuf h_rp, h_label, h_index, h_value, h_value2, h_ticket, h_misc, h_fbi, h_fbv, h_fbt;
  OwnPlan own;  // ordered mode: the chunk's owner plan
  DevBuf d_rp, d_label, d_index, d_value, d_value2, d_ticket, d_misc, d_fbi, d_fbv, d_fbt, d_pred;
  cudaEvent_t done = nullptr;    // last kernel that read this slot
  cudaEvent_t copied = nullptr;  // the slot's H2D copies have landed (recorded on the copy stream)
  cudaEvent_t planned = nullptr;  // ordered mode: the slot's arrays and plan are complete (recorded on the plan stream)
  bool used = false;
};

namespace svdk {
struct DeltaSeg {
  float *cur;     // live slab segment
  long long n;    // floats
  long long off;  // offset in the packed buffers
};
struct DeltaPlan {
  DeltaSeg seg[5];
  int nseg;
  long long total;
};
// A "unit" is one user's consecutive blocks: a DEFAULT block, or START..END
// (apex_svd_data.h:353-371).  unit_off[u]..unit_off[u+1] are its blocks.
struct DevUgroup {
  const int *unit_off;     // [num_unit+1] block ranges
  const int *blk_row_off;  // [num_block+1]
  const int *blk_fb_off;   // [num_block+1]
  const unsigned *fb_index;
  const float *fb_value;
  const unsigned *fb_ticket;  // ordered mode
  const int *order;           // hogwild: unit processing order (longest first), may be null
  int row_base;               // blk_row_off values are absolute rows; csr arrays start at row_base
  int fb_base;                // blk_fb_off values are absolute; fb arrays start at fb_base
  int has_fb;                 // any feedback entry in the launch (SVD++): bounds the Hogwild concurrency
};
struct Geometry {
  int lanes, vec;
};
}  // namespace svdk

struct svdgpu_rank_state;  // svdgpu_rank.cu
struct svdgpu_comm;        // svdgpu_comm.cu: NCCL communicator of a handle

struct svdgpu {
  svdgpu_shape shape;
  svdgpu_hparams hp;
  bool hp_set = false;
  int device = 0;
  int num_sm = 0;
  int mode = SVDGPU_MODE_HOGWILD;
  int scatter_user = svdk::SCATTER_RED, scatter_item = svdk::SCATTER_RED;
  int exact_dot = 0;  // Hogwild default: tree-order dot (the ordered mode always uses the reference order)
  int lanes_opt = 0;
  int chunk_rows = 1 << 20;
  int ctas_per_sm = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  // host-pointer calls stage chunk c+1 on this stream while the kernels of chunk c run on `stream`
  cudaStream_t copy_stream = nullptr;
  // ordered host-pointer calls build the plan of chunk c+1 here while k_own trains chunk c on `stream`
  cudaStream_t plan_stream = nullptr;
  cudaEvent_t ev_plan = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_copy = nullptr;
  svdk::DevModel dm;
  svdk::DevHP dhp{};
  size_t rows = 0;
  int *d_err = nullptr;
  unsigned *d_counter = nullptr;
  double *d_eval = nullptr;  // {sum of squared errors, count} of an svdgpu_eval_* call
  bool eval_on = false;
  float eval_scale = 1.0f;
  unsigned *d_row_mask = nullptr;  // Hogwild: rows the fast pass left to the generic pass (1 bit per row)
  size_t row_mask_cap = 0;
  size_t flag_for_generic = 0;  // which of the three flag words gates the generic pass of this launch
  size_t any_left_at = 0;  // index of the LAST of the three "a fast pass left something" words inside d_row_mask
  int pass1 = 3;  // option "pass1": 0 sends every row through the generic pass, 1 first fast pass only,
                  // 2 first and second fast pass (rows with two item features), 3 (default) also the third
                  // (basic rows with up to 16 global features)

  int svdpp_fast = 1;    // option "svdpp_fast": 0 keeps every user unit in k_ugroup
  unsigned char *d_unit_kind = nullptr;  // per unit of a launch: 1 = taken by k_svdpp
  size_t unit_kind_cap = 0;
  int ugroup_units = 0;  // option "ugroup_units": user units in flight in Hogwild user-group training (0 = auto: 64 with
                         // feedback lists -- more diverges, tools/hogwild_parity.py --svdpp -- else the occupancy limit)
  int l2_ahead = -1;   // option "l2_ahead": generic pass prefetches the next tile's rows into L2 (-1 auto)
  int mfg = 1;         // option "mfg": third fast pass variant: 0 three stages / ring depth 4 / one CTA per SM
                       // (configs[4]: 0.79 G inst/s), 1 (default) two stages with late window requests / ring
                       // depth 2 / two CTAs per SM (1.17 G inst/s: the pass is short of warps, not of bytes in flight)
  int stream_tile = 0; // option "stream_tile": rows per tile of the generic pass (0 = auto: 64 / 32 / 16 / 8 by row width)
  int ring_depth = 0;  // option "ring_depth": k_mf ring depth (0 = default 4)
  int compact_h2d = 1;  // option "compact_h2d" (1: Hogwild / predict, 2: ordered host-pointer calls too): they do not copy a chunk's
                        // row_ptr when every row has the same feature counts, nor its values when all are 1.0f
                        // (checked on host threads while earlier chunks are copied; rebuilt on the device)
  int scan_threads = 0;            // option "scan_threads": host threads of that check (0 = min(cores, 16))
  int compact_min_rows = 1 << 18;  // option "compact_min_rows": calls with fewer rows skip the check
  int exact_owner = 1;  // option "exact_owner": ordered mode sends basic-MF rows under plain L2 decay through the
                        // item-owner kernel (k_own, svdgpu_own.cu); 0 keeps every row in k_exact
  int own_min_rows = 4096;   // option "own_min_rows": launches with fewer rows keep k_exact (no plan to build)
  int own_urgent_gap = 4096;   // option "own_urgent_gap": a user whose next rating follows within this many rows
                               // is published at once instead of with the owner's batch
  int own_batch = 24;        // option "own_batch": version publishes the most loaded owner holds back (<= 32)
  int own_slots = 0;         // option "own_slots": item rows an owner keeps in shared memory (0 = auto, <= 32)
  int own_depth = 8;         // option "own_depth": ring slots per owner (8 or 16)
  int own_fast = 1;          // option "own_fast": 0 keeps the generic link for every shape (testing)
  int own_acquire = 0;       // option "own_acquire": loaders poll versions with ld.acquire.gpu (adds CCTL.IVALL per poll)
  int own_isolate = 200;     // option "own_isolate": owners of items above this % of the mean owner load get an issue port to themselves (0: off)
  int own_plan_beside = 1;   // option "own_plan_beside": host-pointer calls build the plan of chunk c+1 while k_own trains chunk c (0: after it; measured slower)
  int own_redeal = 25;       // option "own_redeal": a plan keeps the previous plan's deal of items to owners while its heaviest owner is within this % of the best possible (0: deal every time)
  int own_isolate_full = 75; // option "own_isolate_full": ... and those above this % of the hottest item's count a whole SM (0: off)
  int own_reverse = 1;       // option "own_reverse": busiest owners on the highest warp ids (the arbiter prefers them)
  int own_spare_sms = 8;     // option "own_spare_sms": SMs an ordered host-pointer call leaves to the plan kernels and
                             // fills of the next chunk (k_own then runs on num_sm - this many CTAs)
  int own_poll_ns = 50;      // option "own_poll_ns": loader warps sleep this long between polls without progress
  int own_partner = 0;       // option "own_partner": 1 = split link (k_own2: owner + partner warp) where the shape allows.
                             // Measured, 20 M rows of configs[1]: 22.4 ms against k_own's 19.9 ms (the partner releases the
                             // ring slots later, the hot owners wait for them; alone on the GPU a link takes 0.304 vs 0.316 us)
  int own_stats = 0;         // option "own_stats": k_own records per-owner cycle counters (svdgpu_own_stats)
  OwnScratch own;
  unsigned *d_abort = nullptr;  // k_own: set when a wait timed out, every warp leaves
  int exact_opt = 5;   // option "exact_opt": k_exact hand-off variants (bit mask, svdgpu_ordered.cu):
                       // 1 no per-lane fence before the release, 2 spin before sleeping, 4 staged slice
  // Hogwild stability guard: N instances in flight that touch a row with probability p apply ~N*p stale
  // steps of size lr to it at once; beyond N*p*lr ~ 2 asynchronous SGD on that row diverges (measured: NaN at
  // k = 32, lr = 0.01, 3000 items -- tools/hogwild_stability.py).  Training launches are capped at
  // inflight_cap = hog_safety / (lr * share of the hottest item), 0 = no cap.
  long long inflight_cap = 0;
  int hog_safety_permille = 1000;  // option "hogwild_safety" (per mille; 0 disables the guard)
  unsigned *d_hist = nullptr;      // item histogram + max word of the guard
  size_t hist_cap = 0;
  double hot_frac = -1.0;          // share of the hottest item the current cap was derived from
  double hot_call = -1.0;          // ... as last measured on a host-pointer call (-1: never)
  unsigned *h_hot = nullptr;       // pinned {count of the hottest item, rows} of the measurement in flight
  cudaEvent_t ev_hot = nullptr;
  int hot_pending = 0;
  int mf_ctas = 0;     // option "mf_ctas": k_mf CTAs per SM the register allocation aims at (0 = default 2)
  static constexpr int NSLOT = 3;
  Slot slot[NSLOT];
  int cur_slot = 0;
  // bulk ingest: pinned chunk buffers (kept across calls) and where the time went
  HostBuf ing_rp, ing_label, ing_index, ing_value;
  double ingest_read_s = 0.0, ingest_call_s = 0.0;
  // side features (feature_user / feature_item, base.h:98-99): index -> extra (index, value) pairs
  struct Side {
    std::vector<unsigned> rp, idx;
    std::vector<float> val;
    bool on() const { return rp.size() > 1; }
  } side_u, side_i;
  // SVDFeatureRanker state (svdgpu_rank_init)
  svdgpu_rank_state *rank = nullptr;
  int rank_force_sort = 0;  // option "rank_force_sort": top-k by the radix sort even for small top_k (testing)
  // multi-GPU exchange
  svdgpu_comm *comm = nullptr;
  long long n_coll = 0, n_coll_bytes = 0;  // NCCL collectives issued and the bytes they reduced
  svdk::DeltaPlan plan;
  float *d_snap = nullptr, *d_delta = nullptr;
  // host scratch for tickets
  std::vector<unsigned> cnt_ui, cnt_g;
  std::string err;
  long long n_launch = 0, n_inst = 0, n_h2d = 0, n_d2h = 0;
  long long n_own = 0, n_own_rows = 0;  // k_own launches and the rows they trained
  long long n_deal = 0, n_redeal = 0;  // plans that dealt the items out anew / carried the deal over
  long long own_lpt_us = 0, own_cntwait_us = 0;  // plan: host time dealing items out / waiting for the item counts
};

// an instance batch resident in HBM (svdgpu_batch_create / svdgpu_batch_sample_pairs)
struct svdgpu_batch {
  int num_row = 0;
  long long num_val = 0;
  DevBuf d_rp, d_label, d_index, d_value, d_value2, d_ticket, d_pred;
  bool has_ticket = false, has_value2 = false;
  double hot_frac = -1.0;  // share of the hottest item among the batch's rows (Hogwild guard), -1: not measured yet
  OwnPlan own;  // ordered mode through k_own: the batch's owner plan (valid when the rows qualify)
  // user-group structure
  bool ugroup = false, has_fb = false;
  int num_block = 0, num_unit = 0;
  DevBuf d_unit_off, d_blk_row_off, d_blk_fb_off, d_fbi, d_fbv, d_fbt, d_order;
  std::vector<int> unit_off;  // host copy: block range of each unit
};

namespace svdk {
int fail(svdgpu *h, const char *fmt, ...);

#define CU(h, call)                                                                                 \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess)                                                                          \
      return svdk::fail(h, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

template <typename K>
int grid_for(svdgpu *h, K kernel, int threads, long long work_items, int *grid, size_t dyn_smem = 0) {
  int per_sm = 0;
  CU(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem));
  if (per_sm < 1) per_sm = 1;
  if (h->ctas_per_sm > 0) per_sm = std::min(per_sm, h->ctas_per_sm);
  long long gmax = (long long)per_sm * h->num_sm;
  *grid = (int)std::max(1LL, std::min(gmax, work_items));
  return 0;
}

// launchers (one translation unit each, so nvcc runs in parallel)
int launch_stream(svdgpu *h, const Geometry &g, const DevCsr &csr, int r0, int r1, bool train, float *pred);
int launch_mf(svdgpu *h, const Geometry &g, const DevCsr &csr, int r0, int r1, bool train, float *pred,
              int which, unsigned *flag_out, const unsigned *flag_gate);
int launch_exact(svdgpu *h, const Geometry &g, const DevCsr &csr, int r0, int r1);
// ordered mode with item-owner warps (svdgpu_own.cu)
bool own_supported(const svdgpu *h);
int own_owners_per_cta(const svdgpu *h);
// ctas: CTAs (= SMs) the launch will occupy, 12 owners each; 0 = every SM
int own_plan_build(svdgpu *h, const DevCsr &csr, int r0, int n, OwnPlan &p, cudaStream_t st, int *bad, int ctas = 0);
int launch_own(svdgpu *h, const OwnPlan &p, cudaStream_t st);
void own_plan_free(OwnPlan &p);
void own_scratch_free(OwnScratch &s);
int launch_ugroup(svdgpu *h, const Geometry &g, const DevCsr &csr, const DevUgroup &ug, int u0, int u1,
                  bool train, bool ordered, float *pred);
int launch_delta(svdgpu *h, int mode, float scale);
void rank_free(svdgpu *h);
int launch_svdpp(svdgpu *h, const DevCsr &csr, const DevUgroup &ug, int u0, int u1, int warps,
                 const unsigned char **kind_out);
}  // namespace svdk
