/* svdgpu.h -- C ABI of the B200 SGD trainer for SVDFeature's hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.
 * It is what a reference-side ISVDTrainer implementation binds (see
 * svdfeature_b200/csrc/gpu_trainer.cpp and INTEGRATION.md).  All reference
 * citations are relative to the reference root; "base.h" is
 * solvers/base-solver/apex_svd_base.h, "model.h" is apex_svd_model.h.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     available from svdgpu_last_error().  (The reference's own convention --
 *     print + exit(-1), apex-utils/apex_utils.h:47-50 -- is re-established by
 *     the C++ trainer on top of this ABI.)
 *   - host pointers are borrowed for the duration of the call only.
 *   - one handle = one CUDA device = one host thread at a time (the reference's
 *     trainer is single-threaded too, svd_feature.cpp:231-247).
 *   - there is NO CPU fallback: without a CUDA device svdgpu_create fails.
 */
#ifndef SVDGPU_H_
#define SVDGPU_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct svdgpu svdgpu_t;             /* one trainer on one GPU            */
typedef struct svdgpu_batch svdgpu_batch_t; /* an instance batch resident in HBM */

/* Model shape: the fields of SVDModelParam (model.h:373-450) and SVDTypeParam
 * (model.h:242-261) the hot path depends on. */
typedef struct {
  int num_user;
  int num_item;
  int num_global;
  int num_ufeedback; /* rows of W_ufeedback; used only when format_type == 1 */
  int num_factor;    /* k */
  int no_user_bias;
  int active_type; /* model.h:61-79: 0 linear, 1 sigmoid-L2, 2 sigmoid-likelihood,
                      3 sigmoid-rank, 5 smooth hinge, 6 hinge-L2, 7 = 3          */
  int format_type; /* model.h:50-57: 0 random order, 1 user grouped (SVD++)     */
} svdgpu_shape;

/* Training hyper-parameters: SVDTrainParam (model.h:291-344) + base_score
 * (already passed through calc_base_score, model.h:220-237, as the model file
 * stores it). */
typedef struct {
  float learning_rate;
  float wd_user, wd_item;
  float wd_user_bias, wd_item_bias;
  float wd_global;
  int reg_method;  /* base.h:211-283: 0 L2 decay, 1 L1 soft threshold, 2 projection onto
                      |w|^2 <= wd, 3 L1 on user rows + L2 decay on item rows.  The lazy
                      variants 4/5 are rejected (broken in the reference: base.h:195,226) */
  int reg_global;  /* base.h:188-210: 0 L2 decay, 1 L1 (4/5 rejected)              */
  unsigned num_regfree_global;
  float scale_lr_ufeedback, wd_ufeedback, wd_ufeedback_bias; /* base.h:512-520 */
  float base_score;
  int user_nonnegative; /* SVDModelParam::user_nonnegative: clamp user rows at 0 after
                           every update (base.h:242-245)                            */
} svdgpu_hparams;

/* How instances of one call are ordered against each other. */
enum {
  /* Ordered data-flow: every row (user/item/global/feedback) carries a version
   * counter and each instance waits for exactly the writers that precede it in
   * input order.  Result is bit-identical to the reference's sequential loop
   * (base.h:456-462) for active_type 0/5/6; 1-ulp-of-expf close for sigmoid types. */
  SVDGPU_MODE_EXACT = 0,
  /* Hogwild: one lane group per instance, no ordering between instances of a
   * call; per-instance arithmetic is unchanged.  Throughput mode. */
  SVDGPU_MODE_HOGWILD = 1
};

/* ---- lifetime ---------------------------------------------------------- */
/* replaces: SVDFeature ctor + SVDModel::alloc_space (base.h:102-109, model.h:511-556) */
int svdgpu_create(svdgpu_t **out, const svdgpu_shape *shape, int device);
void svdgpu_destroy(svdgpu_t *h);
/* message of the last failure on this handle (h may be NULL for create failures) */
const char *svdgpu_last_error(const svdgpu_t *h);

/* ---- configuration ----------------------------------------------------- */
/* replaces: SVDTrainParam::set_param (model.h:350-368) + set_round's lr decay (base.h:470-478) */
int svdgpu_set_hparams(svdgpu_t *h, const svdgpu_hparams *hp);
int svdgpu_set_mode(svdgpu_t *h, int mode);
/* Tunables, by name:
 *   "scatter_user", "scatter_item" : 0 plain store, 1 red.global.add of the delta (hogwild)
 *   "exact_dot"   : 1 keep the reference's 4-lane dot order in hogwild mode (default 0)
 *   "pass1"       : 0 sends every row through the generic pass (default 1: basic-MF rows take the fast pass)
 *   "ring_depth", "mf_ctas" : fast-pass tuning (gather ring depth 2|4, CTAs per SM 2|3)
 *   "lanes"       : lanes per instance in hogwild mode (0 = auto = pitch/4, max 32)
 *   "chunk_rows"  : rows per launch for host-pointer calls
 *   "ctas_per_sm" : persistent CTAs per SM (0 = auto)
 *   "compact_h2d" : 1 (default): Hogwild / predict host-pointer calls of at least "compact_min_rows"
 *                   rows (default 262144) do not copy a chunk's row_ptr when all its rows have the same
 *                   feature counts, nor its values when all are 1.0f; host threads ("scan_threads",
 *                   0 = this process's share of the cores, at most 16; fewer than 5: no scan) verify every element while earlier chunks are copied, the
 *                   arrays are rebuilt on the device.  What the kernels read is identical either way.
 *                   Ordered-mode calls take this path when four or more ranks share the host
 *                   (LOCAL_WORLD_SIZE) or with 2 (alone the item-owner kernel, not the bus, sets the pace,
 *                   and copying everything measured faster); 0: never.
 *   "hogwild_safety": Hogwild stability guard, per mille (default 1000; 0 = off).  N instances in flight that
 *                   touch a row with probability p apply ~N*p stale steps of size lr to it at once; beyond
 *                   N*p*lr ~ 2 asynchronous SGD on that row diverges.  Training launches of the fast passes
 *                   keep at most safety / (lr * share of the hottest item) instances in flight (the share is
 *                   measured on the device: once per resident batch, on the first chunk of a host call);
 *                   counters "inflight_cap", "hot_item_ppm".  configs[1] is far below the bound: no effect.
 *   "exact_opt"   : ordered kernel hand-off variants, bit mask (default 5): 1 release without a
 *                   per-lane fence, 2 poll back to back, 4 instance slice staged in shared memory
 *   "exact_owner" : 1 (default): in the ordered mode, launches made only of basic-MF rows (0 | 1 | 1
 *                   features) under plain L2 decay take the item-owner kernel (k_own: every item row
 *                   stays on one warp, user rows travel under version counters; bit-identical to
 *                   k_exact and to the reference); 0 keeps every row in k_exact.  Tuning:
 *                   "own_min_rows" (4096: smaller launches keep k_exact), "own_batch" (24: version
 *                   publishes the busiest owner holds back per release fence, <= 32),
 *                   "own_urgent_gap" (4096: a user whose next rating follows within this many rows
 *                   is published at once), "own_slots" (0 = auto: item rows per owner kept in
 *                   shared memory, <= 32; the rest of an owner's rows stay in L2), "own_partner" (0; 1 = the
 *                   variant that splits a link over an owner and a partner warp, k_own2: same results,
 *                   measured slower on configs[1]), "own_redeal" (25: a plan keeps the
 *                   deal of items to owners of the plan before it while the heaviest owner stays within this %
 *                   of the mean -- a host-pointer call plans every chunk, and dealing anew costs the host
 *                   2.3 ms; 0 = deal every time; counters "own_deals", "own_redeals"), "own_isolate" (200: an item that
 *                   carries more than this % of the mean owner load gets the issue port of its owner warp
 *                   to itself -- the two owner warps that share the port stay empty; 0 = off) and
 *                   "own_isolate_full" (75: ... and one within this % of the hottest item's count a whole
 *                   SM, for at most an eighth of the SMs; 0 = off): the chain of the hottest item is what
 *                   the launch waits for, and a link of it is slower in company */
int svdgpu_set_option(svdgpu_t *h, const char *name, long long value);
/* Launch on this CUDA stream (a cudaStream_t) instead of the handle's own. */
int svdgpu_set_stream(svdgpu_t *h, void *cuda_stream);

/* Side features (the reference's feature_user / feature_item files, SparseFeatureArray<float>,
 * apex-utils/apex_utils.h:141-196): feature index r of the user (which = 0) or item (which = 1)
 * space expands to the extra pairs (index[j], value[j]), j in [row_ptr[r], row_ptr[r+1]);
 * indices >= num_row have none.  num_row = 0 clears.  Every later update/predict call applies
 * the expansion exactly as base.h:298-308,330-349,365-379,399-422 do (on the host, per chunk).
 * replaces: SparseFeatureArray::load + the "extra feature" loops of SVDFeature. */
int svdgpu_set_side_features(svdgpu_t *h, int which, int num_row, const unsigned *row_ptr,
                             const unsigned *index, const float *value);

/* Ranged weight decay (the reference's "up:" / "ip:" / "uip:" / "gp:" keys, ParameterSet,
 * base.h:33-75): indices of the user (which = 0), item (1) or global (2) space below bound[0]
 * take wd[0], those in [bound[j-1], bound[j]) take wd[j] instead of wd_user / wd_item /
 * wd_global; bound[] is what the configuration file gives ("up:bound = 100": EXCLUSIVE, strictly
 * ascending, non-zero).  An index >= bound[n-1] met during training is the reference's
 * "bound set err" (reported by svdgpu_sync).  n = 0 clears; n <= 8.
 * replaces: ParameterSet::set_param / get_wd as used by reg_global / reg_user / reg_item
 * (base.h:189,213,253). */
int svdgpu_set_wd_ranges(svdgpu_t *h, int which, int n, const unsigned *bound, const float *wd);

/* ---- model transfer ---------------------------------------------------- */
/* Layout = SVDModel's slabs (model.h:481-556): ui_bias[rows], W_uiset[rows][pitch],
 * g_bias[num_global] with rows = ustart + num_user + num_item, ustart =
 * num_ufeedback when format_type == 1 else 0; feedback rows first, then user
 * rows, then item rows.  pitch_floats is the host row stride in floats
 * (>= num_factor).  replaces: rand_init / load_from_file / save_to_file targets. */
int svdgpu_upload_model(svdgpu_t *h, const float *ui_bias, const float *W_uiset,
                        size_t pitch_floats, const float *g_bias);
int svdgpu_download_model(svdgpu_t *h, float *ui_bias, float *W_uiset, size_t pitch_floats,
                          float *g_bias);

/* ---- the hot path, host buffers ---------------------------------------- */
/* One SVDFeatureCSR batch (apex_svd_data.h:34-231): row r owns feature segments
 * global [row_ptr[3r],row_ptr[3r+1]), user [..3r+1,3r+2), item [..3r+2,3r+3) of
 * index/value.  replaces: for each row ISVDTrainer::update(Elem) ->
 * SVDFeature::update_inner (base.h:456-466). */
int svdgpu_update_csr(svdgpu_t *h, int num_row, const int *row_ptr, const float *label,
                      const unsigned *index, const float *value);
/* replaces: for each row ISVDTrainer::predict(Elem) -> SVDFeature::pred (base.h:445-454,467-469) */
int svdgpu_predict_csr(svdgpu_t *h, int num_row, const int *row_ptr, const float *label,
                       const unsigned *index, const float *value, float *out);
/* User-grouped input (SVDPlusBlock, apex_svd_data.h:376-466): block b owns rows
 * [blk_row_off[b], blk_row_off[b+1]) of the CSR and feedback entries
 * [blk_fb_off[b], blk_fb_off[b+1]); blk_tag (svdpp_tag, may be NULL = DEFAULT).
 * replaces: for each block SVDPPFeature::update(SVDPlusBlock) (base.h:568-582). */
int svdgpu_update_ugroup(svdgpu_t *h, int num_block, const int *blk_row_off,
                         const int *blk_fb_off, const int *blk_tag, const unsigned *fb_index,
                         const float *fb_value, const int *row_ptr, const float *label,
                         const unsigned *index, const float *value);
/* replaces: SVDPPFeature::predict(vector<float>&, SVDPlusBlock) (base.h:583-591) */
int svdgpu_predict_ugroup(svdgpu_t *h, int num_block, const int *blk_row_off,
                          const int *blk_fb_off, const int *blk_tag, const unsigned *fb_index,
                          const float *fb_value, const int *row_ptr, const float *label,
                          const unsigned *index, const float *value, float *out);

/* ---- evaluation on the device ------------------------------------------- */
/* Predict and accumulate the squared error on the device; only two doubles come back.
 * *sum_sq = sum over rows of ((pred - label) * scale)^2, *count = rows.
 * replaces: the predict + RMSEEvaluator::add_eval loop of svd_feature_infer.cpp:38-56,243-277
 * (the caller prints sqrt(sum_sq / count)). */
int svdgpu_eval_csr(svdgpu_t *h, int num_row, const int *row_ptr, const float *label,
                    const unsigned *index, const float *value, float scale, double *sum_sq,
                    long long *count);
int svdgpu_eval_ugroup(svdgpu_t *h, int num_block, const int *blk_row_off, const int *blk_fb_off,
                       const int *blk_tag, const unsigned *fb_index, const float *fb_value,
                       const int *row_ptr, const float *label, const unsigned *index,
                       const float *value, float scale, double *sum_sq, long long *count);

/* ---- bulk ingest of the reference's binary buffer files ------------------ */
/* One pass over a buffer file written by tools/make_feature_buffer (format_type 0) or
 * tools/make_ugroup_buffer (format_type 1): batches are read into pinned memory, concatenated
 * into chunks of "chunk_rows" rows and handed to the hot path.
 * replaces: SVDFeatureCSRFactory / SVDPlusBlockFactory::load_next (apex_svd_data.cpp:218-248,
 * 556-640; apex_svd_data.h:220-230,437-450) + the per-row loop svd_feature.cpp:231-247. */
int svdgpu_update_buffer_file(svdgpu_t *h, const char *path, long long *num_row);
int svdgpu_predict_buffer_file(svdgpu_t *h, const char *path, float *out, long long out_cap,
                               long long *num_row);
int svdgpu_eval_buffer_file(svdgpu_t *h, const char *path, float scale, double *sum_sq,
                            long long *num_row);

/* ---- the hot path, batches resident in HBM ----------------------------- */
/* Copy a CSR batch to the device once and train on it every round (the
 * reference re-reads its buffer file each round, svd_feature.cpp:245-246). */
int svdgpu_batch_create(svdgpu_t *h, svdgpu_batch_t **out, int num_row, const int *row_ptr,
                        const float *label, const unsigned *index, const float *value);
/* optional: attach user-group structure to a resident batch */
int svdgpu_batch_set_ugroup(svdgpu_t *h, svdgpu_batch_t *b, int num_block, const int *blk_row_off,
                            const int *blk_fb_off, const int *blk_tag, const unsigned *fb_index,
                            const float *fb_value);
/* rows [row_begin,row_end) (blocks [begin,end) for a user-grouped batch) */
int svdgpu_batch_update(svdgpu_t *h, svdgpu_batch_t *b, int begin, int end);
/* (Re)build the ordered mode's item-owner plan of a resident batch (svdgpu_batch_create builds it
 * already when the trainer is in the ordered mode; this entry exists to time the plan and to plan a
 * batch created in Hogwild mode).  *planned = 1 when the batch qualifies (basic-MF rows, plain L2
 * decay), else 0 and the ordered mode keeps k_exact. */
int svdgpu_batch_plan(svdgpu_t *h, svdgpu_batch_t *b, int *planned);
/* out_host may be NULL (predictions stay on the device; see svdgpu_batch_pred_ptr) */
int svdgpu_batch_predict(svdgpu_t *h, svdgpu_batch_t *b, int begin, int end, float *out_host);
/* squared error of rows [begin,end) of a resident batch (see svdgpu_eval_csr) */
int svdgpu_batch_eval(svdgpu_t *h, svdgpu_batch_t *b, int begin, int end, float scale, double *sum_sq,
                      long long *count);
void svdgpu_batch_destroy(svdgpu_t *h, svdgpu_batch_t *b);

/* ---- pairwise-rank sample generation on the device ----------------------- */
/* PairwiseRankGenerator's parameters (apex_svd_data.cpp:967-989), same names and defaults. */
typedef struct {
  int rank_sample_method;    /* 0: positives vs negatives (sample_posneg); 1: every row against one random row
                                of its block whose label differs by more than rank_sample_gap (sample_cmp).
                                +10: label = p.label - n.label instead of 1 (genpair's second branch; the
                                reference's own switch rejects 10/11, apex_svd_data.cpp:1007-1011)        */
  int rank_sample_num;       /* pairs per block; <= 0: one per negative row (reference: -1)     */
  int rank_sample_max;       /* cap on pairs per block; <= 0: none (reference: INT_MAX)         */
  int rank_sample_pointwise; /* emit (p, label 1) and (n, label 0) instead of the merged pair   */
  float pos_sample_lowerb;   /* reference default 0.8  */
  float neg_sample_upperb;   /* reference default 1e-6 */
  float rank_sample_gap;     /* method 1 only; reference default 0.0001, must be > 0 */
  unsigned long long seed;   /* change it every round: the reference reshuffles on every pass   */
} svdgpu_pair_params;
/* From a resident user-grouped batch of rated rows make the resident batch of pair rows the
 * reference's host sampler would feed the trainer (same blocks and feedback lists, rows replaced).
 * Train on it with svdgpu_batch_update; free it with svdgpu_batch_destroy.
 * replaces: PairwiseRankGenerator::next -> sample_posneg / sample_cmp / genpair / merge
 * (apex_svd_data.cpp:828-965, 1001-1018); the rand()-driven shuffles become keyed
 * permutations, so parity with the host sampler is structural, not bit-level. */
int svdgpu_batch_sample_pairs(svdgpu_t *h, svdgpu_batch_t *src, const svdgpu_pair_params *params,
                              svdgpu_batch_t **out);
/* Copy a resident batch back to the host (any output may be NULL): row_ptr[3*num_row+1],
 * label[num_row], index/value[num_val], blk_row_off[num_block+1]. */
int svdgpu_batch_download(svdgpu_t *h, svdgpu_batch_t *b, int *num_row, long long *num_val,
                          int *row_ptr, float *label, unsigned *index, float *value,
                          int *blk_row_off);

/* ---- ranking a fixed item set for many users (SVDFeatureRanker) ---------- */
/* The reference's ranker (base.h:597-813) reads a TAGGED instance stream: the label field of a row
 * holds a svdranker_tag (apex_svd.h:115-152): 0 ITEM (item + global features of the next candidate),
 * 2 USER (user features; opens a user section), 1 POS / -1 BAN (candidate positions, in the
 * user-feature field), 3 SPEC (candidate position in the user-feature field + extra item / global
 * features scored for this user only), 4 PROCESS (rank now).  Every PROCESS row appends to the
 * result: top_k > 0: the positions of the top_k candidates by score; top_k = 0: the rank position
 * of every POS candidate, in the order they were given.  Equal scores: lower position first (the
 * reference's std::sort leaves it unspecified).  The model is the one in the handle
 * (svdgpu_upload_model); side features (svdgpu_set_side_features) apply as in base.h:698-702,731-734.
 *
 * replaces: SVDFeatureRanker::init_ranker (base.h:666-685; top_k is its "top_k" parameter) */
int svdgpu_rank_init(svdgpu_t *h, int num_item_set, int top_k);
/* replaces: for each row ISVDRanker::process(result, Elem) (base.h:795-797; the loop of
 * svd_feature_infer.cpp:352-360).  The stream may be cut anywhere between calls; sections closed by
 * a PROCESS row inside the call are ranked together on the device.  *num_result receives the
 * number of results; if they do not fit result_cap the call fails and keeps them for the next call. */
int svdgpu_rank_csr(svdgpu_t *h, int num_row, const int *row_ptr, const float *label, const unsigned *index,
                    const float *value, int *result, long long result_cap, long long *num_result);
/* replaces: for each block ISVDRanker::process(result, SVDPlusBlock) (base.h:798-812): a DEFAULT or
 * START block's feedback list becomes the implicit-feedback part of the following USER rows. */
int svdgpu_rank_ugroup(svdgpu_t *h, int num_block, const int *blk_row_off, const int *blk_fb_off,
                       const int *blk_tag, const unsigned *fb_index, const float *fb_value,
                       const int *row_ptr, const float *label, const unsigned *index, const float *value,
                       int *result, long long result_cap, long long *num_result);

/* ---- synchronisation, timing, introspection ---------------------------- */
/* wait for all queued work; reports device-side input errors (index out of
 * bound -- the reference's assert_true at base.h:320,327,343) */
int svdgpu_sync(svdgpu_t *h);
/* CUDA-event stopwatch on the launch stream */
int svdgpu_timer_start(svdgpu_t *h);
int svdgpu_timer_stop(svdgpu_t *h, float *elapsed_ms);
/* "kernel_launches", "instances", "h2d_bytes", "d2h_bytes", "num_sm", "lanes", "own_launches", "own_rows",
 * "collectives", "collective_bytes" (NCCL all-reduces issued by the handle and the bytes they reduced),
 * "ingest_read_us", "ingest_call_us" (bulk ingest: time reading the file / inside the hot-path calls) */
long long svdgpu_get_counter(const svdgpu_t *h, const char *name);
/* Diagnostics of the last item-owner launch (option "own_stats" = 1 before it): per owner warp
 * {cycles in its queue loop, cycles waiting for a ring slot, cycles in publish fences, number of
 * waits}; out holds 4 * cap_owners values, *num_owner receives the owner count. */
int svdgpu_own_stats(svdgpu_t *h, long long *out, int cap_owners, int *num_owner);
/* Bandwidth probe on this handle's device and stream: which = 0 streams `bytes` of device memory
 * `iters` times with 16-byte loads, 1 copies it (read + write counted).  A buffer that fits L2
 * (<= 64 MB) measures the L2 rate the gather kernels live on, a multi-GB one the HBM rate. */
int svdgpu_microbench(svdgpu_t *h, int which, size_t bytes, int iters, double *gbytes_per_s);
/* device pointers of the model slabs (for peer / collective plumbing): 0 ui_bias,
 * 1 W_uiset, 2 g_bias; *pitch_floats receives the device row stride */
void *svdgpu_device_ptr(svdgpu_t *h, int which, size_t *pitch_floats);

/* ---- multi-GPU: replicated item side, user rows sharded by hash(user) ---- */
/* Snapshot the replicated slabs (item rows, item bias, g_bias[, feedback rows]). */
int svdgpu_items_snapshot(svdgpu_t *h);
/* Pack delta = current - snapshot into one contiguous device buffer; returns its
 * device pointer and length in floats (the buffer a NCCL all-reduce runs on). */
int svdgpu_items_pack_delta(svdgpu_t *h, void **dev_ptr, size_t *num_floats);
/* current = snapshot + scale * (reduced delta); then re-snapshot. */
int svdgpu_items_apply_delta(svdgpu_t *h, float scale);


/* ---- multi-GPU: the exchange itself (NCCL over NVLink, bound at run time) -------------------
 * One handle per GPU -- one process per GPU, or several handles in one process.  The reference has
 * no distributed code; these entry points serve its training loop (svd_feature.cpp:231-247) when
 * it runs once per GPU on the rows of that GPU's users (user id mod world == rank).
 *
 * svdgpu_comm_id: rank 0 makes the 128-byte rendezvous id (ncclGetUniqueId) and hands it to the
 * other ranks by any means (a file, MPI, torch.distributed); svdgpu_comm_init joins (collective). */
int svdgpu_comm_id(char *id128);
int svdgpu_comm_init(svdgpu_t *h, int world, int rank, const char *id128);
int svdgpu_comm_destroy(svdgpu_t *h);
int svdgpu_comm_rank(const svdgpu_t *h, int *world, int *rank);
/* Exchange the replicated slabs: delta = current - snapshot, ONE ncclAllReduce(sum) on the launch
 * stream, current = snapshot + scale * sum, re-snapshot (scale 1 keeps the whole step mass of all
 * ranks, 1/world averages).  The first call only takes the snapshot.  Collective. */
int svdgpu_allreduce_items(svdgpu_t *h, float scale);
/* Complete the user side on every rank (before save_model / evaluation): rows of users other ranks
 * own are zeroed and the user slab is summed in place.  Collective. */
int svdgpu_allgather_users(svdgpu_t *h);
/* Several handles of ONE process (hs[i] on its own device): the same exchange as one NCCL group;
 * communicators are created on first use (ncclCommInitAll).   */
int svdgpu_allreduce_items_group(svdgpu_t **hs, int n, float scale);

#ifdef __cplusplus
}
#endif
#endif /* SVDGPU_H_ */
