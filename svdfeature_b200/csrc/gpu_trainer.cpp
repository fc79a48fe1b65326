// gpu_trainer.cpp -- the reference's solver seam, implemented on the B200.
//
// `GpuSVDFeature` is an apex_svd::ISVDTrainer (apex_svd.h:33-107) and this file
// defines apex_svd::create_svd_trainer (apex_svd.h:212), i.e. it takes the place
// of the reference's apex_svd.cpp at link time exactly as
// solvers/example/Makefile:17-23 documents for custom solvers.  It covers what
// SVDFeature and SVDPPFeature (solvers/base-solver/apex_svd_base.h:79-592) do:
// same set_param keys, same call order, same model-file bytes
// (apex_svd_model.h:570-660), same error convention (message on stderr, exit(-1)).
//
// The class owns no arithmetic.  Model initialisation and file I/O are host work
// (they are not on the hot path and must reproduce glibc rand() draws); every
// update/predict goes through the C ABI of include/svdgpu.h to the CUDA kernels.
#ifdef SVDGPU_WITH_REFERENCE_HEADERS
#include "apex_svd.h"
#else
#include "apex_compat.h"
#endif

#include "svdgpu.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>

namespace {

// On-disk model header: SVDModelParam, apex_svd_model.h:373-450 (17 fields in
// declaration order, then int reserved[247]; 1056 bytes).
struct ModelHeader {
  int32_t num_user, num_item, num_factor, num_global;
  float u_init_sigma, i_init_sigma, base_score;
  int32_t no_user_bias, num_ufeedback;
  float ufeedback_init_sigma;
  int32_t num_randinit_ufactor, num_randinit_ifactor;
  int32_t common_latent_space, user_nonnegative, common_feedback_space, extend_flag, item_nonnegative;
  int32_t reserved[247];
};
static_assert(sizeof(ModelHeader) == 1056, "model header must stay 1056 bytes");

// apex-tensor/apex_random.h:53-77 (uniform in (0,1) from rand(); Marsaglia polar,
// second variate discarded)
inline double next_double2() { return ((double)rand() + 1.0) / ((double)RAND_MAX + 2.0); }
inline double sample_normal() {
  double x, y, s;
  do {
    x = 2 * next_double2() - 1.0;
    y = 2 * next_double2() - 1.0;
    s = x * x + y * y;
  } while (s >= 1.0 || s == 0.0);
  return x * std::sqrt(-2.0 * std::log(s) / s);
}

// apex_svd_model.h:220-237
float transform_base_score(float base_score, int active_type) {
  switch (active_type) {
    case 0: case 5: case 6: return base_score;
    case 1: case 2: case 3: case 7:
      apex_utils::assert_true(base_score > 0.0f && base_score < 1.0f, "sigmoid range constrain");
      return -logf(1.0f / base_score - 1.0f);
    default: apex_utils::error("unkown active type"); return 0.0f;
  }
}

void check(svdgpu_t *h, int rc) {
  if (rc != 0) apex_utils::error(svdgpu_last_error(h));
}

// growable SoA staging for the per-row / per-block API.  push() is the per-instance cost of the
// reference's calling convention (one virtual update(Elem) per rating, svd_feature.cpp:231-247): the
// arrays grow geometrically and a row is appended with plain stores (no per-segment vector::insert).
struct CsrStage {
  std::vector<int> row_ptr;
  std::vector<float> label;
  std::vector<unsigned> index;
  std::vector<float> value;
  size_t nrow_ = 0, nval_ = 0;  // used prefixes of the arrays (their size() is the capacity in use + slack)
  CsrStage() { row_ptr.assign(1, 0); }
  void clear() {
    nrow_ = nval_ = 0;
    row_ptr.resize(1);
    row_ptr[0] = 0;
    label.clear();
    index.clear();
    value.clear();
  }
  int num_row() const { return (int)nrow_; }
  void reserve_rows(size_t rows, size_t vals_per_row) {
    row_ptr.reserve(3 * rows + 1);
    label.reserve(rows);
    index.reserve(rows * vals_per_row);
    value.reserve(rows * vals_per_row);
  }
  void push(const apex_svd::SVDFeatureCSR::Elem &e) {
    const size_t n = (size_t)e.num_global + (size_t)e.num_ufactor + (size_t)e.num_ifactor;
    if (label.size() < nrow_ + 1) {
      const size_t cap = std::max<size_t>(1024, 2 * (nrow_ + 1));
      label.resize(cap);
      row_ptr.resize(3 * cap + 1);
    }
    if (index.size() < nval_ + n) {
      const size_t cap = std::max<size_t>(4096, 2 * (nval_ + n));
      index.resize(cap);
      value.resize(cap);
    }
    label[nrow_] = e.label;
    unsigned *ix = index.data() + nval_;
    float *vl = value.data() + nval_;
    int *rp = row_ptr.data() + 3 * nrow_;  // rp[0] is this row's first position, already set
    size_t k = 0;
    for (int i = 0; i < e.num_global; ++i, ++k) ix[k] = e.index_global[i], vl[k] = e.value_global[i];
    rp[1] = (int)(nval_ + k);
    for (int i = 0; i < e.num_ufactor; ++i, ++k) ix[k] = e.index_ufactor[i], vl[k] = e.value_ufactor[i];
    rp[2] = (int)(nval_ + k);
    for (int i = 0; i < e.num_ifactor; ++i, ++k) ix[k] = e.index_ifactor[i], vl[k] = e.value_ifactor[i];
    rp[3] = (int)(nval_ + k);
    nval_ += n;
    nrow_ += 1;
  }
};

// ParameterSet (base.h:33-75): "<prefix>wd" / "<prefix>bound" pairs given in order; the bounds are
// kept as the configuration gives them (the C ABI takes them exclusive).
struct RangedWd {
  std::string prefix_a, prefix_b;
  std::vector<float> wd;
  std::vector<unsigned> bound;
  RangedWd(const char *a, const char *b) : prefix_a(a), prefix_b(b) {}
  void set_param(const char *name, const char *val) {
    if (!strncmp(name, prefix_a.c_str(), prefix_a.size())) name += prefix_a.size();
    else if (!strncmp(name, prefix_b.c_str(), prefix_b.size())) name += prefix_b.size();
    else return;
    if (!strcmp("bound", name)) {
      const unsigned bd = (unsigned)atoi(val);
      apex_utils::assert_true(bd > 0, "can't give 0 as bound");
      apex_utils::assert_true(bound.empty() || bound.back() < bd, "bound must be given in order");
      apex_utils::assert_true(bound.size() + 1 == wd.size(), "must specifiy wd in each range");
      bound.push_back(bd);
    }
    if (!strcmp("wd", name)) {
      apex_utils::assert_true(wd.size() == bound.size(), "setting must be exactly");
      wd.push_back((float)atof(val));
    }
  }
};

}  // namespace

namespace apex_svd {

class GpuSVDFeature : public ISVDTrainer {
 public:
  explicit GpuSVDFeature(const SVDTypeParam &mtype) : mtype_(mtype) {
    memset(&mp_, 0, sizeof(mp_));
    mp_.u_init_sigma = mp_.i_init_sigma = 0.01f;  // model.h:436-450
    mp_.base_score = 0.5f;
    memset(&hp_, 0, sizeof(hp_));
    hp_.learning_rate = 0.01f;  // model.h:334-344
    hp_.scale_lr_ufeedback = 1.0f;
    svdpp_ = (mtype.extend_type == 1 || mtype.format_type == svd_type::USER_GROUP_FORMAT);
    blk_row_off_.push_back(0);
    blk_fb_off_.push_back(0);
  }
  virtual ~GpuSVDFeature() {
    if (h_) svdgpu_destroy(h_);
  }

  // ---- model related interface (base.h:126-173) ---------------------------
  virtual void set_param(const char *name, const char *val) {
    if (!strcmp(name, "feature_user")) name_feat_user_ = val;  // base.h:127-128
    if (!strcmp(name, "feature_item")) name_feat_item_ = val;
    // ranged weight decay (base.h:136-138): every key is offered to all three sets
    u_param_.set_param(name, val);
    i_param_.set_param(name, val);
    g_param_.set_param(name, val);
    // SVDTrainParam::set_param, model.h:350-368
    if (!strcmp("learning_rate", name)) hp_.learning_rate = (float)atof(val);
    if (!strcmp("wd_user", name)) hp_.wd_user = (float)atof(val);
    if (!strcmp("wd_item", name)) hp_.wd_item = (float)atof(val);
    if (!strcmp("wd_uiset", name)) hp_.wd_user = hp_.wd_item = (float)atof(val);
    if (!strcmp("wd_user_bias", name)) hp_.wd_user_bias = (float)atof(val);
    if (!strcmp("wd_item_bias", name)) hp_.wd_item_bias = (float)atof(val);
    if (!strcmp("wd_uiset_bias", name)) hp_.wd_user_bias = hp_.wd_item_bias = (float)atof(val);
    if (!strcmp("wd_global", name)) hp_.wd_global = (float)atof(val);
    if (!strcmp("reg_method", name)) hp_.reg_method = atoi(val);
    if (!strcmp("reg_global", name)) hp_.reg_global = atoi(val);
    if (!strcmp("num_regfree_global", name)) hp_.num_regfree_global = (unsigned)atoi(val);
    if (!strcmp("decay_learning_rate", name)) decay_learning_rate_ = atoi(val);
    if (!strcmp("decay_rate", name)) decay_rate_ = (float)atof(val);
    if (!strcmp("scale_lr_ufeedback", name)) hp_.scale_lr_ufeedback = (float)atof(val);
    if (!strcmp("wd_ufeedback", name)) hp_.wd_ufeedback = (float)atof(val);
    if (!strcmp("wd_ufeedback_bias", name)) hp_.wd_ufeedback_bias = (float)atof(val);
    // GPU-side keys (new; the reference ignores unknown keys, so one config serves both)
    if (!strcmp("gpu:device", name)) device_ = atoi(val);
    if (!strcmp("gpu:mode", name)) {
      if (!strcmp(val, "exact") || !strcmp(val, "0")) mode_ = SVDGPU_MODE_EXACT;
      else if (!strcmp(val, "hogwild") || !strcmp(val, "1")) mode_ = SVDGPU_MODE_HOGWILD;
      else apex_utils::error("gpu:mode must be exact or hogwild");
    }
    if (!strcmp("gpu:batch", name)) batch_rows_ = atoi(val) > 0 ? atoi(val) : batch_rows_;
    // multi-GPU: one process of the unchanged driver per GPU, every one reading the same input;
    // a process trains the rows of its users (first user feature mod world == rank) and the item
    // side is all-reduced over NCCL (svdgpu_comm.cu)
    if (!strcmp("gpu:world", name)) world_ = atoi(val) > 0 ? atoi(val) : 1;
    if (!strcmp("gpu:rank", name)) rank_ = atoi(val);
    if (!strcmp("gpu:nccl_id", name)) nccl_id_file_ = val;
    if (!strcmp("gpu:allreduce_rows", name)) xchg_rows_ = atoll(val) > 0 ? atoll(val) : xchg_rows_;
    if (!strcmp("gpu:allreduce_scale", name)) xchg_scale_ = !strcmp(val, "mean") ? -1.0f : (float)atof(val);
    if (!strncmp("gpu:opt:", name, 8)) options_.push_back(std::make_pair(std::string(name + 8), atoll(val)));
    if (h_) {
      flush();  // rows staged so far were handed over under the old hyper-parameters (the reference applies them per row)
      push_hparams();
      if (!strcmp("gpu:mode", name)) check(h_, svdgpu_set_mode(h_, mode_));
      if (!strncmp("gpu:opt:", name, 8)) check(h_, svdgpu_set_option(h_, name + 8, atoll(val)));
    }
    if (space_allocated_) return;  // base.h:133-135
    // SVDModelParam::set_param, model.h:456-476
    if (!strcmp("num_user", name)) mp_.num_user = atoi(val);
    if (!strcmp("num_item", name)) mp_.num_item = atoi(val);
    if (!strcmp("num_uiset", name)) mp_.num_user = mp_.num_item = atoi(val);
    if (!strcmp("num_global", name)) mp_.num_global = atoi(val);
    if (!strcmp("num_factor", name)) mp_.num_factor = atoi(val);
    if (!strcmp("u_init_sigma", name)) mp_.u_init_sigma = (float)atof(val);
    if (!strcmp("i_init_sigma", name)) mp_.i_init_sigma = (float)atof(val);
    if (!strcmp("ui_init_sigma", name)) mp_.u_init_sigma = mp_.i_init_sigma = (float)atof(val);
    if (!strcmp("base_score", name)) mp_.base_score = (float)atof(val);
    if (!strcmp("no_user_bias", name)) mp_.no_user_bias = atoi(val);
    if (!strcmp("num_ufeedback", name)) mp_.num_ufeedback = atoi(val);
    if (!strcmp("num_randinit_ufactor", name)) mp_.num_randinit_ufactor = atoi(val);
    if (!strcmp("num_randinit_ifactor", name)) mp_.num_randinit_ifactor = atoi(val);
    if (!strcmp("num_randinit_uifactor", name)) mp_.num_randinit_ifactor = mp_.num_randinit_ufactor = atoi(val);
    if (!strcmp("ufeedback_init_sigma", name)) mp_.ufeedback_init_sigma = (float)atof(val);
    if (!strcmp("common_latent_space", name)) mp_.common_latent_space = atoi(val);
    if (!strcmp("common_feedback_space", name)) mp_.common_feedback_space = atoi(val);
    if (!strcmp("user_nonnegative", name)) mp_.user_nonnegative = atoi(val);
    if (!strcmp("item_nonnegative", name)) mp_.item_nonnegative = atoi(val);
  }

  virtual void load_model(FILE *fi) {  // model.h:570-633
    flush();
    if (fread(&mp_, sizeof(ModelHeader), 1, fi) != 1) {
      printf("error loading CF SVD model\n");
      exit(-1);
    }
    alloc_host();
    read_1d(&ui_bias_[ustart_], mp_.num_user, fi);
    read_2d(&W_[(size_t)ustart_ * pitch_], mp_.num_user, fi);
    read_1d(&ui_bias_[ustart_ + mp_.num_user], mp_.num_item, fi);
    read_2d(&W_[(size_t)(ustart_ + mp_.num_user) * pitch_], mp_.num_item, fi);
    read_1d(g_bias_.data(), mp_.num_global, fi);
    if (mtype_.format_type == svd_type::USER_GROUP_FORMAT) {
      read_1d(ui_bias_.data(), mp_.num_ufeedback, fi);
      read_2d(W_.data(), mp_.num_ufeedback, fi);
    }
    if (h_) {
      ensure_handle();  // shape may have changed
      upload();
    }
  }

  virtual void save_model(FILE *fo) {  // model.h:638-660
    flush();
    if (h_ && world_ > 1 && comm_on_) {  // every process saves at the same points of the driver's loop
      exchange();
      check(h_, svdgpu_allgather_users(h_));  // the file holds every user's rows, whoever trained them
    }
    if (h_) check(h_, svdgpu_download_model(h_, ui_bias_.data(), W_.data(), (size_t)pitch_, g_bias_.data()));
    fwrite(&mp_, sizeof(ModelHeader), 1, fo);
    write_1d(&ui_bias_[ustart_], mp_.num_user, fo);
    write_2d(&W_[(size_t)ustart_ * pitch_], mp_.num_user, fo);
    write_1d(&ui_bias_[ustart_ + mp_.num_user], mp_.num_item, fo);
    write_2d(&W_[(size_t)(ustart_ + mp_.num_user) * pitch_], mp_.num_item, fo);
    write_1d(g_bias_.data(), mp_.num_global, fo);
    if (mtype_.format_type == svd_type::USER_GROUP_FORMAT) {
      write_1d(ui_bias_.data(), mp_.num_ufeedback, fo);
      write_2d(W_.data(), mp_.num_ufeedback, fo);
    }
  }

  virtual void init_model(void) {  // base.h:146-149; model.h:665-705 rand_init
    alloc_host();
    mp_.base_score = transform_base_score(mp_.base_score, mtype_.active_type);
    const int k = mp_.num_factor;
    {
      const int ny = mp_.num_randinit_ufactor != 0 ? mp_.num_randinit_ufactor : mp_.num_user;
      float *base = &W_[(size_t)ustart_ * pitch_];
      for (int y = 0; y < ny; ++y)
        for (int x = 0; x < k; ++x) base[(size_t)y * pitch_ + x] = (float)sample_normal() * mp_.u_init_sigma;
      if (mp_.user_nonnegative)  // model.h:678-683 (all num_user rows)
        for (int y = 0; y < mp_.num_user; ++y)
          for (int x = 0; x < k; ++x) base[(size_t)y * pitch_ + x] = fabsf(base[(size_t)y * pitch_ + x]);
    }
    {
      const int ny = mp_.num_randinit_ifactor != 0 ? mp_.num_randinit_ifactor : mp_.num_item;
      float *base = &W_[(size_t)(ustart_ + mp_.num_user) * pitch_];
      for (int y = 0; y < ny; ++y)
        for (int x = 0; x < k; ++x) base[(size_t)y * pitch_ + x] = (float)sample_normal() * mp_.i_init_sigma;
      if (mp_.item_nonnegative)  // model.h:695-700 (the randomly initialised rows)
        for (int y = 0; y < ny; ++y)
          for (int x = 0; x < k; ++x) base[(size_t)y * pitch_ + x] = fabsf(base[(size_t)y * pitch_ + x]);
    }
    if (mtype_.format_type == svd_type::USER_GROUP_FORMAT) {
      for (int y = 0; y < mp_.num_ufeedback; ++y)
        for (int x = 0; x < k; ++x)
          W_[(size_t)y * pitch_ + x] = (float)sample_normal() * mp_.ufeedback_init_sigma;
    }
  }

  virtual void init_trainer(void) {  // base.h:151-173
    apex_utils::assert_true(space_allocated_ != 0, "init_trainer: no model (call init_model or load_model first)");
    ensure_handle();
    load_side_features(0, name_feat_user_);  // base.h:152-153
    load_side_features(1, name_feat_item_);
    upload();
    join_comm();
    init_end_ = 1;
  }

  // ---- training interface ---------------------------------------------------
  virtual void set_round(int nround) {  // base.h:470-478
    if (decay_learning_rate_ != 0) {
      apex_utils::assert_true(round_counter_ <= nround, "round counter restriction");
      flush();
      while (round_counter_ < nround) {
        hp_.learning_rate *= decay_rate_;
        round_counter_++;
      }
      if (h_) push_hparams();
    }
  }
  virtual void finish_round(void) {
    flush();
    exchange();  // (every process finishes the round: the collective lines up)
  }

  virtual void update(const SVDFeatureCSR::Elem &feature) {  // base.h:464-466
    apex_utils::assert_true(init_end_ != 0, "update before init_trainer");
    if (svdpp_) apex_utils::error("GPU trainer: a user-grouped model takes SVDPlusBlock input");
    if (mine(feature))
      rows_.push(feature);  // the Elem aliases the loader's buffer: copy now (apex_buffer_loader.h:212-226)
    if (rows_.num_row() >= batch_rows_) flush();
    tick();
  }
  virtual float predict(const SVDFeatureCSR::Elem &feature) {  // base.h:467-469
    apex_utils::assert_true(init_end_ != 0, "predict before init_trainer");
    if (svdpp_) apex_utils::error("GPU trainer: a user-grouped model takes SVDPlusBlock input");
    flush();
    CsrStage one;
    one.push(feature);
    float out = 0.0f;
    check(h_, svdgpu_predict_csr(h_, 1, one.row_ptr.data(), one.label.data(), one.index.data(),
                                 one.value.data(), &out));
    return out;
  }

  virtual void update(const SVDPlusBlock &data) {  // base.h:568-582
    apex_utils::assert_true(init_end_ != 0, "update before init_trainer");
    if (!svdpp_) apex_utils::error("GPU trainer: a random-order model takes SVDFeatureCSR::Elem input");
    // (a user's blocks follow each other and carry the user in their rows: the whole unit goes to one process)
    if (data.data.num_row == 0 || mine(data.data[0])) push_block(data);
    const bool closed = data.extend_tag == svdpp_tag::DEFAULT || data.extend_tag == svdpp_tag::END_TAG;
    if (closed && rows_.num_row() >= batch_rows_) flush();
    if (closed) tick(data.data.num_row);
    else seen_ += data.data.num_row;
  }
  virtual void predict(std::vector<float> &pred, const SVDPlusBlock &data) {  // base.h:583-591
    apex_utils::assert_true(init_end_ != 0, "predict before init_trainer");
    if (!svdpp_) apex_utils::error("GPU trainer: a random-order model takes SVDFeatureCSR::Elem input");
    flush();
    // The reference prepares the feedback sum on DEFAULT / START blocks only and predicts the rows
    // of MIDDLE / END blocks with the state the last of those left (base.h:583-591): keep that
    // block's feedback list and predict every block of the user with it.
    const bool opens = data.extend_tag == svdpp_tag::DEFAULT || data.extend_tag == svdpp_tag::START_TAG;
    if (opens) {
      pred_fb_index_.assign(data.index_ufeedback, data.index_ufeedback + data.num_ufeedback);
      pred_fb_value_.assign(data.value_ufeedback, data.value_ufeedback + data.num_ufeedback);
    }
    // (own buffers: the training stage may hold the blocks of a run that is still open)
    CsrStage rows;
    for (int r = 0; r < data.data.num_row; ++r) rows.push(data.data[r]);
    const int bro[2] = {0, rows.num_row()}, bfo[2] = {0, (int)pred_fb_index_.size()}, tag[1] = {svdpp_tag::DEFAULT};
    pred.resize((size_t)rows.num_row());
    check(h_, svdgpu_predict_ugroup(h_, 1, bro, bfo, tag, pred_fb_index_.data(), pred_fb_value_.data(),
                                    rows.row_ptr.data(), rows.label.data(), rows.index.data(), rows.value.data(),
                                    pred.data()));
  }

  // ---- bulk extensions (whole batches in one call) ----------
  void update_batch(const SVDFeatureCSR &c) {
    apex_utils::assert_true(init_end_ != 0, "update before init_trainer");
    if (svdpp_) apex_utils::error("GPU trainer: a user-grouped model takes SVDPlusBlock input");
    flush();
    if (world_ > 1) {  // this process's share of the batch, through the staging path
      for (int r = 0; r < c.num_row; ++r) update(c[r]);
      return;
    }
    check(h_, svdgpu_update_csr(h_, c.num_row, c.row_ptr, c.row_label, c.feat_index, c.feat_value));
  }
  void predict_batch(const SVDFeatureCSR &c, float *out) {
    apex_utils::assert_true(init_end_ != 0, "predict before init_trainer");
    flush();
    check(h_, svdgpu_predict_csr(h_, c.num_row, c.row_ptr, c.row_label, c.feat_index, c.feat_value, out));
  }
  void update_ugroup_batch(int num_block, const int *blk_row_off, const int *blk_fb_off, const int *blk_tag,
                           const unsigned *fb_index, const float *fb_value, const SVDFeatureCSR &c) {
    apex_utils::assert_true(init_end_ != 0, "update before init_trainer");
    flush();
    check(h_, svdgpu_update_ugroup(h_, num_block, blk_row_off, blk_fb_off, blk_tag, fb_index, fb_value,
                                   c.row_ptr, c.row_label, c.feat_index, c.feat_value));
  }
  void predict_ugroup_batch(int num_block, const int *blk_row_off, const int *blk_fb_off, const int *blk_tag,
                            const unsigned *fb_index, const float *fb_value, const SVDFeatureCSR &c,
                            float *out) {
    apex_utils::assert_true(init_end_ != 0, "predict before init_trainer");
    flush();
    check(h_, svdgpu_predict_ugroup(h_, num_block, blk_row_off, blk_fb_off, blk_tag, fb_index, fb_value,
                                    c.row_ptr, c.row_label, c.feat_index, c.feat_value, out));
  }
  svdgpu_t *handle() { return h_; }
  void sync() {
    flush();
    if (h_) check(h_, svdgpu_sync(h_));
  }

 private:
  void alloc_host() {  // model.h:511-556 (separate index spaces only)
    if (mp_.common_latent_space != 0 || mp_.common_feedback_space != 0)
      apex_utils::error("common_latent_space / common_feedback_space are not supported by the GPU trainer");
    ustart_ = (mtype_.format_type == svd_type::USER_GROUP_FORMAT) ? mp_.num_ufeedback : 0;
    rows_total_ = (size_t)ustart_ + (size_t)mp_.num_user + (size_t)mp_.num_item;
    pitch_ = ((mp_.num_factor + 3) >> 2) << 2;
    ui_bias_.assign(rows_total_ ? rows_total_ : 1, 0.0f);
    W_.assign((rows_total_ ? rows_total_ : 1) * (size_t)(pitch_ ? pitch_ : 1), 0.0f);
    g_bias_.assign(mp_.num_global > 0 ? (size_t)mp_.num_global : 1, 0.0f);
    space_allocated_ = 1;
  }
  void ensure_handle() {
    svdgpu_shape s;
    s.num_user = mp_.num_user;
    s.num_item = mp_.num_item;
    s.num_global = mp_.num_global;
    s.num_ufeedback = mp_.num_ufeedback;
    s.num_factor = mp_.num_factor;
    s.no_user_bias = mp_.no_user_bias;
    s.active_type = mtype_.active_type;
    s.format_type = svdpp_ ? 1 : 0;
    if (h_ && !memcmp(&s, &shape_, sizeof(s))) return;
    if (h_) svdgpu_destroy(h_);
    h_ = NULL;
    if (svdpp_ && mtype_.format_type != svd_type::USER_GROUP_FORMAT)
      apex_utils::error("GPU trainer: extend_type=1 needs format_type=1 (user-grouped input)");
    if (svdgpu_create(&h_, &s, device_) != 0) apex_utils::error(svdgpu_last_error(NULL));
    shape_ = s;
    check(h_, svdgpu_set_mode(h_, mode_));
    for (size_t i = 0; i < options_.size(); ++i)
      check(h_, svdgpu_set_option(h_, options_[i].first.c_str(), options_[i].second));
    push_hparams();
  }
  // SparseFeatureArray<float>::load (apex-utils/apex_utils.h:176-195): per line "n idx:val x n"
  void load_side_features(int which, const std::string &fname) {
    if (fname == "NULL") return;
    FILE *fi = fopen(fname.c_str(), "r");
    if (!fi) {
      fprintf(stderr, "can not open file \"%s\"\n", fname.c_str());
      exit(-1);
    }
    std::vector<unsigned> rp(1, 0u), idx;
    std::vector<float> val;
    int n;
    while (fscanf(fi, "%d", &n) == 1) {
      rp.push_back(rp.back() + (unsigned)n);
      for (int i = 0; i < n; ++i) {
        unsigned id;
        float v;
        apex_utils::assert_true(fscanf(fi, "%u:%f", &id, &v) == 2, "load sparse feature");
        idx.push_back(id);
        val.push_back(v);
      }
    }
    fclose(fi);
    check(h_, svdgpu_set_side_features(h_, which, (int)rp.size() - 1, rp.data(), idx.data(), val.data()));
  }
  // ---- multi-GPU (gpu:world > 1) ---------------------------------------------------------
  bool mine(const SVDFeatureCSR::Elem &e) const {
    if (world_ <= 1) return true;
    const unsigned u = e.num_ufactor > 0 ? e.index_ufactor[0] : 0u;  // rows without a user feature: rank 0
    return (int)(u % (unsigned)world_) == rank_;
  }
  // `n` more rows of the COMMON input stream have gone by: every process counts all of them, so
  // the exchange points are the same everywhere whatever share of the rows a process keeps
  void tick(long long n = 1) {
    if (world_ <= 1) return;
    const long long before = seen_ / xchg_rows_;
    seen_ += n;
    if (seen_ / xchg_rows_ != before) {
      flush();
      exchange();
    }
  }
  void exchange() {
    if (world_ <= 1 || !h_ || !comm_on_) return;
    check(h_, svdgpu_allreduce_items(h_, xchg_scale_ < 0.0f ? 1.0f / (float)world_ : xchg_scale_));
  }
  // rank 0 makes the NCCL id and leaves it in the file gpu:nccl_id names; the others wait for it
  void join_comm() {
    if (world_ <= 1 || comm_on_) return;
    apex_utils::assert_true(rank_ >= 0 && rank_ < world_, "gpu:rank must be in [0, gpu:world)");
    apex_utils::assert_true(!nccl_id_file_.empty(), "gpu:world > 1 needs gpu:nccl_id=<file shared by the processes>");
    char id[128];
    if (rank_ == 0) {
      check(h_, svdgpu_comm_id(id));
      const std::string tmp = nccl_id_file_ + ".tmp";
      FILE *fo = fopen(tmp.c_str(), "wb");
      apex_utils::assert_true(fo != NULL, "can not write gpu:nccl_id file");
      fwrite(id, 1, sizeof(id), fo);
      fclose(fo);
      apex_utils::assert_true(rename(tmp.c_str(), nccl_id_file_.c_str()) == 0, "can not publish gpu:nccl_id file");
    } else {
      bool got = false;
      for (int t = 0; t < 6000 && !got; ++t) {  // up to 10 minutes
        FILE *fi = fopen(nccl_id_file_.c_str(), "rb");
        if (fi) {
          got = fread(id, 1, sizeof(id), fi) == sizeof(id);
          fclose(fi);
        }
        if (!got) usleep(100000);
      }
      apex_utils::assert_true(got, "gpu:nccl_id file did not appear");
    }
    check(h_, svdgpu_comm_init(h_, world_, rank_, id));
    comm_on_ = true;
    exchange();  // (first call: snapshot of the replicated slabs)
  }

  void push_hparams() {
    hp_.base_score = mp_.base_score;
    hp_.user_nonnegative = mp_.user_nonnegative;
    check(h_, svdgpu_set_hparams(h_, &hp_));
    const RangedWd *sets[3] = {&u_param_, &i_param_, &g_param_};
    for (int w = 0; w < 3; ++w)  // get_wd consults the ranges only when a bound was given (base.h:70)
      check(h_, svdgpu_set_wd_ranges(h_, w, (int)sets[w]->bound.size(), sets[w]->bound.data(), sets[w]->wd.data()));
  }
  void upload() { check(h_, svdgpu_upload_model(h_, ui_bias_.data(), W_.data(), (size_t)pitch_, g_bias_.data())); }

  void push_block(const SVDPlusBlock &b) { stage_block(b, b.index_ufeedback, b.value_ufeedback, b.num_ufeedback); }
  void stage_block(const SVDPlusBlock &b, const unsigned *fbi, const float *fbv, int nfb) {
    fb_index_.insert(fb_index_.end(), fbi, fbi + nfb);
    fb_value_.insert(fb_value_.end(), fbv, fbv + nfb);
    blk_fb_off_.push_back((int)fb_index_.size());
    for (int r = 0; r < b.data.num_row; ++r) rows_.push(b.data[r]);
    blk_row_off_.push_back(rows_.num_row());
    blk_tag_.push_back(b.extend_tag);
  }
  // drop the first nb staged blocks, keep the (open) rest
  void keep_open_run(int nb) {
    const int r0 = blk_row_off_[(size_t)nb], f0 = blk_fb_off_[(size_t)nb], v0 = rows_.row_ptr[3 * (size_t)r0];
    CsrStage rest;
    const size_t nr = rows_.nrow_, nv = rows_.nval_;
    rest.label.assign(rows_.label.begin() + r0, rows_.label.begin() + nr);
    rest.index.assign(rows_.index.begin() + v0, rows_.index.begin() + nv);
    rest.value.assign(rows_.value.begin() + v0, rows_.value.begin() + nv);
    rest.row_ptr.clear();
    for (size_t i = 3 * (size_t)r0; i <= 3 * nr; ++i) rest.row_ptr.push_back(rows_.row_ptr[i] - v0);
    rest.nrow_ = nr - (size_t)r0;
    rest.nval_ = nv - (size_t)v0;
    rows_ = rest;
    std::vector<int> bro(1, 0), bfo(1, 0);
    for (size_t b = (size_t)nb + 1; b < blk_row_off_.size(); ++b) {
      bro.push_back(blk_row_off_[b] - r0);
      bfo.push_back(blk_fb_off_[b] - f0);
    }
    blk_row_off_ = bro;
    blk_fb_off_ = bfo;
    blk_tag_.erase(blk_tag_.begin(), blk_tag_.begin() + nb);
    fb_index_.erase(fb_index_.begin(), fb_index_.begin() + f0);
    fb_value_.erase(fb_value_.begin(), fb_value_.begin() + f0);
  }
  void clear_stage() {
    rows_.clear();
    blk_row_off_.assign(1, 0);
    blk_fb_off_.assign(1, 0);
    blk_tag_.clear();
    fb_index_.clear();
    fb_value_.clear();
  }
  void flush() {
    if (!h_) return;
    if (svdpp_) {
      // a user unit (START .. END) is trained as a whole: blocks of a run that has not been closed
      // yet stay staged until its END block arrives
      int nb = (int)blk_tag_.size();
      while (nb > 0 && blk_tag_[(size_t)nb - 1] != svdpp_tag::DEFAULT && blk_tag_[(size_t)nb - 1] != svdpp_tag::END_TAG) --nb;
      if (nb > 0)
        check(h_, svdgpu_update_ugroup(h_, nb, blk_row_off_.data(), blk_fb_off_.data(), blk_tag_.data(),
                                       fb_index_.data(), fb_value_.data(), rows_.row_ptr.data(),
                                       rows_.label.data(), rows_.index.data(), rows_.value.data()));
      if (nb < (int)blk_tag_.size()) {
        keep_open_run(nb);
        return;
      }
    } else if (rows_.num_row() > 0) {
      check(h_, svdgpu_update_csr(h_, rows_.num_row(), rows_.row_ptr.data(), rows_.label.data(),
                                  rows_.index.data(), rows_.value.data()));
    }
    clear_stage();
  }

  // tensor records, apex-tensor/apex_tensor_cpu_inline_common.h:72-88
  void write_1d(const float *p, int n, FILE *fo) {
    int32_t h = n;
    fwrite(&h, 4, 1, fo);
    if (n > 0) fwrite(p, 4, (size_t)n, fo);
  }
  void write_2d(const float *p, int y_max, FILE *fo) {
    int32_t h[2] = {mp_.num_factor, y_max};
    fwrite(h, 4, 2, fo);
    for (int y = 0; y < y_max; ++y) fwrite(p + (size_t)y * pitch_, 4, (size_t)mp_.num_factor, fo);
  }
  void read_1d(float *p, int n, FILE *fi) {
    int32_t h = 0;
    apex_utils::assert_true(fread(&h, 4, 1, fi) == 1 && h == n, "tensor::load_from_file");
    if (n > 0) apex_utils::assert_true(fread(p, 4, (size_t)n, fi) == (size_t)n, "tensor::load_from_file");
  }
  void read_2d(float *p, int y_max, FILE *fi) {
    int32_t h[2] = {0, 0};
    apex_utils::assert_true(fread(h, 4, 2, fi) == 2 && h[0] == mp_.num_factor && h[1] == y_max,
                            "tensor::load_from_file");
    for (int y = 0; y < y_max; ++y)
      if (mp_.num_factor > 0)
        apex_utils::assert_true(fread(p + (size_t)y * pitch_, 4, (size_t)mp_.num_factor, fi) == (size_t)mp_.num_factor,
                                "tensor::load_from_file");
  }

  SVDTypeParam mtype_;
  ModelHeader mp_;
  svdgpu_hparams hp_;
  svdgpu_shape shape_;
  int decay_learning_rate_ = 0;
  float decay_rate_ = 1.0f;
  int round_counter_ = 0;
  int init_end_ = 0, space_allocated_ = 0;
  bool svdpp_ = false;
  int device_ = 0, mode_ = SVDGPU_MODE_EXACT;
  int batch_rows_ = 1 << 20;
  std::vector<std::pair<std::string, long long> > options_;
  std::string name_feat_user_ = "NULL", name_feat_item_ = "NULL";
  RangedWd u_param_{"up:", "uip:"}, i_param_{"ip:", "uip:"}, g_param_{"gp:", "gp:"};  // base.h:102-103
  svdgpu_t *h_ = NULL;
  // host mirror of the model (init, load/save)
  int ustart_ = 0, pitch_ = 0;
  size_t rows_total_ = 0;
  std::vector<float> ui_bias_, W_, g_bias_;
  // staged input
  CsrStage rows_;
  std::vector<int> blk_row_off_, blk_fb_off_, blk_tag_;
  std::vector<unsigned> fb_index_;
  std::vector<float> fb_value_;
  // multi-GPU
  int world_ = 1, rank_ = 0;
  bool comm_on_ = false;
  std::string nccl_id_file_;
  long long xchg_rows_ = 1 << 22, seen_ = 0;  // exchange every this many rows of the common stream
  float xchg_scale_ = -1.0f;                  // < 0: mean over the processes (the default: tracks the sequential run at
                                              // any exchange frequency, profiles/r2_convergence_sweep.jsonl)
  // feedback list of the last DEFAULT / START block handed to predict (base.h:583-591)
  std::vector<unsigned> pred_fb_index_;
  std::vector<float> pred_fb_value_;
};

// apex_svd.cpp:32-45 -- the link-time seam
ISVDTrainer *create_svd_trainer(SVDTypeParam mtype) {
  if (mtype.extend_type == 2 || mtype.extend_type == 15 || mtype.extend_type == 30 || mtype.extend_type == 31)
    apex_utils::error("GPU trainer: extend_type 2/15/30/31 solvers (multi-imfb, bilinear, gbrt) are CPU-only in the reference and not provided here");
  return new GpuSVDFeature(mtype);
}

// SVDFeatureRanker (base.h:597-813) on the GPU: the model, side features and device handle are the
// trainer's; every row goes straight to the C ABI's stream state machine (svdgpu_rank_csr), which
// ranks a user section on the device when its PROCESS row arrives.
class GpuSVDRanker : public ISVDRanker {
 public:
  explicit GpuSVDRanker(const SVDTypeParam &mtype) : tr_(mtype), mtype_(mtype) {}
  virtual void load_model(FILE *fi) { tr_.load_model(fi); }  // base.h:662-664
  virtual void set_param(const char *name, const char *val) {  // base.h:656-660 (+ the gpu: keys)
    if (!strcmp(name, "top_k")) top_k_ = atoi(val);
    if (!strcmp(name, "feature_user") || !strcmp(name, "feature_item") || !strncmp(name, "gpu:", 4))
      tr_.set_param(name, val);
  }
  virtual void init_ranker(int num_item_set) {  // base.h:666-685
    tr_.init_trainer();
    check(tr_.handle(), svdgpu_rank_init(tr_.handle(), num_item_set, top_k_));
  }
  virtual void process(std::vector<int> &result, const SVDFeatureCSR::Elem &e) {  // base.h:795-797
    CsrStage one;
    one.push(e);
    note(e);
    // (buf() may grow buf_: take the pointer and the capacity in this order, not as two arguments
    // of one call, whose evaluation order is unspecified)
    int *out = buf();
    const long long cap = (long long)buf_.size();
    collect(result, svdgpu_rank_csr(tr_.handle(), 1, one.row_ptr.data(), one.label.data(), one.index.data(),
                                    one.value.data(), out, cap, &n_));
  }
  virtual void process(std::vector<int> &result, const SVDPlusBlock &b) {  // base.h:798-812
    CsrStage rows;
    for (int r = 0; r < b.data.num_row; ++r) {
      rows.push(b.data[r]);
      note(b.data[r]);
    }
    const int bro[2] = {0, rows.num_row()}, bfo[2] = {0, b.num_ufeedback}, tag[1] = {b.extend_tag};
    int *out = buf();
    const long long cap = (long long)buf_.size();
    collect(result, svdgpu_rank_ugroup(tr_.handle(), 1, bro, bfo, tag, b.index_ufeedback, b.value_ufeedback,
                                       rows.row_ptr.data(), rows.label.data(), rows.index.data(), rows.value.data(),
                                       out, cap, &n_));
  }

 private:
  // results of one call never exceed top_k per PROCESS row or the POS entries seen since the last USER row
  void note(const SVDFeatureCSR::Elem &e) {
    const int tag = (int)e.label;
    if (tag == svdranker_tag::USER_TAG) pending_ = 0;
    if (tag == svdranker_tag::POS_SAMPLE) pending_ += e.num_ufactor;
    if (tag == svdranker_tag::PROCESS_TAG) need_ += top_k_ > 0 ? top_k_ : pending_;
  }
  int *buf() {
    if (buf_.size() < (size_t)need_ + 16) buf_.resize((size_t)need_ + 16);
    need_ = 0;
    return buf_.data();
  }
  void collect(std::vector<int> &result, int rc) {
    check(tr_.handle(), rc);
    result.insert(result.end(), buf_.begin(), buf_.begin() + n_);
  }
  GpuSVDFeature tr_;
  SVDTypeParam mtype_;
  int top_k_ = 0, pending_ = 0, need_ = 0;
  long long n_ = 0;
  std::vector<int> buf_;
};

// apex_svd.cpp:44-46
ISVDRanker *create_svd_ranker(SVDTypeParam mtype) {
  return new GpuSVDRanker(mtype);
}

}  // namespace apex_svd

// ---- bulk C entry points (GPU trainer only; same handle as trainer_cabi.cpp) ----
namespace {
struct ShimHandle {  // must match trainer_cabi.cpp's Handle
  apex_svd::SVDTypeParam mtype;
  apex_svd::ISVDTrainer *tr;
};
apex_svd::GpuSVDFeature *gpu_of(void *hv) {
  return static_cast<apex_svd::GpuSVDFeature *>(static_cast<ShimHandle *>(hv)->tr);
}
apex_svd::SVDFeatureCSR view(int num_row, const int *row_ptr, const float *label, const unsigned *index,
                             const float *value) {
  apex_svd::SVDFeatureCSR c;
  c.num_row = num_row;
  c.num_val = num_row > 0 ? row_ptr[3 * (size_t)num_row] - row_ptr[0] : 0;
  c.row_ptr = const_cast<int *>(row_ptr);
  c.row_label = const_cast<float *>(label);
  c.feat_index = const_cast<unsigned *>(index);
  c.feat_value = const_cast<float *>(value);
  return c;
}
}  // namespace

extern "C" {
void svdtr_update_csr_bulk(void *hv, int num_row, const int *row_ptr, const float *label,
                           const unsigned *index, const float *value) {
  gpu_of(hv)->update_batch(view(num_row, row_ptr, label, index, value));
}
void svdtr_predict_csr_bulk(void *hv, int num_row, const int *row_ptr, const float *label,
                            const unsigned *index, const float *value, float *out) {
  gpu_of(hv)->predict_batch(view(num_row, row_ptr, label, index, value), out);
}
void svdtr_update_ugroup_bulk(void *hv, int num_block, const int *blk_row_off, const int *blk_fb_off,
                              const int *blk_tag, const unsigned *fb_index, const float *fb_value,
                              const int *row_ptr, const float *label, const unsigned *index,
                              const float *value) {
  const int n = num_block > 0 ? blk_row_off[num_block] : 0;
  gpu_of(hv)->update_ugroup_batch(num_block, blk_row_off, blk_fb_off, blk_tag, fb_index, fb_value,
                                  view(n, row_ptr, label, index, value));
}
void svdtr_predict_ugroup_bulk(void *hv, int num_block, const int *blk_row_off, const int *blk_fb_off,
                               const int *blk_tag, const unsigned *fb_index, const float *fb_value,
                               const int *row_ptr, const float *label, const unsigned *index,
                               const float *value, float *out) {
  const int n = num_block > 0 ? blk_row_off[num_block] : 0;
  gpu_of(hv)->predict_ugroup_batch(num_block, blk_row_off, blk_fb_off, blk_tag, fb_index, fb_value,
                                   view(n, row_ptr, label, index, value), out);
}
void svdtr_sync(void *hv) { gpu_of(hv)->sync(); }
void *svdtr_gpu_handle(void *hv) { return gpu_of(hv)->handle(); }
}
