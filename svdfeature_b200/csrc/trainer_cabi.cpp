// trainer_cabi.cpp -- flat C entry points over apex_svd::ISVDTrainer.
//
// The reference's drivers (svd_feature.cpp:194-288, svd_feature_infer.cpp)
// talk to a solver only through ISVDTrainer (apex_svd.h:33-107) obtained from
// create_svd_trainer (apex_svd.h:212).  This file wraps exactly that seam in
// `extern "C"` functions so that Python (ctypes) can drive ANY implementation
// of the seam the same way:
//
//   * linked with the reference's own apex_svd.cpp  -> oracle/_ref/libsvdf_ref.so
//     (the unmodified CPU reference; test oracle and CPU baseline only)
//   * linked with gpu_trainer.cpp                    -> libsvdf_gpu.so
//     (the B200 trainer; the product)
//
// Every call below maps 1:1 to one ISVDTrainer virtual; the batch helpers
// simply loop the per-row virtual the way svd_feature.cpp:231-247 does.
#ifdef SVDGPU_WITH_REFERENCE_HEADERS
#include "apex_svd.h"
#else
#include "apex_compat.h"
#endif

#include <cstdio>
#include <cstdlib>
#include <vector>

using apex_svd::ISVDRanker;
using apex_svd::ISVDTrainer;
using apex_svd::SVDFeatureCSR;
using apex_svd::SVDPlusBlock;
using apex_svd::SVDTypeParam;

namespace {
struct Handle {
  SVDTypeParam mtype;
  ISVDTrainer *tr;
};
struct RankHandle {
  SVDTypeParam mtype;
  ISVDRanker *rk;
};

inline SVDFeatureCSR make_csr(int num_row, const int *row_ptr, const float *label,
                              const unsigned *index, const float *value) {
  SVDFeatureCSR c;
  c.num_row = num_row;
  c.num_val = num_row > 0 ? row_ptr[3 * num_row] - row_ptr[0] : 0;
  c.row_ptr = const_cast<int *>(row_ptr);
  c.row_label = const_cast<float *>(label);
  c.feat_index = const_cast<unsigned *>(index);
  c.feat_value = const_cast<float *>(value);
  return c;
}

inline SVDPlusBlock make_block(int b, const int *blk_row_off, const int *blk_fb_off,
                               const int *blk_tag, const unsigned *fb_index, const float *fb_value,
                               const int *row_ptr, const float *label, const unsigned *index,
                               const float *value) {
  SVDPlusBlock blk;
  const int r0 = blk_row_off[b], r1 = blk_row_off[b + 1];
  blk.num_ufeedback = blk_fb_off[b + 1] - blk_fb_off[b];
  blk.extend_tag = blk_tag ? blk_tag[b] : apex_svd::svdpp_tag::DEFAULT;
  blk.index_ufeedback = const_cast<unsigned *>(fb_index) + blk_fb_off[b];
  blk.value_ufeedback = const_cast<float *>(fb_value) + blk_fb_off[b];
  blk.data = make_csr(r1 - r0, row_ptr + 3 * r0, label + r0, index, value);
  return blk;
}
}  // namespace

extern "C" {

// apex_random::seed (apex-tensor/apex_random.h:40-43) is srand(); the CLI
// seeds with 10 (svd_feature.cpp:293).
void svdtr_seed(unsigned seed) { srand(seed); }

// create_svd_trainer(mtype); format_type is resolved like svd_feature.cpp:118-125.
void *svdtr_create(int format_type, int active_type, int extend_type) {
  Handle *h = new Handle();
  h->mtype.format_type = (uint8_t)format_type;
  h->mtype.active_type = (uint8_t)active_type;
  h->mtype.extend_type = (uint8_t)extend_type;
  h->mtype.decide_format();
  h->tr = apex_svd::create_svd_trainer(h->mtype);
  return h;
}

void svdtr_destroy(void *hv) {
  Handle *h = static_cast<Handle *>(hv);
  if (!h) return;
  delete h->tr;
  delete h;
}

void svdtr_set_param(void *hv, const char *name, const char *val) {
  static_cast<Handle *>(hv)->tr->set_param(name, val);
}
void svdtr_init_model(void *hv) { static_cast<Handle *>(hv)->tr->init_model(); }
void svdtr_init_trainer(void *hv) { static_cast<Handle *>(hv)->tr->init_trainer(); }
void svdtr_set_round(void *hv, int r) { static_cast<Handle *>(hv)->tr->set_round(r); }
void svdtr_finish_round(void *hv) { static_cast<Handle *>(hv)->tr->finish_round(); }

// model file = [SVDTypeParam 4 B][ISVDTrainer::save_model], svd_feature.cpp:184-191
int svdtr_save_model(void *hv, const char *path) {
  Handle *h = static_cast<Handle *>(hv);
  FILE *fo = std::fopen(path, "wb");
  if (!fo) return -1;
  std::fwrite(&h->mtype, sizeof(SVDTypeParam), 1, fo);
  h->tr->save_model(fo);
  std::fclose(fo);
  return 0;
}

// svd_feature.cpp:175-182 (skip the 4-byte type header, then load_model)
int svdtr_load_model(void *hv, const char *path) {
  Handle *h = static_cast<Handle *>(hv);
  FILE *fi = std::fopen(path, "rb");
  if (!fi) return -1;
  SVDTypeParam t;
  if (std::fread(&t, sizeof(SVDTypeParam), 1, fi) != 1) {
    std::fclose(fi);
    return -2;
  }
  h->tr->load_model(fi);
  std::fclose(fi);
  return 0;
}

// for each row: update(Elem) -- the hot loop of svd_feature.cpp:231-247
void svdtr_update_csr(void *hv, int num_row, const int *row_ptr, const float *label,
                      const unsigned *index, const float *value) {
  ISVDTrainer *tr = static_cast<Handle *>(hv)->tr;
  const SVDFeatureCSR c = make_csr(num_row, row_ptr, label, index, value);
  for (int r = 0; r < num_row; ++r) tr->update(c[r]);
}

void svdtr_predict_csr(void *hv, int num_row, const int *row_ptr, const float *label,
                       const unsigned *index, const float *value, float *out) {
  ISVDTrainer *tr = static_cast<Handle *>(hv)->tr;
  const SVDFeatureCSR c = make_csr(num_row, row_ptr, label, index, value);
  for (int r = 0; r < num_row; ++r) out[r] = tr->predict(c[r]);
}

// user-grouped input: block b owns rows [blk_row_off[b], blk_row_off[b+1]) and
// feedback entries [blk_fb_off[b], blk_fb_off[b+1]); blk_tag may be NULL.
void svdtr_update_ugroup(void *hv, int num_block, const int *blk_row_off, const int *blk_fb_off,
                         const int *blk_tag, const unsigned *fb_index, const float *fb_value,
                         const int *row_ptr, const float *label, const unsigned *index,
                         const float *value) {
  ISVDTrainer *tr = static_cast<Handle *>(hv)->tr;
  for (int b = 0; b < num_block; ++b) {
    const SVDPlusBlock blk = make_block(b, blk_row_off, blk_fb_off, blk_tag, fb_index, fb_value,
                                        row_ptr, label, index, value);
    tr->update(blk);
  }
}

void svdtr_predict_ugroup(void *hv, int num_block, const int *blk_row_off, const int *blk_fb_off,
                          const int *blk_tag, const unsigned *fb_index, const float *fb_value,
                          const int *row_ptr, const float *label, const unsigned *index,
                          const float *value, float *out) {
  ISVDTrainer *tr = static_cast<Handle *>(hv)->tr;
  std::vector<float> p;
  for (int b = 0; b < num_block; ++b) {
    const SVDPlusBlock blk = make_block(b, blk_row_off, blk_fb_off, blk_tag, fb_index, fb_value,
                                        row_ptr, label, index, value);
    tr->predict(p, blk);
    for (size_t i = 0; i < p.size(); ++i) out[blk_row_off[b] + (int)i] = p[i];
  }
}

// ---- ISVDRanker (apex_svd.h:160-197), driven like svd_feature_infer.cpp:126-154,233-241,349-371 ----
// The model file's 4-byte SVDTypeParam decides the ranker's type (svd_feature_infer.cpp:128-137).
void *svdrk_create_from_model(const char *path) {
  FILE *fi = std::fopen(path, "rb");
  if (!fi) return NULL;
  RankHandle *h = new RankHandle();
  if (std::fread(&h->mtype, sizeof(SVDTypeParam), 1, fi) != 1) {
    std::fclose(fi);
    delete h;
    return NULL;
  }
  h->rk = apex_svd::create_svd_ranker(h->mtype);
  h->rk->load_model(fi);
  std::fclose(fi);
  return h;
}
void svdrk_destroy(void *hv) {
  RankHandle *h = static_cast<RankHandle *>(hv);
  if (!h) return;
  delete h->rk;
  delete h;
}
void svdrk_set_param(void *hv, const char *name, const char *val) {
  static_cast<RankHandle *>(hv)->rk->set_param(name, val);
}
void svdrk_init_ranker(void *hv, int num_item_set) { static_cast<RankHandle *>(hv)->rk->init_ranker(num_item_set); }

// for each row: process(result, Elem) -- the loop of svd_feature_infer.cpp:352-360; results are
// appended in stream order; returns their number (those beyond cap are counted, not stored)
long svdrk_rank_csr(void *hv, int num_row, const int *row_ptr, const float *label, const unsigned *index,
                    const float *value, int *result, long cap) {
  ISVDRanker *rk = static_cast<RankHandle *>(hv)->rk;
  const SVDFeatureCSR c = make_csr(num_row, row_ptr, label, index, value);
  long n = 0;
  std::vector<int> p;
  for (int r = 0; r < num_row; ++r) {
    p.clear();
    rk->process(p, c[r]);
    for (size_t i = 0; i < p.size(); ++i, ++n)
      if (n < cap) result[n] = p[i];
  }
  return n;
}
// for each block: process(result, SVDPlusBlock) -- svd_feature_infer.cpp:362-370
long svdrk_rank_ugroup(void *hv, int num_block, const int *blk_row_off, const int *blk_fb_off, const int *blk_tag,
                       const unsigned *fb_index, const float *fb_value, const int *row_ptr, const float *label,
                       const unsigned *index, const float *value, int *result, long cap) {
  ISVDRanker *rk = static_cast<RankHandle *>(hv)->rk;
  long n = 0;
  std::vector<int> p;
  for (int b = 0; b < num_block; ++b) {
    const SVDPlusBlock blk = make_block(b, blk_row_off, blk_fb_off, blk_tag, fb_index, fb_value, row_ptr, label,
                                        index, value);
    p.clear();
    rk->process(p, blk);
    for (size_t i = 0; i < p.size(); ++i, ++n)
      if (n < cap) result[n] = p[i];
  }
  return n;
}

}  // extern "C"
