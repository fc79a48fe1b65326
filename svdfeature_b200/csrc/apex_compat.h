// apex_compat.h -- stand-alone mirror of the reference's solver seam.
//
// The GPU trainer (gpu_trainer.cpp) implements the reference's plug-in
// interface `apex_svd::ISVDTrainer`.  When the reference tree is available the
// trainer is compiled straight against the reference's own `apex_svd.h`
// (-DSVDGPU_WITH_REFERENCE_HEADERS -I<reference>), which is what makes it a
// link-time drop-in for `apex_svd.o`.  On a machine
// without the reference tree (the GPU box) the same source is compiled against
// this header instead; it declares only the types the seam needs, with the
// same names, field order and virtual-method order, so objects built either
// way have the same ABI.
//
// What is mirrored (reference file:line, relative to the reference root):
//   apex_svd::SVDFeatureCSR / ::Elem      apex_svd_data.h:34-231
//   apex_svd::svdpp_tag                   apex_svd_data.h:353-371
//   apex_svd::SVDPlusBlock                apex_svd_data.h:376-466
//   apex_svd::svd_type / SVDTypeParam     apex_svd_model.h:50-57, 242-287
//   apex_svd::ISVDTrainer                 apex_svd.h:33-107
//   apex_svd::svdranker_tag / ISVDRanker  apex_svd.h:115-152, 160-197
//   apex_svd::create_svd_trainer/_ranker  apex_svd.h:212,222
//   apex_utils::error / assert_true       apex-utils/apex_utils.h:47-58
//
// Nothing here computes; the arithmetic lives in svdgpu_kernels.cu.
#ifndef SVDGPU_APEX_COMPAT_H_
#define SVDGPU_APEX_COMPAT_H_

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace apex_utils {
// errors are fatal, reported on stderr: the reference has no exceptions and
// no return codes on this path.
inline void error(const char *msg) {
  std::fprintf(stderr, "%s\n", msg);
  std::exit(-1);
}
inline void assert_true(bool ok, const char *msg) {
  if (!ok) error(msg);
}
}  // namespace apex_utils

namespace apex_svd {

namespace svd_type {
const int RANDOM_ORDER_FORMAT = 0;
const int USER_GROUP_FORMAT = 1;
const int AUTO_DETECT = 2;
}  // namespace svd_type

namespace svdpp_tag {
const int DEFAULT = 0;
const int START_TAG = 1;
const int END_TAG = 2;
const int MIDDLE_TAG = 3;
}  // namespace svdpp_tag

// One batch of training rows.  Row r owns three consecutive segments of
// feat_index/feat_value: global [row_ptr[3r],row_ptr[3r+1]), user
// [row_ptr[3r+1],row_ptr[3r+2]), item [row_ptr[3r+2],row_ptr[3r+3]).
struct SVDFeatureCSR {
  struct Elem {
    float label;
    int num_global;
    int num_ufactor;
    int num_ifactor;
    unsigned *index_global;
    unsigned *index_ufactor;
    unsigned *index_ifactor;
    float *value_global;
    float *value_ufactor;
    float *value_ifactor;
    inline int total_num() const { return num_global + num_ufactor + num_ifactor; }
  };
  int num_row;
  int num_val;
  float *row_label;
  int *row_ptr;
  unsigned *feat_index;
  float *feat_value;

  inline Elem operator[](int r) const {
    const int *p = row_ptr + 3 * r;
    Elem e;
    e.label = row_label[r];
    e.num_global = p[1] - p[0];
    e.num_ufactor = p[2] - p[1];
    e.num_ifactor = p[3] - p[2];
    e.index_global = feat_index + p[0];
    e.value_global = feat_value + p[0];
    e.index_ufactor = feat_index + p[1];
    e.value_ufactor = feat_value + p[1];
    e.index_ifactor = feat_index + p[2];
    e.value_ifactor = feat_value + p[2];
    return e;
  }
};

// One user's rows plus that user's implicit-feedback list.
struct SVDPlusBlock {
  int num_ufeedback;
  int extend_tag;
  int extra_info;
  unsigned *index_ufeedback;
  float *value_ufeedback;
  SVDFeatureCSR data;
  SVDPlusBlock() : extend_tag(svdpp_tag::DEFAULT), extra_info(0) {}
};

// Four bytes that open every model file.
struct SVDTypeParam {
  uint8_t format_type;
  uint8_t active_type;
  uint8_t extend_type;
  uint8_t variant_type;
  SVDTypeParam()
      : format_type(svd_type::AUTO_DETECT), active_type(0), extend_type(0), variant_type(0) {}
  inline void set_param(const char *name, const char *val) {
    const uint8_t v = (uint8_t)std::atoi(val);
    if (!std::strcmp(name, "model_type") || !std::strcmp(name, "format_type")) format_type = v;
    if (!std::strcmp(name, "active_type")) active_type = v;
    if (!std::strcmp(name, "extend_type")) extend_type = v;
    if (!std::strcmp(name, "variant_type")) variant_type = v;
  }
  inline void decide_format(int fmt = svd_type::AUTO_DETECT) {
    if (format_type != svd_type::AUTO_DETECT) return;
    format_type = (uint8_t)fmt;
    if (format_type != svd_type::AUTO_DETECT) return;
    format_type = (uint8_t)(extend_type == 0 ? svd_type::RANDOM_ORDER_FORMAT
                                             : svd_type::USER_GROUP_FORMAT);
  }
};

// The plug-in interface.  Virtual order must not change: it is the vtable the
// reference's drivers call through.
class ISVDTrainer {
 public:
  virtual void set_param(const char *name, const char *val) = 0;
  virtual void load_model(FILE *fi) = 0;
  virtual void save_model(FILE *fo) = 0;
  virtual void init_model(void) = 0;
  virtual void init_trainer(void) = 0;

 public:
  virtual void set_round(int nround) { apex_utils::error("not implemented 3"); }
  virtual void finish_round(void) {}
  virtual void update(const SVDFeatureCSR::Elem &feature) { apex_utils::error("not implemented 2"); }
  virtual float predict(const SVDFeatureCSR::Elem &feature) {
    apex_utils::error("not implemented 1");
    return 0.0f;
  }

 public:
  virtual void update(const SVDPlusBlock &data) { apex_utils::error("not implemented"); }
  virtual void predict(std::vector<float> &pred, const SVDPlusBlock &data) {
    apex_utils::error("not implemented");
  }

 public:
  virtual ~ISVDTrainer() {}
};

// Tags of the ranker's input rows, carried in the label field (apex_svd.h:115-152).
namespace svdranker_tag {
const int ITEM_TAG = 0;
const int USER_TAG = 2;
const int POS_SAMPLE = 1;
const int BAN_SAMPLE = -1;
const int SPEC_SAMPLE = 3;
const int PROCESS_TAG = 4;
}  // namespace svdranker_tag

// Ranking utility interface (apex_svd.h:160-197); implemented by GpuSVDRanker (gpu_trainer.cpp).
class ISVDRanker {
 public:
  virtual void load_model(FILE *fi) = 0;
  virtual void init_ranker(int num_item_set) = 0;
  virtual void set_param(const char *name, const char *val) = 0;

 public:
  virtual void process(std::vector<int> &result, const SVDFeatureCSR::Elem &feature) {
    apex_utils::error("not implemented");
  }
  virtual void process(std::vector<int> &result, const SVDPlusBlock &data) {
    apex_utils::error("not implemented");
  }

 public:
  virtual ~ISVDRanker() {}
};

ISVDTrainer *create_svd_trainer(SVDTypeParam mtype);
ISVDRanker *create_svd_ranker(SVDTypeParam mtype);

}  // namespace apex_svd

#endif  // SVDGPU_APEX_COMPAT_H_
