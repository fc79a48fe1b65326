/* svdf_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded CPU restatement of the reference's SGD hot path
 * (SVDFeature / SVDPPFeature in solvers/base-solver/apex_svd_base.h) used as
 * the parity oracle for the CUDA path.  Nothing under svdfeature_b200/ may
 * include, link or call this.  Pinned bit-for-bit against the compiled
 * reference (oracle/_ref/libsvdf_ref.so) by tests/test_oracle_vs_ref.py and
 * against committed golden vectors by tests/test_oracle_golden.py.
 */
#ifndef SVDF_ORACLE_H_
#define SVDF_ORACLE_H_
#ifdef __cplusplus
extern "C" {
#endif

typedef struct svdo svdo_t;

svdo_t *svdo_create(int format_type, int active_type, int extend_type);
void svdo_destroy(svdo_t *m);
void svdo_seed(unsigned seed);
/* same (name, value) strings the reference's set_param chain understands */
void svdo_set_param(svdo_t *m, const char *name, const char *val);
void svdo_init_model(svdo_t *m);
void svdo_init_trainer(svdo_t *m);
void svdo_set_round(svdo_t *m, int nround);
int svdo_save_model(svdo_t *m, const char *path);
int svdo_load_model(svdo_t *m, const char *path);

void svdo_update_csr(svdo_t *m, int num_row, const int *row_ptr, const float *label,
                     const unsigned *index, const float *value);
void svdo_predict_csr(svdo_t *m, int num_row, const int *row_ptr, const float *label,
                      const unsigned *index, const float *value, float *out);
void svdo_update_ugroup(svdo_t *m, int num_block, const int *blk_row_off, const int *blk_fb_off,
                        const int *blk_tag, const unsigned *fb_index, const float *fb_value,
                        const int *row_ptr, const float *label, const unsigned *index,
                        const float *value);
void svdo_predict_ugroup(svdo_t *m, int num_block, const int *blk_row_off, const int *blk_fb_off,
                         const int *blk_tag, const unsigned *fb_index, const float *fb_value,
                         const int *row_ptr, const float *label, const unsigned *index,
                         const float *value, float *out);

/* SVDFeatureRanker (base.h:597-813): load_model, set_param("top_k" ...), init_ranker, then a tagged
 * instance stream; results (item indices for top_k > 0, else rank positions of the POS items) are
 * appended in stream order.  Return value: number of results (beyond cap: counted, not stored). */
void svdo_init_ranker(svdo_t *m, int num_item_set);
long svdo_rank_csr(svdo_t *m, int num_row, const int *row_ptr, const float *label, const unsigned *index,
                   const float *value, int *result, long cap);
long svdo_rank_ugroup(svdo_t *m, int num_block, const int *blk_row_off, const int *blk_fb_off,
                      const int *blk_tag, const unsigned *fb_index, const float *fb_value,
                      const int *row_ptr, const float *label, const unsigned *index, const float *value,
                      int *result, long cap);

/* raw views for tests: which = 0 ui_bias, 1 W_uiset, 2 g_bias */
float *svdo_data(svdo_t *m, int which);
/* what = 0 rows of W_uiset, 1 pitch in floats, 2 ustart (feedback rows), 3 num_factor,
 * 4 num_user, 5 num_item, 6 num_global, 7 num_ufeedback */
long svdo_info(svdo_t *m, int what);
float svdo_base_score(svdo_t *m);
float svdo_learning_rate(svdo_t *m);

#ifdef __cplusplus
}
#endif
#endif
