/* svdf_oracle.c -- TEST INFRASTRUCTURE ONLY (see svdf_oracle.h).
 *
 * CPU restatement, in plain C, of the reference's per-instance SGD step.
 * Every function names the reference code it follows (paths relative to the
 * reference root; "base.h" = solvers/base-solver/apex_svd_base.h, "model.h" =
 * apex_svd_model.h, "sse.h" = apex-tensor/apex_tensor_sse.h).
 *
 * Arithmetic contract reproduced here (and by the CUDA "exact" kernels):
 *   - model state and row arithmetic are fp32; every row op is a separate
 *     multiply followed by a separate add (the reference is built -msse2:
 *     mulps/addps, never FMA) -> compile this file with -ffp-contract=off;
 *   - bias + dot are accumulated in fp64 and cast to fp32 once (base.h:317,446);
 *   - a row times a scalar s with |s-1| <= 1e-6 is NOT multiplied (sse.h:231-242);
 *   - the dot product keeps 4 lane-strided partial sums and combines them as
 *     (l0+l2)+(l1+l3), then adds the k%4 tail serially (sse.h:88-97,289-317).
 *
 * Pinning: tests/test_oracle_vs_ref.py trains this oracle and the compiled
 * reference (oracle/_ref) on the same seeded inputs and demands identical
 * model bytes; tests/test_oracle_golden.py checks committed reference outputs.
 */
#include "svdf_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { FMT_RANDOM_ORDER = 0, FMT_USER_GROUP = 1, FMT_AUTO = 2 };
enum { TAG_DEFAULT = 0, TAG_START = 1, TAG_END = 2, TAG_MIDDLE = 3 };

/* base.h:33-75 ParameterSet: per-index-range weight decay ("up:", "ip:", "gp:") */
typedef struct {
  char prefix_a[16], prefix_b[16];
  int n_wd, n_bound;
  float wd[64];
  unsigned bound[64];
} range_wd_t;

struct svdo {
  /* model.h:242-261 SVDTypeParam */
  int format_type, active_type, extend_type;
  /* model.h:373-450 SVDModelParam (fields used on the path) */
  int num_user, num_item, num_factor, num_global;
  float u_init_sigma, i_init_sigma, base_score;
  int no_user_bias, num_ufeedback;
  float ufeedback_init_sigma;
  int num_randinit_ufactor, num_randinit_ifactor;
  int common_latent_space, user_nonnegative, common_feedback_space, extend_flag, item_nonnegative;
  /* model.h:291-344 SVDTrainParam */
  float learning_rate;
  int decay_learning_rate;
  float decay_rate, min_learning_rate;
  float wd_user, wd_item, wd_user_bias, wd_item_bias;
  int reg_method;
  float wd_global;
  int reg_global;
  unsigned num_regfree_global;
  float scale_lr_ufeedback, wd_ufeedback_user, wd_ufeedback, wd_ufeedback_bias;
  range_wd_t u_param, i_param, g_param;
  /* storage, model.h:511-556: one slab, rows = ustart + num_user + num_item */
  int space_allocated;
  int ustart, pitch; /* pitch in floats: row bytes rounded up to 16 (sse.h:26-27) */
  long rows;
  float *ui_bias, *W_uiset, *g_bias;
  float *u_bias, *W_user, *i_bias, *W_item, *ufeedback_bias, *W_ufeedback;
  /* trainer scratch, base.h:85,486-488 */
  int init_end, round_counter;
  unsigned sample_counter;
  unsigned *ref_user, *ref_item, *ref_global;
  /* side features (base.h:98-99; apex-utils/apex_utils.h:141-196 SparseFeatureArray<float>) */
  char name_feat_user[256], name_feat_item[256];
  struct { unsigned num_row; unsigned *row_ptr; unsigned *index; float *value; } feat_user, feat_item;
  float *tmp_ufactor, *tmp_ifactor, *tmp_ufeedback, *old_ufeedback;
  float norm_ufeedback, tmp_ufeedback_bias, old_ufeedback_bias;
  int is_svdpp; /* apex_svd.cpp:32-45: extend_type==1 or USER_GROUP -> SVDPPFeature */
  /* SVDFeatureRanker state (base.h:605-631) */
  struct {
    int init_end, num_item_set, num_item_processed, top_k;
    float *tmp_ifactors, *bias_ifactors; /* [num_item_set][pitch], [num_item_set] */
    float *tmp_ufactor, *tmp_ifactor, *tmp_ufeedback, *item_score;
    int *item_tag, *pos_item, num_pos, cap_pos;
  } rk;
};

static void die(const char *msg) { /* apex-utils/apex_utils.h:47-50 */
  fprintf(stderr, "%s\n", msg);
  exit(-1);
}

/* ---------------- tensor primitives (apex-tensor) ---------------- */

/* sse.h:231-242 ScalarOptimizer<Mul>: scalar "equal to one" test */
static int scalar_is_one(float s) { return !(fabs((double)(s - 1.0f)) > 1e-6); }

/* dst += src * s : ContainerExp::operator+=(CompositeExp) -> scalar_map<AddTo,Mul>
 * (apex_exp_template.h:388-392,471-474; sse.h:261-272) */
static void row_add_scaled(float *dst, const float *src, float s, int n) {
  int i;
  if (scalar_is_one(s)) {
    for (i = 0; i < n; ++i) dst[i] = dst[i] + src[i];
  } else {
    for (i = 0; i < n; ++i) {
      float p = src[i] * s;
      dst[i] = dst[i] + p;
    }
  }
}
/* dst *= s : operator*=(double) -> dst = dst * s -> scalar_map<SaveTo,Mul>
 * (apex_exp_template.h:340-344); a no-op when s is "one" */
static void row_scale(float *dst, float s, int n) {
  int i;
  if (scalar_is_one(s)) return;
  for (i = 0; i < n; ++i) dst[i] = dst[i] * s;
}
static void row_fill(float *dst, float v, int n) {
  int i;
  for (i = 0; i < n; ++i) dst[i] = v;
}
/* sse.h:289-317 sdot + sse.h:88-97 sum_all */
static float row_dot(const float *a, const float *b, int n) {
  const int len = (n >> 2) << 2;
  float l0 = 0.0f, l1 = 0.0f, l2 = 0.0f, l3 = 0.0f, sum;
  int i;
  for (i = 0; i < len; i += 4) {
    float p0 = a[i] * b[i], p1 = a[i + 1] * b[i + 1];
    float p2 = a[i + 2] * b[i + 2], p3 = a[i + 3] * b[i + 3];
    l0 = l0 + p0;
    l1 = l1 + p1;
    l2 = l2 + p2;
    l3 = l3 + p3;
  }
  {
    float a02 = l0 + l2, a13 = l1 + l3;
    sum = a02 + a13;
  }
  for (i = len; i < n; ++i) {
    float p = a[i] * b[i];
    sum = sum + p;
  }
  return sum;
}
/* apex_tensor_cpu_inline_common.h:168-175 */
static void row_reg_L1(float *w, float eps, int n) {
  int i;
  for (i = 0; i < n; ++i) {
    if (w[i] > eps) w[i] -= eps;
    else if (w[i] < -eps) w[i] += eps;
    else w[i] = 0.0f;
  }
}
/* base.h:175-180 */
static void scalar_reg_L1(float *w, float wd) {
  if (*w > wd) *w -= wd;
  else if (*w < -wd) *w += wd;
  else *w = 0.0f;
}
/* base.h:181-186 */
static void row_project(float *w, float B, int n) {
  float sum = row_dot(w, w, n);
  if (sum > B) row_scale(w, sqrtf(B / sum), n);
}

/* ---------------- active_type (model.h:61-238) ---------------- */

static float smooth_hinge_grad(float z) { /* model.h:90-94 */
  if (z > 1.0f) return 0.0f;
  if (z < 0.0f) return 1.0f;
  return 1.0f - z;
}
static float map_active(float sum, int type) { /* model.h:112-123 */
  switch (type) {
    case 0: return sum;
    case 1:
    case 2: return 1.0f / (1.0f + expf(-sum));
    case 3:
    case 5:
    case 6:
    case 7: return sum;
    default: die("unkown active type"); return 0.0f;
  }
}
static float cal_grad(float r, float pred, int type) { /* model.h:132-156 */
  switch (type) {
    case 0: return r - pred;
    case 1: return (r - pred) * pred * (1 - pred);
    case 2: return r - pred;
    case 7:
    case 3: return r - 1.0f / (1.0f + expf(-pred));
    case 5:
      if (r > 0.5f) return smooth_hinge_grad(pred - 0.5f);
      else return -smooth_hinge_grad(0.5f - pred);
    case 6:
      if (r > 0.5f) {
        if (pred > 1.0f) return 0.0f;
        else return r - pred;
      } else {
        if (pred < 0.0f) return 0.0f;
        else return r - pred;
      }
    default: die("unkown active type"); return 0.0f;
  }
}
static float calc_base_score(float base_score, int type) { /* model.h:220-237 */
  switch (type) {
    case 0:
    case 6:
    case 5: return base_score;
    case 1:
    case 2:
    case 3:
    case 7:
      if (!(base_score > 0.0f && base_score < 1.0f)) die("sigmoid range constrain");
      return -logf(1.0f / base_score - 1.0f);
    default: die("unkown active type"); return 0.0f;
  }
}

/* ---------------- RNG (apex-tensor/apex_random.h:40-77) ---------------- */

void svdo_seed(unsigned seed) { srand(seed); }
static double next_double2(void) { return ((double)rand() + 1.0) / ((double)RAND_MAX + 2.0); }
static double sample_normal(void) {
  double x, y, s;
  do {
    x = 2 * next_double2() - 1.0;
    y = 2 * next_double2() - 1.0;
    s = x * x + y * y;
  } while (s >= 1.0 || s == 0.0);
  return x * sqrt(-2.0 * log(s) / s);
}
/* apex_tensor_cpu_inline_common.h:249-253 */
static void sample_gaussian_rows(float *base, int pitch, int y_max, int x_max, float sd) {
  int y, x;
  for (y = 0; y < y_max; ++y)
    for (x = 0; x < x_max; ++x) base[(size_t)y * pitch + x] = (float)sample_normal() * sd;
}

/* ---------------- parameters ---------------- */

static void range_wd_init(range_wd_t *p, const char *a, const char *b) {
  memset(p, 0, sizeof(*p));
  strcpy(p->prefix_a, a);
  strcpy(p->prefix_b, b);
}
static void range_wd_set(range_wd_t *p, const char *name, const char *val) { /* base.h:48-68 */
  size_t la = strlen(p->prefix_a), lb = strlen(p->prefix_b);
  if (!strncmp(name, p->prefix_a, la)) name += la;
  else if (!strncmp(name, p->prefix_b, lb)) name += lb;
  else return;
  if (!strcmp("bound", name)) {
    unsigned bd = (unsigned)atoi(val);
    if (!(bd > 0)) die("can't give 0 as bound");
    if (!(p->n_bound == 0 || p->bound[p->n_bound - 1] < bd)) die("bound must be given in order");
    if (!(p->n_bound + 1 == p->n_wd)) die("must specifiy wd in each range");
    p->bound[p->n_bound++] = bd - 1;
  }
  if (!strcmp("wd", name)) {
    if (!(p->n_wd == p->n_bound)) die("setting must be exactly");
    p->wd[p->n_wd++] = (float)atof(val);
  }
}
static float range_wd_get(const range_wd_t *p, unsigned id, float wd_default) { /* base.h:69-74 */
  int i;
  if (p->n_bound == 0) return wd_default;
  for (i = 0; i < p->n_bound; ++i)
    if (!(p->bound[i] < id)) return p->wd[i]; /* std::lower_bound */
  die("bound set err");
  return 0.0f;
}

svdo_t *svdo_create(int format_type, int active_type, int extend_type) {
  svdo_t *m = (svdo_t *)calloc(1, sizeof(svdo_t));
  m->format_type = format_type;
  m->active_type = active_type;
  m->extend_type = extend_type;
  /* model.h:279-286 decide_format */
  if (m->format_type == FMT_AUTO) m->format_type = (extend_type == 0 ? FMT_RANDOM_ORDER : FMT_USER_GROUP);
  m->is_svdpp = (extend_type == 1 || m->format_type == FMT_USER_GROUP);
  /* model.h:436-450 */
  m->u_init_sigma = m->i_init_sigma = 0.01f;
  m->base_score = 0.5f;
  /* model.h:334-344 */
  m->learning_rate = 0.01f;
  m->decay_rate = 1.0f;
  m->scale_lr_ufeedback = 1.0f;
  range_wd_init(&m->u_param, "up:", "uip:");
  range_wd_init(&m->i_param, "ip:", "uip:");
  range_wd_init(&m->g_param, "gp:", "gp:");
  strcpy(m->name_feat_user, "NULL"); /* base.h:105-106 */
  strcpy(m->name_feat_item, "NULL");
  return m;
}

static void free_model(svdo_t *m) {
  if (!m->space_allocated) return;
  free(m->ui_bias);
  free(m->W_uiset);
  free(m->g_bias);
  m->space_allocated = 0;
}
void svdo_destroy(svdo_t *m) {
  if (!m) return;
  free_model(m);
  free(m->tmp_ufactor);
  free(m->tmp_ifactor);
  free(m->tmp_ufeedback);
  free(m->old_ufeedback);
  free(m->ref_user);
  free(m->ref_item);
  free(m->ref_global);
  free(m->feat_user.row_ptr); free(m->feat_user.index); free(m->feat_user.value);
  free(m->feat_item.row_ptr); free(m->feat_item.index); free(m->feat_item.value);
  free(m->rk.tmp_ifactors); free(m->rk.bias_ifactors); free(m->rk.tmp_ufactor); free(m->rk.tmp_ifactor);
  free(m->rk.tmp_ufeedback); free(m->rk.item_score); free(m->rk.item_tag); free(m->rk.pos_item);
  free(m);
}

void svdo_set_param(svdo_t *m, const char *name, const char *val) { /* base.h:126-136 */
  if (!strcmp(name, "feature_user")) strncpy(m->name_feat_user, val, 255); /* base.h:127-128 */
  if (!strcmp(name, "feature_item")) strncpy(m->name_feat_item, val, 255);
  /* model.h:350-368 SVDTrainParam::set_param */
  if (!strcmp("learning_rate", name)) m->learning_rate = (float)atof(val);
  if (!strcmp("wd_user", name)) m->wd_user = (float)atof(val);
  if (!strcmp("wd_item", name)) m->wd_item = (float)atof(val);
  if (!strcmp("wd_uiset", name)) m->wd_user = m->wd_item = (float)atof(val);
  if (!strcmp("wd_user_bias", name)) m->wd_user_bias = (float)atof(val);
  if (!strcmp("wd_item_bias", name)) m->wd_item_bias = (float)atof(val);
  if (!strcmp("wd_uiset_bias", name)) m->wd_user_bias = m->wd_item_bias = (float)atof(val);
  if (!strcmp("wd_global", name)) m->wd_global = (float)atof(val);
  if (!strcmp("reg_method", name)) m->reg_method = atoi(val);
  if (!strcmp("reg_global", name)) m->reg_global = atoi(val);
  if (!strcmp("num_regfree_global", name)) m->num_regfree_global = (unsigned)atoi(val);
  if (!strcmp("decay_learning_rate", name)) m->decay_learning_rate = atoi(val);
  if (!strcmp("min_learning_rate", name)) m->min_learning_rate = (float)atof(val);
  if (!strcmp("decay_rate", name)) m->decay_rate = (float)atof(val);
  if (!strcmp("scale_lr_ufeedback", name)) m->scale_lr_ufeedback = (float)atof(val);
  if (!strcmp("wd_ufeedback", name)) m->wd_ufeedback = (float)atof(val);
  if (!strcmp("wd_ufeedback_bias", name)) m->wd_ufeedback_bias = (float)atof(val);
  if (!strcmp(name, "top_k")) m->rk.top_k = atoi(val); /* SVDFeatureRanker::set_param, base.h:656-660 */
  range_wd_set(&m->u_param, name, val);
  range_wd_set(&m->i_param, name, val);
  range_wd_set(&m->g_param, name, val);
  if (m->space_allocated) return; /* base.h:133-135 */
  /* model.h:456-476 SVDModelParam::set_param */
  if (!strcmp("num_user", name)) m->num_user = atoi(val);
  if (!strcmp("num_item", name)) m->num_item = atoi(val);
  if (!strcmp("num_uiset", name)) m->num_user = m->num_item = atoi(val);
  if (!strcmp("num_global", name)) m->num_global = atoi(val);
  if (!strcmp("num_factor", name)) m->num_factor = atoi(val);
  if (!strcmp("u_init_sigma", name)) m->u_init_sigma = (float)atof(val);
  if (!strcmp("i_init_sigma", name)) m->i_init_sigma = (float)atof(val);
  if (!strcmp("ui_init_sigma", name)) m->u_init_sigma = m->i_init_sigma = (float)atof(val);
  if (!strcmp("base_score", name)) m->base_score = (float)atof(val);
  if (!strcmp("no_user_bias", name)) m->no_user_bias = atoi(val);
  if (!strcmp("num_ufeedback", name)) m->num_ufeedback = atoi(val);
  if (!strcmp("num_randinit_ufactor", name)) m->num_randinit_ufactor = atoi(val);
  if (!strcmp("num_randinit_ifactor", name)) m->num_randinit_ifactor = atoi(val);
  if (!strcmp("num_randinit_uifactor", name)) m->num_randinit_ifactor = m->num_randinit_ufactor = atoi(val);
  if (!strcmp("ufeedback_init_sigma", name)) m->ufeedback_init_sigma = (float)atof(val);
  if (!strcmp("common_latent_space", name)) m->common_latent_space = atoi(val);
  if (!strcmp("common_feedback_space", name)) m->common_feedback_space = atoi(val);
  if (!strcmp("user_nonnegative", name)) m->user_nonnegative = atoi(val);
  if (!strcmp("item_nonnegative", name)) m->item_nonnegative = atoi(val);
}

/* model.h:511-556 alloc_space (common_latent_space / common_feedback_space = 0 only) */
static void alloc_model(svdo_t *m) {
  if (m->common_latent_space != 0 || m->common_feedback_space != 0)
    die("oracle: common_latent_space/common_feedback_space are out of scope");
  m->ustart = (m->format_type == FMT_USER_GROUP) ? m->num_ufeedback : 0;
  m->rows = (long)m->ustart + m->num_user + m->num_item;
  m->pitch = ((m->num_factor + 3) >> 2) << 2;
  m->ui_bias = (float *)calloc((size_t)(m->rows > 0 ? m->rows : 1), sizeof(float));
  m->W_uiset = (float *)calloc((size_t)(m->rows > 0 ? m->rows : 1) * (m->pitch > 0 ? m->pitch : 1), sizeof(float));
  m->g_bias = (float *)calloc((size_t)(m->num_global > 0 ? m->num_global : 1), sizeof(float));
  m->u_bias = m->ui_bias + m->ustart;
  m->W_user = m->W_uiset + (size_t)m->ustart * m->pitch;
  m->i_bias = m->ui_bias + m->ustart + m->num_user;
  m->W_item = m->W_uiset + (size_t)(m->ustart + m->num_user) * m->pitch;
  m->ufeedback_bias = m->ui_bias;
  m->W_ufeedback = m->W_uiset;
  m->space_allocated = 1;
}

void svdo_init_model(svdo_t *m) { /* base.h:146-149; model.h:665-705 rand_init */
  int y, x;
  alloc_model(m);
  m->base_score = calc_base_score(m->base_score, m->active_type);
  {
    int ny = m->num_randinit_ufactor != 0 ? m->num_randinit_ufactor : m->num_user;
    sample_gaussian_rows(m->W_user, m->pitch, ny, m->num_factor, m->u_init_sigma);
    if (m->user_nonnegative)
      for (y = 0; y < m->num_user; ++y)
        for (x = 0; x < m->num_factor; ++x)
          m->W_user[(size_t)y * m->pitch + x] = fabsf(m->W_user[(size_t)y * m->pitch + x]);
  }
  {
    int ny = m->num_randinit_ifactor != 0 ? m->num_randinit_ifactor : m->num_item;
    sample_gaussian_rows(m->W_item, m->pitch, ny, m->num_factor, m->i_init_sigma);
    if (m->item_nonnegative)
      for (y = 0; y < ny; ++y)
        for (x = 0; x < m->num_factor; ++x)
          m->W_item[(size_t)y * m->pitch + x] = fabsf(m->W_item[(size_t)y * m->pitch + x]);
  }
  if (m->format_type == FMT_USER_GROUP)
    sample_gaussian_rows(m->W_ufeedback, m->pitch, m->num_ufeedback, m->num_factor, m->ufeedback_init_sigma);
}

/* SparseFeatureArray<float>::load (apex-utils/apex_utils.h:176-195): per line "n idx:val x n" */
static void sparse_load(const char *fname, unsigned *num_row, unsigned **row_ptr, unsigned **index, float **value) {
  FILE *fi = fopen(fname, "r");
  size_t cap_r = 16, cap_e = 64, ne = 0;
  int n;
  if (!fi) die("can not open file");
  *num_row = 0;
  *row_ptr = (unsigned *)malloc(cap_r * sizeof(unsigned));
  *index = (unsigned *)malloc(cap_e * sizeof(unsigned));
  *value = (float *)malloc(cap_e * sizeof(float));
  (*row_ptr)[0] = 0;
  while (fscanf(fi, "%d", &n) == 1) {
    int i;
    if (*num_row + 2 > cap_r) *row_ptr = (unsigned *)realloc(*row_ptr, (cap_r *= 2) * sizeof(unsigned));
    (*row_ptr)[*num_row + 1] = (*row_ptr)[*num_row] + (unsigned)n;
    (*num_row)++;
    for (i = 0; i < n; ++i) {
      if (ne + 1 > cap_e) {
        cap_e *= 2;
        *index = (unsigned *)realloc(*index, cap_e * sizeof(unsigned));
        *value = (float *)realloc(*value, cap_e * sizeof(float));
      }
      if (fscanf(fi, "%u:%f", &(*index)[ne], &(*value)[ne]) != 2) die("load sparse feature");
      ne++;
    }
  }
  fclose(fi);
}
/* SparseFeatureArray::operator[] (apex_utils.h:163-172): rows beyond the file are empty */
#define SIDE_BEGIN(f, id) ((id) < (f).num_row ? (f).row_ptr[(id)] : 0u)
#define SIDE_END(f, id) ((id) < (f).num_row ? (f).row_ptr[(id) + 1] : 0u)

void svdo_init_trainer(svdo_t *m) { /* base.h:151-173, 499-503 */
  size_t n = (size_t)(m->pitch > 0 ? m->pitch : 1);
  if (strcmp(m->name_feat_user, "NULL"))
    sparse_load(m->name_feat_user, &m->feat_user.num_row, &m->feat_user.row_ptr, &m->feat_user.index, &m->feat_user.value);
  if (strcmp(m->name_feat_item, "NULL"))
    sparse_load(m->name_feat_item, &m->feat_item.num_row, &m->feat_item.row_ptr, &m->feat_item.index, &m->feat_item.value);
  m->tmp_ufactor = (float *)calloc(n, sizeof(float));
  m->tmp_ifactor = (float *)calloc(n, sizeof(float));
  m->tmp_ufeedback = (float *)calloc(n, sizeof(float));
  m->old_ufeedback = (float *)calloc(n, sizeof(float));
  m->sample_counter = 0;
  if (m->reg_global >= 4) m->ref_global = (unsigned *)calloc((size_t)m->num_global + 1, sizeof(unsigned));
  if (m->reg_method >= 4) {
    m->ref_user = (unsigned *)calloc((size_t)m->num_user + 1, sizeof(unsigned));
    m->ref_item = (unsigned *)calloc((size_t)m->num_item + 1, sizeof(unsigned));
  }
  m->init_end = 1;
}

void svdo_set_round(svdo_t *m, int nround) { /* base.h:470-478 */
  if (m->decay_learning_rate != 0) {
    if (!(m->round_counter <= nround)) die("round counter restriction");
    while (m->round_counter < nround) {
      m->learning_rate *= m->decay_rate;
      m->round_counter++;
    }
  }
}

/* ---------------- model file (model.h:570-660, cpu_inline_common.h:72-88) ---------------- */

static void write_param(const svdo_t *m, FILE *fo) { /* model.h:373-450: 17 fields + reserved[247] */
  int32_t buf[264];
  memset(buf, 0, sizeof(buf));
  buf[0] = m->num_user;
  buf[1] = m->num_item;
  buf[2] = m->num_factor;
  buf[3] = m->num_global;
  memcpy(&buf[4], &m->u_init_sigma, 4);
  memcpy(&buf[5], &m->i_init_sigma, 4);
  memcpy(&buf[6], &m->base_score, 4);
  buf[7] = m->no_user_bias;
  buf[8] = m->num_ufeedback;
  memcpy(&buf[9], &m->ufeedback_init_sigma, 4);
  buf[10] = m->num_randinit_ufactor;
  buf[11] = m->num_randinit_ifactor;
  buf[12] = m->common_latent_space;
  buf[13] = m->user_nonnegative;
  buf[14] = m->common_feedback_space;
  buf[15] = m->extend_flag;
  buf[16] = m->item_nonnegative;
  fwrite(buf, sizeof(buf), 1, fo);
}
static int read_param(svdo_t *m, FILE *fi) {
  int32_t buf[264];
  if (fread(buf, sizeof(buf), 1, fi) != 1) return -1;
  m->num_user = buf[0];
  m->num_item = buf[1];
  m->num_factor = buf[2];
  m->num_global = buf[3];
  memcpy(&m->u_init_sigma, &buf[4], 4);
  memcpy(&m->i_init_sigma, &buf[5], 4);
  memcpy(&m->base_score, &buf[6], 4);
  m->no_user_bias = buf[7];
  m->num_ufeedback = buf[8];
  memcpy(&m->ufeedback_init_sigma, &buf[9], 4);
  m->num_randinit_ufactor = buf[10];
  m->num_randinit_ifactor = buf[11];
  m->common_latent_space = buf[12];
  m->user_nonnegative = buf[13];
  m->common_feedback_space = buf[14];
  m->extend_flag = buf[15];
  m->item_nonnegative = buf[16];
  return 0;
}
static void write_1d(const float *p, int x_max, FILE *fo) {
  int32_t h = x_max;
  fwrite(&h, 4, 1, fo);
  if (x_max > 0) fwrite(p, 4, (size_t)x_max, fo);
}
static void write_2d(const float *p, int pitch, int y_max, int x_max, FILE *fo) {
  int32_t h[2];
  int y;
  h[0] = x_max;
  h[1] = y_max;
  fwrite(h, 4, 2, fo);
  for (y = 0; y < y_max; ++y) fwrite(p + (size_t)y * pitch, 4, (size_t)x_max, fo);
}
static int read_1d(float *p, int expect, FILE *fi) {
  int32_t h;
  if (fread(&h, 4, 1, fi) != 1 || h != expect) return -1;
  if (h > 0 && fread(p, 4, (size_t)h, fi) != (size_t)h) return -1;
  return 0;
}
static int read_2d(float *p, int pitch, int y_expect, int x_expect, FILE *fi) {
  int32_t h[2];
  int y;
  if (fread(h, 4, 2, fi) != 2 || h[0] != x_expect || h[1] != y_expect) return -1;
  for (y = 0; y < y_expect; ++y)
    if (x_expect > 0 && fread(p + (size_t)y * pitch, 4, (size_t)x_expect, fi) != (size_t)x_expect) return -1;
  return 0;
}

int svdo_save_model(svdo_t *m, const char *path) { /* svd_feature.cpp:184-191 + model.h:638-660 */
  FILE *fo = fopen(path, "wb");
  uint8_t t[4];
  if (!fo) return -1;
  t[0] = (uint8_t)m->format_type;
  t[1] = (uint8_t)m->active_type;
  t[2] = (uint8_t)m->extend_type;
  t[3] = 0;
  fwrite(t, 4, 1, fo);
  write_param(m, fo);
  write_1d(m->u_bias, m->num_user, fo);
  write_2d(m->W_user, m->pitch, m->num_user, m->num_factor, fo);
  write_1d(m->i_bias, m->num_item, fo);
  write_2d(m->W_item, m->pitch, m->num_item, m->num_factor, fo);
  write_1d(m->g_bias, m->num_global, fo);
  if (m->format_type == FMT_USER_GROUP) {
    write_1d(m->ufeedback_bias, m->num_ufeedback, fo);
    write_2d(m->W_ufeedback, m->pitch, m->num_ufeedback, m->num_factor, fo);
  }
  fclose(fo);
  return 0;
}

int svdo_load_model(svdo_t *m, const char *path) { /* model.h:570-633 */
  FILE *fi = fopen(path, "rb");
  uint8_t t[4];
  int rc = 0;
  if (!fi) return -1;
  if (fread(t, 4, 1, fi) != 1) { fclose(fi); return -2; }
  if (read_param(m, fi)) { fclose(fi); return -3; }
  free_model(m);
  alloc_model(m);
  rc |= read_1d(m->u_bias, m->num_user, fi);
  rc |= read_2d(m->W_user, m->pitch, m->num_user, m->num_factor, fi);
  rc |= read_1d(m->i_bias, m->num_item, fi);
  rc |= read_2d(m->W_item, m->pitch, m->num_item, m->num_factor, fi);
  rc |= read_1d(m->g_bias, m->num_global, fi);
  if (m->format_type == FMT_USER_GROUP) {
    rc |= read_1d(m->ufeedback_bias, m->num_ufeedback, fi);
    rc |= read_2d(m->W_ufeedback, m->pitch, m->num_ufeedback, m->num_factor, fi);
  }
  fclose(fi);
  return rc ? -4 : 0;
}

/* ---------------- the hot path ---------------- */

typedef struct {
  float label;
  int ng, nu, ni;
  const unsigned *gi, *ui, *ii;
  const float *gv, *uv, *iv;
} elem_t;

static elem_t csr_row(int r, const int *row_ptr, const float *label, const unsigned *index,
                      const float *value) { /* apex_svd_data.h:129-142 */
  elem_t e;
  const int *p = row_ptr + 3 * r;
  e.label = label[r];
  e.ng = p[1] - p[0];
  e.nu = p[2] - p[1];
  e.ni = p[3] - p[2];
  e.gi = index + p[0]; e.gv = value + p[0];
  e.ui = index + p[1]; e.uv = value + p[1];
  e.ii = index + p[2]; e.iv = value + p[2];
  return e;
}

static void reg_global(svdo_t *m, unsigned gid) { /* base.h:188-210 */
  float lambda = m->learning_rate * range_wd_get(&m->g_param, gid, m->wd_global);
  if (gid >= m->num_regfree_global) {
    switch (m->reg_global) {
      case 0: m->g_bias[gid] *= (1.0f - lambda); break;
      case 1: scalar_reg_L1(&m->g_bias[gid], lambda); break;
      case 4: {
        float k = (float)(m->ref_global[gid] - m->sample_counter);
        m->g_bias[gid] *= expf(logf(1.0f - lambda) * k);
        m->ref_global[gid] = m->sample_counter;
        break;
      }
      case 5: {
        float k = (float)(m->ref_global[gid] - m->sample_counter);
        scalar_reg_L1(&m->g_bias[gid], lambda * k);
        m->ref_global[gid] = m->sample_counter;
        break;
      }
      default: die("unknown global decay method");
    }
  }
}
static void reg_user(svdo_t *m, unsigned uid) { /* base.h:211-250 */
  const int k = m->num_factor;
  float *w = m->W_user + (size_t)uid * m->pitch;
  float wd = range_wd_get(&m->u_param, uid, m->wd_user);
  float lambda = m->learning_rate * wd;
  switch (m->reg_method) {
    case 0: row_scale(w, 1.0f - lambda, k); break;
    case 3:
    case 1: row_reg_L1(w, lambda, k); break;
    case 2: row_project(w, wd, k); break;
    case 4: {
      float kk = (float)(m->ref_user[uid] - m->sample_counter);
      row_scale(w, expf(logf(1.0f - lambda) * kk), k);
      m->ref_user[uid] = m->sample_counter;
      break;
    }
    case 5: {
      float kk = (float)(m->ref_user[uid] - m->sample_counter);
      row_reg_L1(w, lambda * kk, k);
      m->ref_user[uid] = m->sample_counter;
      break;
    }
    default: die("unknown reg_method");
  }
  if (m->user_nonnegative) {
    int i;
    for (i = 0; i < k; ++i)
      if (w[i] <= 0.0f) w[i] = 0.0f;
  }
  if (m->no_user_bias == 0) m->u_bias[uid] *= (1.0f - m->learning_rate * m->wd_user_bias);
}
static void reg_item(svdo_t *m, unsigned iid) { /* base.h:251-283 */
  const int k = m->num_factor;
  float *w = m->W_item + (size_t)iid * m->pitch;
  float wd = range_wd_get(&m->i_param, iid, m->wd_item);
  float lambda = m->learning_rate * wd;
  switch (m->reg_method) {
    case 3:
    case 0: row_scale(w, 1.0f - lambda, k); break;
    case 1: row_reg_L1(w, lambda, k); break;
    case 2: row_project(w, wd, k); break;
    case 4: {
      float kk = (float)(m->ref_item[iid] - m->sample_counter);
      row_scale(w, expf(logf(1.0f - lambda) * kk), k);
      m->ref_item[iid] = m->sample_counter;
      break;
    }
    case 5: {
      float kk = (float)(m->ref_item[iid] - m->sample_counter);
      row_reg_L1(w, lambda * kk, k);
      m->ref_item[iid] = m->sample_counter;
      break;
    }
    default: die("unknown reg_method");
  }
  m->i_bias[iid] *= (1.0f - m->learning_rate * m->wd_item_bias);
}
static void regularize(svdo_t *m, const elem_t *e, int is_after) { /* base.h:286-311 */
  int i;
  if ((is_after && m->reg_global < 4) || (!is_after && m->reg_global >= 4))
    for (i = 0; i < e->ng; ++i) reg_global(m, e->gi[i]);
  if ((is_after && m->reg_method < 4) || (!is_after && m->reg_method >= 4)) {
    unsigned j;
    for (i = 0; i < e->nu; ++i) { /* base.h:298-303 */
      reg_user(m, e->ui[i]);
      for (j = SIDE_BEGIN(m->feat_user, e->ui[i]); j < SIDE_END(m->feat_user, e->ui[i]); ++j) reg_user(m, m->feat_user.index[j]);
    }
    for (i = 0; i < e->ni; ++i) { /* base.h:304-309 */
      reg_item(m, e->ii[i]);
      for (j = SIDE_BEGIN(m->feat_item, e->ii[i]); j < SIDE_END(m->feat_item, e->ii[i]); ++j) reg_item(m, m->feat_item.index[j]);
    }
  }
}

static double calc_bias(svdo_t *m, const elem_t *e) { /* base.h:313-353 */
  double sum = 0.0f;
  int i;
  for (i = 0; i < e->ng; ++i) {
    unsigned gid = e->gi[i];
    float p;
    if (!(gid < (unsigned)m->num_global)) die("global feature index exceed setting");
    p = e->gv[i] * m->g_bias[gid];
    sum += p;
  }
  if (m->no_user_bias == 0) {
    for (i = 0; i < e->nu; ++i) {
      unsigned uid = e->ui[i];
      float p;
      if (!(uid < (unsigned)m->num_user)) die("user feature index exceed bound");
      p = e->uv[i] * m->u_bias[uid];
      sum += p;
      { /* extra feature, base.h:329-333 */
        unsigned j;
        for (j = SIDE_BEGIN(m->feat_user, uid); j < SIDE_END(m->feat_user, uid); ++j) {
          float q = m->u_bias[m->feat_user.index[j]] * m->feat_user.value[j];
          sum += q;
        }
      }
    }
    sum += (m->is_svdpp ? m->tmp_ufeedback_bias : 0.0f); /* base.h:433-435, 509-511 */
  }
  sum += 0.0f; /* get_bias_plugin, base.h:436-438 */
  for (i = 0; i < e->ni; ++i) {
    unsigned iid = e->ii[i];
    float p;
    if (!(iid < (unsigned)m->num_item)) die("item feature index exceed bound");
    p = e->iv[i] * m->i_bias[iid];
    sum += p;
    { /* extra feature, base.h:345-349: (i_bias * value) * ival, left to right in float */
      unsigned j;
      for (j = SIDE_BEGIN(m->feat_item, iid); j < SIDE_END(m->feat_item, iid); ++j) {
        float q = m->i_bias[m->feat_item.index[j]] * m->feat_item.value[j];
        q = q * e->iv[i];
        sum += q;
      }
    }
  }
  return sum;
}

static void prepare_tmp(svdo_t *m, const elem_t *e) { /* base.h:354-381 */
  const int k = m->num_factor;
  int i;
  if (m->is_svdpp) memcpy(m->tmp_ufactor, m->tmp_ufeedback, sizeof(float) * (size_t)k); /* base.h:506-508 */
  else row_fill(m->tmp_ufactor, 0.0f, k);                                            /* base.h:430-432 */
  row_fill(m->tmp_ifactor, 0.0f, k);
  for (i = 0; i < e->nu; ++i) {
    unsigned uid = e->ui[i];
    if (!(uid < (unsigned)m->num_user)) die("user feature index exceed bound");
    row_add_scaled(m->tmp_ufactor, m->W_user + (size_t)uid * m->pitch, e->uv[i], k);
    { /* extra feature, base.h:364-368 */
      unsigned j;
      for (j = SIDE_BEGIN(m->feat_user, uid); j < SIDE_END(m->feat_user, uid); ++j)
        row_add_scaled(m->tmp_ufactor, m->W_user + (size_t)m->feat_user.index[j] * m->pitch, m->feat_user.value[j], k);
    }
  }
  for (i = 0; i < e->ni; ++i) {
    unsigned iid = e->ii[i];
    if (!(iid < (unsigned)m->num_item)) die("item feature index exceed bound");
    row_add_scaled(m->tmp_ifactor, m->W_item + (size_t)iid * m->pitch, e->iv[i], k);
    { /* extra feature, base.h:375-379: W * value * ival -- the two scalars are folded in double by
         operator*(ScalarMapExp, double) (apex_exp_template.h:500-503) and cast to float once */
      unsigned j;
      for (j = SIDE_BEGIN(m->feat_item, iid); j < SIDE_END(m->feat_item, iid); ++j) {
        float sc = (float)((double)m->feat_item.value[j] * (double)e->iv[i]);
        row_add_scaled(m->tmp_ifactor, m->W_item + (size_t)m->feat_item.index[j] * m->pitch, sc, k);
      }
    }
  }
}

static float pred(svdo_t *m, const elem_t *e) { /* base.h:445-454 */
  double sum = m->base_score + calc_bias(m, e);
  prepare_tmp(m, e);
  sum += row_dot(m->tmp_ufactor, m->tmp_ifactor, m->num_factor);
  return map_active((float)sum, m->active_type);
}

static void update_svdpp(svdo_t *m, float err) { /* base.h:512-520 */
  const int k = m->num_factor;
  float lr = m->learning_rate * m->scale_lr_ufeedback;
  row_add_scaled(m->tmp_ufeedback, m->tmp_ifactor, lr * err * m->norm_ufeedback, k);
  row_scale(m->tmp_ufeedback, 1.0f - lr * m->wd_ufeedback, k);
  if (m->no_user_bias == 0) {
    m->tmp_ufeedback_bias += lr * err * m->norm_ufeedback;
    m->tmp_ufeedback_bias *= (1.0f - lr * m->wd_ufeedback_bias);
  }
}

static void update_no_decay(svdo_t *m, float err, const elem_t *e) { /* base.h:383-427 */
  const int k = m->num_factor;
  int i;
  for (i = 0; i < e->ng; ++i) m->g_bias[e->gi[i]] += m->learning_rate * err * e->gv[i];
  for (i = 0; i < e->nu; ++i) {
    unsigned uid = e->ui[i];
    float scale = m->learning_rate * err * e->uv[i];
    row_add_scaled(m->W_user + (size_t)uid * m->pitch, m->tmp_ifactor, scale, k);
    if (m->no_user_bias == 0) m->u_bias[uid] += scale;
    { /* extra feature, base.h:399-407 */
      unsigned j;
      for (j = SIDE_BEGIN(m->feat_user, uid); j < SIDE_END(m->feat_user, uid); ++j) {
        float sc = m->learning_rate * err * m->feat_user.value[j];
        row_add_scaled(m->W_user + (size_t)m->feat_user.index[j] * m->pitch, m->tmp_ifactor, sc, k);
        if (m->no_user_bias == 0) m->u_bias[m->feat_user.index[j]] += sc;
      }
    }
  }
  for (i = 0; i < e->ni; ++i) {
    unsigned iid = e->ii[i];
    float scale = m->learning_rate * err * e->iv[i];
    row_add_scaled(m->W_item + (size_t)iid * m->pitch, m->tmp_ufactor, scale, k);
    m->i_bias[iid] += scale;
    { /* extra feature, base.h:417-422 */
      unsigned j;
      for (j = SIDE_BEGIN(m->feat_item, iid); j < SIDE_END(m->feat_item, iid); ++j) {
        float sc = m->learning_rate * err * m->feat_item.value[j] * e->iv[i];
        m->i_bias[m->feat_item.index[j]] += sc;
        row_add_scaled(m->W_item + (size_t)m->feat_item.index[j] * m->pitch, m->tmp_ufactor, sc, k);
      }
    }
  }
  if (m->is_svdpp) update_svdpp(m, err);
}

static void update_inner(svdo_t *m, const elem_t *e) { /* base.h:456-462 */
  float err;
  regularize(m, e, 0);
  err = cal_grad(e->label, pred(m, e), m->active_type) * 1.0f;
  update_no_decay(m, err, e);
  m->sample_counter++;
  regularize(m, e, 1);
}

void svdo_update_csr(svdo_t *m, int num_row, const int *row_ptr, const float *label,
                     const unsigned *index, const float *value) { /* base.h:464-466 */
  int r;
  for (r = 0; r < num_row; ++r) {
    elem_t e = csr_row(r, row_ptr, label, index, value);
    update_inner(m, &e);
  }
}
void svdo_predict_csr(svdo_t *m, int num_row, const int *row_ptr, const float *label,
                      const unsigned *index, const float *value, float *out) { /* base.h:467-469 */
  int r;
  for (r = 0; r < num_row; ++r) {
    elem_t e = csr_row(r, row_ptr, label, index, value);
    out[r] = pred(m, &e);
  }
}

static void prepare_ufeedback(svdo_t *m, int nfb, const unsigned *fi, const float *fv) { /* base.h:523-538 */
  const int k = m->num_factor;
  int i;
  m->norm_ufeedback = 0.0f;
  row_fill(m->tmp_ufeedback, 0.0f, k);
  m->tmp_ufeedback_bias = 0.0f;
  for (i = 0; i < nfb; ++i) {
    unsigned fid = fi[i];
    float val = fv[i];
    if (!(fid < (unsigned)m->num_ufeedback)) die("ufeedback id exceed bound");
    row_add_scaled(m->tmp_ufeedback, m->W_ufeedback + (size_t)fid * m->pitch, val, k);
    m->norm_ufeedback += val * val;
    if (m->no_user_bias == 0) m->tmp_ufeedback_bias += m->ufeedback_bias[fid] * val;
  }
}
static void update_ufeedback(svdo_t *m, int nfb, const unsigned *fi, const float *fv) { /* base.h:539-554 */
  const int k = m->num_factor;
  int i;
  if (nfb == 0) return;
  for (i = 0; i < k; ++i) m->tmp_ufeedback[i] = m->tmp_ufeedback[i] - m->old_ufeedback[i];
  m->tmp_ufeedback_bias -= m->old_ufeedback_bias;
  row_scale(m->tmp_ufeedback, 1.0f / m->norm_ufeedback, k);
  m->tmp_ufeedback_bias *= 1.0f / m->norm_ufeedback;
  for (i = 0; i < nfb; ++i) {
    unsigned fid = fi[i];
    float val = fv[i];
    row_add_scaled(m->W_ufeedback + (size_t)fid * m->pitch, m->tmp_ufeedback, val, k);
    if (m->no_user_bias == 0) m->ufeedback_bias[fid] += m->tmp_ufeedback_bias * val;
  }
}

void svdo_update_ugroup(svdo_t *m, int num_block, const int *blk_row_off, const int *blk_fb_off,
                        const int *blk_tag, const unsigned *fb_index, const float *fb_value,
                        const int *row_ptr, const float *label, const unsigned *index,
                        const float *value) { /* base.h:568-582 */
  int b, r;
  for (b = 0; b < num_block; ++b) {
    int tag = blk_tag ? blk_tag[b] : TAG_DEFAULT;
    int nfb = blk_fb_off[b + 1] - blk_fb_off[b];
    const unsigned *fi = fb_index + blk_fb_off[b];
    const float *fv = fb_value + blk_fb_off[b];
    if (tag == TAG_DEFAULT || tag == TAG_START) {
      prepare_ufeedback(m, nfb, fi, fv);
      m->old_ufeedback_bias = m->tmp_ufeedback_bias;
      memcpy(m->old_ufeedback, m->tmp_ufeedback, sizeof(float) * (size_t)m->num_factor);
    }
    for (r = blk_row_off[b]; r < blk_row_off[b + 1]; ++r) {
      elem_t e = csr_row(r, row_ptr, label, index, value);
      update_inner(m, &e);
    }
    if (tag == TAG_DEFAULT || tag == TAG_END) update_ufeedback(m, nfb, fi, fv);
  }
}
void svdo_predict_ugroup(svdo_t *m, int num_block, const int *blk_row_off, const int *blk_fb_off,
                         const int *blk_tag, const unsigned *fb_index, const float *fb_value,
                         const int *row_ptr, const float *label, const unsigned *index,
                         const float *value, float *out) { /* base.h:583-591 */
  int b, r;
  for (b = 0; b < num_block; ++b) {
    int tag = blk_tag ? blk_tag[b] : TAG_DEFAULT;
    if (tag == TAG_DEFAULT || tag == TAG_START)
      prepare_ufeedback(m, blk_fb_off[b + 1] - blk_fb_off[b], fb_index + blk_fb_off[b],
                        fb_value + blk_fb_off[b]);
    for (r = blk_row_off[b]; r < blk_row_off[b + 1]; ++r) {
      elem_t e = csr_row(r, row_ptr, label, index, value);
      out[r] = pred(m, &e);
    }
  }
}

float *svdo_data(svdo_t *m, int which) {
  switch (which) {
    case 0: return m->ui_bias;
    case 1: return m->W_uiset;
    case 2: return m->g_bias;
    default: return NULL;
  }
}
long svdo_info(svdo_t *m, int what) {
  switch (what) {
    case 0: return m->rows;
    case 1: return m->pitch;
    case 2: return m->ustart;
    case 3: return m->num_factor;
    case 4: return m->num_user;
    case 5: return m->num_item;
    case 6: return m->num_global;
    case 7: return m->num_ufeedback;
    default: return -1;
  }
}
float svdo_base_score(svdo_t *m) { return m->base_score; }
float svdo_learning_rate(svdo_t *m) { return m->learning_rate; }

/* ======================================================================================
 * SVDFeatureRanker (base.h:597-813): rank a fixed item set for a stream of user sections.
 * The input is a tagged instance stream (svdranker_tag, apex_svd.h:115-152): the label field
 * holds the tag, item indices of POS/BAN/SPEC rows sit in the user-feature field.
 * ====================================================================================== */
enum { RK_ITEM = 0, RK_POS = 1, RK_USER = 2, RK_SPEC = 3, RK_PROCESS = 4, RK_BAN = -1 };

void svdo_init_ranker(svdo_t *m, int num_item_set) { /* base.h:666-685 */
  size_t n = (size_t)(m->pitch > 0 ? m->pitch : 1), ns = (size_t)(num_item_set > 0 ? num_item_set : 1);
  if (strcmp(m->name_feat_user, "NULL"))
    sparse_load(m->name_feat_user, &m->feat_user.num_row, &m->feat_user.row_ptr, &m->feat_user.index, &m->feat_user.value);
  if (strcmp(m->name_feat_item, "NULL"))
    sparse_load(m->name_feat_item, &m->feat_item.num_row, &m->feat_item.row_ptr, &m->feat_item.index, &m->feat_item.value);
  m->rk.num_item_processed = 0;
  m->rk.num_item_set = num_item_set;
  m->rk.tmp_ufactor = (float *)calloc(n, sizeof(float));
  m->rk.tmp_ifactor = (float *)calloc(n, sizeof(float));
  m->rk.tmp_ufeedback = (float *)calloc(n, sizeof(float));
  m->rk.tmp_ifactors = (float *)calloc(n * ns, sizeof(float));
  m->rk.bias_ifactors = (float *)calloc(ns, sizeof(float));
  m->rk.item_score = (float *)calloc(ns, sizeof(float));
  m->rk.item_tag = (int *)calloc(ns, sizeof(int));
  m->rk.init_end = 1;
}

static void rk_prepare_ifactor(svdo_t *m, float *ifactor, float *bias_out, const elem_t *e) { /* base.h:687-710 */
  const int k = m->num_factor;
  float bias = 0.0f;
  int i;
  row_fill(ifactor, 0.0f, k);
  for (i = 0; i < e->ni; ++i) {
    const unsigned iid = e->ii[i];
    const float ival = e->iv[i];
    unsigned j;
    if (!(iid < (unsigned)m->num_item)) die("item feature index exceed setting");
    row_add_scaled(ifactor, m->W_item + (size_t)iid * m->pitch, ival, k);
    {
      float p = m->i_bias[iid] * ival;
      bias = bias + p;
    }
    for (j = SIDE_BEGIN(m->feat_item, iid); j < SIDE_END(m->feat_item, iid); ++j) {
      /* W * value * ival: the two scalars fold in double (apex_exp_template.h:500-503), see prepare_tmp */
      float sc = (float)((double)m->feat_item.value[j] * (double)ival);
      float p = m->i_bias[m->feat_item.index[j]] * m->feat_item.value[j];
      row_add_scaled(ifactor, m->W_item + (size_t)m->feat_item.index[j] * m->pitch, sc, k);
      p = p * ival;
      bias = bias + p;
    }
  }
  for (i = 0; i < e->ng; ++i) {
    const unsigned gid = e->gi[i];
    float p;
    if (!(gid < (unsigned)m->num_global)) die("global feature index exceed setting");
    p = e->gv[i] * m->g_bias[gid];
    bias = bias + p;
  }
  *bias_out = bias;
}

static void rk_proc_item(svdo_t *m, const elem_t *e) { /* base.h:712-717 */
  const int idx = m->rk.num_item_processed++;
  if (!(m->rk.num_item_processed <= m->rk.num_item_set)) die("item instance exceed specified item set size");
  rk_prepare_ifactor(m, m->rk.tmp_ifactors + (size_t)idx * m->pitch, &m->rk.bias_ifactors[idx], e);
}

static void rk_proc_user(svdo_t *m, const elem_t *e) { /* base.h:719-740 */
  const int k = m->num_factor;
  int i;
  if (m->format_type == FMT_USER_GROUP) memcpy(m->rk.tmp_ufactor, m->rk.tmp_ufeedback, sizeof(float) * (size_t)k);
  else row_fill(m->rk.tmp_ufactor, 0.0f, k);
  for (i = 0; i < e->nu; ++i) {
    const unsigned uid = e->ui[i];
    unsigned j;
    if (!(uid < (unsigned)m->num_user)) die("user feature index exceed bound");
    row_add_scaled(m->rk.tmp_ufactor, m->W_user + (size_t)uid * m->pitch, e->uv[i], k);
    for (j = SIDE_BEGIN(m->feat_user, uid); j < SIDE_END(m->feat_user, uid); ++j)
      row_add_scaled(m->rk.tmp_ufactor, m->W_user + (size_t)m->feat_user.index[j] * m->pitch, m->feat_user.value[j], k);
  }
  m->rk.num_pos = 0;
  for (i = 0; i < m->rk.num_item_set; ++i) m->rk.item_score[i] = 0.0f;
  for (i = 0; i < m->rk.num_item_processed; ++i) m->rk.item_tag[i] = 0;
}

static void rk_proc_tag(svdo_t *m, const elem_t *e, int tag) { /* base.h:741-749 */
  int i;
  for (i = 0; i < e->nu; ++i) {
    const int idx = (int)e->ui[i];
    if (!(idx < m->rk.num_item_processed)) die("sample item index exceed bound");
    if (!(m->rk.item_tag[idx] == 0)) die("each pos sample item can not occur in baned sample list");
    m->rk.item_tag[idx] = tag;
    if (tag == RK_POS) {
      if (m->rk.num_pos == m->rk.cap_pos) {
        m->rk.cap_pos = m->rk.cap_pos ? 2 * m->rk.cap_pos : 16;
        m->rk.pos_item = (int *)realloc(m->rk.pos_item, sizeof(int) * (size_t)m->rk.cap_pos);
      }
      m->rk.pos_item[m->rk.num_pos++] = idx;
    }
  }
}

static void rk_proc_spec(svdo_t *m, const elem_t *e) { /* base.h:750-758 */
  float bias, d;
  int idx;
  if (!(e->nu == 1)) die("must specify item index of sample in user feature field\n");
  idx = (int)e->ui[0];
  if (!(idx < m->rk.num_item_processed)) die("sample item index exceed bound");
  rk_prepare_ifactor(m, m->rk.tmp_ifactor, &bias, e);
  d = row_dot(m->rk.tmp_ufactor, m->rk.tmp_ifactor, m->num_factor);
  m->rk.item_score[idx] = bias + d;
}

typedef struct { int iid; float score; } rk_entry;
/* Entry::operator< (base.h:622): higher score first.  std::sort leaves the order of equal scores
 * to the implementation; this restatement puts the lower item index first (stable). */
static int rk_cmp(const void *a, const void *b) {
  const rk_entry *x = (const rk_entry *)a, *y = (const rk_entry *)b;
  if (x->score > y->score) return -1;
  if (y->score > x->score) return 1;
  return x->iid < y->iid ? -1 : (x->iid > y->iid ? 1 : 0);
}

static long rk_proc_rank(svdo_t *m, int *rst, long cap, long n_out) { /* base.h:759-782 */
  const int n = m->rk.num_item_processed;
  rk_entry *entry = (rk_entry *)malloc(sizeof(rk_entry) * (size_t)(n > 0 ? n : 1));
  int ne = 0, i;
  for (i = 0; i < n; ++i) {
    float t;
    if (m->rk.item_tag[i] == RK_BAN) continue;
    t = row_dot(m->rk.tmp_ufactor, m->rk.tmp_ifactors + (size_t)i * m->pitch, m->num_factor);
    t = m->rk.bias_ifactors[i] + t;
    m->rk.item_score[i] = m->rk.item_score[i] + t;
    entry[ne].iid = i;
    entry[ne].score = m->rk.item_score[i];
    ne++;
  }
  qsort(entry, (size_t)ne, sizeof(rk_entry), rk_cmp);
  if (m->rk.top_k > 0) {
    if (!(ne >= m->rk.top_k)) die("k can not exceed candidate size");
    for (i = 0; i < m->rk.top_k; ++i) {
      if (n_out < cap) rst[n_out] = entry[i].iid;
      n_out++;
    }
  } else {
    for (i = 0; i < ne; ++i) m->rk.item_tag[entry[i].iid] = i;
    for (i = 0; i < m->rk.num_pos; ++i) {
      if (n_out < cap) rst[n_out] = m->rk.item_tag[m->rk.pos_item[i]];
      n_out++;
    }
  }
  free(entry);
  return n_out;
}

static long rk_proc(svdo_t *m, const elem_t *e, int *rst, long cap, long n_out) { /* base.h:783-793 */
  const int tag = (int)e->label;
  switch (tag) {
    case RK_ITEM: rk_proc_item(m, e); break;
    case RK_USER: rk_proc_user(m, e); break;
    case RK_POS:
    case RK_BAN: rk_proc_tag(m, e, tag); break;
    case RK_SPEC: rk_proc_spec(m, e); break;
    case RK_PROCESS: n_out = rk_proc_rank(m, rst, cap, n_out); break;
    default: break;
  }
  return n_out;
}

/* for each row ISVDRanker::process(result, Elem) (base.h:795-797); returns the number of results
 * (those beyond cap are counted but not stored) */
long svdo_rank_csr(svdo_t *m, int num_row, const int *row_ptr, const float *label, const unsigned *index,
                   const float *value, int *result, long cap) {
  long n_out = 0;
  int r;
  for (r = 0; r < num_row; ++r) {
    elem_t e = csr_row(r, row_ptr, label, index, value);
    n_out = rk_proc(m, &e, result, cap, n_out);
  }
  return n_out;
}

/* for each block ISVDRanker::process(result, SVDPlusBlock) (base.h:798-812) */
long svdo_rank_ugroup(svdo_t *m, int num_block, const int *blk_row_off, const int *blk_fb_off, const int *blk_tag,
                      const unsigned *fb_index, const float *fb_value, const int *row_ptr, const float *label,
                      const unsigned *index, const float *value, int *result, long cap) {
  long n_out = 0;
  int b, r, i;
  for (b = 0; b < num_block; ++b) {
    const int tag = blk_tag ? blk_tag[b] : TAG_DEFAULT;
    if (tag == TAG_DEFAULT || tag == TAG_START) {
      row_fill(m->rk.tmp_ufeedback, 0.0f, m->num_factor);
      for (i = blk_fb_off[b]; i < blk_fb_off[b + 1]; ++i) {
        const unsigned fid = fb_index[i];
        if (!(fid < (unsigned)m->num_ufeedback)) die("ufeedback id exceed bound");
        row_add_scaled(m->rk.tmp_ufeedback, m->W_ufeedback + (size_t)fid * m->pitch, fb_value[i], m->num_factor);
      }
    }
    for (r = blk_row_off[b]; r < blk_row_off[b + 1]; ++r) {
      elem_t e = csr_row(r, row_ptr, label, index, value);
      n_out = rk_proc(m, &e, result, cap, n_out);
    }
  }
  return n_out;
}
