// oracle/ref_pairs_shim.cpp -- TEST INFRASTRUCTURE ONLY.  Drives the UNMODIFIED reference's
// PairwiseRankGenerator (apex_svd_data.cpp:812-1025) through its own factory
// (create_plus_iterator(BINARY_BUFFER_RANK), apex_svd_data.cpp:1313-1334) on a user-group buffer
// file and hands the generated blocks out through a C interface, so that
// tests/golden/make_pairs_golden.py can record what the reference's sampler does with given blocks.
// Compiled only by `make -C oracle ref` (needs /root/reference); never linked by the product.
#include "apex_svd_data.h"
#include "apex-tensor/apex_random.h"

#include <cstring>

using namespace apex_svd;

extern "C" {

void *refpairs_open(const char *buffer_path, int n_kv, const char **keys, const char **vals, unsigned seed) {
  IDataIterator<SVDPlusBlock> *it = create_plus_iterator(input_type::BINARY_BUFFER_RANK);
  it->set_param("buffer_feature", buffer_path);
  for (int i = 0; i < n_kv; ++i) it->set_param(keys[i], vals[i]);
  apex_random::seed(seed);
  it->init();
  it->before_first();
  return it;
}

// next generated block: returns 0 at the end of the file, 1 otherwise; -1 if the caps are too small
int refpairs_next(void *h, int cap_row, int cap_val, int *num_row, int *num_val, int *row_ptr, float *label,
                  unsigned *index, float *value) {
  IDataIterator<SVDPlusBlock> *it = static_cast<IDataIterator<SVDPlusBlock> *>(h);
  SVDPlusBlock b;
  if (!it->next(b)) return 0;
  *num_row = b.data.num_row;
  *num_val = b.data.num_val;
  if (b.data.num_row > cap_row || b.data.num_val > cap_val) return -1;
  memcpy(row_ptr, b.data.row_ptr, sizeof(int) * (3 * (size_t)b.data.num_row + 1));
  if (b.data.num_row) {
    memcpy(label, b.data.row_label, sizeof(float) * (size_t)b.data.num_row);
    memcpy(index, b.data.feat_index, sizeof(unsigned) * (size_t)b.data.num_val);
    memcpy(value, b.data.feat_value, sizeof(float) * (size_t)b.data.num_val);
  }
  return 1;
}

void refpairs_close(void *h) { delete static_cast<IDataIterator<SVDPlusBlock> *>(h); }

}  // extern "C"
