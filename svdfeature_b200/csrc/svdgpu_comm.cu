// svdgpu_comm.cu -- the multi-GPU exchange inside the library: one handle
// per GPU (one process per GPU, or several handles in one process), NCCL over NVLink.
//
//   * instances and user rows are partitioned by (user id mod world): a rank trains only the rows
//     of its users, so user rows are never communicated during training;
//   * the item side (W_item, i_bias, g_bias, and the item-indexed feedback rows) is replicated:
//     svdgpu_allreduce_items packs delta = current - snapshot into one contiguous buffer, sums it
//     over the ranks with ONE ncclAllReduce on the launch stream, and applies snapshot + scale*sum;
//   * svdgpu_allgather_users completes the model on every rank before it is saved or evaluated:
//     the rows of users a rank does not own are zeroed and the user slab is summed in place.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): inside a PyTorch process that is the library
// torch already loaded, in the C++ trainer the system's.  The reference has no distributed code at
// all; the caller these entry points serve is its training loop (svd_feature.cpp:231-247).
#include "svdgpu_internal.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

using namespace svdk;

namespace {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

NcclApi &nccl() {
  static NcclApi api;
  if (api.lib || !api.err.empty()) return api;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) {
    api.err = std::string("NCCL is not available: ") + dlerror();
    return api;
  }
#define SYM(field, name)                                              \
  *(void **)(&api.field) = dlsym(api.lib, name);                      \
  if (!api.field) api.err = std::string("NCCL symbol missing: ") + name;
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommInitAll, "ncclCommInitAll")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllReduce, "ncclAllReduce")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return api;
}

#define NC(h, call)                                                                          \
  do {                                                                                       \
    ncclResult_t r_ = (call);                                                                \
    if (r_ != ncclSuccess) return fail(h, "%s failed: %s", #call, nccl().GetErrorString(r_)); \
  } while (0)

// rows of users another rank owns (user id mod world != rank) are zeroed: W rows and biases
__global__ void k_zero_foreign_users(float *W, float *bias, int num_user, int pitch, int world, int rank) {
  const long long total = (long long)num_user * pitch;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int u = (int)(i / pitch);
    if (u % world != rank) {
      W[i] = 0.0f;
      if (i % pitch == 0) bias[u] = 0.0f;
    }
  }
}

}  // namespace

struct svdgpu_comm {
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0;
};

extern "C" {

int svdgpu_comm_id(char *id128) {
  if (!id128) return 1;
  NcclApi &n = nccl();
  if (!n.err.empty()) return fail(nullptr, "%s", n.err.c_str());
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "the id is handed around as 128 bytes");
  NC(nullptr, n.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int svdgpu_comm_init(svdgpu_t *h, int world, int rank, const char *id128) {
  if (!h) return 1;
  if (world < 1 || rank < 0 || rank >= world) return fail(h, "comm_init: rank %d of %d", rank, world);
  if (h->comm) return fail(h, "comm_init: the handle already has a communicator");
  CU(h, cudaSetDevice(h->device));
  svdgpu_comm *c = new svdgpu_comm();
  c->world = world;
  c->rank = rank;
  if (world > 1) {
    if (!id128) {
      delete c;
      return fail(h, "comm_init: null id");
    }
    NcclApi &n = nccl();
    if (!n.err.empty()) {
      delete c;
      return fail(h, "%s", n.err.c_str());
    }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclResult_t r = n.CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
      delete c;
      return fail(h, "ncclCommInitRank failed: %s", n.GetErrorString(r));
    }
  }
  h->comm = c;
  return 0;
}

int svdgpu_comm_destroy(svdgpu_t *h) {
  if (!h || !h->comm) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->comm->comm) nccl().CommDestroy(h->comm->comm);
  delete h->comm;
  h->comm = nullptr;
  return 0;
}

int svdgpu_comm_rank(const svdgpu_t *h, int *world, int *rank) {
  if (!h) return 1;
  if (world) *world = h->comm ? h->comm->world : 1;
  if (rank) *rank = h->comm ? h->comm->rank : 0;
  return 0;
}

int svdgpu_allreduce_items(svdgpu_t *h, float scale) {
  if (!h) return 1;
  if (!h->comm) return fail(h, "allreduce_items: call svdgpu_comm_init first");
  CU(h, cudaSetDevice(h->device));
  if (!h->d_snap) return svdgpu_items_snapshot(h);  // first call: nothing to exchange yet
  void *buf = nullptr;
  size_t nf = 0;
  if (svdgpu_items_pack_delta(h, &buf, &nf)) return 1;
  if (h->comm->world > 1 && nf > 0) {
    NC(h, nccl().AllReduce(buf, buf, nf, ncclFloat, ncclSum, h->comm->comm, h->stream));
    h->n_coll++;
    h->n_coll_bytes += (long long)nf * 4;
  }
  return svdgpu_items_apply_delta(h, scale);
}

int svdgpu_allgather_users(svdgpu_t *h) {
  if (!h) return 1;
  if (!h->comm) return fail(h, "allgather_users: call svdgpu_comm_init first");
  CU(h, cudaSetDevice(h->device));
  const DevModel &m = h->dm;
  if (h->comm->world <= 1 || m.num_user == 0) return 0;
  float *W = m.W + (size_t)m.user_off * m.pitch, *b = m.bias + m.user_off;
  k_zero_foreign_users<<<h->num_sm * 8, 256, 0, h->stream>>>(W, b, m.num_user, m.pitch, h->comm->world, h->comm->rank);
  CU(h, cudaGetLastError());
  h->n_launch++;
  NcclApi &n = nccl();
  NC(h, n.GroupStart());
  NC(h, n.AllReduce(W, W, (size_t)m.num_user * m.pitch, ncclFloat, ncclSum, h->comm->comm, h->stream));
  NC(h, n.AllReduce(b, b, (size_t)m.num_user, ncclFloat, ncclSum, h->comm->comm, h->stream));
  NC(h, n.GroupEnd());
  h->n_coll += 2;
  h->n_coll_bytes += ((long long)m.num_user * m.pitch + m.num_user) * 4;
  return 0;
}

// Several handles in ONE process (one per device): communicators are made with ncclCommInitAll on
// first use; the exchange of all handles is one NCCL group.
int svdgpu_allreduce_items_group(svdgpu_t **hs, int n, float scale) {
  if (!hs || n < 1) return 1;
  svdgpu *h0 = hs[0];
  bool have = true;
  for (int i = 0; i < n; ++i) have = have && hs[i] && hs[i]->comm;
  if (!have) {
    for (int i = 0; i < n; ++i)
      if (!hs[i] || hs[i]->comm) return fail(h0, "allreduce_items_group: either every handle has a communicator or none");
    NcclApi &api = nccl();
    if (n > 1 && !api.err.empty()) return fail(h0, "%s", api.err.c_str());
    std::vector<ncclComm_t> comms((size_t)n, nullptr);
    std::vector<int> devs((size_t)n);
    for (int i = 0; i < n; ++i) devs[(size_t)i] = hs[i]->device;
    if (n > 1) NC(h0, api.CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) {
      svdgpu_comm *c = new svdgpu_comm();
      c->comm = comms[(size_t)i];
      c->world = n;
      c->rank = i;
      hs[i]->comm = c;
    }
  }
  bool first = false;
  for (int i = 0; i < n; ++i) first = first || !hs[i]->d_snap;
  if (first) {
    for (int i = 0; i < n; ++i)
      if (svdgpu_items_snapshot(hs[i])) return 1;
    return 0;
  }
  std::vector<void *> buf((size_t)n, nullptr);
  std::vector<size_t> nf((size_t)n, 0);
  for (int i = 0; i < n; ++i)
    if (svdgpu_items_pack_delta(hs[i], &buf[(size_t)i], &nf[(size_t)i])) return 1;
  if (n > 1) {
    NcclApi &api = nccl();
    NC(h0, api.GroupStart());
    for (int i = 0; i < n; ++i) {
      cudaSetDevice(hs[i]->device);
      ncclResult_t r = api.AllReduce(buf[(size_t)i], buf[(size_t)i], nf[(size_t)i], ncclFloat, ncclSum, hs[i]->comm->comm,
                                     hs[i]->stream);
      if (r != ncclSuccess) {
        api.GroupEnd();
        return fail(h0, "ncclAllReduce failed: %s", api.GetErrorString(r));
      }
      hs[i]->n_coll++;
      hs[i]->n_coll_bytes += (long long)nf[(size_t)i] * 4;
    }
    NC(h0, api.GroupEnd());
  }
  for (int i = 0; i < n; ++i)
    if (svdgpu_items_apply_delta(hs[i], scale)) return 1;
  return 0;
}

}  // extern "C"
