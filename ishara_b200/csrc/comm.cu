// ishara_b200 — data-parallel gradient exchange inside the library (SURVEY.md §8b `ishara_model_comm_init`, §8e).
//
// The one real exchange step of the path is the gradient all-reduce of the training step (the reference's only
// counterpart is nn.DataParallel, integration.py:1058-1060). One process per GPU; every rank holds the same weights,
// runs forward/backward on its shard and sums gradients over NVLink/NVSwitch with NCCL. The exchange is bucketed per
// module and issued on a second stream as soon as a module's backward has finished, so it overlaps the rest of the
// backward pass; nothing is read back to the host before it (train.cu).
//
// NCCL is bound at RUN time (dlopen), not linked: the library loads on a host without NCCL / without a GPU (the C-ABI
// tests do that), and inside a process that already loaded a libnccl.so.2 (PyTorch's) the same instance is reused.
#include <dlfcn.h>
#include <nccl.h>  // types and enums only

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "model_internal.h"

namespace ishara {

void train_drop_graph(ishara_model* m);  // train.cu

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  std::string error;
  bool ok = false;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* env = getenv("ISHARA_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (n == nullptr || *n == 0) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle != nullptr) break;
    }
    if (api.handle == nullptr) {
      api.error = "NCCL not found (tried ISHARA_NCCL_LIB, libnccl.so.2, libnccl.so)";
      return;
    }
    auto sym = [&](const char* s) { return dlsym(api.handle, s); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
    if (!api.ok) api.error = "libnccl is missing one of ncclGetUniqueId/CommInitRank/CommDestroy/AllReduce/GetErrorString";
  });
  return &api;
}

#define ISHARA_NCCL_OK(expr)                                                                         \
  do {                                                                                               \
    ncclResult_t _r = (expr);                                                                        \
    if (_r != ncclSuccess) {                                                                         \
      set_last_error(std::string(#expr) + ": " + nccl_api()->GetErrorString(_r));                    \
      return ISHARA_ERR_COMM;                                                                        \
    }                                                                                                \
  } while (0)

}  // namespace

// ---- bucket plan (pure host logic; exported for the CPU tests as ishara_comm_bucket_plan) --------------------------
// Modules run their backward in REVERSE order. hi[k] = end offset (in floats) of the highest gradient module k writes.
// A gradient is final once every module that writes it has run, i.e. after the backward of module k everything at or
// above max_{j<k} hi[j] is final. Consecutive final ranges are merged until they hold at least min_elems values.
// Output: for each module k (forward order) the half-open range [lo[k], up[k]) to reduce right after ITS backward
// (empty when lo == up). The ranges tile [0, n_train) exactly once.
void comm_bucket_plan(const int64_t* hi, int n, int64_t n_train, int64_t min_elems, int64_t* lo_out, int64_t* up_out) {
  std::vector<int64_t> bound(static_cast<size_t>(n) + 1, 0);  // bound[k] = max_{j<k} hi[j]
  for (int k = 0; k < n; ++k) bound[k + 1] = std::max(bound[k], std::min(hi[k], n_train));
  int64_t pending_up = n_train;  // everything in [bound[k], pending_up) is final after module k
  for (int k = n - 1; k >= 0; --k) {
    const int64_t lo = bound[k];
    lo_out[k] = up_out[k] = 0;
    if (lo >= pending_up) continue;
    if (pending_up - lo >= min_elems || k == 0) {
      lo_out[k] = k == 0 ? 0 : lo;
      up_out[k] = pending_up;
      pending_up = lo_out[k];
    }
  }
}

int comm_unique_id(void* out128) {
  NcclApi* api = nccl_api();
  if (!api->ok) { set_last_error(api->error); return ISHARA_ERR_COMM; }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes in every NCCL 2.x");
  ncclUniqueId id;
  ISHARA_NCCL_OK(api->GetUniqueId(&id));
  std::memcpy(out128, &id, sizeof(id));
  return 0;
}

int model_comm_init(ishara_model* m, const void* id128, int rank, int world) {
  if (id128 == nullptr || world < 1 || rank < 0 || rank >= world) { set_last_error("comm_init: bad arguments"); return ISHARA_ERR_INVALID; }
  NcclApi* api = nccl_api();
  if (!api->ok) { set_last_error(api->error); return ISHARA_ERR_COMM; }
  if (m->comm != nullptr) { set_last_error("comm_init: communicator already initialised (call comm_destroy first)"); return ISHARA_ERR_STATE; }
  ISHARA_CUDA_OK(cudaSetDevice(m->device));
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  train_drop_graph(m);  // a step captured without the exchange is no longer the step to replay
  ncclComm_t comm = nullptr;
  ISHARA_NCCL_OK(api->CommInitRank(&comm, world, id, rank));
  m->comm = comm;
  m->comm_rank = rank;
  m->comm_world = world;
  if (m->comm_stream == nullptr) ISHARA_CUDA_OK(cudaStreamCreateWithFlags(&m->comm_stream, cudaStreamNonBlocking));
  if (m->comm_ready == nullptr) {
    ISHARA_CUDA_OK(cudaEventCreateWithFlags(&m->comm_ready, cudaEventDisableTiming));
    ISHARA_CUDA_OK(cudaEventCreateWithFlags(&m->comm_done, cudaEventDisableTiming));
  }
  return 0;
}

int model_comm_destroy(ishara_model* m) {
  if (m->comm != nullptr) {
    cudaSetDevice(m->device);
    if (m->comm_stream) cudaStreamSynchronize(m->comm_stream);
    train_drop_graph(m);  // graphs that captured NCCL collectives must go before the communicator does
    nccl_api()->CommDestroy(static_cast<ncclComm_t>(m->comm));
    m->comm = nullptr;
  }
  m->comm_world = 1;
  m->comm_rank = 0;
  return 0;
}

// In-place sum of buf[0, count) over all ranks, on the handle's communication stream, ordered after everything
// enqueued on `after` so far. The caller later makes its stream wait on comm_done (comm_join).
int comm_allreduce_after(ishara_model* m, float* buf, int64_t count, cudaStream_t after) {
  if (m->comm == nullptr || count <= 0) return 0;
  NcclApi* api = nccl_api();
  ISHARA_CUDA_OK(cudaEventRecord(m->comm_ready, after));
  ISHARA_CUDA_OK(cudaStreamWaitEvent(m->comm_stream, m->comm_ready, 0));
  ISHARA_NCCL_OK(api->AllReduce(buf, buf, static_cast<size_t>(count), ncclFloat, ncclSum, static_cast<ncclComm_t>(m->comm), m->comm_stream));
  return 0;
}

// `stream` waits until every exchange issued so far has finished
int comm_join(ishara_model* m, cudaStream_t stream) {
  if (m->comm == nullptr) return 0;
  ISHARA_CUDA_OK(cudaEventRecord(m->comm_done, m->comm_stream));
  ISHARA_CUDA_OK(cudaStreamWaitEvent(stream, m->comm_done, 0));
  return 0;
}

int nccl_version() {
  NcclApi* api = nccl_api();
  int v = 0;
  if (api->ok && api->GetVersion != nullptr) api->GetVersion(&v);
  return v;
}

}  // namespace ishara
