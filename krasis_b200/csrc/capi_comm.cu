// C-ABI for the expert-parallel communicator (include/krasis_b200.h, "multi-GPU" section).
//
// The reference's EP exchange is Python: every GPU gets all tokens through a pinned-host bounce and GPU0 adds the partial
// sums (python/krasis/model.py:3086-3211, KrasisEngine.reduce_sum_bf16 src/moe.rs:2505).  Here one process drives one GPU
// and the exchange is NCCL over NVLink, owned by this library so a Rust / C host needs no torch: the host only moves the
// 128-byte NCCL unique id between its processes (any side channel), everything else is behind these calls.
// NCCL is bound at run time (dlopen of libnccl.so.2 — the copy already loaded by the host process if there is one), so the
// library itself has no link-time dependency and single-GPU users never touch it.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/krasis_b200.h"

extern "C" int kb2_set_error_(int code, const char* msg);   // capi.cu

namespace {

typedef struct ncclComm* ncclComm_t;
struct NcclUniqueId { char internal[128]; };
enum { kNcclUint8 = 1, kNcclBfloat16 = 9 };
enum { kNcclSum = 0 };

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*ReduceScatter)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);     // the host's copy (e.g. torch's bundled NCCL) if loaded
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
#define KB2_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name))
    KB2_SYM(GetUniqueId, "ncclGetUniqueId");
    KB2_SYM(CommInitRank, "ncclCommInitRank");
    KB2_SYM(CommDestroy, "ncclCommDestroy");
    KB2_SYM(AllGather, "ncclAllGather");
    KB2_SYM(ReduceScatter, "ncclReduceScatter");
    KB2_SYM(AllReduce, "ncclAllReduce");
    KB2_SYM(Broadcast, "ncclBroadcast");
    KB2_SYM(Reduce, "ncclReduce");
    KB2_SYM(GetErrorString, "ncclGetErrorString");
#undef KB2_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.ReduceScatter && api.AllReduce &&
             api.Broadcast && api.Reduce && api.GetErrorString;
  });
  return api;
}

int failf(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  return kb2_set_error_(code, buf);
}

}  // namespace

struct kb2_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1, device = 0;
  int* barrier_word = nullptr;        // 4-byte device buffer of kb2_comm_barrier
};

#define NCCL_TRY(expr)                                                                                   \
  do {                                                                                                   \
    int _r = (expr);                                                                                     \
    if (_r != 0) return failf(KB2_ERR_CUDA, "%s: %s", #expr, nccl().GetErrorString(_r));                  \
  } while (0)

extern "C" {

KB2_API int kb2_comm_unique_id(void* out128) {
  if (!out128) return failf(KB2_ERR_VALUE, "null argument");
  if (!nccl().ok) return failf(KB2_ERR_STATE, "NCCL (libnccl.so.2) could not be loaded");
  NcclUniqueId id;
  NCCL_TRY(nccl().GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return KB2_OK;
}

KB2_API int kb2_comm_init(const void* unique_id128, int32_t rank, int32_t num_ranks, int32_t device, kb2_comm** out) {
  if (!unique_id128 || !out) return failf(KB2_ERR_VALUE, "null argument");
  if (num_ranks < 1 || rank < 0 || rank >= num_ranks) return failf(KB2_ERR_VALUE, "bad rank %d/%d", rank, num_ranks);
  if (!nccl().ok) return failf(KB2_ERR_STATE, "NCCL (libnccl.so.2) could not be loaded");
  if (cudaSetDevice(device) != cudaSuccess) return failf(KB2_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  NcclUniqueId id;
  memcpy(&id, unique_id128, sizeof(id));
  kb2_comm* c = new kb2_comm();
  c->rank = rank; c->nranks = num_ranks; c->device = device;
  int r = nccl().CommInitRank(&c->comm, num_ranks, id, rank);
  if (r != 0) { delete c; return failf(KB2_ERR_CUDA, "ncclCommInitRank: %s", nccl().GetErrorString(r)); }
  *out = c;
  return KB2_OK;
}

KB2_API void kb2_comm_destroy(kb2_comm* c) {
  if (!c) return;
  if (c->barrier_word) cudaFree(c->barrier_word);
  if (c->comm) nccl().CommDestroy(c->comm);
  delete c;
}

KB2_API int kb2_comm_all_gather(kb2_comm* c, const void* send_dev, void* recv_dev, size_t bytes_per_rank, void* stream) {
  if (!c || !send_dev || !recv_dev) return failf(KB2_ERR_VALUE, "null argument");
  NCCL_TRY(nccl().AllGather(send_dev, recv_dev, bytes_per_rank, kNcclUint8, c->comm, (cudaStream_t)stream));
  return KB2_OK;
}

KB2_API int kb2_comm_reduce_scatter_bf16(kb2_comm* c, const void* send_dev, void* recv_dev, size_t elems_per_rank, void* stream) {
  if (!c || !send_dev || !recv_dev) return failf(KB2_ERR_VALUE, "null argument");
  NCCL_TRY(nccl().ReduceScatter(send_dev, recv_dev, elems_per_rank, kNcclBfloat16, kNcclSum, c->comm, (cudaStream_t)stream));
  return KB2_OK;
}

KB2_API int kb2_comm_all_reduce_bf16(kb2_comm* c, const void* send_dev, void* recv_dev, size_t elems, void* stream) {
  if (!c || !send_dev || !recv_dev) return failf(KB2_ERR_VALUE, "null argument");
  NCCL_TRY(nccl().AllReduce(send_dev, recv_dev, elems, kNcclBfloat16, kNcclSum, c->comm, (cudaStream_t)stream));
  return KB2_OK;
}

KB2_API int kb2_comm_broadcast(kb2_comm* c, void* buf_dev, size_t bytes, int32_t root, void* stream) {
  if (!c || !buf_dev) return failf(KB2_ERR_VALUE, "null argument");
  NCCL_TRY(nccl().Broadcast(buf_dev, buf_dev, bytes, kNcclUint8, root, c->comm, (cudaStream_t)stream));
  return KB2_OK;
}

KB2_API int kb2_comm_reduce_bf16(kb2_comm* c, const void* send_dev, void* recv_dev, size_t elems, int32_t root, void* stream) {
  if (!c || !send_dev) return failf(KB2_ERR_VALUE, "null argument");
  if (root < 0 || root >= c->nranks) return failf(KB2_ERR_VALUE, "reduce: root %d outside [0, %d)", root, c->nranks);
  if (c->rank == root && !recv_dev) return failf(KB2_ERR_VALUE, "reduce: the root needs a receive buffer");
  NCCL_TRY(nccl().Reduce(send_dev, recv_dev, elems, kNcclBfloat16, kNcclSum, root, c->comm, (cudaStream_t)stream));
  return KB2_OK;
}

// Cross-rank barrier on `stream`: a 4-byte all-reduce.  A rank leaves it only after every rank's stream has reached it, i.e. after
// every kernel the peers enqueued before it (and its peer-memory stores) has completed.
KB2_API int kb2_comm_barrier(kb2_comm* c, void* stream) {
  if (!c) return failf(KB2_ERR_VALUE, "null argument");
  if (!c->barrier_word) {
    if (cudaMalloc(&c->barrier_word, 64) != cudaSuccess) return failf(KB2_ERR_CUDA, "cudaMalloc failed (barrier word)");
    cudaMemset(c->barrier_word, 0, 64);
  }
  NCCL_TRY(nccl().AllReduce(c->barrier_word, c->barrier_word, 1, kNcclUint8, kNcclSum, c->comm, (cudaStream_t)stream));
  return KB2_OK;
}

// Collective: every rank allocates `bytes` of device memory (cudaMalloc, zero-filled), the CUDA IPC handles travel by all-gather, and
// ptrs_out_host[r] receives a pointer through which THIS rank can load / store rank r's buffer (ptrs_out_host[rank] = the local one).
// One process per GPU on one node (NVLink / NVSwitch peer access).  Free with kb2_comm_peer_free on every rank.
KB2_API int kb2_comm_peer_alloc(kb2_comm* c, size_t bytes, void** ptrs_out_host) {
  if (!c || !ptrs_out_host || bytes == 0) return failf(KB2_ERR_VALUE, "null argument");
  if (cudaSetDevice(c->device) != cudaSuccess) return failf(KB2_ERR_CUDA, "cudaSetDevice failed");
  void* local = nullptr;
  if (cudaMalloc(&local, bytes) != cudaSuccess) return failf(KB2_ERR_CUDA, "cudaMalloc(%zu) failed (peer buffer)", bytes);
  cudaMemset(local, 0, bytes);
  cudaIpcMemHandle_t mine;
  if (cudaIpcGetMemHandle(&mine, local) != cudaSuccess) { cudaFree(local); return failf(KB2_ERR_CUDA, "cudaIpcGetMemHandle failed"); }
  const size_t hs = sizeof(cudaIpcMemHandle_t);
  unsigned char* dbuf = nullptr;
  if (cudaMalloc(&dbuf, hs * c->nranks) != cudaSuccess) { cudaFree(local); return failf(KB2_ERR_CUDA, "cudaMalloc failed (handle exchange)"); }
  cudaMemcpy(dbuf + hs * c->rank, &mine, hs, cudaMemcpyHostToDevice);
  int rc = nccl().AllGather(dbuf + hs * c->rank, dbuf, hs, kNcclUint8, c->comm, (cudaStream_t)0);
  if (rc == 0 && cudaStreamSynchronize(0) != cudaSuccess) rc = -1;
  std::vector<cudaIpcMemHandle_t> all(c->nranks);
  if (rc == 0) cudaMemcpy(all.data(), dbuf, hs * c->nranks, cudaMemcpyDeviceToHost);
  cudaFree(dbuf);
  if (rc != 0) { cudaFree(local); return failf(KB2_ERR_CUDA, "IPC handle exchange failed"); }
  for (int r = 0; r < c->nranks; ++r) {
    if (r == c->rank) { ptrs_out_host[r] = local; continue; }
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      return failf(KB2_ERR_CUDA, "cudaIpcOpenMemHandle failed for rank %d (no peer access between the devices?)", r);
    }
    ptrs_out_host[r] = p;
  }
  return KB2_OK;
}

KB2_API int kb2_comm_peer_free(kb2_comm* c, void** ptrs_host) {
  if (!c || !ptrs_host) return failf(KB2_ERR_VALUE, "null argument");
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < c->nranks; ++r) {
    if (!ptrs_host[r]) continue;
    if (r == c->rank) cudaFree(ptrs_host[r]);
    else cudaIpcCloseMemHandle(ptrs_host[r]);
    ptrs_host[r] = nullptr;
  }
  return KB2_OK;
}

}  // extern "C"
