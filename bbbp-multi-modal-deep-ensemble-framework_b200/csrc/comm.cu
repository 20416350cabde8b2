// bbbp_comm_*: the two multi-GPU exchanges of the path over NCCL.
//
// Screening shards by whole reference batches and ends with ONE all-gather of fp32 scores; data-parallel training runs
// replicas and averages one flat gradient buffer.  Host code only.  NCCL is bound at first use with
// dlopen("libnccl.so.2") so that (a) the library has no link-time dependency on it and (b) a process that already carries
// an NCCL (torch's bundled copy) shares that one instance instead of loading a second.
#include <dlfcn.h>
#include <cstring>
#include <mutex>
#include "common.cuh"

namespace {

// the slice of nccl.h this file needs (ABI-stable since NCCL 2.10: ncclAvg)
struct UniqueId { char internal[128]; };
constexpr int kNcclFloat32 = 7, kNcclSum = 0, kNcclAvg = 4;

struct Nccl {
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    n.AllGather = reinterpret_cast<decltype(n.AllGather)>(dlsym(h, "ncclAllGather"));
    n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(dlsym(h, "ncclAllReduce"));
    n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllGather && n.AllReduce && n.GetErrorString;
  });
  return n;
}

int need_nccl(const char* what) {
  if (nccl().ok) return BBBP_OK;
  const char* why = dlerror();
  bbbp::set_error("%s: libnccl.so.2 not found (%s)", what, why ? why : "dlopen failed");
  return BBBP_EUNSUPPORTED;
}

int status(const char* what, int rc) {
  if (rc == 0) return BBBP_OK;
  bbbp::set_error("%s: NCCL error %d: %s", what, rc, nccl().GetErrorString(rc));
  return BBBP_ECUDA;
}

}  // namespace

extern "C" int bbbp_comm_unique_id(void* id128) {
  BBBP_CHECK_ARG(id128 != nullptr, "bbbp_comm_unique_id: NULL");
  if (int rc = need_nccl("bbbp_comm_unique_id")) return rc;
  return status("ncclGetUniqueId", nccl().GetUniqueId(static_cast<UniqueId*>(id128)));
}

extern "C" int bbbp_comm_init_rank(bbbp_comm_t* comm, int nranks, const void* id128, int rank) {
  BBBP_CHECK_ARG(comm && id128 && nranks >= 1 && rank >= 0 && rank < nranks, "bbbp_comm_init_rank: rank %d of %d", rank, nranks);
  if (int rc = need_nccl("bbbp_comm_init_rank")) return rc;
  UniqueId id;
  memcpy(&id, id128, sizeof(id));
  return status("ncclCommInitRank", nccl().CommInitRank(comm, nranks, id, rank));
}

extern "C" int bbbp_comm_destroy(bbbp_comm_t comm) {
  if (!comm) return BBBP_OK;
  if (int rc = need_nccl("bbbp_comm_destroy")) return rc;
  return status("ncclCommDestroy", nccl().CommDestroy(comm));
}

extern "C" int bbbp_comm_gather_scores(bbbp_comm_t comm, const float* send, float* recv, size_t count, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(comm && send && recv, "bbbp_comm_gather_scores: NULL argument");
  if (int rc = need_nccl("bbbp_comm_gather_scores")) return rc;
  return status("ncclAllGather", nccl().AllGather(send, recv, count, kNcclFloat32, comm, bbbp::as_stream(stream)));
}

extern "C" int bbbp_comm_average_gradients(bbbp_comm_t comm, float* flat, size_t count, bbbp_stream_t stream) {
  BBBP_CHECK_ARG(comm && flat, "bbbp_comm_average_gradients: NULL argument");
  if (int rc = need_nccl("bbbp_comm_average_gradients")) return rc;
  (void)kNcclSum;
  return status("ncclAllReduce", nccl().AllReduce(flat, flat, count, kNcclFloat32, kNcclAvg, comm, bbbp::as_stream(stream)));
}
