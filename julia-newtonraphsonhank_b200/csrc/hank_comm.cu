// hank_comm.cu — the one collective of the design: all-gather of Jacobian / JVP column blocks
// computed by lane-sharded ranks (one process per GPU).  NCCL is resolved with dlopen at first
// use so single-GPU users need no NCCL at load time and the library binds to whichever
// libnccl.so.2 the host process already loaded (e.g. PyTorch's bundled one).
#include <dlfcn.h>
#include <cstring>
#include "hank_ctx.h"
#include "../../include/hankb200.h"

namespace {

typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId { char internal[128]; };
typedef int ncclResult_t;
enum { ncclFloat64 = 8 };

struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      api.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.h) break;
    }
    if (api.h) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.h, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.h, "ncclCommInitRank");
      api.AllGather = (decltype(api.AllGather))dlsym(api.h, "ncclAllGather");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.h, "ncclCommDestroy");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.h, "ncclGetErrorString");
      if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy) api.h = nullptr;
    }
  }
  return api.h ? &api : nullptr;
}

int nccl_fail(hank_ctx* c, ncclResult_t r, const char* what) {
  NcclApi* a = nccl();
  std::string msg = std::string(what) + ": NCCL error";
  if (a && a->GetErrorString) msg += std::string(" ") + a->GetErrorString(r);
  return hank::set_error(c, HANK_ERR_CUDA, msg);
}

}  // namespace

extern "C" {

int hank_comm_unique_id(void* id128) {
  NcclApi* a = nccl();
  if (!a || !id128) return HANK_ERR_CUDA;
  ncclUniqueId id;
  if (a->GetUniqueId(&id) != 0) return HANK_ERR_CUDA;
  std::memcpy(id128, &id, 128);
  return HANK_OK;
}

int hank_comm_init(hank_ctx* c, int nranks, int rank, const void* id128) {
  if (!c || nranks < 1 || rank < 0 || rank >= nranks) return HANK_ERR_ARG;
  c->nranks = nranks; c->rank = rank;
  if (nranks == 1) return HANK_OK;
  NcclApi* a = nccl();
  if (!a) return hank::set_error(c, HANK_ERR_CUDA, "libnccl.so.2 could not be loaded");
  if (!id128) return hank::set_error(c, HANK_ERR_ARG, "missing NCCL unique id");
  int rc = hank::cuda_check(c, cudaSetDevice(c->device), "cudaSetDevice");
  if (rc) return rc;
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  ncclComm_t comm = nullptr;
  ncclResult_t r = a->CommInitRank(&comm, nranks, id, rank);
  if (r != 0) return nccl_fail(c, r, "ncclCommInitRank");
  c->nccl_comm = comm;
  return HANK_OK;
}

int hank_allgather_columns_dev(hank_ctx* c, const double* local, size_t count, double* all) {
  if (!c) return HANK_ERR_ARG;
  int rc = hank::cuda_check(c, cudaSetDevice(c->device), "cudaSetDevice");
  if (rc) return rc;
  if (c->nranks == 1) {
    if (all != local)
      return hank::cuda_check(c, cudaMemcpyAsync(all, local, count * sizeof(double), cudaMemcpyDeviceToDevice, c->stream),
                              "cudaMemcpyAsync");
    return HANK_OK;
  }
  NcclApi* a = nccl();
  if (!a || !c->nccl_comm) return hank::set_error(c, HANK_ERR_STATE, "hank_comm_init has not been called");
  ncclResult_t r = a->AllGather(local, all, count, ncclFloat64, (ncclComm_t)c->nccl_comm, c->stream);
  if (r != 0) return nccl_fail(c, r, "ncclAllGather");
  return HANK_OK;
}

// Host-pointer form: the block is staged through the context's pinned / device scratch, gathered, and copied back.
int hank_allgather_columns(hank_ctx* c, const double* local, size_t count, double* all) {
  if (!c || !local || !all) return HANK_ERR_ARG;
  int rc = hank::cuda_check(c, cudaSetDevice(c->device), "cudaSetDevice");
  if (rc) return rc;
  const size_t need = (size_t)(c->nranks + 1) * count * sizeof(double);
  if (c->gather_bytes < need) {
    if (c->d_gather) cudaFree(c->d_gather);
    c->d_gather = nullptr; c->gather_bytes = 0;
    rc = hank::cuda_check(c, cudaMalloc((void**)&c->d_gather, need), "cudaMalloc");
    if (rc) return rc;
    c->gather_bytes = need;
  }
  double* d_loc = c->d_gather; double* d_all = c->d_gather + count;
  rc = hank::cuda_check(c, cudaMemcpyAsync(d_loc, local, count * sizeof(double), cudaMemcpyHostToDevice, c->stream), "cudaMemcpyAsync");
  if (rc) return rc;
  rc = hank_allgather_columns_dev(c, d_loc, count, d_all);
  if (rc) return rc;
  rc = hank::cuda_check(c, cudaMemcpyAsync(all, d_all, (size_t)c->nranks * count * sizeof(double), cudaMemcpyDeviceToHost, c->stream),
                        "cudaMemcpyAsync");
  if (rc) return rc;
  return hank::cuda_check(c, cudaStreamSynchronize(c->stream), "cudaStreamSynchronize");
}

int hank_comm_destroy(hank_ctx* c) {
  if (!c) return HANK_ERR_ARG;
  if (c->nccl_comm) {
    NcclApi* a = nccl();
    if (a) a->CommDestroy((ncclComm_t)c->nccl_comm);
    c->nccl_comm = nullptr;
  }
  c->nranks = 1; c->rank = 0;
  return HANK_OK;
}

}  // extern "C"
