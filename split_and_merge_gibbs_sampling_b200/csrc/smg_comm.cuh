// smg_comm.cuh -- the reductions over GPUs of SURVEY 8(e), in the C++ library (extern "C": smg_comm_* / smg_chains_*).
//
// Chains are independent: nothing is exchanged while sampling.  At the end of a run (or at a flush point) NCCL over
// NVLink reduces
//   * the posterior similarity matrix: ncclReduceScatter(sum, int32) so that rank g keeps the row block
//     [g n/G, (g+1) n/G) -- or ncclAllReduce when every rank wants the whole matrix;
//   * the histogram of the number of clusters: ncclAllReduce(sum, int64);
//   * the (count, mean, M2) moments of the two halves of every chain's trace: ncclAllGather -> split-R-hat
//     (Gelman et al., BDA3 section 11.4) on every rank.
// The reference has none of this (single process; its R scripts call mcclust.ext::comp.psm and LaplacesDemon::ESS
// after the run, realdata_analysis/zoo_simulator.R:205-215,339).
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy torch has already mapped when the caller is a torch
// process, the system one otherwise), so the library itself has no link-time dependency on it.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "smg_chain.cuh"

namespace smg {

struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      api.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.h) break;
    }
    if (!api.h) {
      api.err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
      return;
    }
    auto sym = [&](const char* s) -> void* {
      void* p = dlsym(api.h, s);
      if (!p && api.err.empty()) api.err = std::string("NCCL symbol missing: ") + s;
      return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.ReduceScatter = (decltype(api.ReduceScatter))sym("ncclReduceScatter");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  });
  return &api;
}

#define SMG_NCCL(call)                                                                                          \
  do {                                                                                                          \
    ncclResult_t _r = (call);                                                                                   \
    if (_r != ncclSuccess)                                                                                      \
      return smg::fail(SMG_ERR_CUDA, std::string("NCCL error: ") + smg::nccl_api()->GetErrorString(_r) + " at " + \
                                         __FILE__ + ":" + std::to_string(__LINE__));                            \
  } while (0)

}  // namespace smg

struct smg_comm {
  int rank = 0, world = 1, device = 0;
  ncclComm_t comm = nullptr;
  cudaStream_t st = nullptr;
  cudaEvent_t ev[2] = {};
  int* scratch = nullptr;  // [world * 256] device words for the warm-up / rendezvous collectives
};
