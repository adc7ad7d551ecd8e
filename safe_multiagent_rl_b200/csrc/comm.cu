// The one exchange of the multi-GPU path in the C ABI: an all-reduce (sum) of the additive stats vector over the
// ranks' env shards, feeding MetaAgent.update (safe_multi_agent_RL/meta_agent.py:32-39; main.py:65-68 calls it once
// per meta cycle).  NCCL is bound at run time (dlopen of libnccl.so.2, the SONAME both the system package and the
// PyTorch wheel install -- inside a PyTorch process this resolves to the copy that is already loaded), so
// libsmarl.so itself has no link-time dependency and single-GPU hosts never touch NCCL.
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

namespace smarl {

// the slice of nccl.h this file needs (stable since NCCL 2.0)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[SMARL_COMM_ID_BYTES]; } ncclUniqueId;
typedef int ncclResult_t;
enum { kNcclSuccess = 0, kNcclFloat64 = 8, kNcclSum = 0 };

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (api.handle) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
      api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
      api.GetVersion = (decltype(api.GetVersion))dlsym(api.handle, "ncclGetVersion");
      if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce) api.handle = nullptr;
    }
  }
  return api.handle ? &api : nullptr;
}

struct Comm {
  ncclComm_t nccl;
  int rank, world;
};

#define SMARL_NCCL(api, call)                                                                           \
  do {                                                                                                  \
    ncclResult_t r__ = (call);                                                                          \
    if (r__ != kNcclSuccess) {                                                                          \
      set_error("%s failed: %s", #call, (api)->GetErrorString ? (api)->GetErrorString(r__) : "NCCL error"); \
      return SMARL_ECUDA;                                                                               \
    }                                                                                                   \
  } while (0)

}  // namespace smarl

using namespace smarl;

extern "C" int smarl_comm_get_unique_id(void* id_out) {
  SMARL_REQUIRE(id_out != nullptr, "id_out is NULL");
  NcclApi* api = nccl_api();
  if (!api) {
    set_error("libnccl.so.2 could not be loaded (%s)", dlerror());
    return SMARL_EUNSUPPORTED;
  }
  ncclUniqueId id;
  SMARL_NCCL(api, api->GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return SMARL_OK;
}

extern "C" int smarl_comm_init_from_unique_id(SmarlComm** out, const void* id, int32_t rank, int32_t world_size) {
  SMARL_REQUIRE(out != nullptr && id != nullptr, "null pointer");
  SMARL_REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, "bad rank %d of %d", rank, world_size);
  NcclApi* api = nccl_api();
  if (!api) {
    set_error("libnccl.so.2 could not be loaded (%s)", dlerror());
    return SMARL_EUNSUPPORTED;
  }
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  Comm* c = new Comm{nullptr, rank, world_size};
  ncclResult_t r = api->CommInitRank(&c->nccl, world_size, uid, rank);     // collective: every rank calls it
  if (r != kNcclSuccess) {
    set_error("ncclCommInitRank failed: %s", api->GetErrorString ? api->GetErrorString(r) : "NCCL error");
    delete c;
    return SMARL_ECUDA;
  }
  *out = reinterpret_cast<SmarlComm*>(c);
  return SMARL_OK;
}

extern "C" void smarl_comm_destroy(SmarlComm* comm) {
  if (!comm) return;
  Comm* c = reinterpret_cast<Comm*>(comm);
  if (NcclApi* api = nccl_api()) api->CommDestroy(c->nccl);
  delete c;
}

extern "C" int smarl_comm_nccl_version(void) {
  NcclApi* api = nccl_api();
  int v = 0;
  if (api && api->GetVersion) api->GetVersion(&v);
  return v;
}

extern "C" int smarl_stats_allreduce(SmarlComm* comm, double* stats, int32_t n, smarl_stream_t stream) {
  SMARL_REQUIRE(comm != nullptr && stats != nullptr, "null pointer");
  SMARL_REQUIRE(n >= 1 && n <= 4 * SMARL_MAX_AGENTS + 1, "stats length %d outside 1..%d", n, 4 * SMARL_MAX_AGENTS + 1);
  Comm* c = reinterpret_cast<Comm*>(comm);
  NcclApi* api = nccl_api();
  SMARL_REQUIRE(api != nullptr, "NCCL is not loaded");
  // in place; every slot is a sum over envs (integer-valued cost sums and counts are exact in f64), so the result
  // equals the single-GPU vector over all envs and lambda comes out identical on every rank
  SMARL_NCCL(api, api->AllReduce(stats, stats, (size_t)n, kNcclFloat64, kNcclSum, c->nccl, (cudaStream_t)stream));
  return SMARL_OK;
}
