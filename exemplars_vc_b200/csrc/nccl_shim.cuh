// Minimal NCCL binding resolved at run time from the libnccl.so.2 already loaded in the process
// (torch bundles 2.28.9).  Only what the exemplar-sharded path needs: one communicator per rank and
// an in-place fp32 sum all-reduce of the partial A*H on the caller's stream.
#pragma once
#include "evc_common.cuh"
#include <dlfcn.h>

namespace evc {
namespace nccl {

struct UniqueId { char internal[128]; };
typedef void* Comm;

struct Api {
  void* lib = nullptr;
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
  int (*CommDestroy)(Comm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

inline int load(Api** out) {
  static Api api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      api.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD);
      if (!api.lib) api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (api.lib) {
      api.GetUniqueId = (int (*)(UniqueId*))dlsym(api.lib, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(Comm*, int, UniqueId, int))dlsym(api.lib, "ncclCommInitRank");
      api.AllReduce = (int (*)(const void*, void*, size_t, int, int, Comm, cudaStream_t))dlsym(api.lib, "ncclAllReduce");
      api.CommDestroy = (int (*)(Comm))dlsym(api.lib, "ncclCommDestroy");
      api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
    }
  }
  if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy)
    return fail(EVC_ERR_COMM, "libnccl.so.2 is not loadable in this process (%s)", dlerror() ? dlerror() : "missing symbol");
  *out = &api;
  return EVC_OK;
}

#define EVC_NCCL(api, call)                                                                     \
  do {                                                                                          \
    int r__ = (call);                                                                           \
    if (r__ != 0)                                                                               \
      return fail(EVC_ERR_COMM, "%s -> %s", #call, (api)->GetErrorString ? (api)->GetErrorString(r__) : "nccl error"); \
  } while (0)

inline int unique_id(char id_out[128]) {
  Api* a;
  EVC_TRY(load(&a));
  UniqueId id;
  EVC_NCCL(a, a->GetUniqueId(&id));
  memcpy(id_out, id.internal, 128);
  return EVC_OK;
}

inline int init_rank(Comm* c, const char id[128], int rank, int world) {
  Api* a;
  EVC_TRY(load(&a));
  UniqueId uid;
  memcpy(uid.internal, id, 128);
  EVC_NCCL(a, a->CommInitRank(c, world, uid, rank));
  return EVC_OK;
}

inline int all_reduce_sum(Comm c, float* buf, size_t count, cudaStream_t s) {
  Api* a;
  EVC_TRY(load(&a));
  EVC_NCCL(a, a->AllReduce(buf, buf, count, /*ncclFloat32*/ 7, /*ncclSum*/ 0, c, s));
  return EVC_OK;
}

inline void destroy(Comm c) {
  Api* a;
  if (load(&a) == EVC_OK && c) a->CommDestroy(c);
}

}  // namespace nccl
}  // namespace evc
