// Shared plumbing for libevc_b200: status/error strings, launch counting, small helpers.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <vector>

#include "../../include/evc.h"

namespace evc {

// sklearn _nmf.py:32 -- EPSILON = np.finfo(np.float32).eps
static constexpr float kEpsilon = 1.1920928955078125e-07f;

inline thread_local char g_err[512] = "";
inline std::atomic<long long> g_launches{0};

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define EVC_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return evc::fail(EVC_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,              \
                       cudaGetErrorString(e__));                                               \
  } while (0)

#define EVC_TRY(call)                \
  do {                               \
    int s__ = (call);                \
    if (s__ != EVC_OK) return s__;   \
  } while (0)

// Every kernel launch goes through this so evc_kernel_launch_count() is the truth.
#define EVC_LAUNCH_CHECK()                                                                      \
  do {                                                                                         \
    evc::g_launches.fetch_add(1, std::memory_order_relaxed);                                   \
    cudaError_t e__ = cudaGetLastError();                                                      \
    if (e__ != cudaSuccess)                                                                    \
      return evc::fail(EVC_ERR_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__,          \
                       cudaGetErrorString(e__));                                               \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }
inline size_t round_up_sz(size_t a, size_t b) { return (a + b - 1) / b * b; }

// A grow-only device buffer owned by a handle.
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int reserve(size_t need) {
    if (need <= bytes) return EVC_OK;
    if (p) {
      EVC_CUDA(cudaFree(p));
      p = nullptr;
      bytes = 0;
    }
    need = round_up_sz(need, 256);
    EVC_CUDA(cudaMalloc(&p, need));
    bytes = need;
    return EVC_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};


// Optional per-kernel-class device timing (evc_profile_enable / evc_profile_read).
struct Profiler {
  bool on = false;
  std::vector<cudaEvent_t> ev;
  size_t used = 0;
  struct Rec { int cls; size_t a, b; };
  std::vector<Rec> recs;
  size_t mark(cudaStream_t s) {
    if (used == ev.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ev.push_back(e);
    }
    cudaEventRecord(ev[used], s);
    return used++;
  }
  void release() {
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    ev.clear(); recs.clear(); used = 0;
  }
};
inline thread_local Profiler* g_prof = nullptr;
struct ProfScope {
  int cls; cudaStream_t s; size_t a = 0; bool on;
  ProfScope(int c, cudaStream_t st) : cls(c), s(st), on(g_prof && g_prof->on) { if (on) a = g_prof->mark(s); }
  ~ProfScope() { if (on) { size_t b = g_prof->mark(s); g_prof->recs.push_back({cls, a, b}); } }
};

}  // namespace evc
