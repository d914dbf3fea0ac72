// All-reduce of the partial A*H over NVLink peer memory (exemplar sharding) -- our own kernel
// instead of ncclAllReduce: each rank owns an IPC-shared buffer [flags | send | recv]; one launch per iteration
//   1. tells every peer "my partial is in my send region" (release store of the epoch into the peers' flags),
//   2. waits until all peers said so,
//   3. reduce-scatter + all-gather in one pass: rank r sums slice r of all ranks' send regions in rank order
//      (peer loads over NVLink, deterministic and identical on every rank) and stores the result into every
//      rank's recv region (peer stores),
//   4. publishes "my slice is written everywhere" and waits for the same from all peers, so the kernel ends
//      only when this rank's recv region is complete and nobody reads its send region any more.
// The message is small (T x F fp32, 1-4 MB): two flag round trips + ~(world-1)/world of the message over NVLink.
#pragma once
#include "evc_common.cuh"
#include <algorithm>
#include <cstdlib>

namespace evc {
namespace p2p {

constexpr int kMaxWorld = 8;
constexpr size_t kFlagBytes = 4096;  // ready[8] at +0, done[8] at +256 (unsigned int epochs)

struct Args {
  float* send[kMaxWorld];
  float* recv[kMaxWorld];
  unsigned int* flags[kMaxWorld];
  int rank, world;
  size_t n4;  // float4 elements in the message
  unsigned int epoch;
  unsigned int* block_counter;  // local words: [0] block counter, [1] / [2] go-words, [3] error (1 + rank waited for)
  long long timeout_cycles;     // a peer that does not show up within this many SM cycles is reported, not trapped on
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// Poll a flag written by a peer over NVLink.  Only a handful of threads per GPU do this (with back-off): heavy
// system-scope polling from every block slowed the peers' remote flag writes down by hundreds of microseconds.
// A peer that never arrives (it raised, its host stalled) must not take the CUDA context down: after the time-out the
// wait gives up, records which rank it waited for in the error word and lets the kernel finish (with a meaningless
// sum); the host reads the word at its next synchronisation point and returns EVC_ERR_COMM.
__device__ __forceinline__ void wait_flags(const unsigned int* flags, int idx, unsigned int epoch, long long timeout,
                                           unsigned int* err) {
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(flags + idx) - epoch) < 0) {
    __nanosleep(64);
    if (clock64() - t0 > timeout) {
      atomicExch(err, 1u + (unsigned int)idx);
      return;
    }
  }
}
// Local (same GPU) go-word: one block polls the peers, the others wait here.
// (the launch is cooperative, so the polling block is resident and always sets the word -- at the latest when its own
// wait times out)
__device__ __forceinline__ void wait_local(const unsigned int* word, unsigned int epoch) {
  unsigned int v;
  do {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(word) : "memory");
    if ((int)(v - epoch) >= 0) break;
    __nanosleep(32);
  } while (true);
}
__device__ __forceinline__ void set_local(unsigned int* word, unsigned int epoch) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(word), "r"(epoch) : "memory");
}

// local words: [0] block counter, [1] "all partials are ready", [2] "all slices are written"
__global__ void __launch_bounds__(256) allreduce_kernel(const Args a) {
  unsigned int* my_flags = a.flags[a.rank];
  unsigned int* words = a.block_counter;
  if (blockIdx.x == 0) {
    if (threadIdx.x < a.world) {
      __threadfence_system();  // the partial (written by the previous kernel in the stream) before the flag
      st_release_sys(a.flags[threadIdx.x] + a.rank, a.epoch);
      wait_flags(my_flags, threadIdx.x, a.epoch, a.timeout_cycles, words + 3);
    }
    __syncthreads();
    if (threadIdx.x == 0) set_local(words + 1, a.epoch);
  } else {
    if (threadIdx.x == 0) wait_local(words + 1, a.epoch);
    __syncthreads();
  }

  const size_t begin = a.n4 * a.rank / a.world, end = a.n4 * (a.rank + 1) / a.world;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += 2 * stride) {
    // all peer loads of two elements are issued before any is consumed (NVLink latency is ~2 us)
    const size_t i2 = i + stride;
    const bool two = i2 < end;
    float4 v[kMaxWorld], w[kMaxWorld];
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r)
      if (r < a.world) {
        v[r] = ld_peer(reinterpret_cast<const float4*>(a.send[r]) + i);
        if (two) w[r] = ld_peer(reinterpret_cast<const float4*>(a.send[r]) + i2);
      }
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), u = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r)  // fixed rank order: every rank gets bit-identical sums
      if (r < a.world) {
        s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w;
        if (two) { u.x += w[r].x; u.y += w[r].y; u.z += w[r].z; u.w += w[r].w; }
      }
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r)
      if (r < a.world) {
        reinterpret_cast<float4*>(a.recv[r])[i] = s;
        if (two) reinterpret_cast<float4*>(a.recv[r])[i2] = u;
      }
  }

  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(words, 1u);
    last = (prev == gridDim.x - 1);
    if (last) *words = 0u;  // for the next launch
  }
  __syncthreads();
  if (last) {
    // every block of this rank has written its share: tell the peers, wait for theirs, release the local blocks
    if (threadIdx.x < a.world) {
      __threadfence_system();
      st_release_sys(a.flags[threadIdx.x] + 64 + a.rank, a.epoch);
      wait_flags(my_flags + 64, threadIdx.x, a.epoch, a.timeout_cycles, words + 3);
    }
    __syncthreads();
    if (threadIdx.x == 0) set_local(words + 2, a.epoch);
  } else {
    if (threadIdx.x == 0) wait_local(words + 2, a.epoch);
    __syncthreads();
  }
}

struct State {
  int rank = 0, world = 1, t_max = 0;
  size_t region_bytes = 0;
  void* local = nullptr;
  void* peer[kMaxWorld] = {};
  unsigned int* block_counter = nullptr;
  unsigned int epoch = 0;
  bool attached = false;
  float* send_of(int r) const { return reinterpret_cast<float*>(static_cast<char*>(peer[r]) + kFlagBytes); }
  float* recv_of(int r) const { return reinterpret_cast<float*>(static_cast<char*>(peer[r]) + kFlagBytes + region_bytes); }
  float* send() const { return reinterpret_cast<float*>(static_cast<char*>(local) + kFlagBytes); }
  float* recv() const { return reinterpret_cast<float*>(static_cast<char*>(local) + kFlagBytes + region_bytes); }
};

inline int alloc_local(State* st, int max_frames, int ldWH, char handle_out[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (st->local) return fail(EVC_ERR_INVALID_ARGUMENT, "p2p exchange buffer already allocated");
  st->t_max = max_frames;
  st->region_bytes = round_up_sz((size_t)max_frames * ldWH * sizeof(float), 4096);
  const size_t bytes = kFlagBytes + 2 * st->region_bytes;
  EVC_CUDA(cudaMalloc(&st->local, bytes));
  EVC_CUDA(cudaMemset(st->local, 0, bytes));
  EVC_CUDA(cudaMalloc(&st->block_counter, 4 * sizeof(unsigned int)));
  EVC_CUDA(cudaMemset(st->block_counter, 0, 4 * sizeof(unsigned int)));
  cudaIpcMemHandle_t h;
  EVC_CUDA(cudaIpcGetMemHandle(&h, st->local));
  memcpy(handle_out, &h, 64);
  return EVC_OK;
}

inline int attach(State* st, const char* handles, int rank, int world) {
  if (!st->local) return fail(EVC_ERR_INVALID_ARGUMENT, "p2p: allocate the local exchange buffer first");
  if (world < 2 || world > kMaxWorld || rank < 0 || rank >= world)
    return fail(EVC_ERR_UNSUPPORTED, "p2p all-reduce supports 2..%d ranks on one node", kMaxWorld);
  st->rank = rank; st->world = world;
  for (int r = 0; r < world; ++r) {
    if (r == rank) { st->peer[r] = st->local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * 64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(&st->peer[r], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (int q = 0; q < r; ++q)
        if (q != rank && st->peer[q]) { cudaIpcCloseMemHandle(st->peer[q]); st->peer[q] = nullptr; }
      return fail(EVC_ERR_COMM, "cudaIpcOpenMemHandle(rank %d) -> %s", r, cudaGetErrorString(e));
    }
  }
  st->attached = true;
  return EVC_OK;
}

// Unmap the peers' buffers and go back to NCCL (the local buffer stays allocated): used when not every rank could attach.
inline void detach(State* st) {
  for (int r = 0; r < st->world; ++r)
    if (r != st->rank && st->peer[r]) { cudaIpcCloseMemHandle(st->peer[r]); st->peer[r] = nullptr; }
  st->attached = false;
}

// Host side of the device time-out: call after the stream has been synchronised.
inline int poll_error(State* st) {
  if (!st->attached || !st->block_counter) return EVC_OK;
  unsigned int e = 0;
  EVC_CUDA(cudaMemcpy(&e, st->block_counter + 3, sizeof(e), cudaMemcpyDeviceToHost));
  if (e) {
    cudaMemset(st->block_counter + 3, 0, sizeof(e));
    return fail(EVC_ERR_COMM, "peer-memory all-reduce: rank %u did not arrive within the time-out (EVC_P2P_TIMEOUT_S)", e - 1);
  }
  return EVC_OK;
}

inline void release(State* st) {
  for (int r = 0; r < st->world; ++r)
    if (r != st->rank && st->peer[r]) cudaIpcCloseMemHandle(st->peer[r]);
  if (st->local) cudaFree(st->local);
  if (st->block_counter) cudaFree(st->block_counter);
  *st = State{};
}

// In: this rank's partial in send(); out: the sum over ranks in recv() (on every rank).  `count` floats, multiple of 4.
inline int all_reduce(State* st, size_t count, cudaStream_t s) {
  Args a{};
  for (int r = 0; r < st->world; ++r) {
    a.send[r] = st->send_of(r);
    a.recv[r] = st->recv_of(r);
    a.flags[r] = reinterpret_cast<unsigned int*>(st->peer[r]);
  }
  a.rank = st->rank; a.world = st->world; a.n4 = count / 4; a.epoch = ++st->epoch; a.block_counter = st->block_counter;
  static const double timeout_s = getenv("EVC_P2P_TIMEOUT_S") ? atof(getenv("EVC_P2P_TIMEOUT_S")) : 60.0;
  a.timeout_cycles = (long long)(timeout_s * 2.0e9);
  const size_t per_rank = (a.n4 + st->world - 1) / st->world;
  int blocks = (int)((per_rank + 255) / 256);
  blocks = (blocks + 1) / 2;  // two elements per thread
  // the blocks wait on one another (go-words): a cooperative launch guarantees they are all resident, and the grid
  // is bounded by what the device can hold at once
  int dev = 0, sms = 0, per_sm = 0;
  EVC_CUDA(cudaGetDevice(&dev));
  EVC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  EVC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, allreduce_kernel, 256, 0));
  const int cap = std::max(1, std::min(120, sms * std::max(per_sm, 1)));
  blocks = blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
  void* params[] = {&a};
  EVC_CUDA(cudaLaunchCooperativeKernel((void*)allreduce_kernel, dim3(blocks), dim3(256), params, 0, s));
  EVC_LAUNCH_CHECK();
  return EVC_OK;
}

}  // namespace p2p
}  // namespace evc
