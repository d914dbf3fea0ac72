// Probe (not product): what ONE thread pays per TMA instruction (mbarrier expect_tx + cp.async.bulk.tensor.2d), with
// trivial coordinate arithmetic, and what a CTA ingests when 1 / 2 / 4 / 8 warps issue the same boxes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_issue_probe tools/probe/tma_issue_probe.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

constexpr int kSlots = 16;
// Every issuing warp owns kSlots/issuers slots; per slot: wait for the previous load of the slot, expect_tx, issue.
// per_stage = TMA instructions that share one barrier (a GEMM stage of per_stage boxes).
__global__ void __launch_bounds__(288, 1)
issue_probe(const __grid_constant__ CUtensorMap tm, int box_cols, int box_rows, int issuers, int per_stage, int rounds,
            long long* cycles, long long* issue_cycles) {
  extern __shared__ __align__(1024) uint8_t ring[];
  __shared__ __align__(8) uint64_t bar[kSlots];
  const uint32_t box_bytes = (uint32_t)box_cols * box_rows * 2u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) mbar_init(smem_u32(&bar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < issuers) {
    const uint32_t r0 = smem_u32(ring);
    const int stages = kSlots / per_stage;           // barriers in use
    int row = (blockIdx.x * 7 + w * 3) & 255;
    long long t_issue = 0;
    const long long t0 = clock64();
    uint32_t phase = 0;
    for (int r = 0; r < rounds; ++r) {
      for (int st = w; st < stages; st += issuers) {
        if (r > 0) mbar_wait(smem_u32(&bar[st]), phase ^ 1u);
        const long long a = clock64();
        mbar_expect(smem_u32(&bar[st]), box_bytes * per_stage);
#pragma unroll 1
        for (int j = 0; j < per_stage; ++j) {
          tma_2d(r0 + (st * per_stage + j) * box_bytes, &tm, ((r + j) & 15) * box_cols, row * box_rows, smem_u32(&bar[st]));
          row = (row + 37) & 255;
        }
        t_issue += clock64() - a;
      }
      phase ^= 1u;
    }
    for (int st = w; st < stages; st += issuers) mbar_wait(smem_u32(&bar[st]), phase ^ 1u);
    if (w == 0) { cycles[blockIdx.x] = clock64() - t0; issue_cycles[blockIdx.x] = t_issue; }
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int rows = 40960, cols = 576;
  const size_t pitch = (size_t)cols * 2, total = (size_t)rows * pitch;
  uint8_t* buf;
  cudaMalloc(&buf, total);
  cudaMemset(buf, 1, total);
  long long *cyc, *icyc;
  cudaMalloc(&cyc, 256 * sizeof(long long));
  cudaMalloc(&icyc, 256 * sizeof(long long));
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  PFN_encodeTiled enc = (PFN_encodeTiled)fp;
  cudaFuncSetAttribute(issue_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("# 16 slots in flight; 'per instr' = issuing thread's cycles per (share of expect_tx + one TMA instruction)\n");
  printf("# %-20s %7s %9s | %9s %10s %8s\n", "box", "issuers", "per_stage", "B/clk/SM", "per instr", "TB/s");
  struct Case { const char* name; int box_cols, box_rows; };
  const Case cases[] = {{"64B x 128 rows", 32, 128}, {"128B x 64 rows", 64, 64}, {"64B x 32 rows", 32, 32}};
  for (const Case& c : cases) {
    const int box_bytes = c.box_cols * c.box_rows * 2;
    CUtensorMap tm{};
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {(cuuint32_t)c.box_cols, (cuuint32_t)c.box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     c.box_cols * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
    for (int per_stage : {1, 4}) {
      for (int issuers : {1, 2, 4, 8}) {
        if (kSlots / per_stage < issuers) continue;
        const int rounds = 400;
        float best = 1e30f; long long bc = 0, bi = 0;
        for (int rep = 0; rep < 3; ++rep) {
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0); cudaEventCreate(&e1);
          cudaEventRecord(e0);
          issue_probe<<<148, 288, kSlots * box_bytes + 1024>>>(tm, c.box_cols, c.box_rows, issuers, per_stage, rounds, cyc, icyc);
          cudaEventRecord(e1);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          long long h[148], hi[148];
          cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
          cudaMemcpy(hi, icyc, sizeof(hi), cudaMemcpyDeviceToHost);
          long long mx = 0, mi = 0;
          for (int i = 0; i < 148; ++i) { mx = h[i] > mx ? h[i] : mx; mi += hi[i]; }
          if (ms < best) { best = ms; bc = mx; bi = mi / 148; }
        }
        const double bytes_cta = (double)rounds * kSlots * box_bytes;
        const double instr_warp0 = (double)rounds * (kSlots / per_stage + issuers - 1) / issuers * per_stage;  // approx
        printf("  %-20s %7d %9d | %9.1f %10.0f %8.2f\n", c.name, issuers, per_stage, bytes_cta / bc, bi / instr_warp0,
               bytes_cta * 148 / (best * 1e-3) / 1e12);
      }
    }
  }
  return 0;
}
