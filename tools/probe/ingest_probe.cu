// Probe (not product): how many bytes per clock one SM can take in from L2 through TMA, as a function of the row
// segment a box fetches (64-byte rows = the 32-element K-blocks of the fp32-accurate mode, 128-byte rows, contiguous
// 1-D bulk copies), the bytes in flight and the number of CTAs.  No consumer: a slot's load is re-issued as soon as
// it lands.  The buffer is the size of contraction 2's dictionary operand (47 MB: L2 resident after the first sweep).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ingest_probe tools/probe/ingest_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
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
__device__ __forceinline__ void bulk_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar) : "memory");
}

constexpr int kMaxStages = 64;
// mode 0: 1-D bulk copies of box_bytes; mode 1: 2-D boxes (box_cols x box_rows elements of 2 bytes)
__global__ void __launch_bounds__(128, 1)
ingest(const __grid_constant__ CUtensorMap tm, const uint8_t* base, size_t total_bytes, int mode, int box_cols, int box_rows,
       int kblocks, int row_blocks, int stages, int loads, int issuers, int lanes_mode, long long* cycles, long long* burst) {
  extern __shared__ __align__(1024) uint8_t ring[];
  __shared__ __align__(8) uint64_t bar[kMaxStages];
  const uint32_t box_bytes = (uint32_t)box_cols * box_rows * 2u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(smem_u32(&bar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // `issuers` threads (one per warp) each own the slots s = w, w + issuers, ...
  const int w = lanes_mode ? (int)threadIdx.x : (int)(threadIdx.x >> 5);
  if ((lanes_mode ? threadIdx.x < 32 : (threadIdx.x & 31) == 0) && w < issuers) {
    const long long t0 = clock64();
    const uint32_t r0 = smem_u32(ring);
    auto issue = [&](int i, int s) {
      // load i of this CTA: walk the (row block, k-block) tiles like a persistent GEMM CTA does
      const long long tile = (long long)blockIdx.x + (long long)(i / kblocks) * gridDim.x;
      const int rb = (int)(tile % row_blocks), kb = i % kblocks;
      mbar_expect(smem_u32(&bar[s]), box_bytes);
      if (mode == 1) tma_2d(r0 + s * box_bytes, &tm, kb * box_cols, rb * box_rows, smem_u32(&bar[s]));
      else bulk_1d(r0 + s * box_bytes, base + (((size_t)tile * kblocks + kb) * box_bytes) % (total_bytes - box_bytes) / 16 * 16, box_bytes, smem_u32(&bar[s]));
    };
    int n_mine = 0;
    for (int s = w; s < stages; s += issuers) issue(s, s), ++n_mine;
    if (w == 0) burst[blockIdx.x] = (clock64() - t0) / n_mine;  // cycles per back-to-back issue (no waits in between)
    uint32_t phase = 0;
    for (int i = stages; i < loads; i += stages) {
      for (int s = w; s < stages; s += issuers) {
        mbar_wait(smem_u32(&bar[s]), phase);
        if (i + s < loads) issue(i + s, s);
      }
      phase ^= 1u;
    }
    // drain
    for (int s = w; s < stages; s += issuers) mbar_wait(smem_u32(&bar[s]), phase);
    if (w == 0) cycles[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int rows = 40960, cols = 576;  // 2 planes x 20480 exemplars, pitch 1152 bytes (bf16): contraction 2's operand
  const size_t pitch = (size_t)cols * 2, total = (size_t)rows * pitch;
  uint8_t* buf;
  cudaMalloc(&buf, total);
  cudaMemset(buf, 1, total);
  long long *cyc, *burst;
  cudaMalloc(&cyc, 256 * sizeof(long long));
  cudaMalloc(&burst, 256 * sizeof(long long));
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  PFN_encodeTiled enc = (PFN_encodeTiled)fp;
  cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("# buffer %.1f MB, device max clock %d MHz\n", total / 1e6, clk_khz / 1000);
  printf("# issuers: nW = n warps (one lane each), nL = n lanes of one warp; burst = cycles per back-to-back issue of the first `stages` loads\n");
  printf("# %-22s %5s %6s %7s %6s | %8s %9s %12s %8s\n", "pattern", "CTAs", "stages", "inflt", "issuer", "TB/s", "B/clk/SM", "clk(MHz,est)", "burst");
  struct Case { const char* name; int mode, box_cols, box_rows; };
  const Case cases[] = {
      {"2D 64B x 128 rows", 1, 32, 128},   {"2D 128B x 128 rows", 1, 64, 128}, {"2D 128B x 64 rows", 1, 64, 64},
      {"2D 64B x 256 rows", 1, 32, 256},   {"2D 64B x 32 rows", 1, 32, 32},    {"1D bulk 2 KB", 0, 1024, 1},
      {"1D bulk 16 KB", 0, 8192, 1},
  };
  for (const Case& c : cases) {
    const int box_bytes = c.box_cols * c.box_rows * 2;
    for (int ctas : {148}) {
      for (int inflight_kb : {64, 192}) {
        for (int issuers : {1, 2, 4, 8, -4}) {
          const int lanes_mode = issuers < 0;
          if (lanes_mode) issuers = -issuers;
          const int stages = inflight_kb * 1024 / box_bytes;
          if (stages < 1 || stages > kMaxStages || (issuers > stages)) continue;
          if (ctas != 148 && (inflight_kb != 192 || issuers != 1)) continue;
          CUtensorMap tm{};
          int kblocks = 1, row_blocks = 1;
          if (c.mode == 1) {
            cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
            cuuint64_t strides[1] = {(cuuint64_t)pitch};
            cuuint32_t box[2] = {(cuuint32_t)c.box_cols, (cuuint32_t)c.box_rows};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             c.box_cols * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
            kblocks = cols / c.box_cols; row_blocks = rows / c.box_rows;
          } else {
            kblocks = 16; row_blocks = 1 << 20;
          }
          const int loads = (int)(24.0e6 / box_bytes / 1) / stages * stages;  // 24 MB per CTA
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0); cudaEventCreate(&e1);
          float best = 1e30f; long long best_cyc = 0, best_burst = 0;
          for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            ingest<<<ctas, 128, stages * box_bytes + 1024>>>(tm, buf, total, c.mode, c.box_cols, c.box_rows, kblocks, row_blocks, stages,
                                                            loads, issuers, lanes_mode, cyc, burst);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long h[256]; cudaMemcpy(h, cyc, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
            if (ms < best) { best = ms; best_cyc = mx; cudaMemcpy(&best_burst, burst, sizeof(long long), cudaMemcpyDeviceToHost); }
          }
          const double bytes = (double)loads * box_bytes * ctas;
          printf("  %-22s %5d %6d %5dKB %5d%s | %8.2f %9.1f %12.0f %8lld\n", c.name, ctas, stages, inflight_kb, issuers, lanes_mode ? "L" : "W",
                 bytes / (best * 1e-3) / 1e12, (double)loads * box_bytes / (double)best_cyc, best_cyc / (best * 1e-3) / 1e6, best_burst);
        }
      }
    }
  }
  return 0;
}
