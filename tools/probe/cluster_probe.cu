// Probe (not product): shared::cluster address format of CTA-local shared addresses, mapa results, and how many
// clusters of 2 / 4 / 8 CTAs with ~224 KB of dynamic shared memory each are co-resident on this device.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__global__ void probe(int csize) {
  extern __shared__ uint8_t dyn[];
  __shared__ uint64_t bar;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (threadIdx.x == 0 && blockIdx.x < (unsigned)csize) {
    uint32_t a = smem_u32(&bar);
    printf("csize %d block %d rank %u smid %u: &bar=0x%08x dyn=0x%08x mapa(r0)=0x%08x mapa(r1)=0x%08x mapa(r%d)=0x%08x\n", csize, blockIdx.x, rank, smid,
           a, smem_u32(dyn), mapa(a, 0), mapa(a, 1), csize - 1, mapa(a, csize - 1));
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
int main() {
  const int smem = 224 * 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe, &cfg);
    printf("cluster size %2d: max active clusters %d (%d CTAs) [%s]\n", cs, n, n * cs, cudaGetErrorString(e));
    if (cs <= 8) {
      e = cudaLaunchKernelEx(&cfg, probe, cs);
      cudaError_t e2 = cudaDeviceSynchronize();
      printf("  launch: %s / %s\n", cudaGetErrorString(e), cudaGetErrorString(e2));
    }
  }
  return 0;
}
