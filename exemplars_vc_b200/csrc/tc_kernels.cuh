// Tensor-core kernels (sm_100a): both contractions of the multiplicative update as tcgen05.mma
// GEMMs with TMEM accumulators, operands staged by TMA through an mbarrier ring.
//
// Orientation (shared by every kernel here): the DICTIONARY dimension is the MMA M dimension (TMEM
// lanes), the FRAME dimension is the MMA N dimension (TMEM columns), both operands are K-major:
//
//   contraction 1 (sklearn _nmf.py:554, WH = W@A):     D[f, t] = sum_n AT[f, n] * H[t, n]     K = N
//   contraction 2 (sklearn _nmf.py:585, R@A.T):        D[n, t] = sum_f A [n, f] * R[t, f]     K = F
//   conversion    (04_align_n_nmf.py:391, H.T@B):      D[f, t] = sum_n BT[f, n] * H[t, n]     K = N
//
// so an epilogue thread owns one dictionary row (one TMEM lane) and walks frames; with H, WH and R
// stored frame-major (T, ld) a warp touches 32 consecutive floats per frame: coalesced.
//
// fp32-accurate mode (3xTF32): x = hi + lo with hi = x truncated to tf32 (the tensor core ignores the
// low 13 mantissa bits of a 32-bit operand, so the raw fp32 tile IS the hi operand) and
// lo = x - hi (exact in fp32); D += M_lo*N_hi + M_hi*N_lo + M_hi*N_hi, dropping lo*lo (2^-22 relative).
#pragma once
#include "evc_common.cuh"
#include "umma.cuh"

namespace evc {
namespace tc {

using namespace umma;

inline int k_pitch(int F) { return round_up(F, 8); }

enum TcEpilogue { TEPI_PARTIAL = 0, TEPI_MU_KL = 1, TEPI_MU_FRO = 2 };

struct GemmParams {
  int M_total;  // dictionary-side rows: F for contraction 1 / conversion, N for contraction 2
  int T;        // frames
  int K;        // reduction length
  int num_m_groups, num_t_tiles, num_splits, kblocks_per_split, kblocks_total;
  float* out;    // PARTIAL: [split][t][m] with pitch ld_out;  MU_*: the activations H (T, ld_out)
  int ld_out;
  float* out_lo;  // MU_*: lo part of H for the 3xTF32 operand (nullable)
  const float* colsum;
  const float* num0;  // MU_FRO: cached numerator X A^T, same pitch as H
  float lam, eps;
  const unsigned char* row_active;
};

__device__ __forceinline__ float tf32_lo(float x) {
  return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
}

constexpr int kThreads = 192;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int kSmemBudget = 227 * 1024 - 2048;

template <int kMTiles, int kBlockT, int kBlockK, bool kSplit3>
struct TileCfg {
  static constexpr int kRowBytes = kBlockK * 4;
  static constexpr int kMTileBytes = 128 * kRowBytes;
  static constexpr int kNTileBytes = kBlockT * kRowBytes;
  static constexpr int kCopies = kSplit3 ? 2 : 1;
  static constexpr int kStageBytes = kCopies * (kMTiles * kMTileBytes + kNTileBytes);
  static constexpr int kStagesRaw = kSmemBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kAccCols = kMTiles * kBlockT;
  static constexpr int kAccStages = (512 / kAccCols) >= 2 ? 2 : 1;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024;  // + slack to align the ring to 1024 B
  static_assert(kRowBytes == 64 || kRowBytes == 128, "K block must be one 64B or 128B swizzle row");
  static_assert(kStages >= 2, "tile does not fit twice in shared memory");
  static_assert(kAccCols <= 512, "accumulators exceed TMEM");
  static_assert(kBlockT % 32 == 0 && kBlockT >= 32 && kBlockT <= 256, "bad frame tile");
};

template <int kMTiles, int kBlockT, int kBlockK, bool kSplit3, int kEpi>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmMlo,
               const __grid_constant__ CUtensorMap tmN, const __grid_constant__ CUtensorMap tmNlo, const GemmParams p) {
  using Cfg = TileCfg<kMTiles, kBlockT, kBlockK, kSplit3>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kAccStages = Cfg::kAccStages;
  constexpr uint32_t kIdesc = make_idesc(kFmtTF32, 128, kBlockT);

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kStages];
  __shared__ __align__(8) uint64_t bar_empty[kStages];
  __shared__ __align__(8) uint64_t bar_acc_full[kAccStages];
  __shared__ __align__(8) uint64_t bar_acc_empty[kAccStages];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmM);
    tma_prefetch_desc(&tmN);
    if (kSplit3) {
      tma_prefetch_desc(&tmMlo);
      tma_prefetch_desc(&tmNlo);
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(smem_u32(&bar_full[i]), 1);
      mbar_init(smem_u32(&bar_empty[i]), 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(smem_u32(&bar_acc_full[i]), 1);
      mbar_init(smem_u32(&bar_acc_empty[i]), 4);  // one elected lane of each epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_smem), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int num_items = p.num_m_groups * p.num_t_tiles * p.num_splits;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int t_tile = item % p.num_t_tiles;
        const int rest = item / p.num_t_tiles;
        const int m_group = rest % p.num_m_groups;
        const int split = rest / p.num_m_groups;
        const int m0 = m_group * (128 * kMTiles);
        const int t0 = t_tile * kBlockT;
        const int kb0 = split * p.kblocks_per_split;
        const int kb1 = min(kb0 + p.kblocks_per_split, p.kblocks_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          const uint32_t full = smem_u32(&bar_full[stage]);
          mbar_arrive_expect_tx(full, (uint32_t)Cfg::kStageBytes);
          const uint32_t sbase = ring + stage * Cfg::kStageBytes;
          const int kc = kb * kBlockK;
#pragma unroll
          for (int i = 0; i < kMTiles; ++i) {
            tma_load_2d(sbase + i * Cfg::kMTileBytes, &tmM, kc, m0 + i * 128, full, kEvictNormal);
            if (kSplit3)
              tma_load_2d(sbase + (kMTiles + i) * Cfg::kMTileBytes, &tmMlo, kc, m0 + i * 128, full, kEvictNormal);
          }
          const uint32_t nbase = sbase + Cfg::kCopies * kMTiles * Cfg::kMTileBytes;
          tma_load_2d(nbase, &tmN, kc, t0, full, kEvictNormal);
          if (kSplit3) tma_load_2d(nbase + Cfg::kNTileBytes, &tmNlo, kc, t0, full, kEvictNormal);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int split = item / (p.num_t_tiles * p.num_m_groups);
        const int kb0 = split * p.kblocks_per_split;
        const int kb1 = min(kb0 + p.kblocks_per_split, p.kblocks_total);
        mbar_wait(smem_u32(&bar_acc_empty[acc]), acc_phase ^ 1u);
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          const uint32_t sbase = ring + stage * Cfg::kStageBytes;
          const uint32_t nbase = sbase + Cfg::kCopies * kMTiles * Cfg::kMTileBytes;
          const int kvalid = min(kBlockK, p.K - kb * kBlockK);
          const int ksteps = (kvalid + 7) >> 3;
#pragma unroll
          for (int i = 0; i < kMTiles; ++i) {
            const uint32_t d = tmem_base + (uint32_t)((acc * kMTiles + i) * kBlockT);
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint32_t first = (kb > kb0 || ks > 0) ? 1u : 0u;
              const uint64_t a_hi = make_smem_desc(sbase + i * Cfg::kMTileBytes + ks * 32, Cfg::kRowBytes);
              const uint64_t b_hi = make_smem_desc(nbase + ks * 32, Cfg::kRowBytes);
              if (kSplit3) {
                const uint64_t a_lo =
                    make_smem_desc(sbase + (kMTiles + i) * Cfg::kMTileBytes + ks * 32, Cfg::kRowBytes);
                const uint64_t b_lo = make_smem_desc(nbase + Cfg::kNTileBytes + ks * 32, Cfg::kRowBytes);
                mma_tf32(d, a_lo, b_hi, kIdesc, first);
                mma_tf32(d, a_hi, b_lo, kIdesc, 1u);
                mma_tf32(d, a_hi, b_hi, kIdesc, 1u);
              } else {
                mma_tf32(d, a_hi, b_hi, kIdesc, first);
              }
            }
          }
          mma_commit(smem_u32(&bar_empty[stage]));  // frees the smem slot when these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        mma_commit(smem_u32(&bar_acc_full[acc]));  // accumulator complete -> epilogue
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ================= epilogue (4 warps; warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32)) =================
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int t_tile = item % p.num_t_tiles;
      const int rest = item / p.num_t_tiles;
      const int m_group = rest % p.num_m_groups;
      const int split = rest / p.num_m_groups;
      const int t0 = t_tile * kBlockT;
      mbar_wait(smem_u32(&bar_acc_full[acc]), acc_phase);
      tc_fence_after();
#pragma unroll
      for (int i = 0; i < kMTiles; ++i) {
        const int m = m_group * (128 * kMTiles) + i * 128 + quarter * 32 + lane;
        const bool m_ok = m < p.M_total;
        float den = 1.f;
        if (kEpi == TEPI_MU_KL) {
          den = (m_ok ? p.colsum[m] : 1.f) + p.lam;
          if (den == 0.f) den = p.eps;
        }
        for (int c = 0; c < kBlockT / 32; ++c) {
          const int tb = t0 + c * 32;
          if (tb >= p.T) break;  // warp-uniform
          uint32_t v[32];
          const uint32_t taddr =
              tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((acc * kMTiles + i) * kBlockT + c * 32);
          tmem_ld_32x32(taddr, v);
          tmem_ld_wait();
          if (kEpi == TEPI_PARTIAL) {
            float* o = p.out + ((size_t)split * p.T + tb) * p.ld_out + m;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (m_ok && tb + j < p.T) o[(size_t)j * p.ld_out] = __uint_as_float(v[j]);
          } else {
            float h[32];
            float* o = p.out + (size_t)tb * p.ld_out + m;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const bool ok = m_ok && (tb + j < p.T) && (!p.row_active || p.row_active[tb + j]);
              h[j] = ok ? o[(size_t)j * p.ld_out] : __int_as_float(0x7fc00000);
            }
            if (kEpi == TEPI_MU_FRO) {
              const float* q = p.num0 + (size_t)tb * p.ld_out + m;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (h[j] == h[j]) {
                  float dn = __uint_as_float(v[j]) + p.lam;
                  if (dn == 0.f) dn = p.eps;
                  h[j] = h[j] * __fdiv_rn(q[(size_t)j * p.ld_out], dn);
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) h[j] = h[j] * __fdiv_rn(__uint_as_float(v[j]), den);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const bool ok = m_ok && (tb + j < p.T) && (!p.row_active || p.row_active[tb + j]);
              if (ok) {
                o[(size_t)j * p.ld_out] = h[j];
                if (p.out_lo) p.out_lo[(size_t)(tb + j) * p.ld_out + m] = tf32_lo(h[j]);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[acc]));
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- memory-bound helpers ------------------------------------------------------------------------

// lo[i] = x[i] - trunc_tf32(x[i]) over a (rows, ld) matrix (pad columns included: they are zero).
__global__ void split_lo_kernel(const float* __restrict__ x, float* __restrict__ lo, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) lo[i] = tf32_lo(x[i]);
}

// WH[t,f] = sum_s P[s][t][f]   (fixed order: deterministic)
__global__ void reduce_partials_kernel(const float* __restrict__ P, int S, int T, int ldp, int F,
                                       float* __restrict__ WH, int ldwh) {
  const int f = blockIdx.y * blockDim.x + threadIdx.x;
  const int t = blockIdx.x;
  if (t >= T || f >= ldwh) return;
  float s = 0.f;
  if (f < F)
    for (int k = 0; k < S; ++k) s += P[((size_t)k * T + t) * ldp + f];
  WH[(size_t)t * ldwh + f] = s;
}

// R = X / max(WH, eps) with zeroed pad columns (+ lo part for 3xTF32); `copy` = 1 stores WH itself
// (Frobenius: the second contraction multiplies A^T with A H).
__global__ void ratio_split_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ WH, int ldwh,
                                   float* __restrict__ R, float* __restrict__ R_lo, int ldr, int T, int F, float eps,
                                   int copy) {
  const int f = blockIdx.y * blockDim.x + threadIdx.x;
  const int t = blockIdx.x;
  if (t >= T || f >= ldr) return;
  float r = 0.f;
  if (f < F) {
    const float wh = WH[(size_t)t * ldwh + f];
    r = copy ? wh : __fdiv_rn(X[(size_t)t * ldx + f], fmaxf(wh, eps));
  }
  R[(size_t)t * ldr + f] = r;
  if (R_lo) R_lo[(size_t)t * ldr + f] = tf32_lo(r);
}

// ---- host side -----------------------------------------------------------------------------------

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline int get_encode(PFN_encodeTiled* out) {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    EVC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess)
      return fail(EVC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    fn = (PFN_encodeTiled)p;
  }
  *out = fn;
  return EVC_OK;
}

// Row-major fp32 matrix (rows, cols) with pitch ld floats; box = box_rows x box_cols, box_cols*4 in {64,128}.
inline int make_tmap(CUtensorMap* m, const float* base, int rows, int cols, int ld, int box_cols, int box_rows) {
  PFN_encodeTiled enc;
  EVC_TRY(get_encode(&enc));
  if (((uintptr_t)base & 15) || (ld & 3))
    return fail(EVC_ERR_INVALID_ARGUMENT, "tensor-core modes need 16-byte aligned matrices with a pitch multiple of 4 floats");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = (box_cols * 4 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(EVC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d box=%dx%d", (int)r, rows, cols,
                ld, box_rows, box_cols);
  return EVC_OK;
}

inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

inline int check_device(int dev) {
  int major = 0;
  EVC_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return fail(EVC_ERR_UNSUPPORTED, "tensor-core modes need an sm_100 device (compute capability %d.x found)", major);
  return EVC_OK;
}

inline int check_alignment(int mode, const float* H, int ldH) {
  if (mode == EVC_MODE_FP32) return EVC_OK;
  if (((uintptr_t)H & 15) || (ldH & 3))
    return fail(EVC_ERR_INVALID_ARGUMENT, "tensor-core modes need H 16-byte aligned with ldH a multiple of 4");
  return EVC_OK;
}

template <int kMTiles, int kBlockT, int kBlockK, bool kSplit3, int kEpi>
inline int launch_tc(const CUtensorMap& tmM, const CUtensorMap& tmMlo, const CUtensorMap& tmN, const CUtensorMap& tmNlo,
                     const GemmParams& p, cudaStream_t s) {
  using Cfg = TileCfg<kMTiles, kBlockT, kBlockK, kSplit3>;
  auto kern = tc_gemm_kernel<kMTiles, kBlockT, kBlockK, kSplit3, kEpi>;
  static bool configured = false;
  if (!configured) {
    EVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int items = p.num_m_groups * p.num_t_tiles * p.num_splits;
  if (items <= 0) return EVC_OK;
  const int grid = items < num_sms() ? items : num_sms();
  kern<<<grid, kThreads, Cfg::kSmemBytes, s>>>(tmM, tmMlo, tmN, tmNlo, p);
  EVC_LAUNCH_CHECK();
  return EVC_OK;
}

// Tile shapes per contraction.  K block = one swizzle row: 16 fp32 (64 B) when both hi and lo tiles
// are staged (3xTF32), 32 fp32 (128 B) otherwise, so a ring stage is 48-64 KB and 3-4 stages fit.
constexpr int kC1MTiles = 2, kC1BlockT = 256;  // contraction 1 / conversion: 256 dictionary rows x 256 frames, split-K
constexpr int kC2MTiles = 1, kC2BlockT = 256;  // contraction 2: 128 exemplars x 256 frames, 2 accumulator stages
constexpr int kBlockK3 = 16, kBlockK1 = 32;

struct DictOperands {
  int F = 0, N = 0, ldA = 0, ldN = 0;
  bool has_target = false;
  const float* A = nullptr;  // (N, ldA), borrowed from the handle; raw fp32 = hi operand
  float* A_lo = nullptr;     // (N, ldA)
  float* AT = nullptr;       // (F, ldN) transposed copy: K-major operand of contraction 1
  float* AT_lo = nullptr;
  float* BT = nullptr;       // (F, ldN) transposed target dictionary: operand of the conversion
  float* BT_lo = nullptr;
  CUtensorMap tmA, tmA_lo, tmAT, tmAT_lo, tmBT, tmBT_lo;
  // per-solve state inside the workspace
  size_t hlo_floats = 0;     // lo part of H lives at the start of the workspace
  const float* hlo_for = nullptr;
  void release() {
    cudaFree(A_lo); cudaFree(AT); cudaFree(AT_lo); cudaFree(BT); cudaFree(BT_lo);
    A_lo = AT = AT_lo = BT = BT_lo = nullptr;
  }
};

inline int launch_split_lo(const float* x, float* lo, size_t n, cudaStream_t s) {
  if (!n) return EVC_OK;
  size_t blocks = (n + 1023) / 1024;
  if (blocks > 148 * 16) blocks = 148 * 16;
  split_lo_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, lo, n);
  EVC_LAUNCH_CHECK();
  return EVC_OK;
}

inline int build_operands(DictOperands* o, int mode, const float* A, const float* B, int ldA, int F, int N, cudaStream_t s) {
  if (mode == EVC_MODE_BF16) return fail(EVC_ERR_UNSUPPORTED, "EVC_MODE_BF16 is not implemented yet");
  const bool split3 = (mode == EVC_MODE_3XTF32);
  const int bk = split3 ? kBlockK3 : kBlockK1;
  o->F = F; o->N = N; o->ldA = ldA; o->ldN = round_up(N, 4); o->A = A; o->has_target = (B != nullptr);
  const size_t at_bytes = (size_t)F * o->ldN * sizeof(float);
  const size_t a_bytes = (size_t)N * ldA * sizeof(float);
  dim3 tb(32, 8), tg(ceil_div(N, 32), ceil_div(F, 32));
  EVC_CUDA(cudaMalloc(&o->AT, at_bytes));
  EVC_CUDA(cudaMemsetAsync(o->AT, 0, at_bytes, s));
  simt::transpose_kernel<<<tg, tb, 0, s>>>(A, ldA, o->AT, o->ldN, N, F);
  EVC_LAUNCH_CHECK();
  EVC_TRY(make_tmap(&o->tmA, A, N, F, ldA, bk, 128));
  EVC_TRY(make_tmap(&o->tmAT, o->AT, F, N, o->ldN, bk, 128));
  o->tmA_lo = o->tmA; o->tmAT_lo = o->tmAT;
  if (split3) {
    EVC_CUDA(cudaMalloc(&o->A_lo, a_bytes));
    EVC_CUDA(cudaMalloc(&o->AT_lo, at_bytes));
    EVC_TRY(launch_split_lo(A, o->A_lo, (size_t)N * ldA, s));
    EVC_TRY(launch_split_lo(o->AT, o->AT_lo, (size_t)F * o->ldN, s));
    EVC_TRY(make_tmap(&o->tmA_lo, o->A_lo, N, F, ldA, bk, 128));
    EVC_TRY(make_tmap(&o->tmAT_lo, o->AT_lo, F, N, o->ldN, bk, 128));
  }
  if (B) {
    EVC_CUDA(cudaMalloc(&o->BT, at_bytes));
    EVC_CUDA(cudaMemsetAsync(o->BT, 0, at_bytes, s));
    simt::transpose_kernel<<<tg, tb, 0, s>>>(B, ldA, o->BT, o->ldN, N, F);
    EVC_LAUNCH_CHECK();
    EVC_TRY(make_tmap(&o->tmBT, o->BT, F, N, o->ldN, bk, 128));
    o->tmBT_lo = o->tmBT;
    if (split3) {
      EVC_CUDA(cudaMalloc(&o->BT_lo, at_bytes));
      EVC_TRY(launch_split_lo(o->BT, o->BT_lo, (size_t)F * o->ldN, s));
      EVC_TRY(make_tmap(&o->tmBT_lo, o->BT_lo, F, N, o->ldN, bk, 128));
    }
  }
  return EVC_OK;
}

// Split-K plan of contraction 1: fill the SMs with (m_group, t_tile, split) work items.
struct C1Plan {
  int m_groups, t_tiles, splits, kb_total, kb_per_split, ldp;
};
inline C1Plan plan_c1(int F, int N, int T, int bk) {
  C1Plan pl;
  pl.m_groups = ceil_div(F, 128 * kC1MTiles);
  pl.t_tiles = ceil_div(T, kC1BlockT);
  pl.kb_total = ceil_div(N, bk);
  int base = pl.m_groups * pl.t_tiles;
  int s = num_sms() / base;
  if (s < 1) s = 1;
  if (s > pl.kb_total) s = pl.kb_total;
  pl.kb_per_split = ceil_div(pl.kb_total, s);
  pl.splits = ceil_div(pl.kb_total, pl.kb_per_split);
  pl.ldp = round_up(F, 32);
  return pl;
}

// Workspace layout: [ H_lo (T*ldH floats, 3xTF32 only) | split-K partials ].
inline size_t ws_partials_offset(const DictOperands& o, int mode, int T, int ldH) {
  return (mode == EVC_MODE_3XTF32) ? round_up_sz((size_t)T * ldH, 64) : 0;
}

// Called whenever H was (re)written by something other than the update kernel: (re)derive the lo
// operand of H and make sure the workspace can hold it plus the split-K partials.
inline int after_h_written(DictOperands& o, int mode, const float* H, int ldH, int T, DevBuf* ws, cudaStream_t s) {
  if (mode == EVC_MODE_FP32) return EVC_OK;
  const int bk = (mode == EVC_MODE_3XTF32) ? kBlockK3 : kBlockK1;
  const C1Plan pl = plan_c1(o.F, o.N, T, bk);
  const size_t need = ws_partials_offset(o, mode, T, ldH) + (size_t)pl.splits * T * pl.ldp;
  EVC_TRY(ws->reserve(need * sizeof(float)));
  ProfScope ps(3, s);
  if (mode == EVC_MODE_3XTF32) EVC_TRY(launch_split_lo(H, ws->as<float>(), (size_t)T * ldH, s));
  return EVC_OK;
}

template <bool kSplit3>
inline int contract_wh_t(DictOperands& o, int mode, const float* H, int ldH, int T, float* WH, int ldWH, bool target,
                         DevBuf* ws, cudaStream_t s) {
  constexpr int bk = kSplit3 ? kBlockK3 : kBlockK1;
  const C1Plan pl = plan_c1(o.F, o.N, T, bk);
  float* hlo = ws->as<float>();
  float* partials = ws->as<float>() + ws_partials_offset(o, mode, T, ldH);
  CUtensorMap tmH, tmHlo;
  EVC_TRY(make_tmap(&tmH, H, T, o.N, ldH, bk, kC1BlockT));
  tmHlo = tmH;
  if (kSplit3) EVC_TRY(make_tmap(&tmHlo, hlo, T, o.N, ldH, bk, kC1BlockT));
  GemmParams p{};
  p.M_total = o.F; p.T = T; p.K = o.N;
  p.num_m_groups = pl.m_groups; p.num_t_tiles = pl.t_tiles; p.num_splits = pl.splits;
  p.kblocks_per_split = pl.kb_per_split; p.kblocks_total = pl.kb_total;
  p.out = partials; p.ld_out = pl.ldp;
  {
    ProfScope ps(0, s);
    EVC_TRY((launch_tc<kC1MTiles, kC1BlockT, bk, kSplit3, TEPI_PARTIAL>(target ? o.tmBT : o.tmAT,
                                                                      target ? o.tmBT_lo : o.tmAT_lo, tmH, tmHlo, p, s)));
  }
  ProfScope ps(1, s);
  dim3 g(T, ceil_div(ldWH, 128));
  reduce_partials_kernel<<<g, 128, 0, s>>>(partials, pl.splits, T, pl.ldp, o.F, WH, ldWH);
  EVC_LAUNCH_CHECK();
  return EVC_OK;
}

inline int contract_wh(DictOperands& o, int mode, const float* H, int ldH, int T, float* WH, int ldWH, bool target,
                       DevBuf* ws, cudaStream_t s) {
  if (target && !o.has_target) return fail(EVC_ERR_INVALID_ARGUMENT, "no target dictionary");
  if (mode == EVC_MODE_3XTF32) return contract_wh_t<true>(o, mode, H, ldH, T, WH, ldWH, target, ws, s);
  return contract_wh_t<false>(o, mode, H, ldH, T, WH, ldWH, target, ws, s);
}

// Second contraction with a fused epilogue.  `Rsrc` is what multiplies A^T: the ratio (KL) or A H (Frobenius).
template <bool kSplit3, int kEpi>
inline int contract2_t(DictOperands& o, int T, const float* R, const float* R_lo, int ldR, GemmParams p, cudaStream_t s) {
  constexpr int bk = kSplit3 ? kBlockK3 : kBlockK1;
  CUtensorMap tmR, tmRlo;
  EVC_TRY(make_tmap(&tmR, R, T, o.F, ldR, bk, kC2BlockT));
  tmRlo = tmR;
  if (kSplit3) EVC_TRY(make_tmap(&tmRlo, R_lo, T, o.F, ldR, bk, kC2BlockT));
  p.M_total = o.N; p.T = T; p.K = o.F;
  p.num_m_groups = ceil_div(o.N, 128 * kC2MTiles); p.num_t_tiles = ceil_div(T, kC2BlockT); p.num_splits = 1;
  p.kblocks_total = ceil_div(o.F, bk); p.kblocks_per_split = p.kblocks_total;
  ProfScope ps(2, s);
  return launch_tc<kC2MTiles, kC2BlockT, bk, kSplit3, kEpi>(o.tmA, o.tmA_lo, tmR, tmRlo, p, s);
}

inline int launch_ratio(const float* X, int ldX, const float* WH, int ldWH, float* R, float* R_lo, int ldR, int T, int F,
                        float eps, int copy, cudaStream_t s) {
  ProfScope ps(1, s);
  dim3 g(T, ceil_div(ldR, 128));
  ratio_split_kernel<<<g, 128, 0, s>>>(X, ldX, WH, ldWH, R, R_lo, ldR, T, F, eps, copy);
  EVC_LAUNCH_CHECK();
  return EVC_OK;
}

inline int update_kl(DictOperands& o, int mode, const float* X, int ldX, int T, const float* WH, int ldWH, float* R,
                     int ldR, float* H, int ldH, const float* colsum, float lam, float eps,
                     const unsigned char* row_active, DevBuf* ws, cudaStream_t s) {
  const bool split3 = (mode == EVC_MODE_3XTF32);
  float* R_lo = split3 ? R + (size_t)T * ldR : nullptr;
  EVC_TRY(launch_ratio(X, ldX, WH, ldWH, R, R_lo, ldR, T, o.F, eps, 0, s));
  GemmParams p{};
  p.out = H; p.ld_out = ldH; p.out_lo = split3 ? ws->as<float>() : nullptr;
  p.colsum = colsum; p.lam = lam; p.eps = eps; p.row_active = row_active;
  if (split3) return contract2_t<true, TEPI_MU_KL>(o, T, R, R_lo, ldR, p, s);
  return contract2_t<false, TEPI_MU_KL>(o, T, R, R_lo, ldR, p, s);
}

inline int update_fro(DictOperands& o, int mode, int T, const float* WH, int ldWH, float* R, int ldR, float* H, int ldH,
                      const float* num0, float lam, float eps, const unsigned char* row_active, DevBuf* ws,
                      cudaStream_t s) {
  const bool split3 = (mode == EVC_MODE_3XTF32);
  float* R_lo = split3 ? R + (size_t)T * ldR : nullptr;
  EVC_TRY(launch_ratio(nullptr, 0, WH, ldWH, R, R_lo, ldR, T, o.F, eps, 1, s));
  GemmParams p{};
  p.out = H; p.ld_out = ldH; p.out_lo = split3 ? ws->as<float>() : nullptr;
  p.num0 = num0; p.lam = lam; p.eps = eps; p.row_active = row_active;
  if (split3) return contract2_t<true, TEPI_MU_FRO>(o, T, R, R_lo, ldR, p, s);
  return contract2_t<false, TEPI_MU_FRO>(o, T, R, R_lo, ldR, p, s);
}

// NUM0 (T, ldH) = X A^T : the second contraction with a plain store ([split=0][t][n] layout == (T, ldH)).
inline int frob_numerator(DictOperands& o, int mode, const float* X, int ldX, int T, float* R, int ldR, float* num0,
                          int ldH, DevBuf* ws, cudaStream_t s) {
  const bool split3 = (mode == EVC_MODE_3XTF32);
  float* R_lo = split3 ? R + (size_t)T * ldR : nullptr;
  // stage X into the zero-padded K-operand buffer (copy mode of the ratio kernel with WH := X)
  EVC_TRY(launch_ratio(nullptr, 0, X, ldX, R, R_lo, ldR, T, o.F, 0.f, 1, s));
  GemmParams p{};
  p.out = num0; p.ld_out = ldH;
  if (split3) return contract2_t<true, TEPI_PARTIAL>(o, T, R, R_lo, ldR, p, s);
  return contract2_t<false, TEPI_PARTIAL>(o, T, R, R_lo, ldR, p, s);
}

}  // namespace tc
}  // namespace evc
