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
// low 13 mantissa bits of a 32-bit operand, so the raw fp32 tile TMA brings IS the hi operand) and
// lo = x - hi (exact in fp32), written next to it in shared memory by split warps;
// D += M_lo*N_hi + M_hi*N_lo + M_hi*N_hi, dropping lo*lo (2^-22 relative).  HBM and L2 only ever carry one
// fp32 copy of each operand.
//
// kSplit3 == 2 ("cross16") keeps hi*hi on kind::tf32 and runs the two small cross terms on kind::f16 with bf16
// operands, M_lo16*N_hi16 + M_hi16*N_lo16 (hi16 = bf16(x), lo16 = bf16(x - trunc_tf32(x))): a bf16 MMA covers
// 16 K elements in the time a tf32 MMA covers 8, so a product costs 2 MMA passes instead of 3 and a third less
// operand traffic out of shared memory, for a per-term error of ~2^-20 (lo is <= 2^-10 |x| and is itself kept to
// 2^-9) on top of the dropped lo*lo.  The split warps write the two bf16 tiles (32-byte rows, 32B swizzle) where
// the fp32 lo tile used to be.
//
// BF16 fast mode (kBf16): the same kernel with tcgen05.mma.kind::f16 on bf16 operand copies -- the dictionary is
// converted once, the ratio is emitted in bf16 by the reduction pass, and the fused update writes a bf16 shadow of
// the activations next to the fp32 master copy (the multiplicative update itself stays fp32).  A K-block is still
// one 128-byte swizzle row, i.e. 64 bf16 elements, and one MMA still advances 32 bytes of K (16 elements), so the
// shared-memory layout, descriptors and barrier protocol are unchanged.
#pragma once
#include "evc_common.cuh"
#include "umma.cuh"
#include <cuda_bf16.h>
#include <algorithm>
#include <cstdlib>

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
  // The last m_group may hold fewer 128-row sub-tiles than the others (F = 513 -> 2 + 2 + 1); it then gets
  // fewer, longer K splits so every CTA carries the same number of MMAs.  items_main = work items of the
  // other groups; 0 splits_last means "no special last group".
  int items_main, splits_last, kblocks_per_split_last;
  int direct_store;  // fused update: write H straight from registers (1) instead of shared memory + TMA store (0)
  // Tail balancing of contraction 2: tiles with index >= half_from are processed as two half-width items (t_cols =
  // kBlockT/2 frames each) so the last, partly filled round costs half a tile time.  half_from = items_main: off.
  int half_from;
  int m_fastest;  // work-item order: 1 = consecutive CTAs take consecutive dictionary-row groups of one frame tile
  float* out;    // PARTIAL: [split][t][m] with pitch ld_out;  MU_*: the activations H (T, ld_out)
  int ld_out;
  __nv_bfloat16* out16;  // BF16 mode, MU_*: bf16 shadow of H (T, ld_out16), the K operand of the next contraction 1
  int ld_out16;
  const float* colsum;
  const float* num0;  // MU_FRO: cached numerator X A^T, same pitch as H
  float lam, eps;
  const unsigned char* row_active;
  // Dictionary rows F_main..F-1 that contraction 1 does not run through the tensor cores (F = 513 = 4*128 + 1):
  // the fused update accumulates their share of the NEXT A*H, sum_n h_new[t,n] * A[n, F_main+l], per warp.
  const float* left_a;  // (n_left, left_lda): rows F_main.. of the transposed dictionary (contiguous in n)
  int left_lda, n_left;
  float* left_out;      // [l][t][row] partial sums, row = 4*(128-exemplar block) + lane quarter
  int left_ld, left_rows;  // frames pitch, rows pitch
  long long* dbg_cycles;  // EVC_DEBUG_TIMING: per-CTA [8] cycle counters of the role threads (nullptr = off)
  int debug_flags;  // timing experiments only (EVC_DEBUG_FLAGS): 1 skip split, 2 skip MMA, 4 skip TMA, 8 skip epilogue memory ops, 16 enable the L2 look-ahead prefetch,
                    // 32 / 64 skip the frame-side / dictionary-side operand loads, 128 skip the H chunk loads and stores of the fused update
};

// vals[j] (j = 0..31) per lane -> returns, in lane L, the sum over all 32 lanes of vals[L]  (31 shuffles).
__device__ __forceinline__ float warp_transpose_sum(float (&vals)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? vals[i] : vals[i + off];
      const float keep = up ? vals[i + off] : vals[i];
      vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return vals[0];
}

struct WorkItem {
  int m_group, t_tile, split, kb0, kb1;
  int t_off, t_cols;  // frame offset inside the frame tile and number of frame columns of this item
};
__host__ __device__ __forceinline__ WorkItem decode_item(const GemmParams& p, int item, int block_t) {
  WorkItem w;
  w.t_off = 0;
  w.t_cols = block_t;
  if (item >= p.half_from && p.half_from < p.items_main) {
    // half-width tail items of contraction 2 (no split-K there): tile = half_from + h/2, half = h & 1
    const int h = item - p.half_from;
    item = p.half_from + (h >> 1);
    w.t_cols = block_t / 2;
    w.t_off = (h & 1) * w.t_cols;
  }
  if (item < p.items_main) {
    const int groups = p.splits_last ? p.num_m_groups - 1 : p.num_m_groups;
    if (p.m_fastest) {
      w.m_group = item % groups;
      const int rest = item / groups;
      w.t_tile = rest % p.num_t_tiles;
      w.split = rest / p.num_t_tiles;
    } else {
      w.t_tile = item % p.num_t_tiles;
      const int rest = item / p.num_t_tiles;
      w.m_group = rest % groups;
      w.split = rest / groups;
    }
    w.kb0 = w.split * p.kblocks_per_split;
    w.kb1 = min(w.kb0 + p.kblocks_per_split, p.kblocks_total);
  } else {
    const int it = item - p.items_main;
    w.t_tile = it % p.num_t_tiles;
    w.split = it / p.num_t_tiles;
    w.m_group = p.num_m_groups - 1;
    w.kb0 = w.split * p.kblocks_per_split_last;
    w.kb1 = min(w.kb0 + p.kblocks_per_split_last, p.kblocks_total);
  }
  return w;
}

__device__ __forceinline__ float tf32_lo(float x) {
  return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
}

constexpr int kEpiWarps = 8;    // two warps per TMEM lane quarter, interleaved over 32-column chunks
constexpr int kXformWarps = 8;  // dedicated hi/lo split warps (only when the epilogue overlaps the main loop)
#ifndef EVC_WARPS_PER_STAGE
#define EVC_WARPS_PER_STAGE 2
#endif
constexpr int kWarpsPerStage = EVC_WARPS_PER_STAGE;  // split warps that share one ring stage (groups take K-blocks round-robin)
constexpr int kSmemBudget = 227 * 1024 - 2048;
// back-off (ns) between polls of the many-warp waits; 0 = spin
#ifndef EVC_SPLIT_SLEEP_NS
#define EVC_SPLIT_SLEEP_NS 0
#endif
// MMA-issuing thread, per K-block (A/B knobs, both measured and rejected: 113.3 vs 109.9 us for contraction 2):
// EVC_KLOOP_FENCE 0 drops the tcgen05.fence::after_thread_sync after every stage wait, EVC_KLOOP_PEEK 1 polls the
// NEXT stage's barrier before issuing this stage's MMAs.
// EVC_EARLY_HIHI 1 (cross16 only): the tf32 hi*hi MMAs of a K-block are issued as soon as its TMA bytes have
// landed in both CTAs ("landed" barrier), the bf16 cross terms when the split warps are done ("ready"), so the
// split overlaps tensor work of the same stage instead of preceding it.
#ifndef EVC_EARLY_HIHI
#define EVC_EARLY_HIHI 0
#endif
#ifndef EVC_KLOOP_FENCE
#define EVC_KLOOP_FENCE 1
#endif
#ifndef EVC_KLOOP_PEEK
#define EVC_KLOOP_PEEK 0
#endif
#ifndef EVC_EPI_SLEEP_NS
#define EVC_EPI_SLEEP_NS 0
#endif

// The fused-update epilogue stages H through shared memory in [32 frames x 128 exemplars] chunks moved by TMA
// (loads prefetched by a loader warp, stores issued by a storer warp): per-lane 128-byte global accesses
// from the epilogue warps were limited by the SM's outstanding-miss capacity, bulk copies are not.
#ifndef EVC_H_BUFS
#define EVC_H_BUFS 4
#endif
constexpr int kHChunkT = 32, kHBufBytes = kHChunkT * 128 * 4, kHBufs = EVC_H_BUFS;

// kCG = CTA group size of the MMA.  1: one SM per tile.  2: a CTA pair (cluster of 2) shares each tile --
// tcgen05.mma.cta_group::2 with M = 256 (128 dictionary rows per CTA) and the frame (N) operand split in halves
// between the two CTAs' shared memories, which halves the per-SM shared-memory reads of the frame operand
// (the 3xTF32 main loop of the 1-CTA version is shared-memory-bandwidth bound).
template <int kMTiles, int kBlockT, int kBlockK, int kSplit3, bool kStageH, int kCG, bool kStageQ = false>
struct TileCfg {
  static constexpr int kRowBytes = kBlockK * 4;
  static constexpr int kMTileBytes = 128 * kRowBytes;              // this CTA's 128 rows of one dictionary sub-tile
  static constexpr int kNRows = kBlockT / kCG;                     // frame rows this CTA loads
  static constexpr int kNTileBytes = kNRows * kRowBytes;
  static constexpr int kCopies = kSplit3 ? 2 : 1;
  static constexpr int kMBytes = kMTiles * kMTileBytes;            // hi tiles of the dictionary operand
  static constexpr int kLoadBytes = kMBytes + kNTileBytes;         // what TMA brings per stage (hi only)
  static constexpr int kStageBytes = kCopies * kLoadBytes;         // + the lo tiles computed in place
  static constexpr int kOffMlo = kMBytes;
  static constexpr int kOffN = kCopies * kMBytes;
  static constexpr int kOffNlo = kOffN + kNTileBytes;
  // Frobenius also stages the cached numerator X A^T next to H: half as many, twice as large buffers
  static constexpr int kHBufsUsed = kStageQ ? kHBufs / 2 : kHBufs;
  static constexpr int kHBufStride = kStageQ ? 2 * kHBufBytes : kHBufBytes;
  static constexpr int kHBytes = kStageH ? kHBufs * kHBufBytes : 0;
  static constexpr int kStagesRaw = (kSmemBudget - kHBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kAccCols = kMTiles * kBlockT;
  static constexpr int kAccStages = (512 / kAccCols) >= 2 ? 2 : 1;
  // With two accumulator stages the epilogue runs concurrently with the next tile's main loop, so the
  // hi/lo split needs its own warps; with one stage the (idle) epilogue warps do it.
  static constexpr bool kDedicatedXform = kSplit3 && kAccStages == 2;
  static constexpr int kXformThreads = kSplit3 ? (kDedicatedXform ? kXformWarps * 32 : kEpiWarps * 32) : 0;
  // A pair without split warps still needs someone to tell the leader that the peer's TMA bytes landed.
  static constexpr bool kRelay = (kCG == 2) && !kSplit3;
  // arrivals per CTA on the leader's "stage ready" barrier
  static constexpr int kSplitWarps = kXformThreads > 0 ? kXformThreads / 32 : 1;
  static constexpr int kSplitGroups = kSplit3 ? kSplitWarps / kWarpsPerStage : 1;  // groups take K-blocks round-robin
  static constexpr int kReadyArrivals = kSplit3 ? kWarpsPerStage : 1;  // the warps of the group that owns the stage
  static constexpr int kFirstXformWarp = 2 + kEpiWarps;
  static constexpr int kRelayWarp = kFirstXformWarp + (kDedicatedXform ? kXformWarps : 0);
  static constexpr int kLoaderWarp = kRelayWarp + (kRelay ? 1 : 0);  // H chunk loader, then storer
  static constexpr int kThreads = (kLoaderWarp + (kStageH ? 2 : 0)) * 32;
  static constexpr int kOffH = kStages * kStageBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + kHBytes + 1024;  // + slack to align the ring to 1024 B
  static constexpr int kRowsPerSub = 128 * kCG;                   // dictionary rows of one MMA (M = 128 or 256)
  static_assert(kRowBytes == 64 || kRowBytes == 128, "K block must be one 64B or 128B swizzle row");
  static_assert(kStages >= 2, "tile does not fit twice in shared memory");
  static_assert(kAccCols <= 512, "accumulators exceed TMEM");
  static_assert(kBlockT % 32 == 0 && kBlockT >= 32 && kBlockT <= 256, "bad frame tile");
  static_assert(!kStageH || kMTiles == 1, "H staging assumes one sub-tile per work item");
  static_assert(kCG == 1 || kCG == 2, "CTA group is 1 or 2");
  static_assert(kSplit3 != 2 || kRowBytes == 64, "the bf16 cross-term tiles are derived from 64-byte rows");
};

// lo = x - trunc_tf32(x) for one ring stage: element-wise on raw bytes, so the swizzled layout TMA wrote
// carries over unchanged to the lo tiles.  ONE warp owns a whole stage (the split warps take K-blocks
// round-robin): the per-stage cost is dominated by fixed latencies (barrier wake-up, proxy fence, arrive),
// so several stages are split concurrently instead of all warps sharing one.  16 B per lane per access,
// batches of 8 loads in flight.
template <class Cfg>
__device__ __forceinline__ void split_region(const uint8_t* hi, uint8_t* lo, int bytes, int lane, int part) {
  // this warp's share: a contiguous 1/kWarpsPerStage of the region
  const int share = bytes / kWarpsPerStage;
  const float4* src = reinterpret_cast<const float4*>(hi + part * share) + lane;
  float4* dst = reinterpret_cast<float4*>(lo + part * share) + lane;
  const int n = share / 16 / 32;  // float4 per lane (a multiple of 4: tiles are >= 4 KB)
  for (int q0 = 0; q0 < n; q0 += 4) {
    float4 x[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) x[q] = src[(q0 + q) * 32];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      dst[(q0 + q) * 32] = make_float4(tf32_lo(x[q].x), tf32_lo(x[q].y), tf32_lo(x[q].z), tf32_lo(x[q].w));
  }
}
// cross16: from one fp32 tile (64-byte rows, 64B swizzle: 16-byte chunk c of row r sits at chunk c ^ ((r >> 1) & 3))
// derive hi16 = bf16(x) and lo16 = bf16(x - trunc_tf32(x)) as tiles with 32-byte rows in the 32B-swizzle layout
// (chunk c of row r at chunk c ^ ((r >> 2) & 1)).  A unit is half a row: 8 floats in, 16 + 16 bytes out.
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
template <int kRows>
__device__ __forceinline__ void split16_tile(const uint8_t* src, uint8_t* hi16, uint8_t* lo16, int lane, int part) {
  constexpr int kShare = kRows * 2 / kWarpsPerStage, kPerLane = kShare / 32;
  constexpr int kBatch = kPerLane < 4 ? kPerLane : 4;  // units in flight per lane
  static_assert(kPerLane >= 1 && kPerLane % kBatch == 0, "tile too small for the split warps");
#pragma unroll
  for (int b = 0; b < kPerLane; b += kBatch) {
    float4 x[kBatch][2];
#pragma unroll
    for (int q = 0; q < kBatch; ++q) {
      const int u = part * kShare + (b + q) * 32 + lane, row = u >> 1, h = u & 1, sw = (row >> 1) & 3;
      const uint8_t* r = src + row * 64;
      x[q][0] = *reinterpret_cast<const float4*>(r + (((2 * h) ^ sw) << 4));
      x[q][1] = *reinterpret_cast<const float4*>(r + (((2 * h + 1) ^ sw) << 4));
    }
#pragma unroll
    for (int q = 0; q < kBatch; ++q) {
      const int u = part * kShare + (b + q) * 32 + lane, row = u >> 1, h = u & 1;
      const int off = row * 32 + ((h ^ ((row >> 2) & 1)) << 4);
      const float4 a = x[q][0], c = x[q][1];
      *reinterpret_cast<uint4*>(hi16 + off) =
          make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(c.x, c.y), pack_bf16(c.z, c.w));
      *reinterpret_cast<uint4*>(lo16 + off) =
          make_uint4(pack_bf16(tf32_lo(a.x), tf32_lo(a.y)), pack_bf16(tf32_lo(a.z), tf32_lo(a.w)),
                     pack_bf16(tf32_lo(c.x), tf32_lo(c.y)), pack_bf16(tf32_lo(c.z), tf32_lo(c.w)));
    }
  }
}
template <class Cfg, int kMTiles>
__device__ __forceinline__ void split16_stage(uint8_t* stage, int lane, int part) {
  // derived region of dictionary sub-tile i: [hi16 | lo16] where the fp32 lo tile of the classic split would be
#pragma unroll
  for (int i = 0; i < kMTiles; ++i) {
    uint8_t* d = stage + Cfg::kOffMlo + i * Cfg::kMTileBytes;
    split16_tile<128>(stage + i * Cfg::kMTileBytes, d, d + Cfg::kMTileBytes / 2, lane, part);
  }
  uint8_t* d = stage + Cfg::kOffNlo;
  split16_tile<Cfg::kNRows>(stage + Cfg::kOffN, d, d + Cfg::kNTileBytes / 2, lane, part);
}

template <class Cfg>
__device__ __forceinline__ void split_stage(uint8_t* stage, int lane, int part) {
  static_assert(Cfg::kMBytes % (2048 * kWarpsPerStage) == 0 && Cfg::kNTileBytes % (2048 * kWarpsPerStage) == 0,
                "every split warp must get a multiple of 2 KB per region");
  split_region<Cfg>(stage, stage + Cfg::kOffMlo, Cfg::kMBytes, lane, part);
  split_region<Cfg>(stage + Cfg::kOffN, stage + Cfg::kOffNlo, Cfg::kNTileBytes, lane, part);
}

template <int kMTiles, int kBlockT, int kBlockK, int kSplit3, int kEpi, int kCG, bool kBf16 = false>
__global__ void __launch_bounds__(
    (TileCfg<kMTiles, kBlockT, kBlockK, kSplit3, kEpi == TEPI_MU_KL || kEpi == TEPI_MU_FRO, kCG, kEpi == TEPI_MU_FRO>::kThreads), 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmN,
               const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmQ, const GemmParams p) {
  constexpr bool kFro = (kEpi == TEPI_MU_FRO);
  constexpr bool kStageH = (kEpi == TEPI_MU_KL) || kFro;
  using Cfg = TileCfg<kMTiles, kBlockT, kBlockK, kSplit3, kStageH, kCG, kFro>;
  constexpr int kHB = Cfg::kHBufsUsed;  // chunk buffers in use
  constexpr int kStages = Cfg::kStages;
  constexpr int kAccStages = Cfg::kAccStages;
  static_assert(!(kBf16 && kSplit3), "the hi/lo split belongs to the TF32 path");
  constexpr uint32_t kFmt = kBf16 ? kFmtBF16 : kFmtTF32;
  constexpr int kKE = kBf16 ? 2 * kBlockK : kBlockK;  // K elements per K-block (one swizzle row)
  constexpr int kKStep = kBf16 ? 16 : 8;              // K elements per MMA (32 bytes either way)
  constexpr uint32_t kIdesc = make_idesc(kFmt, 128 * kCG, kBlockT);
  constexpr bool kCross16 = (kSplit3 == 2);
  constexpr bool kEarly = kCross16 && (EVC_EARLY_HIHI != 0);
  // does the MMA warp wait on the "ready" barrier (split and/or pair) or directly on the TMA barrier?
  constexpr bool kUseReady = kSplit3 || kCG == 2;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kStages];   // this CTA's TMA bytes landed
  __shared__ __align__(8) uint64_t bar_ready[kStages];  // (leader's copy is used) stage usable by the MMA: lo tiles
                                                        // written / both CTAs of the pair loaded
  __shared__ __align__(8) uint64_t bar_landed[kStages];  // (leader's copy is used) the TMA bytes of the stage landed in
                                                         // every CTA of the pair (EVC_EARLY_HIHI)
  __shared__ __align__(8) uint64_t bar_empty[kStages];  // MMAs that read the stage retired (both CTAs' copies fire)
  __shared__ __align__(8) uint64_t bar_acc_full[kAccStages];
  __shared__ __align__(8) uint64_t bar_acc_empty[kAccStages];  // (leader's copy is used)
  __shared__ __align__(8) uint64_t bar_hfull[kHBufs];   // H chunk landed in shared memory
  __shared__ __align__(8) uint64_t bar_hready[kHBufs];  // the 4 epilogue warps of a chunk wrote the updated values
  __shared__ __align__(8) uint64_t bar_hempty[kHBufs];  // the TMA store has read the buffer
  __shared__ uint32_t tmem_base_smem;
  __shared__ long long dbg_t_issue[kStages];  // EVC_DEBUG_TIMING: clock at TMA issue per stage

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // cycle counter of the EVC_DEBUG_TIMING instrumentation: not read at all in normal runs (a CS2R per pipeline step
  // of the single-thread roles is not free)
  const bool dbg_on = p.dbg_cycles != nullptr;
  auto clk = [dbg_on]() -> long long { return dbg_on ? clock64() : 0ll; };
  const long long t_entry = clk();
  const uint32_t rank = (kCG == 2) ? cluster_ctarank() : 0u;  // 0 = leader (issues the MMAs)
  uint8_t* ring_ptr = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t ring = smem_u32(ring_ptr);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmM);
    tma_prefetch_desc(&tmN);
    if (kStageH) tma_prefetch_desc(&tmH);
    if (kFro) tma_prefetch_desc(&tmQ);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(smem_u32(&bar_full[i]), 1);
      mbar_init(smem_u32(&bar_ready[i]), Cfg::kReadyArrivals * kCG);
      mbar_init(smem_u32(&bar_landed[i]), kCG);
      mbar_init(smem_u32(&bar_empty[i]), 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(smem_u32(&bar_acc_full[i]), 1);
      mbar_init(smem_u32(&bar_acc_empty[i]), kEpiWarps * kCG);  // one elected lane of each epilogue warp
    }
    for (int i = 0; i < kHBufs; ++i) {
      mbar_init(smem_u32(&bar_hfull[i]), 1);
      mbar_init(smem_u32(&bar_hready[i]), 4);
      mbar_init(smem_u32(&bar_hempty[i]), p.direct_store ? 4 : 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kCG == 2) { tmem_alloc2(smem_u32(&tmem_base_smem), 512); tmem_relinquish2(); }
    else { tmem_alloc(smem_u32(&tmem_base_smem), 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if (kCG == 2) cluster_sync_all(); else __syncthreads();  // peer's barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // everything above (barrier init, TMEM allocation, descriptor prefetch) may overlap the tail of the previous
  // kernel in the stream; from here on its results are needed
  const long long t_setup = clk();
  pdl_wait();
  pdl_launch_dependents();
  const long long t_go = clk();

  const int num_items = p.items_main + p.splits_last * p.num_t_tiles + (p.items_main - p.half_from);
  const int first_item = blockIdx.x / kCG, item_stride = gridDim.x / kCG;  // both CTAs of a pair walk the same items

  // "stage ready" arrive: local barrier for a single CTA, the leader's copy for a pair
  auto ready_arrive = [&](int stage) {
    if (kCG == 2) mbar_arrive_cluster(mapa_rank(smem_u32(&bar_ready[stage]), 0));
    else mbar_arrive(smem_u32(&bar_ready[stage]));
  };

  auto landed_arrive = [&](int stage) {
    if (kCG == 2) mbar_arrive_cluster(mapa_rank(smem_u32(&bar_landed[stage]), 0));
    else mbar_arrive(smem_u32(&bar_landed[stage]));
  };

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long c_empty = 0;
      const long long c_start = clk();
      // L2 look-ahead: the tiles of K-block (current + kAhead) are prefetched into L2 when the current one is
      // loaded, so the ring refill sees L2 latency instead of HBM latency (the ring itself is only 3-6 deep).
      constexpr int kAhead = 2 * kStages;
      int la_item = first_item, la_kb = 0, la_kb1 = 0, la_m0 = 0, la_t0 = 0, la_count = 0, issued = 0;
      bool la_valid = la_item < num_items;
      if (la_valid) {
        const WorkItem w = decode_item(p, la_item, kBlockT);
        la_kb = w.kb0; la_kb1 = w.kb1;
        la_m0 = w.m_group * (Cfg::kRowsPerSub * kMTiles) + (int)rank * 128;
        la_t0 = w.t_tile * kBlockT + w.t_off + (int)rank * (w.t_cols / kCG);
      }
      // measured: no gain (the ring refill is not HBM-latency bound) and the extra issue slots slow the
      // single producer thread down, so the look-ahead is off unless debug flag 16 asks for it
      const bool use_la = (p.debug_flags & 16) && !(p.debug_flags & 4);
      for (int item = first_item; item < num_items; item += item_stride) {
        const WorkItem w = decode_item(p, item, kBlockT);
        const int m0 = w.m_group * (Cfg::kRowsPerSub * kMTiles) + (int)rank * 128;
        // (a half-width item still loads a kNRows-row box: the narrower MMA never reads the surplus rows)
        const int t0 = w.t_tile * kBlockT + w.t_off + (int)rank * (w.t_cols / kCG);
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          while (use_la && la_valid && la_count < issued + kAhead) {
            if (la_count >= issued + kStages) {  // (the first kStages blocks are about to be loaded anyway)
              const int kc = la_kb * kKE;
#pragma unroll
              for (int i = 0; i < kMTiles; ++i)
                if (la_m0 + i * Cfg::kRowsPerSub < p.M_total) tma_prefetch_l2_2d(&tmM, kc, la_m0 + i * Cfg::kRowsPerSub);
              if (la_t0 < p.T) tma_prefetch_l2_2d(&tmN, kc, la_t0);
            }
            ++la_count;
            if (++la_kb >= la_kb1) {
              la_item += item_stride;
              la_valid = la_item < num_items;
              if (la_valid) {
                const WorkItem w2 = decode_item(p, la_item, kBlockT);
                la_kb = w2.kb0; la_kb1 = w2.kb1;
                la_m0 = w2.m_group * (Cfg::kRowsPerSub * kMTiles) + (int)rank * 128;
                la_t0 = w2.t_tile * kBlockT + w2.t_off + (int)rank * (w2.t_cols / kCG);
              }
            }
          }
          ++issued;
          const long long c0 = clk();
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          c_empty += clk() - c0;
          const uint32_t full = smem_u32(&bar_full[stage]);
          if (p.debug_flags & 4) {
            mbar_arrive(full);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
            continue;
          }
          const bool skip_m = (p.debug_flags & 64) != 0, skip_n = (p.debug_flags & 32) != 0;  // timing experiments
          mbar_arrive_expect_tx(full, (uint32_t)((skip_m ? 0 : Cfg::kMBytes) + (skip_n ? 0 : Cfg::kNTileBytes)));
          if (p.dbg_cycles) dbg_t_issue[stage] = clk();
          const uint32_t sbase = ring + stage * Cfg::kStageBytes;
          const int kc = kb * kKE;
          // (sub-tiles past M_total are still loaded: TMA zero-fills them and the byte count stays constant)
#pragma unroll
          for (int i = 0; i < kMTiles; ++i)
            if (!skip_m) tma_load_2d(sbase + i * Cfg::kMTileBytes, &tmM, kc, m0 + i * Cfg::kRowsPerSub, full, kEvictNormal);
          if (!skip_n) tma_load_2d(sbase + Cfg::kOffN, &tmN, kc, t0, full, kEvictNormal);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
      if (p.dbg_cycles) {
        long long* o = p.dbg_cycles + (size_t)blockIdx.x * 8;
        o[3] = clk() - c_start; o[4] = c_empty;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA of a pair only) =================
    if (lane == 0 && rank == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      long long c_acc = 0, c_ready = 0, c_issue = 0, c_commit = 0;
      const long long c_start = clk();
      for (int item = first_item; item < num_items; item += item_stride) {
        const WorkItem w = decode_item(p, item, kBlockT);
        const int m0 = w.m_group * (Cfg::kRowsPerSub * kMTiles);
        const int kb0 = w.kb0, kb1 = w.kb1;
        const uint32_t idesc = (w.t_cols == kBlockT) ? kIdesc : make_idesc(kFmt, 128 * kCG, (uint32_t)w.t_cols);
        const uint32_t idesc16 = make_idesc(kFmtBF16, 128 * kCG, (uint32_t)w.t_cols);  // cross16: the bf16 cross terms
        long long c0 = clk();
        if (kCG == 2) mbar_wait_cluster(smem_u32(&bar_acc_empty[acc]), acc_phase ^ 1u);
        else mbar_wait(smem_u32(&bar_acc_empty[acc]), acc_phase ^ 1u);
        c_acc += clk() - c0;
        tc_fence_after();
        bool peeked = false;  // the current stage's barrier was already seen complete
        for (int kb = kb0; kb < kb1; ++kb) {
          // 3xTF32 / pairs: the ready barrier fires after the TMA barrier(s) and after the lo tiles are visible
          c0 = clk();
          if (kEarly) {
            if (kCG == 2) mbar_wait_cluster(smem_u32(&bar_landed[stage]), phase);
            else mbar_wait(smem_u32(&bar_landed[stage]), phase);
          } else if (!peeked) {
            if (kCG == 2) mbar_wait_cluster(smem_u32(&bar_ready[stage]), phase);
            else mbar_wait(smem_u32(kUseReady ? &bar_ready[stage] : &bar_full[stage]), phase);
          }
          if (EVC_KLOOP_PEEK && kb + 1 < kb1) {
            const int ns = (stage + 1 == kStages) ? 0 : stage + 1;
            const uint32_t np = (ns == 0) ? (phase ^ 1u) : phase;
            peeked = mbar_try_wait(smem_u32((kUseReady || kCG == 2) ? &bar_ready[ns] : &bar_full[ns]), np);
          } else {
            peeked = false;
          }
          c_ready += clk() - c0;
          if (EVC_KLOOP_FENCE) tc_fence_after();
          c0 = clk();
          const uint32_t sbase = ring + stage * Cfg::kStageBytes;
          const uint32_t nbase = sbase + Cfg::kOffN;
          const int kvalid = min(kKE, p.K - kb * kKE);
          const int ksteps = (kvalid + kKStep - 1) / kKStep;
#pragma unroll
          for (int i = 0; i < kMTiles; ++i) {
            if (m0 + i * Cfg::kRowsPerSub >= p.M_total) break;  // pure padding: no MMAs, the epilogue skips it too
            if (p.debug_flags & 2) break;
            const uint32_t d = tmem_base + (uint32_t)((acc * kMTiles + i) * kBlockT);
            if (kEarly) {
              // hi*hi now (only needs the TMA bytes); the cross terms of all sub-tiles follow below, after "ready"
              for (int ks = 0; ks < ksteps; ++ks) {
                const uint64_t a_hi = make_smem_desc(sbase + i * Cfg::kMTileBytes + ks * 32, Cfg::kRowBytes);
                const uint64_t b_hi = make_smem_desc(nbase + ks * 32, Cfg::kRowBytes);
                mma_issue<false, kCG>(d, a_hi, b_hi, idesc, (kb > kb0 || ks > 0) ? 1u : 0u);
              }
            } else if (kCross16) {
              // small terms first: lo16*hi16 + hi16*lo16 over the whole K-block (one 16-element bf16 MMA each;
              // a K tail was zero-filled by TMA), then hi*hi in tf32
              const uint32_t a16 = sbase + Cfg::kOffMlo + i * Cfg::kMTileBytes, b16 = nbase + Cfg::kNTileBytes;
              const uint64_t a_hi16 = make_smem_desc(a16, 32), a_lo16 = make_smem_desc(a16 + Cfg::kMTileBytes / 2, 32);
              const uint64_t b_hi16 = make_smem_desc(b16, 32), b_lo16 = make_smem_desc(b16 + Cfg::kNTileBytes / 2, 32);
              mma_issue<true, kCG>(d, a_lo16, b_hi16, idesc16, kb > kb0 ? 1u : 0u);
              mma_issue<true, kCG>(d, a_hi16, b_lo16, idesc16, 1u);
              for (int ks = 0; ks < ksteps; ++ks) {
                const uint64_t a_hi = make_smem_desc(sbase + i * Cfg::kMTileBytes + ks * 32, Cfg::kRowBytes);
                const uint64_t b_hi = make_smem_desc(nbase + ks * 32, Cfg::kRowBytes);
                mma_issue<false, kCG>(d, a_hi, b_hi, idesc, 1u);
              }
            } else
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint32_t accum = (kb > kb0 || ks > 0) ? 1u : 0u;
              const uint64_t a_hi = make_smem_desc(sbase + i * Cfg::kMTileBytes + ks * 32, Cfg::kRowBytes);
              const uint64_t b_hi = make_smem_desc(nbase + ks * 32, Cfg::kRowBytes);
              if (kSplit3) {
                const uint64_t a_lo =
                    make_smem_desc(sbase + Cfg::kOffMlo + i * Cfg::kMTileBytes + ks * 32, Cfg::kRowBytes);
                const uint64_t b_lo = make_smem_desc(nbase + Cfg::kNTileBytes + ks * 32, Cfg::kRowBytes);
                if (kCG == 2) {
                  mma_tf32_2cta(d, a_lo, b_hi, idesc, accum);
                  mma_tf32_2cta(d, a_hi, b_lo, idesc, 1u);
                  mma_tf32_2cta(d, a_hi, b_hi, idesc, 1u);
                } else {
                  mma_tf32(d, a_lo, b_hi, idesc, accum);
                  mma_tf32(d, a_hi, b_lo, idesc, 1u);
                  mma_tf32(d, a_hi, b_hi, idesc, 1u);
                }
              } else {
                mma_issue<kBf16, kCG>(d, a_hi, b_hi, idesc, accum);
              }
            }
          }
          if (kEarly) {
            const long long cw = clk();
            if (kCG == 2) mbar_wait_cluster(smem_u32(&bar_ready[stage]), phase);
            else mbar_wait(smem_u32(&bar_ready[stage]), phase);
            c_ready += clk() - cw;
            tc_fence_after();
#pragma unroll
            for (int i = 0; i < kMTiles; ++i) {
              if (m0 + i * Cfg::kRowsPerSub >= p.M_total) break;
              if (p.debug_flags & 2) break;
              const uint32_t d = tmem_base + (uint32_t)((acc * kMTiles + i) * kBlockT);
              const uint32_t a16 = sbase + Cfg::kOffMlo + i * Cfg::kMTileBytes, b16 = nbase + Cfg::kNTileBytes;
              mma_issue<true, kCG>(d, make_smem_desc(a16 + Cfg::kMTileBytes / 2, 32), make_smem_desc(b16, 32), idesc16, 1u);
              mma_issue<true, kCG>(d, make_smem_desc(a16, 32), make_smem_desc(b16 + Cfg::kNTileBytes / 2, 32), idesc16, 1u);
            }
          }
          // frees the smem slot (in both CTAs of a pair) when these MMAs retire
          const long long c1 = clk();
          c_issue += c1 - c0;
          if (kCG == 2) mma_commit_2cta(smem_u32(&bar_empty[stage])); else mma_commit(smem_u32(&bar_empty[stage]));
          c_commit += clk() - c1;
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        // accumulator complete -> epilogue(s)
        if (kCG == 2) mma_commit_2cta(smem_u32(&bar_acc_full[acc])); else mma_commit(smem_u32(&bar_acc_full[acc]));
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
      }
      if (p.dbg_cycles) {
        long long* o = p.dbg_cycles + (size_t)blockIdx.x * 8;
        o[0] = clk() - c_start; o[1] = c_acc; o[2] = c_ready;
        long long* o2 = p.dbg_cycles + (size_t)(gridDim.x + blockIdx.x) * 8;
        o2[3] = c_issue; o2[4] = c_commit;
      }
    }
  } else if (kStageH && warp == Cfg::kLoaderWarp) {
    // ================= H chunk loader: prefetches the activations the epilogue will update =================
    if (lane == 0) {
      int hbase = 0;
      for (int item = first_item; item < num_items; item += item_stride) {
        const WorkItem w = decode_item(p, item, kBlockT);
        const int t0 = w.t_tile * kBlockT + w.t_off, n0 = w.m_group * Cfg::kRowsPerSub + (int)rank * 128;
        const int nch = max(0, min(w.t_cols / kHChunkT, (p.T - t0 + kHChunkT - 1) / kHChunkT));
        for (int c = 0; c < nch; ++c) {
          const int seq = hbase + c, b = seq % kHB;
          const uint32_t ph = (uint32_t)(seq / kHB) & 1u;
          mbar_wait(smem_u32(&bar_hempty[b]), ph ^ 1u);
          const uint32_t full = smem_u32(&bar_hfull[b]);
          if (p.debug_flags & 128) { mbar_arrive(full); continue; }  // timing experiments: no H stream
          mbar_arrive_expect_tx(full, (uint32_t)Cfg::kHBufStride);
          tma_load_2d(ring + Cfg::kOffH + b * Cfg::kHBufStride, &tmH, n0, t0 + c * kHChunkT, full, kEvictFirst);
          if (kFro)
            tma_load_2d(ring + Cfg::kOffH + b * Cfg::kHBufStride + kHBufBytes, &tmQ, n0, t0 + c * kHChunkT, full, kEvictFirst);
        }
        hbase += nch;
      }
    }
  } else if (kStageH && warp == Cfg::kLoaderWarp + 1) {
    // ================= H chunk storer =================
    if (lane == 0 && !p.direct_store) {
      int hbase = 0;
      for (int item = first_item; item < num_items; item += item_stride) {
        const WorkItem w = decode_item(p, item, kBlockT);
        const int t0 = w.t_tile * kBlockT + w.t_off, n0 = w.m_group * Cfg::kRowsPerSub + (int)rank * 128;
        const int nch = max(0, min(w.t_cols / kHChunkT, (p.T - t0 + kHChunkT - 1) / kHChunkT));
        for (int c = 0; c < nch; ++c) {
          const int seq = hbase + c, b = seq % kHB;
          const uint32_t ph = (uint32_t)(seq / kHB) & 1u;
          mbar_wait(smem_u32(&bar_hready[b]), ph);
          if (!(p.debug_flags & 128)) {
            tma_store_2d(&tmH, n0, t0 + c * kHChunkT, ring + Cfg::kOffH + b * Cfg::kHBufStride);
            tma_store_commit();
            tma_store_wait_read();
          }
          mbar_arrive(smem_u32(&bar_hempty[b]));
        }
        hbase += nch;
      }
      tma_store_wait_all();
    }
  } else if (Cfg::kRelay && warp == Cfg::kRelayWarp) {
    // ================= pair without split warps: tell the leader when this CTA's stage has landed =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = first_item; item < num_items; item += item_stride) {
        const WorkItem w = decode_item(p, item, kBlockT);
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          ready_arrive(stage);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= Cfg::kFirstXformWarp) {
    // ================= dedicated hi/lo split warps (3xTF32 with an overlapped epilogue) =================
    if (Cfg::kDedicatedXform) {
      const int me = warp - Cfg::kFirstXformWarp;  // this warp owns K-blocks me, me + kSplitWarps, ...
      int stage = 0, seq = 0;
      uint32_t phase = 0;
      long long d_tma = 0, d_split = 0, d_n = 0;
      for (int item = first_item; item < num_items; item += item_stride) {
        const WorkItem w = decode_item(p, item, kBlockT);
        for (int kb = w.kb0; kb < w.kb1; ++kb, ++seq) {
          // every warp observes every phase of the barrier (a waiter that skipped phases could be fooled by
          // parity aliasing two ring passes later); only the owner of the K-block does the work
          mbar_wait_backoff(smem_u32(&bar_full[stage]), phase, EVC_SPLIT_SLEEP_NS);
          if (kEarly && me == 0 && lane == 0) landed_arrive(stage);
          if (seq % Cfg::kSplitGroups == me / kWarpsPerStage) {
            const long long t1 = clk();
            if (p.dbg_cycles && lane == 0) { d_tma += t1 - dbg_t_issue[stage]; ++d_n; }
            if (!(p.debug_flags & 1)) {
              if (kCross16) split16_stage<Cfg, kMTiles>(ring_ptr + stage * Cfg::kStageBytes, lane, me % kWarpsPerStage);
              else split_stage<Cfg>(ring_ptr + stage * Cfg::kStageBytes, lane, me % kWarpsPerStage);
            }
            fence_proxy_async_smem();  // generic-proxy stores -> visible to tcgen05.mma's operand reads
            __syncwarp();
            if (lane == 0) ready_arrive(stage);
            if (p.dbg_cycles && lane == 0) d_split += clk() - t1;
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
      if (p.dbg_cycles && lane == 0) {
        unsigned long long* o = reinterpret_cast<unsigned long long*>(p.dbg_cycles + (size_t)(gridDim.x + blockIdx.x) * 8);
        atomicAdd(o + 0, (unsigned long long)d_tma); atomicAdd(o + 1, (unsigned long long)d_split);
        atomicAdd(o + 2, (unsigned long long)d_n);
      }
    }
  } else {
    // ================= epilogue: 8 warps.  Warp w may touch TMEM lanes [32*(w%4), +32); the two warps
    // of a lane quarter take alternate 32-column chunks, so each scheduler has two warps to overlap
    // the memory round trips of the fused update. =================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    int acc = 0, stage = 0, hbase = 0, split_seq = 0;
    uint32_t acc_phase = 0, phase = 0;
    long long c_accfull = 0, c_hfull = 0, d_tma = 0, d_split = 0, d_n = 0;
    const long long c_estart = clk();
    for (int item = first_item; item < num_items; item += item_stride) {
      const WorkItem w = decode_item(p, item, kBlockT);
      const int m_group = w.m_group, split = w.split;
      const int t0 = w.t_tile * kBlockT + w.t_off;
      if (kSplit3 && !Cfg::kDedicatedXform) {
        // single accumulator stage: these warps have nothing to drain during the main loop, so they
        // produce the lo tiles, one warp per K-block round-robin
        const int me = warp - 2;
        for (int kb = w.kb0; kb < w.kb1; ++kb, ++split_seq) {
          mbar_wait_backoff(smem_u32(&bar_full[stage]), phase, EVC_SPLIT_SLEEP_NS);  // (all warps see all phases; see the dedicated warps)
          if (kEarly && me == 0 && lane == 0) landed_arrive(stage);
          if (split_seq % Cfg::kSplitGroups == me / kWarpsPerStage) {
            const long long t1 = clk();
            if (p.dbg_cycles && lane == 0) { d_tma += t1 - dbg_t_issue[stage]; ++d_n; }
            if (!(p.debug_flags & 1)) {
              if (kCross16) split16_stage<Cfg, kMTiles>(ring_ptr + stage * Cfg::kStageBytes, lane, me % kWarpsPerStage);
              else split_stage<Cfg>(ring_ptr + stage * Cfg::kStageBytes, lane, me % kWarpsPerStage);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ready_arrive(stage);
            if (p.dbg_cycles && lane == 0) d_split += clk() - t1;
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
      long long c0 = clk();
      mbar_wait_backoff(smem_u32(&bar_acc_full[acc]), acc_phase, EVC_EPI_SLEEP_NS);
      c_accfull += clk() - c0;
      tc_fence_after();
      if (kStageH) {
        // ---- fused multiplicative update through the shared-memory H chunks ----
        const int m = m_group * Cfg::kRowsPerSub + (int)rank * 128 + quarter * 32 + lane;
        float den = ((m < p.M_total && !kFro) ? p.colsum[m] : 1.f) + p.lam;
        if (den == 0.f) den = p.eps;
        const float inv_den = __frcp_rn(den);
        const int nch = max(0, min(w.t_cols / kHChunkT, (p.T - t0 + kHChunkT - 1) / kHChunkT));
        for (int c = half; c < nch; c += 2) {
          const int seq = hbase + c, b = seq % kHB;
          const uint32_t ph = (uint32_t)(seq / kHB) & 1u;
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kBlockT + c * 32), v);
          c0 = clk();
          mbar_wait(smem_u32(&bar_hfull[b]), ph);
          c_hfull += clk() - c0;
          float* hb = reinterpret_cast<float*>(ring_ptr + Cfg::kOffH + b * Cfg::kHBufStride) + quarter * 32 + lane;
          float h[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) h[j] = hb[j * 128];
          tmem_ld_wait();
          const int tbm = t0 + c * 32;
          if (kFro) {
            // Frobenius: H <- H * (X A^T) / (A^T (A H) + lambda); the numerator chunk sits behind the H chunk
            const float* qb = hb + kHBufBytes / 4;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (p.row_active == nullptr || (tbm + j < p.T && p.row_active[tbm + j])) {
                float dn = __uint_as_float(v[j]) + p.lam;
                if (dn == 0.f) dn = p.eps;
                h[j] = h[j] * __fdividef(qb[j * 128], dn);
              }
            }
          } else if (p.row_active == nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) h[j] = h[j] * (__uint_as_float(v[j]) * inv_den);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (tbm + j < p.T && p.row_active[tbm + j]) h[j] = h[j] * (__uint_as_float(v[j]) * inv_den);
          }
          if (kBf16 && p.out16 != nullptr && m < p.M_total && !(p.debug_flags & 8)) {
            // bf16 shadow of the updated activations: 64 contiguous bytes per warp per frame, straight from registers
            __nv_bfloat16* o16 = p.out16 + (size_t)tbm * p.ld_out16 + m;
            const int rows = min(32, p.T - tbm);
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < rows) o16[(size_t)j * p.ld_out16] = __float2bfloat16_rn(h[j]);
          }
          if (p.direct_store) {
            // registers -> global, 128 B per warp per frame; the chunk buffer is free as soon as it was read
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_hempty[b]));
            if (m < p.M_total && !(p.debug_flags & 8)) {
              float* o = p.out + (size_t)(t0 + c * 32) * p.ld_out + m;
              const int rows = min(32, p.T - (t0 + c * 32));
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < rows) o[(size_t)j * p.ld_out] = h[j];
            }
          } else {
            if (!(p.debug_flags & 8)) {
#pragma unroll
              for (int j = 0; j < 32; ++j) hb[j * 128] = h[j];
            }
            fence_proxy_async_smem();  // generic-proxy stores -> visible to the TMA store
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_hready[b]));
          }
          for (int l = 0; l < p.n_left; ++l) {
            // (rows past T and exemplars past N were zero-filled by TMA: they add nothing)
            const float a = (m < p.M_total) ? p.left_a[(size_t)l * p.left_lda + m] : 0.f;
            float sv[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) sv[j] = h[j] * a;
            const float tot = warp_transpose_sum(sv, lane);
            const int prow = (m_group * kCG + (int)rank) * 4 + quarter;  // one partial row per 32 exemplars
            p.left_out[((size_t)l * p.left_ld + (t0 + c * 32 + lane)) * p.left_rows + prow] = tot;
          }
        }
        hbase += nch;
      }
#pragma unroll
      for (int i = 0; i < (kStageH ? 0 : kMTiles); ++i) {
        const int mrow0 = m_group * (Cfg::kRowsPerSub * kMTiles) + i * Cfg::kRowsPerSub + (int)rank * 128;
        if (m_group * (Cfg::kRowsPerSub * kMTiles) + i * Cfg::kRowsPerSub >= p.M_total) break;  // whole MMA is padding
        const int m = mrow0 + quarter * 32 + lane;
        const bool m_ok = m < p.M_total;
        const bool rows_full = (mrow0 + 128 <= p.M_total) && (p.row_active == nullptr);
        for (int c = half; c < w.t_cols / 32; c += 2) {
          const int tb = t0 + c * 32;
          if (tb >= p.T) break;  // warp-uniform
          if (p.debug_flags & 8) continue;
          uint32_t v[32];
          const uint32_t taddr =
              tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((acc * kMTiles + i) * kBlockT + c * 32);
          tmem_ld_32x32(taddr, v);
          tmem_ld_wait();
          // whole chunk in range, every lane a real row, no frozen utterances: straight-line code
          const bool fast = (tb + 32 <= p.T) && rows_full;
          if (kEpi == TEPI_PARTIAL) {
            float* o = p.out + ((size_t)split * p.T + tb) * p.ld_out + m;
            if (fast) {
#pragma unroll
              for (int j = 0; j < 32; ++j) o[(size_t)j * p.ld_out] = __uint_as_float(v[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (m_ok && tb + j < p.T) o[(size_t)j * p.ld_out] = __uint_as_float(v[j]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCG == 2) mbar_arrive_cluster(mapa_rank(smem_u32(&bar_acc_empty[acc]), 0));
        else mbar_arrive(smem_u32(&bar_acc_empty[acc]));
      }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
    }
    if (p.dbg_cycles && warp == 2 && lane == 0) {
      long long* o = p.dbg_cycles + (size_t)blockIdx.x * 8;
      o[5] = clk() - c_estart; o[6] = c_accfull; o[7] = c_hfull;
    }
    if (p.dbg_cycles && lane == 0 && kSplit3 && !Cfg::kDedicatedXform) {
      unsigned long long* o = reinterpret_cast<unsigned long long*>(p.dbg_cycles + (size_t)(gridDim.x + blockIdx.x) * 8);
      atomicAdd(o + 0, (unsigned long long)d_tma); atomicAdd(o + 1, (unsigned long long)d_split);
      atomicAdd(o + 2, (unsigned long long)d_n);
    }
  }

  tc_fence_before();
  // nobody leaves (and frees shared memory / TMEM the leader's MMAs may still read) before both CTAs are done
  if (kCG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if (kCG == 2) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    if (p.dbg_cycles && lane == 0) {
      // CTA lifetime: entry -> barriers/TMEM set up -> predecessor grid done (PDL) -> ... -> TMEM released
      long long* o2 = p.dbg_cycles + (size_t)(gridDim.x + blockIdx.x) * 8;
      o2[5] = t_setup - t_entry; o2[6] = t_go - t_setup; o2[7] = clk() - t_entry;
    }
  }
}

// ---- memory-bound helpers ------------------------------------------------------------------------

// WH[t,f] = sum_s P[s][t][f] for the tensor-core rows f < F_main (fixed order: deterministic);
// WH[t,F_main+l] = sum_r L[l][t][r] from the fused update's per-warp partials when `left_rows` > 0 (the block
// that owns those columns reduces the `left_rows` contiguous partials of its frame first).
// With `R` != nullptr it also emits the ratio R = X / max(WH, eps) (zero pad columns) in the same pass.
__global__ void __launch_bounds__(128)
reduce_partials_kernel(const float* __restrict__ P, int S, int S_last, int f_last, int T, int ldp, int F, int F_main,
                       float* __restrict__ WH, int ldwh, const float* __restrict__ L, int left_rows, int n_left,
                       int left_ld, const float* __restrict__ X, int ldx, float* __restrict__ R, int ldr, float eps,
                       __nv_bfloat16* __restrict__ R16, int ldr16) {
  // one block per frame; a thread owns groups of 4 consecutive columns (16-byte loads of the partials)
  __shared__ float s_left[8];
  __shared__ float s_warp[4];
  pdl_wait();
  pdl_launch_dependents();
  const int t = blockIdx.x;
  if (t >= T) return;
  if (left_rows > 0) {
    for (int l = 0; l < n_left; ++l) {
      const float* row = L + ((size_t)l * left_ld + t) * left_rows;
      float a = 0.f;
      for (int r = threadIdx.x; r < left_rows; r += blockDim.x) a += row[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = a;
      __syncthreads();
      if (threadIdx.x == 0) s_left[l] = (s_warp[0] + s_warp[1]) + (s_warp[2] + s_warp[3]);
      __syncthreads();
    }
  }
  const int cols = max(ldwh, max(R ? ldr : 0, R16 ? ldr16 : 0));
  for (int f0 = threadIdx.x * 4; f0 < cols; f0 += blockDim.x * 4) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    bool have[4] = {false, false, false, false};
    if (f0 + 4 <= F_main && (f0 >= f_last || f0 + 4 <= f_last)) {
      const int n = (f0 >= f_last) ? S_last : S;  // columns of the last row group have their own split count
      const float* p0 = P + (size_t)t * ldp + f0;
      // batches of 6 independent 16-byte loads in flight, summed in split order (deterministic)
      int k = 0;
      for (; k + 6 <= n; k += 6) {
        float4 v[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) v[q] = *reinterpret_cast<const float4*>(p0 + (size_t)(k + q) * T * ldp);
#pragma unroll
        for (int q = 0; q < 6; ++q) { s[0] += v[q].x; s[1] += v[q].y; s[2] += v[q].z; s[3] += v[q].w; }
      }
      for (; k < n; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(p0 + (size_t)k * T * ldp);
        s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      }
      have[0] = have[1] = have[2] = have[3] = true;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int f = f0 + j;
        if (f < F_main) {
          const int n = (f >= f_last) ? S_last : S;
          for (int k = 0; k < n; ++k) s[j] += P[((size_t)k * T + t) * ldp + f];
          have[j] = true;
        } else if (f < F && left_rows > 0) {
          s[j] = s_left[f - F_main];
          have[j] = true;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int f = f0 + j;
      if (f < ldwh && (have[j] || f >= F)) WH[(size_t)t * ldwh + f] = s[j];
      if ((R && f < ldr) || (R16 && f < ldr16)) {
        const float r = (have[j] && f < F) ? __fdiv_rn(X[(size_t)t * ldx + f], fmaxf(s[j], eps)) : 0.f;
        if (R && f < ldr) R[(size_t)t * ldr + f] = r;
        if (R16 && f < ldr16) R16[(size_t)t * ldr16 + f] = __float2bfloat16_rn(r);
      }
    }
  }
}

// Standalone leftover rows: WH[t, F_main+l] = sum_n H[t,n] * a[l][n].  One block per frame; used whenever the
// fused update has not just produced the partials (first iteration, objective of a given H, conversion).
__global__ void __launch_bounds__(256)
leftover_rows_kernel(const float* __restrict__ H, int ldh, int T, int N, const float* __restrict__ a, int lda,
                     int n_left, float* __restrict__ WH, int ldwh, int F_main) {
  __shared__ float red[8][8];
  const int t = blockIdx.x;
  float acc[8];
#pragma unroll
  for (int l = 0; l < 8; ++l) acc[l] = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float h = H[(size_t)t * ldh + n];
#pragma unroll
    for (int l = 0; l < 8; ++l)
      if (l < n_left) acc[l] = fmaf(h, a[(size_t)l * lda + n], acc[l]);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int l = 0; l < 8; ++l) {
    float v = acc[l];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][l] = v;
  }
  __syncthreads();
  if (threadIdx.x < n_left) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    WH[(size_t)t * ldwh + F_main + threadIdx.x] = v;
  }
}

// R = X / max(WH, eps) with zeroed pad columns; `copy` = 1 stores WH itself (Frobenius: the second
// contraction multiplies A^T with A H; also used to stage X for the Frobenius numerator).
// With `R16` the result is stored as bf16 (pitch ldr16) instead: the K operand of the BF16 mode.
__global__ void ratio_pad_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ WH, int ldwh,
                                 float* __restrict__ R, int ldr, int T, int F, float eps, int copy,
                                 __nv_bfloat16* __restrict__ R16, int ldr16) {
  const int f = blockIdx.y * blockDim.x + threadIdx.x;
  const int t = blockIdx.x;
  if (t >= T || f >= (R16 ? ldr16 : ldr)) return;
  float r = 0.f;
  if (f < F) {
    const float wh = WH[(size_t)t * ldwh + f];
    r = copy ? wh : __fdiv_rn(X[(size_t)t * ldx + f], fmaxf(wh, eps));
  }
  if (R16) R16[(size_t)t * ldr16 + f] = __float2bfloat16_rn(r);
  else R[(size_t)t * ldr + f] = r;
}

// dst (rows, ldd) bf16 = src (rows, lds) fp32, round to nearest even; pad columns [cols, ldd) are zeroed.
__global__ void to_bf16_kernel(const float* __restrict__ src, int lds, __nv_bfloat16* __restrict__ dst, int ldd,
                               int rows, int cols) {
  const int r = blockIdx.x;
  const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (r >= rows || c0 >= ldd) return;
  const float* sp = src + (size_t)r * lds;
  __nv_bfloat16* dp = dst + (size_t)r * ldd;
  if (c0 + 4 <= cols && (lds & 3) == 0 && (ldd & 3) == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0) {
    const float4 v = *reinterpret_cast<const float4*>(sp + c0);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(dp + c0) = pk;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + j;
      if (c < ldd) dp[c] = __float2bfloat16_rn(c < cols ? sp[c] : 0.f);
    }
  }
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

// Row-major matrix (rows, cols) of `esize`-byte elements (4: fp32, 2: bf16) with pitch ld elements;
// box = box_rows x box_cols elements, box_cols*esize in {64,128} when swizzled.
inline int make_tmap_any(CUtensorMap* m, const void* base, int esize, int rows, int cols, int ld, int box_cols,
                         int box_rows, bool swizzle) {
  PFN_encodeTiled enc;
  EVC_TRY(get_encode(&enc));
  if (((uintptr_t)base & 15) || (((size_t)ld * esize) & 15))
    return fail(EVC_ERR_INVALID_ARGUMENT, "tensor-core modes need 16-byte aligned matrices with a 16-byte multiple pitch");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esize};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = !swizzle ? CU_TENSOR_MAP_SWIZZLE_NONE
                                : (box_cols * esize == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(m, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(EVC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d box=%dx%d", (int)r, rows, cols,
                ld, box_rows, box_cols);
  return EVC_OK;
}
inline int make_tmap(CUtensorMap* m, const float* base, int rows, int cols, int ld, int box_cols, int box_rows,
                     bool swizzle = true) {
  return make_tmap_any(m, base, 4, rows, cols, ld, box_cols, box_rows, swizzle);
}
inline int make_tmap16(CUtensorMap* m, const __nv_bfloat16* base, int rows, int cols, int ld, int box_cols, int box_rows) {
  return make_tmap_any(m, base, 2, rows, cols, ld, box_cols, box_rows, true);
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

// Programmatic dependent launch between the kernels of an iteration (EVC_NO_PDL=1 turns it off).
inline bool use_pdl() {
  static const bool on = getenv("EVC_NO_PDL") == nullptr;
  return on;
}

template <int kMTiles, int kBlockT, int kBlockK, int kSplit3, int kEpi, int kCG, bool kBf16 = false>
inline int launch_tc(const CUtensorMap& tmM, const CUtensorMap& tmN, const CUtensorMap& tmH, const CUtensorMap& tmQ,
                     const GemmParams& p, cudaStream_t s) {
  using Cfg = TileCfg<kMTiles, kBlockT, kBlockK, kSplit3, kEpi == TEPI_MU_KL || kEpi == TEPI_MU_FRO, kCG, kEpi == TEPI_MU_FRO>;
  auto kern = tc_gemm_kernel<kMTiles, kBlockT, kBlockK, kSplit3, kEpi, kCG, kBf16>;
  static bool configured = false;
  if (!configured) {
    EVC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured = true;
  }
  const int items = p.items_main + p.splits_last * p.num_t_tiles + (p.items_main - p.half_from);
  if (items <= 0) return EVC_OK;
  const int slots = num_sms() / kCG;  // CTAs (kCG = 1) or CTA pairs (kCG = 2) resident at once
  const int grid = (items < slots ? items : slots) * kCG;
  static const int dbg = getenv("EVC_DEBUG_FLAGS") ? atoi(getenv("EVC_DEBUG_FLAGS")) : 0;
  GemmParams q = p;
  q.debug_flags = dbg;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Cfg::kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl() ? 2 : 1;
  static const bool timing = getenv("EVC_DEBUG_TIMING") != nullptr;
  static long long* dbuf = nullptr;
  static int prints = 0;
  if (timing) {
    if (!dbuf) EVC_CUDA(cudaMalloc(&dbuf, 8 * sizeof(long long) * 1024));
    EVC_CUDA(cudaMemsetAsync(dbuf, 0, 8 * sizeof(long long) * grid * 2, s));
    q.dbg_cycles = dbuf;
  }
  EVC_CUDA(cudaLaunchKernelEx(&cfg, kern, tmM, tmN, tmH, tmQ, q));
  EVC_LAUNCH_CHECK();
  if (timing && prints < 6) {
    std::vector<long long> h((size_t)grid * 16);
    EVC_CUDA(cudaStreamSynchronize(s));
    EVC_CUDA(cudaMemcpy(h.data(), dbuf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    double a[8] = {0}; int nl = 0;
    for (int b = 0; b < grid; b += kCG) { for (int k = 0; k < 8; ++k) a[k] += (double)h[(size_t)b * 8 + k]; ++nl; }
    fprintf(stderr, "[evc timing] kernel<%d,%d,%d,%d,%d,cg%d%s> grid %d (leaders avg, cycles): mma loop %.0f (wait acc_empty %.0f, wait ready %.0f) | "
            "producer loop %.0f (wait empty %.0f) | epilogue warp loop %.0f (wait acc_full %.0f, wait hfull %.0f)\n",
            kMTiles, kBlockT, kBlockK, (int)kSplit3, kEpi, kCG, kBf16 ? ",bf16" : "", grid, a[0] / nl, a[1] / nl, a[2] / nl, a[3] / nl, a[4] / nl,
            a[5] / nl, a[6] / nl, a[7] / nl);
    double tma = 0, spl = 0, cnt = 0;
    for (int b = 0; b < grid; ++b) { tma += (double)h[(size_t)(grid + b) * 8]; spl += (double)h[(size_t)(grid + b) * 8 + 1]; cnt += (double)h[(size_t)(grid + b) * 8 + 2]; }
    {
      double iss = 0, com = 0;
      for (int b = 0; b < grid; b += kCG) { iss += (double)h[(size_t)(grid + b) * 8 + 3]; com += (double)h[(size_t)(grid + b) * 8 + 4]; }
      fprintf(stderr, "[evc timing]    MMA thread: issuing MMAs %.0f cycles, commits %.0f cycles (leaders avg)\n", iss / nl, com / nl);
    }
    {
      double su = 0, pw = 0, life = 0, lmax = 0;
      for (int b = 0; b < grid; ++b) {
        su += (double)h[(size_t)(grid + b) * 8 + 5]; pw += (double)h[(size_t)(grid + b) * 8 + 6];
        life += (double)h[(size_t)(grid + b) * 8 + 7]; lmax = std::max(lmax, (double)h[(size_t)(grid + b) * 8 + 7]);
      }
      fprintf(stderr, "[evc timing]    CTA lifetime (all CTAs): set-up %.0f cycles, PDL wait %.0f, total avg %.0f max %.0f\n",
              su / grid, pw / grid, life / grid, lmax);
    }
    if (cnt > 0) fprintf(stderr, "[evc timing]    per K-block: TMA issue -> landed %.0f cycles, landed -> split done + ready arrive %.0f cycles (%.0f blocks)\n", tma / cnt, spl / cnt, cnt);
    ++prints;
  }
  return EVC_OK;
}

// Tile shapes per contraction.  K block = one swizzle row: 16 fp32 (64 B) when both hi and lo tiles
// sit in the ring (3xTF32), 32 fp32 (128 B) otherwise, so a ring stage is 48-64 KB and 3-4 stages fit.
constexpr int kC1MTiles = 2, kC1BlockT = 256;  // contraction 1 / conversion: 256 dictionary rows x 256 frames, split-K
constexpr int kC2MTiles = 1, kC2BlockT = 256;  // contraction 2: 128 exemplars x 256 frames, 2 accumulator stages
constexpr int kBlockK3 = 16, kBlockK1 = 32;
// K elements per K-block of a mode (BF16: a 128-byte swizzle row holds 64 elements)
inline int bk_elems(int mode) {
  return mode == EVC_MODE_3XTF32 ? kBlockK3 : mode == EVC_MODE_BF16 ? 2 * kBlockK1 : kBlockK1;
}

// CTA group of the MMAs: 2 (CTA pairs, default) or 1 (EVC_CTA_GROUP=1: single-CTA kernels, kept for A/B runs).
// fp32-accurate mode: 2 = tf32 hi*hi + two bf16 cross-term MMAs (default), 1 = three tf32 MMAs per product
// (EVC_SPLIT_CROSS16=0; kept for A/B runs and as the accuracy reference)
inline int split_flavor() {
  static const int f = (getenv("EVC_SPLIT_CROSS16") && atoi(getenv("EVC_SPLIT_CROSS16")) == 0) ? 1 : 2;
  return f;
}
inline int cta_group() {
  static const int cg = (getenv("EVC_CTA_GROUP") && atoi(getenv("EVC_CTA_GROUP")) == 1) ? 1 : 2;
  return cg;
}

// Resident tensor-core operands of one dictionary.  All fp32: the hi operand of 3xTF32 is the raw value
// (the MMA truncates it), the lo operand is derived in shared memory, so HBM holds one copy per layout.
struct DictOperands {
  int F = 0, N = 0, ldA = 0, ldN = 0;
  bool has_target = false;
  const float* A = nullptr;  // (N, ldA), borrowed from the handle: K-major operand of contraction 2
  float* AT = nullptr;       // (F, ldN) transposed copy: K-major operand of contraction 1
  float* BT = nullptr;       // (F, ldN) transposed target dictionary: operand of the conversion
  CUtensorMap tmA, tmAT, tmBT;
  int F_main = 0, n_left = 0;  // contraction 1 runs rows [0, F_main) on the tensor cores; n_left = F - F_main <= 8
  bool left_valid = false;     // the workspace holds leftover partials of the CURRENT activations
  // BF16 mode: bf16 copies of the three operands, and the per-solve bf16 shadows of H and of the ratio
  int ldA16 = 0, ldN16 = 0;
  __nv_bfloat16 *A16 = nullptr, *AT16 = nullptr, *BT16 = nullptr;
  CUtensorMap tmA16, tmAT16, tmBT16;
  DevBuf h16, r16;
  void release() {
    cudaFree(AT); cudaFree(BT); cudaFree(A16); cudaFree(AT16); cudaFree(BT16);
    AT = BT = nullptr;
    A16 = AT16 = BT16 = nullptr;
    h16.release(); r16.release();
  }
};

inline int launch_to_bf16(const float* src, int lds, __nv_bfloat16* dst, int ldd, int rows, int cols, cudaStream_t s) {
  if (rows <= 0) return EVC_OK;
  dim3 g(rows, ceil_div(ldd, 4 * 256));
  to_bf16_kernel<<<g, 256, 0, s>>>(src, lds, dst, ldd, rows, cols);
  EVC_LAUNCH_CHECK();
  return EVC_OK;
}

inline int build_operands(DictOperands* o, int mode, const float* A, const float* B, int ldA, int F, int N, cudaStream_t s) {
  const int bk = (mode == EVC_MODE_3XTF32) ? kBlockK3 : kBlockK1;  // (fp32 maps; BF16 builds its own below)
  o->F = F; o->N = N; o->ldA = ldA; o->ldN = round_up(N, 4); o->A = A; o->has_target = (B != nullptr);
  // A few rows past a multiple of 128 (the Nyquist bin of a 513-bin spectrum) would cost a whole 128-row MMA
  // tile; they are handled as dot products on the CUDA cores instead.
  o->n_left = (F > 128 && (F % 128) <= 8 && !getenv("EVC_NO_LEFTOVER")) ? F % 128 : 0;
  o->F_main = F - o->n_left;
  const size_t at_bytes = (size_t)F * o->ldN * sizeof(float);
  dim3 tb(32, 8), tg(ceil_div(N, 32), ceil_div(F, 32));
  EVC_CUDA(cudaMalloc(&o->AT, at_bytes));
  EVC_CUDA(cudaMemsetAsync(o->AT, 0, at_bytes, s));
  simt::transpose_kernel<<<tg, tb, 0, s>>>(A, ldA, o->AT, o->ldN, N, F);
  EVC_LAUNCH_CHECK();
  EVC_TRY(make_tmap(&o->tmA, A, N, F, ldA, bk, 128));
  EVC_TRY(make_tmap(&o->tmAT, o->AT, o->F_main, N, o->ldN, bk, 128));
  if (B) {
    EVC_CUDA(cudaMalloc(&o->BT, at_bytes));
    EVC_CUDA(cudaMemsetAsync(o->BT, 0, at_bytes, s));
    simt::transpose_kernel<<<tg, tb, 0, s>>>(B, ldA, o->BT, o->ldN, N, F);
    EVC_LAUNCH_CHECK();
    EVC_TRY(make_tmap(&o->tmBT, o->BT, o->F_main, N, o->ldN, bk, 128));
  }
  if (mode == EVC_MODE_BF16) {
    const int bk16 = bk_elems(mode);
    o->ldA16 = round_up(F, 8); o->ldN16 = round_up(N, 8);
    EVC_CUDA(cudaMalloc(&o->A16, (size_t)N * o->ldA16 * sizeof(__nv_bfloat16)));
    EVC_CUDA(cudaMalloc(&o->AT16, (size_t)F * o->ldN16 * sizeof(__nv_bfloat16)));
    EVC_TRY(launch_to_bf16(A, ldA, o->A16, o->ldA16, N, F, s));
    EVC_TRY(launch_to_bf16(o->AT, o->ldN, o->AT16, o->ldN16, F, N, s));
    EVC_TRY(make_tmap16(&o->tmA16, o->A16, N, F, o->ldA16, bk16, 128));
    EVC_TRY(make_tmap16(&o->tmAT16, o->AT16, o->F_main, N, o->ldN16, bk16, 128));
    if (B) {
      EVC_CUDA(cudaMalloc(&o->BT16, (size_t)F * o->ldN16 * sizeof(__nv_bfloat16)));
      EVC_TRY(launch_to_bf16(o->BT, o->ldN, o->BT16, o->ldN16, F, N, s));
      EVC_TRY(make_tmap16(&o->tmBT16, o->BT16, o->F_main, N, o->ldN16, bk16, 128));
    }
  }
  return EVC_OK;
}

// Split-K plan of contraction 1: fill the SMs with (m_group, t_tile, split) work items carrying equal MMA
// counts.  Groups of kC1MTiles sub-tiles get `splits` K ranges; a last group with fewer sub-tiles gets
// `splits_last` longer ones.
struct C1Plan {
  int m_groups, t_tiles, kb_total, ldp;
  int splits, kb_per_split;            // full groups
  int splits_last, kb_per_split_last;  // last, partial group (0 = none)
  int f_last;                          // first dictionary row of the last group
  int max_splits;
};
inline C1Plan plan_c1(int F, int N, int T, int bk) {
  C1Plan pl{};
  const int cg = cta_group();
  const int sub_rows = 128 * cg;  // dictionary rows of one MMA
  const int tiles = ceil_div(F, sub_rows);
  pl.m_groups = ceil_div(tiles, kC1MTiles);
  pl.t_tiles = ceil_div(T, kC1BlockT);
  pl.kb_total = ceil_div(N, bk);
  pl.ldp = round_up(F, 32);
  const int tiles_last = tiles - (pl.m_groups - 1) * kC1MTiles;
  const bool partial = tiles_last < kC1MTiles;
  const int full_groups = partial ? pl.m_groups - 1 : pl.m_groups;
  pl.f_last = partial ? full_groups * kC1MTiles * sub_rows : F;
  // w = sub-tile K-blocks per CTA; grow it until the items fit the SMs
  const int slots = num_sms() / cg;  // CTAs or CTA pairs resident at once
  long long total = (long long)tiles * pl.kb_total * pl.t_tiles;
  int w = (int)std::max<long long>(1, (total + slots - 1) / slots);
  for (;; ++w) {
    const int per = std::max(1, ceil_div(w, kC1MTiles));
    const int sf = full_groups ? ceil_div(pl.kb_total, std::min(per, pl.kb_total)) : 0;
    const int per_l = std::max(1, ceil_div(w, tiles_last));
    const int sl = partial ? ceil_div(pl.kb_total, std::min(per_l, pl.kb_total)) : 0;
    const long long items = (long long)pl.t_tiles * (full_groups * sf + sl);
    if (items <= slots || (sf <= 1 && sl <= 1)) {
      pl.splits = sf; pl.kb_per_split = std::min(per, pl.kb_total);
      pl.splits_last = sl; pl.kb_per_split_last = std::min(per_l, pl.kb_total);
      break;
    }
  }
  pl.max_splits = std::max(pl.splits, pl.splits_last);
  return pl;
}

// Called before a solve / product: make sure the workspace can hold the split-K partials.
// Workspace layout: [ split-K partials | leftover-row partials of the fused update ].
inline size_t ws_left_offset(const C1Plan& pl, int T) { return round_up_sz((size_t)pl.max_splits * T * pl.ldp, 64); }
inline int left_rows(const DictOperands& o) { return round_up(ceil_div(o.N, 128), cta_group()) * 4; }
inline int left_ld(int T) { return round_up(T, kC2BlockT); }

inline int after_h_written(DictOperands& o, int mode, const float* H, int ldH, int T, DevBuf* ws, cudaStream_t s) {
  if (mode == EVC_MODE_FP32) return EVC_OK;
  const C1Plan pl = plan_c1(o.F_main, o.N, T, bk_elems(mode));
  o.left_valid = false;
  const size_t left = (size_t)left_rows(o) * o.n_left * left_ld(T);
  EVC_TRY(ws->reserve((ws_left_offset(pl, T) + left) * sizeof(float)));
  if (mode == EVC_MODE_BF16) {
    // the bf16 shadow of these activations (afterwards the fused update keeps it current) and room for the ratio
    EVC_TRY(o.h16.reserve((size_t)T * o.ldN16 * sizeof(__nv_bfloat16)));
    EVC_TRY(o.r16.reserve((size_t)T * o.ldA16 * sizeof(__nv_bfloat16)));
    EVC_TRY(launch_to_bf16(H, ldH, o.h16.as<__nv_bfloat16>(), o.ldN16, T, o.N, s));
  }
  return EVC_OK;
}

struct RatioArgs {  // fuse R = X / max(WH, eps) into the split-K reduction
  const float* X; int ldX; float* R; int ldR; float eps;
};

inline int launch_ratio(const float* X, int ldX, const float* WH, int ldWH, float* R, int ldR, int T, int F,
                        float eps, int copy, cudaStream_t s, __nv_bfloat16* R16 = nullptr, int ldR16 = 0) {
  ProfScope ps(1, s);
  dim3 g(T, ceil_div(R16 ? ldR16 : ldR, 128));
  ratio_pad_kernel<<<g, 128, 0, s>>>(X, ldX, WH, ldWH, R, ldR, T, F, eps, copy, R16, ldR16);
  EVC_LAUNCH_CHECK();
  return EVC_OK;
}

template <int kSplit3, int kCG, bool kBf16 = false>
inline int contract_wh_t(DictOperands& o, const float* H, int ldH, int T, float* WH, int ldWH, bool target,
                         DevBuf* ws, cudaStream_t s, const RatioArgs* ra) {
  constexpr int bk = kSplit3 ? kBlockK3 : kBlockK1;   // K-block in 4-byte words (the kernel's template argument)
  constexpr int bke = kBf16 ? 2 * bk : bk;            // ... in elements
  const C1Plan pl = plan_c1(o.F_main, o.N, T, bke);
  float* partials = ws->as<float>();
  float* leftp = ws->as<float>() + ws_left_offset(pl, T);
  CUtensorMap tmH;
  // BF16: the K operand is the shadow of H (made by after_h_written, kept current by the fused update)
  if (kBf16) EVC_TRY(make_tmap16(&tmH, o.h16.as<__nv_bfloat16>(), T, o.N, o.ldN16, bke, kC1BlockT / kCG));
  else EVC_TRY(make_tmap(&tmH, H, T, o.N, ldH, bk, kC1BlockT / kCG));
  GemmParams p{};
  p.M_total = o.F_main; p.T = T; p.K = o.N;
  p.num_m_groups = pl.m_groups; p.num_t_tiles = pl.t_tiles; p.num_splits = pl.splits;
  p.kblocks_per_split = pl.kb_per_split; p.kblocks_total = pl.kb_total;
  p.splits_last = pl.splits_last; p.kblocks_per_split_last = pl.kb_per_split_last;
  p.items_main = (pl.splits_last ? pl.m_groups - 1 : pl.m_groups) * pl.t_tiles * pl.splits;
  p.half_from = p.items_main;
  p.out = partials; p.ld_out = pl.ldp;
  {
    ProfScope ps(0, s);
    const CUtensorMap& tmD = kBf16 ? (target ? o.tmBT16 : o.tmAT16) : (target ? o.tmBT : o.tmAT);
    EVC_TRY((launch_tc<kC1MTiles, kC1BlockT, bk, kSplit3, TEPI_PARTIAL, kCG, kBf16>(tmD, tmH, tmH, tmH, p, s)));
  }
  // leftover rows: from the fused update's partials when they describe this H, else a dot-product pass over H
  const bool from_partials = o.n_left > 0 && !target && o.left_valid;
  const bool standalone = o.n_left > 0 && !from_partials;
  {
    ProfScope ps(1, s);
    dim3 g(T, 1);
    const bool fuse = ra && !standalone;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = g; cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = use_pdl() ? 1 : 0;
    EVC_CUDA(cudaLaunchKernelEx(&cfg, reduce_partials_kernel, (const float*)partials, pl.splits, pl.splits_last, pl.f_last, T,
                                pl.ldp, o.F, o.F_main, WH, ldWH, (const float*)leftp, from_partials ? left_rows(o) : 0,
                                o.n_left, left_ld(T), fuse ? ra->X : (const float*)nullptr, fuse ? ra->ldX : 0,
                                (fuse && !kBf16) ? ra->R : (float*)nullptr, fuse ? ra->ldR : 0, fuse ? ra->eps : 0.f,
                                (fuse && kBf16) ? o.r16.as<__nv_bfloat16>() : (__nv_bfloat16*)nullptr, o.ldA16));
    EVC_LAUNCH_CHECK();
    if (standalone) {
      const float* rows = (target ? o.BT : o.AT) + (size_t)o.F_main * o.ldN;
      leftover_rows_kernel<<<T, 256, 0, s>>>(H, ldH, T, o.N, rows, o.ldN, o.n_left, WH, ldWH, o.F_main);
      EVC_LAUNCH_CHECK();
    }
  }
  if (ra && standalone)
    EVC_TRY(launch_ratio(ra->X, ra->ldX, WH, ldWH, ra->R, ra->ldR, T, o.F, ra->eps, 0, s,
                         kBf16 ? o.r16.as<__nv_bfloat16>() : nullptr, o.ldA16));
  return EVC_OK;
}

inline int contract_wh(DictOperands& o, int mode, const float* H, int ldH, int T, float* WH, int ldWH, bool target,
                       DevBuf* ws, cudaStream_t s, const RatioArgs* ra = nullptr) {
  if (target && !o.has_target) return fail(EVC_ERR_INVALID_ARGUMENT, "no target dictionary");
  if (cta_group() == 2) {
    if (mode == EVC_MODE_3XTF32 && split_flavor() == 2) return contract_wh_t<2, 2>(o, H, ldH, T, WH, ldWH, target, ws, s, ra);
    if (mode == EVC_MODE_3XTF32) return contract_wh_t<1, 2>(o, H, ldH, T, WH, ldWH, target, ws, s, ra);
    if (mode == EVC_MODE_BF16) return contract_wh_t<false, 2, true>(o, H, ldH, T, WH, ldWH, target, ws, s, ra);
    return contract_wh_t<false, 2>(o, H, ldH, T, WH, ldWH, target, ws, s, ra);
  }
  if (mode == EVC_MODE_3XTF32 && split_flavor() == 2) return contract_wh_t<2, 1>(o, H, ldH, T, WH, ldWH, target, ws, s, ra);
  if (mode == EVC_MODE_3XTF32) return contract_wh_t<1, 1>(o, H, ldH, T, WH, ldWH, target, ws, s, ra);
  if (mode == EVC_MODE_BF16) return contract_wh_t<false, 1, true>(o, H, ldH, T, WH, ldWH, target, ws, s, ra);
  return contract_wh_t<false, 1>(o, H, ldH, T, WH, ldWH, target, ws, s, ra);
}

// Second contraction with a fused epilogue.  `R` is what multiplies A^T: the ratio (KL) or A H (Frobenius).
template <int kSplit3, int kEpi, int kCG, bool kBf16>
inline int contract2_cg(DictOperands& o, int T, const float* R, int ldR, GemmParams p, cudaStream_t s) {
  constexpr int bk = kSplit3 ? kBlockK3 : kBlockK1;
  constexpr int bke = kBf16 ? 2 * bk : bk;
  CUtensorMap tmR;
  if (kBf16) EVC_TRY(make_tmap16(&tmR, o.r16.as<__nv_bfloat16>(), T, o.F, o.ldA16, bke, kC2BlockT / kCG));
  else EVC_TRY(make_tmap(&tmR, R, T, o.F, ldR, bk, kC2BlockT / kCG));
  p.M_total = o.N; p.T = T; p.K = o.F;
  p.num_m_groups = ceil_div(o.N, 128 * kC2MTiles * kCG); p.num_t_tiles = ceil_div(T, kC2BlockT); p.num_splits = 1;
  p.kblocks_total = ceil_div(o.F, bke); p.kblocks_per_split = p.kblocks_total;
  p.items_main = p.num_m_groups * p.num_t_tiles; p.splits_last = 0; p.kblocks_per_split_last = 0;
  {
    // tail balancing: the tiles of a last, less-than-half-filled round run as two half-width items each
    const int slots = num_sms() / kCG, rem = p.items_main % slots;
    static const bool allow = getenv("EVC_NO_HALF_TILES") == nullptr;
    p.half_from = (allow && p.items_main > slots && rem > 0 && 2 * rem <= slots) ? p.items_main - rem : p.items_main;
  }
  // neighbouring CTAs update neighbouring 512-byte runs of the same H rows: DRAM pages stay open
  p.m_fastest = getenv("EVC_T_FASTEST") ? 0 : 1;
  p.direct_store = getenv("EVC_DIRECT_STORE") ? 1 : 0;
  CUtensorMap tmHc = tmR, tmQc = tmR;  // the fused updates stage H (and the Frobenius numerator) through shared memory
  if (kBf16 && (kEpi == TEPI_MU_KL || kEpi == TEPI_MU_FRO)) { p.out16 = o.h16.as<__nv_bfloat16>(); p.ld_out16 = o.ldN16; }
  if (kEpi == TEPI_MU_KL || kEpi == TEPI_MU_FRO) EVC_TRY(make_tmap(&tmHc, p.out, T, o.N, p.ld_out, 128, kHChunkT, false));
  if (kEpi == TEPI_MU_FRO) EVC_TRY(make_tmap(&tmQc, p.num0, T, o.N, p.ld_out, 128, kHChunkT, false));
  ProfScope ps(2, s);
  return launch_tc<kC2MTiles, kC2BlockT, bk, kSplit3, kEpi, kCG, kBf16>(kBf16 ? o.tmA16 : o.tmA, tmR, tmHc, tmQc, p, s);
}

template <int kEpi>
inline int contract2_t(DictOperands& o, int mode, int T, const float* R, int ldR, const GemmParams& p, cudaStream_t s) {
  if (cta_group() == 2) {
    if (mode == EVC_MODE_3XTF32 && split_flavor() == 2) return contract2_cg<2, kEpi, 2, false>(o, T, R, ldR, p, s);
    if (mode == EVC_MODE_3XTF32) return contract2_cg<1, kEpi, 2, false>(o, T, R, ldR, p, s);
    if (mode == EVC_MODE_BF16) return contract2_cg<false, kEpi, 2, true>(o, T, R, ldR, p, s);
    return contract2_cg<false, kEpi, 2, false>(o, T, R, ldR, p, s);
  }
  if (mode == EVC_MODE_3XTF32 && split_flavor() == 2) return contract2_cg<2, kEpi, 1, false>(o, T, R, ldR, p, s);
  if (mode == EVC_MODE_3XTF32) return contract2_cg<1, kEpi, 1, false>(o, T, R, ldR, p, s);
  if (mode == EVC_MODE_BF16) return contract2_cg<false, kEpi, 1, true>(o, T, R, ldR, p, s);
  return contract2_cg<false, kEpi, 1, false>(o, T, R, ldR, p, s);
}

inline int update_kl(DictOperands& o, int mode, const float* X, int ldX, int T, const float* WH, int ldWH, float* R,
                     int ldR, float* H, int ldH, const float* colsum, float lam, float eps,
                     const unsigned char* row_active, DevBuf* ws, cudaStream_t s, bool ratio_done = false) {
  __nv_bfloat16* R16 = (mode == EVC_MODE_BF16) ? o.r16.as<__nv_bfloat16>() : nullptr;
  if (!ratio_done) EVC_TRY(launch_ratio(X, ldX, WH, ldWH, R, ldR, T, o.F, eps, 0, s, R16, o.ldA16));
  GemmParams p{};
  p.out = H; p.ld_out = ldH;
  p.colsum = colsum; p.lam = lam; p.eps = eps; p.row_active = row_active;
  if (o.n_left > 0) {
    const C1Plan pl = plan_c1(o.F_main, o.N, T, bk_elems(mode));
    p.left_a = o.AT + (size_t)o.F_main * o.ldN; p.left_lda = o.ldN; p.n_left = o.n_left;
    p.left_out = ws->as<float>() + ws_left_offset(pl, T); p.left_ld = left_ld(T); p.left_rows = left_rows(o);
    o.left_valid = true;  // (stream order: the partials are complete before the next contraction 1 reads them)
  }
  return contract2_t<TEPI_MU_KL>(o, mode, T, R, ldR, p, s);
}

inline int update_fro(DictOperands& o, int mode, int T, const float* WH, int ldWH, float* R, int ldR, float* H, int ldH,
                      const float* num0, float lam, float eps, const unsigned char* row_active, DevBuf* ws,
                      cudaStream_t s) {
  EVC_TRY(launch_ratio(nullptr, 0, WH, ldWH, R, ldR, T, o.F, eps, 1, s,
                       (mode == EVC_MODE_BF16) ? o.r16.as<__nv_bfloat16>() : nullptr, o.ldA16));
  GemmParams p{};
  p.out = H; p.ld_out = ldH;
  p.num0 = num0; p.lam = lam; p.eps = eps; p.row_active = row_active;
  if (o.n_left > 0) {
    const C1Plan pl = plan_c1(o.F_main, o.N, T, bk_elems(mode));
    p.left_a = o.AT + (size_t)o.F_main * o.ldN; p.left_lda = o.ldN; p.n_left = o.n_left;
    p.left_out = ws->as<float>() + ws_left_offset(pl, T); p.left_ld = left_ld(T); p.left_rows = left_rows(o);
    o.left_valid = true;
  }
  return contract2_t<TEPI_MU_FRO>(o, mode, T, R, ldR, p, s);
}

// NUM0 (T, ldH) = X A^T : the second contraction with a plain store ([split=0][t][n] layout == (T, ldH)).
inline int frob_numerator(DictOperands& o, int mode, const float* X, int ldX, int T, float* R, int ldR, float* num0,
                          int ldH, DevBuf* ws, cudaStream_t s) {
  // stage X into the zero-padded K-operand buffer (copy mode of the ratio kernel with WH := X)
  EVC_TRY(launch_ratio(nullptr, 0, X, ldX, R, ldR, T, o.F, 0.f, 1, s,
                       (mode == EVC_MODE_BF16) ? o.r16.as<__nv_bfloat16>() : nullptr, o.ldA16));
  GemmParams p{};
  p.out = num0; p.ld_out = ldH;
  return contract2_t<TEPI_PARTIAL>(o, mode, T, R, ldR, p, s);
}

}  // namespace tc
}  // namespace evc
