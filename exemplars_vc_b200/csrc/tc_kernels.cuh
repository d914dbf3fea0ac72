// Tensor-core kernels (sm_100a): both contractions of the multiplicative update as tcgen05.mma GEMMs on CTA pairs
// (cta_group::2) with TMEM accumulators, operands staged by TMA through an mbarrier ring.
//
// Orientation (shared by every kernel here): the DICTIONARY dimension is the MMA M dimension (TMEM lanes), the
// FRAME dimension is the MMA N dimension (TMEM columns), both operands are K-major:
//
//   contraction 1 (sklearn _nmf.py:554, WH = W@A):     D[f, t] = sum_n AT[f, n] * H[t, n]     K = N
//   contraction 2 (sklearn _nmf.py:585, R@A.T):        D[n, t] = sum_f A [n, f] * R[t, f]     K = F
//   conversion    (04_align_n_nmf.py:391, H.T@B):      D[f, t] = sum_n BT[f, n] * H[t, n]     K = N
//
// so an epilogue thread owns one dictionary row (one TMEM lane) and walks frames; with H, WH and R stored
// frame-major (T, ld) a warp touches 32 consecutive floats per frame: coalesced.
//
// Arithmetic of a product (template argument kPrec):
//
//   PREC_SPLIT  (EVC_MODE_3XTF32, the fp32-accurate mode): every operand element is x = x1 + x2 with
//               x1 = bf16_rn(x), x2 = bf16_rn(x - x1) (16 mantissa bits, round-to-nearest, so the representation
//               error is <= 2^-17 |x| and unbiased) and  D += M2*N1 + M1*N2 + M1*N1  on kind::f16 (bf16 operands,
//               fp32 accumulate in TMEM), dropping M2*N2 (<= 2^-18).  Three bf16 MMAs cost 1.5 tf32 MMA-times per
//               product (the previous tf32 hi*hi + two bf16 cross terms cost 2, three tf32 passes 3).  The planes
//               (x1 | x2) are made ONCE where the data is produced -- the dictionary at evc_dict_create, the ratio
//               R by the kernel that forms it -- and cost the same 4 bytes per element in HBM, L2 and shared memory
//               as the fp32 value, so contraction 2 is a pure TMA -> MMA pipeline.  Only the activations H (an
//               fp32 master copy that the fused update rewrites every iteration) are split in shared memory, by
//               the otherwise idle epilogue warps of contraction 1 (kSplitN).
//   PREC_TF32   (EVC_MODE_TF32): one kind::tf32 MMA per product on the raw fp32 tiles.
//   PREC_BF16   (EVC_MODE_BF16): one kind::f16 MMA per product on bf16 copies (dictionary converted once, the ratio
//               emitted in bf16, a bf16 shadow of H kept by the fused update).
//
// A K-block is one 128-byte swizzle row: 32 elements in PREC_SPLIT (their 32 hi and 32 lo bf16 values side by side:
// "interleaved planes", one TMA box with full 128-byte row segments brings both), 32 fp32 or 64 bf16
// (128-byte rows) in the fast modes; one MMA advances 32 bytes of K.
//
// CTA pairs: M = 256 (128 rows per CTA), N = 256 with the frame operand split in halves between the two CTAs'
// shared memories.  Both CTAs issue TMA loads for their halves with cta_group::2, crediting the bytes to the LEADER's
// "full" barrier; the leader's single MMA thread waits on it, issues, and tcgen05.commit...multicast frees the ring
// slot in both CTAs.
#pragma once
#include "evc_common.cuh"
#include "umma.cuh"
#include <cuda_bf16.h>
#include <algorithm>
#include <cstdlib>

namespace evc {
namespace tc {

using namespace umma;

// Row pitch (elements) of every K-major operand: a multiple of 128 bytes, so that no 64- or 128-byte row segment a
// TMA box fetches straddles a 128-byte L2 line (a 1056-byte pitch cost contraction 2 25 % extra L2->SM traffic:
// profiles/r2_pitch_before_fix_ncu.txt).
inline int k_pitch(int F) { return round_up(F, 32); }
inline int k_pitch16(int F) { return round_up(F, 64); }
// ... and of an operand with interleaved hi / lo planes (fp32-accurate mode): per 32 K elements, 32 hi then 32 lo bf16
// values -- one 128-byte row segment per K-block; element k of plane p sits at column (k / 32) * 64 + p * 32 + k % 32.
inline int k_pitch_i(int K) { return 2 * round_up(K, 32); }
__host__ __device__ __forceinline__ int col_i(int k) { return ((k >> 5) << 6) + (k & 31); }

enum TcEpilogue { TEPI_PARTIAL = 0, TEPI_MU_KL = 1, TEPI_MU_FRO = 2 };
enum TcPrec { PREC_SPLIT = 0, PREC_TF32 = 1, PREC_BF16 = 2 };

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
// x -> (x1, x2) = (bf16_rn(x), bf16_rn(x - x1)), two values at a time
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = pack_bf16(a - hf.x, b - hf.y);
}
// The K operand of contraction 2 in the mode's format: fp32 R (pitch ldr, pad columns zeroed), bf16 R16, or the
// two bf16 planes R12 (plane 1 at + plane elements) of the fp32-accurate split.
struct ROut {
  float* R; int ldr;
  __nv_bfloat16* R16; int ldr16;
  __nv_bfloat16* R12; int ldr12, k12;  // interleaved planes: pitch, logical columns (F rounded up to whole K-blocks)
  __host__ __device__ int cols() const { return R12 ? k12 : (R16 ? ldr16 : (R ? ldr : 0)); }
};
__device__ __forceinline__ void store_r(const ROut& o, int t, int f, float r) {
  if (o.R && f < o.ldr) o.R[(size_t)t * o.ldr + f] = r;
  if (o.R16 && f < o.ldr16) o.R16[(size_t)t * o.ldr16 + f] = __float2bfloat16_rn(r);
  if (o.R12 && f < o.k12) {
    const __nv_bfloat16 r1 = __float2bfloat16_rn(r);
    __nv_bfloat16* q = o.R12 + (size_t)t * o.ldr12 + col_i(f);
    q[0] = r1;
    q[32] = __float2bfloat16_rn(r - __bfloat162float(r1));
  }
}
// four consecutive columns f..f+3 (f a multiple of 4, every pitch a multiple of 4): 8- / 16-byte stores
__device__ __forceinline__ void store_r4(const ROut& o, int t, int f, float4 r) {
  if (o.R && f < o.ldr) *reinterpret_cast<float4*>(o.R + (size_t)t * o.ldr + f) = r;
  if (o.R16 && f < o.ldr16) {
    uint2 v;
    v.x = pack_bf16(r.x, r.y); v.y = pack_bf16(r.z, r.w);
    *reinterpret_cast<uint2*>(o.R16 + (size_t)t * o.ldr16 + f) = v;
  }
  if (o.R12 && f < o.k12) {
    uint2 hi, lo;
    split2(r.x, r.y, hi.x, lo.x);
    split2(r.z, r.w, hi.y, lo.y);
    __nv_bfloat16* q = o.R12 + (size_t)t * o.ldr12 + col_i(f);
    *reinterpret_cast<uint2*>(q) = hi;
    *reinterpret_cast<uint2*>(q + 32) = lo;
  }
}

constexpr int kRedCols = 256, kRedThreads = 64, kRedMaxBatch = 32, kRedMaxTiles = 1024;

// The split-K sum of contraction 1 (and the ratio R = X / max(A H, eps) formed from it) done by the contraction's OWN
// CTAs -- the first GEMM's epilogue, no second launch (sklearn _nmf.py:554-571).  Every CTA stores its partial tile,
// arrives on the frame tile's counter and, once all `contributors` CTAs of that frame tile have arrived, sums a
// 1/contributors share of the tile's (frame, 256-column) units over the splits in fixed order (bit-identical to
// reduce_partials_kernel), staging the partials through the now idle operand ring with bulk copies.  All work items
// are resident at once whenever K is split (plan_c1), so the wait cannot deadlock.  counter[2*tile] counts arrivals,
// counter[2*tile+1] the CTAs that have passed the wait; the last of those resets both, so nothing is reset by the host.
struct FusedReduce {
  int enabled;        // 1: tile barrier + sum over the splits; 2: K not split, the epilogue writes A*H / the ratio from TMEM
  unsigned int* counter;
  int contributors;   // CTAs that store partials of one frame tile
  int nslots;         // staging slots (2: the next batch's copies overlap the current sum)
  int nb;             // frames per batch (1, 2 or 4)
  int slot_floats;    // floats per slot: [chunk][S_max][nb][kRedCols] partials | [n_left][nb][left_rows] leftover-row partials
  int S, S_last, f_last, F;
  float* WH; int ldwh;
  const float* L; int left_rows, left_ld;   // leftover-row partials of the fused update (left_rows = 0: none)
  const float* Lp; int lp_splits;            // ... or contraction 1's own per-split sums [split][row][frame] (fp32-accurate mode)
  int n_left;
  const float* X; int ldx;                  // X != nullptr: emit the ratio too
  int cols;                                 // columns to write per frame: max(ldwh, ratio pitch)
  ROut ro;
};

struct GemmParams {
  int M_total;  // dictionary-side rows: F for contraction 1 / conversion, N for contraction 2
  int T;        // frames
  int K;        // reduction length
  int num_m_groups, num_t_tiles, num_splits, kblocks_per_split, kblocks_total;
  // The last m_group may hold fewer 256-row sub-tiles than the others; it then gets fewer, longer K splits so every
  // CTA pair carries the same number of MMAs.  items_main = work items of the other groups; 0 splits_last means
  // "no special last group".
  int items_main, splits_last, kblocks_per_split_last;
  // Tail balancing of contraction 2: the tiles with index >= half_from (those of a last, partly filled round) are cut
  // into tail_parts narrower items of whole 32-frame chunks each (plan_tail), so that round costs a fraction of a tile
  // time.  half_from = items_main: off.
  int half_from, tail_parts;
  int m_fastest;  // work-item order: 1 = consecutive CTA pairs take consecutive dictionary-row groups of one frame tile
  float* out;    // PARTIAL: [split][t][m] with pitch ld_out;  MU_*: the activations H (T, ld_out)
  int ld_out;
  const float* colsum;
  float lam, eps;
  const unsigned char* row_active;
  // Dictionary rows F_main..F-1 that contraction 1 does not run through the tensor cores (F = 513 = 4*128 + 1):
  // the fused update accumulates their share of the NEXT A*H, sum_n h_new[t,n] * A[n, F_main+l], per warp.
  const float* left_a;  // (n_left, left_lda): rows F_main.. of the transposed dictionary (contiguous in n)
  int left_lda, n_left;
  float* left_out;      // [l][t][row] partial sums, row = 4*(128-exemplar block) + lane quarter
  int left_ld, left_rows;  // frames pitch, rows pitch
  int out_keep_l2;      // PARTIAL: store with the L2 evict-last hint (split-K partials: the reduction reads them next)
  int debug_flags;      // -DEVC_INSTRUMENT builds only (tools/flag_sweep.sh); always 0 and never read otherwise
  FusedReduce red;      // PARTIAL: split-K sum (+ ratio) inside this launch
  int direct_store;     // MU_*: the updated activations leave by per-lane global stores instead of the staged TMA store
};

// Timing experiments (results are garbage when a flag is set).  The default build compiles every test to `false`.
//   1 no plane split (contraction 1)   2 no MMA issue   4 no operand TMA   8 no update arithmetic in the epilogue
//  16 no H chunk loads / stores       32 no leftover-row partials        64 no TMEM loads
#ifdef EVC_INSTRUMENT
#define EVC_DBG(p, bit) (((p).debug_flags & (bit)) != 0)
// per-role wait / work cycle counters, printed by the first two CTAs when flag 128 is set
#define EVC_CLK_DECL(...) long long __VA_ARGS__
#define EVC_CLK(v) v = clock64()
#define EVC_CLK_ADD(acc, since) acc += clock64() - (since)
#define EVC_CLK_PRINT(p, ...) do { if (((p).debug_flags & 128) && blockIdx.x < 2) printf(__VA_ARGS__); } while (0)
#else
#define EVC_DBG(p, bit) false
#define EVC_CLK_DECL(...)
#define EVC_CLK(v)
#define EVC_CLK_ADD(acc, since)
#define EVC_CLK_PRINT(p, ...)
#endif

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
    // narrow tail items of contraction 2 (no split-K there): tile = half_from + h / parts, piece = h % parts; the
    // tile's 32-frame chunks are dealt to the pieces as evenly as possible (256 frames in 3 pieces: 96 + 96 + 64)
    const int parts = p.tail_parts, h = item - p.half_from, part = h % parts;
    item = p.half_from + h / parts;
    const int chunks = block_t / 32, base = chunks / parts, extra = chunks % parts;
    w.t_cols = (base + (part < extra ? 1 : 0)) * 32;
    w.t_off = (part * base + (part < extra ? part : extra)) * 32;
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

__host__ __device__ __forceinline__ int num_items_of(const GemmParams& p) {
  return p.items_main + p.splits_last * p.num_t_tiles + (p.items_main - p.half_from) * (p.tail_parts - 1);
}
// The `rem` tiles of a last, partly filled round are cut into as many pieces (at most 4: 64 frames) as still fit ONE
// extra round of the `slots` resident CTA pairs: 316 tiles on 74 pairs -> 20 tail tiles x 3 pieces, 0.375 instead of
// 1 (whole tiles) or 0.5 (halves) tile times for the round.
inline void plan_tail(GemmParams& p, int slots, int block_t, bool allow, int max_parts = 4) {
  p.half_from = p.items_main;
  p.tail_parts = 1;
  const int rem = slots > 0 ? p.items_main % slots : 0;
  if (!allow || p.items_main <= slots || rem <= 0) return;
  const int parts = std::min(std::min(max_parts, block_t / 32), slots / rem);
  if (parts >= 2) { p.half_from = p.items_main - rem; p.tail_parts = parts; }
}

constexpr int kCG = 2;          // CTAs per MMA (cta_group::2)
constexpr int kEpiWarps = 8;    // two warps per TMEM lane quarter, interleaved over 32-column chunks
constexpr int kSmemBudget = 227 * 1024 - 2048;

// The fused-update epilogue stages H through shared memory in [32 frames x 128 exemplars] chunks moved by TMA
// (loads prefetched by a loader warp, stores issued by a storer warp): per-lane 128-byte global accesses
// from the epilogue warps were limited by the SM's outstanding-miss capacity, bulk copies are not.
#ifndef EVC_H_BUFS
#define EVC_H_BUFS 4
#endif
#ifndef EVC_MAX_STAGES
#define EVC_MAX_STAGES 8
#endif
constexpr int kHChunkT = 32, kHBufBytes = kHChunkT * 128 * 4, kHBufs = EVC_H_BUFS;

template <int kMTiles, int kBlockT, int kPrec, bool kSplitN, bool kStageH, bool kStageQ>
struct TileCfg {
  // One K-block = one 128-byte swizzle row of every operand that TMA brings: 32 fp32 (tf32), 64 bf16 (bf16), or -- in
  // the fp32-accurate split mode -- the 32 hi and the 32 lo bf16 values of 32 K elements side by side ("interleaved
  // planes": [hi 0..31 | lo 0..31]), so one box with full 128-byte row segments brings both planes (separate 64-byte
  // rows per plane capped what an SM ingests at ~55 B/clk against ~85 with 128-byte rows: tools/probe/tma_issue_probe.cu).
  // Only the frame operand that contraction 1 derives in shared memory (kSplitN) keeps two 64-byte-row planes.
  static constexpr int kPlanes = (kPrec == PREC_SPLIT) ? 2 : 1;
  static constexpr int kElemBytes = (kPrec == PREC_TF32) ? 4 : 2;
  static constexpr int kKE = (kPrec == PREC_BF16) ? 64 : 32;             // K elements per K-block: 32 / 32 / 64
  static constexpr int kKStep = 32 / kElemBytes;                         // K elements per MMA: 16 / 8 / 16
  static constexpr int kBoxCols = 128 / kElemBytes;                      // tensor-map columns of one K-block: 64 / 32 / 64
  static constexpr int kMRowBytes = 128;
  static constexpr int kMSubBytes = 128 * kMRowBytes;                    // this CTA's 128 rows of one sub-tile (both planes)
  static constexpr int kMBytes = kMTiles * kMSubBytes;
  static constexpr int kMPlaneOff = 64;                                  // split mode: lo values start 64 bytes into the row
  static constexpr int kNRows = kBlockT / kCG;                           // frame rows this CTA holds
  static constexpr int kNRowBytes = kSplitN ? 64 : 128;
  static constexpr int kNPlaneBytes = kNRows * kNRowBytes;               // kSplitN: one derived plane
  static constexpr int kNBytes = kSplitN ? 2 * kNPlaneBytes : kNRows * 128;
  static constexpr int kNPlaneOff = kSplitN ? kNPlaneBytes : 64;         // where the lo plane starts (split mode)
  // kSplitN: the fp32 frame tile TMA brings; the two planes are derived IN PLACE (same 16 KB: every thread reads its
  // quarter rows, a barrier, then writes), which buys a fourth ring stage -- with three the TMA -> split -> MMA chain
  // left the tensor pipe idle 16 % of contraction 1's main loop
  static constexpr int kRawBytes = kSplitN ? kNRows * 128 : 0;
  static_assert(!kSplitN || kRawBytes == kNBytes, "in-place split: the planes take exactly the raw tile's bytes");
  // kSplitN: this K-block of the dictionary rows that stay off the tensor cores (F_main.., at most 8): [8][32] fp32
  static constexpr int kLeftBytes = kSplitN ? 8 * 128 : 0;
  static constexpr int kOffN = kMBytes, kOffRaw = kMBytes, kOffLeft = kMBytes + kNBytes;
  static constexpr int kStageBytes = kMBytes + kNBytes + kLeftBytes;
  // bytes per CTA per stage that TMA credits to the leader's "full" barrier
  static constexpr int kTxBytes = kMBytes + (kSplitN ? 0 : kNBytes);
  // arrivals on the leader's "full" barrier: its producer's expect_tx + (kSplitN) every split warp of both CTAs
  static constexpr int kFullArrivals = 1 + (kSplitN ? kEpiWarps * kCG : 0);
  // Frobenius also stages the cached numerator X A^T next to H: half as many, twice as large buffers
  static constexpr int kHBufsUsed = kStageQ ? kHBufs / 2 : kHBufs;
  static constexpr int kHBufStride = kStageQ ? 2 * kHBufBytes : kHBufBytes;
  static constexpr int kHBytes = kStageH ? kHBufs * kHBufBytes : 0;
  // PREC_BF16: the bf16 shadow of the updated chunk is staged next to it and leaves by TMA too (2-byte per-lane global
  // stores from the epilogue warps were what bounded the fast mode: profiles/r1_bf16_ncu_summary.txt)
  static constexpr bool kShadow = kStageH && (kPrec == PREC_BF16);
  static constexpr int kSBufBytes = kHChunkT * 128 * 2;
  static constexpr int kSBytes = kShadow ? kHBufs * kSBufBytes : 0;
  static constexpr int kStagesRaw = (kSmemBudget - kHBytes - kSBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > EVC_MAX_STAGES ? EVC_MAX_STAGES : kStagesRaw;
  static constexpr int kAccCols = kMTiles * kBlockT;
  static constexpr int kAccStages = (512 / kAccCols) >= 2 ? 2 : 1;
  static constexpr int kLoaderWarp = 2 + kEpiWarps;  // H chunk loader, then storer
  static constexpr int kThreads = (kLoaderWarp + (kStageH ? 2 : 0)) * 32;
  static constexpr int kOffH = kStages * kStageBytes;
  static constexpr int kOffS = kOffH + kHBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + kHBytes + kSBytes + 1024;  // + slack to align the ring to 1024 B
  static constexpr int kRowsPerSub = 128 * kCG;                   // dictionary rows of one MMA (M = 256)
  static_assert(kStages >= 2, "tile does not fit twice in shared memory");
  static_assert(kAccCols <= 512, "accumulators exceed TMEM");
  static_assert(kBlockT % 32 == 0 && kBlockT >= 32 && kBlockT <= 256, "bad frame tile");
  static_assert(!kSplitN || kPrec == PREC_SPLIT, "only the split mode derives planes in shared memory");
  static_assert(!kSplitN || kAccStages == 1, "the epilogue warps split during the main loop: one accumulator stage");
  static_assert(!kSplitN || (kNRows * 4) % (kEpiWarps * 32) == 0, "every split thread gets whole units");
};

// a / d with the reciprocal r = rn(1/d) of the (per-exemplar, constant) denominator at hand: q0 = a*r corrected by
// one FMA residual step, the fast path of an IEEE division -- the correctly rounded quotient whenever it is a
// normal number (sklearn does `numerator /= denominator; W *= numerator`, _nmf.py:617-624).
__device__ __forceinline__ float quotient(float a, float d, float r) {
  const float q0 = a * r;
  const float q = fmaf(fmaf(-q0, d, a), r, q0);
  return (q == q) ? q : q0;  // a*r overflowed or a is inf: keep the uncorrected value instead of inf - inf
}

// barrier among `threads` threads of the CTA (ids 1..: 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// One frame tile of a ring stage: raw = [kRows x 32 fp32] as TMA wrote it with the 128B swizzle (16-byte chunk c
// of row r sits at chunk c ^ (r & 7)); p1 / p2 = [kRows x 32 bf16] in the 64B-swizzle layout the MMA descriptors
// expect (chunk c of row r at chunk c ^ ((r >> 1) & 3)).  A unit is a quarter row: 8 floats in, 16 + 16 bytes out;
// the 256 threads of the 8 epilogue warps take kRows*4/256 units each.  Bank-conflict free: a quarter-warp reads /
// writes eight distinct 16-byte columns of two adjacent rows.
//
// With n_left > 0 the same registers also feed the dictionary rows that stay off the tensor cores (the Nyquist bin
// of a 513-bin spectrum): lacc[q][l] += sum_e x[8c + e] * aleft[l][8c + e] for this thread's quarter rows -- the
// activations pass through these registers anyway, and these warps have time to spare (they wait for TMA two thirds
// of the main loop).  aleft = [8][32] fp32, this K-block of rows F_main.. of the transposed dictionary.
template <int kRows>
__device__ __forceinline__ void split_planes(const uint8_t* raw, uint8_t* p1, uint8_t* p2, int tid, const float* aleft,
                                             int n_left, float (&lacc)[kRows * 4 / (kEpiWarps * 32)][8]) {
  constexpr int kPer = kRows * 4 / (kEpiWarps * 32);
  float4 x[kPer][2];
#pragma unroll
  for (int q = 0; q < kPer; ++q) {
    const int u = q * (kEpiWarps * 32) + tid, row = u >> 2, c = u & 3, sw = row & 7;
    const uint8_t* r = raw + row * 128;
    x[q][0] = *reinterpret_cast<const float4*>(r + (((2 * c) ^ sw) << 4));
    x[q][1] = *reinterpret_cast<const float4*>(r + (((2 * c + 1) ^ sw) << 4));
  }
  if (n_left > 0) {
    const int c = tid & 3;  // (kEpiWarps * 32 is a multiple of 4: the quarter index does not depend on q)
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      if (l >= n_left) break;
      const float4 a0 = *reinterpret_cast<const float4*>(aleft + l * 32 + 8 * c);
      const float4 a1 = *reinterpret_cast<const float4*>(aleft + l * 32 + 8 * c + 4);
#pragma unroll
      for (int q = 0; q < kPer; ++q) {
        float a = lacc[q][l];
        a = fmaf(x[q][0].x, a0.x, a); a = fmaf(x[q][0].y, a0.y, a); a = fmaf(x[q][0].z, a0.z, a); a = fmaf(x[q][0].w, a0.w, a);
        a = fmaf(x[q][1].x, a1.x, a); a = fmaf(x[q][1].y, a1.y, a); a = fmaf(x[q][1].z, a1.z, a); a = fmaf(x[q][1].w, a1.w, a);
        lacc[q][l] = a;
      }
    }
  }
  named_bar_sync(1, kEpiWarps * 32);  // in place: every quarter row is in registers before any plane is written
#pragma unroll
  for (int q = 0; q < kPer; ++q) {
    const int u = q * (kEpiWarps * 32) + tid, row = u >> 2, c = u & 3;
    const int off = row * 64 + ((c ^ ((row >> 1) & 3)) << 4);
    uint4 hi, lo;
    split2(x[q][0].x, x[q][0].y, hi.x, lo.x);
    split2(x[q][0].z, x[q][0].w, hi.y, lo.y);
    split2(x[q][1].x, x[q][1].y, hi.z, lo.z);
    split2(x[q][1].z, x[q][1].w, hi.w, lo.w);
    *reinterpret_cast<uint4*>(p1 + off) = hi;
    *reinterpret_cast<uint4*>(p2 + off) = lo;
  }
}

// ---- split-K sum + ratio inside contraction 1 (FusedReduce) ------------------------------------------------------
__device__ __forceinline__ unsigned int atom_add_release_gpu(unsigned int* p, unsigned int v) {
  unsigned int old;
  asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// generic-proxy accesses to global memory <-> the async proxy (bulk copies) that reads the same bytes next
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// Called by the 256 threads of the 8 epilogue warps (tid = 0..255) of every CTA after its partial tile is stored.
// `slot` = this CTA's index among the `contributors` of frame tile `t_tile`; `stage` = the idle operand ring;
// `bars` = 2 mbarriers (count 1), `sh` = 4 x 16 floats of scratch; tmP = the partials as a 3-D tensor
// (column, frame, split) with boxes of 256 columns x nb frames x S_max splits.
//
// The frame tile is cut into batches of `nb` frames; batch b belongs to contributor b % contributors.  One thread
// brings a whole batch -- every split's rows of those frames -- with ONE box load per 256 columns (per-(split, frame)
// 1 KB bulk copies were bound by the copy issue rate: 390 per CTA), double buffered when two batches fit the ring.
template <int kBlockT>
__device__ __forceinline__ void fused_reduce_phase(const GemmParams& p, const CUtensorMap* tmP, int t_tile, int slot,
                                                   float* stage, uint64_t* bars, float* sh, int tid) {
  const FusedReduce& r = p.red;
  EVC_CLK_DECL(c_0 = 0, c_1 = 0, c_2 = 0, c_3 = 0, c_a = 0, c_w = 0, c_left = 0, c_main = 0, c_tail = 0, c_bar = 0, c_iss = 0);
  EVC_CLK(c_0);
  // this CTA's batches of the tile: (2) below, after the tile barrier (1)
  const int g = tid >> 6, t64 = tid & 63;
  const int t_base = t_tile * kBlockT;
  const int frames = min(kBlockT, p.T - t_base);
  const int nb = r.nb, S_max = max(r.S, r.S_last);
  const int n_batches = frames > 0 ? (frames + nb - 1) / nb : 0;
  const int ldp = p.ld_out, F_main = p.M_total;
  const int n_chunks = (ldp + kRedCols - 1) / kRedCols;           // 256-column boxes per partial row
  const int chunk_floats = S_max * nb * kRedCols;                   // one box in shared memory: [split][frame][256]
  const int left_floats = r.left_rows > 0 ? r.n_left * nb * r.left_rows : 0;
  const uint32_t bar0 = smem_u32(&bars[0]);
  float* s_left = sh + 16 * g;       // [8] leftover sums of frame g of the batch | [2] per-warp partials
  float* s_warp = s_left + 8;
  const bool issuer = (tid == kEpiWarps * 32 - 1);  // a thread whose 64-thread group has the least other work
  auto issue = [&](int b, int s) {   // one thread: box loads of batch b into slot s
    const int t = t_base + b * nb;
    float* dst = stage + (size_t)s * r.slot_floats;
    const uint32_t bar = bar0 + 8u * s;
    mbar_arrive_expect_tx(bar, (uint32_t)((n_chunks * chunk_floats + left_floats) * 4));
    for (int c = 0; c < n_chunks; ++c)
      tma_load_3d(smem_u32(dst + (size_t)c * chunk_floats), tmP, c * kRedCols, t, 0, bar, kEvictFirst);
    if (left_floats > 0)  // [l][frame][row]: the frames of a batch are contiguous in the leftover partials
      for (int l = 0; l < r.n_left; ++l)
        bulk_load_1d(smem_u32(dst + (size_t)n_chunks * chunk_floats + (size_t)l * nb * r.left_rows),
                     r.L + ((size_t)l * r.left_ld + t) * r.left_rows, (uint32_t)(nb * r.left_rows) * 4u, bar, kEvictFirst);
  };
  // X entries of this thread's first float4 of a batch and of its first tail column (prefetched one batch ahead)
  auto load_x4 = [&](int t, int f) {
    float4 x4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!r.X || t >= p.T) return x4;
    if (f + 3 < r.F && ((r.ldx & 3) == 0) && ((((uintptr_t)r.X) & 15) == 0))
      return __ldg(reinterpret_cast<const float4*>(r.X + (size_t)t * r.ldx + f));
    float* xs = reinterpret_cast<float*>(&x4);
#pragma unroll
    for (int j = 0; j < 4; ++j) if (f + j < r.F) xs[j] = r.X[(size_t)t * r.ldx + f + j];
    return x4;
  };
  float4 x_main = make_float4(0.f, 0.f, 0.f, 0.f);
  float x_tail = 0.f;
  auto prefetch_x = [&](int bb) {
    if (!r.X || bb >= n_batches) return;
    const int c4 = ldp >> 2, tb = t_base + bb * nb;
    if (tid < nb * c4) {
      const int fr = tid / c4, f = (tid - fr * c4) * 4;
      if (f < F_main) x_main = load_x4(tb + fr, f);
    }
    if (g < nb && tb + g < p.T && F_main + t64 < r.F && (r.left_rows > 0 || r.Lp)) x_tail = r.X[(size_t)(tb + g) * r.ldx + F_main + t64];
  };
  int b = slot, k = 0;
  uint32_t phases = 0;
  prefetch_x(b);  // (the inputs do not depend on the other CTAs: in flight across the tile barrier)

  // (1) publish this CTA's partials, wait for the other contributors of the frame tile
  __threadfence();
  fence_proxy_async_global();
  named_bar_sync(1, kEpiWarps * 32);
  EVC_CLK(c_1);
  if (tid == 0) {
    unsigned int* ctr = r.counter + 2 * t_tile;
    atom_add_release_gpu(ctr, 1u);
    const long long t0 = clock64();
    while (ld_acquire_gpu(ctr) < (unsigned int)r.contributors) {
      if (clock64() - t0 > 4000000000ll) {
        printf("evc: split-K tile barrier timed out (block %d, frame tile %d: %u of %d arrived)\n", (int)blockIdx.x,
               t_tile, ld_acquire_gpu(ctr), r.contributors);
        __trap();
      }
    }
    // everybody is past the point of incrementing; the last CTA through resets both counters for the next launch
    if (atom_add_release_gpu(ctr + 1, 1u) == (unsigned int)r.contributors - 1u) {
      ctr[0] = 0u;
      ctr[1] = 0u;
    }
    __threadfence();
    fence_proxy_async_global();
  }
  named_bar_sync(1, kEpiWarps * 32);
  EVC_CLK(c_2);

  if (issuer) {
    if (b < n_batches) issue(b, 0);
    if (r.nslots == 2 && b + r.contributors < n_batches) issue(b + r.contributors, 1);
  }
  for (; b < n_batches; b += r.contributors, ++k) {
    const int s = (r.nslots == 2) ? (k & 1) : 0;
    const float* st = stage + (size_t)s * r.slot_floats;
    const int t_b = t_base + b * nb;
    EVC_CLK(c_a);
    mbar_wait(bar0 + 8u * s, (phases >> s) & 1u);
    EVC_CLK_ADD(c_w, c_a);
    phases ^= 1u << s;
    EVC_CLK(c_a);
    // leftover rows (the Nyquist bin): 64-thread group g sums frame g of the batch, in reduce_partials_kernel's order
    if (left_floats > 0 && g < nb) {
      const float* lst = st + (size_t)n_chunks * chunk_floats;
      for (int l = 0; l < r.n_left; ++l) {
        const float* row = lst + ((size_t)l * nb + g) * r.left_rows;
        float a = 0.f;
        for (int q0 = t64; q0 < r.left_rows; q0 += 8 * kRedThreads) {  // loads first, adds in order (same sum)
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = (q0 + e * kRedThreads < r.left_rows) ? row[q0 + e * kRedThreads] : 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) if (q0 + e * kRedThreads < r.left_rows) a += v[e];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if ((t64 & 31) == 0) s_warp[t64 >> 5] = a;
        named_bar_sync(2 + g, kRedThreads);
        if (t64 == 0) s_left[l] = s_warp[0] + s_warp[1];
        named_bar_sync(2 + g, kRedThreads);
      }
    }
    EVC_CLK_ADD(c_left, c_a);
    EVC_CLK(c_a);
    // tensor-core columns: float4 per thread, splits summed in order (deterministic)
    const int c4_per_frame = ldp >> 2;
    for (int idx = tid, it = 0; idx < nb * c4_per_frame; idx += kEpiWarps * 32, ++it) {
      const int fr = idx / c4_per_frame, f = (idx - fr * c4_per_frame) * 4, t = t_b + fr;
      if (t >= p.T || f >= F_main) continue;
      const int c = f / kRedCols;
      const int n = (f >= r.f_last) ? r.S_last : r.S;
      const float4 x4 = (it == 0) ? x_main : load_x4(t, f);
      const float* src = st + (size_t)c * chunk_floats + (size_t)fr * kRedCols + (f - c * kRedCols);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const size_t kstride = (size_t)nb * kRedCols;
      for (int k0 = 0; k0 < n; k0 += 6) {  // six splits' loads in flight, added in split order
        float4 v[6];
#pragma unroll
        for (int e = 0; e < 6; ++e)
          v[e] = (k0 + e < n) ? *reinterpret_cast<const float4*>(src + (size_t)(k0 + e) * kstride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int e = 0; e < 6; ++e)
          if (k0 + e < n) { acc.x += v[e].x; acc.y += v[e].y; acc.z += v[e].z; acc.w += v[e].w; }
      }
      if (f + 3 < F_main) {  // whole group inside the tensor-core rows: vector stores
        if (f < r.ldwh) *reinterpret_cast<float4*>(r.WH + (size_t)t * r.ldwh + f) = acc;
        if (r.X)
          store_r4(r.ro, t, f, make_float4(__fdiv_rn(x4.x, fmaxf(acc.x, p.eps)), __fdiv_rn(x4.y, fmaxf(acc.y, p.eps)),
                                           __fdiv_rn(x4.z, fmaxf(acc.z, p.eps)), __fdiv_rn(x4.w, fmaxf(acc.w, p.eps))));
      } else {
        const float sv[4] = {acc.x, acc.y, acc.z, acc.w};
        const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (f + j >= F_main) break;
          if (f + j < r.ldwh) r.WH[(size_t)t * r.ldwh + f + j] = sv[j];
          if (r.X) store_r(r.ro, t, f + j, (f + j < r.F) ? __fdiv_rn(xv[j], fmaxf(sv[j], p.eps)) : 0.f);
        }
      }
    }
    EVC_CLK_ADD(c_main, c_a);
    EVC_CLK(c_a);
    const float x_tail_now = x_tail;
    // this thread's X entries of the CTA's next batch, fetched under the rest of this batch and the next wait
    prefetch_x(b + r.contributors);
    // columns past the tensor-core rows: the leftover rows, then zero padding
    if (g < nb && t_b + g < p.T) {
      const int t = t_b + g;
      for (int f = F_main + t64, it = 0; f < r.cols; f += kRedThreads, ++it) {
        const bool have = f < r.F && (r.left_rows > 0 || r.Lp);
        float sum = 0.f;
        if (have && r.Lp) {
          for (int k0 = 0; k0 < r.lp_splits; k0 += 8) {  // eight loads in flight, added in split order
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              v[e] = (k0 + e < r.lp_splits) ? __ldcg(r.Lp + ((size_t)(k0 + e) * r.n_left + (f - F_main)) * r.left_ld + t) : 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) if (k0 + e < r.lp_splits) sum += v[e];
          }
        } else if (have) sum = s_left[f - F_main];
        if (f < r.ldwh && (have || f >= r.F)) r.WH[(size_t)t * r.ldwh + f] = sum;
        if (r.X) store_r(r.ro, t, f, have ? __fdiv_rn(it == 0 ? x_tail_now : r.X[(size_t)t * r.ldx + f], fmaxf(sum, p.eps)) : 0.f);
      }
    }
    EVC_CLK_ADD(c_tail, c_a);
    EVC_CLK(c_a);
    named_bar_sync(1, kEpiWarps * 32);  // the slot (and s_left) may be overwritten
    EVC_CLK_ADD(c_bar, c_a);
    EVC_CLK(c_a);
    const int bn = b + r.nslots * r.contributors;
    if (issuer && bn < n_batches) issue(bn, s);
    EVC_CLK_ADD(c_iss, c_a);
  }
  EVC_CLK(c_3);
  if (tid == 0)
    EVC_CLK_PRINT(p, "clk cta %d fused reduce: fence %lld barrier %lld reduce %lld (copies %lld left %lld main %lld tail %lld bar %lld issue %lld, %d batches)\n",
                  (int)blockIdx.x, c_1 - c_0, c_2 - c_1, c_3 - c_2, c_w, c_left, c_main, c_tail, c_bar, c_iss, k);
}

// kP = CTA pairs per cluster.  With kP > 1 the pairs of a cluster work on items that share one operand tile -- the
// frame tile of the ratio in contraction 2 (kShareM = false: pair q takes dictionary-row group P*g + q), the
// dictionary tile in contraction 1 (kShareM = true: pair q takes frame tile P*s + q) -- and every CTA fetches only a
// 1/kP slice of that tile, multicasting it to the CTAs of the other pairs that hold the same half: L2 -> SM traffic
// of the shared operand drops by kP (both contractions are bound by that traffic, profiles/r2_ncu_summary.txt).  The ring slots
// then move in lock step across the cluster: a slot is free when EVERY pair's MMAs have read it (kP commits).
template <int kMTiles, int kBlockT, int kPrec, bool kSplitN, int kEpi, int kP = 1, bool kShareM = false>
__global__ void __launch_bounds__(
    (TileCfg<kMTiles, kBlockT, kPrec, kSplitN, kEpi != TEPI_PARTIAL, kEpi == TEPI_MU_FRO>::kThreads), 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmN,
               const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmQ,
               const __grid_constant__ CUtensorMap tmS, const GemmParams p) {
  static_assert(kP == 1 || kP == 2 || kP == 4, "1, 2 or 4 CTA pairs per cluster");
  constexpr bool kFro = (kEpi == TEPI_MU_FRO);
  constexpr bool kStageH = (kEpi != TEPI_PARTIAL);
  using Cfg = TileCfg<kMTiles, kBlockT, kPrec, kSplitN, kStageH, kFro>;
  constexpr int kHB = Cfg::kHBufsUsed;  // chunk buffers in use
  constexpr int kStages = Cfg::kStages;
  constexpr int kAccStages = Cfg::kAccStages;
  constexpr uint32_t kFmt = (kPrec == PREC_TF32) ? kFmtTF32 : kFmtBF16;
  constexpr uint32_t kIdesc = make_idesc(kFmt, 128 * kCG, kBlockT);

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kStages];   // (leader's copy is used) operands of the stage are in place
  __shared__ __align__(8) uint64_t bar_raw[kStages];    // kSplitN: this CTA's fp32 frame tile landed
  __shared__ __align__(8) uint64_t bar_empty[kStages];  // MMAs that read the stage retired (both CTAs' copies fire)
  __shared__ __align__(8) uint64_t bar_acc_full[kAccStages];
  __shared__ __align__(8) uint64_t bar_acc_empty[kAccStages];  // (leader's copy is used)
  __shared__ __align__(8) uint64_t bar_hfull[kHBufs];   // H chunk landed in shared memory
  __shared__ __align__(8) uint64_t bar_hready[kHBufs];  // the 4 epilogue warps of a chunk wrote the updated values
  __shared__ __align__(8) uint64_t bar_hempty[kHBufs];  // the TMA store has read the buffer
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(8) uint64_t bar_red[2];          // FusedReduce: the partials of a batch landed (one per slot)
  __shared__ float red_sh[64];

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform: role branches do not diverge
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const uint32_t rank = crank & 1u;         // rank inside the CTA pair: 0 = leader (issues the MMAs)
  const int q = (int)(crank >> 1);           // pair inside the cluster
  const uint32_t leader = crank & ~1u;      // cluster rank of this pair's leader
  uint8_t* ring_ptr = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t ring = smem_u32(ring_ptr);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmM);
    tma_prefetch_desc(&tmN);
    if (kStageH) tma_prefetch_desc(&tmH);
    if (Cfg::kShadow || (kSplitN && p.n_left > 0)) tma_prefetch_desc(&tmS);
    if (kFro || (kEpi == TEPI_PARTIAL && p.red.enabled == 1)) tma_prefetch_desc(&tmQ);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(smem_u32(&bar_full[i]), Cfg::kFullArrivals);
      mbar_init(smem_u32(&bar_raw[i]), 1);
      mbar_init(smem_u32(&bar_empty[i]), kP);  // one commit per pair of the cluster
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(smem_u32(&bar_acc_full[i]), 1);
      mbar_init(smem_u32(&bar_acc_empty[i]), kEpiWarps * kCG);  // one elected lane of each epilogue warp
    }
    for (int i = 0; i < kHBufs; ++i) {
      mbar_init(smem_u32(&bar_hfull[i]), 1);
      mbar_init(smem_u32(&bar_hready[i]), 4);
      mbar_init(smem_u32(&bar_hempty[i]), (p.direct_store && !Cfg::kShadow) ? 4 : 1);  // the storer's TMA store read it | the 4 warps consumed it
    }
    if (kEpi == TEPI_PARTIAL)
      for (int i = 0; i < 2; ++i) mbar_init(smem_u32(&bar_red[i]), 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc2(smem_u32(&tmem_base_smem), 512); tmem_relinquish2(); }
  tc_fence_before();
  cluster_sync_all();  // the peer's barriers are initialised before any remote arrive / remote transaction
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // everything above (barrier init, TMEM allocation, descriptor prefetch) may overlap the tail of the previous
  // kernel in the stream; from here on its results are needed
  pdl_wait();
  pdl_launch_dependents();

  const int num_items = num_items_of(p);
  // all CTAs of a cluster walk the same (cluster-level) items; pair q takes its own share of each (item_of)
  const int first_item = blockIdx.x / (kCG * kP), item_stride = gridDim.x / (kCG * kP);
  auto item_of = [&](int item) {
    WorkItem w = decode_item(p, item, kBlockT);
    if (kP > 1) {
      if (kShareM) w.t_tile = w.t_tile * kP + q; else w.m_group = w.m_group * kP + q;
    }
    return w;
  };
  // dictionary sub-tiles of row group g that hold at least one real row (the others are pure padding: no MMAs, no update)
  auto valid_subtiles = [&](int g) {
    const int rows_left = p.M_total - g * (Cfg::kRowsPerSub * kMTiles);
    return max(0, min(kMTiles, (rows_left + Cfg::kRowsPerSub - 1) / Cfg::kRowsPerSub));
  };
  // the pair leader's "full" barriers as shared::cluster addresses (stage i at + 8 i); == local address & ~(1 << 24)
  const uint32_t full_leader = mapa_rank(smem_u32(&bar_full[0]), leader);
  // cluster ranks of the CTAs that hold the same half of the shared operand tile as this one; all CTAs; this pair
  uint16_t mc_mask = 0;
#pragma unroll
  for (int qq = 0; qq < kP; ++qq) mc_mask |= (uint16_t)(1u << (2 * qq + (int)rank));
  constexpr uint16_t kAllMask = (uint16_t)((1u << (kCG * kP)) - 1u);
  const uint16_t pair_mask = (uint16_t)(3u << (2 * q));
  constexpr int kSlice = 128 / kP;  // rows of a shared-operand box this CTA fetches (and multicasts)

  if (warp == 0) {
    // ================= TMA producer (one thread in each CTA of the pair) =================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      EVC_CLK_DECL(c_t0 = 0, c_a = 0, c_wait = 0);
      EVC_CLK(c_t0);
      for (int item = first_item; item < num_items; item += item_stride) {
        const WorkItem w = item_of(item);
        const int m0 = w.m_group * (Cfg::kRowsPerSub * kMTiles) + (int)rank * 128;
        // (a half-width item still loads a kNRows-row box: the narrower MMA never reads the surplus rows)
        const int t0 = w.t_tile * kBlockT + w.t_off + (int)rank * (w.t_cols / kCG);
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          EVC_CLK(c_a);
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          EVC_CLK_ADD(c_wait, c_a);
          const uint32_t full = full_leader + (uint32_t)stage * 8u;
          if (EVC_DBG(p, 4)) {
            if (rank == 0) mbar_arrive(smem_u32(&bar_full[stage]));
            if (kSplitN) mbar_arrive(smem_u32(&bar_raw[stage]));
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
            continue;
          }
          // the leader expects both CTAs' bytes (boxes past the matrix edge are zero-filled and still counted)
          if (rank == 0) mbar_arrive_expect_tx(smem_u32(&bar_full[stage]), (uint32_t)(kCG * Cfg::kTxBytes));
          const uint32_t sbase = ring + stage * Cfg::kStageBytes;
          const int kc = kb * Cfg::kKE;        // K element / column of the fp32 tiles
          const int kcb = kb * Cfg::kBoxCols;  // tensor-map column of the K-block's box
#pragma unroll
          for (int i = 0; i < kMTiles; ++i) {
            const uint32_t dst = sbase + i * Cfg::kMSubBytes;
            const int row = m0 + i * Cfg::kRowsPerSub;
            if (kP > 1 && kShareM)
              tma_load_2d_pair_mc(dst + q * kSlice * Cfg::kMRowBytes, &tmM, kcb, row + q * kSlice, full, mc_mask, kEvictNormal);
            else
              tma_load_2d_pair(dst, &tmM, kcb, row, full, kEvictNormal);
          }
          if (kSplitN) {
            const uint32_t rawb = smem_u32(&bar_raw[stage]);
            mbar_arrive_expect_tx(rawb, (uint32_t)(Cfg::kRawBytes + (p.n_left > 0 ? Cfg::kLeftBytes : 0)));
            tma_load_2d(sbase + Cfg::kOffRaw, &tmN, kc, t0, rawb, kEvictFirst);  // H is streamed: keep A^T in L2
            // rows F_main.. of the transposed dictionary for this K-block (rows past n_left are zero-filled)
            if (p.n_left > 0) tma_load_2d(sbase + Cfg::kOffLeft, &tmS, kc, 0, rawb, kEvictNormal);
          } else {
            constexpr int kNSlice = Cfg::kNRows / kP;
            const uint32_t dst = sbase + Cfg::kOffN;
            if (kP > 1 && !kShareM)
              tma_load_2d_pair_mc(dst + q * kNSlice * 128, &tmN, kcb, t0 + q * kNSlice, full, mc_mask, kEvictNormal);
            else
              tma_load_2d_pair(dst, &tmN, kcb, t0, full, kEvictNormal);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
      EVC_CLK_PRINT(p, "clk cta %d producer: total %lld wait_empty %lld\n", (int)blockIdx.x, clock64() - c_t0, c_wait);
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA of the pair only) =================
    if (rank == 0 && elect_one()) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      EVC_CLK_DECL(c_t0 = 0, c_a = 0, c_wacc = 0, c_wfull = 0);
      EVC_CLK(c_t0);
      for (int item = first_item; item < num_items; item += item_stride) {
        const WorkItem w = item_of(item);
        const int m0 = w.m_group * (Cfg::kRowsPerSub * kMTiles);
        const int kb0 = w.kb0, kb1 = w.kb1;
        const uint32_t idesc = (w.t_cols == kBlockT) ? kIdesc : make_idesc(kFmt, 128 * kCG, (uint32_t)w.t_cols);
        EVC_CLK(c_a);
        mbar_wait(smem_u32(&bar_acc_empty[acc]), acc_phase ^ 1u);
        EVC_CLK_ADD(c_wacc, c_a);
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb) {
          EVC_CLK(c_a);
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          EVC_CLK_ADD(c_wfull, c_a);
          tc_fence_after();
          const uint32_t sbase = ring + stage * Cfg::kStageBytes;
          const uint32_t nbase = sbase + Cfg::kOffN;
          const int kvalid = min(Cfg::kKE, p.K - kb * Cfg::kKE);
          const int ksteps = (kvalid + Cfg::kKStep - 1) / Cfg::kKStep;
#pragma unroll
          for (int i = 0; i < kMTiles; ++i) {
            if (m0 + i * Cfg::kRowsPerSub >= p.M_total) break;  // pure padding: no MMAs, the epilogue skips it too
            if (EVC_DBG(p, 2)) break;
            const uint32_t d = tmem_base + (uint32_t)((acc * kMTiles + i) * kBlockT);
            const uint32_t abase = sbase + i * Cfg::kMSubBytes;
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint32_t accum = (kb > kb0 || ks > 0) ? 1u : 0u;
              const uint64_t a1 = make_smem_desc(abase + ks * 32, Cfg::kMRowBytes);
              const uint64_t b1 = make_smem_desc(nbase + ks * 32, Cfg::kNRowBytes);
              if (kPrec == PREC_SPLIT) {
                // small terms first, then the leading one
                const uint64_t a2 = make_smem_desc(abase + Cfg::kMPlaneOff + ks * 32, Cfg::kMRowBytes);
                const uint64_t b2 = make_smem_desc(nbase + Cfg::kNPlaneOff + ks * 32, Cfg::kNRowBytes);
                mma_f16_2cta(d, a2, b1, idesc, accum);
                mma_f16_2cta(d, a1, b2, idesc, 1u);
                mma_f16_2cta(d, a1, b1, idesc, 1u);
              } else if (kPrec == PREC_TF32) {
                mma_tf32_2cta(d, a1, b1, idesc, accum);
              } else {
                mma_f16_2cta(d, a1, b1, idesc, accum);
              }
            }
          }
          // frees the smem slot when these MMAs retire -- in every CTA of the cluster (the other pairs multicast into
          // this pair's slots, so they have to see them free too)
          mma_commit_mc(smem_u32(&bar_empty[stage]), kAllMask);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        // accumulator complete -> epilogue warps of both CTAs of this pair
        mma_commit_mc(smem_u32(&bar_acc_full[acc]), pair_mask);
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
      }
      EVC_CLK_PRINT(p, "clk cta %d mma: total %lld wait_acc_empty %lld wait_full %lld\n", (int)blockIdx.x, clock64() - c_t0, c_wacc, c_wfull);
    }
  } else if (kStageH && warp == Cfg::kLoaderWarp) {
    // ================= H chunk loader: prefetches the activations the epilogue will update =================
    if (elect_one()) {
      int hbase = 0;
      EVC_CLK_DECL(c_t0 = 0, c_a = 0, c_wait = 0);
      EVC_CLK(c_t0);
      for (int item = first_item; item < num_items; item += item_stride) {
        const WorkItem w = item_of(item);
        const int t0 = w.t_tile * kBlockT + w.t_off;
        const int nch = max(0, min(w.t_cols / kHChunkT, (p.T - t0 + kHChunkT - 1) / kHChunkT));
        const int nsub = valid_subtiles(w.m_group);
        // chunk cc of the item = chunk cc % nch of dictionary sub-tile cc / nch
        for (int cc = 0; cc < nsub * nch; ++cc) {
          const int i = cc / nch, c = cc - i * nch;
          const int n0 = (w.m_group * kMTiles + i) * Cfg::kRowsPerSub + (int)rank * 128;
          const int seq = hbase + cc, b = seq % kHB;
          const uint32_t ph = (uint32_t)(seq / kHB) & 1u;
          EVC_CLK(c_a);
          mbar_wait(smem_u32(&bar_hempty[b]), ph ^ 1u);
          EVC_CLK_ADD(c_wait, c_a);
          const uint32_t full = smem_u32(&bar_hfull[b]);
          if (EVC_DBG(p, 16)) { mbar_arrive(full); continue; }
          mbar_arrive_expect_tx(full, (uint32_t)Cfg::kHBufStride);
          tma_load_2d(ring + Cfg::kOffH + b * Cfg::kHBufStride, &tmH, n0, t0 + c * kHChunkT, full, kEvictFirst);
          if (kFro)
            tma_load_2d(ring + Cfg::kOffH + b * Cfg::kHBufStride + kHBufBytes, &tmQ, n0, t0 + c * kHChunkT, full, kEvictFirst);
        }
        hbase += nsub * nch;
      }
      EVC_CLK_PRINT(p, "clk cta %d hloader: total %lld wait_hempty %lld\n", (int)blockIdx.x, clock64() - c_t0, c_wait);
    }
  } else if (kStageH && warp == Cfg::kLoaderWarp + 1) {
    // ================= H chunk storer =================
    if (!(p.direct_store && !Cfg::kShadow) && elect_one()) {
      int hbase = 0;
      EVC_CLK_DECL(c_t0 = 0, c_a = 0, c_wait = 0, c_wread = 0);
      EVC_CLK(c_t0);
      for (int item = first_item; item < num_items; item += item_stride) {
        const WorkItem w = item_of(item);
        const int t0 = w.t_tile * kBlockT + w.t_off;
        const int nch = max(0, min(w.t_cols / kHChunkT, (p.T - t0 + kHChunkT - 1) / kHChunkT));
        const int nsub = valid_subtiles(w.m_group);
        for (int cc = 0; cc < nsub * nch; ++cc) {
          const int i = cc / nch, c = cc - i * nch;
          const int n0 = (w.m_group * kMTiles + i) * Cfg::kRowsPerSub + (int)rank * 128;
          const int seq = hbase + cc, b = seq % kHB;
          const uint32_t ph = (uint32_t)(seq / kHB) & 1u;
          EVC_CLK(c_a);
          mbar_wait(smem_u32(&bar_hready[b]), ph);
          EVC_CLK_ADD(c_wait, c_a);
          if (!EVC_DBG(p, 16)) {
            tma_store_2d(&tmH, n0, t0 + c * kHChunkT, ring + Cfg::kOffH + b * Cfg::kHBufStride);
            if (Cfg::kShadow) tma_store_2d(&tmS, n0, t0 + c * kHChunkT, ring + Cfg::kOffS + b * Cfg::kSBufBytes);
            tma_store_commit();
            EVC_CLK(c_a);
            tma_store_wait_read();
            EVC_CLK_ADD(c_wread, c_a);
          }
          mbar_arrive(smem_u32(&bar_hempty[b]));
        }
        hbase += nsub * nch;
      }
      tma_store_wait_all();
      EVC_CLK_PRINT(p, "clk cta %d hstorer: total %lld wait_hready %lld wait_store_read %lld\n", (int)blockIdx.x, clock64() - c_t0, c_wait, c_wread);
    }
  } else {
    // ================= epilogue: 8 warps.  Warp w may touch TMEM lanes [32*(w%4), +32); the two warps
    // of a lane quarter take alternate 32-column chunks, so each scheduler has two warps to overlap
    // the memory round trips of the fused update. =================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    int acc = 0, stage = 0, hbase = 0;
    uint32_t acc_phase = 0, phase = 0;
    EVC_CLK_DECL(c_t0 = 0, c_a = 0, c_wacc = 0, c_wh = 0, c_wraw = 0, c_red = 0);
    EVC_CLK(c_t0);
    for (int item = first_item; item < num_items; item += item_stride) {
      const WorkItem w = item_of(item);
      const int m_group = w.m_group, split = w.split;
      const int t0 = w.t_tile * kBlockT + w.t_off;
      if (kSplitN) {
        // single accumulator stage: these warps have nothing to drain during the main loop, so they derive the bf16
        // planes of the frame operand (the activations) from the fp32 tile TMA brought -- all 8 warps on every
        // stage, so every warp observes every phase of bar_raw
        const int tid = threadIdx.x - 64;
        constexpr int kPer = Cfg::kNRows * 4 / (kEpiWarps * 32);
        // the pairs of dictionary-row group 0 also carry the rows that stay off the tensor cores (same K range)
        const int n_left = (m_group == 0) ? p.n_left : 0;
        float lacc[kPer][8];
#pragma unroll
        for (int q = 0; q < kPer; ++q)
#pragma unroll
          for (int l = 0; l < 8; ++l) lacc[q][l] = 0.f;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          EVC_CLK(c_a);
          mbar_wait(smem_u32(&bar_raw[stage]), phase);
          EVC_CLK_ADD(c_wraw, c_a);
          uint8_t* sb = ring_ptr + stage * Cfg::kStageBytes;
          if (!EVC_DBG(p, 1))
            split_planes<Cfg::kNRows>(sb + Cfg::kOffRaw, sb + Cfg::kOffN, sb + Cfg::kOffN + Cfg::kNPlaneBytes, tid,
                                      reinterpret_cast<const float*>(sb + Cfg::kOffLeft), n_left, lacc);
          fence_proxy_async_smem();  // generic-proxy stores -> visible to tcgen05.mma's operand reads
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(full_leader + (uint32_t)stage * 8u);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        // this split's share of (A H)[t, F_main + l]: the four quarter-row threads of a frame are adjacent lanes
#pragma unroll
        for (int l = 0; l < 8; ++l) {
          if (l >= n_left) break;
#pragma unroll
          for (int q = 0; q < kPer; ++q) {
            float v = lacc[q][l];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            const int t = t0 + (int)rank * Cfg::kNRows + ((q * (kEpiWarps * 32) + tid) >> 2);
            if ((tid & 3) == 0 && t < p.left_ld) p.left_out[((size_t)split * p.n_left + l) * p.left_ld + t] = v;
          }
        }
      }
      EVC_CLK(c_a);
      mbar_wait(smem_u32(&bar_acc_full[acc]), acc_phase);
      EVC_CLK_ADD(c_wacc, c_a);
      tc_fence_after();
      if (kStageH) {
        // ---- fused multiplicative update through the shared-memory H chunks ----
        const int nch = max(0, min(w.t_cols / kHChunkT, (p.T - t0 + kHChunkT - 1) / kHChunkT));
        const int nsub = valid_subtiles(m_group);
        int cur_i = -1, m = 0;
        float den = 1.f, inv_den = 1.f;
        float la[8];  // this lane's entries of the leftover dictionary rows (n_left <= 8)
        for (int cc = half; cc < nsub * nch; cc += 2) {
          const int i = cc / nch, c = cc - i * nch;
          if (i != cur_i) {  // per dictionary sub-tile: this lane's row, its denominator, its leftover entries
            cur_i = i;
            m = (m_group * kMTiles + i) * Cfg::kRowsPerSub + (int)rank * 128 + quarter * 32 + lane;
            den = ((m < p.M_total && !kFro) ? p.colsum[m] : 1.f) + p.lam;
            if (den == 0.f) den = p.eps;
            inv_den = __frcp_rn(den);
#pragma unroll
            for (int l = 0; l < 8; ++l) la[l] = (l < p.n_left && m < p.M_total) ? p.left_a[(size_t)l * p.left_lda + m] : 0.f;
          }
          const int seq = hbase + cc, b = seq % kHB;
          const uint32_t ph = (uint32_t)(seq / kHB) & 1u;
          uint32_t v[32];
          if (!EVC_DBG(p, 64))
            tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((acc * kMTiles + i) * kBlockT + c * 32), v);
          EVC_CLK(c_a);
          mbar_wait(smem_u32(&bar_hfull[b]), ph);
          EVC_CLK_ADD(c_wh, c_a);
          float* hb = reinterpret_cast<float*>(ring_ptr + Cfg::kOffH + b * Cfg::kHBufStride) + quarter * 32 + lane;
          float h[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) h[j] = hb[j * 128];
          tmem_ld_wait();
          const int tbm = t0 + c * 32;
          if (EVC_DBG(p, 8)) {
            // (timing experiments: activations pass through unchanged)
          } else if (kFro) {
            // Frobenius: H <- H * (X A^T) / (A^T (A H) + lambda); the numerator chunk sits behind the H chunk
            const float* qb = hb + kHBufBytes / 4;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (p.row_active == nullptr || (tbm + j < p.T && p.row_active[tbm + j])) {
                float dn = __uint_as_float(v[j]) + p.lam;
                if (dn == 0.f) dn = p.eps;
                h[j] = h[j] * __fdiv_rn(qb[j * 128], dn);
              }
            }
          } else if (p.row_active == nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) h[j] = h[j] * quotient(__uint_as_float(v[j]), den, inv_den);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (tbm + j < p.T && p.row_active[tbm + j]) h[j] = h[j] * quotient(__uint_as_float(v[j]), den, inv_den);
          }
          if (Cfg::kShadow) {
            // bf16 shadow of the updated activations (the K operand of the next contraction 1), staged like the chunk
            __nv_bfloat16* sbuf = reinterpret_cast<__nv_bfloat16*>(ring_ptr + Cfg::kOffS + b * Cfg::kSBufBytes) + quarter * 32 + lane;
#pragma unroll
            for (int j = 0; j < 32; ++j) sbuf[j * 128] = __float2bfloat16_rn(h[j]);
          }
          if (p.direct_store && !Cfg::kShadow) {
            // straight to global memory: a warp writes 32 consecutive exemplars of one frame (128 bytes) per store;
            // the chunk buffer is free as soon as its values are in registers
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_hempty[b]));
            if (m < p.M_total) {
              float* g = p.out + (size_t)tbm * p.ld_out + m;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (tbm + j < p.T) g[(size_t)j * p.ld_out] = h[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) hb[j * 128] = h[j];
            fence_proxy_async_smem();  // generic-proxy stores -> visible to the TMA store
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_hready[b]));
          }
#pragma unroll
          for (int l = 0; l < 8; ++l) {
            if (l >= p.n_left || EVC_DBG(p, 32)) break;
            // (rows past T and exemplars past N were zero-filled by TMA: they add nothing)
            const float a = la[l];
            float sv[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) sv[j] = h[j] * a;
            const float tot = warp_transpose_sum(sv, lane);
            const int prow = ((m_group * kMTiles + i) * kCG + (int)rank) * 4 + quarter;  // one partial row per 32 exemplars
            p.left_out[((size_t)l * p.left_ld + (t0 + c * 32 + lane)) * p.left_rows + prow] = tot;
          }
        }
        hbase += nsub * nch;
      }
#pragma unroll
      for (int i = 0; i < (kStageH ? 0 : kMTiles); ++i) {
        const int mrow0 = m_group * (Cfg::kRowsPerSub * kMTiles) + i * Cfg::kRowsPerSub + (int)rank * 128;
        if (m_group * (Cfg::kRowsPerSub * kMTiles) + i * Cfg::kRowsPerSub >= p.M_total) break;  // whole MMA is padding
        const int m = mrow0 + quarter * 32 + lane;
        const bool m_ok = m < p.M_total;
        const bool rows_full = (mrow0 + 128 <= p.M_total);
        for (int c = half; c < w.t_cols / 32; c += 2) {
          const int tb = t0 + c * 32;
          if (tb >= p.T) break;  // warp-uniform
          uint32_t v[32];
          const uint32_t taddr =
              tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((acc * kMTiles + i) * kBlockT + c * 32);
          tmem_ld_32x32(taddr, v);
          if (p.red.enabled == 2) {
            // K is not split: this accumulator IS A*H -- the ratio X / max(A H, eps) (and / or A*H itself) leaves
            // straight from TMEM, no partials, no second pass (sklearn _nmf.py:554-571)
            const FusedReduce& r = p.red;
            float x[32];
            if (r.X) {
#pragma unroll
              for (int j = 0; j < 32; ++j) x[j] = (m_ok && tb + j < p.T) ? __ldg(r.X + (size_t)(tb + j) * r.ldx + m) : 0.f;
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (!m_ok || tb + j >= p.T) continue;
              const float sum = __uint_as_float(v[j]);
              if (r.WH) r.WH[(size_t)(tb + j) * r.ldwh + m] = sum;
              if (r.X) store_r(r.ro, tb + j, m, __fdiv_rn(x[j], fmaxf(sum, p.eps)));
            }
            continue;
          }
          tmem_ld_wait();
          // whole chunk in range and every lane a real row: straight-line code
          const bool fast = (tb + 32 <= p.T) && rows_full;
          float* o = p.out + ((size_t)split * p.T + tb) * p.ld_out + m;
          if (fast && p.out_keep_l2) {
            // split-K partials: 37 MB that the reduction kernel reads right after this one -- keep them in L2 (they
            // were evicted by the streamed operands and came back from DRAM: 16 % L2 hits, profiles/r2_reduce_before_bulk_ncu.txt)
#pragma unroll
            for (int j = 0; j < 32; ++j) st_global_hint(o + (size_t)j * p.ld_out, __uint_as_float(v[j]), kEvictLast);
          } else if (fast) {
#pragma unroll
            for (int j = 0; j < 32; ++j) o[(size_t)j * p.ld_out] = __uint_as_float(v[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (m_ok && tb + j < p.T) o[(size_t)j * p.ld_out] = __uint_as_float(v[j]);
          }
        }
      }
      if (kEpi == TEPI_PARTIAL && p.red.enabled == 1) {
        // split-K sum (+ ratio) by the contraction's own CTAs; one item per pair in this mode, the ring is idle
        const int groups = p.splits_last ? p.num_m_groups - 1 : p.num_m_groups;
        const int idx = (item < p.items_main) ? w.m_group + groups * split : groups * p.num_splits + split;
        EVC_CLK(c_a);
        fused_reduce_phase<kBlockT>(p, &tmQ, w.t_tile, idx * kCG + (int)rank, reinterpret_cast<float*>(ring_ptr), bar_red,
                                    red_sh, (int)threadIdx.x - 64);
        EVC_CLK_ADD(c_red, c_a);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_rank(smem_u32(&bar_acc_empty[acc]), leader));
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0 && (warp == 2 || warp == 6))
      EVC_CLK_PRINT(p, "clk cta %d epilogue warp %d: total %lld wait_acc_full %lld wait_hfull %lld wait_raw %lld fused_reduce %lld\n",
                    (int)blockIdx.x, warp, clock64() - c_t0, c_wacc, c_wh, c_wraw, c_red);
  }

  tc_fence_before();
  // nobody leaves (and frees shared memory / TMEM the leader's MMAs may still read) before both CTAs are done
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// ---- memory-bound helpers ------------------------------------------------------------------------

// WH[t,f] = sum_s P[s][t][f] for the tensor-core rows f < F_main (fixed order: deterministic);
// WH[t,F_main+l] = sum_r L[l][t][r] from the fused update's per-warp partials when `left_rows` > 0.
// With `X` != nullptr it also emits the ratio R = X / max(WH, eps) (zero pad columns) in the same pass.
//
// One block per (frame, 512-column chunk).  The split-K partials of the chunk -- S segments of 2 KB, megabytes apart
// in the workspace -- are staged in shared memory by cp.async.bulk copies issued by one thread and summed from there:
// with per-lane 16-byte loads this kernel ran at 2.0 TB/s (the SM's outstanding-miss capacity, profiles/r2_reduce_before_bulk_ncu.txt),
// bulk copies are not subject to that limit.
// 256 columns x 64 threads per block: 18 KB of staging at the headline shape, so eleven blocks share an SM and the
// 3 000 blocks of a launch run in under two waves (512 columns: five per SM, 2.7 waves, 17.6 us -- profiles/r2_ncu_summary.txt).
__global__ void __launch_bounds__(kRedThreads)
reduce_partials_kernel(const __grid_constant__ CUtensorMap tmP,  // P as a (column, frame, split) tensor, boxes 256 x 1 x batch
                       const float* __restrict__ P, int S, int S_last, int f_last, int T, int ldp, int F, int F_main,
                       float* __restrict__ WH, int ldwh, const float* __restrict__ L, int left_rows, int n_left,
                       int left_ld, const float* __restrict__ X, int ldx, float eps, ROut ro, int batch, int chunk0,
                       const float* __restrict__ Lp, int lp_splits) {
  extern __shared__ __align__(128) float red_smem[];  // [batch][kRedCols] staged partials | [n_left][left_rows]
  __shared__ __align__(8) uint64_t bar;
  __shared__ float s_left[8];
  __shared__ float s_warp[kRedThreads / 32];
  const int t = blockIdx.x, c0 = ((int)blockIdx.y + chunk0) * kRedCols;  // chunk0 > 0: only the columns past the tensor-core rows
  float* stage = red_smem;
  float* lstage = red_smem + (size_t)batch * kRedCols;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmP);
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  const int pcols = max(0, min(kRedCols, min(ldp, F_main) - c0));  // tensor-core columns of this chunk (a multiple of 4)
  const int pcols_ld = max(0, min(kRedCols, ldp - c0));             // ... including the partial buffer's pad columns
  const int n = (c0 >= f_last) ? S_last : S;                        // the last row group has its own split count
  const bool has_left = left_rows > 0 && F_main >= c0 && F_main < c0 + kRedCols;
  const int col = threadIdx.x * 4;  // this thread's 4 columns of the chunk
  // the frame's X entries do not depend on the kernel before this one: in flight across the dependency wait
  const bool x_vec = X && col < pcols && c0 + col + 3 < F_main && (ldx & 3) == 0 && (((uintptr_t)X) & 15) == 0;
  float4 x4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (x_vec) x4 = __ldg(reinterpret_cast<const float4*>(X + (size_t)t * ldx + c0 + col));
  __syncthreads();
  pdl_wait();
  pdl_launch_dependents();
  uint32_t phase = 0;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k0 = 0; k0 < n || (k0 == 0 && has_left); k0 += batch) {
    const int nb = max(0, min(batch, n - k0));
    if (threadIdx.x == 0) {
      // ONE box load brings the chunk's rows of all `batch` splits (per-split 1 KB bulk copies were bound by their
      // issue rate); splits / columns past the tensor are zero-filled and still counted
      uint32_t bytes = (pcols_ld > 0 && nb > 0) ? (uint32_t)batch * kRedCols * 4u : 0u;
      if (k0 == 0 && has_left) bytes += (uint32_t)n_left * left_rows * 4u;
      mbar_arrive_expect_tx(smem_u32(&bar), bytes);
      if (pcols_ld > 0 && nb > 0) tma_load_3d(smem_u32(stage), &tmP, c0, t, k0, smem_u32(&bar), kEvictFirst);
      if (k0 == 0 && has_left)
        for (int l = 0; l < n_left; ++l)
          bulk_load_1d(smem_u32(lstage + (size_t)l * left_rows), L + ((size_t)l * left_ld + t) * left_rows,
                       (uint32_t)left_rows * 4u, smem_u32(&bar), kEvictFirst);
    }
    mbar_wait(smem_u32(&bar), phase);
    phase ^= 1u;
    if (col < pcols)
      for (int kq = 0; kq < nb; kq += 6) {  // six splits' loads in flight, added in split order: deterministic
        float4 v[6];
#pragma unroll
        for (int e = 0; e < 6; ++e)
          v[e] = (kq + e < nb) ? *reinterpret_cast<const float4*>(stage + (size_t)(kq + e) * kRedCols + col) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int e = 0; e < 6; ++e)
          if (kq + e < nb) { acc.x += v[e].x; acc.y += v[e].y; acc.z += v[e].z; acc.w += v[e].w; }
      }
    if (k0 == 0 && has_left) {
      for (int l = 0; l < n_left; ++l) {
        float a = 0.f;
        for (int r = threadIdx.x; r < left_rows; r += blockDim.x) a += lstage[(size_t)l * left_rows + r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = a;
        __syncthreads();
        if (threadIdx.x == 0) {
          float tot = 0.f;
#pragma unroll
          for (int w = 0; w < kRedThreads / 32; ++w) tot += s_warp[w];
          s_left[l] = tot;
        }
        __syncthreads();
      }
    }
    __syncthreads();  // the stage is free for the next batch
  }
  const float sv[4] = {acc.x, acc.y, acc.z, acc.w};
  const int cols = max(ldwh, X ? ro.cols() : 0);
  if (c0 + col + 3 < F_main && (!X || x_vec) && (ldwh & 3) == 0 && (((uintptr_t)WH) & 15) == 0) {
    // four tensor-core columns: vector stores
    const int f = c0 + col;
    if (f < ldwh) *reinterpret_cast<float4*>(WH + (size_t)t * ldwh + f) = acc;
    if (X)
      store_r4(ro, t, f, make_float4(__fdiv_rn(x4.x, fmaxf(acc.x, eps)), __fdiv_rn(x4.y, fmaxf(acc.y, eps)),
                                     __fdiv_rn(x4.z, fmaxf(acc.z, eps)), __fdiv_rn(x4.w, fmaxf(acc.w, eps))));
    return;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int f = c0 + col + j;
    if (f >= cols) break;
    float s = 0.f;
    bool have = false;
    if (f < F_main) { s = sv[j]; have = true; }
    else if (f < F && Lp) {  // rows kept off the tensor cores: contraction 1's per-split sums [split][row][frame]
      for (int k0 = 0; k0 < lp_splits; k0 += 8) {  // eight loads in flight, added in split order
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e)
          v[e] = (k0 + e < lp_splits) ? __ldcg(Lp + ((size_t)(k0 + e) * n_left + (f - F_main)) * left_ld + t) : 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) if (k0 + e < lp_splits) s += v[e];
      }
      have = true;
    }
    else if (f < F && left_rows > 0) { s = s_left[f - F_main]; have = true; }
    if (f < ldwh && (have || f >= F)) WH[(size_t)t * ldwh + f] = s;
    if (X) {
      const float r = (have && f < F) ? __fdiv_rn(X[(size_t)t * ldx + f], fmaxf(s, eps)) : 0.f;
      store_r(ro, t, f, r);
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
  const float* hrow = H + (size_t)t * ldh;
  // 16-byte loads, four per thread in flight (rows are 16-byte aligned for every H / A^T this library allocates)
  const bool vec = ((((uintptr_t)hrow) | ((uintptr_t)a)) & 15) == 0 && (lda & 3) == 0;
  const int n4 = vec ? (N >> 2) : 0;
  for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * blockDim.x) {
    float4 h[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i0 + q * blockDim.x;
      h[q] = i < n4 ? __ldcs(reinterpret_cast<const float4*>(hrow) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      if (l >= n_left) break;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = i0 + q * blockDim.x;
        if (i < n4) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(a + (size_t)l * lda) + i);
          acc[l] = fmaf(h[q].x, w.x, acc[l]); acc[l] = fmaf(h[q].y, w.y, acc[l]);
          acc[l] = fmaf(h[q].z, w.z, acc[l]); acc[l] = fmaf(h[q].w, w.w, acc[l]);
        }
      }
    }
  }
  for (int n = (n4 << 2) + threadIdx.x; n < N; n += blockDim.x) {
    const float h = hrow[n];
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
__global__ void ratio_pad_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ WH, int ldwh, int T,
                                 int F, float eps, int copy, ROut ro) {
  const int f = blockIdx.y * blockDim.x + threadIdx.x;
  const int t = blockIdx.x;
  if (t >= T || f >= ro.cols()) return;
  float r = 0.f;
  if (f < F) {
    const float wh = WH[(size_t)t * ldwh + f];
    r = copy ? wh : __fdiv_rn(X[(size_t)t * ldx + f], fmaxf(wh, eps));
  }
  store_r(ro, t, f, r);
}

// dst (rows, ldd) bf16 = src (rows, lds) fp32, round to nearest even; pad columns [cols, ldd) are zeroed.
// With `interleave` the destination holds both planes of the fp32-accurate split, x1 = bf16_rn(x) and
// x2 = bf16_rn(x - x1), side by side per 32 columns (k_pitch_i / col_i).
__global__ void to_bf16_kernel(const float* __restrict__ src, int lds, __nv_bfloat16* __restrict__ dst, int ldd,
                               int rows, int cols, int interleave) {
  const int r = blockIdx.x;
  const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  const int lcols = interleave ? ldd / 2 : ldd;  // logical columns
  if (r >= rows || c0 >= lcols) return;
  const float* sp = src + (size_t)r * lds;
  __nv_bfloat16* dp = dst + (size_t)r * ldd;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + j;
    if (c >= lcols) break;
    const float x = c < cols ? sp[c] : 0.f;
    const __nv_bfloat16 x1 = __float2bfloat16_rn(x);
    if (interleave) {
      dp[col_i(c)] = x1;
      dp[col_i(c) + 32] = __float2bfloat16_rn(x - __bfloat162float(x1));
    } else {
      dp[c] = x1;
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
inline int make_tmap_any(CUtensorMap* m, const void* base, int esize, long long rows, int cols, int ld, int box_cols,
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
    return fail(EVC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%d ld=%d box=%dx%d", (int)r, rows, cols,
                ld, box_rows, box_cols);
  return EVC_OK;
}
inline int make_tmap(CUtensorMap* m, const float* base, int rows, int cols, int ld, int box_cols, int box_rows,
                     bool swizzle = true) {
  return make_tmap_any(m, base, 4, rows, cols, ld, box_cols, box_rows, swizzle);
}
inline int make_tmap16(CUtensorMap* m, const __nv_bfloat16* base, long long rows, int cols, int ld, int box_cols,
                       int box_rows) {
  return make_tmap_any(m, base, 2, rows, cols, ld, box_cols, box_rows, true);
}

// fp32 tensor (d2, d1, d0) with d0 contiguous, row pitch ld0 elements and plane pitch ld1 elements; box b0 x b1 x b2,
// no swizzle (the split-K partials [split][frame][column] read back by contraction 1's own CTAs).
inline int make_tmap3d(CUtensorMap* m, const float* base, int d0, int d1, int d2, size_t ld0, size_t ld1, int b0, int b1,
                       int b2) {
  PFN_encodeTiled enc;
  EVC_TRY(get_encode(&enc));
  if (((uintptr_t)base & 15) || ((ld0 * 4) & 15) || ((ld1 * 4) & 15) || b0 > 256 || b1 > 256 || b2 > 256)
    return fail(EVC_ERR_INVALID_ARGUMENT, "bad 3-D tensor map (pitches %zu %zu, box %d %d %d)", ld0, ld1, b0, b1, b2);
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)ld0 * 4, (cuuint64_t)ld1 * 4};
  cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(EVC_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed (%d)", (int)r);
  return EVC_OK;
}

// SM count of the CURRENT device (a process may hold dictionaries on several GPUs)
inline int num_sms() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cache[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}
inline int cta_group() { return kCG; }

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

// CTA pairs per cluster asked for by the environment (A/B runs) or the measured default; 1, 2 or 4.
inline int env_pairs(const char* name, int dflt) {
  const char* v = getenv(name);
  const int pr = v ? atoi(v) : dflt;
  return (pr == 2 || pr == 4) ? pr : 1;
}

// How many clusters of kCG * P CTAs of this kernel the CURRENT device holds at once (B200: 74 pairs, 33 clusters of
// two pairs, 15 of four: tools/probe/cluster_probe.cu); without a device (host-only tests) the SM count decides.
template <class Kern>
inline int active_clusters(Kern kern, int threads, int smem, int P) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(kCG * P * 64);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCG * P;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = 0;
  }
  return n;
}

template <int kMTiles, int kBlockT, int kPrec, bool kSplitN, int kEpi, int kP = 1, bool kShareM = false>
struct TcLaunch {
  using Cfg = TileCfg<kMTiles, kBlockT, kPrec, kSplitN, kEpi != TEPI_PARTIAL, kEpi == TEPI_MU_FRO>;
  // configure (per device) and return the number of clusters resident at once; 0 = this cluster size cannot run here
  static int slots() {
    static int cache[64];
    static bool configured[64] = {false};
    auto kern = tc_gemm_kernel<kMTiles, kBlockT, kPrec, kSplitN, kEpi, kP, kShareM>;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return num_sms() / (kCG * kP); }
    if (!configured[dev]) {
      // the > 48 KB dynamic shared memory opt-in is per device
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) {
        cudaGetLastError();
        return num_sms() / (kCG * kP);
      }
      cache[dev] = (kP == 1) ? num_sms() / kCG : active_clusters(kern, Cfg::kThreads, Cfg::kSmemBytes, kP);
      configured[dev] = true;
    }
    return cache[dev];
  }
  static int launch(const CUtensorMap& tmM, const CUtensorMap& tmN, const CUtensorMap& tmH, const CUtensorMap& tmQ,
                    const CUtensorMap& tmS, const GemmParams& p, cudaStream_t s) {
    auto kern = tc_gemm_kernel<kMTiles, kBlockT, kPrec, kSplitN, kEpi, kP, kShareM>;
    const int nslots = slots();
    if (nslots <= 0) return fail(EVC_ERR_UNSUPPORTED, "clusters of %d CTA pairs cannot be scheduled on this device", kP);
    const int items = num_items_of(p);
    if (items <= 0) return EVC_OK;
    const int grid = (items < nslots ? items : nslots) * kCG * kP;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(Cfg::kThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCG * kP;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl() ? 2 : 1;
#ifdef EVC_INSTRUMENT
    static const int dbg = getenv("EVC_DEBUG_FLAGS") ? atoi(getenv("EVC_DEBUG_FLAGS")) : 0;
    GemmParams qp = p;
    qp.debug_flags = dbg;
    EVC_CUDA(cudaLaunchKernelEx(&cfg, kern, tmM, tmN, tmH, tmQ, tmS, qp));
#else
    EVC_CUDA(cudaLaunchKernelEx(&cfg, kern, tmM, tmN, tmH, tmQ, tmS, p));
#endif
    EVC_LAUNCH_CHECK();
    return EVC_OK;
  }
};

// Tile shapes per contraction.
constexpr int kC1MTiles = 2, kC1BlockT = 256;  // contraction 1 / conversion: 512 dictionary rows x 256 frames, split-K
constexpr int kC2MTiles = 1, kC2BlockT = 256;  // contraction 2: 256 exemplars x 256 frames, 2 accumulator stages
// K elements per K-block of a mode (one swizzle row)
inline int bk_elems(int mode) { return mode == EVC_MODE_BF16 ? 64 : 32; }
// rows reserved per bf16 plane in a stacked operand: a multiple of every row-group size, so no TMA box straddles planes
inline long long plane_rows(int rows) { return (long long)round_up(rows, 1024); }

// Resident tensor-core operands of one dictionary.
struct DictOperands {
  int F = 0, N = 0, ldA = 0, ldN = 0;
  bool has_target = false;
  const float* A = nullptr;  // (N, ldA), borrowed from the handle
  // EVC_MODE_TF32: fp32 K-major operands
  float* AT = nullptr;       // (F, ldN) transposed copy: operand of contraction 1
  float* BT = nullptr;       // (F, ldN) transposed target dictionary: operand of the conversion
  CUtensorMap tmA, tmAT, tmBT;
  int F_main = 0, n_left = 0;  // contraction 1 runs rows [0, F_main) on the tensor cores; n_left = F - F_main <= 8
  bool left_valid = false;     // the workspace holds leftover partials of the CURRENT activations
  float* ATleft = nullptr;     // (n_left, ldN) fp32 rows F_main.. of A^T and B^T for the CUDA-core leftover rows
  float* BTleft = nullptr;
  // EVC_MODE_3XTF32: stacked bf16 planes (plane 1 at row *_rows); EVC_MODE_BF16: plane 0 only
  int ldA16 = 0, ldN16 = 0;
  long long a_rows = 0, at_rows = 0;
  __nv_bfloat16 *A16 = nullptr, *AT16 = nullptr, *BT16 = nullptr;
  CUtensorMap tmA16, tmAT16, tmBT16;
  CUtensorMap tmAT16_s[2], tmBT16_s[2];  // the same with 64- and 32-row boxes: the slices clusters of 2 / 4 pairs multicast
  DevBuf h16, r16;  // per solve: bf16 shadow of H (BF16 mode); the ratio as bf16 / as two bf16 planes
  long long r_rows = 0;  // rows per plane of r16 in the split mode
  DevBuf red_ctr;        // FusedReduce: arrival / passed counters per frame tile (zero between launches)
  void release() {
    cudaFree(AT); cudaFree(BT); cudaFree(A16); cudaFree(AT16); cudaFree(BT16); cudaFree(ATleft); cudaFree(BTleft);
    AT = BT = ATleft = BTleft = nullptr;
    A16 = AT16 = BT16 = nullptr;
    h16.release(); r16.release(); red_ctr.release();
  }
};

// dst (bf16, or interleaved hi / lo planes) <- src fp32 (rows, lds); the destination was zeroed, so pad rows are zero.
inline int launch_to_bf16(const float* src, int lds, __nv_bfloat16* dst, int ldd, int rows, int cols, bool interleave,
                          cudaStream_t s) {
  if (rows <= 0) return EVC_OK;
  dim3 g(rows, ceil_div(interleave ? ldd / 2 : ldd, 4 * 256));
  to_bf16_kernel<<<g, 256, 0, s>>>(src, lds, dst, ldd, rows, cols, interleave ? 1 : 0);
  EVC_LAUNCH_CHECK();
  return EVC_OK;
}

inline int build_operands(DictOperands* o, int mode, const float* A, const float* B, int ldA, int F, int N, cudaStream_t s) {
  o->F = F; o->N = N; o->ldA = ldA; o->ldN = round_up(N, 4); o->A = A; o->has_target = (B != nullptr);
  // A few rows past a multiple of 128 (the Nyquist bin of a 513-bin spectrum) would cost a whole 128-row MMA
  // tile; they are handled as dot products on the CUDA cores instead.
  o->n_left = (F > 128 && (F % 128) <= 8 && !getenv("EVC_NO_LEFTOVER")) ? F % 128 : 0;
  o->F_main = F - o->n_left;
  const size_t at_bytes = (size_t)F * o->ldN * sizeof(float);
  EVC_TRY(o->red_ctr.reserve(2 * kRedMaxTiles * sizeof(unsigned int)));
  EVC_CUDA(cudaMemsetAsync(o->red_ctr.p, 0, o->red_ctr.bytes, s));
  dim3 tb(32, 8), tg(ceil_div(N, 32), ceil_div(F, 32));
  // fp32 transposes: resident operands in TF32 mode, staging for the bf16 copies (and the leftover rows) otherwise
  EVC_CUDA(cudaMalloc(&o->AT, at_bytes));
  EVC_CUDA(cudaMemsetAsync(o->AT, 0, at_bytes, s));
  simt::transpose_kernel<<<tg, tb, 0, s>>>(A, ldA, o->AT, o->ldN, N, F);
  EVC_LAUNCH_CHECK();
  if (B) {
    EVC_CUDA(cudaMalloc(&o->BT, at_bytes));
    EVC_CUDA(cudaMemsetAsync(o->BT, 0, at_bytes, s));
    simt::transpose_kernel<<<tg, tb, 0, s>>>(B, ldA, o->BT, o->ldN, N, F);
    EVC_LAUNCH_CHECK();
  }
  if (o->n_left > 0) {
    const size_t lb = (size_t)o->n_left * o->ldN * sizeof(float);
    EVC_CUDA(cudaMalloc(&o->ATleft, lb));
    EVC_CUDA(cudaMemcpyAsync(o->ATleft, o->AT + (size_t)o->F_main * o->ldN, lb, cudaMemcpyDeviceToDevice, s));
    if (B) {
      EVC_CUDA(cudaMalloc(&o->BTleft, lb));
      EVC_CUDA(cudaMemcpyAsync(o->BTleft, o->BT + (size_t)o->F_main * o->ldN, lb, cudaMemcpyDeviceToDevice, s));
    }
  }
  if (mode == EVC_MODE_TF32) {
    EVC_TRY(make_tmap(&o->tmA, A, N, F, ldA, 32, 128));
    EVC_TRY(make_tmap(&o->tmAT, o->AT, o->F_main, N, o->ldN, 32, 128));
    if (B) EVC_TRY(make_tmap(&o->tmBT, o->BT, o->F_main, N, o->ldN, 32, 128));
    return EVC_OK;
  }
  // 16-bit operands: bf16 copies (bf16 mode) or interleaved hi / lo planes (fp32-accurate mode); either way one
  // K-block is 64 tensor-map columns = one 128-byte row segment
  const bool il = (mode == EVC_MODE_3XTF32);
  o->ldA16 = il ? k_pitch_i(F) : k_pitch16(F); o->ldN16 = il ? k_pitch_i(N) : k_pitch16(N);
  o->a_rows = plane_rows(N); o->at_rows = plane_rows(o->F_main);
  const size_t a_elems = (size_t)o->a_rows * o->ldA16, at_elems = (size_t)o->at_rows * o->ldN16;
  EVC_CUDA(cudaMalloc(&o->A16, a_elems * sizeof(__nv_bfloat16)));
  EVC_CUDA(cudaMalloc(&o->AT16, at_elems * sizeof(__nv_bfloat16)));
  EVC_CUDA(cudaMemsetAsync(o->A16, 0, a_elems * sizeof(__nv_bfloat16), s));
  EVC_CUDA(cudaMemsetAsync(o->AT16, 0, at_elems * sizeof(__nv_bfloat16), s));
  EVC_TRY(launch_to_bf16(A, ldA, o->A16, o->ldA16, N, F, il, s));
  EVC_TRY(launch_to_bf16(o->AT, o->ldN, o->AT16, o->ldN16, o->F_main, N, il, s));
  const int cA = il ? o->ldA16 : F, cN = il ? o->ldN16 : N;  // tensor-map columns (interleaved: every column is data or zero)
  EVC_TRY(make_tmap16(&o->tmA16, o->A16, o->a_rows, cA, o->ldA16, 64, 128));
  EVC_TRY(make_tmap16(&o->tmAT16, o->AT16, o->at_rows, cN, o->ldN16, 64, 128));
  EVC_TRY(make_tmap16(&o->tmAT16_s[0], o->AT16, o->at_rows, cN, o->ldN16, 64, 64));
  EVC_TRY(make_tmap16(&o->tmAT16_s[1], o->AT16, o->at_rows, cN, o->ldN16, 64, 32));
  if (B) {
    EVC_CUDA(cudaMalloc(&o->BT16, at_elems * sizeof(__nv_bfloat16)));
    EVC_CUDA(cudaMemsetAsync(o->BT16, 0, at_elems * sizeof(__nv_bfloat16), s));
    EVC_TRY(launch_to_bf16(o->BT, o->ldN, o->BT16, o->ldN16, o->F_main, N, il, s));
    EVC_TRY(make_tmap16(&o->tmBT16, o->BT16, o->at_rows, cN, o->ldN16, 64, 128));
    EVC_TRY(make_tmap16(&o->tmBT16_s[0], o->BT16, o->at_rows, cN, o->ldN16, 64, 64));
    EVC_TRY(make_tmap16(&o->tmBT16_s[1], o->BT16, o->at_rows, cN, o->ldN16, 64, 32));
  }
  // the fp32 transposes were staging only (stream order: the conversions above read them first)
  EVC_CUDA(cudaFreeAsync(o->AT, s));
  o->AT = nullptr;
  if (o->BT) { EVC_CUDA(cudaFreeAsync(o->BT, s)); o->BT = nullptr; }
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
// `pairs` = CTA pairs per cluster (they take `pairs` neighbouring frame tiles of one K range and share the dictionary
// tile): the plan is made over frame SUPER-tiles and `slots` = clusters resident at once.
inline C1Plan plan_c1(int F, int N, int T, int bk, int pairs = 1, int slots_in = 0) {
  C1Plan pl{};
  const int sub_rows = 128 * kCG;  // dictionary rows of one MMA
  const int tiles = ceil_div(F, sub_rows);
  pl.m_groups = ceil_div(tiles, kC1MTiles);
  pl.t_tiles = ceil_div(ceil_div(T, kC1BlockT), pairs);
  pl.kb_total = ceil_div(N, bk);
  pl.ldp = round_up(F, 32);
  const int tiles_last = tiles - (pl.m_groups - 1) * kC1MTiles;
  const bool partial = tiles_last < kC1MTiles;
  const int full_groups = partial ? pl.m_groups - 1 : pl.m_groups;
  pl.f_last = partial ? full_groups * kC1MTiles * sub_rows : F;
  // w = sub-tile K-blocks per CTA pair; grow it until the items fit the SMs
  const int slots = slots_in > 0 ? slots_in : num_sms() / (kCG * pairs);  // clusters resident at once
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

// CTA pairs per cluster of the two contractions (fp32-accurate mode; the fast modes run plain pairs).  Defaults are
// the measured best (profiles/r2_multicast_ab.txt); EVC_C1_PAIRS / EVC_C2_PAIRS = 1 | 2 | 4 override for A/B runs.
#ifndef EVC_C1_PAIRS_DEFAULT
#define EVC_C1_PAIRS_DEFAULT 1
#endif
#ifndef EVC_C2_PAIRS_DEFAULT
#define EVC_C2_PAIRS_DEFAULT 1
#endif
inline int c1_slots_for(int P) {
  if (P == 2) return TcLaunch<kC1MTiles, kC1BlockT, PREC_SPLIT, true, TEPI_PARTIAL, 2, true>::slots();
  if (P == 4) return TcLaunch<kC1MTiles, kC1BlockT, PREC_SPLIT, true, TEPI_PARTIAL, 4, true>::slots();
  return num_sms() / kCG;
}
inline int c1_pairs(int mode) {
  static const int want = env_pairs("EVC_C1_PAIRS", EVC_C1_PAIRS_DEFAULT);
  if (mode != EVC_MODE_3XTF32 || want == 1) return 1;
  return c1_slots_for(want) > 0 ? want : 1;  // a device that cannot co-schedule such clusters runs plain pairs
}
#ifndef EVC_C2_MTILES_DEFAULT
#define EVC_C2_MTILES_DEFAULT 1
#endif
// dictionary sub-tiles per work item of contraction 2 (fp32-accurate mode, plain pairs): EVC_C2_MTILES = 1 | 2
inline int c2_mtiles() {
  static const int v = getenv("EVC_C2_MTILES") ? atoi(getenv("EVC_C2_MTILES")) : EVC_C2_MTILES_DEFAULT;
  return v == 2 ? 2 : 1;
}
inline int c2_slots_for(int P) {
  if (P == 2) return TcLaunch<kC2MTiles, kC2BlockT, PREC_SPLIT, false, TEPI_MU_KL, 2, false>::slots();
  if (P == 4) return TcLaunch<kC2MTiles, kC2BlockT, PREC_SPLIT, false, TEPI_MU_KL, 4, false>::slots();
  return num_sms() / kCG;
}
inline int c2_pairs(int mode) {
  static const int want = env_pairs("EVC_C2_PAIRS", EVC_C2_PAIRS_DEFAULT);
  if (mode != EVC_MODE_3XTF32 || want == 1) return 1;
  return c2_slots_for(want) > 0 ? want : 1;
}
struct DictOperands;
inline C1Plan c1_plan(int F_main, int N, int T, int mode);

// Called before a solve / product: make sure the workspace can hold the split-K partials.
// Workspace layout: [ split-K partials | leftover-row partials of the fused update ].
inline size_t ws_left_offset(const C1Plan& pl, int T) { return round_up_sz((size_t)pl.max_splits * T * pl.ldp, 64); }
// One partial per 32 exemplars, for every 128-row block some CTA of contraction 2 works on -- with P pairs per
// cluster the row groups come in runs of P, so the padding groups of the last run write (zero) partials too; every
// entry the reduction sums is therefore written by every fused update.
inline int left_rows(const DictOperands& o, int mode) { return round_up(ceil_div(o.N, 128), kCG * c2_pairs(mode)) * 4; }
inline int left_ld(int T) { return round_up(T, kC2BlockT); }

// Room for the K operand of contraction 2 (the ratio) in the mode's format; pad rows start out zero.
inline int reserve_ratio(DictOperands& o, int mode, int T, cudaStream_t s) {
  if (mode != EVC_MODE_3XTF32 && mode != EVC_MODE_BF16) return EVC_OK;
  const long long rows = plane_rows(T);
  const size_t need = (size_t)rows * o.ldA16 * sizeof(__nv_bfloat16);  // (ldA16: bf16 pitch, or both interleaved planes)
  if (need > o.r16.bytes || rows != o.r_rows) {
    EVC_TRY(o.r16.reserve(need));
    EVC_CUDA(cudaMemsetAsync(o.r16.p, 0, o.r16.bytes, s));
    o.r_rows = rows;
  }
  return EVC_OK;
}
inline ROut ratio_out(DictOperands& o, int mode, float* R, int ldR) {
  ROut ro{};
  if (mode == EVC_MODE_3XTF32) {
    ro.R12 = o.r16.as<__nv_bfloat16>(); ro.ldr12 = o.ldA16; ro.k12 = round_up(o.F, 32);
  } else if (mode == EVC_MODE_BF16) {
    ro.R16 = o.r16.as<__nv_bfloat16>(); ro.ldr16 = o.ldA16;
  } else {
    ro.R = R; ro.ldr = ldR;
  }
  return ro;
}

inline C1Plan c1_plan(int F_main, int N, int T, int mode) {
  const int P = c1_pairs(mode);
  return plan_c1(F_main, N, T, bk_elems(mode), P, c1_slots_for(P));
}

inline int after_h_written(DictOperands& o, int mode, const float* H, int ldH, int T, DevBuf* ws, cudaStream_t s) {
  if (mode == EVC_MODE_FP32) return EVC_OK;
  const C1Plan pl = c1_plan(o.F_main, o.N, T, mode);
  o.left_valid = false;
  // leftover-row partials: one per 32 exemplars (fused update, fast modes) or one per K split (contraction 1)
  const size_t left = (size_t)std::max(left_rows(o, mode), pl.max_splits) * o.n_left * left_ld(T);
  EVC_TRY(ws->reserve((ws_left_offset(pl, T) + left) * sizeof(float)));
  EVC_TRY(reserve_ratio(o, mode, T, s));
  if (mode == EVC_MODE_BF16) {
    // the bf16 shadow of these activations (afterwards the fused update keeps it current)
    EVC_TRY(o.h16.reserve((size_t)T * o.ldN16 * sizeof(__nv_bfloat16)));
    EVC_TRY(launch_to_bf16(H, ldH, o.h16.as<__nv_bfloat16>(), o.ldN16, T, o.N, false, s));
  }
  return EVC_OK;
}

struct RatioArgs {  // fuse R = X / max(WH, eps) into the split-K reduction
  const float* X; int ldX; float* R; int ldR; float eps;
};

inline int launch_ratio(DictOperands& o, int mode, const float* X, int ldX, const float* WH, int ldWH, float* R, int ldR,
                        int T, int F, float eps, int copy, cudaStream_t s) {
  ProfScope ps(1, s);
  const ROut ro = ratio_out(o, mode, R, ldR);
  dim3 g(T, ceil_div(ro.cols(), 128));
  ratio_pad_kernel<<<g, 128, 0, s>>>(X, ldX, WH, ldWH, T, F, eps, copy, ro);
  EVC_LAUNCH_CHECK();
  return EVC_OK;
}

template <int kPrec, int kP = 1>
inline int contract_wh_t(DictOperands& o, int mode, const float* H, int ldH, int T, float* WH, int ldWH, bool target,
                         DevBuf* ws, cudaStream_t s, const RatioArgs* ra) {
  constexpr bool kSplitN = (kPrec == PREC_SPLIT);
  const int bke = bk_elems(mode);
  const C1Plan pl = c1_plan(o.F_main, o.N, T, mode);
  float* partials = ws->as<float>();
  float* leftp = ws->as<float>() + ws_left_offset(pl, T);
  CUtensorMap tmH;
  // the frame operand: the fp32 activations (split in shared memory / used as tf32), or their bf16 shadow
  if (kPrec == PREC_BF16) EVC_TRY(make_tmap16(&tmH, o.h16.as<__nv_bfloat16>(), T, o.N, o.ldN16, bke, kC1BlockT / kCG));
  else EVC_TRY(make_tmap(&tmH, H, T, o.N, ldH, 32, kC1BlockT / kCG));
  GemmParams p{};
  p.M_total = o.F_main; p.T = T; p.K = o.N;
  p.num_m_groups = pl.m_groups; p.num_t_tiles = pl.t_tiles; p.num_splits = pl.splits;
  p.kblocks_per_split = pl.kb_per_split; p.kblocks_total = pl.kb_total;
  p.splits_last = pl.splits_last; p.kblocks_per_split_last = pl.kb_per_split_last;
  p.items_main = (pl.splits_last ? pl.m_groups - 1 : pl.m_groups) * pl.t_tiles * pl.splits;
  p.half_from = p.items_main; p.tail_parts = 1;
  p.out = partials; p.ld_out = pl.ldp;
  p.out_keep_l2 = getenv("EVC_NO_KEEP_L2") ? 0 : 1;
  // leftover rows (F_main..): in the fp32-accurate mode contraction 1's split warps carry them along (per-split sums
  // in the workspace, summed with the partials); in the fast modes they come from the fused update's partials when
  // those describe this H, else from a dot-product pass over H
  const bool own_left = kSplitN && o.n_left > 0;
  const bool from_partials = !own_left && o.n_left > 0 && !target && o.left_valid;
  const bool standalone = !own_left && o.n_left > 0 && !from_partials;
  const int lp_splits = (pl.splits_last && pl.m_groups == 1) ? pl.splits_last : pl.splits;  // of dictionary-row group 0
  CUtensorMap tmL = tmH;
  if (own_left) {
    EVC_TRY(make_tmap(&tmL, target ? o.BTleft : o.ATleft, o.n_left, o.N, o.ldN, 32, 8, false));
    p.n_left = o.n_left; p.left_out = leftp; p.left_ld = left_ld(T);
  }
  const bool fuse = ra && !standalone;  // the ratio is formed in the same pass as the split-K sum
  const ROut ro = fuse ? ratio_out(o, mode, ra->R, ra->ldR) : ROut{};
  const int cols = std::max(ldWH, fuse ? ro.cols() : 0);
  const int lrows = from_partials ? left_rows(o, mode) : 0;
  // The sum (+ ratio) runs inside the contraction itself when every work item is resident at once (always the case
  // when K is split) and a unit's partials fit the staging slots the idle operand ring offers; else a second launch.
  using L1 = TcLaunch<kC1MTiles, kC1BlockT, kPrec, kSplitN, TEPI_PARTIAL, kP, true>;
  bool in_kernel = false, direct = false;
  CUtensorMap tmP = tmH;  // the partials as a (column, frame, split) tensor when the sum runs in the kernel
  {
    // EVC_NO_FUSED_REDUCE=1: always the separate pass.  EVC_FUSED_REDUCE=1: also when K is split (tile barrier + sum
    // by the contraction's own CTAs) -- bit-identical, but measured slower than the separate pass at the headline
    // shape (74 us against 59 + 11 us: fence + tile barrier cost ~7k cycles and 256 threads per SM hide the copy and
    // shared-memory latencies worse than eleven resident blocks of the reduction kernel), so it is not the default.
    static const bool allow = getenv("EVC_NO_FUSED_REDUCE") == nullptr;
    static const bool allow_split = allow && getenv("EVC_FUSED_REDUCE") != nullptr && atoi(getenv("EVC_FUSED_REDUCE")) != 0;
    const int items = p.items_main + p.splits_last * p.num_t_tiles;
    const size_t ring_floats = (size_t)L1::Cfg::kStages * L1::Cfg::kStageBytes / sizeof(float);
    const int n_chunks = ceil_div(pl.ldp, kRedCols);
    // frames per batch / slots: the most frames per box load such that two batches (double buffering) fit the ring
    int nb = 0, nslots = 0;
    size_t slot_floats = 0;
    for (int cand_slots = 2; cand_slots >= 1 && !nb; --cand_slots)
      for (int cand = 4; cand >= 1 && !nb; cand >>= 1) {
        const size_t sf = round_up_sz((size_t)n_chunks * pl.max_splits * cand * kRedCols + (size_t)o.n_left * cand * lrows, 32);
        if (cand_slots * sf <= ring_floats) { nb = cand; nslots = cand_slots; slot_floats = sf; }
      }
    FusedReduce& r = p.red;
    if (allow && pl.max_splits == 1 && o.F_main % kRedCols == 0) {
      // K is not split (many frames): A*H is complete in TMEM -- the epilogue writes the ratio (with `ra`) or A*H
      // (without) itself; the separate pass below only sees the columns past the tensor-core rows
      direct = true;
      r.enabled = 2;
      r.WH = fuse ? nullptr : WH; r.ldwh = ldWH;  // nobody reads A*H of an iteration whose ratio is formed here
      r.X = fuse ? ra->X : nullptr; r.ldx = fuse ? ra->ldX : 0; r.ro = ro;
      p.eps = fuse ? ra->eps : 0.f;
    } else if (allow_split && nb > 0 && items <= L1::slots() && pl.t_tiles * kP <= kRedMaxTiles && o.red_ctr.p && pl.max_splits <= 256) {
      in_kernel = true;
      EVC_TRY(make_tmap3d(&tmP, partials, pl.ldp, T, pl.max_splits, (size_t)pl.ldp, (size_t)T * pl.ldp, kRedCols, nb,
                          pl.max_splits));
      r.enabled = 1; r.counter = o.red_ctr.as<unsigned int>();
      r.contributors = ((pl.splits_last ? pl.m_groups - 1 : pl.m_groups) * pl.splits + pl.splits_last) * kCG;
      r.nslots = nslots; r.nb = nb; r.slot_floats = (int)slot_floats;
      r.S = pl.splits; r.S_last = pl.splits_last; r.f_last = pl.f_last; r.F = o.F;
      r.WH = WH; r.ldwh = ldWH;
      r.L = leftp; r.left_rows = lrows; r.left_ld = left_ld(T); r.n_left = o.n_left;
      r.Lp = own_left ? leftp : nullptr; r.lp_splits = lp_splits;
      r.X = fuse ? ra->X : nullptr; r.ldx = fuse ? ra->ldX : 0;
      r.cols = cols; r.ro = ro;
      p.eps = fuse ? ra->eps : 0.f;
    }
  }
  {
    ProfScope ps(0, s);
    const CUtensorMap& tmD = (kPrec == PREC_TF32) ? (target ? o.tmBT : o.tmAT)
                             : (kP == 1)          ? (target ? o.tmBT16 : o.tmAT16)
                                                  : (target ? o.tmBT16_s[kP / 4] : o.tmAT16_s[kP / 4]);
    EVC_TRY((L1::launch(tmD, tmH, tmH, tmP, tmL, p, s)));
  }
  if (!in_kernel && !(direct && ceil_div(cols, kRedCols) <= o.F_main / kRedCols)) {
    ProfScope ps(1, s);
    const int batch = std::max(1, std::min(kRedMaxBatch, pl.max_splits));
    const size_t smem = ((size_t)batch * kRedCols + (size_t)o.n_left * lrows) * sizeof(float);
    static bool configured[64] = {false};
    int dev = 0;
    EVC_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
      EVC_CUDA(cudaFuncSetAttribute(reduce_partials_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    if (smem > 160 * 1024) return fail(EVC_ERR_UNSUPPORTED, "split-K reduction: %zu bytes of staging exceed shared memory", smem);
    // (after the direct epilogue: only the column chunks past the tensor-core rows -- leftover rows and zero padding)
    const int chunk0 = direct ? o.F_main / kRedCols : 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(T, ceil_div(cols, kRedCols) - chunk0); cfg.blockDim = dim3(kRedThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = use_pdl() ? 1 : 0;
    CUtensorMap tmPb;
    EVC_TRY(make_tmap3d(&tmPb, partials, pl.ldp, T, pl.max_splits, (size_t)pl.ldp, (size_t)T * pl.ldp, kRedCols, 1, batch));
    EVC_CUDA(cudaLaunchKernelEx(&cfg, reduce_partials_kernel, tmPb, (const float*)partials, pl.splits, pl.splits_last, pl.f_last, T,
                                pl.ldp, o.F, o.F_main, WH, ldWH, (const float*)leftp, lrows, o.n_left, left_ld(T),
                                fuse ? ra->X : (const float*)nullptr, fuse ? ra->ldX : 0, fuse ? ra->eps : 0.f, ro, batch, chunk0,
                                own_left ? (const float*)leftp : (const float*)nullptr, lp_splits));
    EVC_LAUNCH_CHECK();
  }
  if (standalone) {
    ProfScope ps(1, s);
    const float* rows = (mode == EVC_MODE_TF32) ? (target ? o.BT : o.AT) + (size_t)o.F_main * o.ldN
                                                : (target ? o.BTleft : o.ATleft);
    leftover_rows_kernel<<<T, 256, 0, s>>>(H, ldH, T, o.N, rows, o.ldN, o.n_left, WH, ldWH, o.F_main);
    EVC_LAUNCH_CHECK();
  }
  if (ra && standalone) EVC_TRY(launch_ratio(o, mode, ra->X, ra->ldX, WH, ldWH, ra->R, ra->ldR, T, o.F, ra->eps, 0, s));
  return EVC_OK;
}

inline int contract_wh(DictOperands& o, int mode, const float* H, int ldH, int T, float* WH, int ldWH, bool target,
                       DevBuf* ws, cudaStream_t s, const RatioArgs* ra = nullptr) {
  if (target && !o.has_target) return fail(EVC_ERR_INVALID_ARGUMENT, "no target dictionary");
  if (mode == EVC_MODE_3XTF32) {
    const int P = c1_pairs(mode);
    if (P == 2) return contract_wh_t<PREC_SPLIT, 2>(o, mode, H, ldH, T, WH, ldWH, target, ws, s, ra);
    if (P == 4) return contract_wh_t<PREC_SPLIT, 4>(o, mode, H, ldH, T, WH, ldWH, target, ws, s, ra);
    return contract_wh_t<PREC_SPLIT>(o, mode, H, ldH, T, WH, ldWH, target, ws, s, ra);
  }
  if (mode == EVC_MODE_BF16) return contract_wh_t<PREC_BF16>(o, mode, H, ldH, T, WH, ldWH, target, ws, s, ra);
  return contract_wh_t<PREC_TF32>(o, mode, H, ldH, T, WH, ldWH, target, ws, s, ra);
}

// Second contraction with a fused epilogue.  `R` is what multiplies A^T: the ratio (KL) or A H (Frobenius); in the
// split / bf16 modes it lives in o.r16 (made by the reduction / ratio kernels).
// kMT = dictionary sub-tiles (of 256 exemplars) per work item.  kMT = 2 makes an item 512 exemplars x 256 frames: the
// two sub-tiles share every K-block of the ratio tile, so a pair ingests 3 operand tiles per 2 products instead of
// 4 (both GEMMs are bound by what an SM can ingest) -- at the price of all of TMEM for one item, i.e.
// no overlap of the fused update with the next item's MMAs.
template <int kPrec, int kEpi, int kP = 1, int kMT = kC2MTiles>
inline int contract2_p(DictOperands& o, int mode, int T, const float* R, int ldR, GemmParams p, const float* num0,
                       cudaStream_t s) {
  using Launch = TcLaunch<kMT, kC2BlockT, kPrec, false, kEpi, kP, false>;
  if (kP > 1 && Launch::slots() <= 0)  // this device cannot co-schedule such clusters: plain pairs
    return contract2_p<kPrec, kEpi, 1, kMT>(o, mode, T, R, ldR, p, num0, s);
  const int bke = bk_elems(mode);
  CUtensorMap tmR;
  // (clusters of kP pairs: every CTA fetches a 1/kP slice of its half of the frame tile and multicasts it)
  if (kPrec == PREC_TF32) EVC_TRY(make_tmap(&tmR, R, T, o.F, ldR, 32, kC2BlockT / kCG));
  else EVC_TRY(make_tmap16(&tmR, o.r16.as<__nv_bfloat16>(), (long long)T, kPrec == PREC_SPLIT ? o.ldA16 : o.F, o.ldA16, 64,
                           kC2BlockT / kCG / kP));
  p.M_total = o.N; p.T = T; p.K = o.F;
  // row groups of 256 exemplars, dealt to the pairs of a cluster in runs of kP: the work items are cluster-level
  p.num_m_groups = ceil_div(ceil_div(o.N, 128 * kMT * kCG), kP); p.num_t_tiles = ceil_div(T, kC2BlockT); p.num_splits = 1;
  p.kblocks_total = ceil_div(o.F, bke); p.kblocks_per_split = p.kblocks_total;
  p.items_main = p.num_m_groups * p.num_t_tiles; p.splits_last = 0; p.kblocks_per_split_last = 0;
  {
    // tail balancing: the tiles of a last, partly filled round run as 2-4 narrower items each
    static const bool allow = getenv("EVC_NO_HALF_TILES") == nullptr;
    static const int max_parts = getenv("EVC_TAIL_PARTS") ? std::max(2, atoi(getenv("EVC_TAIL_PARTS"))) : 4;
    plan_tail(p, Launch::slots(), kC2BlockT, allow, max_parts);
  }
  // neighbouring CTA pairs update neighbouring 512-byte runs of the same H rows: DRAM pages stay open
  p.m_fastest = 1;
  {
    static const bool direct = getenv("EVC_C2_DIRECT_STORE") != nullptr && atoi(getenv("EVC_C2_DIRECT_STORE")) != 0;
    p.direct_store = direct ? 1 : 0;
  }
  CUtensorMap tmHc = tmR, tmQc = tmR, tmSc = tmR;  // the fused updates stage H (and the Frobenius numerator) through shared memory
  if (kPrec == PREC_BF16 && kEpi != TEPI_PARTIAL)  // ... and the bf16 shadow of H leaves the same way
    EVC_TRY(make_tmap_any(&tmSc, o.h16.as<__nv_bfloat16>(), 2, T, o.N, o.ldN16, 128, kHChunkT, false));
  if (kEpi != TEPI_PARTIAL) EVC_TRY(make_tmap(&tmHc, p.out, T, o.N, p.ld_out, 128, kHChunkT, false));
  if (kEpi == TEPI_MU_FRO) EVC_TRY(make_tmap(&tmQc, num0, T, o.N, p.ld_out, 128, kHChunkT, false));
  ProfScope ps(2, s);
  return Launch::launch(kPrec == PREC_TF32 ? o.tmA : o.tmA16, tmR, tmHc, tmQc, tmSc, p, s);
}

template <int kEpi>
inline int contract2_t(DictOperands& o, int mode, int T, const float* R, int ldR, const GemmParams& p, const float* num0,
                       cudaStream_t s) {
  if (mode == EVC_MODE_3XTF32) {
    const int P = c2_pairs(mode);
    if (P == 2) return contract2_p<PREC_SPLIT, kEpi, 2>(o, mode, T, R, ldR, p, num0, s);
    if (P == 4) return contract2_p<PREC_SPLIT, kEpi, 4>(o, mode, T, R, ldR, p, num0, s);
    if (c2_mtiles() == 2) return contract2_p<PREC_SPLIT, kEpi, 1, 2>(o, mode, T, R, ldR, p, num0, s);
    return contract2_p<PREC_SPLIT, kEpi>(o, mode, T, R, ldR, p, num0, s);
  }
  if (mode == EVC_MODE_BF16) return contract2_p<PREC_BF16, kEpi>(o, mode, T, R, ldR, p, num0, s);
  return contract2_p<PREC_TF32, kEpi>(o, mode, T, R, ldR, p, num0, s);
}

inline void left_args(DictOperands& o, int mode, int T, DevBuf* ws, GemmParams& p) {
  if (o.n_left <= 0 || mode == EVC_MODE_3XTF32) return;  // (fp32-accurate mode: contraction 1 carries those rows itself)
  const C1Plan pl = c1_plan(o.F_main, o.N, T, mode);
  p.left_a = (mode == EVC_MODE_TF32) ? o.AT + (size_t)o.F_main * o.ldN : o.ATleft;
  p.left_lda = o.ldN; p.n_left = o.n_left;
  p.left_out = ws->as<float>() + ws_left_offset(pl, T); p.left_ld = left_ld(T); p.left_rows = left_rows(o, mode);
  o.left_valid = true;  // (stream order: the partials are complete before the next contraction 1 reads them)
}

inline int update_kl(DictOperands& o, int mode, const float* X, int ldX, int T, const float* WH, int ldWH, float* R,
                     int ldR, float* H, int ldH, const float* colsum, float lam, float eps,
                     const unsigned char* row_active, DevBuf* ws, cudaStream_t s, bool ratio_done = false) {
  if (!ratio_done) EVC_TRY(launch_ratio(o, mode, X, ldX, WH, ldWH, R, ldR, T, o.F, eps, 0, s));
  GemmParams p{};
  p.out = H; p.ld_out = ldH;
  p.colsum = colsum; p.lam = lam; p.eps = eps; p.row_active = row_active;
  left_args(o, mode, T, ws, p);
  return contract2_t<TEPI_MU_KL>(o, mode, T, R, ldR, p, nullptr, s);
}

inline int update_fro(DictOperands& o, int mode, int T, const float* WH, int ldWH, float* R, int ldR, float* H, int ldH,
                      const float* num0, float lam, float eps, const unsigned char* row_active, DevBuf* ws,
                      cudaStream_t s) {
  EVC_TRY(launch_ratio(o, mode, nullptr, 0, WH, ldWH, R, ldR, T, o.F, eps, 1, s));
  GemmParams p{};
  p.out = H; p.ld_out = ldH;
  p.lam = lam; p.eps = eps; p.row_active = row_active;
  left_args(o, mode, T, ws, p);
  return contract2_t<TEPI_MU_FRO>(o, mode, T, R, ldR, p, num0, s);
}

// NUM0 (T, ldH) = X A^T : the second contraction with a plain store ([split=0][t][n] layout == (T, ldH)).
inline int frob_numerator(DictOperands& o, int mode, const float* X, int ldX, int T, float* R, int ldR, float* num0,
                          int ldH, DevBuf* ws, cudaStream_t s) {
  // stage X into the zero-padded K-operand buffer (copy mode of the ratio kernel with WH := X)
  EVC_TRY(reserve_ratio(o, mode, T, s));
  EVC_TRY(launch_ratio(o, mode, nullptr, 0, X, ldX, R, ldR, T, o.F, 0.f, 1, s));
  GemmParams p{};
  p.out = num0; p.ld_out = ldH;
  return contract2_t<TEPI_PARTIAL>(o, mode, T, R, ldR, p, nullptr, s);
}

}  // namespace tc
}  // namespace evc
