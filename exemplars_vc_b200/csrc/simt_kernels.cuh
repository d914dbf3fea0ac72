// CUDA-core kernels: the exact-fp32 mode (EVC_MODE_FP32) of the two contractions, plus the
// memory-bound pieces every mode shares (A^T 1, H0 fill, ratio, objective rows).
//
// Reference arithmetic restated here (sklearn/decomposition/_nmf.py, 1.9.0):
//   WH = W@A ; WH[WH<eps]=eps ; R = X/WH                    :554-571
//   num = R@A.T ; den = A.sum(1) (+l1) ; den[den==0]=eps     :585-615
//   W *= num/den                                             :617-624
//   objective                                                :139-154
#pragma once
#include "evc_common.cuh"

namespace evc {
namespace simt {

enum Epilogue { EPI_STORE = 0, EPI_RATIO = 1, EPI_MU_KL = 2, EPI_MU_FRO = 3 };

struct EpiArgs {
  float* C;                         // STORE/RATIO: output (M,N); MU_*: the activations H, updated in place
  int ldc;
  const float* X;                   // RATIO: the frames X; MU_FRO: the cached numerator X A^T
  int ldx;
  const float* colsum;              // MU_KL: A^T 1
  float lam;                        // penalty added to the denominator this iteration
  float eps;
  const unsigned char* row_active;  // MU_*: frames of utterances that already stopped are left alone
};

constexpr int BM = 64, BN = 64, BK = 16;

// C[m,n] = sum_k P[m*ldp + k] * Q(k,n);  Q(k,n) = QT ? Q[n*ldq + k] : Q[k*ldq + n]
template <int EPI, bool QT>
__global__ void __launch_bounds__(256)
gemm_kernel(int M, int N, int K, const float* __restrict__ P, int ldp, const float* __restrict__ Q, int ldq,
            EpiArgs e) {
  __shared__ float Ps[BK][BM + 4];
  __shared__ float Qs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    {  // P tile: k contiguous in memory
      const int m = tid >> 2, kq = (tid & 3) * 4;
      const int gm = m0 + m;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gk = k0 + kq + i;
        Ps[kq + i][m] = (gm < M && gk < K) ? P[(size_t)gm * ldp + gk] : 0.f;
      }
    }
    if (QT) {  // Q tile: k contiguous
      const int n = tid >> 2, kq = (tid & 3) * 4;
      const int gn = n0 + n;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gk = k0 + kq + i;
        Qs[kq + i][n] = (gn < N && gk < K) ? Q[(size_t)gn * ldq + gk] : 0.f;
      }
    } else {  // Q tile: n contiguous
      const int k = tid >> 4, nq = (tid & 15) * 4;
      const int gk = k0 + k;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int gn = n0 + nq + i;
        Qs[k][nq + i] = (gn < N && gk < K) ? Q[(size_t)gk * ldq + gn] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&Ps[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Qs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    if ((EPI == EPI_MU_KL || EPI == EPI_MU_FRO) && e.row_active && !e.row_active[m]) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      const float v = acc[i][j];
      float* c = e.C + (size_t)m * e.ldc + n;
      if (EPI == EPI_STORE) {
        *c = v;
      } else if (EPI == EPI_RATIO) {
        *c = __fdiv_rn(e.X[(size_t)m * e.ldx + n], fmaxf(v, e.eps));
      } else if (EPI == EPI_MU_KL) {
        float den = e.colsum[n] + e.lam;
        if (den == 0.f) den = e.eps;
        *c = *c * __fdiv_rn(v, den);
      } else {
        float den = v + e.lam;
        if (den == 0.f) den = e.eps;
        *c = *c * __fdiv_rn(e.X[(size_t)m * e.ldx + n], den);
      }
    }
  }
}

template <int EPI, bool QT>
inline int launch_gemm(int M, int N, int K, const float* P, int ldp, const float* Q, int ldq, const EpiArgs& e,
                       cudaStream_t s) {
  if (M <= 0 || N <= 0) return EVC_OK;
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM));
  gemm_kernel<EPI, QT><<<grid, 256, 0, s>>>(M, N, K, P, ldp, Q, ldq, e);
  EVC_LAUNCH_CHECK();
  return EVC_OK;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per row: out[r] = sum_c M[r*ld + c], accumulated in double (used for mean(X)).
__global__ void row_sum_kernel(const float* __restrict__ Mx, int ld, int rows, int cols, double* __restrict__ out) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for (int c = lane; c < cols; c += 32) s += (double)Mx[(size_t)r * ld + c];
  s = warp_sum(s);
  if (lane == 0) out[r] = s;
}

// A^T 1 (sklearn's H_sum) + validation flags: flags[0] = any negative entry, flags[1] = any non-zero entry.
// One warp per exemplar.  fp32 pairwise-ish accumulation in double then rounded once to fp32.
__global__ void colsum_kernel(const float* __restrict__ A, int lda, int N, int F, float* __restrict__ colsum,
                              int* __restrict__ flags) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  bool neg = false, nz = false;
  for (int f = lane; f < F; f += 32) {
    const float a = A[(size_t)n * lda + f];
    neg |= (a < 0.f) || (a != a);
    nz |= (a != 0.f);
    s += (double)a;
  }
  s = warp_sum(s);
  if (lane == 0) colsum[n] = (float)s;
  if (__any_sync(0xffffffffu, neg) && lane == 0) atomicOr(&flags[0], 1);
  if (__any_sync(0xffffffffu, nz) && lane == 0) atomicOr(&flags[1], 1);
}

// H[t, :] = w0[t]  (sklearn _nmf.py:1225-1226, one value per utterance).  One block per frame, 16-byte stores when the
// row is 16-byte aligned (it is for every H this library allocates): an 80 MB fill runs at HBM write speed instead of
// being bound by the launch shape of 79 000 one-float-per-thread blocks (profiles/r1_ncu_memory_bound_summary.txt).
__global__ void __launch_bounds__(256) fill_rows_kernel(float* __restrict__ H, int ldh, int T, int N, const float* __restrict__ w0) {
  const int t = blockIdx.x;
  if (t >= T) return;
  const float v = w0[t];
  float* row = H + (size_t)t * ldh;
  if ((((uintptr_t)row) & 15) == 0) {
    const int n4 = N >> 2;
    const float4 v4 = make_float4(v, v, v, v);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) reinterpret_cast<float4*>(row)[i] = v4;
    for (int n = (n4 << 2) + threadIdx.x; n < N; n += blockDim.x) row[n] = v;
  } else {
    for (int n = threadIdx.x; n < N; n += blockDim.x) row[n] = v;
  }
}

// R = X / max(WH, eps)   (sklearn _nmf.py:568-571); columns [F, ldr) of R are zeroed so that R can be a
// zero-padded K operand of the second contraction.
__global__ void ratio_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ WH, int ldwh,
                             float* __restrict__ R, int ldr, int T, int F, float eps) {
  const int f = blockIdx.y * blockDim.x + threadIdx.x;
  const int t = blockIdx.x;
  if (t >= T || f >= ldr) return;
  float r = 0.f;
  if (f < F) r = __fdiv_rn(X[(size_t)t * ldx + f], fmaxf(WH[(size_t)t * ldwh + f], eps));
  R[(size_t)t * ldr + f] = r;
}

// Row terms of the objective, in double.  loss == KL (sklearn _nmf.py:139-154):
//   sum_{X>eps} X log(X / max(WH,eps)) - sum_{X>eps} X + sum_all WH
// (sklearn forms sum_all WH as dot(W.sum(0), A.sum(1)); it is the same number up to rounding.)
// loss == FROBENIUS: sum (X - WH)^2.
// Four warps per frame (a double-precision log per element: with one warp per frame a 513-bin frame was a serial
// chain of 16 logs per lane and the kernel took 17 us, profiles/r2_ncu_launch_shares.txt); the four warp sums are added in a fixed order.
constexpr int kObjWarpsPerRow = 4, kObjRowsPerBlock = 2;
__global__ void __launch_bounds__(kObjWarpsPerRow * kObjRowsPerBlock * 32)
objective_rows_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ WH, int ldwh, int T, int F,
                      float eps, int loss, double* __restrict__ rowobj) {
  __shared__ double part[kObjRowsPerBlock][kObjWarpsPerRow];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = warp / kObjWarpsPerRow, w = warp % kObjWarpsPerRow;
  const int t = blockIdx.x * kObjRowsPerBlock + r;
  double s = 0.0;
  if (t < T) {
    for (int f = w * 32 + lane; f < F; f += 32 * kObjWarpsPerRow) {
      const float x = X[(size_t)t * ldx + f];
      const float wh = WH[(size_t)t * ldwh + f];
      if (loss == EVC_LOSS_KL) {
        double term = (double)wh;
        if (x > eps) {
          const double whc = (double)fmaxf(wh, eps);
          term += (double)x * log((double)x / whc) - (double)x;
        }
        s += term;
      } else {
        const double d = (double)x - (double)wh;
        s += d * d;
      }
    }
  }
  s = warp_sum(s);
  if (lane == 0) part[r][w] = s;
  __syncthreads();
  if (w == 0 && lane == 0 && t < T) rowobj[t] = (part[r][0] + part[r][1]) + (part[r][2] + part[r][3]);
}

// dst (rows, ld_dst) <- src (rows, ld_src), zero-filling the pad columns [cols, ld_dst).
__global__ void repitch_kernel(const float* __restrict__ src, int ld_src, float* __restrict__ dst, int ld_dst,
                               int rows, int cols) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  const int r = blockIdx.x;
  if (r >= rows || c >= ld_dst) return;
  dst[(size_t)r * ld_dst + c] = (c < cols) ? src[(size_t)r * ld_src + c] : 0.f;
}

// dst (cols, ld_dst) <- transpose of src (rows, ld_src); pad columns [rows, ld_dst) zeroed.
__global__ void transpose_kernel(const float* __restrict__ src, int ld_src, float* __restrict__ dst, int ld_dst,
                                 int rows, int cols) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(size_t)r * ld_src + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;  // dst row = c, dst col = r
    if (c < cols && r < ld_dst) dst[(size_t)c * ld_dst + r] = tile[threadIdx.x][i];
  }
}

// out[k, (d+c)*F + f] = frames[clamp(idx[k]+d, lo[k], hi[k]-1), f]: aligned-frame gather + context stacking.
__global__ void gather_stack_kernel(const float* __restrict__ frames, int ld, int F, const int* __restrict__ idx,
                                    const int* __restrict__ lo, const int* __restrict__ hi, int n_out, int context,
                                    float* __restrict__ out, int ld_out) {
  const int k = blockIdx.x;
  if (k >= n_out) return;
  const int width = (2 * context + 1) * F;
  const int base = idx[k], l = lo[k], h = hi[k] - 1;
  for (int c = blockIdx.y * blockDim.x + threadIdx.x; c < width; c += gridDim.y * blockDim.x) {
    const int d = c / F - context, f = c - (c / F) * F;
    int r = base + d;
    r = r < l ? l : (r > h ? h : r);
    out[(size_t)k * ld_out + c] = frames[(size_t)r * ld + f];
  }
}

}  // namespace simt
}  // namespace evc
