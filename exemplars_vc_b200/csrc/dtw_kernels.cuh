// DTW alignment of parallel utterances on the device: the producer of the exemplar index paths
// (01_make_dict_parallel.py:215-249, `dtw(feat_A.T, feat_B.T, lambda x, y: sum(np.square(x - y)))` per file pair in a
// multiprocessing.Pool -- the README's "this consumes most of the time": 70 s for 20 files in the reference's logs).
//
// The arithmetic is that of the un-vendored `dtw` package the reference imports (pierre-rouanet/dtw, dtw.dtw with its
// default step pattern, warp = 1):
//   D[i][j] = c(i,j) + min(D[i-1][j-1], D[i-1][j], D[i][j-1])   with D[-1][-1] = 0 and an infinite border,
//   c(i,j)  = sum_d (a[i][d] - b[j][d])^2 accumulated left to right in float64 (python's sum over np.square),
//   traceback from (r-1, c-1): argmin over (diagonal, up, left), the FIRST minimum wins, until (0, 0).
// One block per file pair sweeps the anti-diagonals (three of them live in shared memory), recording for every cell
// the direction the traceback would take (same argmin, same tie-break), so only one byte per cell goes to global
// memory; one thread per pair then walks the directions.  Double precision, no FMA contraction in the cost, so the
// accumulated costs -- and therefore the paths -- are bit-identical to the package's.
#pragma once
#include "evc_common.cuh"
#include <cmath>

namespace evc {
namespace dtw {

constexpr int kThreads = 512;

// A (sum_r, dim), B (sum_c, dim) row-major doubles; file f owns rows [a_off[f], a_off[f+1]) / [b_off[f], b_off[f+1]).
// dirs: r*c bytes per file at dir_off[f] (0 = diagonal, 1 = up (i-1), 2 = left (j-1)).  dist[f] = D[r-1][c-1] / (r + c).
__global__ void __launch_bounds__(kThreads)
forward_kernel(const double* __restrict__ A, const long long* __restrict__ a_off, const double* __restrict__ B,
               const long long* __restrict__ b_off, int dim, unsigned char* __restrict__ dirs,
               const long long* __restrict__ dir_off, double* __restrict__ dist, int max_diag) {
  extern __shared__ double dsm[];  // three anti-diagonals of max_diag cells each, indexed by i
  const int f = blockIdx.x;
  const long long a0 = a_off[f], b0 = b_off[f];
  const int r = (int)(a_off[f + 1] - a0), c = (int)(b_off[f + 1] - b0);
  if (r <= 0 || c <= 0) { if (threadIdx.x == 0) dist[f] = 0.0; return; }
  unsigned char* dr = dirs + dir_off[f];
  double* d0 = dsm;                 // diagonal k
  double* d1 = dsm + max_diag;      // diagonal k-1
  double* d2 = dsm + 2 * max_diag;  // diagonal k-2
  const double inf = INFINITY;
  for (int k = 0; k < r + c - 1; ++k) {
    const int i_lo = max(0, k - (c - 1)), i_hi = min(r - 1, k);
    for (int i = i_lo + (int)threadIdx.x; i <= i_hi; i += blockDim.x) {
      const int j = k - i;
      const double* a = A + (a0 + i) * dim;
      const double* b = B + (b0 + j) * dim;
      double cost = 0.0;
      for (int d = 0; d < dim; ++d) {
        const double e = __dsub_rn(a[d], b[d]);
        cost = __dadd_rn(cost, __dmul_rn(e, e));
      }
      // predecessors: (i-1, j-1) on diagonal k-2, (i-1, j) and (i, j-1) on diagonal k-1 (indexed by their own i)
      const double diag = (i > 0 && j > 0) ? d2[i - 1] : ((i == 0 && j == 0) ? 0.0 : inf);
      const double up = (i > 0) ? d1[i - 1] : inf;
      const double left = (j > 0) ? d1[i] : inf;
      double m = diag;
      unsigned char dir = 0;
      if (up < m) { m = up; dir = 1; }
      if (left < m) { m = left; dir = 2; }
      d0[i] = __dadd_rn(cost, m);
      dr[(size_t)i * c + j] = dir;
    }
    __syncthreads();
    double* t = d2; d2 = d1; d1 = d0; d0 = t;
  }
  if (threadIdx.x == 0) dist[f] = d1[r - 1] / (double)(r + c);   // (after the last swap the final diagonal is d1)
}

// One thread per file: walk the directions back from (r-1, c-1); the path is written in REVERSE into
// path_a/path_b[path_off[f] ...] (at most r + c - 1 entries), its length into path_len[f].
__global__ void traceback_kernel(const long long* __restrict__ a_off, const long long* __restrict__ b_off,
                                 const unsigned char* __restrict__ dirs, const long long* __restrict__ dir_off,
                                 int* __restrict__ path_a, int* __restrict__ path_b,
                                 const long long* __restrict__ path_off, int* __restrict__ path_len, int n_files) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_files) return;
  const int r = (int)(a_off[f + 1] - a_off[f]), c = (int)(b_off[f + 1] - b_off[f]);
  if (r <= 0 || c <= 0) { path_len[f] = 0; return; }
  const unsigned char* dr = dirs + dir_off[f];
  int* pa = path_a + path_off[f];
  int* pb = path_b + path_off[f];
  int i = r - 1, j = c - 1, n = 0;
  pa[n] = i; pb[n] = j; ++n;
  while (i > 0 || j > 0) {
    const unsigned char d = dr[(size_t)i * c + j];
    if (d == 0) { --i; --j; } else if (d == 1) { --i; } else { --j; }
    pa[n] = i; pb[n] = j; ++n;
  }
  path_len[f] = n;
}

}  // namespace dtw
}  // namespace evc
