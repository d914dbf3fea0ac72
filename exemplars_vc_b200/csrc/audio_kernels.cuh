// The step right after the activation path: the WORLD-branch residual compensation of
// 04_align_n_nmf.py:292-294 / 363-373 as elementwise epilogues of the two products, and the Griffin-Lim vocoder of
// the STFT branch (zz_audio_utilities.py:181-292; 300 iterations of STFT -> keep the phase, impose the magnitude ->
// inverse STFT -> overlap-add, called at 04_align_n_nmf.py:187).
//
// Griffin-Lim runs in DOUBLE precision like the reference's numpy code (the iteration is a fixed-point search over
// phases; the reference seeds it with noise, np.random.randn, so parity is defined for a given start signal x0).
// The reference's frame length is 400 = 2^4 * 5^2 (04_align_n_nmf.py:46), not a power of two: each frame's DFT is
// done directly against a twiddle table in shared memory -- O(fft^2) per frame, 0.6 MFMA, fully parallel over the T
// frames -- and STFT -> phase -> inverse STFT of ONE frame happen inside one block (frames are independent until the
// overlap-add), so the spectrogram never exists in global memory during the iteration.
#pragma once
#include "evc_common.cuh"
#include <cmath>

namespace evc {
namespace audio {

// ---- WORLD-branch residual (04_align_n_nmf.py:292-294, 363-373) ------------------------------------------------
// R = log(H^T A - X): NaN where H^T A < X, -inf where equal (IEEE log), exactly like np.log on the difference.
__global__ void residual_kernel(const float* __restrict__ WH, int ldwh, const float* __restrict__ X, int ldx,
                                float* __restrict__ R, int ldr, int T, int F) {
  const int f = blockIdx.y * blockDim.x + threadIdx.x, t = blockIdx.x;
  if (t >= T || f >= F) return;
  R[(size_t)t * ldr + f] = logf(WH[(size_t)t * ldwh + f] - X[(size_t)t * ldx + f]);
}
// converted = exp(log(H^T B) + log(r)) with r = residual after `residual[isnan(residual)] = 0` (:363-365).
// In exact arithmetic that is (H^T B) * r for r > 0, 0 for r == 0 (log 0 = -inf, exp(-inf) = 0) and NaN for r < 0
// (log of a negative number) -- formed as the product, which is closer to the float64 reference than exp(log + log)
// evaluated in fp32.
__global__ void apply_residual_kernel(const float* __restrict__ Yin, int ldyin, const float* __restrict__ R, int ldr,
                                      float* __restrict__ Y, int ldy, int T, int F) {
  const int f = blockIdx.y * blockDim.x + threadIdx.x, t = blockIdx.x;
  if (t >= T || f >= F) return;
  const float y = Yin[(size_t)t * ldyin + f];
  float r = R[(size_t)t * ldr + f];
  if (r != r) r = 0.f;
  Y[(size_t)t * ldy + f] = (r < 0.f) ? nanf("") : y * r;
}

// ---- Griffin-Lim ----------------------------------------------------------------------------------------------------
enum GlMode { GL_ITERATE = 0, GL_STFT = 1, GL_ISTFT = 2 };

// One block per frame t.
//   GL_ITERATE: frame t of x (windowed) -> DFT -> proposal = mag[t] * exp(i angle) -> inverse DFT -> window -> frames[t]
//   GL_STFT:    frame t of x (windowed) -> DFT -> spec[t]  (zz_audio_utilities.py:181-196)
//   GL_ISTFT:   spec[t] -> inverse DFT -> window -> frames[t]  (zz_audio_utilities.py:199-218, before the overlap-add)
// spec is (T, bins, 2) doubles (re, im), bins = fft/2 + 1; fft must be even.
__global__ void __launch_bounds__(256)
gl_frame_kernel(int mode, const double* __restrict__ x, const double* __restrict__ window, const float* __restrict__ mag,
                int ldm, double* __restrict__ spec, int T, int fft, int hop, double* __restrict__ frames) {
  extern __shared__ double gl_sm[];
  const int bins = fft / 2 + 1;
  double* tw_cos = gl_sm;            // cos(2 pi n / fft)
  double* tw_sin = tw_cos + fft;     // sin(2 pi n / fft)
  double* frame = tw_sin + fft;      // windowed samples
  double* s_re = frame + fft;        // spectrum of the proposal
  double* s_im = s_re + bins;
  const int t = blockIdx.x;
  for (int n = threadIdx.x; n < fft; n += blockDim.x) {
    double s, c;
    sincospi(2.0 * (double)n / (double)fft, &s, &c);
    tw_cos[n] = c; tw_sin[n] = s;
    if (mode != GL_ISTFT) frame[n] = window[n] * x[(size_t)t * hop + n];
  }
  if (mode == GL_ISTFT)
    for (int k = threadIdx.x; k < bins; k += blockDim.x) {
      s_re[k] = spec[((size_t)t * bins + k) * 2];
      s_im[k] = spec[((size_t)t * bins + k) * 2 + 1];
    }
  __syncthreads();
  if (mode != GL_ISTFT) {
    for (int k = threadIdx.x; k < bins; k += blockDim.x) {
      // rfft: X_k = sum_n x_n exp(-2 pi i k n / fft)
      double re = 0.0, im = 0.0;
      int idx = 0;
      for (int n = 0; n < fft; ++n) {
        re = fma(frame[n], tw_cos[idx], re);
        im = fma(-frame[n], tw_sin[idx], im);
        idx += k;
        if (idx >= fft) idx -= fft;
      }
      if (mode == GL_STFT) {
        spec[((size_t)t * bins + k) * 2] = re;
        spec[((size_t)t * bins + k) * 2 + 1] = im;
      } else {
        // keep the phase, impose the magnitude (np.angle(0) = 0, so a zero bin becomes the real magnitude)
        const double m = hypot(re, im), M = (double)mag[(size_t)t * ldm + k];
        s_re[k] = m > 0.0 ? M * (re / m) : M;
        s_im[k] = m > 0.0 ? M * (im / m) : 0.0;
      }
    }
    if (mode == GL_STFT) return;
    __syncthreads();
  }
  // irfft (numpy ignores the imaginary parts of bins 0 and fft/2):
  //   x_n = (1/fft) [ X_0 + (-1)^n X_{fft/2} + 2 sum_{k=1}^{fft/2-1} (Re X_k cos(2 pi k n/fft) - Im X_k sin(2 pi k n/fft)) ]
  const double inv = 1.0 / (double)fft;
  for (int n = threadIdx.x; n < fft; n += blockDim.x) {
    double acc = s_re[0];
    int idx = 0;
    for (int k = 1; k < bins - 1; ++k) {
      idx += n;
      if (idx >= fft) idx -= fft;
      acc = fma(2.0 * s_re[k], tw_cos[idx], acc);
      acc = fma(-2.0 * s_im[k], tw_sin[idx], acc);
    }
    acc += (n & 1) ? -s_re[bins - 1] : s_re[bins - 1];
    frames[(size_t)t * fft + n] = window[n] * (acc * inv);
  }
}

// x_new[s] = sum over the frames that cover sample s, in ascending frame order (the order of the reference's loop,
// zz_audio_utilities.py:216-217); optionally accumulates sum (x_new - x_prev)^2 for the RMSE the reference prints.
__global__ void gl_overlap_add_kernel(const double* __restrict__ frames, int T, int fft, int hop, long long len,
                                      double* __restrict__ x_new, const double* __restrict__ x_prev,
                                      double* __restrict__ sq_diff) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double d2 = 0.0;
  if (s < len) {
    long long t_hi = s / hop;
    if (t_hi > T - 1) t_hi = T - 1;
    const long long t_lo = (s - fft + 1 <= 0) ? 0 : (s - fft + 1 + hop - 1) / hop;  // smallest t with s - t*hop < fft
    double acc = 0.0;
    for (long long t = t_lo; t <= t_hi; ++t) acc += frames[(size_t)t * fft + (size_t)(s - t * hop)];
    if (x_prev && sq_diff) { const double d = acc - x_prev[s]; d2 = d * d; }
    x_new[s] = acc;
  }
  if (sq_diff) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    if ((threadIdx.x & 31) == 0 && d2 != 0.0) atomicAdd(sq_diff, d2);
  }
}

inline size_t gl_smem(int fft) { return (size_t)(3 * fft + 2 * (fft / 2 + 1)) * sizeof(double); }

inline int gl_check(int T, int fft, int hop) {
  if (T < 1 || fft < 2 || (fft & 1) || hop < 1 || hop > fft)
    return fail(EVC_ERR_INVALID_ARGUMENT, "Griffin-Lim: need T >= 1, an even fft_size >= 2 and 1 <= hop <= fft_size");
  if (gl_smem(fft) > 200 * 1024) return fail(EVC_ERR_UNSUPPORTED, "Griffin-Lim: fft_size %d does not fit in shared memory", fft);
  return EVC_OK;
}

inline int gl_launch_frames(int mode, const double* x, const double* window, const float* mag, int ldm, double* spec,
                            int T, int fft, int hop, double* frames, cudaStream_t s) {
  static bool configured[64] = {false};
  int dev = 0;
  EVC_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    EVC_CUDA(cudaFuncSetAttribute(gl_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  gl_frame_kernel<<<T, 256, gl_smem(fft), s>>>(mode, x, window, mag, ldm, spec, T, fft, hop, frames);
  EVC_LAUNCH_CHECK();
  return EVC_OK;
}

}  // namespace audio
}  // namespace evc
