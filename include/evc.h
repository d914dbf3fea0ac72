/*
 * evc.h -- C ABI of the B200-native activation-estimation path (libevc_b200.so).
 *
 * The reference (entn-at/exemplars_vc) has no FFI: its boundary is the Python call
 *
 *   _W, _H, n_iter = non_negative_factorization(X=X, H=W, init="custom", update_H=False,
 *                        n_components=W.shape[0], beta_loss=..., solver='mu', tol=tol,
 *                        max_iter=150)                        04_align_n_nmf.py:212-213
 *   converted = np.matmul(H.T, B)                             04_align_n_nmf.py:391
 *
 * The entry points below are what a binding for that seam calls (INTEGRATION.md shows the
 * ctypes stub).  Conventions: plain C symbols, int status return (0 = ok), no exceptions
 * cross the boundary, every matrix is ROW-MAJOR IN THE REFERENCE'S ORIENTATION (frames and
 * exemplars are rows):
 *
 *     X (T,F) frames          == conv_sp / conv_stft      04_align_n_nmf.py:231,315
 *     A (N,F) source dict     == A_sp / W of _factorize   04_align_n_nmf.py:242,194
 *     B (N,F) target dict     == B_sp / B_stft            04_align_n_nmf.py:359,390
 *     H (T,N) activations     == sklearn's W (_W); the reference returns the view _W.T
 *     Y (T,F) converted       == H.T @ B                  04_align_n_nmf.py:391
 *
 * Unless a function name ends in _host, all data pointers are DEVICE pointers owned by
 * the caller, and the call is asynchronous on `stream` (a cudaStream_t passed as void*)
 * except where it must read a scalar back (evc_solve with tol > 0 synchronises the stream
 * every `check_every` iterations, like the reference's stop rule, sklearn _nmf.py:867-879).
 * A handle must not be used from two host threads at once; distinct handles are independent.
 */
#ifndef EVC_H_
#define EVC_H_

#ifdef __cplusplus
extern "C" {
#endif

#define EVC_VERSION 100 /* major*10000 + minor*100 + patch */

typedef struct evc_dict* evc_dict_t;
typedef struct evc_comm* evc_comm_t;

enum evc_status {
  EVC_OK = 0,
  EVC_ERR_INVALID_ARGUMENT = 1, /* bad shape / null pointer / bad enum               */
  EVC_ERR_CUDA = 2,             /* a CUDA runtime/driver call failed                  */
  EVC_ERR_UNSUPPORTED = 3,      /* mode not available for this shape / device         */
  EVC_ERR_VALUE = 4,            /* negative or all-zero dictionary (sklearn _nmf.py:61-76 raises ValueError) */
  EVC_ERR_COMM = 5              /* NCCL failure or libnccl not loadable               */
};

/* Arithmetic of the two contractions A*H and A^T*R (and of Y = B*H). */
enum evc_mode {
  EVC_MODE_FP32 = 0,   /* fp32 FFMA on CUDA cores: exact fp32, any shape (also F = 1, the f0 track) */
  EVC_MODE_3XTF32 = 1, /* fp32-accurate: every operand x = x1 + x2 (two bf16 planes, made once where the data is
                          produced); three tcgen05 kind::f16 MMAs per product (x2*y1 + x1*y2 + x1*y1), fp32
                          accumulate in TMEM.  (The enumerator keeps its round-1 name for ABI stability.)       */
  EVC_MODE_TF32 = 2,   /* tcgen05 kind::tf32, one MMA per product: fast mode                           */
  EVC_MODE_BF16 = 3    /* tcgen05 kind::f16 on bf16 copies of A, of the ratio and of H (shadow kept by the fused
                          update), fp32 accumulate and fp32 multiplicative update: fastest, ~1e-3 on H     */
};

enum evc_loss {
  EVC_LOSS_KL = 1,       /* beta = 1, sklearn _nmf.py:551-591 : H <- H * A^T(X / AH) / (A^T 1 + lambda)      */
  EVC_LOSS_FROBENIUS = 2 /* beta = 2, sklearn _nmf.py:535-549 : H <- H * A^T X / (A^T (A H) + lambda)        */
};

/* How H is initialised. */
enum evc_init {
  EVC_INIT_SKLEARN = 0, /* H0 = sqrt(mean(X)/N) everywhere (sklearn _nmf.py:1224-1226)          */
  EVC_INIT_GIVEN = 1    /* H holds the caller's initial activations (nmf_tool/nmf.py:28-31 style) */
};

typedef struct evc_solve_params {
  int loss;          /* enum evc_loss                                                                */
  int init;          /* enum evc_init                                                                */
  int max_iter;      /* reference: 150 (04_align_n_nmf.py:213)                                       */
  int check_every;   /* objective / stop rule period; reference: 10 (sklearn _nmf.py:867)            */
  float tol;         /* stop when (prev - err) / err_init < tol; 0 disables (sklearn _nmf.py:867-879) */
  float lambda;      /* constant L1 penalty added to the denominator (north star formula)            */
  float lambda_step; /* extra penalty added per iteration: den_k = A^T1 + lambda + k*lambda_step.
                        lambda = 0, lambda_step = l1_reg_W reproduces sklearn 1.9.0's in-place
                        accumulation; 0 for the north-star formula            */
  float epsilon;     /* clamp for A*H and for a zero denominator; <= 0 selects the reference's
                        value 1.1920929e-07 (sklearn _nmf.py:32)                                      */
} evc_solve_params;

typedef struct evc_solve_result {
  int n_iter;               /* iterations run (1-based count at break, else max_iter)   */
  int converged;            /* 1 when the stop rule fired                               */
  double objective;         /* sqrt(2 KL(X || H^T A)) or ||X - H^T A||_F of the result  */
  double objective_at_init; /* the same for H0                                          */
} evc_solve_result;

int evc_version(void);
/* Thread-local message for the last non-zero status returned on this thread. */
const char* evc_last_error_string(void);

/* Fill `p` with the reference's defaults: KL, sklearn init, max_iter 150, check 10, tol 1e-4, lambda 0. */
void evc_default_params(evc_solve_params* p);

/*
 * Make a device-resident dictionary from A (N,F) (row pitch ldA floats) and optionally the
 * paired target dictionary B (N,F) (row pitch ldB): re-pitches them for TMA, builds the
 * transposed / hi-lo / bf16 operand copies the mode needs, and computes A^T 1
 * (sklearn's cached H_sum, _nmf.py:588-590).  Replaces the per-call
 * `A_sp = np.asarray(A_sp)` staging of 04_align_n_nmf.py:230-246.  Synchronises `stream`
 * once to validate A >= 0 and not all-zero (sklearn _nmf.py:61-76 -> EVC_ERR_VALUE).
 */
int evc_dict_create(const float* A, int ldA, const float* B, int ldB, int F, int N, int mode,
                    void* stream, evc_dict_t* out);
int evc_dict_destroy(evc_dict_t d);
/* `mode` reports the mode that RUNS: a problem below one MMA operand tile (F < 64 or N < 128, e.g. the F = 1 f0 track
 * of 04_align_n_nmf.py:288) runs on the exact-fp32 CUDA-core kernels whatever tensor-core mode was asked for. */
int evc_dict_info(evc_dict_t d, int* F, int* N, int* mode, int* has_target);
/* Copy A^T 1 (N floats) to a device buffer. */
int evc_dict_colsum(evc_dict_t d, float* out, void* stream);

/*
 * Activation solve: replaces the sklearn call at 04_align_n_nmf.py:212-213.
 * X (T,F) pitch ldX; H (T,N) pitch ldH is written (and read first when init = GIVEN).
 */
int evc_solve(evc_dict_t d, const float* X, int ldX, int T, float* H, int ldH,
              const evc_solve_params* p, evc_solve_result* res, void* stream);

/*
 * Batched-utterance solve: X_stacked holds n_utt utterances stacked along T;
 * utterance u owns rows [t_offsets[u], t_offsets[u+1]) (host array of n_utt+1 ints).
 * Each utterance gets its own H0 (from its own mean) and, when per_utterance_stop != 0, its
 * own stop decision exactly as n_utt separate reference calls would; `res` has n_utt entries.
 */
int evc_solve_batched(evc_dict_t d, const float* X_stacked, int ldX, const int* t_offsets, int n_utt,
                      float* H, int ldH, const evc_solve_params* p, int per_utterance_stop,
                      evc_solve_result* res, void* stream);

/* Conversion product Y (T,F) = H (T,N) * B (N,F): 04_align_n_nmf.py:391. */
int evc_convert(evc_dict_t d, const float* H, int ldH, int T, float* Y, int ldY, void* stream);

/* Reconstruction WH (T,F) = H (T,N) * A (N,F): the first contraction on its own; the reference forms it for
 * the residual log(H^T A - X) at 04_align_n_nmf.py:292-294. */
int evc_reconstruct(evc_dict_t d, const float* H, int ldH, int T, float* WH, int ldWH, void* stream);

/* sqrt(2 KL) / Frobenius objective of a given H (sklearn _beta_divergence, _nmf.py:78-182). Synchronises. */
int evc_objective(evc_dict_t d, const float* X, int ldX, int T, const float* H, int ldH, int loss,
                  float epsilon, double* out, void* stream);

/*
 * Host-buffer convenience: X, H, Y are HOST pointers (H and/or Y may be NULL to skip the
 * copy back).  Does H2D, evc_solve, evc_convert (if Y != NULL and the dictionary has a
 * target) and D2H on `stream`, then synchronises.
 */
int evc_factorize_convert_host(evc_dict_t d, const float* X, int ldX, int T, float* H, int ldH,
                               float* Y, int ldY, const evc_solve_params* p, evc_solve_result* res,
                               void* stream);

/*
 * Exemplar (N) sharding across GPUs: each rank builds its dictionary from its own rows
 * A[n_begin:n_end], attaches a communicator, and evc_solve / evc_convert / evc_objective
 * then all-reduce the partial A*H (T,F) inside the loop.  The NCCL library
 * is the one already loaded in the process (libnccl.so.2), resolved with dlopen.
 */
int evc_comm_unique_id(char id_out[128]);
int evc_comm_create(const char id[128], int rank, int world, evc_comm_t* out);
int evc_comm_destroy(evc_comm_t c);
int evc_dict_attach_comm(evc_dict_t d, evc_comm_t c, int n_total);

/*
 * Dictionary construction on the device: aligned-frame gather with optional +-context frame
 * stacking.  Replaces the Python double loop that gathers frames by DTW index (04_align_n_nmf.py:113-124) and
 * the list.extend / np.asarray stacking (04_align_n_nmf.py:230-246); context > 0 builds the stacked exemplars of
 * BASELINE.json's F = 2565 configuration ((2*context+1) * F features per row).
 *   out[k, (d+context)*F + f] = frames[clamp(idx[k] + d, lo[k], hi[k]-1), f]      d = -context..context
 * frames (n_frames, F) pitch ld; idx / lo / hi: n_out device ints (frame index and the [lo, hi) range of the file
 * the frame belongs to, so stacking never crosses a file boundary); out (n_out, (2*context+1)*F) pitch ld_out.
 */
int evc_gather_stack(const float* frames, int ld, int n_frames, int F, const int* idx, const int* lo, const int* hi,
                     int n_out, int context, float* out, int ld_out, void* stream);

/*
 * The step right after the path.
 *
 * WORLD-branch residual compensation as epilogues of the two products:
 *   evc_residual:          R (T,F) = log(H A - X)                       04_align_n_nmf.py:292-294
 *   evc_convert_residual:  Y (T,F) = exp(log(H B) + log(R')), R' = R with NaN -> 0       04_align_n_nmf.py:363-373
 * (IEEE semantics of the reference's numpy expression: NaN where H A < X, Y = 0 where R' = 0, NaN where R' < 0).
 *
 * Griffin-Lim vocoder of the STFT branch (zz_audio_utilities.py:258-292, called with fft_size 400, hop 80,
 * 300 iterations at 04_align_n_nmf.py:187), in double precision like the reference: device pointers, `window` =
 * fft_size doubles (np.hanning(fft_size)), signals of T*hop + fft_size samples, x0 = the start signal (the
 * reference draws np.random.randn), mag (T, fft_size/2+1) float.  sq_diff (iterations doubles, may be NULL) receives
 * sum (x_new - x_prev)^2 per iteration (the RMSE the reference prints is sqrt(sq_diff / len)).  evc_stft / evc_istft
 * are zz_audio_utilities.py:181-196 / 199-218; spec is (T, fft_size/2+1, 2) doubles (re, im).
 */
int evc_residual(evc_dict_t d, const float* X, int ldX, int T, const float* H, int ldH, float* R, int ldR, void* stream);
int evc_convert_residual(evc_dict_t d, const float* H, int ldH, int T, const float* R, int ldR, float* Y, int ldY,
                         void* stream);
int evc_griffin_lim(const float* mag, int ldm, int T, int fft_size, int hop, int iterations, const double* window,
                    const double* x0, double* x_out, double* sq_diff, void* stream);
int evc_stft(const double* x, long long len, int fft_size, int hop, const double* window, double* spec, void* stream);
int evc_istft(const double* spec, int T, int fft_size, int hop, const double* window, double* x_out, void* stream);

/*
 * DTW alignment of n_files parallel utterance pairs (01_make_dict_parallel.py:215-249: the `dtw`
 * package with the squared-L2 local cost of :226, default step pattern).  Device pointers.  A (sum_r, dim) and
 * B (sum_c, dim) hold the per-file feature frames (float64, row-major) back to back; file f owns rows
 * [a_off[f], a_off[f+1]) / [b_off[f], b_off[f+1]) (n_files + 1 offsets each); max_frames >= every file's frame count
 * on the A side.  dirs = workspace of sum_f r_f * c_f bytes, file f at dir_off[f].  Outputs: the alignment path of
 * file f, REVERSED (from (r-1, c-1) back to (0, 0)), in path_a / path_b[path_off[f] ...] (room for r_f + c_f - 1
 * entries), its length in path_len[f], and dist[f] = accumulated cost / (r + c) as the package returns it.
 */
int evc_dtw(const double* A, const long long* a_off, const double* B, const long long* b_off, int n_files, int dim,
            int max_frames, unsigned char* dirs, const long long* dir_off, int* path_a, int* path_b,
            const long long* path_off, int* path_len, double* dist, void* stream);

/*
 * Optional: replace the NCCL all-reduce of the exemplar-sharded path by libevc_b200's own kernel over NVLink peer
 * memory (one node, 2..8 ranks, one process per GPU).  Every rank calls evc_p2p_alloc (allocates the exchange buffer
 * for up to max_frames frames and returns a 64-byte CUDA IPC handle), the handles of all ranks are exchanged by the
 * host (rank-major, world*64 bytes) and passed to evc_p2p_attach.  evc_dict_attach_comm must have been called first.
 */
int evc_p2p_alloc(evc_dict_t d, int max_frames, char handle_out[64]);
int evc_p2p_attach(evc_dict_t d, const char* handles, int rank, int world);
/* Unmap the peers' buffers again: the handle goes back to the NCCL all-reduce.  Used when not EVERY rank could attach
 * (all ranks must use the same exchange).  A peer that does not arrive within EVC_P2P_TIMEOUT_S seconds (default 60)
 * makes the next synchronising call return EVC_ERR_COMM instead of hanging or trapping. */
int evc_p2p_detach(evc_dict_t d);

/* Diagnostics: host milliseconds the last evc_solve* on this thread spent ENQUEUEING its iteration loop (if this
 * approaches the device time of the loop, the GPU is waiting for the host). */
double evc_last_enqueue_ms(void);

/* Diagnostics: number of kernels this library has launched in this process. */
long long evc_kernel_launch_count(void);

/* Diagnostics: MMAs one logical product costs in `mode`: EVC_MODE_3XTF32 -> 3 (bf16 MMAs: x2*y1 + x1*y2 + x1*y1),
 * EVC_MODE_TF32 -> 1 (a tf32 MMA), EVC_MODE_BF16 -> 1 (a bf16 MMA), EVC_MODE_FP32 -> 0 (no tensor cores).  bench.py
 * multiplies the algorithmic TFLOP/s by this to get the executed figure ncu's tensor pipe sees (against the dense
 * BF16 rate for the first and third mode, the dense TF32 rate for the second). */
int evc_mma_passes_per_product(int mode);

/*
 * Diagnostics: per-kernel-class device timing.  After evc_profile_enable(d, 1) every launch of the handle is
 * bracketed by CUDA events on the caller's stream; evc_profile_read synchronises, sums the elapsed times per
 * class, returns them and clears the log.  Classes: 0 = contraction 1 (A*H, B*H: tensor/FFMA GEMM),
 * 1 = split-K reduction + ratio (memory-bound), 2 = contraction 2 with the fused multiplicative update,
 * 3 = objective / init helpers, 4 = the per-iteration exchange of the partial A*H when the exemplar dimension
 * is sharded.  ms[EVC_PROFILE_CLASSES] and launches[EVC_PROFILE_CLASSES] are host arrays.
 */
#define EVC_PROFILE_CLASSES 5
int evc_profile_enable(evc_dict_t d, int on);
int evc_profile_read(evc_dict_t d, double* ms, int* launches);

#ifdef __cplusplus
}
#endif
#endif /* EVC_H_ */
