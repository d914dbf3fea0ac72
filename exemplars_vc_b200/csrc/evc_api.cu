// libevc_b200: C ABI (include/evc.h) over the CUDA kernels.  One translation unit.
//
// Host-side control flow restates sklearn/decomposition/_nmf.py (1.9.0), the dependency that
// holds the reference's arithmetic (04_align_n_nmf.py:212-213):
//   _check_w_h (update_H=False, solver 'mu')      :1205-1226   -> init_activations()
//   _fit_multiplicative_update                    :726-886     -> solve_impl() loop + stop rule
//   _multiplicative_update_w                      :521-626     -> step_kl()/step_fro()
//   _beta_divergence                              :78-182      -> objective()
#include "evc_common.cuh"
#include "simt_kernels.cuh"
#include "tc_kernels.cuh"
#include "nccl_shim.cuh"
#include "p2p_allreduce.cuh"
#include "audio_kernels.cuh"
#include "dtw_kernels.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <new>

using namespace evc;

struct evc_comm {
  nccl::Comm comm;
  int rank = 0, world = 1;
};

struct evc_dict {
  int F = 0, N = 0, mode = 0, device = 0;
  int mode_requested = 0;  // what the caller asked for; `mode` is what runs (problems below one MMA tile -> FFMA)
  int n_total = 0;  // exemplars across all shards (== N when not sharded): H0 uses it (sklearn n_components)
  bool has_target = false;
  // fp32 mode operands
  int ldA = 0;               // pitch of A / B copies (floats)
  float* A = nullptr;        // (N, ldA) zero padded
  float* B = nullptr;        // (N, ldA)
  float* colsum = nullptr;   // A^T 1  (N)
  tc::DictOperands tc_ops;   // tensor-core operand copies + tensor maps (modes 1..3)
  evc_comm* comm = nullptr;  // exemplar sharding: all-reduce of partial A*H
  p2p::State p2p;            // ... by our own kernel over NVLink peer memory when attached (else NCCL)
  // per-solve workspace, grow-only
  DevBuf WH, R, rowd, w0, active, num0, tcws;
  DevBuf hostX, hostH, hostY;   // device staging of evc_factorize_convert_host (grow-only: no cudaMalloc per call)
  double* host_rows = nullptr;  // pinned mirror of rowd
  size_t host_rows_cap = 0;
  float* host_w0 = nullptr;     // pinned
  unsigned char* host_active = nullptr;
  size_t host_T_cap = 0;
  int ldWH = 0, ldR = 0;
  Profiler prof;
};

namespace {

thread_local double g_last_enqueue_ms = 0.0;
constexpr int kMinTensorF = 64, kMinTensorN = 128;

// A*H of the current activations.  With the peer-memory all-reduce attached it lives in the IPC exchange buffer
// (the reduced result lands there directly); otherwise in the handle's own workspace.
inline float* wh_buf(evc_dict* d) { return d->p2p.attached ? d->p2p.recv() : d->WH.as<float>(); }

int reserve_workspace(evc_dict* d, int T, int ldH, bool need_num0) {
  d->ldWH = round_up(d->F, 4);
  d->ldR = tc::k_pitch(d->F);
  if (d->p2p.attached && T > d->p2p.t_max)
    return fail(EVC_ERR_INVALID_ARGUMENT, "T = %d exceeds the %d frames the peer-memory exchange buffer was sized for", T,
                d->p2p.t_max);
  EVC_TRY(d->WH.reserve((size_t)T * d->ldWH * sizeof(float)));
  EVC_TRY(d->R.reserve((size_t)T * d->ldR * sizeof(float)));
  EVC_TRY(d->rowd.reserve((size_t)T * sizeof(double)));
  EVC_TRY(d->w0.reserve((size_t)T * sizeof(float)));
  EVC_TRY(d->active.reserve((size_t)T));
  if (need_num0) EVC_TRY(d->num0.reserve((size_t)T * ldH * sizeof(float)));
  if ((size_t)T > d->host_T_cap) {
    if (d->host_rows) cudaFreeHost(d->host_rows);
    if (d->host_w0) cudaFreeHost(d->host_w0);
    if (d->host_active) cudaFreeHost(d->host_active);
    d->host_rows = nullptr; d->host_w0 = nullptr; d->host_active = nullptr; d->host_T_cap = 0;
    EVC_CUDA(cudaMallocHost(&d->host_rows, (size_t)T * sizeof(double)));
    EVC_CUDA(cudaMallocHost(&d->host_w0, (size_t)T * sizeof(float)));
    EVC_CUDA(cudaMallocHost(&d->host_active, (size_t)T));
    d->host_T_cap = T;
  }
  return EVC_OK;
}

// ---- the three contractions, dispatched on mode -------------------------------------------------

// WH (T, ldWH) = H (T,N) * A (N,F)         [first contraction; sklearn :554]
int contract_wh(evc_dict* d, const float* H, int ldH, int T, float* WH, int ldWH, bool target, cudaStream_t s,
                const tc::RatioArgs* ra = nullptr) {
  const bool sharded = d->comm && d->comm->world > 1;
  // peer-memory all-reduce: the local partial goes to the send region, the sum arrives in the recv region (== WH)
  const bool use_p2p = sharded && d->p2p.attached && WH == d->p2p.recv() && ldWH == d->ldWH;
  float* local_out = use_p2p ? d->p2p.send() : WH;
  if (d->mode == EVC_MODE_FP32) {
    ProfScope ps(0, s);
    simt::EpiArgs e{};
    e.C = local_out; e.ldc = ldWH;
    EVC_TRY((simt::launch_gemm<simt::EPI_STORE, false>(T, d->F, d->N, H, ldH, target ? d->B : d->A, d->ldA, e, s)));
  } else {
    EVC_TRY(tc::contract_wh(d->tc_ops, d->mode, H, ldH, T, local_out, ldWH, target, &d->tcws, s, ra));
  }
  if (sharded) {
    ProfScope ps(4, s);  // exemplar sharding: the per-iteration exchange of the partial A*H
    if (use_p2p) EVC_TRY(p2p::all_reduce(&d->p2p, (size_t)T * ldWH, s));
    else EVC_TRY(nccl::all_reduce_sum(d->comm->comm, WH, (size_t)T * ldWH, s));
  }
  return EVC_OK;
}

// KL: H *= (R A^T) / (A^T1 + lam)   [second contraction + multiplicative update; sklearn :585-624]
int update_kl(evc_dict* d, const float* X, int ldX, int T, float* H, int ldH, float lam, float eps,
              const unsigned char* row_active, cudaStream_t s, bool ratio_done = false) {
  float* R = d->R.as<float>();
  if (d->mode == EVC_MODE_FP32) {
    {
      ProfScope ps(1, s);
      dim3 g(T, ceil_div(d->ldR, 256));
      simt::ratio_kernel<<<g, 256, 0, s>>>(X, ldX, wh_buf(d), d->ldWH, R, d->ldR, T, d->F, eps);
      EVC_LAUNCH_CHECK();
    }
    ProfScope ps(2, s);
    simt::EpiArgs e{};
    e.C = H; e.ldc = ldH; e.colsum = d->colsum; e.lam = lam; e.eps = eps; e.row_active = row_active;
    EVC_TRY((simt::launch_gemm<simt::EPI_MU_KL, true>(T, d->N, d->F, R, d->ldR, d->A, d->ldA, e, s)));
    return EVC_OK;
  }
  return tc::update_kl(d->tc_ops, d->mode, X, ldX, T, wh_buf(d), d->ldWH, R, d->ldR, H, ldH, d->colsum,
                       lam, eps, row_active, &d->tcws, s, ratio_done);
}

// Frobenius: H *= NUM0 / (WH A^T + lam)   [sklearn :535-549, with A^T(A H) instead of the N x N Gram]
int update_fro(evc_dict* d, int T, float* H, int ldH, const float* num0, float lam, float eps,
               const unsigned char* row_active, cudaStream_t s) {
  if (d->mode == EVC_MODE_FP32) {
    ProfScope ps(2, s);
    simt::EpiArgs e{};
    e.C = H; e.ldc = ldH; e.X = num0; e.ldx = ldH; e.lam = lam; e.eps = eps; e.row_active = row_active;
    EVC_TRY((simt::launch_gemm<simt::EPI_MU_FRO, true>(T, d->N, d->F, wh_buf(d), d->ldWH, d->A, d->ldA, e, s)));
    return EVC_OK;
  }
  return tc::update_fro(d->tc_ops, d->mode, T, wh_buf(d), d->ldWH, d->R.as<float>(), d->ldR, H, ldH, num0,
                        lam, eps, row_active, &d->tcws, s);
}

// NUM0 (T, ldH) = X A^T  (Frobenius numerator, computed once: sklearn :537-543)
int frob_numerator(evc_dict* d, const float* X, int ldX, int T, float* num0, int ldH, cudaStream_t s) {
  if (d->mode == EVC_MODE_FP32) {
    simt::EpiArgs e{};
    e.C = num0; e.ldc = ldH;
    EVC_TRY((simt::launch_gemm<simt::EPI_STORE, true>(T, d->N, d->F, X, ldX, d->A, d->ldA, e, s)));
    return EVC_OK;
  }
  return tc::frob_numerator(d->tc_ops, d->mode, X, ldX, T, d->R.as<float>(), d->ldR, num0, ldH, &d->tcws, s);
}

// Per-segment objective: err[u] = sqrt(2 * sum rows) (KL) or sqrt(sum rows) (Frobenius). Synchronises.
int objective_segments(evc_dict* d, const float* X, int ldX, int T, const float* H, int ldH, int loss, float eps,
                       const std::vector<int>& seg, std::vector<double>& err, cudaStream_t s) {
  EVC_TRY(contract_wh(d, H, ldH, T, wh_buf(d), d->ldWH, false, s));
  {
    ProfScope ps(3, s);
    simt::objective_rows_kernel<<<ceil_div(T, simt::kObjRowsPerBlock), simt::kObjWarpsPerRow * simt::kObjRowsPerBlock * 32, 0, s>>>(
        X, ldX, wh_buf(d), d->ldWH, T, d->F, eps, loss, d->rowd.as<double>());
    EVC_LAUNCH_CHECK();
  }
  EVC_CUDA(cudaMemcpyAsync(d->host_rows, d->rowd.p, (size_t)T * sizeof(double), cudaMemcpyDeviceToHost, s));
  EVC_CUDA(cudaStreamSynchronize(s));
  EVC_TRY(p2p::poll_error(&d->p2p));  // a peer that never arrived (exemplar sharding) is an error, not a hang
  const int nseg = (int)seg.size() - 1;
  err.assign(nseg, 0.0);
  for (int u = 0; u < nseg; ++u) {
    double acc = 0.0;
    for (int t = seg[u]; t < seg[u + 1]; ++t) acc += d->host_rows[t];
    acc = acc > 0.0 ? acc : 0.0;  // sklearn :177 max(res, 0)
    err[u] = (loss == EVC_LOSS_KL) ? std::sqrt(2.0 * acc) : std::sqrt(acc);
  }
  return EVC_OK;
}

int solve_impl(evc_dict* d, const float* X, int ldX, const int* t_offsets, int n_utt, float* H, int ldH,
               const evc_solve_params* p, int per_utterance_stop, evc_solve_result* res, cudaStream_t s) {
  if (!d || !X || !H || !p || !t_offsets || n_utt < 1)
    return fail(EVC_ERR_INVALID_ARGUMENT, "evc_solve: null argument or n_utt < 1");
  const int T = t_offsets[n_utt];
  if (t_offsets[0] != 0) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_solve: t_offsets[0] must be 0");
  for (int u = 0; u < n_utt; ++u)
    if (t_offsets[u + 1] < t_offsets[u]) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_solve: t_offsets not monotone");
  if (ldX < d->F || ldH < d->N) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_solve: ldX < F or ldH < N");
  if (p->loss != EVC_LOSS_KL && p->loss != EVC_LOSS_FROBENIUS)
    return fail(EVC_ERR_INVALID_ARGUMENT, "evc_solve: loss must be EVC_LOSS_KL or EVC_LOSS_FROBENIUS");
  if (p->max_iter < 0 || p->check_every < 1 || p->tol < 0.f || p->lambda < 0.f || p->lambda_step < 0.f)
    return fail(EVC_ERR_INVALID_ARGUMENT, "evc_solve: max_iter/check_every/tol/lambda out of range");
  const float eps = p->epsilon > 0.f ? p->epsilon : kEpsilon;
  const int loss = p->loss;
  g_prof = &d->prof;
  if (T == 0) {
    for (int u = 0; u < n_utt && res; ++u) res[u] = evc_solve_result{0, 0, 0.0, 0.0};
    return EVC_OK;
  }
  EVC_TRY(reserve_workspace(d, T, ldH, loss == EVC_LOSS_FROBENIUS));
  EVC_TRY(tc::check_alignment(d->mode, H, ldH));

  // Stop-rule segments: per utterance, or the whole stack as one matrix (a single reference call).
  std::vector<int> seg;
  if (per_utterance_stop) seg.assign(t_offsets, t_offsets + n_utt + 1);
  else seg = {0, T};
  const int nseg = (int)seg.size() - 1;

  // H0 = sqrt(mean(X_u) / n_components)   sklearn :1225-1226
  if (p->init == EVC_INIT_SKLEARN) {
    simt::row_sum_kernel<<<ceil_div(T, 8), 256, 0, s>>>(X, ldX, T, d->F, d->rowd.as<double>());
    EVC_LAUNCH_CHECK();
    EVC_CUDA(cudaMemcpyAsync(d->host_rows, d->rowd.p, (size_t)T * sizeof(double), cudaMemcpyDeviceToHost, s));
    EVC_CUDA(cudaStreamSynchronize(s));
    for (int u = 0; u < nseg; ++u) {
      double acc = 0.0;
      for (int t = seg[u]; t < seg[u + 1]; ++t) acc += d->host_rows[t];
      const int rows = seg[u + 1] - seg[u];
      const double mean = rows > 0 ? acc / ((double)rows * d->F) : 0.0;
      const float w0 = (float)std::sqrt(mean / (double)d->n_total);
      for (int t = seg[u]; t < seg[u + 1]; ++t) d->host_w0[t] = w0;
    }
    EVC_CUDA(cudaMemcpyAsync(d->w0.p, d->host_w0, (size_t)T * sizeof(float), cudaMemcpyHostToDevice, s));
    simt::fill_rows_kernel<<<T, 256, 0, s>>>(H, ldH, T, d->N, d->w0.as<float>());
    EVC_LAUNCH_CHECK();
  } else if (p->init != EVC_INIT_GIVEN) {
    return fail(EVC_ERR_INVALID_ARGUMENT, "evc_solve: init must be EVC_INIT_SKLEARN or EVC_INIT_GIVEN");
  }
  EVC_TRY(tc::after_h_written(d->tc_ops, d->mode, H, ldH, T, &d->tcws, s));

  std::vector<double> err0, err, prev;
  EVC_TRY(objective_segments(d, X, ldX, T, H, ldH, loss, eps, seg, err0, s));  // sklearn :822
  prev = err0;
  std::vector<int> n_iter(nseg, p->max_iter), conv(nseg, 0);
  std::vector<char> active(nseg, 1);
  int n_active = nseg;
  bool any_frozen = false;
  // an utterance without frames has nothing to solve: converged before the first iteration (its err0 = 0 would make
  // the stop test NaN and keep the loop alive until max_iter)
  for (int u = 0; u < nseg; ++u)
    if (seg[u + 1] == seg[u]) { active[u] = 0; conv[u] = 1; n_iter[u] = 0; --n_active; }

  float* num0 = d->num0.as<float>();
  if (loss == EVC_LOSS_FROBENIUS) EVC_TRY(frob_numerator(d, X, ldX, T, num0, ldH, s));

  // tensor-core modes on one GPU: the split-K reduction of contraction 1 emits the ratio in the same pass
  const bool fuse_ratio = d->mode != EVC_MODE_FP32 && loss == EVC_LOSS_KL && !(d->comm && d->comm->world > 1);
  const tc::RatioArgs ra{X, ldX, d->R.as<float>(), d->ldR, eps};
  bool wh_fresh = true;  // WH = H A of the current H is in the workspace (the objective at init just made it)
  const auto enq0 = std::chrono::steady_clock::now();
  double sync_ms = 0.0;
  for (int k = 1; k <= p->max_iter && n_active > 0; ++k) {
    const float lam = p->lambda + (float)k * p->lambda_step;
    const unsigned char* mask = any_frozen ? d->active.as<unsigned char>() : nullptr;
    bool ratio_done = false;
    if (!wh_fresh) {
      EVC_TRY(contract_wh(d, H, ldH, T, wh_buf(d), d->ldWH, false, s, fuse_ratio ? &ra : nullptr));
      ratio_done = fuse_ratio;
    }
    wh_fresh = false;
    if (loss == EVC_LOSS_KL) EVC_TRY(update_kl(d, X, ldX, T, H, ldH, lam, eps, mask, s, ratio_done));
    else EVC_TRY(update_fro(d, T, H, ldH, num0, lam, eps, mask, s));

    if (p->tol > 0.f && k % p->check_every == 0) {  // sklearn :867-879
      const auto c0 = std::chrono::steady_clock::now();
      EVC_TRY(objective_segments(d, X, ldX, T, H, ldH, loss, eps, seg, err, s));
      sync_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - c0).count();
      wh_fresh = true;  // the next iteration reuses this A*H (sklearn recomputes it, _nmf.py:554 after :868)
      bool changed = false;
      for (int u = 0; u < nseg; ++u) {
        if (!active[u]) continue;
        if ((prev[u] - err[u]) / err0[u] < (double)p->tol) {
          active[u] = 0; conv[u] = 1; n_iter[u] = k; --n_active; changed = true;
        }
        prev[u] = err[u];
      }
      if (changed && n_active > 0) {
        for (int u = 0; u < nseg; ++u)
          memset(d->host_active + seg[u], active[u] ? 1 : 0, (size_t)(seg[u + 1] - seg[u]));
        EVC_CUDA(cudaMemcpyAsync(d->active.p, d->host_active, (size_t)T, cudaMemcpyHostToDevice, s));
        any_frozen = true;
      }
    }
  }

  g_last_enqueue_ms =
      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - enq0).count() - sync_ms;

  if (res) {
    EVC_TRY(objective_segments(d, X, ldX, T, H, ldH, loss, eps, seg, err, s));
    for (int u = 0; u < n_utt; ++u) {
      const int q = per_utterance_stop ? u : 0;
      res[u].n_iter = n_iter[q];
      res[u].converged = conv[q];
      res[u].objective = err[q];
      res[u].objective_at_init = err0[q];
    }
  }
  return EVC_OK;
}

}  // namespace

// ---- C ABI ---------------------------------------------------------------------------------------

extern "C" {

int evc_version(void) { return EVC_VERSION; }
const char* evc_last_error_string(void) { return g_err; }
long long evc_kernel_launch_count(void) { return g_launches.load(); }
double evc_last_enqueue_ms(void) { return g_last_enqueue_ms; }
int evc_mma_passes_per_product(int mode) {
  // MMAs issued per product: 3 bf16 (kind::f16) MMAs in the fp32-accurate split mode, 1 in the fast modes
  if (mode == EVC_MODE_3XTF32) return 3;
  return (mode == EVC_MODE_TF32 || mode == EVC_MODE_BF16) ? 1 : 0;
}

void evc_default_params(evc_solve_params* p) {
  if (!p) return;
  p->loss = EVC_LOSS_KL;
  p->init = EVC_INIT_SKLEARN;
  p->max_iter = 150;
  p->check_every = 10;
  p->tol = 1e-4f;
  p->lambda = 0.f;
  p->lambda_step = 0.f;
  p->epsilon = 0.f;
}

int evc_dict_destroy(evc_dict_t d) {
  if (!d) return EVC_OK;
  cudaFree(d->A); cudaFree(d->B); cudaFree(d->colsum);
  d->tc_ops.release();
  d->WH.release(); d->R.release(); d->rowd.release(); d->w0.release(); d->active.release();
  d->num0.release(); d->tcws.release(); d->prof.release(); p2p::release(&d->p2p);
  d->hostX.release(); d->hostH.release(); d->hostY.release();
  if (g_prof == &d->prof) g_prof = nullptr;
  if (d->host_rows) cudaFreeHost(d->host_rows);
  if (d->host_w0) cudaFreeHost(d->host_w0);
  if (d->host_active) cudaFreeHost(d->host_active);
  delete d;
  return EVC_OK;
}

int evc_dict_create(const float* A, int ldA, const float* B, int ldB, int F, int N, int mode, void* stream,
                    evc_dict_t* out) {
  if (!out) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_dict_create: out is null");
  *out = nullptr;
  if (!A || F < 1 || N < 1 || ldA < F || (B && ldB < F))
    return fail(EVC_ERR_INVALID_ARGUMENT, "evc_dict_create: bad shape (F=%d N=%d ldA=%d ldB=%d)", F, N, ldA, ldB);
  if (mode < EVC_MODE_FP32 || mode > EVC_MODE_BF16)
    return fail(EVC_ERR_INVALID_ARGUMENT, "evc_dict_create: unknown mode %d", mode);
  cudaStream_t s = (cudaStream_t)stream;
  evc_dict* d = new (std::nothrow) evc_dict();
  if (!d) return fail(EVC_ERR_CUDA, "evc_dict_create: out of host memory");
  d->mode_requested = mode;
  // A problem that does not fill one 128 x 64 operand tile (the F = 1 f0 track of 04_align_n_nmf.py:288, toy cases)
  // gains nothing from the tensor cores and leaves the split products too few terms to average their rounding:
  // it runs on the exact-fp32 CUDA-core kernels whatever tensor-core mode was asked for.
  if (mode != EVC_MODE_FP32 && (F < kMinTensorF || N < kMinTensorN)) mode = EVC_MODE_FP32;
  d->F = F; d->N = N; d->n_total = N; d->mode = mode; d->has_target = (B != nullptr);
  int st = [&]() -> int {
    EVC_CUDA(cudaGetDevice(&d->device));
    if (mode != EVC_MODE_FP32) EVC_TRY(tc::check_device(d->device));
    d->ldA = tc::k_pitch(F);
    const size_t bytes = (size_t)N * d->ldA * sizeof(float);
    EVC_CUDA(cudaMalloc(&d->A, bytes));
    EVC_CUDA(cudaMalloc(&d->colsum, (size_t)N * sizeof(float)));
    dim3 g(N, ceil_div(d->ldA, 256));
    simt::repitch_kernel<<<g, 256, 0, s>>>(A, ldA, d->A, d->ldA, N, F);
    EVC_LAUNCH_CHECK();
    if (B) {
      EVC_CUDA(cudaMalloc(&d->B, bytes));
      simt::repitch_kernel<<<g, 256, 0, s>>>(B, ldB, d->B, d->ldA, N, F);
      EVC_LAUNCH_CHECK();
    }
    int* flags = nullptr;
    EVC_CUDA(cudaMalloc(&flags, 2 * sizeof(int)));
    EVC_CUDA(cudaMemsetAsync(flags, 0, 2 * sizeof(int), s));
    simt::colsum_kernel<<<ceil_div(N, 8), 256, 0, s>>>(d->A, d->ldA, N, F, d->colsum, flags);
    EVC_LAUNCH_CHECK();
    int hflags[2] = {0, 0};
    EVC_CUDA(cudaMemcpyAsync(hflags, flags, sizeof(hflags), cudaMemcpyDeviceToHost, s));
    EVC_CUDA(cudaStreamSynchronize(s));
    cudaFree(flags);
    // sklearn _check_init / check_non_negative (_nmf.py:61-76): ValueError on negative or all-zero dictionary
    if (hflags[0]) return fail(EVC_ERR_VALUE, "Negative values in data passed to NMF (input H)");
    if (!hflags[1]) return fail(EVC_ERR_VALUE, "Array passed to NMF (input H) is full of zeros.");
    if (mode != EVC_MODE_FP32) EVC_TRY(tc::build_operands(&d->tc_ops, mode, d->A, d->B, d->ldA, F, N, s));
    return EVC_OK;
  }();
  if (st != EVC_OK) {
    char keep[sizeof(g_err)];
    memcpy(keep, g_err, sizeof(keep));
    evc_dict_destroy(d);
    memcpy(g_err, keep, sizeof(keep));
    return st;
  }
  *out = d;
  return EVC_OK;
}

int evc_dict_info(evc_dict_t d, int* F, int* N, int* mode, int* has_target) {
  if (!d) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_dict_info: null handle");
  if (F) *F = d->F;
  if (N) *N = d->N;
  if (mode) *mode = d->mode;  // the mode that runs
  if (has_target) *has_target = d->has_target ? 1 : 0;
  return EVC_OK;
}

int evc_dict_colsum(evc_dict_t d, float* out, void* stream) {
  if (!d || !out) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_dict_colsum: null argument");
  EVC_CUDA(cudaMemcpyAsync(out, d->colsum, (size_t)d->N * sizeof(float), cudaMemcpyDeviceToDevice,
                           (cudaStream_t)stream));
  return EVC_OK;
}

int evc_solve(evc_dict_t d, const float* X, int ldX, int T, float* H, int ldH, const evc_solve_params* p,
              evc_solve_result* res, void* stream) {
  if (T < 0) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_solve: T < 0");
  const int off[2] = {0, T};
  return solve_impl(d, X, ldX, off, 1, H, ldH, p, 0, res, (cudaStream_t)stream);
}

int evc_solve_batched(evc_dict_t d, const float* X, int ldX, const int* t_offsets, int n_utt, float* H, int ldH,
                      const evc_solve_params* p, int per_utterance_stop, evc_solve_result* res, void* stream) {
  return solve_impl(d, X, ldX, t_offsets, n_utt, H, ldH, p, per_utterance_stop, res, (cudaStream_t)stream);
}

static int product_impl(evc_dict_t d, const float* H, int ldH, int T, float* Y, int ldY, bool target, void* stream,
                        const char* who) {
  if (!d || !H || !Y || T < 0) return fail(EVC_ERR_INVALID_ARGUMENT, "%s: null argument", who);
  if (target && !d->has_target)
    return fail(EVC_ERR_INVALID_ARGUMENT, "%s: dictionary was created without a target B", who);
  if (ldH < d->N || ldY < d->F) return fail(EVC_ERR_INVALID_ARGUMENT, "%s: ldH < N or ldY < F", who);
  if (T == 0) return EVC_OK;
  cudaStream_t s = (cudaStream_t)stream;
  g_prof = &d->prof;
  EVC_TRY(tc::check_alignment(d->mode, H, ldH));
  EVC_TRY(reserve_workspace(d, T, ldH, false));
  EVC_TRY(tc::after_h_written(d->tc_ops, d->mode, H, ldH, T, &d->tcws, s));
  EVC_TRY(contract_wh(d, H, ldH, T, wh_buf(d), d->ldWH, target, s));
  EVC_CUDA(cudaMemcpy2DAsync(Y, (size_t)ldY * sizeof(float), wh_buf(d), (size_t)d->ldWH * sizeof(float),
                             (size_t)d->F * sizeof(float), T, cudaMemcpyDeviceToDevice, s));
  return EVC_OK;
}

int evc_convert(evc_dict_t d, const float* H, int ldH, int T, float* Y, int ldY, void* stream) {
  return product_impl(d, H, ldH, T, Y, ldY, true, stream, "evc_convert");
}

int evc_reconstruct(evc_dict_t d, const float* H, int ldH, int T, float* WH, int ldWH, void* stream) {
  return product_impl(d, H, ldH, T, WH, ldWH, false, stream, "evc_reconstruct");
}

// ---- the step right after the path: residual compensation epilogues and the Griffin-Lim vocoder ----

int evc_residual(evc_dict_t d, const float* X, int ldX, int T, const float* H, int ldH, float* R, int ldR, void* stream) {
  if (!d || !X || !H || !R || T < 0) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_residual: null argument");
  if (ldX < d->F || ldH < d->N || ldR < d->F) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_residual: pitch smaller than the row length");
  if (T == 0) return EVC_OK;
  cudaStream_t s = (cudaStream_t)stream;
  g_prof = &d->prof;
  EVC_TRY(tc::check_alignment(d->mode, H, ldH));
  EVC_TRY(reserve_workspace(d, T, ldH, false));
  EVC_TRY(tc::after_h_written(d->tc_ops, d->mode, H, ldH, T, &d->tcws, s));
  EVC_TRY(contract_wh(d, H, ldH, T, wh_buf(d), d->ldWH, false, s));
  dim3 g(T, ceil_div(d->F, 128));
  audio::residual_kernel<<<g, 128, 0, s>>>(wh_buf(d), d->ldWH, X, ldX, R, ldR, T, d->F);
  EVC_LAUNCH_CHECK();
  return EVC_OK;
}

int evc_convert_residual(evc_dict_t d, const float* H, int ldH, int T, const float* R, int ldR, float* Y, int ldY,
                         void* stream) {
  if (!d || !H || !R || !Y || T < 0) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_convert_residual: null argument");
  if (!d->has_target) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_convert_residual: dictionary was created without a target B");
  if (ldH < d->N || ldR < d->F || ldY < d->F)
    return fail(EVC_ERR_INVALID_ARGUMENT, "evc_convert_residual: pitch smaller than the row length");
  if (T == 0) return EVC_OK;
  cudaStream_t s = (cudaStream_t)stream;
  g_prof = &d->prof;
  EVC_TRY(tc::check_alignment(d->mode, H, ldH));
  EVC_TRY(reserve_workspace(d, T, ldH, false));
  EVC_TRY(tc::after_h_written(d->tc_ops, d->mode, H, ldH, T, &d->tcws, s));
  EVC_TRY(contract_wh(d, H, ldH, T, wh_buf(d), d->ldWH, true, s));
  dim3 g(T, ceil_div(d->F, 128));
  audio::apply_residual_kernel<<<g, 128, 0, s>>>(wh_buf(d), d->ldWH, R, ldR, Y, ldY, T, d->F);
  EVC_LAUNCH_CHECK();
  return EVC_OK;
}

int evc_stft(const double* x, long long len, int fft_size, int hop, const double* window, double* spec, void* stream) {
  if (!x || !window || !spec || len <= fft_size) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_stft: bad argument");
  const int T = (int)((len - fft_size + hop - 1) / hop);  // range(0, len - fft_size, hop)
  EVC_TRY(audio::gl_check(T, fft_size, hop));
  return audio::gl_launch_frames(audio::GL_STFT, x, window, nullptr, 0, spec, T, fft_size, hop, nullptr, (cudaStream_t)stream);
}

int evc_istft(const double* spec, int T, int fft_size, int hop, const double* window, double* x_out, void* stream) {
  if (!spec || !window || !x_out) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_istft: null argument");
  EVC_TRY(audio::gl_check(T, fft_size, hop));
  cudaStream_t s = (cudaStream_t)stream;
  const long long len = (long long)T * hop + fft_size;
  double* frames = nullptr;
  EVC_CUDA(cudaMallocAsync(&frames, (size_t)T * fft_size * sizeof(double), s));
  int st = audio::gl_launch_frames(audio::GL_ISTFT, nullptr, window, nullptr, 0, const_cast<double*>(spec), T, fft_size, hop,
                                   frames, s);
  if (st == EVC_OK) {
    audio::gl_overlap_add_kernel<<<(unsigned)((len + 255) / 256), 256, 0, s>>>(frames, T, fft_size, hop, len, x_out, nullptr, nullptr);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (cudaGetLastError() != cudaSuccess) st = fail(EVC_ERR_CUDA, "evc_istft: overlap-add launch failed");
  }
  cudaFreeAsync(frames, s);
  return st;
}

int evc_griffin_lim(const float* mag, int ldm, int T, int fft_size, int hop, int iterations, const double* window,
                    const double* x0, double* x_out, double* sq_diff, void* stream) {
  if (!mag || !window || !x0 || !x_out || iterations < 0) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_griffin_lim: bad argument");
  EVC_TRY(audio::gl_check(T, fft_size, hop));
  if (ldm < fft_size / 2 + 1) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_griffin_lim: ldm < fft_size/2 + 1");
  cudaStream_t s = (cudaStream_t)stream;
  const long long len = (long long)T * hop + fft_size;
  double *frames = nullptr, *xa = nullptr, *xb = nullptr;
  EVC_CUDA(cudaMallocAsync(&frames, (size_t)T * fft_size * sizeof(double), s));
  EVC_CUDA(cudaMallocAsync(&xa, (size_t)len * sizeof(double), s));
  EVC_CUDA(cudaMallocAsync(&xb, (size_t)len * sizeof(double), s));
  int st = [&]() -> int {
    EVC_CUDA(cudaMemcpyAsync(xa, x0, (size_t)len * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (sq_diff && iterations > 0) EVC_CUDA(cudaMemsetAsync(sq_diff, 0, (size_t)iterations * sizeof(double), s));
    double *cur = xa, *nxt = xb;
    for (int it = 0; it < iterations; ++it) {
      EVC_TRY(audio::gl_launch_frames(audio::GL_ITERATE, cur, window, mag, ldm, nullptr, T, fft_size, hop, frames, s));
      audio::gl_overlap_add_kernel<<<(unsigned)((len + 255) / 256), 256, 0, s>>>(frames, T, fft_size, hop, len, nxt, cur,
                                                                                sq_diff ? sq_diff + it : nullptr);
      EVC_LAUNCH_CHECK();
      std::swap(cur, nxt);
    }
    EVC_CUDA(cudaMemcpyAsync(x_out, cur, (size_t)len * sizeof(double), cudaMemcpyDeviceToDevice, s));
    return EVC_OK;
  }();
  cudaFreeAsync(frames, s); cudaFreeAsync(xa, s); cudaFreeAsync(xb, s);
  return st;
}

int evc_dtw(const double* A, const long long* a_off, const double* B, const long long* b_off, int n_files, int dim,
            int max_frames, unsigned char* dirs, const long long* dir_off, int* path_a, int* path_b,
            const long long* path_off, int* path_len, double* dist, void* stream) {
  if (!A || !a_off || !B || !b_off || !dirs || !dir_off || !path_a || !path_b || !path_off || !path_len || !dist ||
      n_files < 0 || dim < 1 || max_frames < 1)
    return fail(EVC_ERR_INVALID_ARGUMENT, "evc_dtw: bad argument");
  if (n_files == 0) return EVC_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = (size_t)3 * max_frames * sizeof(double);
  if (smem > 200 * 1024) return fail(EVC_ERR_UNSUPPORTED, "evc_dtw: files longer than %d frames are not supported", (int)(200 * 1024 / 24));
  static bool configured[64] = {false};
  int dev = 0;
  EVC_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    EVC_CUDA(cudaFuncSetAttribute(dtw::forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  dtw::forward_kernel<<<n_files, dtw::kThreads, smem, s>>>(A, a_off, B, b_off, dim, dirs, dir_off, dist, max_frames);
  EVC_LAUNCH_CHECK();
  dtw::traceback_kernel<<<ceil_div(n_files, 64), 64, 0, s>>>(a_off, b_off, dirs, dir_off, path_a, path_b, path_off, path_len, n_files);
  EVC_LAUNCH_CHECK();
  return EVC_OK;
}

int evc_objective(evc_dict_t d, const float* X, int ldX, int T, const float* H, int ldH, int loss, float epsilon,
                  double* out, void* stream) {
  if (!d || !X || !H || !out || T < 1) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_objective: bad argument");
  if (ldX < d->F || ldH < d->N) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_objective: ldX < F or ldH < N");
  if (loss != EVC_LOSS_KL && loss != EVC_LOSS_FROBENIUS) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_objective: bad loss");
  cudaStream_t s = (cudaStream_t)stream;
  g_prof = &d->prof;
  EVC_TRY(reserve_workspace(d, T, ldH, false));
  EVC_TRY(tc::check_alignment(d->mode, H, ldH));
  EVC_TRY(tc::after_h_written(d->tc_ops, d->mode, H, ldH, T, &d->tcws, s));
  std::vector<int> seg = {0, T};
  std::vector<double> err;
  EVC_TRY(objective_segments(d, X, ldX, T, H, ldH, loss, epsilon > 0.f ? epsilon : kEpsilon, seg, err, s));
  *out = err[0];
  return EVC_OK;
}

int evc_factorize_convert_host(evc_dict_t d, const float* X, int ldX, int T, float* H, int ldH, float* Y, int ldY,
                               const evc_solve_params* p, evc_solve_result* res, void* stream) {
  if (!d || !X || !p || T < 0) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_factorize_convert_host: null argument");
  if (ldX < d->F || (H && ldH < d->N) || (Y && ldY < d->F))
    return fail(EVC_ERR_INVALID_ARGUMENT, "evc_factorize_convert_host: pitch smaller than the row length");
  if (Y && !d->has_target) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_factorize_convert_host: no target dictionary");
  if (p->init == EVC_INIT_GIVEN && !H) return fail(EVC_ERR_INVALID_ARGUMENT, "init = GIVEN needs H");
  if (T == 0) return EVC_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int ldXd = round_up(d->F, 4), ldHd = round_up(d->N, 32), ldYd = round_up(d->F, 4);
  EVC_TRY(d->hostX.reserve((size_t)T * ldXd * sizeof(float)));
  EVC_TRY(d->hostH.reserve((size_t)T * ldHd * sizeof(float)));
  if (Y) EVC_TRY(d->hostY.reserve((size_t)T * ldYd * sizeof(float)));
  float *dX = d->hostX.as<float>(), *dH = d->hostH.as<float>(), *dY = d->hostY.as<float>();
  EVC_CUDA(cudaMemcpy2DAsync(dX, (size_t)ldXd * 4, X, (size_t)ldX * 4, (size_t)d->F * 4, T, cudaMemcpyHostToDevice, s));
  if (p->init == EVC_INIT_GIVEN)
    EVC_CUDA(cudaMemcpy2DAsync(dH, (size_t)ldHd * 4, H, (size_t)ldH * 4, (size_t)d->N * 4, T, cudaMemcpyHostToDevice, s));
  EVC_TRY(evc_solve(d, dX, ldXd, T, dH, ldHd, p, res, s));
  if (Y) {
    EVC_TRY(evc_convert(d, dH, ldHd, T, dY, ldYd, s));
    EVC_CUDA(cudaMemcpy2DAsync(Y, (size_t)ldY * 4, dY, (size_t)ldYd * 4, (size_t)d->F * 4, T, cudaMemcpyDeviceToHost, s));
  }
  if (H)
    EVC_CUDA(cudaMemcpy2DAsync(H, (size_t)ldH * 4, dH, (size_t)ldHd * 4, (size_t)d->N * 4, T, cudaMemcpyDeviceToHost, s));
  EVC_CUDA(cudaStreamSynchronize(s));
  return EVC_OK;
}

int evc_comm_unique_id(char id_out[128]) { return nccl::unique_id(id_out); }

int evc_comm_create(const char id[128], int rank, int world, evc_comm_t* out) {
  if (!out || !id || world < 1 || rank < 0 || rank >= world)
    return fail(EVC_ERR_INVALID_ARGUMENT, "evc_comm_create: bad argument");
  evc_comm* c = new (std::nothrow) evc_comm();
  if (!c) return fail(EVC_ERR_CUDA, "out of host memory");
  c->rank = rank; c->world = world;
  int st = nccl::init_rank(&c->comm, id, rank, world);
  if (st != EVC_OK) { delete c; return st; }
  *out = c;
  return EVC_OK;
}

int evc_comm_destroy(evc_comm_t c) {
  if (!c) return EVC_OK;
  nccl::destroy(c->comm);
  delete c;
  return EVC_OK;
}

int evc_dict_attach_comm(evc_dict_t d, evc_comm_t c, int n_total) {
  if (!d) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_dict_attach_comm: null handle");
  if (c && n_total < d->N) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_dict_attach_comm: n_total < local N");
  d->comm = c;
  d->n_total = c ? n_total : d->N;
  return EVC_OK;
}

int evc_p2p_alloc(evc_dict_t d, int max_frames, char handle_out[64]) {
  if (!d || !handle_out || max_frames < 1) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_p2p_alloc: bad argument");
  return p2p::alloc_local(&d->p2p, max_frames, round_up(d->F, 4), handle_out);
}

int evc_p2p_attach(evc_dict_t d, const char* handles, int rank, int world) {
  if (!d || !handles) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_p2p_attach: null argument");
  if (!d->comm || d->comm->world != world || d->comm->rank != rank)
    return fail(EVC_ERR_INVALID_ARGUMENT, "evc_p2p_attach: call evc_dict_attach_comm with the same rank/world first");
  return p2p::attach(&d->p2p, handles, rank, world);
}

int evc_p2p_detach(evc_dict_t d) {
  if (!d) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_p2p_detach: null handle");
  p2p::detach(&d->p2p);
  return EVC_OK;
}

int evc_gather_stack(const float* frames, int ld, int n_frames, int F, const int* idx, const int* lo, const int* hi,
                     int n_out, int context, float* out, int ld_out, void* stream) {
  if (!frames || !idx || !lo || !hi || !out || F < 1 || n_frames < 1 || n_out < 0 || context < 0 || ld < F ||
      ld_out < (2 * context + 1) * F)
    return fail(EVC_ERR_INVALID_ARGUMENT, "evc_gather_stack: bad argument");
  if (n_out == 0) return EVC_OK;
  const int width = (2 * context + 1) * F;
  dim3 g(n_out, ceil_div(width, 1024) < 1 ? 1 : ceil_div(width, 1024));
  simt::gather_stack_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(frames, ld, F, idx, lo, hi, n_out, context, out, ld_out);
  EVC_LAUNCH_CHECK();
  return EVC_OK;
}

int evc_profile_enable(evc_dict_t d, int on) {
  if (!d) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_profile_enable: null handle");
  d->prof.on = on != 0;
  d->prof.recs.clear();
  d->prof.used = 0;
  return EVC_OK;
}

int evc_profile_read(evc_dict_t d, double* ms, int* launches) {
  if (!d || !ms || !launches) return fail(EVC_ERR_INVALID_ARGUMENT, "evc_profile_read: null argument");
  for (int c = 0; c < EVC_PROFILE_CLASSES; ++c) { ms[c] = 0.0; launches[c] = 0; }
  if (d->prof.used) EVC_CUDA(cudaEventSynchronize(d->prof.ev[d->prof.used - 1]));
  for (const Profiler::Rec& r : d->prof.recs) {
    float t = 0.f;
    EVC_CUDA(cudaEventElapsedTime(&t, d->prof.ev[r.a], d->prof.ev[r.b]));
    if (r.cls >= 0 && r.cls < EVC_PROFILE_CLASSES) { ms[r.cls] += t; launches[r.cls] += 1; }
  }
  d->prof.recs.clear();
  d->prof.used = 0;
  return EVC_OK;
}

}  // extern "C"
