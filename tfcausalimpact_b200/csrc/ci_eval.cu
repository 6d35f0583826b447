// K1-K3: joint log-prob + gradient of the fixed-layout structural time-series model.
//   k_params_forward : u -> (R, ql, qd, prior+logdet), beta                     [thread per lane]
//   k_offsets        : r_t = y_t - mean - X_t . beta                            [lane x time tile]
//   k_kalman<S>      : Kalman log-lik + reverse-time adjoint, registers only    [thread per lane]
//   k_beta_bar       : dL/dbeta_j = -sum_t rbar_t X_tj                          [lane x covariate tile]
//   k_params_backward: VJP through horseshoe / softplus / sqrt + prior gradient [thread per lane]
// Lanes are the fastest axis of every array, so each warp access is one 128 B line; the packed
// design matrix is time-contiguous so a thread pulls 4 steps of one covariate per 16 B load.
#include "ci_abi_common.cuh"
#include "ci_kalman.cuh"

namespace ci {

char* last_error_buf() {
  static thread_local char buf[512] = "";
  return buf;
}

__global__ void k_params_forward(int N, int G, int k, int S, const float* __restrict__ sconst,
                                 const float* __restrict__ u, float* __restrict__ beta,
                                 float* __restrict__ scal) {
  const long B = (long)N * G;
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = (int)(b / G);
  LaneScalars ls = params_forward_lane(u + b, B, k, S, sconst + n, N, beta + b);
  scal[0 * B + b] = ls.R;
  scal[1 * B + b] = ls.ql;
  scal[2 * B + b] = ls.qd;
  scal[3 * B + b] = ls.lp0;
}

// out[t - t_lo][b] for t in [t_lo, t_hi):  with_y ? y[n][t] - mean[n] - sum_j beta[j][b] X[n][j][t]
//                                                 : mean[n] + sum_j beta[j][b] X[n][j][t];
// 4 time steps of one covariate per 16-byte X load.
__global__ void k_offsets(ci_layout L, const float* __restrict__ y, const float* __restrict__ X,
                          const float* __restrict__ sconst, const float* __restrict__ beta,
                          float* __restrict__ out, int t_lo, int t_hi, int t_chunk, int with_y) {
  const long B = (long)L.N * L.G;
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = (int)(b / L.G);
  const int tbeg = (t_lo / 8) * 8 + blockIdx.y * t_chunk;
  const int tend = min(t_hi, tbeg + t_chunk);
  const float mean = sconst[SC_MEAN * L.N + n];
  const float* __restrict__ Xn = X + (long)n * L.k * L.Tpad;
  const float* __restrict__ yn = y + (long)n * L.T;
  for (int t0 = tbeg; t0 < tend; t0 += 8) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = -mean;
    const bool second = (t0 + 4 < L.Tpad);
    for (int j = 0; j < L.k; ++j) {
      const float bj = beta[(long)j * B + b];
      const float4 x0 = __ldg(reinterpret_cast<const float4*>(Xn + (long)j * L.Tpad + t0));
      a[0] = fmaf(-bj, x0.x, a[0]); a[1] = fmaf(-bj, x0.y, a[1]);
      a[2] = fmaf(-bj, x0.z, a[2]); a[3] = fmaf(-bj, x0.w, a[3]);
      if (second) {
        const float4 x1 = __ldg(reinterpret_cast<const float4*>(Xn + (long)j * L.Tpad + t0 + 4));
        a[4] = fmaf(-bj, x1.x, a[4]); a[5] = fmaf(-bj, x1.y, a[5]);
        a[6] = fmaf(-bj, x1.z, a[6]); a[7] = fmaf(-bj, x1.w, a[7]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int t = t0 + i;
      if (t >= t_lo && t < tend) out[(long)(t - t_lo) * B + b] = with_y ? __ldg(yn + t) + a[i] : -a[i];
    }
  }
}

void launch_offsets(const ci_layout* L, const ci_data* D, const float* beta, float* out, int t_lo, int t_hi,
                    int with_y, cudaStream_t st) {
  const long B = (long)L->N * L->G;
  const int bs = 128;
  const unsigned gb = (unsigned)((B + bs - 1) / bs);
  // enough (lane-block x time-chunk) CTAs to cover the machine a few times over
  const int span = t_hi - (t_lo / 8) * 8;
  int chunks = 1;
  while ((long)gb * chunks < 148 * 8 && chunks * 8 < span) chunks *= 2;
  const int t_chunk = ((span + chunks - 1) / chunks + 7) / 8 * 8;
  chunks = (span + t_chunk - 1) / t_chunk;
  k_offsets<<<dim3(gb, chunks), bs, 0, st>>>(*L, D->d_y, D->d_X, D->d_sconst, beta, out, t_lo, t_hi, t_chunk, with_y);
}

void launch_params_forward(const ci_layout* L, const ci_data* D, const float* d_u, float* beta, float* scal,
                           cudaStream_t st) {
  const long B = (long)L->N * L->G;
  k_params_forward<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(L->N, L->G, L->k, L->S, D->d_sconst, d_u, beta, scal);
}

template <int S>
__global__ void __launch_bounds__(128)
k_kalman(int N, int G, int T, int r, const float* __restrict__ sconst, float* __restrict__ scal,
         float* __restrict__ resid, float* __restrict__ tape) {
  const long B = (long)N * G;
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = (int)(b / G);
  float ll, Rb, qlb, qdb;
  kalman_logp_grad_lane<S>(resid + b, tape + b, B, T, r, scal[0 * B + b], scal[1 * B + b], scal[2 * B + b],
                           sconst[SC_M0 * N + n], sconst[SC_P0 * N + n], ll, Rb, qlb, qdb);
  scal[4 * B + b] = ll;
  scal[5 * B + b] = Rb;
  scal[6 * B + b] = qlb;
  scal[7 * B + b] = qdb;
}

// beta_bar[j][b] = -sum_t rbar[t][b] X[n][j][t]; JT covariates per thread.
template <int JT>
__global__ void k_beta_bar(ci_layout L, const float* __restrict__ X, const float* __restrict__ rbar,
                           float* __restrict__ beta_bar) {
  const long B = (long)L.N * L.G;
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = (int)(b / L.G);
  const int j0 = blockIdx.y * JT;
  const float* __restrict__ Xn = X + ((long)n * L.k + j0) * L.Tpad;
  float acc[JT];
#pragma unroll
  for (int i = 0; i < JT; ++i) acc[i] = 0.f;
  for (int t0 = 0; t0 < L.T; t0 += 4) {
    float r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = (t0 + i < L.T) ? rbar[(long)(t0 + i) * B + b] : 0.f;
#pragma unroll
    for (int jj = 0; jj < JT; ++jj) {
      if (j0 + jj < L.k) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(Xn + (long)jj * L.Tpad + t0));
        acc[jj] = fmaf(-r[0], x.x, acc[jj]);
        acc[jj] = fmaf(-r[1], x.y, acc[jj]);
        acc[jj] = fmaf(-r[2], x.z, acc[jj]);
        acc[jj] = fmaf(-r[3], x.w, acc[jj]);
      }
    }
  }
#pragma unroll
  for (int jj = 0; jj < JT; ++jj)
    if (j0 + jj < L.k) beta_bar[(long)(j0 + jj) * B + b] = acc[jj];
}

__global__ void k_params_backward(int N, int G, int k, int S, const float* __restrict__ sconst,
                                  const float* __restrict__ u, const float* __restrict__ beta_bar,
                                  const float* __restrict__ scal, float* __restrict__ logp,
                                  float* __restrict__ grad) {
  const long B = (long)N * G;
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = (int)(b / G);
  params_backward_lane(u + b, B, k, S, sconst + n, N, beta_bar + b, scal[5 * B + b], scal[6 * B + b],
                       scal[7 * B + b], grad + b);
  logp[b] = scal[3 * B + b] + scal[4 * B + b];
}

template <int S>
static void launch_kalman(const ci_layout* L, const ci_data* D, const EvalWs& ws, cudaStream_t st) {
  const long B = (long)L->N * L->G;
  const int bs = 64;
  k_kalman<S><<<(unsigned)((B + bs - 1) / bs), bs, 0, st>>>(L->N, L->G, L->T, L->r, D->d_sconst, ws.scal,
                                                             ws.resid, ws.tape);
}

int launch_eval(const ci_layout* L, const ci_data* D, const float* d_u, float* d_logp, float* d_grad,
                const EvalWs& ws, cudaStream_t st) {
  const long B = (long)L->N * L->G;
  const int bs = 128;
  const unsigned gb = (unsigned)((B + bs - 1) / bs);
  k_params_forward<<<gb, bs, 0, st>>>(L->N, L->G, L->k, L->S, D->d_sconst, d_u, ws.beta, ws.scal);
  launch_offsets(L, D, ws.beta, ws.resid, 0, L->T, 1, st);
  CI_DISPATCH_S(L->S, (launch_kalman<S_>(L, D, ws, st)));
  if (L->k > 0) {
    constexpr int JT = 8;
    k_beta_bar<JT><<<dim3(gb, (L->k + JT - 1) / JT), bs, 0, st>>>(*L, D->d_X, ws.resid, ws.beta);
  }
  k_params_backward<<<gb, bs, 0, st>>>(L->N, L->G, L->k, L->S, D->d_sconst, d_u, ws.beta, ws.scal, d_logp,
                                       d_grad);
  CI_CHECK_LAUNCH();
  return CI_OK;
}

// ---- statistics / packing ----------------------------------------------------------------------

__global__ void k_series_stats(int N, int T, int k, int S, double prior_level_sd,
                               const float* __restrict__ y, float* __restrict__ sconst) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* yn = y + (long)n * T;
  double sum = 0.0;
  int cnt = 0;
  float first = 0.f;
  bool have_first = false;
  for (int t = 0; t < T; ++t) {
    const float v = yn[t];
    if (v == v) {
      sum += v;
      ++cnt;
      if (!have_first) { first = v; have_first = true; }
    }
  }
  const double mean = sum / (double)cnt;
  double ss = 0.0;
  for (int t = 0; t < T; ++t) {
    const float v = yn[t];
    if (v == v) ss += ((double)v - mean) * ((double)v - mean);
  }
  const double sd = sqrt(ss / (double)cnt);
  const double y0c = (double)first - mean;
  const double s0 = fabs(y0c) + sd;
  const double a = 16.0;
  const double obs_b = a * (sd * 0.5) * (sd * 0.5);
  const double level_b = a * prior_level_sd * prior_level_sd;
  double c = (a * log(obs_b) - lgamma(a)) + (a * log(level_b) - lgamma(a)) + 2.0 * log(2.0) + 2.0 * log(sd);
  if (k > 0) {
    c += (1.0 + k) * (0.5 * log(0.5) - lgamma(0.5)) + (1.0 + k) * 0.5 * log(2.0 / 3.141592653589793) -
         k * 0.5 * 1.8378770664093453;
  }
  if (S > 1) c += -log(3.0) - 0.5 * 1.8378770664093453 + log(sd);
  sconst[SC_MEAN * N + n] = (float)mean;
  sconst[SC_SD * N + n] = (float)sd;
  sconst[SC_M0 * N + n] = (float)y0c;
  sconst[SC_P0 * N + n] = (float)(s0 * s0);
  sconst[SC_OBS_B * N + n] = (float)obs_b;
  sconst[SC_LEVEL_B * N + n] = (float)level_b;
  sconst[SC_DRIFT_MU * N + n] = (float)log(0.01 * sd);
  sconst[SC_PRIOR_CONST * N + n] = (float)c;
}

__global__ void k_pack_design(ci_layout L, const float* __restrict__ Xrm, float* __restrict__ Xp) {
  const long total = (long)L.N * L.k * L.Tpad;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int t = (int)(i % L.Tpad);
  const long nj = i / L.Tpad;
  const int j = (int)(nj % L.k);
  const long n = nj / L.k;
  float v = 0.f;
  if (t < L.T + L.Tf) {
    v = Xrm[(n * (L.T + L.Tf) + t) * L.k + j];
    if (!(v == v)) v = 0.f;
  }
  Xp[i] = v;
}

__global__ void k_constrain(int N, int G, int k, int S, const float* __restrict__ sconst,
                            const float* __restrict__ u, float* __restrict__ theta) {
  const long B = (long)N * G;
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = (int)(b / G);
  params_constrain_lane(u + b, B, k, S, sconst[SC_SD * N + n], theta + b, B);
}

}  // namespace ci

using namespace ci;

extern "C" {

int ci_version(void) { return 100; }
const char* ci_last_error_string(void) { return last_error_buf(); }
int ci_num_params(int32_t k, int32_t S) { return num_params(k, S); }

int ci_series_stats(const ci_layout* L, const float* d_y, float prior_level_sd, float* d_sconst, void* stream) {
  if (int e = check_layout(L)) return e;
  CI_CHECK_ARG(d_y && d_sconst, "NULL pointer");
  CI_CHECK_ARG(prior_level_sd > 0.f, "prior_level_sd must be positive");
  k_series_stats<<<(L->N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(L->N, L->T, L->k, L->S,
                                                                        (double)prior_level_sd, d_y, d_sconst);
  CI_CHECK_LAUNCH();
  return CI_OK;
}

int ci_pack_design(const ci_layout* L, const float* d_X_rowmajor, float* d_X_packed, void* stream) {
  if (int e = check_layout(L)) return e;
  if (L->k == 0) return CI_OK;
  CI_CHECK_ARG(d_X_rowmajor && d_X_packed, "NULL pointer");
  const long total = (long)L->N * L->k * L->Tpad;
  k_pack_design<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*L, d_X_rowmajor, d_X_packed);
  CI_CHECK_LAUNCH();
  return CI_OK;
}

size_t ci_eval_workspace_bytes(const ci_layout* L) {
  if (check_layout(L)) return 0;
  return eval_ws_floats(L, nullptr, nullptr);
}

int ci_kalman_logp_grad(const ci_layout* L, const ci_data* D, const float* d_u, float* d_logp, float* d_grad,
                        void* d_workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_layout(L)) return e;
  CI_CHECK_ARG(D && D->d_y && D->d_sconst && (L->k == 0 || D->d_X), "NULL data pointer");
  CI_CHECK_ARG(d_u && d_logp && d_grad && d_workspace, "NULL pointer");
  if (L->S > CI_MAX_REG_S) return fail(CI_ERR_UNSUPPORTED, "nseasons=%d not supported by this build", L->S);
  if (workspace_bytes < ci_eval_workspace_bytes(L))
    return fail(CI_ERR_WORKSPACE_TOO_SMALL, "workspace needs %zu bytes", ci_eval_workspace_bytes(L));
  Carver c(d_workspace);
  EvalWs ws;
  eval_ws_floats(L, &ws, &c);
  return launch_eval(L, D, d_u, d_logp, d_grad, ws, (cudaStream_t)stream);
}

int ci_constrain(const ci_layout* L, const ci_data* D, const float* d_u, float* d_theta, void* stream) {
  if (int e = check_layout(L)) return e;
  CI_CHECK_ARG(D && D->d_sconst && d_u && d_theta, "NULL pointer");
  const long B = (long)L->N * L->G;
  k_constrain<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(L->N, L->G, L->k, L->S, D->d_sconst,
                                                                            d_u, d_theta);
  CI_CHECK_LAUNCH();
  return CI_OK;
}

}  // extern "C"

// ---- FP32-FMA peak probe (roofline denominator for the CUDA-core-bound kernels) -----------------
namespace ci {
__global__ void __launch_bounds__(256) k_fma_probe(int iters, float seed, float* __restrict__ out) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + (float)(threadIdx.x + i) * 1e-3f;
  const float m = 0.999f + seed * 1e-6f, c = 1e-3f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456f) out[0] = s;
}
}  // namespace ci

extern "C" int ci_bench_fma(int32_t iters, int32_t blocks, float* d_out, void* stream, double* flops_out) {
  CI_CHECK_ARG(iters > 0 && blocks > 0 && d_out, "bad arguments");
  ci::k_fma_probe<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 0.5f, d_out);
  if (flops_out) *flops_out = 2.0 * 16.0 * (double)iters * 256.0 * (double)blocks;
  CI_CHECK_LAUNCH();
  return CI_OK;
}
