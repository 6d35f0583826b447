// K6 / K7: predictive distributions per posterior draw and Philox-seeded posterior-predictive
// sampling.
//
//   ci_filter_predict   tfp.sts.one_step_predictive  (/root/reference/causalimpact/model.py:414-418)
//                       + the forward_filter half of tfp.sts.forecast (model.py:446-451)
//   ci_onestep_sample   one_step_dist.sample(n) / .mean()     (causalimpact/inferences.py:108,119)
//   ci_forecast_sample  posterior_dist.sample(n) / .mean()    (inferences.py:113,134; main.py:377)
//
// Sampling semantics: a sample first picks one posterior draw j uniformly (mixture
// over draws), then -- one-step: y_t ~ N(mean_jt, var_jt) independently over t; forecast:
// x_0 ~ N(m_j, P_j), then the state-space model is rolled forward Tf steps with transition and
// observation noise (one joint trajectory).  Outputs are written in ORIGINAL units
// (y * sigma + mu, causalimpact/misc.py:75-77) when d_mu_sig is given.
#include "ci_abi_common.cuh"
#include "ci_kalman.cuh"

namespace ci {

struct PredWs {
  float* beta;  // [max(k,1)][B]
  float* scal;  // [8][B]
  float* off;   // [T][B] offsets, then residuals in place
};
static size_t pred_ws_floats(const ci_layout* L, PredWs* ws, Carver* c) {
  const size_t B = (size_t)L->N * L->G;
  const size_t nb = (size_t)(L->k > 0 ? L->k : 1) * B, ns = 8 * B, no = (size_t)L->T * B;
  if (ws && c) {
    ws->beta = c->floats(nb);
    ws->scal = c->floats(ns);
    ws->off = c->floats(no);
  }
  return align_up(nb * 4, 256) + align_up(ns * 4, 256) + align_up(no * 4, 256);
}

struct FcWs {
  float* beta;  // [max(k,1)][B]
  float* scal;  // [8][B]
  float* off;   // [Tf][B] forecast-period offsets per draw
  float* pm;    // [Tf][B] per-draw forecast mean path
};
static size_t fc_ws_floats(const ci_layout* L, FcWs* ws, Carver* c) {
  const size_t B = (size_t)L->N * L->G;
  const size_t nb = (size_t)(L->k > 0 ? L->k : 1) * B, ns = 8 * B, no = (size_t)(L->Tf > 0 ? L->Tf : 1) * B;
  if (ws && c) {
    ws->beta = c->floats(nb);
    ws->scal = c->floats(ns);
    ws->off = c->floats(no);
    ws->pm = c->floats(no);
  }
  return align_up(nb * 4, 256) + align_up(ns * 4, 256) + 2 * align_up(no * 4, 256);
}

// Forward filter per draw lane.  off_resid: in = offset_t (mean + X_t beta), used as scratch.
// pred_mean[t][b] = H m_t|t-1 + offset_t ; pred_var[t][b] = s_t ; fstate = m, P (packed upper),
// lower Cholesky factor of P (packed, row-major lower) of the predicted state for step T.
template <int S>
__global__ void __launch_bounds__(128)
k_filter(int N, int G, int T, int r, const float* __restrict__ y, const float* __restrict__ sconst,
         const float* __restrict__ scal, const float* __restrict__ off, float* __restrict__ pred_mean,
         float* __restrict__ pred_var, float* __restrict__ fstate) {
  constexpr int NP = S * (S + 1) / 2;
  const long B = (long)N * G;
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = (int)(b / G);
  const float R = scal[0 * B + b], ql = scal[1 * B + b], qd = scal[2 * B + b];
  const float* __restrict__ yn = y + (long)n * T;
  KState<S> k;
  kf_init<S>(k, sconst[SC_M0 * N + n], sconst[SC_P0 * N + n]);
  int phase = 0;
  for (int t = 0; t < T; ++t) {
    const float o = off[(long)t * B + b];
    const float yt = __ldg(yn + t);
    float hm = k.m[0], s = k.P[sym<S>(0, 0)] + R;
    if (S > 1) {
      hm += k.m[1];
      s += 2.f * k.P[sym<S>(0, 1)] + k.P[sym<S>(1, 1)];
    }
    pred_mean[(long)t * B + b] = hm + o;
    pred_var[(long)t * B + b] = s;
    if (yt == yt) {
      float g[S], s2, v, is;
      kf_update<S>(k, yt - o, R, g, s2, v, is);
    }
    const bool change = (phase == r - 1);
    kf_predict<S>(k, ql, qd, change);
    phase = change ? 0 : phase + 1;
  }
#pragma unroll
  for (int i = 0; i < S; ++i) fstate[(long)i * B + b] = k.m[i];
#pragma unroll
  for (int i = 0; i < NP; ++i) fstate[(long)(S + i) * B + b] = k.P[i];
  // guarded Cholesky (a non-positive pivot zeroes its column), row-major packed lower factor
  float Lf[NP];
#pragma unroll
  for (int i = 0; i < S; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      float a = k.P[sym<S>(i, j)];
#pragma unroll
      for (int q = 0; q < j; ++q) a = fmaf(-Lf[i * (i + 1) / 2 + q], Lf[j * (j + 1) / 2 + q], a);
      if (i == j) {
        Lf[i * (i + 1) / 2 + j] = a > 0.f ? sqrtf(a) : 0.f;
      } else {
        const float dj = Lf[j * (j + 1) / 2 + j];
        Lf[i * (i + 1) / 2 + j] = dj > 0.f ? a / dj : 0.f;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) fstate[(long)(S + NP + i) * B + b] = Lf[i];
}

__device__ __forceinline__ uint32_t pick_draw(uint32_t sample_id, uint32_t stream, int G, uint64_t seed) {
  const float un = philox_uniform4(sample_id, 0u, stream, TAG_PICK, seed).x;
  const int j = (int)(un * (float)G);
  return (uint32_t)(j < G ? j : G - 1);
}

// pre_sims[n][i][t]: 128 samples of one series per CTA, time tiled by 32 through shared memory.  The
// lanes of a warp read the (mean, variance) of their picked draws of the SAME series and step (<= G
// adjacent floats: a few sectors, L1-resident) and the [nsamp][T] rows are written as full 128 B lines.
__global__ void __launch_bounds__(128)
k_onestep_sample(int N, int G, int T, int nsamp, uint64_t seed, const float* __restrict__ pm,
                 const float* __restrict__ pv, const float* __restrict__ mu_sig, float* __restrict__ sims) {
  __shared__ float tile[128][33];
  const int n = blockIdx.y;
  const int i = blockIdx.x * 128 + threadIdx.x;
  const long B = (long)N * G;
  const long si = (long)n * nsamp + (i < nsamp ? i : 0);
  const long b = (long)n * G + pick_draw((uint32_t)si, 0u, G, seed);
  const float mu = mu_sig ? mu_sig[n] : 0.f, sg = mu_sig ? mu_sig[N + n] : 1.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int tn = min(32, T - t0);
#pragma unroll
    for (int q4 = 0; q4 < 8; ++q4) {
      // counter = (sample, time quad): the same stream as one Philox call per 4 consecutive steps
      const f32x4 z = philox_normal4_fast((uint32_t)si, (uint32_t)((t0 >> 2) + q4), 0u, TAG_ONESTEP, seed);
      const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int q = q4 * 4 + e;
        if (q < tn) {
          const long o = (long)(t0 + q) * B + b;
          tile[threadIdx.x][q] = fmaf(fmaf(sqrtf(pv[o]), zz[e], pm[o]), sg, mu);
        }
      }
    }
    __syncthreads();
    for (int row = warp; row < 128; row += 4) {
      const int ii = blockIdx.x * 128 + row;
      if (ii < nsamp && lane < tn) sims[((long)n * nsamp + ii) * T + t0 + lane] = tile[row][lane];
    }
    __syncthreads();
  }
}

// out[n][t] = mean over the G draws of a[t][n*G + j], unstandardised.
__global__ void k_mean_over_draws(int N, int G, int W, const float* __restrict__ a,
                                  const float* __restrict__ mu_sig, float* __restrict__ out) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)N * W) return;
  const int t = (int)(idx % W);
  const int n = (int)(idx / W);
  const long B = (long)N * G;
  float acc = 0.f;
  for (int j = 0; j < G; ++j) acc += a[(long)t * B + (long)n * G + j];
  acc /= (float)G;
  const float mu = mu_sig ? mu_sig[n] : 0.f, sg = mu_sig ? mu_sig[N + n] : 1.f;
  out[idx] = fmaf(acc, sg, mu);
}

// Deterministic forecast mean path per draw: pm[t][b] = H F^t m_b + offset_{T+t}.
template <int S>
__global__ void k_forecast_mean(int N, int G, int T, int Tf, int r, const float* __restrict__ fstate,
                                const float* __restrict__ off, float* __restrict__ pm) {
  const long B = (long)N * G;
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  KState<S> k;
#pragma unroll
  for (int i = 0; i < S; ++i) k.m[i] = fstate[(long)i * B + b];
#pragma unroll
  for (int i = 0; i < KState<S>::NP; ++i) k.P[i] = 0.f;
  int phase = T % r;
  for (int t = 0; t < Tf; ++t) {
    pm[(long)t * B + b] = k.m[0] + (S > 1 ? k.m[1] : 0.f) + off[(long)t * B + b];
    const bool change = (phase == r - 1);
    if (S > 1 && change) {
      float zs = 0.f;
#pragma unroll
      for (int i = 1; i < S; ++i) zs += k.m[i];
#pragma unroll
      for (int i = 1; i < S - 1; ++i) k.m[i] = k.m[i + 1];
      k.m[S - 1] = -zs;
    }
    phase = change ? 0 : phase + 1;
  }
}

// One joint forecast trajectory per thread; 128 samples of one series per CTA, time tiled by 32
// through shared memory so the [N][nsamp][Tf] rows are written with full 128-byte lines.
template <int S>
__global__ void __launch_bounds__(128)
k_forecast_sample(int N, int G, int T, int Tf, int r, int nsamp, uint64_t seed,
                  const float* __restrict__ fstate, const float* __restrict__ scal,
                  const float* __restrict__ off, const float* __restrict__ mu_sig, float* __restrict__ sims) {
  constexpr int NP = S * (S + 1) / 2;
  __shared__ float tile[128][33];
  const int n = blockIdx.y;
  const int i = blockIdx.x * 128 + threadIdx.x;
  const bool live = i < nsamp;
  const long B = (long)N * G;
  const long si = (long)n * nsamp + (live ? i : 0);
  const long b = (long)n * G + pick_draw((uint32_t)si, 1u, G, seed);
  const float so = sqrtf(scal[0 * B + b]), sl = sqrtf(scal[1 * B + b]), sdr = sqrtf(scal[2 * B + b]);
  const float mu = mu_sig ? mu_sig[n] : 0.f, sg = mu_sig ? mu_sig[N + n] : 1.f;
  // x0 = m + L z
  float x[S];
  {
    float z[S];
#pragma unroll
    for (int q = 0; q < (S + 3) / 4; ++q) {
      const f32x4 zz = philox_normal4_fast((uint32_t)si, (uint32_t)q, 0u, TAG_FORECAST_X0, seed);
      if (4 * q + 0 < S) z[4 * q + 0] = zz.x;
      if (4 * q + 1 < S) z[4 * q + 1] = zz.y;
      if (4 * q + 2 < S) z[4 * q + 2] = zz.z;
      if (4 * q + 3 < S) z[4 * q + 3] = zz.w;
    }
#pragma unroll
    for (int a = 0; a < S; ++a) {
      float acc = fstate[(long)a * B + b];
#pragma unroll
      for (int c = 0; c <= a; ++c) acc = fmaf(fstate[(long)(S + NP + a * (a + 1) / 2 + c) * B + b], z[c], acc);
      x[a] = acc;
    }
  }
  int phase = T % r;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t0 = 0; t0 < Tf; t0 += 32) {
    const int tn = min(32, Tf - t0);
    for (int q = 0; q < tn; ++q) {
      const int t = t0 + q;
      const f32x4 e = philox_normal4_fast((uint32_t)si, (uint32_t)t, 0u, TAG_FORECAST, seed);
      const float yv = x[0] + (S > 1 ? x[1] : 0.f) + off[(long)t * B + b] + so * e.x;
      tile[threadIdx.x][q] = fmaf(yv, sg, mu);
      x[0] = fmaf(sl, e.y, x[0]);
      const bool change = (phase == r - 1);
      if (S > 1 && change) {
        float zs = 0.f;
#pragma unroll
        for (int a = 1; a < S; ++a) zs += x[a];
        const float xi = sdr * e.z;
#pragma unroll
        for (int a = 1; a < S - 1; ++a) x[a] = x[a + 1] + xi;
        x[S - 1] = -zs + xi;
      }
      phase = change ? 0 : phase + 1;
    }
    __syncthreads();
    for (int row = warp; row < 128; row += 4) {
      const int ii = blockIdx.x * 128 + row;
      if (ii < nsamp && lane < tn) sims[((long)n * nsamp + ii) * Tf + t0 + lane] = tile[row][lane];
    }
    __syncthreads();
  }
}

template <int S>
static void launch_filter(const ci_layout* L, const ci_data* D, const PredWs& ws, float* pm, float* pv, float* fs,
                          cudaStream_t st) {
  const long B = (long)L->N * L->G;
  k_filter<S><<<(unsigned)((B + 63) / 64), 64, 0, st>>>(L->N, L->G, L->T, L->r, D->d_y, D->d_sconst, ws.scal, ws.off,
                                                         pm, pv, fs);
}
template <int S>
static void launch_forecast(const ci_layout* L, const FcWs& ws, const float* fs, const float* mu_sig, int nsamp,
                            uint64_t seed, float* sims, cudaStream_t st) {
  const long B = (long)L->N * L->G;
  k_forecast_mean<S><<<(unsigned)((B + 127) / 128), 128, 0, st>>>(L->N, L->G, L->T, L->Tf, L->r, fs, ws.off, ws.pm);
  if (sims)
    k_forecast_sample<S><<<dim3((nsamp + 127) / 128, L->N), 128, 0, st>>>(L->N, L->G, L->T, L->Tf, L->r, nsamp, seed,
                                                                         fs, ws.scal, ws.off, mu_sig, sims);
}

}  // namespace ci

using namespace ci;

extern "C" {

size_t ci_predict_workspace_bytes(const ci_layout* L) {
  if (check_layout(L)) return 0;
  const size_t a = pred_ws_floats(L, nullptr, nullptr), b = fc_ws_floats(L, nullptr, nullptr);
  return a > b ? a : b;
}

int ci_filter_predict(const ci_layout* L, const ci_data* D, const float* d_u, float* d_pred_mean,
                      float* d_pred_var, float* d_fstate, void* d_workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_layout(L)) return e;
  CI_CHECK_ARG(D && D->d_y && D->d_sconst && D->d_X, "NULL data pointer");
  CI_CHECK_ARG(d_u && d_pred_mean && d_pred_var && d_fstate && d_workspace, "NULL pointer");
  if (!ci_reg_path(L) && !big_supported(L))
    return fail(CI_ERR_UNSUPPORTED, "nseasons=%d not supported by this build", L->S);
  if (workspace_bytes < pred_ws_floats(L, nullptr, nullptr))
    return fail(CI_ERR_WORKSPACE_TOO_SMALL, "workspace needs %zu bytes", pred_ws_floats(L, nullptr, nullptr));
  cudaStream_t st = (cudaStream_t)stream;
  Carver c(d_workspace);
  PredWs ws;
  pred_ws_floats(L, &ws, &c);
  if (!ci_reg_path(L)) {
    if (int e = launch_filter_big(L, D, d_u, d_pred_mean, d_pred_var, d_fstate, st)) return e;
    CI_CHECK_LAUNCH();
    return CI_OK;
  }
  launch_params_forward(L, D, d_u, ws.beta, ws.scal, st);
  launch_offsets(L, D, ws.beta, ws.off, 0, L->T, 0, st);
  CI_DISPATCH_S(L->S, (launch_filter<S_>(L, D, ws, d_pred_mean, d_pred_var, d_fstate, st)));
  CI_CHECK_LAUNCH();
  return CI_OK;
}

int ci_onestep_sample(const ci_layout* L, const float* d_pred_mean, const float* d_pred_var, const float* d_mu_sig,
                      int32_t nsamp, uint64_t seed, float* d_pre_sims, float* d_pre_mean, void* stream) {
  if (int e = check_layout(L)) return e;
  CI_CHECK_ARG(d_pred_mean && d_pred_var, "NULL pointer");
  CI_CHECK_ARG(nsamp >= 0 && (nsamp == 0 || d_pre_sims), "d_pre_sims is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (nsamp > 0) {
    k_onestep_sample<<<dim3((nsamp + 127) / 128, L->N), 128, 0, st>>>(L->N, L->G, L->T, nsamp, seed, d_pred_mean,
                                                                      d_pred_var, d_mu_sig, d_pre_sims);
  }
  if (d_pre_mean) {
    const long total = (long)L->N * L->T;
    k_mean_over_draws<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(L->N, L->G, L->T, d_pred_mean, d_mu_sig,
                                                                       d_pre_mean);
  }
  CI_CHECK_LAUNCH();
  return CI_OK;
}

int ci_forecast_sample(const ci_layout* L, const ci_data* D, const float* d_u, const float* d_fstate,
                       const float* d_mu_sig, int32_t nsamp, uint64_t seed, float* d_post_sims,
                       float* d_post_mean, void* d_workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_layout(L)) return e;
  CI_CHECK_ARG(D && D->d_y && D->d_sconst && D->d_X, "NULL data pointer");
  CI_CHECK_ARG(d_u && d_fstate && d_workspace, "NULL pointer");
  CI_CHECK_ARG(L->Tf >= 1, "forecast horizon Tf must be >= 1");
  CI_CHECK_ARG(nsamp >= 0 && (nsamp == 0 || d_post_sims), "d_post_sims is NULL");
  if (!ci_reg_path(L) && !big_supported(L))
    return fail(CI_ERR_UNSUPPORTED, "nseasons=%d not supported by this build", L->S);
  if (workspace_bytes < fc_ws_floats(L, nullptr, nullptr))
    return fail(CI_ERR_WORKSPACE_TOO_SMALL, "workspace needs %zu bytes", fc_ws_floats(L, nullptr, nullptr));
  cudaStream_t st = (cudaStream_t)stream;
  Carver c(d_workspace);
  FcWs ws;
  fc_ws_floats(L, &ws, &c);
  launch_params_forward(L, D, d_u, ws.beta, ws.scal, st);
  launch_offsets(L, D, ws.beta, ws.off, L->T, L->T + L->Tf, 0, st);
  if (!ci_reg_path(L)) {
    launch_forecast_big(L, d_fstate, ws.scal, ws.off, ws.pm, d_mu_sig, nsamp, seed, nsamp > 0 ? d_post_sims : nullptr, st);
  } else {
    CI_DISPATCH_S(L->S, (launch_forecast<S_>(L, ws, d_fstate, d_mu_sig, nsamp, seed, nsamp > 0 ? d_post_sims : nullptr, st)));
  }
  if (d_post_mean) {
    const long total = (long)L->N * L->Tf;
    k_mean_over_draws<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(L->N, L->G, L->Tf, ws.pm, d_mu_sig, d_post_mean);
  }
  CI_CHECK_LAUNCH();
  return CI_OK;
}

}  // extern "C"

/* Floats per lane of the forecast-state array ci_filter_predict writes (layout is path-specific:
 * [nfs][B] lane-fastest for the register-resident kernels, lane-contiguous [B][nfs] otherwise). */
extern "C" size_t ci_fstate_floats(const ci_layout* L) {
  if (check_layout(L)) return 0;
  if (ci_reg_path(L)) return (size_t)L->S + (size_t)L->S * (L->S + 1);
  return big_fstate_floats(L);
}
