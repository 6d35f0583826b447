// Component decomposition: smoothed (pre-period) and forecast (post-period) marginals of every
// structural component, per posterior draw, mixed over the draws.
//
// Replaces: tfp.sts.decompose_by_component / tfp.sts.decompose_forecast_by_component as the
// reference's notebook uses them on `ci.model`, `ci.model_samples`
// (/root/reference/notebooks/getting_started.ipynb:125,144) -- SURVEY 8(f) rank 2.
//
//   k_smooth<S>          forward filter that keeps (predicted, filtered) moments per step in HBM, then a
//                        Rauch-Tung-Striebel backward pass (dense d x d Cholesky solve per step), thread
//                        per draw lane; emits level / seasonal mean and variance per step
//   k_forecast_moments<S> prior propagation of (m, P) from the forecast state over Tf steps
//   k_mix_components      mixture over the G draws of a series: mean, sqrt(E[var + mean^2] - mean^2)
// Layouts outside the register-resident set (large nseasons, LocalLinearTrend, dense regression, a second
// Seasonal component) are served by the warp-per-lane kernels k_smooth_big / k_forecast_moments_big
// (ci_bigd.cu, Durbin-Koopman backward recursion); outputs then carry a 4th row when S2 > 1.
// Not a hot path (lanes = posterior draws, run once per fit): arithmetic is dense and rolled.
#include "ci_abi_common.cuh"
#include "ci_kalman.cuh"

namespace ci {

template <int S>
struct Dense {
  float a[S][S];
};

// F x for the residual-basis transition on a season-change step
template <int S>
__device__ void apply_F(const float (&x)[S], float (&y)[S], bool change) {
  y[0] = x[0];
  if (S > 1) {
    if (change) {
      float zs = 0.f;
      for (int i = 1; i < S; ++i) zs += x[i];
      for (int i = 1; i < S - 1; ++i) y[i] = x[i + 1];
      y[S - 1] = -zs;
    } else {
      for (int i = 1; i < S; ++i) y[i] = x[i];
    }
  }
}

template <int S>
__device__ void unpack(const KState<S>& k, Dense<S>& M) {
  for (int i = 0; i < S; ++i)
    for (int j = 0; j < S; ++j) M.a[i][j] = k.P[sym<S>(i, j)];
}

// hist layout per lane: [T][2][S + NP] (0 = predicted, 1 = filtered), element stride B
template <int S>
__global__ void __launch_bounds__(64)
k_smooth(int N, int G, int T, int r, const float* __restrict__ y, const float* __restrict__ sconst,
         const float* __restrict__ scal, const float* __restrict__ off, float* __restrict__ hist,
         float* __restrict__ cmean, float* __restrict__ cvar) {
  constexpr int NP = S * (S + 1) / 2;
  constexpr int W = S + NP;
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
    float* hp = hist + ((long)t * 2 * W) * B + b;
    for (int i = 0; i < S; ++i) hp[(long)i * B] = k.m[i];
    for (int i = 0; i < NP; ++i) hp[(long)(S + i) * B] = k.P[i];
    const float yt = __ldg(yn + t);
    if (yt == yt) {
      float g[S], s2, v, is;
      kf_update<S>(k, yt - off[(long)t * B + b], R, g, s2, v, is);
    }
    hp += (long)W * B;
    for (int i = 0; i < S; ++i) hp[(long)i * B] = k.m[i];
    for (int i = 0; i < NP; ++i) hp[(long)(S + i) * B] = k.P[i];
    const bool change = (phase == r - 1);
    kf_predict<S>(k, ql, qd, change);
    phase = change ? 0 : phase + 1;
  }
  // backward pass; (ms, Ps) start as the filtered moments of the last step
  float ms[S];
  Dense<S> Ps;
  {
    const float* hf = hist + (((long)(T - 1) * 2 + 1) * W) * B + b;
    KState<S> f;
    for (int i = 0; i < S; ++i) f.m[i] = hf[(long)i * B];
    for (int i = 0; i < NP; ++i) f.P[i] = hf[(long)(S + i) * B];
    for (int i = 0; i < S; ++i) ms[i] = f.m[i];
    unpack<S>(f, Ps);
  }
  auto emit = [&](int t) {
    cmean[((long)0 * T + t) * B + b] = ms[0];
    cvar[((long)0 * T + t) * B + b] = Ps.a[0][0];
    cmean[((long)1 * T + t) * B + b] = S > 1 ? ms[S > 1 ? 1 : 0] : 0.f;
    cvar[((long)1 * T + t) * B + b] = S > 1 ? Ps.a[S > 1 ? 1 : 0][S > 1 ? 1 : 0] : 0.f;
  };
  emit(T - 1);
  phase = (T - 1) % r;
  for (int t = T - 2; t >= 0; --t) {
    phase = (phase == 0) ? r - 1 : phase - 1;            // phase of step t
    const bool change = (phase == r - 1);
    const float* hf = hist + (((long)t * 2 + 1) * W) * B + b;          // filtered at t
    const float* hq = hist + (((long)(t + 1) * 2) * W) * B + b;        // predicted at t + 1
    KState<S> f, q;
    for (int i = 0; i < S; ++i) { f.m[i] = hf[(long)i * B]; q.m[i] = hq[(long)i * B]; }
    for (int i = 0; i < NP; ++i) { f.P[i] = hf[(long)(S + i) * B]; q.P[i] = hq[(long)(S + i) * B]; }
    Dense<S> Pf, Pq, A, Lc, J;
    unpack<S>(f, Pf);
    unpack<S>(q, Pq);
    // A = Pf F^T  (row i of A = F applied to row i of Pf)
    for (int i = 0; i < S; ++i) apply_F<S>(Pf.a[i], A.a[i], change);
    // Cholesky of Pq (SPD in the residual basis)
    for (int i = 0; i < S; ++i)
      for (int j = 0; j <= i; ++j) {
        float a = Pq.a[i][j];
        for (int c = 0; c < j; ++c) a = fmaf(-Lc.a[i][c], Lc.a[j][c], a);
        Lc.a[i][j] = (i == j) ? sqrtf(fmaxf(a, 1e-30f)) : a / Lc.a[j][j];
      }
    // J = A Pq^-1: for every row a of A solve Pq x = a  (L z = a, L^T x = z)
    for (int i = 0; i < S; ++i) {
      float z[S];
      for (int c = 0; c < S; ++c) {
        float a = A.a[i][c];
        for (int e = 0; e < c; ++e) a = fmaf(-Lc.a[c][e], z[e], a);
        z[c] = a / Lc.a[c][c];
      }
      for (int c = S - 1; c >= 0; --c) {
        float a = z[c];
        for (int e = c + 1; e < S; ++e) a = fmaf(-Lc.a[e][c], J.a[i][e], a);
        J.a[i][c] = a / Lc.a[c][c];
      }
    }
    // ms = mf + J (ms_next - m_pred);  Ps = Pf + J (Ps_next - P_pred) J^T
    float dm[S], nm[S];
    for (int i = 0; i < S; ++i) dm[i] = ms[i] - q.m[i];
    for (int i = 0; i < S; ++i) {
      float a = f.m[i];
      for (int c = 0; c < S; ++c) a = fmaf(J.a[i][c], dm[c], a);
      nm[i] = a;
    }
    Dense<S> Dp, Tm;
    for (int i = 0; i < S; ++i)
      for (int j = 0; j < S; ++j) Dp.a[i][j] = Ps.a[i][j] - Pq.a[i][j];
    for (int i = 0; i < S; ++i)
      for (int j = 0; j < S; ++j) {
        float a = 0.f;
        for (int c = 0; c < S; ++c) a = fmaf(J.a[i][c], Dp.a[c][j], a);
        Tm.a[i][j] = a;
      }
    for (int i = 0; i < S; ++i)
      for (int j = 0; j < S; ++j) {
        float a = Pf.a[i][j];
        for (int c = 0; c < S; ++c) a = fmaf(Tm.a[i][c], J.a[j][c], a);
        Ps.a[i][j] = a;
      }
    for (int i = 0; i < S; ++i) ms[i] = nm[i];
    emit(t);
  }
}

// Prior propagation over the forecast horizon from the packed forecast state (m, P).
template <int S>
__global__ void k_forecast_moments(int N, int G, int T, int Tf, int r, const float* __restrict__ fstate,
                                   const float* __restrict__ scal, float* __restrict__ cmean,
                                   float* __restrict__ cvar) {
  constexpr int NP = S * (S + 1) / 2;
  const long B = (long)N * G;
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float ql = scal[1 * B + b], qd = scal[2 * B + b];
  KState<S> k;
  for (int i = 0; i < S; ++i) k.m[i] = fstate[(long)i * B + b];
  for (int i = 0; i < NP; ++i) k.P[i] = fstate[(long)(S + i) * B + b];
  int phase = T % r;
  for (int t = 0; t < Tf; ++t) {
    cmean[((long)0 * Tf + t) * B + b] = k.m[0];
    cvar[((long)0 * Tf + t) * B + b] = k.P[sym<S>(0, 0)];
    cmean[((long)1 * Tf + t) * B + b] = S > 1 ? k.m[S > 1 ? 1 : 0] : 0.f;
    cvar[((long)1 * Tf + t) * B + b] = S > 1 ? k.P[sym<S>(S > 1 ? 1 : 0, S > 1 ? 1 : 0)] : 0.f;
    const bool change = (phase == r - 1);
    kf_predict<S>(k, ql, qd, change);
    phase = change ? 0 : phase + 1;
  }
}

// out_mean / out_sd [ncomp][N][W] (level, regression, seasonal[, second seasonal]) from per-draw moments
// [ncomp - 1][W][B] and the per-draw offsets [W][B] (regression component = offset - mean(y),
// deterministic per draw).
__global__ void k_mix_components(int N, int G, int W, int has_reg, int ncomp, const float* __restrict__ cmean,
                                 const float* __restrict__ cvar, const float* __restrict__ off,
                                 const float* __restrict__ sconst, float* __restrict__ out_mean,
                                 float* __restrict__ out_sd) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)ncomp * N * W) return;
  const int t = (int)(idx % W);
  const int n = (int)((idx / W) % N);
  const int comp = (int)(idx / ((long)N * W));     // 0 level, 1 regression, 2 seasonal
  const long B = (long)N * G;
  double s1 = 0.0, s2 = 0.0;
  for (int j = 0; j < G; ++j) {
    const long b = (long)n * G + j;
    float m, v;
    if (comp == 1) {
      m = has_reg ? off[(long)t * B + b] - sconst[SC_MEAN * N + n] : 0.f;
      v = 0.f;
    } else {
      const int c = comp == 0 ? 0 : comp - 1;
      m = cmean[((long)c * W + t) * B + b];
      v = cvar[((long)c * W + t) * B + b];
    }
    s1 += m;
    s2 += (double)v + (double)m * m;
  }
  const double mean = s1 / G;
  const double var = s2 / G - mean * mean;
  out_mean[idx] = (float)mean;
  out_sd[idx] = (float)sqrt(var > 0.0 ? var : 0.0);
}

struct DecWs {
  float* beta;   // [max(k,1)][B]
  float* scal;   // [8][B]
  float* off;    // [W][B]
  float* cmean;  // [2][W][B]  (warp-per-lane path: [3][W][B] = level, seasonal, second seasonal)
  float* cvar;   // same shape
  float* hist;   // [T][2][S+NP][B]   (smoothing only; warp-per-lane path: [B][T][3 D + 5])
};
static size_t dec_ws_floats(const ci_layout* L, int W, bool smooth, DecWs* ws, Carver* c) {
  const size_t B = (size_t)L->N * L->G;
  const bool reg = ci_reg_path(L);
  const size_t nb = (size_t)(L->k > 0 ? L->k : 1) * B, ns = 8 * B, no = (size_t)W * B, nc = (reg ? 2 : 3) * (size_t)W * B;
  const size_t S = L->S;
  const size_t nh = !smooth ? 0 : (reg ? (size_t)L->T * 2 * (S + S * (S + 1) / 2) * B : big_hist_floats(L));
  if (ws && c) {
    ws->beta = c->floats(nb);
    ws->scal = c->floats(ns);
    ws->off = c->floats(no);
    ws->cmean = c->floats(nc);
    ws->cvar = c->floats(nc);
    ws->hist = smooth ? c->floats(nh) : nullptr;
  }
  return align_up(nb * 4, 256) + align_up(ns * 4, 256) + align_up(no * 4, 256) + 2 * align_up(nc * 4, 256) +
         (smooth ? align_up(nh * 4, 256) : 0);
}

template <int S>
static void launch_smooth(const ci_layout* L, const ci_data* D, const DecWs& ws, cudaStream_t st) {
  const long B = (long)L->N * L->G;
  k_smooth<S><<<(unsigned)((B + 63) / 64), 64, 0, st>>>(L->N, L->G, L->T, L->r, D->d_y, D->d_sconst, ws.scal, ws.off,
                                                         ws.hist, ws.cmean, ws.cvar);
}
template <int S>
static void launch_fmom(const ci_layout* L, const float* fs, const DecWs& ws, cudaStream_t st) {
  const long B = (long)L->N * L->G;
  k_forecast_moments<S><<<(unsigned)((B + 127) / 128), 128, 0, st>>>(L->N, L->G, L->T, L->Tf, L->r, fs, ws.scal,
                                                                      ws.cmean, ws.cvar);
}

}  // namespace ci

using namespace ci;

extern "C" {

size_t ci_decompose_workspace_bytes(const ci_layout* L) {
  if (check_layout(L) || (!ci_reg_path(L) && !big_supported(L))) return 0;
  const size_t a = dec_ws_floats(L, L->T, true, nullptr, nullptr);
  const size_t b = dec_ws_floats(L, L->Tf > 0 ? L->Tf : 1, false, nullptr, nullptr);
  return a > b ? a : b;
}

int ci_decompose_by_component(const ci_layout* L, const ci_data* D, const float* d_u, float* d_comp_mean,
                              float* d_comp_sd, void* d_workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_layout(L)) return e;
  CI_CHECK_ARG(D && D->d_y && D->d_sconst && D->d_X, "NULL data pointer");
  CI_CHECK_ARG(d_u && d_comp_mean && d_comp_sd && d_workspace, "NULL pointer");
  if (!ci_reg_path(L) && !big_supported(L))
    return fail(CI_ERR_UNSUPPORTED, "decomposition: nseasons=%d not supported", L->S);
  if (workspace_bytes < dec_ws_floats(L, L->T, true, nullptr, nullptr))
    return fail(CI_ERR_WORKSPACE_TOO_SMALL, "workspace needs %zu bytes", dec_ws_floats(L, L->T, true, nullptr, nullptr));
  cudaStream_t st = (cudaStream_t)stream;
  Carver c(d_workspace);
  DecWs ws;
  dec_ws_floats(L, L->T, true, &ws, &c);
  launch_params_forward(L, D, d_u, ws.beta, ws.scal, st);
  launch_offsets(L, D, ws.beta, ws.off, 0, L->T, 0, st);
  const int ncomp = 3 + (L->S2 > 1 ? 1 : 0);
  if (ci_reg_path(L)) {
    CI_DISPATCH_S(L->S, (launch_smooth<S_>(L, D, ws, st)));
  } else {
    if (int e = launch_smooth_big(L, D, d_u, ws.hist, ws.cmean, ws.cvar, st)) return e;
  }
  const long tot = (long)ncomp * L->N * L->T;
  k_mix_components<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(L->N, L->G, L->T, L->k > 0, ncomp, ws.cmean, ws.cvar, ws.off,
                                                                  D->d_sconst, d_comp_mean, d_comp_sd);
  CI_CHECK_LAUNCH();
  return CI_OK;
}

int ci_decompose_forecast(const ci_layout* L, const ci_data* D, const float* d_u, const float* d_fstate,
                          float* d_comp_mean, float* d_comp_sd, void* d_workspace, size_t workspace_bytes,
                          void* stream) {
  if (int e = check_layout(L)) return e;
  CI_CHECK_ARG(D && D->d_y && D->d_sconst && D->d_X, "NULL data pointer");
  CI_CHECK_ARG(d_u && d_fstate && d_comp_mean && d_comp_sd && d_workspace, "NULL pointer");
  CI_CHECK_ARG(L->Tf >= 1, "forecast horizon Tf must be >= 1");
  if (!ci_reg_path(L) && !big_supported(L))
    return fail(CI_ERR_UNSUPPORTED, "decomposition: nseasons=%d not supported", L->S);
  if (workspace_bytes < dec_ws_floats(L, L->Tf, false, nullptr, nullptr))
    return fail(CI_ERR_WORKSPACE_TOO_SMALL, "workspace needs %zu bytes", dec_ws_floats(L, L->Tf, false, nullptr, nullptr));
  cudaStream_t st = (cudaStream_t)stream;
  Carver c(d_workspace);
  DecWs ws;
  dec_ws_floats(L, L->Tf, false, &ws, &c);
  launch_params_forward(L, D, d_u, ws.beta, ws.scal, st);
  launch_offsets(L, D, ws.beta, ws.off, L->T, L->T + L->Tf, 0, st);
  const int ncomp = 3 + (L->S2 > 1 ? 1 : 0);
  if (ci_reg_path(L)) {
    CI_DISPATCH_S(L->S, (launch_fmom<S_>(L, d_fstate, ws, st)));
  } else {
    if (int e = launch_forecast_moments_big(L, d_fstate, ws.scal, ws.cmean, ws.cvar, st)) return e;
  }
  const long tot = (long)ncomp * L->N * L->Tf;
  k_mix_components<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(L->N, L->G, L->Tf, L->k > 0, ncomp, ws.cmean, ws.cvar, ws.off,
                                                                  D->d_sconst, d_comp_mean, d_comp_sd);
  CI_CHECK_LAUNCH();
  return CI_OK;
}

}  // extern "C"
