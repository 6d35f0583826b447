// K4 / K5: the two fit drivers.  Both keep every per-lane quantity on the device for the whole
// fit: the host only enqueues launches on the caller's stream (no synchronisation, no copies), so a
// draw never leaves HBM until the caller reads the result arrays.
//
//   ci_vi_fit / ci_vi_draw  mean-field ADVI, Keras-legacy Adam            SURVEY A.7
//       replaces tfp.sts.build_factored_surrogate_posterior + tfp.vi.fit_surrogate_posterior
//       /root/reference/causalimpact/model.py:367-386 (and the 150 x 5 initialisation inside
//       tfp.sts.fit_with_hmc, model.py:361)
//   ci_hmc_run              HMC + dual-averaging step-size adaptation      SURVEY A.6
//       replaces tfp.sts.fit_with_hmc -> mcmc.sample_chain(DualAveragingStepSizeAdaptation(
//       TransformedTransitionKernel(HamiltonianMonteCarlo))), model.py:361-364
//
// All per-lane arrays are [P][B] with the lane the fastest axis; one thread owns one lane (or one
// (parameter, unit) cell for the Adam update) and walks the parameter axis with coalesced accesses.
// Random numbers are counter-based Philox4x32-10 keyed by (lane, parameter, iteration, stream tag),
// so the oracle (oracle/sts_oracle.py::vi_fit / hmc_run) reproduces the same streams.
#include "ci_abi_common.cuh"

namespace ci {

// ------------------------------------------------------------------------------------------------
// VI
// ------------------------------------------------------------------------------------------------
struct ViWs {
  float* u;     // [P][B]  reparameterised draws
  float* xi;    // [P][B]  standard normals behind them
  float* grad;  // [P][B]
  float* logp;  // [B]
  float* adam;  // [4][P][U]  m_mu, v_mu, m_rho, v_rho
  EvalWs ev;
};

static size_t vi_ws_floats(const ci_layout* L, int units_per_series, ViWs* ws, Carver* c) {
  const size_t B = (size_t)L->N * L->G, U = (size_t)L->N * units_per_series;
  const size_t P = layout_num_params(L);
  size_t tot = 3 * align_up(P * B * 4, 256) + align_up(B * 4, 256) + align_up(4 * P * U * 4, 256);
  if (ws && c) {
    ws->u = c->floats(P * B);
    ws->xi = c->floats(P * B);
    ws->grad = c->floats(P * B);
    ws->logp = c->floats(B);
    ws->adam = c->floats(4 * P * U);
  }
  tot += eval_ws_floats(L, ws ? &ws->ev : nullptr, c);
  return tot;
}

__device__ __forceinline__ float softplus_inv(float x) { return x + logf(-expm1f(-x)); }

__global__ void k_vi_init(long U, int P, uint64_t seed, float init_scale, float* __restrict__ mu,
                          float* __restrict__ rho, float* __restrict__ adam) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= U * P) return;
  const long unit = i % U;
  const int p = (int)(i / U);
  const f32x4 un = philox_uniform4((uint32_t)unit, (uint32_t)p, 0u, TAG_VI_INIT, seed);
  mu[i] = -2.0f + 4.0f * un.x;
  rho[i] = softplus_inv(init_scale);
  adam[i] = 0.f;
  adam[U * P + i] = 0.f;
  adam[2 * U * P + i] = 0.f;
  adam[3 * U * P + i] = 0.f;
}

// u[p][b] = mu[p][unit] + softplus(rho[p][unit]) * xi,   lane b = unit * sample_size + sample
__global__ void k_vi_sample(long B, long U, int P, int sample_size, int step, uint64_t seed,
                            const float* __restrict__ mu, const float* __restrict__ rho,
                            float* __restrict__ u, float* __restrict__ xi) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * P) return;
  const long b = i % B;
  const int p = (int)(i / B);
  const long unit = b / sample_size;
  const f32x4 z = philox_normal4((uint32_t)b, (uint32_t)p, (uint32_t)step, TAG_VI_EPS, seed);
  xi[i] = z.x;
  u[i] = fmaf(softplus(rho[(long)p * U + unit]), z.x, mu[(long)p * U + unit]);
}

// Reparameterisation gradient of the negative ELBO + Keras-legacy Adam, one (parameter, unit) cell
// per thread:  d/dmu = -mean_s g ;  d/drho = -(mean_s g xi) sigmoid(rho) - sigmoid(rho)/softplus(rho).
__global__ void k_vi_adam(long B, long U, int P, int sample_size, int step, float lr,
                          const float* __restrict__ grad, const float* __restrict__ xi,
                          float* __restrict__ mu, float* __restrict__ rho, float* __restrict__ adam) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= U * P) return;
  const long unit = i % U;
  const int p = (int)(i / U);
  float sg = 0.f, sgx = 0.f;
  for (int s = 0; s < sample_size; ++s) {
    const long b = unit * sample_size + s;
    const float g = grad[(long)p * B + b];
    sg += g;
    sgx = fmaf(g, xi[(long)p * B + b], sgx);
  }
  const float inv = 1.0f / (float)sample_size;
  const float r = rho[i];
  const float dsig = sigmoid(r), sigma = softplus(r);
  const float g_mu = -sg * inv;
  const float g_rho = (-sgx * inv) * dsig - dsig / sigma;
  const float b1 = 0.9f, b2 = 0.999f, eps = 1e-7f;
  const float t = (float)(step + 1);
  const float lr_t = lr * sqrtf(1.0f - powf(b2, t)) / (1.0f - powf(b1, t));
  float* m_mu = adam + i;
  float* v_mu = adam + U * P + i;
  float* m_rho = adam + 2 * U * P + i;
  float* v_rho = adam + 3 * U * P + i;
  const float mm = b1 * (*m_mu) + (1.f - b1) * g_mu;
  const float vm = b2 * (*v_mu) + (1.f - b2) * g_mu * g_mu;
  const float mr = b1 * (*m_rho) + (1.f - b1) * g_rho;
  const float vr = b2 * (*v_rho) + (1.f - b2) * g_rho * g_rho;
  *m_mu = mm; *v_mu = vm; *m_rho = mr; *v_rho = vr;
  mu[i] -= lr_t * mm / (sqrtf(vm) + eps);
  rho[i] = r - lr_t * mr / (sqrtf(vr) + eps);
}

// Negative-ELBO estimate per unit (diagnostic only): mean_s [ log q(u_s) - log target(u_s) ].
__global__ void k_vi_loss(long B, long U, int P, int sample_size, const float* __restrict__ rho,
                          const float* __restrict__ xi, const float* __restrict__ logp,
                          float* __restrict__ loss) {
  const long unit = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (unit >= U) return;
  float acc = 0.f;
  for (int p = 0; p < P; ++p) {
    const float ls = logf(softplus(rho[(long)p * U + unit]));
    for (int s = 0; s < sample_size; ++s) {
      const float z = xi[(long)p * B + unit * sample_size + s];
      acc += -ls - 0.5f * z * z - 0.5f * CI_LOG_2PI;
    }
  }
  for (int s = 0; s < sample_size; ++s) acc -= logp[unit * sample_size + s];
  loss[unit] = acc / (float)sample_size;
}

__global__ void k_vi_draw(long U, int P, int num_draws, uint64_t seed, const float* __restrict__ mu,
                          const float* __restrict__ rho, float* __restrict__ out) {
  const long B = U * num_draws;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * P) return;
  const long b = i % B;
  const int p = (int)(i / B);
  const long unit = b / num_draws;
  const f32x4 z = philox_normal4((uint32_t)b, (uint32_t)p, 0u, TAG_VI_DRAW, seed);
  out[i] = fmaf(softplus(rho[(long)p * U + unit]), z.x, mu[(long)p * U + unit]);
}

// ------------------------------------------------------------------------------------------------
// HMC
// ------------------------------------------------------------------------------------------------
struct HmcWs {
  float* uu;    // [P][B] proposal position
  float* pp;    // [P][B] momentum
  float* gg;    // [P][B] gradient at the proposal
  float* g;     // [P][B] gradient at the current state
  float* lane;  // [8][B] lp, lp_new, h0, log_fac, log_fac_avg, err_sum, n_acc_warm, n_acc_samp
  EvalWs ev;
};
enum { HL_LP = 0, HL_LPNEW, HL_H0, HL_LOGFAC, HL_LOGFAC_AVG, HL_ERRSUM, HL_ACC_WARM, HL_ACC_SAMP, HL_N };

static size_t hmc_ws_floats(const ci_layout* L, HmcWs* ws, Carver* c) {
  const size_t B = (size_t)L->N * L->G;
  const size_t P = layout_num_params(L);
  size_t tot = 4 * align_up(P * B * 4, 256) + align_up(HL_N * B * 4, 256);
  if (ws && c) {
    ws->uu = c->floats(P * B);
    ws->pp = c->floats(P * B);
    ws->gg = c->floats(P * B);
    ws->g = c->floats(P * B);
    ws->lane = c->floats(HL_N * B);
  }
  tot += eval_ws_floats(L, ws ? &ws->ev : nullptr, c);
  return tot;
}

__global__ void k_hmc_reset(long B, float* __restrict__ lane) {
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  lane[HL_LOGFAC * B + b] = 0.f;
  lane[HL_LOGFAC_AVG * B + b] = 0.f;
  lane[HL_ERRSUM * B + b] = 0.f;
  lane[HL_ACC_WARM * B + b] = 0.f;
  lane[HL_ACC_SAMP * B + b] = 0.f;
}

// One kernel per transition boundary: [end of transition `it`] + [begin of transition `it + 1`].
//   end   : kinetic energy of the proposal, Metropolis accept, state selection, dual averaging, draw output
//   begin : momentum draw, initial Hamiltonian, first half kick and first drift of the next transition
// Block = 32 lanes x HMC_PG parameter groups (parameter p is handled by group p % HMC_PG), so every
// access is a coalesced 128-byte row of the [P][B] arrays and a lane's P parameters are spread over
// HMC_PG threads; the per-lane sums are reduced through shared memory in a fixed order (deterministic).
// Fusing the two halves re-uses the selected (u, g) from registers instead of re-reading them.
constexpr int HMC_PG = 8;

struct HmcTurn {
  int do_end, do_begin;
  int it, num_warmup, num_adapt;    // transition being closed (end half)
  uint32_t it_rng_end, it_rng_begin;
  float target_accept;
  uint64_t seed;
};

__global__ void __launch_bounds__(32 * HMC_PG)
k_hmc_turn(long B, int P, HmcTurn a, const float* __restrict__ step0, float* __restrict__ u, float* __restrict__ g,
           float* __restrict__ uu, float* __restrict__ gg, float* __restrict__ pp, float* __restrict__ lane,
           float* __restrict__ draw) {
  __shared__ float red[HMC_PG][32];
  __shared__ float bc[2][32];    // per-lane broadcast: accept flag, step factor
  const int x = threadIdx.x, y = threadIdx.y;
  const long b = (long)blockIdx.x * 32 + x;
  const bool live = b < B;
  bool acc = false;
  if (a.do_end) {
    float ke = 0.f;
    if (live)
      for (int p = y; p < P; p += HMC_PG) {
        const float m = pp[(long)p * B + b];
        ke = fmaf(m, m, ke);
      }
    red[y][x] = ke;
    __syncthreads();
    if (y == 0 && live) {
      float kes = 0.f;
#pragma unroll
      for (int q = 0; q < HMC_PG; ++q) kes += red[q][x];
      const float lp_new = lane[HL_LPNEW * B + b];
      const float h1 = -lp_new + 0.5f * kes;
      float log_acc = lane[HL_H0 * B + b] - h1;
      if (!isfinite(log_acc)) log_acc = -INFINITY;
      const float un = philox_uniform4((uint32_t)b, 0u, a.it_rng_end, TAG_HMC_ACC, a.seed).x;
      const bool ac = logf(un) < log_acc;
      if (ac) lane[HL_LP * B + b] = lp_new;
      if (a.it < a.num_warmup) lane[HL_ACC_WARM * B + b] += ac ? 1.f : 0.f;
      else lane[HL_ACC_SAMP * B + b] += ac ? 1.f : 0.f;
      if (a.it < a.num_adapt) {
        // tfp.mcmc.DualAveragingStepSizeAdaptation defaults: exploration_shrinkage 0.05,
        // step_count_smoothing 10, decay_rate 0.75, shrinkage target log(10 eps0).
        const float p_acc = expf(fminf(log_acc, 0.f));
        const float err = lane[HL_ERRSUM * B + b] + a.target_accept - p_acc;
        lane[HL_ERRSUM * B + b] = err;
        const float t = (float)a.it + 1.0f;
        const float log_x = 2.302585092994046f - err * sqrtf(t) / ((t + 10.0f) * 0.05f);
        const float eta = powf(t, -0.75f);
        const float avg = eta * log_x + (1.0f - eta) * lane[HL_LOGFAC_AVG * B + b];
        lane[HL_LOGFAC_AVG * B + b] = avg;
        lane[HL_LOGFAC * B + b] = (a.it < a.num_adapt - 1) ? log_x : avg;
      }
      bc[0][x] = ac ? 1.f : 0.f;
    }
    __syncthreads();
    acc = live && bc[0][x] != 0.f;
  }
  if (y == 0 && live && a.do_begin) bc[1][x] = expf(lane[HL_LOGFAC * B + b]);
  __syncthreads();
  const float fac = (a.do_begin && live) ? bc[1][x] : 1.f;
  float ke2 = 0.f;
  if (live)
    for (int p = y; p < P; p += HMC_PG) {
      const long i = (long)p * B + b;
      float un, gn;
      if (acc) {
        un = uu[i];
        gn = gg[i];
        u[i] = un;
        g[i] = gn;
      } else {
        un = u[i];
        gn = g[i];
      }
      if (a.do_end && draw) draw[i] = un;
      if (a.do_begin) {
        const float mom = philox_normal4((uint32_t)b, (uint32_t)p, a.it_rng_begin, TAG_HMC_MOM, a.seed).x;
        ke2 = fmaf(mom, mom, ke2);
        const float eps = step0[i] * fac;
        const float ph = fmaf(0.5f * eps, gn, mom);
        pp[i] = ph;
        uu[i] = fmaf(eps, ph, un);
      }
    }
  if (a.do_begin) {
    __syncthreads();          // red[] is free again (its readers passed the barriers above)
    red[y][x] = ke2;
    __syncthreads();
    if (y == 0 && live) {
      float kes = 0.f;
#pragma unroll
      for (int q = 0; q < HMC_PG; ++q) kes += red[q][x];
      lane[HL_H0 * B + b] = -lane[HL_LP * B + b] + 0.5f * kes;
    }
  }
}

__global__ void k_hmc_stats(long B, int num_warmup, int num_results, const float* __restrict__ lane,
                            float* __restrict__ stats) {
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  stats[0 * B + b] = num_results > 0 ? lane[HL_ACC_SAMP * B + b] / (float)num_results : 0.f;
  stats[1 * B + b] = num_warmup > 0 ? lane[HL_ACC_WARM * B + b] / (float)num_warmup : 0.f;
  stats[2 * B + b] = expf(lane[HL_LOGFAC * B + b]);
  stats[3 * B + b] = lane[HL_LP * B + b];
}

}  // namespace ci

using namespace ci;

static inline unsigned nblk(long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

extern "C" {

size_t ci_vi_workspace_bytes(const ci_layout* L, int32_t units_per_series) {
  if (check_layout(L) || units_per_series < 1) return 0;
  return vi_ws_floats(L, units_per_series, nullptr, nullptr);
}

int ci_vi_fit(const ci_layout* L, const ci_data* D, int32_t units_per_series, int32_t sample_size,
              int32_t num_steps, float learning_rate, float init_scale, uint64_t seed, float* d_mu,
              float* d_rho, float* d_loss, void* d_workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_layout(L)) return e;
  CI_CHECK_ARG(D && D->d_y && D->d_sconst && D->d_X, "NULL data pointer");
  CI_CHECK_ARG(d_mu && d_rho && d_workspace, "NULL pointer");
  CI_CHECK_ARG(units_per_series >= 1 && sample_size >= 1 && num_steps >= 0, "non-positive count");
  CI_CHECK_ARG(L->G == units_per_series * sample_size, "layout G must equal units_per_series * sample_size");
  CI_CHECK_ARG(learning_rate > 0.f && init_scale > 0.f, "learning_rate and init_scale must be positive");
  if (!ci_reg_path(L) && !big_supported(L))
    return fail(CI_ERR_UNSUPPORTED, "nseasons=%d not supported by this build", L->S);
  if (workspace_bytes < ci_vi_workspace_bytes(L, units_per_series))
    return fail(CI_ERR_WORKSPACE_TOO_SMALL, "workspace needs %zu bytes", ci_vi_workspace_bytes(L, units_per_series));
  cudaStream_t st = (cudaStream_t)stream;
  Carver c(d_workspace);
  ViWs ws;
  vi_ws_floats(L, units_per_series, &ws, &c);
  const long B = (long)L->N * L->G, U = (long)L->N * units_per_series;
  const int P = layout_num_params(L);
  k_vi_init<<<nblk(U * P, 256), 256, 0, st>>>(U, P, seed, init_scale, d_mu, d_rho, ws.adam);
  for (int step = 0; step < num_steps; ++step) {
    k_vi_sample<<<nblk(B * P, 256), 256, 0, st>>>(B, U, P, sample_size, step, seed, d_mu, d_rho, ws.u, ws.xi);
    if (int e = launch_eval(L, D, ws.u, ws.logp, ws.grad, ws.ev, st)) return e;
    if (d_loss)
      k_vi_loss<<<nblk(U, 128), 128, 0, st>>>(B, U, P, sample_size, d_rho, ws.xi, ws.logp, d_loss + (long)step * U);
    k_vi_adam<<<nblk(U * P, 256), 256, 0, st>>>(B, U, P, sample_size, step, learning_rate, ws.grad, ws.xi, d_mu,
                                                d_rho, ws.adam);
  }
  CI_CHECK_LAUNCH();
  return CI_OK;
}

int ci_vi_draw(int32_t P, int64_t U, int32_t num_draws, uint64_t seed, const float* d_mu, const float* d_rho,
               float* d_u_out, void* stream) {
  CI_CHECK_ARG(P >= 1 && U >= 1 && num_draws >= 1, "non-positive count");
  CI_CHECK_ARG(d_mu && d_rho && d_u_out, "NULL pointer");
  k_vi_draw<<<nblk((long)U * num_draws * P, 256), 256, 0, (cudaStream_t)stream>>>((long)U, P, num_draws, seed, d_mu,
                                                                                 d_rho, d_u_out);
  CI_CHECK_LAUNCH();
  return CI_OK;
}

size_t ci_hmc_workspace_bytes(const ci_layout* L) {
  if (check_layout(L)) return 0;
  return hmc_ws_floats(L, nullptr, nullptr);
}

int ci_hmc_run(const ci_layout* L, const ci_data* D, float* d_u, const float* d_step0, int32_t num_results,
               int32_t num_warmup, int32_t num_adapt, int32_t num_leapfrog, float target_accept, uint64_t seed,
               int32_t iter_offset, float* d_draws, float* d_stats, void* d_workspace, size_t workspace_bytes,
               void* stream) {
  if (int e = check_layout(L)) return e;
  CI_CHECK_ARG(D && D->d_y && D->d_sconst && D->d_X, "NULL data pointer");
  CI_CHECK_ARG(d_u && d_step0 && d_workspace, "NULL pointer");
  CI_CHECK_ARG(num_results >= 0 && num_warmup >= 0 && num_leapfrog >= 1, "bad iteration counts");
  CI_CHECK_ARG(num_adapt >= 0 && num_adapt <= num_warmup, "num_adapt must lie in [0, num_warmup]");
  CI_CHECK_ARG(num_results == 0 || d_draws, "d_draws is NULL");
  CI_CHECK_ARG(target_accept > 0.f && target_accept < 1.f, "target_accept must lie in (0, 1)");
  if (!ci_reg_path(L) && !big_supported(L))
    return fail(CI_ERR_UNSUPPORTED, "nseasons=%d not supported by this build", L->S);
  if (workspace_bytes < ci_hmc_workspace_bytes(L))
    return fail(CI_ERR_WORKSPACE_TOO_SMALL, "workspace needs %zu bytes", ci_hmc_workspace_bytes(L));
  cudaStream_t st = (cudaStream_t)stream;
  Carver c(d_workspace);
  HmcWs ws;
  hmc_ws_floats(L, &ws, &c);
  const long B = (long)L->N * L->G;
  const int P = layout_num_params(L);
  const unsigned gl = nblk(B, 128);
  k_hmc_reset<<<gl, 128, 0, st>>>(B, ws.lane);
  // log-prob and gradient at the initial state (cached between transitions, as TFP does)
  if (int e = launch_eval(L, D, d_u, ws.lane + HL_LP * B, ws.g, ws.ev, st)) return e;
  const int n_it = num_warmup + num_results;
  const dim3 tb(32, HMC_PG), tg((unsigned)((B + 31) / 32));
  for (int it = 0; it <= n_it; ++it) {
    // boundary kernel: closes transition it - 1 (accept, adaptation, draw) and opens transition it
    HmcTurn a;
    a.do_end = it > 0;
    a.do_begin = it < n_it;
    a.it = it - 1;
    a.num_warmup = num_warmup;
    a.num_adapt = num_adapt;
    a.it_rng_end = (uint32_t)(iter_offset + it - 1);
    a.it_rng_begin = (uint32_t)(iter_offset + it);
    a.target_accept = target_accept;
    a.seed = seed;
    float* draw = (it > 0 && it - 1 >= num_warmup) ? d_draws + (size_t)(it - 1 - num_warmup) * P * B : nullptr;
    k_hmc_turn<<<tg, tb, 0, st>>>(B, P, a, d_step0, d_u, ws.g, ws.uu, ws.gg, ws.pp, ws.lane, draw);
    if (it == n_it) break;
    for (int l = 0; l < num_leapfrog; ++l) {
      // kick (and, unless last, drift) fused into the evaluation kernel's epilogue
      LeapArgs lf;
      lf.pp = ws.pp;
      lf.uu = ws.uu;
      lf.step0 = d_step0;
      lf.logfac = ws.lane + HL_LOGFAC * B;
      lf.last = (l == num_leapfrog - 1);
      if (int e = launch_eval(L, D, ws.uu, ws.lane + HL_LPNEW * B, ws.gg, ws.ev, st, &lf)) return e;
    }
  }
  if (d_stats) k_hmc_stats<<<gl, 128, 0, st>>>(B, num_warmup, num_results, ws.lane, d_stats);
  CI_CHECK_LAUNCH();
  return CI_OK;
}

}  // extern "C"
