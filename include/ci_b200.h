/* ci_b200.h -- C ABI of the B200-native tfcausalimpact fit / forecast / reduction hot path.
 *
 * The reference (WillianFuks/tfcausalimpact v0.0.17) is pure Python and has NO native/FFI boundary
 * of its own: its hot path is a dozen calls into the un-vendored tensorflow-probability package
 * (causalimpact/model.py:216,235,291,306,311,319,361,370,374,414,446; inferences.py:108-145,397).
 * Every entry point below therefore cites the reference *call site* whose work it replaces; the
 * Python stub a maintainer would add on the reference side is shown in INTEGRATION.md.
 *
 * Conventions
 *  - every function returns 0 on success, a negative CI_ERR_* otherwise, and never throws;
 *    ci_last_error_string() describes the last failure on the calling thread.
 *  - every `d_*` pointer is a caller-owned DEVICE pointer to contiguous float32 (unless noted);
 *    nothing is allocated or freed inside (except the *_host convenience entry points, which own
 *    their staging buffers), `stream` is a cudaStream_t passed as void*.  Calls are asynchronous
 *    with respect to the host and re-entrant per stream.
 *  - "lane" = one (series, chain[, sample]) evaluation; lanes are the FASTEST axis of every
 *    per-lane array ([P][B] parameters, [T][B] residuals ...), lane b belongs to series b / G.
 *  - arithmetic is float32, as the reference forces (causalimpact/data.py:274-275).
 */
#ifndef CI_B200_H_
#define CI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CI_OK 0
#define CI_ERR_INVALID_ARGUMENT (-1)
#define CI_ERR_WORKSPACE_TOO_SMALL (-2)
#define CI_ERR_UNSUPPORTED (-3)   /* layout outside the compiled set (e.g. nseasons not supported) */
#define CI_ERR_CUDA (-4)          /* a CUDA runtime call failed; see ci_last_error_string() */

#define CI_SCONST_FIELDS 8

/* Fixed state-space layout of Sum([LocalLevel | LocalLinearTrend, SparseLinearRegression | LinearRegression ?,
 * Seasonal?, Seasonal?]): flags == 0 and S2 <= 1 is the default model built by causalimpact/model.py:239-321
 * (`build_default_model`); `flags`, `S2`, `r2` describe the custom `model=` set (main.py:186-252 examples). */
typedef struct ci_layout {
  int32_t N;    /* series in this call */
  int32_t G;    /* lanes per series (chains x Monte-Carlo samples); B = N * G */
  int32_t T;    /* observed (pre-period) steps */
  int32_t Tf;   /* forecast (post-period) steps */
  int32_t k;    /* covariates (0 = no regression component; model.py:298) */
  int32_t S;    /* nseasons (1 = no seasonal component; model.py:310) */
  int32_t r;    /* season_duration (model.py:313) */
  int32_t Tpad; /* padded length of the packed design matrix: multiple of 4, >= T + Tf */
  int32_t flags; /* CI_FLAG_*; 0 = the default model of causalimpact/model.py:239-321 */
  int32_t S2;   /* num_seasons of a SECOND tfp.sts.Seasonal component (0 or 1 = none); custom `model=` only
                   (notebooks/getting_started.ipynb cell 116: Seasonal(4) + Seasonal(52)); needs S > 1 */
  int32_t r2;   /* its num_steps_per_season */
} ci_layout;

/* Wider model set for custom `model=` (causalimpact/main.py:186-252 examples).
 * Layouts with flags != 0 run on the large-latent (warp-per-lane) kernels. */
#define CI_FLAG_DENSE_REG 1        /* tfp.sts.LinearRegression (StudentT(5,0,10) weights) instead of Sparse */
#define CI_FLAG_LLT 2              /* tfp.sts.LocalLinearTrend (level + slope) instead of LocalLevel       */
#define CI_FLAG_OBS_LOGNORMAL 4    /* observation noise prior LogNormal(log(.01 sd), 2) (tfp.sts.Sum default) */
#define CI_FLAG_LEVEL_LOGNORMAL 8  /* level scale prior LogNormal(log(.05 sd), 3) (tfp.sts.LocalLevel default) */

/* Caller-owned device inputs shared by every call on one batch of series. */
typedef struct ci_data {
  const float* d_y;      /* [N][T]        observed series, NaN = missing step            */
  const float* d_X;      /* [Tpad/4][N][k+1][4] chunk-major covariates + y (ci_pack_design) */
  const float* d_sconst; /* [8][N]        per-series constants (see ci_series_stats)     */
  const float* d_XT;     /* optional (may be NULL): the same design block covariate-major inside a chunk,
                            [Tpad/4][k+1][N][4] (ci_repack_design_cov).  Used by ci_kalman_logp_grad / the fit
                            drivers when a series has fewer than 4 lanes (e.g. 1 chain, the reference's
                            default): every lane then stages its own series' chunk with cp.async and the 32
                            lanes of a warp read one contiguous run per covariate.  */
} ci_data;

int ci_version(void);
const char* ci_last_error_string(void);

/* Number of scalar parameters of the layout: 2 + (k ? 2 + 3k : 0) + (S > 1).
 * Order = tfp `model.parameters` (notebooks/getting_started.ipynb:2228-2236). */
int ci_num_params(int32_t k, int32_t S);
/* ... of a layout with flags (trend 1|2 scales, regression 2+3k | k weights). */
int ci_num_params_layout(const ci_layout* L);

/* Empirical statistics + prior hyper-parameters per series.
 * Replaces: tfp sts_util.empirical_statistics as used by LocalLevel/Seasonal/Sum, and
 * causalimpact/model.py:283-290 (obs_sd, InverseGamma(16, 16 sigma^2) priors). */
int ci_series_stats(const ci_layout* L, const float* d_y, float prior_level_sd, float* d_sconst,
                    void* stream);

/* Row-major covariates [N][T+Tf][k] and the series [N][T] -> chunk-major [Tpad/4][N][k+1][4]:
 * 4 consecutive steps of one covariate are 16 contiguous bytes, the k+1 rows of a series follow each
 * other and consecutive series are adjacent (one TMA bulk copy per warp and chunk); row k holds the
 * centred series y - mean(y) (NaN -> 0 in covariates, zero padded; NaN in y = missing), so
 * d_sconst must come from ci_series_stats on the same d_y.  Required for every layout,
 * also k = 0.  d_X_rowmajor may be NULL when k = 0.
 * Replaces: causalimpact/model.py:303-308 (concat(pre, post).astype(float32).fillna(0)). */
int ci_pack_design(const ci_layout* L, const float* d_X_rowmajor, const float* d_y, const float* d_sconst,
                   float* d_XY_packed, void* stream);

/* 1 when ci_kalman_logp_grad / the fit drivers would use ci_data.d_XT for this layout (fewer than 4 lanes per
 * series and the per-lane staging slices fit in shared memory), else 0: callers build the copy only then. */
int ci_eval_uses_xt(const ci_layout* L);
/* Covariate-major copy of the packed block for ci_data.d_XT: XT[tc][j][n][q] = XY[tc][n][j][q]. */
int ci_repack_design_cov(const ci_layout* L, const float* d_XY_packed, float* d_XT_packed, void* stream);

/* ---- K1-K3: joint log-prob + gradient in unconstrained space ------------------------------
 * Replaces: `model.joint_log_prob(observed_time_series)` + TF autodiff, causalimpact/model.py:375-377
 * (VI) and the target_log_prob_fn built inside tfp.sts.fit_with_hmc, model.py:361-364.
 *   d_u    [P][B] in   unconstrained parameters
 *   d_logp [B]    out  log p(y|theta) + log p(theta) + log|dtheta/du|
 *   d_grad [P][B] out  gradient w.r.t. u                                                    */
size_t ci_eval_workspace_bytes(const ci_layout* L);
int ci_kalman_logp_grad(const ci_layout* L, const ci_data* D, const float* d_u, float* d_logp,
                        float* d_grad, void* d_workspace, size_t workspace_bytes, void* stream);

/* Tuning hook (process-wide, not part of the reference-facing contract; used by bench / A-B scripts):
 *   "coop_threads_per_lane"  -1 (default: chosen by lane count), 8, 4, or 0 = never use the lane-cooperative kernel
 *   "coop_max_lanes"         largest N * G served by it (-1 = default: the measured crossover)
 *   "coop_mma"               1 (default) regression passes of the lane-cooperative kernel on the tensor cores (3xTF32), 0 = FFMA2
 *   "eval_ws"                1 (default) mid-size batches (above the cooperative kernel, up to 28,416 lanes) run on the warp-specialised
 *                            kernel (csrc/ci_eval_ws.cu: filter steps and regression in two warps per 32 lanes)
 *   "eval_ws_max_lanes"      largest N * G served by it (-1 = default)
 *   "eval_rw"                1 (default) few-lane / large-latent evaluations run on the register-resident kernel (csrc/ci_eval_rw.cu)
 *   "eval_rw_warps"          4 (default) or 8 warps per lane in that kernel
 *   "filter_fused"           1 (default) ci_filter_predict runs as one fused kernel, 0 = the three-kernel path
 *   "onestep_tiled"          1 (default) ci_onestep_sample stages each series' predictive moments in shared memory
 * The lane-cooperative kernel (csrc/ci_eval_coop.cu) serves default-model layouts with 2 <= nseasons <= 7,
 * season_duration 1, whose lanes cannot fill the GPU at one thread per lane. */
int ci_set_option(const char* name, long value);

/* ---- K5: mean-field ADVI ---------------------------------------------------------------------
 * Replaces: tfp.sts.build_factored_surrogate_posterior + tfp.vi.fit_surrogate_posterior with
 * Adam(0.1), causalimpact/model.py:367-381 (200 steps, 1 sample) and the 150 x 5 initialisation
 * inside fit_with_hmc (model.py:361).  L->G must equal units_per_series * sample_size; lane
 * b = unit * sample_size + sample, unit = series * units_per_series + chain.
 *   d_mu, d_rho [P][U]  out  variational location / softplus^-1(scale), U = N * units_per_series
 *   d_loss      [num_steps][U] out (may be NULL)  negative ELBO estimate per step             */
size_t ci_vi_workspace_bytes(const ci_layout* L, int32_t units_per_series);
int ci_vi_fit(const ci_layout* L, const ci_data* D, int32_t units_per_series, int32_t sample_size,
              int32_t num_steps, float learning_rate, float init_scale, uint64_t seed, float* d_mu,
              float* d_rho, float* d_loss, void* d_workspace, size_t workspace_bytes, void* stream);

/* `surrogate_posterior.sample(num_draws)` (model.py:384) in unconstrained space:
 *   d_u_out [P][U * num_draws], lane = unit * num_draws + draw.                              */
int ci_vi_draw(int32_t P, int64_t U, int32_t num_draws, uint64_t seed, const float* d_mu,
               const float* d_rho, float* d_u_out, void* stream);

/* ---- K4: HMC with dual-averaging step-size adaptation ------------------------------------------
 * Replaces: tfp.sts.fit_with_hmc -> mcmc.sample_chain(DualAveragingStepSizeAdaptation(
 * TransformedTransitionKernel(HamiltonianMonteCarlo))), causalimpact/model.py:361-364.
 * L->G = chains per series.  One lane = one chain.
 *   d_u      [P][B]  in/out  current unconstrained state
 *   d_step0  [P][B]  in      initial per-coordinate step sizes (VI stddevs, as fit_with_hmc)
 *   d_draws  [num_results][P][B] out  unconstrained draws after warm-up
 *   d_stats  [4][B] out  accept rate (sampling phase), accept rate (warm-up), final step factor,
 *                        last log-prob                                                          */
size_t ci_hmc_workspace_bytes(const ci_layout* L);
int ci_hmc_run(const ci_layout* L, const ci_data* D, float* d_u, const float* d_step0,
               int32_t num_results, int32_t num_warmup, int32_t num_adapt, int32_t num_leapfrog,
               float target_accept, uint64_t seed, int32_t iter_offset, float* d_draws,
               float* d_stats, void* d_workspace, size_t workspace_bytes, void* stream);

/* Constrained parameter values theta [P][B] from u [P][B] (what `model_samples` exposes). */
int ci_constrain(const ci_layout* L, const ci_data* D, const float* d_u, float* d_theta, void* stream);

/* ---- K6: filtering over the observed period per posterior draw ---------------------------------
 * Replaces: tfp.sts.one_step_predictive (model.py:414-418) and the forward_filter half of
 * tfp.sts.forecast (model.py:446-451).  L->G = posterior draws per series.
 *   d_pred_mean, d_pred_var [T][B]  out  one-step-ahead predictive mean / variance of y_t
 *   d_fstate [ci_fstate_floats(L) * B] out  predicted latent mean, covariance and its lower Cholesky
 *                                           factor at step T (opaque; input of ci_forecast_sample) */
size_t ci_predict_workspace_bytes(const ci_layout* L);
/* floats per lane of d_fstate (allocate ci_fstate_floats(L) * N * G floats); the internal layout
 * depends on the kernel path and is only consumed by ci_forecast_sample */
size_t ci_fstate_floats(const ci_layout* L);
int ci_filter_predict(const ci_layout* L, const ci_data* D, const float* d_u, float* d_pred_mean,
                      float* d_pred_var, float* d_fstate, void* d_workspace, size_t workspace_bytes,
                      void* stream);

/* ---- K7: posterior-predictive sampling ---------------------------------------------------------
 * one_step_dist.sample(n) / .mean() and posterior_dist.sample(n) / .mean()
 * (causalimpact/inferences.py:108,113,119,134; main.py:377), in ORIGINAL units
 * (y * sigma + mu, misc.py:75-77; pass d_mu_sig = NULL when standardize=False).
 *   d_mu_sig [2][N] (mu, sigma) or NULL
 *   d_pre_sims  [N][nsamp][T]   d_pre_mean  [N][T]
 *   d_post_sims [N][nsamp][Tf]  d_post_mean [N][Tf]                                            */
int ci_onestep_sample(const ci_layout* L, const float* d_pred_mean, const float* d_pred_var,
                      const float* d_mu_sig, int32_t nsamp, uint64_t seed, float* d_pre_sims,
                      float* d_pre_mean, void* stream);
int ci_forecast_sample(const ci_layout* L, const ci_data* D, const float* d_u, const float* d_fstate,
                       const float* d_mu_sig, int32_t nsamp, uint64_t seed, float* d_post_sims,
                       float* d_post_mean, void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- K8: sample reductions -----------------------------------------------------------------------
 * Replaces: compile_posterior_inferences (inferences.py:119-206), summarize_posterior_inferences
 * (inferences.py:311-352) and compute_p_value (inferences.py:397-408).
 *   d_y_pre [N][T], d_y_post [N][Tf] observed series in original units (NaN in y_post = masked
 *   step, dropped from every post-period statistic as `post_data[mask]` does)
 *   d_inferences [N][16][T+Tf+1] out  the 16 columns of inferences.py:229-246; columns of the
 *                                     pre+post kind use entries [0, T+Tf'), cumulative ones
 *                                     [0, Tf'+1) (zero-prepended), Tf' = number of unmasked steps
 *   d_summary    [N][21] out  10x2 summary table row-major (inferences.py:341-352) + p-value     */
size_t ci_reduce_workspace_bytes(const ci_layout* L, int32_t nsamp);
int ci_reduce_inferences(const ci_layout* L, const float* d_y_pre, const float* d_y_post,
                         const float* d_pre_sims, const float* d_pre_mean, const float* d_post_sims,
                         const float* d_post_mean, int32_t nsamp, float alpha, float* d_inferences,
                         void* d_workspace, size_t workspace_bytes, void* stream);
int ci_reduce_summary(const ci_layout* L, const float* d_y_post, const float* d_post_mean,
                      const float* d_post_sims, int32_t nsamp, float alpha, float* d_summary,
                      void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- component decomposition ----------------------------------------------
 * Replaces: tfp.sts.decompose_by_component / decompose_forecast_by_component as the reference's
 * notebook applies them to ci.model / ci.model_samples (notebooks/getting_started.ipynb:125,144).
 * L->G = posterior draws per series.  Outputs are mixtures over the draws, in MODEL units:
 *   d_comp_mean, d_comp_sd  [C][N][W]   components: 0 trend (LocalLevel / LocalLinearTrend level),
 *                                       1 regression (X_t beta), 2 Seasonal, 3 second Seasonal (C = 4 only
 *                                       when L->S2 > 1, else C = 3)
 *   W = T (smoothed pre-period marginals) or Tf (forecast marginals); absent components are zero. */
size_t ci_decompose_workspace_bytes(const ci_layout* L);
int ci_decompose_by_component(const ci_layout* L, const ci_data* D, const float* d_u, float* d_comp_mean,
                              float* d_comp_sd, void* d_workspace, size_t workspace_bytes, void* stream);
int ci_decompose_forecast(const ci_layout* L, const ci_data* D, const float* d_u, const float* d_fstate,
                          float* d_comp_mean, float* d_comp_sd, void* d_workspace, size_t workspace_bytes,
                          void* stream);

/* Kernel launches one ci_kalman_logp_grad call enqueues (1: the evaluation is a single fused kernel). */
int ci_eval_num_kernels(const ci_layout* L);

/* FP32-FMA throughput probe used by bench.py as the measured CUDA-core roofline denominator
 * (MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only): `blocks` CTAs of 256 threads, 16
 * independent FMA chains per thread, `iters` rounds; *flops_out = FLOPs the launch performs. */
int ci_bench_fma(int32_t iters, int32_t blocks, float* d_out, void* stream, double* flops_out);

/* Same probe with the Blackwell FP32x2 instruction (FFMA2): same FLOPs, half the instructions. */
int ci_bench_fma2(int32_t iters, int32_t blocks, float* d_out, void* stream, double* flops_out);

#ifdef __cplusplus
}
#endif
#endif /* CI_B200_H_ */
