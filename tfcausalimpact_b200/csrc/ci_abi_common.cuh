// Host-side plumbing shared by the C-ABI translation units: error string, argument checks,
// workspace carving.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include "../../include/ci_b200.h"
#include "ci_core.cuh"
#include "ci_dispatch.cuh"
#include "ci_params.cuh"
#include "ci_params_gen.cuh"

namespace ci {

char* last_error_buf();  // thread-local, defined in ci_eval.cu

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define CI_CHECK_ARG(cond, msg) \
  do { if (!(cond)) return ci::fail(CI_ERR_INVALID_ARGUMENT, "%s: %s", __func__, msg); } while (0)

#define CI_CHECK_LAUNCH()                                                                   \
  do {                                                                                      \
    cudaError_t e_ = cudaPeekAtLastError();                                                 \
    if (e_ != cudaSuccess) return ci::fail(CI_ERR_CUDA, "%s: %s", __func__, cudaGetErrorString(e_)); \
  } while (0)

inline int check_layout(const ci_layout* L) {
  if (!L) return fail(CI_ERR_INVALID_ARGUMENT, "layout is NULL");
  if (L->N < 1 || L->G < 1 || L->T < 1 || L->Tf < 0 || L->k < 0 || L->S < 1 || L->r < 1)
    return fail(CI_ERR_INVALID_ARGUMENT, "layout has a non-positive dimension");
  if (L->S == 1 && L->r != 1) return fail(CI_ERR_INVALID_ARGUMENT, "season_duration > 1 needs nseasons > 1");
  if (L->flags & ~15) return fail(CI_ERR_INVALID_ARGUMENT, "unknown layout flags");
  if (L->S2 < 0 || L->r2 < 0) return fail(CI_ERR_INVALID_ARGUMENT, "layout has a negative second-seasonal dimension");
  if (L->S2 > 1 && (L->S <= 1 || L->r2 < 1))
    return fail(CI_ERR_INVALID_ARGUMENT, "a second seasonal component needs nseasons > 1 and season_duration2 >= 1");
  if (L->Tpad % 4 != 0 || L->Tpad < L->T + L->Tf)
    return fail(CI_ERR_INVALID_ARGUMENT, "Tpad must be a multiple of 4 and >= T + Tf");
  return CI_OK;
}

// default model on the register-resident kernels; everything else on the warp-per-lane path
inline bool ci_two_seasonal(const ci_layout* L) { return L->S2 > 1; }
inline bool ci_gen_params(const ci_layout* L) { return L->flags != 0 || L->S2 > 1; }
inline bool ci_reg_path(const ci_layout* L) { return !ci_gen_params(L) && ci_reg_set(L->S); }
inline int layout_num_params(const ci_layout* L) {
  return ci_gen_params(L) ? num_params_gen(L->k, L->S, L->flags, L->S2) : num_params(L->k, L->S);
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace (256-byte aligned slices).
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(reinterpret_cast<char*>(p)) {}
  float* floats(size_t n) {
    float* r = reinterpret_cast<float*>(base + off);
    off += align_up(n * sizeof(float), 256);
    return r;
  }
};

// ---- large-latent (warp-per-lane) path, ci_bigd.cu ---------------------------------------------
bool big_supported(const ci_layout* L);
size_t big_tape_floats(const ci_layout* L);
size_t big_fstate_floats(const ci_layout* L);   // per lane, lane-contiguous [B][D + 2 D^2]
int launch_eval_big(const ci_layout* L, const ci_data* D, const float* d_u, float* d_logp, float* d_grad,
                    float* tape, cudaStream_t stream);
int launch_filter_big(const ci_layout* L, const ci_data* D, const float* d_u, float* pm, float* pv, float* fs,
                      cudaStream_t stream);
size_t big_hist_floats(const ci_layout* L);     // smoother history [B][T][3 D + 5]
int launch_smooth_big(const ci_layout* L, const ci_data* D, const float* d_u, float* hist, float* cmean, float* cvar,
                      cudaStream_t stream);
int launch_forecast_moments_big(const ci_layout* L, const float* fs, const float* scal, float* cmean, float* cvar,
                                cudaStream_t stream);
void launch_forecast_big(const ci_layout* L, const float* fs, const float* scal, const float* off, float* pm,
                         const float* mu_sig, int nsamp, uint64_t seed, float* sims, cudaStream_t stream);

// ---- few lanes, large latent: 4 warps per lane, covariance in registers, ci_eval_rw.cu ----------------
extern int g_eval_rw, g_rw_warps;
bool rw_eligible(const ci_layout* L);
int launch_eval_rw(const ci_layout* L, const ci_data* D, const float* d_u, float* d_logp, float* d_grad, float* tape,
                   float* rscr, cudaStream_t stream);

extern int g_onestep_tiled;                    // ci_predict.cu tuning hook (ci_set_option)
extern int g_filter_fused;                      // ci_predict.cu tuning hook (ci_set_option)

// ---- lane-cooperative evaluation for sub-wave batches, ci_eval_coop.cu -----------------------------
bool coop_eligible(const ci_layout* L);         // layout + lane count served by the cooperative kernel
size_t coop_tape_floats(const ci_layout* L);    // [warps][T][32], lanes padded to 8
extern int g_coop_mma;                          // regression passes of the cooperative kernel on the tensor cores (ci_set_option)
void coop_set_option(int tpl, long max_lanes);  // tuning hook behind ci_set_option
struct LeapArgs;
int launch_eval_coop(const ci_layout* L, const ci_data* D, const float* d_u, float* d_logp, float* d_grad,
                     float* tape, cudaStream_t stream, const LeapArgs& lf);

// ---- warp-specialised thread-per-lane evaluation for mid-size batches, ci_eval_ws.cu ----------------------------
extern int g_eval_ws;                           // ci_set_option("eval_ws", 0 / 1)
extern long g_ws_max_lanes;                     // ci_set_option("eval_ws_max_lanes", lanes; -1 = default)
bool ws_eligible(const ci_layout* L);
int launch_eval_ws(const ci_layout* L, const ci_data* D, const float* d_u, float* d_logp, float* d_grad, float* tape,
                   cudaStream_t stream, const LeapArgs& lf);

// Per-evaluation scratch: only the filter tape leaves the evaluation kernel (see ci_eval.cu).
struct EvalWs {
  float* tape;   // [T][S+1][B] (g_0..g_{S-1}, v), or warp-blocked [T][B/32][S+1][32] (effects path)
};
inline size_t eval_ws_floats(const ci_layout* L, EvalWs* ws, Carver* c) {
  const size_t B = ((size_t)L->N * L->G + 31) / 32 * 32;   // whole warps (warp-blocked tape layout)
  size_t nt = ci_reg_path(L) ? (size_t)L->T * (L->S + 1) * B : big_tape_floats(L);
  if (ci_reg_path(L) && L->r == 1 && L->S >= 2 && L->S <= 7 && coop_tape_floats(L) > nt) nt = coop_tape_floats(L);
  if (ws && c) ws->tape = c->floats(nt);
  return align_up(nt * 4, 256);
}

// Optional leapfrog epilogue fused into the evaluation kernel (HMC): with the fresh gradient,
//   p += (last ? 1/2 : 1) eps g ;  if (!last) u += eps p ,   eps = step0 * exp(logfac[lane]).
struct LeapArgs {
  float* pp = nullptr;            // [P][B] momentum (nullptr = no epilogue)
  float* uu = nullptr;            // [P][B] position (the array being evaluated)
  const float* step0 = nullptr;   // [P][B]
  const float* logfac = nullptr;  // [B]
  int last = 0;
};

// Enqueue one joint log-prob + gradient evaluation (one kernel) on `stream`.
int launch_eval(const ci_layout* L, const ci_data* D, const float* d_u, float* d_logp, float* d_grad,
                const EvalWs& ws, cudaStream_t stream, const LeapArgs* leap = nullptr);

// out[t - t_lo][b] = with_y ? y_t - offset_t : offset_t, offset_t = mean + X_t . beta  (ci_eval.cu)
void launch_offsets(const ci_layout* L, const ci_data* D, const float* beta, float* out, int t_lo, int t_hi,
                    int with_y, cudaStream_t stream);
// u -> beta [k][B], scal rows 0..3 = R, ql, qd, prior + log-det  (ci_eval.cu)
void launch_params_forward(const ci_layout* L, const ci_data* D, const float* d_u, float* beta, float* scal,
                           cudaStream_t stream);

}  // namespace ci
