// One complete joint log-prob + gradient evaluation for one (series, chain[, sample]) lane:
//   u -> (sigma^2's, beta, prior + log-det)                                     [K3 forward]
//   r_t = y_t - mean - X_t . beta, Kalman filter, tape (g_t, s_t, v_t)           [K2 + K1 forward]
//   reverse-time adjoint sweep -> dL/dr_t -> beta_bar = -sum_t rbar_t X_t        [K1 + K2 reverse]
//   VJP through horseshoe / softplus / sqrt + prior gradient                     [K3 reverse]
// in ONE pass per direction over the series.  Nothing but the tape leaves the thread: beta /
// beta_bar live in a per-lane scratch column (shared memory on the device), the 8 residuals of the
// current time chunk in a second one, mean / covariance / adjoints in registers.
//
// Design matrix layout ("chunk-major"): X[n][tc][j][8], tc = t / 8 -- the k covariates of 8
// consecutive time steps are one contiguous 32*k-byte run, so a lane pulls 8 steps of one covariate
// with two 16-byte loads and a CTA streams whole 128-byte lines.
//
// `__host__ __device__`: the TEST-ONLY host shim (tests/host_shim) runs exactly this function on the
// CPU against the oracle.  Math: SURVEY.md Appendix A.2-A.5.
#pragma once
#include "ci_core.cuh"
#include "ci_kalman.cuh"
#include "ci_params.cuh"

namespace ci {

struct f4 {
  float x, y, z, w;
};

CI_HD f4 load4(const float* p) {
#if defined(__CUDA_ARCH__)
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  return f4{v.x, v.y, v.z, v.w};
#else
  return f4{p[0], p[1], p[2], p[3]};
#endif
}
CI_HD float load1(const float* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

// sb   : per-lane scratch column for beta / beta_bar, element j at sb[j * sbld]
// sr   : per-lane scratch for the 8 residuals of a chunk: element i at sr[(i >> 2) * srld + (i & 3)]
// tape : [T][S+2] with element stride ldt
// u, grad : [P] with element stride ld
template <int S>
CI_HD void eval_lane(const float* __restrict__ u, long ld, const float* __restrict__ sc, long scld,
                     const float* __restrict__ yn, const float* __restrict__ Xn, int T, int k, int r_dur,
                     float* __restrict__ sb, int sbld, float* __restrict__ sr, int srld,
                     float* __restrict__ tape, long ldt, float* __restrict__ logp_out,
                     float* __restrict__ grad) {
  constexpr int TW = S + 2;
  const LaneScalars ls = params_forward_lane(u, ld, k, S, sc, scld, sb, sbld);
  const float R = ls.R, ql = ls.ql, qd = ls.qd;
  const float mean = sc[SC_MEAN * scld];
  const int nchunk = (T + 7) >> 3;

  // ---- forward: residual chunk (regression GEMV) + 8 filter steps ------------------------------
  KState<S> ks;
  kf_init<S>(ks, sc[SC_M0 * scld], sc[SC_P0 * scld]);
  float acc = 0.f, comp = 0.f;  // compensated sum of log s + v^2/s
  int nobs = 0;
  int phase = 0;
  for (int tc = 0; tc < nchunk; ++tc) {
    const int t0 = tc << 3;
    const int nstep = (T - t0) < 8 ? (T - t0) : 8;
    {
      float r[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = (i < nstep ? load1(yn + t0 + i) : 0.f) - mean;
      const float* __restrict__ xp = Xn + (long)tc * k * 8;
#pragma unroll 4
      for (int j = 0; j < k; ++j) {
        const float bj = sb[j * sbld];
        const f4 x0 = load4(xp + j * 8), x1 = load4(xp + j * 8 + 4);
        r[0] = fmaf(-bj, x0.x, r[0]); r[1] = fmaf(-bj, x0.y, r[1]);
        r[2] = fmaf(-bj, x0.z, r[2]); r[3] = fmaf(-bj, x0.w, r[3]);
        r[4] = fmaf(-bj, x1.x, r[4]); r[5] = fmaf(-bj, x1.y, r[5]);
        r[6] = fmaf(-bj, x1.z, r[6]); r[7] = fmaf(-bj, x1.w, r[7]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) sr[(i >> 2) * srld + (i & 3)] = r[i];
    }
    for (int i = 0; i < nstep; ++i) {
      const float r = sr[(i >> 2) * srld + (i & 3)];
      float* tp = tape + (long)(t0 + i) * TW * ldt;
      if (r == r) {
        float g[S], s, v, is;
        kf_update<S>(ks, r, R, g, s, v, is);
#pragma unroll
        for (int q = 0; q < S; ++q) tp[(long)q * ldt] = g[q];
        tp[(long)S * ldt] = s;
        tp[(long)(S + 1) * ldt] = v;
        const float term = logf(s) + v * v * is;
        const float yk = term - comp;
        const float tk = acc + yk;
        comp = (tk - acc) - yk;
        acc = tk;
        ++nobs;
      } else {
#pragma unroll
        for (int q = 0; q < TW; ++q) tp[(long)q * ldt] = (q == S) ? -1.f : 0.f;
      }
      const bool change = (phase == r_dur - 1);
      kf_predict<S>(ks, ql, qd, change);
      phase = change ? 0 : phase + 1;
    }
  }
  const float ll = -0.5f * (acc + (float)nobs * CI_LOG_2PI);

  // ---- reverse: 8 adjoint steps + beta_bar accumulation per chunk ------------------------------
  for (int j = 0; j < k; ++j) sb[j * sbld] = 0.f;
  KState<S> b;
#pragma unroll
  for (int i = 0; i < S; ++i) b.m[i] = 0.f;
#pragma unroll
  for (int i = 0; i < KState<S>::NP; ++i) b.P[i] = 0.f;
  float Rb = 0.f, qlb = 0.f, qdb = 0.f;
  phase = (T - 1) % r_dur;
  for (int i = 7; i >= (T & 7) && (T & 7) != 0; --i) sr[(i >> 2) * srld + (i & 3)] = 0.f;
  // software-pipelined tape reads: the entry of step t-1 is requested before step t is processed
  float cur[TW], nxt[TW] = {};
#pragma unroll
  for (int q = 0; q < TW; ++q) cur[q] = tape[((long)(T - 1) * TW + q) * ldt];
  for (int t = T - 1; t >= 0; --t) {
    const int i = t & 7;
    if (t > 0) {
#pragma unroll
      for (int q = 0; q < TW; ++q) nxt[q] = tape[((long)(t - 1) * TW + q) * ldt];
    }
    const bool change = (phase == r_dur - 1);
    kb_predict_adj<S>(b, qlb, qdb, change);
    phase = (phase == 0) ? r_dur - 1 : phase - 1;
    float rb = 0.f;
    if (cur[S] > 0.f) {
      float g[S];
#pragma unroll
      for (int q = 0; q < S; ++q) g[q] = cur[q];
      rb = kb_update_adj<S>(b, g, cur[S], cur[S + 1], Rb);
    }
    sr[(i >> 2) * srld + (i & 3)] = rb;
    if (i == 0 && k > 0) {
      float rbv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) rbv[q] = sr[(q >> 2) * srld + (q & 3)];
      const float* __restrict__ xp = Xn + (long)(t >> 3) * k * 8;
#pragma unroll 4
      for (int j = 0; j < k; ++j) {
        const f4 x0 = load4(xp + j * 8), x1 = load4(xp + j * 8 + 4);
        float a0 = sb[j * sbld], a1 = 0.f;
        a0 = fmaf(-rbv[0], x0.x, a0); a1 = fmaf(-rbv[1], x0.y, a1);
        a0 = fmaf(-rbv[2], x0.z, a0); a1 = fmaf(-rbv[3], x0.w, a1);
        a0 = fmaf(-rbv[4], x1.x, a0); a1 = fmaf(-rbv[5], x1.y, a1);
        a0 = fmaf(-rbv[6], x1.z, a0); a1 = fmaf(-rbv[7], x1.w, a1);
        sb[j * sbld] = a0 + a1;
      }
    }
#pragma unroll
    for (int q = 0; q < TW; ++q) cur[q] = nxt[q];
  }
  params_backward_lane(u, ld, k, S, sc, scld, sb, sbld, Rb, qlb, qdb, grad);
  *logp_out = ll + ls.lp0;
}

}  // namespace ci
