// Per-lane Kalman filter log-likelihood + hand-derived reverse-time adjoint for the fixed layout
//     x = [level | z_1 .. z_{S-1}],  H = [1 | 1 0 .. 0],
//     F = blockdiag(1, A) on season-change steps (A z)_i = z_{i+1}, (A z)_{S-2} = -sum z,
//     Q = blockdiag(ql, qd * 1 1^T) on season-change steps, diag(ql, 0) otherwise.
// One *thread* owns one (series, chain[, sample]) lane: mean, packed-symmetric covariance and
// their adjoints live in registers (S <= CI_MAX_REG_S); lanes are the fastest-moving index of
// every global array so each warp access is one 128-byte line.
//
// Replaces: tfd.LinearGaussianStateSpaceModel.log_prob + TF autodiff as reached from
// /root/reference/causalimpact/model.py:361-364 (fit_with_hmc) and :374-381 (VI).
// Math: SURVEY.md Appendix A.4 (forward) / A.5 (reverse sweep).  The reverse sweep needs only
// (g_t, s_t, v_t) from the forward pass -- (S+2) floats per step -- because P_f = P - g g^T / s is
// linear in P and quadratic in g = P H^T.
#pragma once
#include "ci_core.cuh"

namespace ci {

template <int S>
struct KState {
  static constexpr int NP = S * (S + 1) / 2;
  float m[S];
  float P[NP];
};

// ---- forward pieces --------------------------------------------------------------------------

template <int S>
CI_HD void kf_init(KState<S>& k, float m0_level, float p0) {
#pragma unroll
  for (int i = 0; i < S; ++i) k.m[i] = 0.f;
  k.m[0] = m0_level;
#pragma unroll
  for (int i = 0; i < KState<S>::NP; ++i) k.P[i] = 0.f;
  k.P[tri<S>(0, 0)] = p0;
  if (S > 1) {
    const float off = -p0 * (1.0f / S);
#pragma unroll
    for (int i = 1; i < S; ++i) {
#pragma unroll
      for (int j = i; j < S; ++j) k.P[tri<S>(i, j)] = (i == j) ? (p0 + off) : off;
    }
  }
}

// Measurement update with residual r (caller guarantees !isnan(r)).  Emits g, s, v.
template <int S>
CI_HD void kf_update(KState<S>& k, float r, float R, float (&g)[S], float& s, float& v, float& is) {
#pragma unroll
  for (int i = 0; i < S; ++i) g[i] = k.P[sym<S>(i, 0)] + (S > 1 ? k.P[sym<S>(i, 1)] : 0.f);
  s = g[0] + (S > 1 ? g[1] : 0.f) + R;
  v = r - k.m[0] - (S > 1 ? k.m[1] : 0.f);
  is = rcp(s);
  const float vis = v * is;
#pragma unroll
  for (int i = 0; i < S; ++i) {
    k.m[i] = fmaf(g[i], vis, k.m[i]);
    const float ki = g[i] * is;
#pragma unroll
    for (int j = i; j < S; ++j) k.P[tri<S>(i, j)] = fmaf(-ki, g[j], k.P[tri<S>(i, j)]);
  }
}

// Time update.  `change` = last step of a season (always true when season_duration == 1).
template <int S>
CI_HD void kf_predict(KState<S>& k, float ql, float qd, bool change) {
  k.P[tri<S>(0, 0)] += ql;
  if (S > 1 && change) {
    constexpr int Z = (S > 1) ? S - 1 : 1;  // seasonal block size
    // mean: z'_i = z_{i+1}, z'_{Z-1} = -sum z
    float zs = 0.f;
#pragma unroll
    for (int i = 1; i < S; ++i) zs += k.m[i];
#pragma unroll
    for (int i = 1; i < S - 1; ++i) k.m[i] = k.m[i + 1];
    k.m[S - 1] = -zs;
    // level/seasonal cross row
    float cs = 0.f;
#pragma unroll
    for (int j = 1; j < S; ++j) cs += k.P[tri<S>(0, j)];
#pragma unroll
    for (int j = 1; j < S - 1; ++j) k.P[tri<S>(0, j)] = k.P[tri<S>(0, j + 1)];
    k.P[tri<S>(0, S - 1)] = -cs;
    // seasonal block: row sums of the symmetric block
    float rs[Z];
    float tot = 0.f;
#pragma unroll
    for (int a = 0; a < Z; ++a) {
      float acc = 0.f;
#pragma unroll
      for (int l = 0; l < Z; ++l) acc += k.P[sym<S>(1 + a, 1 + l)];
      rs[a] = acc;
      tot += acc;
    }
#pragma unroll
    for (int i = 0; i < Z - 1; ++i) {
#pragma unroll
      for (int j = i; j < Z - 1; ++j) k.P[tri<S>(1 + i, 1 + j)] = k.P[tri<S>(2 + i, 2 + j)] + qd;
      k.P[tri<S>(1 + i, S - 1)] = qd - rs[i + 1];
    }
    k.P[tri<S>(S - 1, S - 1)] = tot + qd;
  }
}

// ---- reverse pieces --------------------------------------------------------------------------

template <int S>
CI_HD void kb_predict_adj(KState<S>& b, float& qlb, float& qdb, bool change) {
  qlb += b.P[tri<S>(0, 0)];
  if (S > 1 && change) {
    constexpr int Z = (S > 1) ? S - 1 : 1;
    // qd touches every entry of the seasonal block
    float dsum = 0.f, osum = 0.f;
#pragma unroll
    for (int i = 1; i < S; ++i) {
      dsum += b.P[tri<S>(i, i)];
#pragma unroll
      for (int j = i + 1; j < S; ++j) osum += b.P[tri<S>(i, j)];
    }
    qdb += dsum + 2.f * osum;
    // vectors: (A^T w)_0 = -w_last, (A^T w)_j = w_{j-1} - w_last
    {
      const float wl = b.m[S - 1];
#pragma unroll
      for (int j = S - 1; j >= 2; --j) b.m[j] = b.m[j - 1] - wl;
      b.m[1] = -wl;
      const float cl = b.P[tri<S>(0, S - 1)];
#pragma unroll
      for (int j = S - 1; j >= 2; --j) b.P[tri<S>(0, j)] = b.P[tri<S>(0, j - 1)] - cl;
      b.P[tri<S>(0, 1)] = -cl;
    }
    // block: W'[j,l] = W[j-1,l-1] - W[j-1,last] - W[last,l-1] + W[last,last], W[-1,.] = 0
    float wlast[Z];  // W[a, last], a = 0..Z-1
#pragma unroll
    for (int a = 0; a < Z; ++a) wlast[a] = b.P[sym<S>(1 + a, S - 1)];
    const float wll = wlast[Z - 1];
#pragma unroll
    for (int j = Z - 1; j >= 1; --j) {
#pragma unroll
      for (int l = Z - 1; l >= j; --l)
        b.P[tri<S>(1 + j, 1 + l)] = b.P[tri<S>(j, l)] - wlast[j - 1] - wlast[l - 1] + wll;
    }
#pragma unroll
    for (int l = Z - 1; l >= 1; --l) b.P[tri<S>(1, 1 + l)] = wll - wlast[l - 1];
    b.P[tri<S>(1, 1)] = wll;
  }
}

// Adjoint of kf_update given the taped (g, s, v).  Returns rbar = dL/dr_t.
template <int S>
CI_HD float kb_update_adj(KState<S>& b, const float (&g)[S], float s, float v, float& Rb) {
  const float is = rcp(s);
  float a = 0.f;
  float Pg[S];
#pragma unroll
  for (int i = 0; i < S; ++i) {
    a = fmaf(b.m[i], g[i], a);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < S; ++j) acc = fmaf(b.P[sym<S>(i, j)], g[j], acc);
    Pg[i] = acc;
  }
  float gPg = 0.f;
#pragma unroll
  for (int i = 0; i < S; ++i) gPg = fmaf(g[i], Pg[i], gPg);
  const float vis = v * is;
  const float vb = (a - v) * is;
  const float sb = (gPg - a * v) * is * is - 0.5f * (is - vis * vis);
  Rb += sb;
  float gb[S];
#pragma unroll
  for (int i = 0; i < S; ++i) gb[i] = fmaf(b.m[i], vis, fmaf(-2.f * is, Pg[i], (i < 2 ? sb : 0.f)));
  // P += sym(gb H), H = e0 + e1
  if (S == 1) {
    b.P[0] += gb[0];
  } else {
    b.P[tri<S>(0, 0)] += gb[0];
    b.P[tri<S>(1, 1)] += gb[1];
    b.P[tri<S>(0, 1)] += 0.5f * (gb[0] + gb[1]);
#pragma unroll
    for (int j = 2; j < S; ++j) {
      b.P[tri<S>(0, j)] += 0.5f * gb[j];
      b.P[tri<S>(1, j)] += 0.5f * gb[j];
    }
  }
  b.m[0] -= vb;
  if (S > 1) b.m[1] -= vb;
  return vb;
}

// ---- whole-lane driver -----------------------------------------------------------------------
//
// resid : [T] with element stride `ld`; in: r_t = y_t - offset_t (NaN = missing); out: dL/dr_t.
// tape  : [T][S+2] with element stride `ld` (g_0..g_{S-1}, s (negative = masked), v).
// Outputs: log-likelihood and adjoints of R = sigma_obs^2, ql = sigma_level^2, qd = (sigma_drift/S)^2.
template <int S>
CI_HD void kalman_logp_grad_lane(float* __restrict__ resid, float* __restrict__ tape, long ld, int T,
                                 int r_dur, float R, float ql, float qd, float m0_level, float p0,
                                 float& ll_out, float& Rb_out, float& qlb_out, float& qdb_out) {
  constexpr int TW = S + 2;
  KState<S> k;
  kf_init<S>(k, m0_level, p0);
  // log-likelihood accumulated as -0.5 * (sum log s + sum v^2/s) with a compensated sum
  float acc = 0.f, comp = 0.f;
  int nobs = 0;
  int phase = 0;  // t mod r_dur
  for (int t = 0; t < T; ++t) {
    const float r = resid[(long)t * ld];
    float* tp = tape + (long)t * TW * ld;
    if (r == r) {
      float g[S], s, v, is;
      kf_update<S>(k, r, R, g, s, v, is);
#pragma unroll
      for (int i = 0; i < S; ++i) tp[(long)i * ld] = g[i];
      tp[(long)S * ld] = s;
      tp[(long)(S + 1) * ld] = v;
      const float term = logf(s) + v * v * is;
      const float yk = term - comp;
      const float tk = acc + yk;
      comp = (tk - acc) - yk;
      acc = tk;
      ++nobs;
    } else {
      tp[(long)S * ld] = -1.f;
    }
    const bool change = (phase == r_dur - 1);
    kf_predict<S>(k, ql, qd, change);
    phase = change ? 0 : phase + 1;
  }
  ll_out = -0.5f * (acc + (float)nobs * CI_LOG_2PI);

  // reverse sweep
  KState<S> b;
#pragma unroll
  for (int i = 0; i < S; ++i) b.m[i] = 0.f;
#pragma unroll
  for (int i = 0; i < KState<S>::NP; ++i) b.P[i] = 0.f;
  float Rb = 0.f, qlb = 0.f, qdb = 0.f;
  phase = (T - 1) % r_dur;
  for (int t = T - 1; t >= 0; --t) {
    const float* tp = tape + (long)t * TW * ld;
    const bool change = (phase == r_dur - 1);
    kb_predict_adj<S>(b, qlb, qdb, change);
    phase = (phase == 0) ? r_dur - 1 : phase - 1;
    const float s = tp[(long)S * ld];
    float rb = 0.f;
    if (s > 0.f) {
      float g[S];
#pragma unroll
      for (int i = 0; i < S; ++i) g[i] = tp[(long)i * ld];
      const float v = tp[(long)(S + 1) * ld];
      rb = kb_update_adj<S>(b, g, s, v, Rb);
    }
    resid[(long)t * ld] = rb;
  }
  Rb_out = Rb;
  qlb_out = qlb;
  qdb_out = qdb;
}

// Forward-only filter used by the predictive path (one_step_predictive / forecast prior):
// emits per-step predictive mean of the residual (H m) and variance s, and returns the predicted
// state after the last step.  Masked steps still report (H m, s) from the prior, as TFP does.
template <int S>
CI_HD void kalman_filter_lane(const float* __restrict__ resid, long ld_in, float* __restrict__ hm_out,
                              float* __restrict__ s_out, long ld_out, int T, int r_dur, float R, float ql,
                              float qd, float m0_level, float p0, KState<S>& k) {
  kf_init<S>(k, m0_level, p0);
  int phase = 0;
  for (int t = 0; t < T; ++t) {
    const float r = resid[(long)t * ld_in];
    if (r == r) {
      float g[S], s, v, is;
      kf_update<S>(k, r, R, g, s, v, is);
      hm_out[(long)t * ld_out] = r - v;
      s_out[(long)t * ld_out] = s;
    } else {
      float s = k.P[sym<S>(0, 0)] + R;
      float hm = k.m[0];
      if (S > 1) {
        s += 2.f * k.P[sym<S>(0, 1)] + k.P[sym<S>(1, 1)];
        hm += k.m[1];
      }
      hm_out[(long)t * ld_out] = hm;
      s_out[(long)t * ld_out] = s;
    }
    const bool change = (phase == r_dur - 1);
    kf_predict<S>(k, ql, qd, change);
    phase = change ? 0 : phase + 1;
  }
}

}  // namespace ci
