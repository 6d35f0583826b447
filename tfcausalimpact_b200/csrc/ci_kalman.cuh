// Per-lane Kalman filter log-likelihood + hand-derived reverse-time adjoint for the fixed layout
//     x = [level | z_1 .. z_{S-1}],  H = [1 | 1 0 .. 0],
//     F = blockdiag(1, A) on season-change steps (A z)_i = z_{i+1}, (A z)_{S-2} = -sum z,
//     Q = blockdiag(ql, qd * 1 1^T) on season-change steps, diag(ql, 0) otherwise.
// One *thread* owns one (series, chain[, sample]) lane: mean, packed-symmetric covariance and
// their adjoints live in registers (S <= CI_MAX_REG_S).  These are the step functions; the per-lane
// drivers that chain them over time live in ci_eval_lane.cuh (fit) and ci_predict.cu (filter).
// This residual-basis form serves season_duration > 1, nseasons = 1 and nseasons = 12; for
// season_duration == 1 and 2 <= nseasons <= 7 the fit uses ci_kalman_eff.cuh.
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

}  // namespace ci
