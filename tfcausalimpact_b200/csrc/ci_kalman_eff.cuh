// Seasonal-effects formulation of the per-lane Kalman filter + adjoint, used when
// season_duration == 1 (every step is a season change) and 2 <= nseasons <= CI_MAX_EFF_S.
//
// Same model as ci_kalman.cuh in different coordinates.  There the latent holds the first S-1
// zero-sum seasonal residuals in a frame that ROTATES with the season (transition = companion-like
// shift + negated row sums: ~90 adds and moves per step).  Here the latent holds all S seasonal
// effects in a FIXED frame, x = [level | e_0 .. e_{S-1}], with the zero-sum constraint carried
// implicitly by the (singular) covariance:
//     H_t = e_level + e_{c},  c = t mod S          (the season observed at step t)
//     F   = I                                      (nothing moves)
//     Q_t = diag(ql, qd w w^T),  w = 1 - S e_c     (the drift xi is added to every other effect and
//                                                   -(S-1) xi to the one just observed)
//     P_0 = diag(p0, p0 (I - 11^T / S)),  m_0 = [y0c, 0]
// which is the change of basis e = [z; -sum z] applied to SURVEY A.3 (R2E A E2R is the cyclic shift
// of the effects; R2E (qd 11^T) R2E^T = qd w w^T).  With the time loop unrolled by S the season index
// is a compile-time constant in every step, so all register indices are static and the time update
// costs S(S+1)/2 + 1 adds and no data movement.  The filter outputs (s_t, v_t, log-likelihood) are
// identical up to rounding; parity with the dense oracle is checked at the same rtol 1e-4.
#pragma once
#include "ci_core.cuh"

#define CI_MAX_EFF_S 7

namespace ci {

template <int S>
struct EState {
  static constexpr int D = S + 1;
  static constexpr int NP = D * (D + 1) / 2;
  float m[D];
  float P[NP];
};

template <int S>
CI_HD void ef_init(EState<S>& k, float m0_level, float p0) {
  constexpr int D = S + 1;
#pragma unroll
  for (int i = 0; i < D; ++i) k.m[i] = 0.f;
  k.m[0] = m0_level;
#pragma unroll
  for (int i = 0; i < EState<S>::NP; ++i) k.P[i] = 0.f;
  k.P[tri<D>(0, 0)] = p0;
  const float off = -p0 * (1.0f / S);
#pragma unroll
  for (int i = 1; i < D; ++i) {
#pragma unroll
    for (int j = i; j < D; ++j) k.P[tri<D>(i, j)] = (i == j) ? (p0 + off) : off;
  }
}

// Measurement update at season C with residual r; emits g = P H^T, s, v.
template <int S, int C>
CI_HD void ef_update(EState<S>& k, float r, float R, float (&g)[S + 1], float& s, float& v, float& is) {
  constexpr int D = S + 1;
  constexpr int J = 1 + C;
#pragma unroll
  for (int i = 0; i < D; ++i) g[i] = k.P[sym<D>(i, 0)] + k.P[sym<D>(i, J)];
  s = g[0] + g[J] + R;
  v = r - k.m[0] - k.m[J];
  is = rcp(s);
  const float vis = v * is;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    k.m[i] = fmaf(g[i], vis, k.m[i]);
    const float ki = g[i] * is;
#pragma unroll
    for (int j = i; j < D; ++j) k.P[tri<D>(i, j)] = fmaf(-ki, g[j], k.P[tri<D>(i, j)]);
  }
}

// Time update after observing season C: P += diag(ql, qd w w^T), w = 1 - S e_C.
// qw = (qd, -(S-1) qd, (S-1)^2 qd) precomputed by the caller.
template <int S, int C>
CI_HD void ef_predict(EState<S>& k, float ql, float q11, float q1c, float qcc) {
  constexpr int D = S + 1;
  constexpr int J = 1 + C;
  k.P[tri<D>(0, 0)] += ql;
#pragma unroll
  for (int i = 1; i < D; ++i) {
#pragma unroll
    for (int j = i; j < D; ++j) {
      const float q = (i == J && j == J) ? qcc : ((i == J || j == J) ? q1c : q11);
      k.P[tri<D>(i, j)] += q;
    }
  }
}

// Adjoint of ef_predict: accumulates dL/dql and dL/dqd = w^T Pbar_seas w
//   = 1^T W 1 - 2 S (W 1)_c + S^2 W_cc   (W = seasonal block of Pbar).
// `sig` = 1^T W 1 is carried incrementally by eb_update_adj: Pbar only ever changes by sym(gb H^T), which
// adds sum_{j >= 1} gb_j to the block sum, so the S(S+1)/2-term sum is not recomputed every step.
template <int S, int C>
CI_HD void eb_predict_adj(const EState<S>& b, float& qlb, float& qdb, float sig) {
  constexpr int D = S + 1;
  constexpr int J = 1 + C;
  qlb += b.P[tri<D>(0, 0)];
  // two partial sums, seeded with their first terms (0 + x is not folded by the compiler)
  float r0 = b.P[sym<D>(J, 1)], r1 = D > 2 ? b.P[sym<D>(J, 2)] : 0.f;
#pragma unroll
  for (int i = 3; i < D; ++i) {
    if (i & 1) r0 += b.P[sym<D>(J, i)]; else r1 += b.P[sym<D>(J, i)];
  }
  qdb += fmaf((float)(S * S), b.P[tri<D>(J, J)], fmaf(-(2.f * S), r0 + r1, sig));
}

// Adjoint of ef_update given the taped (g, v) and recomputed s.  Returns rbar = dL/dr_t.
template <int S, int C>
CI_HD float eb_update_adj(EState<S>& b, const float (&g)[S + 1], float s, float v, float& Rb, float& sig) {
  constexpr int D = S + 1;
  constexpr int J = 1 + C;
  const float is = rcp(s);
  float a = 0.f;
  float Pg[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    a = fmaf(b.m[i], g[i], a);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) acc = fmaf(b.P[sym<D>(i, j)], g[j], acc);
    Pg[i] = acc;
  }
  float gPg = 0.f;
#pragma unroll
  for (int i = 0; i < D; ++i) gPg = fmaf(g[i], Pg[i], gPg);
  const float vis = v * is;
  const float vb = (a - v) * is;
  const float sb = (gPg - a * v) * is * is - 0.5f * (is - vis * vis);
  Rb += sb;
  float gb[D];
#pragma unroll
  for (int i = 0; i < D; ++i) gb[i] = fmaf(b.m[i], vis, fmaf(-2.f * is, Pg[i], (i == 0 || i == J) ? sb : 0.f));
  {
    float s0 = gb[1], s1 = D > 2 ? gb[2] : 0.f;
#pragma unroll
    for (int j = 3; j < D; ++j) {
      if (j & 1) s0 += gb[j]; else s1 += gb[j];
    }
    sig += s0 + s1;
  }
  // Pbar += sym(gb H^T), H = e_0 + e_J
  b.P[tri<D>(0, 0)] += gb[0];
  b.P[tri<D>(J, J)] += gb[J];
  b.P[tri<D>(0, J)] += 0.5f * (gb[0] + gb[J]);
#pragma unroll
  for (int j = 1; j < D; ++j) {
    if (j != J) {
      b.P[tri<D>(0, j)] += 0.5f * gb[j];
      b.P[sym<D>(J, j)] += 0.5f * gb[j];
    }
  }
  b.m[0] -= vb;
  b.m[J] -= vb;
  return vb;
}

}  // namespace ci
