// Large-latent path: ONE WARP per (series, chain[, sample]) lane, used when nseasons is outside the
// register-resident set (e.g. nseasons = 52, BASELINE.json configs[4]).  Same model, seasonal-effects
// coordinates (ci_kalman_eff.cuh: x = [level | e_0..e_{S-1}], H_t = e_level + e_c, F = I,
// Q_t = diag(ql, qd w w^T) on season changes), any season_duration, runtime S.
//
// The D x D covariance (D = S + 1) and its adjoint live in shared memory as full row-major matrices;
// lane l of the warp owns columns l, l + 32, ... of every row, so row sweeps are conflict-free
// (consecutive lanes -> consecutive banks) and the rank-1 update, the rank-1 drift term and the
// symmetric matvec of the adjoint need no shuffles: the row scalar (g_i, w_i) is a shared-memory
// broadcast and the column scalars sit in registers.  Per step: O(D^2 / 32) FMAs per thread.
//
//   k_eval_big    joint log-prob + gradient (forward filter, tape [B][T][D+1], reverse adjoint sweep,
//                 warp-cooperative regression GEMV / beta_bar)          replaces: as ci_eval.cu
//   k_filter_big  forward filter per posterior draw -> one-step predictive mean / variance,
//                 predicted state + guarded Cholesky                    replaces: as ci_predict.cu K6
//   k_forecast_mean_big / k_forecast_sample_big   forecast mean path and joint trajectories   (K7)
// Forecast-state layout of this path: lane-contiguous [B][D + 2 D^2] = m, P (full), chol(P) (full).
//
// A SECOND seasonal component (ci_layout.S2 > 1; tfp.sts.Sum([..., Seasonal(S), Seasonal(S2)]), the
// notebook's Month + Year model) appends its S2 effects to the latent vector: H_t gains a third one
// (e_{so2 + c2}), Q_t a second rank-1 drift block.  Kernels are templated on TWO so the single-seasonal
// instantiations carry none of it.
#include "ci_abi_common.cuh"
#include "ci_eval_lane.cuh"

namespace ci {

constexpr int BIG_WARPS = 4;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// latent layout: [level | slope (LocalLinearTrend only) | e_0 .. e_{S-1} (S > 1 only) | e'_0 .. e'_{S2-1} (S2 > 1 only)]
__host__ __device__ __forceinline__ int big_nt(const ci_layout& L) { return (L.flags & CI_FLAG_LLT) ? 2 : 1; }
__host__ __device__ __forceinline__ int big_so2(const ci_layout& L) { return big_nt(L) + (L.S > 1 ? L.S : 0); }
__host__ __device__ __forceinline__ int big_D(const ci_layout& L) { return big_so2(L) + (L.S2 > 1 ? L.S2 : 0); }
__host__ __device__ __forceinline__ bool big_gen(const ci_layout& L) { return L.flags != 0 || L.S2 > 1; }

__device__ __forceinline__ int season_of(int t, int S, int r) { return (t / r) % S; }
__device__ __forceinline__ bool season_change(int t, int S, int r) { return S > 1 && (t % r) == r - 1; }
// drift-direction weight of state index i (0 = level) when season c was just observed
// (so = index of the first seasonal effect)
__device__ __forceinline__ float drift_w(int i, int c, int S, int so) {
  return (i < so || i >= so + S) ? 0.f : (i == so + c ? (float)(1 - S) : 1.f);
}

__device__ void big_init(float* P, float* m, int D, int S, int so, int S2, float m0, float p0, float sd, int lane) {
  const float off = S > 1 ? -p0 / (float)S : 0.f;
  const float off2 = S2 > 1 ? -p0 / (float)S2 : 0.f;
  const int so2 = so + (S > 1 ? S : 0);
  for (int i = 0; i < D; ++i)
    for (int j = lane; j < D; j += 32) {
      float v = 0.f;
      if (i < so || j < so) {
        if (i == j) v = (i == 0) ? p0 : sd * sd;      // level: (|y0c| + sd)^2 ; slope: sd^2
      } else if (i < so2 && j < so2) {
        v = (i == j) ? p0 + off : off;
      } else if (i >= so2 && j >= so2) {
        v = (i == j) ? p0 + off2 : off2;
      }
      P[i * D + j] = v;
    }
  for (int j = lane; j < D; j += 32) m[j] = (j == 0) ? m0 : 0.f;
  __syncwarp();
}

// residual of step t for the warp's lane: centred y minus the regression, summed across the warp
__device__ __forceinline__ float big_resid(const float* __restrict__ XY, const ci_layout& L, int n, int t,
                                           const float* beta, int lane) {
  const int stride = (L.k + 1) * CI_TC;
  const float* xp = XY + ((long)(t / CI_TC) * L.N + n) * stride + (t % CI_TC);
  float acc = 0.f;
  for (int j = lane; j < L.k; j += 32) acc = fmaf(beta[j], __ldg(xp + j * CI_TC), acc);
  acc = warp_sum(acc);
  return __ldg(xp + L.k * CI_TC) - acc;
}

// Column slots of a lane: columns lane, lane + 32, ... (NC = ceil(D / 32) of them, compile time so the
// row sweeps are branch-uniform and the column scalars stay in registers).
template <int NC>
struct Cols {
  int j[NC];
  bool ok[NC];
  __device__ __forceinline__ Cols(int lane, int D) {
#pragma unroll
    for (int q = 0; q < NC; ++q) {
      j[q] = lane + 32 * q;
      ok[q] = j[q] < D;
      if (!ok[q]) j[q] = 0;   // harmless in-range address; never written
    }
  }
};

// One filter step on the shared-memory state.  Returns false when the step is masked.
// Emits g (shared) and s, v; the time update (P += Q_t) is fused into the same row sweep.
// season bookkeeping of one step (both components)
struct BigSeason {
  int c, J, c2, J2;        // season index / latent index of the observed effect (J2 = 0 when !TWO)
  bool change, change2;
};
template <bool TWO>
__device__ __forceinline__ BigSeason big_season(const ci_layout& L, int so, int so2, int t) {
  BigSeason z;
  z.c = L.S > 1 ? season_of(t, L.S, L.r) : 0;
  z.J = L.S > 1 ? so + z.c : 0;
  z.change = season_change(t, L.S, L.r);
  z.c2 = TWO ? season_of(t, L.S2, L.r2) : 0;
  z.J2 = TWO ? so2 + z.c2 : 0;
  z.change2 = TWO && season_change(t, L.S2, L.r2);
  return z;
}

// Incremental form of big_season for loops over consecutive steps (no integer divisions per step).
template <bool TWO>
struct SeasonWalk {
  int S, r, S2, r2, so, so2;
  int c, ph, c2, ph2;     // season index and position inside the season, per component
  __device__ __forceinline__ SeasonWalk(const ci_layout& L, int so_, int so2_, int t)
      : S(L.S), r(L.r), S2(TWO ? L.S2 : 1), r2(TWO ? L.r2 : 1), so(so_), so2(so2_) {
    c = S > 1 ? (t / r) % S : 0;
    ph = S > 1 ? t % r : 0;
    c2 = TWO ? (t / r2) % S2 : 0;
    ph2 = TWO ? t % r2 : 0;
  }
  __device__ __forceinline__ BigSeason get() const {
    BigSeason z;
    z.c = c;
    z.J = S > 1 ? so + c : 0;
    z.change = S > 1 && ph == r - 1;
    z.c2 = c2;
    z.J2 = TWO ? so2 + c2 : 0;
    z.change2 = TWO && ph2 == r2 - 1;
    return z;
  }
  __device__ __forceinline__ void next() {
    if (S > 1 && ++ph == r) { ph = 0; c = (c + 1 == S) ? 0 : c + 1; }
    if (TWO && ++ph2 == r2) { ph2 = 0; c2 = (c2 + 1 == S2) ? 0 : c2 + 1; }
  }
  __device__ __forceinline__ void prev() {
    if (S > 1 && --ph < 0) { ph = r - 1; c = (c == 0) ? S - 1 : c - 1; }
    if (TWO && --ph2 < 0) { ph2 = r2 - 1; c2 = (c2 == 0) ? S2 - 1 : c2 - 1; }
  }
};

template <int NC, int RB, bool TWO>
__device__ __forceinline__ bool big_step(float* __restrict__ P, float* __restrict__ m, float* __restrict__ g, int D, int S, int so,
                                         int S2, int so2, const BigSeason z, float resid, float R, float ql, float qs,
                                         float qd, float qd2, float& s, float& v, float& is, int lane) {
  const Cols<NC> cl(lane, D);
  const int c = z.c, J = z.J, J2 = z.J2;
  const bool obs = (resid == resid);
  float gj[NC], wj[NC], w2j[NC];
#pragma unroll
  for (int q = 0; q < NC; ++q) {
    gj[q] = cl.ok[q] ? P[cl.j[q] * D] + (S > 1 ? P[cl.j[q] * D + J] : 0.f) + (TWO ? P[cl.j[q] * D + J2] : 0.f) : 0.f;
    wj[q] = drift_w(cl.j[q], c, S, so);
    w2j[q] = TWO ? drift_w(cl.j[q], z.c2, S2, so2) : 0.f;
    if (cl.ok[q]) g[cl.j[q]] = gj[q];
  }
  __syncwarp();
  s = g[0] + (S > 1 ? g[J] : 0.f) + (TWO ? g[J2] : 0.f) + R;
  v = resid - m[0] - (S > 1 ? m[J] : 0.f) - (TWO ? m[J2] : 0.f);
  is = rcp(s);
  const float vis = v * is;
  __syncwarp();
  const float nis = obs ? -is : 0.f;
  const float qq = z.change ? qd : 0.f;
  const float qq2 = z.change2 ? qd2 : 0.f;
  if (obs) {
#pragma unroll
    for (int q = 0; q < NC; ++q)
      if (cl.ok[q]) m[cl.j[q]] = fmaf(gj[q], vis, m[cl.j[q]]);
  }
  // rows in batches of RB (4 when a lone warp per scheduler is latency-bound, 8 when many lanes keep the
  // SM busy): all loads of a batch are issued before its stores, so shared-memory latency is
  // paid once per batch instead of once per row (the compiler cannot reorder a load of P past a store to P)
  for (int i0 = 0; i0 < D; i0 += RB) {
    float p[RB][NC], nki[RB], qi[RB], q2i[RB];
#pragma unroll
    for (int e = 0; e < RB; ++e) {
      const int i = i0 + e < D ? i0 + e : D - 1;
      nki[e] = g[i] * nis;
      qi[e] = qq * drift_w(i, c, S, so);
      q2i[e] = TWO ? qq2 * drift_w(i, z.c2, S2, so2) : 0.f;
#pragma unroll
      for (int q = 0; q < NC; ++q) p[e][q] = P[i * D + cl.j[q]];
    }
#pragma unroll
    for (int e = 0; e < RB; ++e) {
      if (i0 + e < D) {
#pragma unroll
        for (int q = 0; q < NC; ++q)
          if (cl.ok[q]) {
            float nv = fmaf(qi[e], wj[q], fmaf(nki[e], gj[q], p[e][q]));
            if (TWO) nv = fmaf(q2i[e], w2j[q], nv);
            P[(i0 + e) * D + cl.j[q]] = nv;
          }
      }
    }
  }
  __syncwarp();
  if (so == 2) {   // LocalLinearTrend transition F = I + e_level e_slope^T:  P <- F P F^T,  m <- F m
#pragma unroll
    for (int q = 0; q < NC; ++q)
      if (cl.ok[q]) P[cl.j[q]] += P[D + cl.j[q]];          // row 0 += row 1
    __syncwarp();
#pragma unroll
    for (int q = 0; q < NC; ++q)
      if (cl.ok[q]) P[cl.j[q] * D] += P[cl.j[q] * D + 1];  // column 0 += column 1
    __syncwarp();
    if (lane == 0) {
      m[0] += m[1];
      P[D + 1] += qs;
    }
  }
  if (lane == 0) P[0] += ql;
  __syncwarp();
  return obs;
}

template <int NC, bool TWO>
__global__ void __launch_bounds__(BIG_WARPS * 32)
k_eval_big(ci_layout L, const float* __restrict__ XY, const float* __restrict__ sconst, const float* __restrict__ u,
           float* __restrict__ logp, float* __restrict__ grad, float* __restrict__ tape) {
  extern __shared__ __align__(16) float smem[];
  const int S = L.S, so = big_nt(L), so2 = big_so2(L), S2 = TWO ? L.S2 : 1, D = big_D(L), k = L.k, T = L.T;
  const long B = (long)L.N * L.G;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long b = (long)blockIdx.x * BIG_WARPS + w;
  if (b >= B) return;
  const int n = (int)(b / L.G);
  const int per_warp = 2 * D * D + 4 * D + (k > 0 ? k : 1) + 8;
  float* P = smem + (long)w * per_warp;
  float* Pb = P + D * D;
  float* m = Pb + D * D;
  float* g = m + D;
  float* mb = g + D;
  float* gb = mb + D;
  float* beta = gb + D;
  float* scal = beta + (k > 0 ? k : 1);
  if (lane == 0) {
    if (!big_gen(L)) {
      const LaneScalars ls = params_forward_lane(u + b, B, k, S, sconst + n, L.N, beta, 1);
      scal[0] = ls.R; scal[1] = ls.ql; scal[2] = ls.qd; scal[3] = ls.lp0; scal[4] = 0.f; scal[5] = 0.f;
    } else {
      const GenScalars ls = params_forward_gen(u + b, B, k, S, L.flags, sconst + n, L.N, beta, 1, L.S2);
      scal[0] = ls.R; scal[1] = ls.ql; scal[2] = ls.qd; scal[3] = ls.lp0; scal[4] = ls.qs; scal[5] = ls.qd2;
    }
  }
  __syncwarp();
  const float R = scal[0], ql = scal[1], qd = scal[2], lp0 = scal[3], qs = scal[4], qd2 = scal[5];
  big_init(P, m, D, S, so, S2, sconst[SC_M0 * L.N + n], sconst[SC_P0 * L.N + n], sconst[SC_SD * L.N + n], lane);
  const Cols<NC> cl(lane, D);
  float* tp = tape + b * (long)T * (D + 1);
  float acc = 0.f, comp = 0.f;
  int nobs = 0;
  for (int t = 0; t < T; ++t) {
    const BigSeason z = big_season<TWO>(L, so, so2, t);
    const float resid = big_resid(XY, L, n, t, beta, lane);
    float s, v, is;
    const bool obs = big_step<NC, 4, TWO>(P, m, g, D, S, so, S2, so2, z, resid, R, ql, qs, qd, qd2, s, v, is, lane);
    float* te = tp + (long)t * (D + 1);
#pragma unroll
    for (int q = 0; q < NC; ++q)
      if (cl.ok[q]) te[cl.j[q]] = g[cl.j[q]];
    if (lane == 0) te[D] = obs ? v : __int_as_float(0x7fc00000);
    if (obs) {
      const float term = logf(s) + v * v * is;
      const float yk = term - comp;
      const float tk = acc + yk;
      comp = (tk - acc) - yk;
      acc = tk;
      ++nobs;
    }
    __syncwarp();
  }
  const float ll = -0.5f * (acc + (float)nobs * CI_LOG_2PI);

  // ---- reverse sweep -----------------------------------------------------------------------------
  for (int i = lane; i < D * D; i += 32) Pb[i] = 0.f;
  for (int j = lane; j < D; j += 32) mb[j] = 0.f;
  for (int j = lane; j < k; j += 32) beta[j] = 0.f;   // becomes beta_bar
  __syncwarp();
  float Rb = 0.f, qlb = 0.f, qsb = 0.f, qdb = 0.f, qd2b = 0.f;
  const int stride = (k + 1) * CI_TC;
  // software-pipelined tape reads: the entry of step t-1 is requested before step t is processed
  float cur_g[NC], cur_v, nxt_g[NC], nxt_v = 0.f;
  {
    const float* te = tp + (long)(T - 1) * (D + 1);
    cur_v = te[D];
#pragma unroll
    for (int q = 0; q < NC; ++q) {
      cur_g[q] = cl.ok[q] ? te[cl.j[q]] : 0.f;
      nxt_g[q] = 0.f;
    }
  }
  for (int t = T - 1; t >= 0; --t) {
    const BigSeason z = big_season<TWO>(L, so, so2, t);
    const int c = z.c, J = z.J, J2 = z.J2;
    if (t > 0) {
      const float* tn = tp + (long)(t - 1) * (D + 1);
      nxt_v = tn[D];
#pragma unroll
      for (int q = 0; q < NC; ++q) nxt_g[q] = cl.ok[q] ? tn[cl.j[q]] : 0.f;
    }
    const float v = cur_v;
    const bool obs = (v == v);
    const bool change = z.change, change2 = z.change2;
    float gj[NC], wj[NC], w2j[NC];
#pragma unroll
    for (int q = 0; q < NC; ++q) {
      gj[q] = obs ? cur_g[q] : 0.f;
      wj[q] = cl.ok[q] ? drift_w(cl.j[q], c, S, so) : 0.f;
      w2j[q] = (TWO && cl.ok[q]) ? drift_w(cl.j[q], z.c2, S2, so2) : 0.f;
      if (cl.ok[q]) g[cl.j[q]] = gj[q];
    }
    qlb += Pb[0];
    if (so == 2) {   // adjoint of the LocalLinearTrend transition: Pbar <- F^T Pbar F, mbar <- F^T mbar
      qsb += Pb[D + 1];
      __syncwarp();
#pragma unroll
      for (int q = 0; q < NC; ++q)
        if (cl.ok[q]) Pb[D + cl.j[q]] += Pb[cl.j[q]];            // row 1 += row 0
      __syncwarp();
#pragma unroll
      for (int q = 0; q < NC; ++q)
        if (cl.ok[q]) Pb[cl.j[q] * D + 1] += Pb[cl.j[q] * D];    // column 1 += column 0
      if (lane == 0) mb[1] += mb[0];
    }
    __syncwarp();
    // one sweep over Pbar gives the adjoints of the drift terms (w^T Pbar w) and Pg = Pbar g
    float col[NC], nac[NC], nac2[NC];
#pragma unroll
    for (int q = 0; q < NC; ++q) col[q] = nac[q] = nac2[q] = 0.f;
    if (obs || change || change2) {
#pragma unroll 4
      for (int i = 0; i < D; ++i) {
        const float gi = g[i];
        const float wi = drift_w(i, c, S, so);
        const float w2i = TWO ? drift_w(i, z.c2, S2, so2) : 0.f;
#pragma unroll
        for (int q = 0; q < NC; ++q) {
          const float pb = Pb[i * D + cl.j[q]];
          col[q] = fmaf(pb, gi, col[q]);
          nac[q] = fmaf(pb, wi, nac[q]);
          if (TWO) nac2[q] = fmaf(pb, w2i, nac2[q]);
        }
      }
    }
    if (change) {
      float part = 0.f;
#pragma unroll
      for (int q = 0; q < NC; ++q) part = fmaf(wj[q], nac[q], part);
      qdb += warp_sum(part);
    }
    if (TWO && change2) {
      float part = 0.f;
#pragma unroll
      for (int q = 0; q < NC; ++q) part = fmaf(w2j[q], nac2[q], part);
      qd2b += warp_sum(part);
    }
    if (obs) {
      const float s = g[0] + (S > 1 ? g[J] : 0.f) + (TWO ? g[J2] : 0.f) + R;
      const float is = rcp(s);
      float pa = 0.f, pg = 0.f;
#pragma unroll
      for (int q = 0; q < NC; ++q) {
        if (cl.ok[q]) {
          pa = fmaf(mb[cl.j[q]], gj[q], pa);
          pg = fmaf(gj[q], col[q], pg);
        }
      }
      const float a = warp_sum(pa), gPg = warp_sum(pg);
      const float vis = v * is;
      const float vb = (a - v) * is;
      const float sbar = (gPg - a * v) * is * is - 0.5f * (is - vis * vis);
      Rb += sbar;
      float gbj[NC];
#pragma unroll
      for (int q = 0; q < NC; ++q) {
        const bool inH = cl.j[q] == 0 || (S > 1 && cl.j[q] == J) || (TWO && cl.j[q] == J2);
        gbj[q] = cl.ok[q] ? fmaf(mb[cl.j[q]], vis, fmaf(-2.f * is, col[q], inH ? sbar : 0.f)) : 0.f;
      }
      __syncwarp();
      // Pbar += sym(gb H^T): column/row 0, then column/row J (then J2): separate phases because the
      // crossing entries (0,J), (0,J2), (J,J2) are touched by two of them
#pragma unroll
      for (int q = 0; q < NC; ++q) {
        if (cl.ok[q]) {
          const float h = 0.5f * gbj[q];
          Pb[cl.j[q] * D] += h;
          Pb[cl.j[q]] += h;
        }
      }
      __syncwarp();
      if (S > 1) {
#pragma unroll
        for (int q = 0; q < NC; ++q) {
          if (cl.ok[q]) {
            const float h = 0.5f * gbj[q];
            Pb[cl.j[q] * D + J] += h;
            Pb[J * D + cl.j[q]] += h;
          }
        }
      }
      if (TWO) {
        __syncwarp();
#pragma unroll
        for (int q = 0; q < NC; ++q) {
          if (cl.ok[q]) {
            const float h = 0.5f * gbj[q];
            Pb[cl.j[q] * D + J2] += h;
            Pb[J2 * D + cl.j[q]] += h;
          }
        }
      }
      if (lane == 0) {
        mb[0] -= vb;
        if (S > 1) mb[J] -= vb;
        if (TWO) mb[J2] -= vb;
      }
      // beta_bar_j -= rbar X_tj
      const float* xp = XY + ((long)(t / CI_TC) * L.N + n) * stride + (t % CI_TC);
      for (int j = lane; j < k; j += 32) beta[j] = fmaf(-vb, __ldg(xp + j * CI_TC), beta[j]);
    }
    cur_v = nxt_v;
#pragma unroll
    for (int q = 0; q < NC; ++q) cur_g[q] = nxt_g[q];
    __syncwarp();
  }
  (void)gb;
  if (lane == 0) {
    if (!big_gen(L)) params_backward_lane(u + b, B, k, S, sconst + n, L.N, beta, 1, Rb, qlb, qdb, grad + b);
    else params_backward_gen(u + b, B, k, S, L.flags, sconst + n, L.N, beta, 1, Rb, qlb, qsb, qdb, grad + b, L.S2, qd2b);
    logp[b] = ll + lp0;
  }
}

// Forward filter per draw lane: one-step predictive mean/variance + forecast state.
template <int NC, bool TWO>
__global__ void __launch_bounds__(BIG_WARPS * 32)
k_filter_big(ci_layout L, const float* __restrict__ XY, const float* __restrict__ sconst, const float* __restrict__ u,
             float* __restrict__ pred_mean, float* __restrict__ pred_var, float* __restrict__ fstate) {
  extern __shared__ __align__(16) float smem[];
  const int S = L.S, so = big_nt(L), so2 = big_so2(L), S2 = TWO ? L.S2 : 1, D = big_D(L), k = L.k, T = L.T;
  const long B = (long)L.N * L.G;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long b = (long)blockIdx.x * BIG_WARPS + w;
  if (b >= B) return;
  const int n = (int)(b / L.G);
  const int per_warp = 2 * D * D + 4 * D + (k > 0 ? k : 1) + 8;
  float* P = smem + (long)w * per_warp;
  float* Lc = P + D * D;
  float* m = Lc + D * D;
  float* g = m + D;
  float* beta = g + 3 * D;
  float* scal = beta + (k > 0 ? k : 1);
  if (lane == 0) {
    if (!big_gen(L)) {
      const LaneScalars ls = params_forward_lane(u + b, B, k, S, sconst + n, L.N, beta, 1);
      scal[0] = ls.R; scal[1] = ls.ql; scal[2] = ls.qd; scal[4] = 0.f; scal[5] = 0.f;
    } else {
      const GenScalars ls = params_forward_gen(u + b, B, k, S, L.flags, sconst + n, L.N, beta, 1, L.S2);
      scal[0] = ls.R; scal[1] = ls.ql; scal[2] = ls.qd; scal[4] = ls.qs; scal[5] = ls.qd2;
    }
  }
  __syncwarp();
  const float R = scal[0], ql = scal[1], qd = scal[2], qs = scal[4], qd2 = scal[5];
  const float mean = sconst[SC_MEAN * L.N + n];
  big_init(P, m, D, S, so, S2, sconst[SC_M0 * L.N + n], sconst[SC_P0 * L.N + n], sconst[SC_SD * L.N + n], lane);
  for (int t = 0; t < T; ++t) {
    const BigSeason z = big_season<TWO>(L, so, so2, t);
    const float resid = big_resid(XY, L, n, t, beta, lane);      // (y - mean) - X beta, NaN if masked
    // offset_t = mean + X_t beta = y_t - resid when observed; recompute it for masked steps
    const int stride = (k + 1) * CI_TC;
    const float* xp = XY + ((long)(t / CI_TC) * L.N + n) * stride + (t % CI_TC);
    float xb = 0.f;
    for (int j = lane; j < k; j += 32) xb = fmaf(beta[j], __ldg(xp + j * CI_TC), xb);
    xb = warp_sum(xb);
    float s, v, is;
    const float hm = m[0] + (S > 1 ? m[z.J] : 0.f) + (TWO ? m[z.J2] : 0.f);
    big_step<NC, 8, TWO>(P, m, g, D, S, so, S2, so2, z, resid, R, ql, qs, qd, qd2, s, v, is, lane);
    if (lane == 0) {
      pred_mean[(long)t * B + b] = hm + mean + xb;
      pred_var[(long)t * B + b] = s;
    }
  }
  // forecast state: m, P, guarded Cholesky (relative pivot tolerance: P is singular along the
  // zero-sum direction of the seasonal effects)
  float* fs = fstate + b * (long)(D + 2 * D * D);
  for (int j = lane; j < D; j += 32) fs[j] = m[j];
  for (int i = lane; i < D * D; i += 32) {
    fs[D + i] = P[i];
    Lc[i] = 0.f;
  }
  float dmax = 0.f;
  for (int i = 0; i < D; ++i) dmax = fmaxf(dmax, P[i * D + i]);
  const float tol = 1e-6f * dmax;
  __syncwarp();
  for (int j = 0; j < D; ++j) {
    // column j: d = P[j][j] - sum_q L[j][q]^2 ; L[i][j] = (P[i][j] - sum_q L[i][q] L[j][q]) / sqrt(d)
    float part = 0.f;
    for (int q = lane; q < j; q += 32) part = fmaf(Lc[j * D + q], Lc[j * D + q], part);
    const float d = P[j * D + j] - warp_sum(part);
    const float dj = d > tol ? sqrtf(d) : 0.f;
    const float inv = dj > 0.f ? 1.0f / dj : 0.f;
    for (int i = j + lane; i < D; i += 32) {
      float a = P[i * D + j];
      for (int q = 0; q < j; ++q) a = fmaf(-Lc[i * D + q], Lc[j * D + q], a);
      Lc[i * D + j] = (i == j) ? dj : a * inv;
    }
    __syncwarp();
  }
  for (int i = lane; i < D * D; i += 32) fs[D + D * D + i] = Lc[i];
}

__device__ __forceinline__ uint32_t pick_draw_big(uint32_t sample_id, uint32_t stream, int G, uint64_t seed) {
  const float un = philox_uniform4(sample_id, 0u, stream, TAG_PICK, seed).x;
  const int j = (int)(un * (float)G);
  return (uint32_t)(j < G ? j : G - 1);
}

__global__ void k_forecast_mean_big(ci_layout L, const float* __restrict__ fstate, const float* __restrict__ off,
                                    float* __restrict__ pm) {
  const int S = L.S, so = big_nt(L), D = big_D(L);
  const long B = (long)L.N * L.G;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * L.Tf) return;
  const long b = idx % B;
  const int t = (int)(idx / B);
  const float* fs = fstate + b * (long)(D + 2 * D * D);
  const int c = S > 1 ? season_of(L.T + t, S, L.r) : 0;
  const float second = L.S2 > 1 ? fs[big_so2(L) + season_of(L.T + t, L.S2, L.r2)] : 0.f;
  pm[idx] = fs[0] + (so == 2 ? (float)t * fs[1] : 0.f) + (S > 1 ? fs[so + c] : 0.f) + second + off[idx];
}

#define CI_BIG_MAXD 96
// One joint forecast trajectory per thread (state vector in local memory).
__global__ void __launch_bounds__(128)
k_forecast_sample_big(ci_layout L, int nsamp, uint64_t seed, const float* __restrict__ fstate,
                      const float* __restrict__ scal, const float* __restrict__ off,
                      const float* __restrict__ mu_sig, float* __restrict__ sims) {
  __shared__ float tile[128][33];
  const int S = L.S, so = big_nt(L), D = big_D(L), Tf = L.Tf, r = L.r;
  const int n = blockIdx.y;
  const int i = blockIdx.x * 128 + threadIdx.x;
  const long B = (long)L.N * L.G;
  const long si = (long)n * nsamp + (i < nsamp ? i : 0);
  const long b = (long)n * L.G + pick_draw_big((uint32_t)si, 1u, L.G, seed);
  const float sobs = sqrtf(scal[0 * B + b]), sl = sqrtf(scal[1 * B + b]), sdr = sqrtf(scal[2 * B + b]);
  const float sslope = so == 2 ? sqrtf(scal[4 * B + b]) : 0.f;
  const int S2 = L.S2 > 1 ? L.S2 : 1, so2 = big_so2(L);
  const float sdr2 = S2 > 1 ? sqrtf(scal[5 * B + b]) : 0.f;
  const float mu = mu_sig ? mu_sig[n] : 0.f, sg = mu_sig ? mu_sig[L.N + n] : 1.f;
  const float* fs = fstate + b * (long)(D + 2 * D * D);
  const float* Lc = fs + D + D * D;
  float x[CI_BIG_MAXD];
  for (int a = 0; a < D; ++a) x[a] = fs[a];
  for (int q = 0; q < (D + 3) / 4; ++q) {
    const f32x4 zz = philox_normal4_fast((uint32_t)si, (uint32_t)q, 0u, TAG_FORECAST_X0, seed);
    const float z4[4] = {zz.x, zz.y, zz.z, zz.w};
    for (int e = 0; e < 4; ++e) {
      const int col = 4 * q + e;
      if (col < D)
        for (int a = col; a < D; ++a) x[a] = fmaf(Lc[a * D + col], z4[e], x[a]);
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t0 = 0; t0 < Tf; t0 += 32) {
    const int tn = min(32, Tf - t0);
    for (int q = 0; q < tn; ++q) {
      const int t = t0 + q;
      const int c = S > 1 ? season_of(L.T + t, S, r) : 0;
      const f32x4 e = philox_normal4_fast((uint32_t)si, (uint32_t)t, 0u, TAG_FORECAST, seed);
      const int c2 = S2 > 1 ? season_of(L.T + t, S2, L.r2) : 0;
      const float yv = x[0] + (S > 1 ? x[so + c] : 0.f) + (S2 > 1 ? x[so2 + c2] : 0.f) + off[(long)t * B + b] + sobs * e.x;
      tile[threadIdx.x][q] = fmaf(yv, sg, mu);
      if (so == 2) {                   // level += slope ; slope random walk
        x[0] += x[1];
        x[1] = fmaf(sslope, e.w, x[1]);
      }
      x[0] = fmaf(sl, e.y, x[0]);
      if (season_change(L.T + t, S, r)) {
        const float xi = sdr * e.z;
        for (int a = so; a < so2; ++a) x[a] += xi;
        x[so + c] -= (float)S * xi;    // net -(S-1) xi on the season just observed
      }
      if (season_change(L.T + t, S2, L.r2)) {   // second seasonal component: its own drift draw
        const float xi = sdr2 * philox_normal4_fast((uint32_t)si, (uint32_t)t, 1u, TAG_FORECAST, seed).x;
        for (int a = so2; a < D; ++a) x[a] += xi;
        x[so2 + c2] -= (float)S2 * xi;
      }
    }
    __syncthreads();
    for (int row = warp; row < 128; row += 4) {
      const int ii = blockIdx.x * 128 + row;
      if (ii < nsamp && lane < tn) sims[((long)n * nsamp + ii) * Tf + t0 + lane] = tile[row][lane];
    }
    __syncthreads();
  }
}

// ---- component decomposition on the warp-per-lane path (ci_decompose.cu dispatches here) ---------------
// Smoother in the Durbin-Koopman backward form, which needs no per-step matrix inverse and only the
// columns of the PREDICTED covariance that belong to the observed entries:
//     r_{t-1} = H v/s + L^T r_t,   N_{t-1} = H H^T / s + L^T N_t L,   L = F (I - K H^T),  K = P H / s
//     E[x_t | y]   = m_t + P_t r_{t-1},   Var[x_t | y] = P_t - P_t N_{t-1} P_t
// The forward pass keeps, per step, the columns P_t[:, c] and means m_t[c] for c in {level, current season
// (, current season of the second component)} plus (v, 1/s): hist[B][T][3 D + 5].  N lives in shared memory
// (it re-uses the covariance buffer); L^T N L touches only the rows/columns of H, like the adjoint above.
__host__ __device__ __forceinline__ int big_hist_width(const ci_layout& L) { return 3 * big_D(L) + 5; }

template <int NC, bool TWO>
__global__ void __launch_bounds__(BIG_WARPS * 32)
k_smooth_big(ci_layout L, const float* __restrict__ XY, const float* __restrict__ sconst, const float* __restrict__ u,
             float* __restrict__ hist, float* __restrict__ cmean, float* __restrict__ cvar) {
  extern __shared__ __align__(16) float smem[];
  const int S = L.S, so = big_nt(L), so2 = big_so2(L), S2 = TWO ? L.S2 : 1, D = big_D(L), k = L.k, T = L.T;
  const int HW = big_hist_width(L);
  const long B = (long)L.N * L.G;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long b = (long)blockIdx.x * BIG_WARPS + w;
  if (b >= B) return;
  const int n = (int)(b / L.G);
  const int per_warp = D * D + 8 * D + (k > 0 ? k : 1) + 8;
  float* P = smem + (long)w * per_warp;     // covariance (forward), then N (backward)
  float* m = P + D * D;
  float* g = m + D;                         // g (forward) / K (backward)
  float* rv = g + D;
  float* pc0 = rv + D;
  float* pcJ = pc0 + D;
  float* pc2 = pcJ + D;
  float* beta = pc2 + 3 * D;
  float* scal = beta + (k > 0 ? k : 1);
  if (lane == 0) {
    if (!big_gen(L)) {
      const LaneScalars ls = params_forward_lane(u + b, B, k, S, sconst + n, L.N, beta, 1);
      scal[0] = ls.R; scal[1] = ls.ql; scal[2] = ls.qd; scal[4] = 0.f; scal[5] = 0.f;
    } else {
      const GenScalars ls = params_forward_gen(u + b, B, k, S, L.flags, sconst + n, L.N, beta, 1, L.S2);
      scal[0] = ls.R; scal[1] = ls.ql; scal[2] = ls.qd; scal[4] = ls.qs; scal[5] = ls.qd2;
    }
  }
  __syncwarp();
  const float R = scal[0], ql = scal[1], qd = scal[2], qs = scal[4], qd2 = scal[5];
  big_init(P, m, D, S, so, S2, sconst[SC_M0 * L.N + n], sconst[SC_P0 * L.N + n], sconst[SC_SD * L.N + n], lane);
  const Cols<NC> cl(lane, D);
  float* hb = hist + b * (long)T * HW;
  for (int t = 0; t < T; ++t) {
    const BigSeason z = big_season<TWO>(L, so, so2, t);
    float* he = hb + (long)t * HW;
#pragma unroll
    for (int q = 0; q < NC; ++q) {
      if (cl.ok[q]) {
        const int j = cl.j[q];
        he[j] = P[j * D];
        he[D + j] = S > 1 ? P[j * D + z.J] : 0.f;
        he[2 * D + j] = TWO ? P[j * D + z.J2] : 0.f;
      }
    }
    if (lane == 0) {
      he[3 * D] = m[0];
      he[3 * D + 1] = S > 1 ? m[z.J] : 0.f;
      he[3 * D + 2] = TWO ? m[z.J2] : 0.f;
    }
    const float resid = big_resid(XY, L, n, t, beta, lane);
    float s, v, is;
    const bool obs = big_step<NC, 8, TWO>(P, m, g, D, S, so, S2, so2, z, resid, R, ql, qs, qd, qd2, s, v, is, lane);
    if (lane == 0) {
      he[3 * D + 3] = obs ? v : __int_as_float(0x7fc00000);
      he[3 * D + 4] = is;
    }
    __syncwarp();
  }
  // ---- backward recursion --------------------------------------------------------------------------
  float* Nm = P;
  for (int i = lane; i < D * D; i += 32) Nm[i] = 0.f;
  for (int j = lane; j < D; j += 32) rv[j] = 0.f;
  __syncwarp();
  for (int t = T - 1; t >= 0; --t) {
    const BigSeason z = big_season<TWO>(L, so, so2, t);
    const int J = z.J, J2 = z.J2;
    const float* he = hb + (long)t * HW;
    float p0[NC], pJ[NC], p2[NC];
#pragma unroll
    for (int q = 0; q < NC; ++q) {
      p0[q] = cl.ok[q] ? he[cl.j[q]] : 0.f;
      pJ[q] = cl.ok[q] ? he[D + cl.j[q]] : 0.f;
      p2[q] = cl.ok[q] ? he[2 * D + cl.j[q]] : 0.f;
      if (cl.ok[q]) {
        pc0[cl.j[q]] = p0[q];
        pcJ[cl.j[q]] = pJ[q];
        pc2[cl.j[q]] = p2[q];
      }
    }
    const float m0 = he[3 * D], mJ = he[3 * D + 1], m2 = he[3 * D + 2];
    const float v = he[3 * D + 3], is = he[3 * D + 4];
    const bool obs = (v == v);
    if (so == 2) {   // r <- F^T r,  N <- F^T N F   (LocalLinearTrend transition)
#pragma unroll
      for (int q = 0; q < NC; ++q)
        if (cl.ok[q]) Nm[D + cl.j[q]] += Nm[cl.j[q]];            // row 1 += row 0
      __syncwarp();
#pragma unroll
      for (int q = 0; q < NC; ++q)
        if (cl.ok[q]) Nm[cl.j[q] * D + 1] += Nm[cl.j[q] * D];    // column 1 += column 0
      if (lane == 0) rv[1] += rv[0];
    }
    __syncwarp();
    if (obs) {
      float Kj[NC], aj[NC];
#pragma unroll
      for (int q = 0; q < NC; ++q) {
        Kj[q] = (p0[q] + pJ[q] + p2[q]) * is;
        aj[q] = 0.f;
        if (cl.ok[q]) g[cl.j[q]] = Kj[q];
      }
      __syncwarp();
#pragma unroll 4
      for (int i = 0; i < D; ++i) {
        const float ki = g[i];
#pragma unroll
        for (int q = 0; q < NC; ++q) aj[q] = fmaf(Nm[i * D + cl.j[q]], ki, aj[q]);
      }
      float pk = 0.f, pr = 0.f;
#pragma unroll
      for (int q = 0; q < NC; ++q) {
        if (cl.ok[q]) {
          pk = fmaf(Kj[q], aj[q], pk);
          pr = fmaf(Kj[q], rv[cl.j[q]], pr);
        }
      }
      const float wgt = warp_sum(pk) + is;
      const float dr = v * is - warp_sum(pr);
      __syncwarp();
      // N <- N - H a^T - a H^T + wgt H H^T : rows of H, then columns of H, then the H x H block
#pragma unroll
      for (int q = 0; q < NC; ++q) {
        if (cl.ok[q]) {
          Nm[cl.j[q]] -= aj[q];
          if (S > 1) Nm[J * D + cl.j[q]] -= aj[q];
          if (TWO) Nm[J2 * D + cl.j[q]] -= aj[q];
        }
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < NC; ++q) {
        if (cl.ok[q]) {
          Nm[cl.j[q] * D] -= aj[q];
          if (S > 1) Nm[cl.j[q] * D + J] -= aj[q];
          if (TWO) Nm[cl.j[q] * D + J2] -= aj[q];
        }
      }
      __syncwarp();
      if (lane == 0) {
        int hs[3] = {0, J, J2};
        const int nh = 1 + (S > 1 ? 1 : 0) + (TWO ? 1 : 0);
        for (int c1 = 0; c1 < nh; ++c1)
          for (int c2 = 0; c2 < nh; ++c2) Nm[hs[c1] * D + hs[c2]] += wgt;
        for (int c1 = 0; c1 < nh; ++c1) rv[hs[c1]] += dr;
      }
      __syncwarp();
    }
    // smoothed marginals of the observed entries: mean = m_c + p_c . r,  var = p_c[c] - p_c^T N p_c
    float q0[NC], qJ[NC], q2[NC];
#pragma unroll
    for (int q = 0; q < NC; ++q) q0[q] = qJ[q] = q2[q] = 0.f;
#pragma unroll 4
    for (int i = 0; i < D; ++i) {
      const float a0 = pc0[i], aJ = pcJ[i], a2 = pc2[i];
#pragma unroll
      for (int q = 0; q < NC; ++q) {
        const float nv = Nm[i * D + cl.j[q]];
        q0[q] = fmaf(nv, a0, q0[q]);
        qJ[q] = fmaf(nv, aJ, qJ[q]);
        if (TWO) q2[q] = fmaf(nv, a2, q2[q]);
      }
    }
    float e0 = 0.f, eJ = 0.f, e2 = 0.f, f0 = 0.f, fJ = 0.f, f2 = 0.f;
#pragma unroll
    for (int q = 0; q < NC; ++q) {
      if (cl.ok[q]) {
        const float rj = rv[cl.j[q]];
        e0 = fmaf(p0[q], rj, e0); eJ = fmaf(pJ[q], rj, eJ); e2 = fmaf(p2[q], rj, e2);
        f0 = fmaf(p0[q], q0[q], f0); fJ = fmaf(pJ[q], qJ[q], fJ); f2 = fmaf(p2[q], q2[q], f2);
      }
    }
    e0 = warp_sum(e0); eJ = warp_sum(eJ); e2 = warp_sum(e2);
    f0 = warp_sum(f0); fJ = warp_sum(fJ); f2 = warp_sum(f2);
    if (lane == 0) {
      cmean[((long)0 * T + t) * B + b] = m0 + e0;
      cvar[((long)0 * T + t) * B + b] = fmaxf(pc0[0] - f0, 0.f);
      cmean[((long)1 * T + t) * B + b] = S > 1 ? mJ + eJ : 0.f;
      cvar[((long)1 * T + t) * B + b] = S > 1 ? fmaxf(pcJ[J] - fJ, 0.f) : 0.f;
      cmean[((long)2 * T + t) * B + b] = TWO ? m2 + e2 : 0.f;
      cvar[((long)2 * T + t) * B + b] = TWO ? fmaxf(pc2[J2] - f2, 0.f) : 0.f;
    }
    __syncwarp();
  }
}

// Component moments over the forecast horizon: prior propagation from the forecast state (m, P).
template <int NC, bool TWO>
__global__ void __launch_bounds__(BIG_WARPS * 32)
k_forecast_moments_big(ci_layout L, const float* __restrict__ fstate, const float* __restrict__ scal_g,
                       float* __restrict__ cmean, float* __restrict__ cvar) {
  extern __shared__ __align__(16) float smem[];
  const int S = L.S, so = big_nt(L), so2 = big_so2(L), S2 = TWO ? L.S2 : 1, D = big_D(L), Tf = L.Tf;
  const long B = (long)L.N * L.G;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long b = (long)blockIdx.x * BIG_WARPS + w;
  if (b >= B) return;
  const int per_warp = D * D + 8 * D + (L.k > 0 ? L.k : 1) + 8;
  float* P = smem + (long)w * per_warp;
  float* m = P + D * D;
  float* g = m + D;
  const float* fs = fstate + b * (long)(D + 2 * D * D);
  for (int j = lane; j < D; j += 32) m[j] = fs[j];
  for (int i = lane; i < D * D; i += 32) P[i] = fs[D + i];
  __syncwarp();
  const float R = scal_g[0 * B + b], ql = scal_g[1 * B + b], qd = scal_g[2 * B + b], qs = scal_g[4 * B + b],
              qd2 = scal_g[5 * B + b];
  for (int t = 0; t < Tf; ++t) {
    const BigSeason z = big_season<TWO>(L, so, so2, L.T + t);
    if (lane == 0) {
      cmean[((long)0 * Tf + t) * B + b] = m[0];
      cvar[((long)0 * Tf + t) * B + b] = P[0];
      cmean[((long)1 * Tf + t) * B + b] = S > 1 ? m[z.J] : 0.f;
      cvar[((long)1 * Tf + t) * B + b] = S > 1 ? P[z.J * D + z.J] : 0.f;
      cmean[((long)2 * Tf + t) * B + b] = TWO ? m[z.J2] : 0.f;
      cvar[((long)2 * Tf + t) * B + b] = TWO ? P[z.J2 * D + z.J2] : 0.f;
    }
    __syncwarp();
    float s, v, is;
    big_step<NC, 8, TWO>(P, m, g, D, S, so, S2, so2, z, __int_as_float(0x7fc00000), R, ql, qs, qd, qd2, s, v, is, lane);
  }
}

size_t big_smooth_smem_bytes(const ci_layout* L) {
  const int D = big_D(*L);
  return (size_t)BIG_WARPS * (D * D + 8 * D + (L->k > 0 ? L->k : 1) + 8) * sizeof(float);
}
size_t big_hist_floats(const ci_layout* L) {
  return (size_t)L->N * L->G * L->T * big_hist_width(*L);
}

// ---- CTA-per-lane evaluation: few lanes, long series (BASELINE configs[4]: 100 series x 10,000 steps) ----
// With B << number of SMs a warp per lane leaves the machine empty and every shared-memory round trip of
// the row sweeps is exposed.  Here one CTA of 64 x RG threads owns a lane: thread (tx, ty) holds column
// tx of rows ty, ty + RG, ... of P and of its adjoint, so a sweep is D / RG rows deep instead of D.
// Same arithmetic as k_eval_big (same tape layout), block barriers instead of __syncwarp.
// The regression is hoisted out of the time loop: a parallel pre-pass writes the residuals r_t = y_t - mean -
// X_t beta, the reverse sweep stores rbar_t, and a parallel post-pass forms beta_bar = -sum_t rbar_t X_t.
constexpr int CTA_COLS = 64;

template <int RG, bool TWO>
__global__ void __launch_bounds__(CTA_COLS * RG)
k_eval_cta(ci_layout L, const float* __restrict__ XY, const float* __restrict__ sconst, const float* __restrict__ u,
           float* __restrict__ logp, float* __restrict__ grad, float* __restrict__ tape, float* __restrict__ rscr) {
  extern __shared__ __align__(16) float smem[];
  constexpr int NT = CTA_COLS * RG;
  const int S = L.S, so = big_nt(L), so2 = big_so2(L), S2 = TWO ? L.S2 : 1, D = big_D(L), k = L.k, T = L.T;
  const long B = (long)L.N * L.G;
  const int tid = threadIdx.x, tx = tid & (CTA_COLS - 1), ty = tid >> 6;
  const long b = blockIdx.x;
  const int n = (int)(b / L.G);
  const bool col_ok = tx < D;
  const int kk = k > 0 ? k : 1;
  float* P = smem;
  float* Pb = P + D * D;
  float* m = Pb + D * D;
  float* g = m + D;
  float* mb = g + D;
  float* part = mb + D;                       // [3][RG][64] partial column sums of the reverse sweep
  float* red = part + 3 * RG * CTA_COLS;      // [2][8] block-reduction scratch
  float* beta = red + 16;
  float* scal = beta + kk;
  if (tid == 0) {
    if (!big_gen(L)) {
      const LaneScalars ls = params_forward_lane(u + b, B, k, S, sconst + n, L.N, beta, 1);
      scal[0] = ls.R; scal[1] = ls.ql; scal[2] = ls.qd; scal[3] = ls.lp0; scal[4] = 0.f; scal[5] = 0.f;
    } else {
      const GenScalars ls = params_forward_gen(u + b, B, k, S, L.flags, sconst + n, L.N, beta, 1, L.S2);
      scal[0] = ls.R; scal[1] = ls.ql; scal[2] = ls.qd; scal[3] = ls.lp0; scal[4] = ls.qs; scal[5] = ls.qd2;
    }
  }
  __syncthreads();
  const float R = scal[0], ql = scal[1], qd = scal[2], lp0 = scal[3], qs = scal[4], qd2 = scal[5];
  const int stride = (k + 1) * CI_TC;
  float* rs = rscr + b * (long)(2 * T);       // [T] residuals, then [T] rbar
  // Per-thread constants of the sweeps: this thread's rows i = ty + RG q (q < NQ) and column tx; membership of
  // the seasonal blocks as 0/1 weights, so that the drift direction of a step is ONE compare + select per row
  // (w_i = in-block ? (i == J ? 1 - S : 1) : 0).
  constexpr int NQ = CTA_COLS / RG;
  float in1[NQ], in2[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int i = ty + RG * q;
    in1[q] = (S > 1 && i >= so && i < so + S && i < D) ? 1.f : 0.f;
    in2[q] = (TWO && i >= so2 && i < so2 + S2 && i < D) ? 1.f : 0.f;
  }
  const float in1x = (S > 1 && tx >= so && tx < so + S) ? 1.f : 0.f;
  const float in2x = (TWO && tx >= so2 && tx < so2 + S2) ? 1.f : 0.f;
  const float wJ = (float)(1 - S), wJ2 = (float)(1 - S2);
  // ---- pre-pass: residuals ---------------------------------------------------------------------------
  for (int t = tid; t < T; t += NT) {
    const float* xp = XY + ((long)(t / CI_TC) * L.N + n) * stride + (t % CI_TC);
    float acc = 0.f;
    for (int j = 0; j < k; ++j) acc = fmaf(beta[j], __ldg(xp + j * CI_TC), acc);
    rs[t] = __ldg(xp + k * CI_TC) - acc;
  }
  // ---- initial state ---------------------------------------------------------------------------------
  {
    const float m0 = sconst[SC_M0 * L.N + n], p0 = sconst[SC_P0 * L.N + n], sd = sconst[SC_SD * L.N + n];
    const float off = S > 1 ? -p0 / (float)S : 0.f, off2 = S2 > 1 ? -p0 / (float)S2 : 0.f;
    for (int e = tid; e < D * D; e += NT) {
      const int i = e / D, j = e % D;
      float v = 0.f;
      if (i < so || j < so) {
        if (i == j) v = (i == 0) ? p0 : sd * sd;
      } else if (i < so2 && j < so2) {
        v = (i == j) ? p0 + off : off;
      } else if (i >= so2 && j >= so2) {
        v = (i == j) ? p0 + off2 : off2;
      }
      P[e] = v;
      Pb[e] = 0.f;
    }
    for (int j = tid; j < D; j += NT) {
      m[j] = (j == 0) ? m0 : 0.f;
      mb[j] = 0.f;
    }
  }
  __syncthreads();
  float* tp = tape + b * (long)T * (D + 1);
  float acc = 0.f, quad = 0.f;      // thread 0: sums of log s_t and v_t^2 / s_t
  int nobs = 0;
  float r_next = rs[0];
  SeasonWalk<TWO> walk(L, so, so2, 0);
  for (int t = 0; t < T; ++t, walk.next()) {
    const BigSeason z = walk.get();
    const int J = z.J, J2 = z.J2;
    const float resid = r_next;
    if (t + 1 < T) r_next = rs[t + 1];
    const bool obs = (resid == resid);
    // phase A: g = P H^T (column reads), H m
    float gj = 0.f;
    if (ty == 0 && col_ok) {
      gj = P[tx * D] + (S > 1 ? P[tx * D + J] : 0.f) + (TWO ? P[tx * D + J2] : 0.f);
      g[tx] = gj;
    }
    const float hm = m[0] + (S > 1 ? m[J] : 0.f) + (TWO ? m[J2] : 0.f);
    __syncthreads();
    const float s = g[0] + (S > 1 ? g[J] : 0.f) + (TWO ? g[J2] : 0.f) + R;
    const float v = resid - hm;
    const float is = rcp(s);
    const float nis = obs ? -is : 0.f;
    const float qq = z.change ? qd : 0.f, qq2 = z.change2 ? qd2 : 0.f;
    // phase C: mean update, covariance sweep (measurement + time update), tape
    if (col_ok) {
      if (ty == 0 && obs) m[tx] = fmaf(gj, v * is, m[tx]);
      const float gx = g[tx];
      const float qwx = qq * (tx == J ? wJ : in1x), qw2x = TWO ? qq2 * (tx == J2 ? wJ2 : in2x) : 0.f;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int i = ty + RG * q;
        if (i < D) {
          float nv = fmaf(i == J ? wJ : in1[q], qwx, fmaf(g[i] * nis, gx, P[i * D + tx]));
          if (TWO) nv = fmaf(i == J2 ? wJ2 : in2[q], qw2x, nv);
          P[i * D + tx] = nv;
        }
      }
      if (so != 2 && tid == 0) P[0] += ql;     // level variance (LocalLinearTrend: after F P F^T below)
      if (ty == RG - 1) tp[(long)t * (D + 1) + tx] = gx;
    }
    if (tid == NT - 1) tp[(long)t * (D + 1) + D] = obs ? v : __int_as_float(0x7fc00000);
    if (tid == 0 && obs) {
      acc += logf(s);
      quad = fmaf(v * is, v, quad);
      ++nobs;
    }
    __syncthreads();
    if (so == 2) {   // LocalLinearTrend transition: P <- F P F^T, m <- F m
      if (ty == 0 && col_ok) P[tx] += P[D + tx];
      __syncthreads();
      if (ty == 0 && col_ok) P[tx * D] += P[tx * D + 1];
      if (tid == 0) m[0] += m[1];
      __syncthreads();
      if (tid == 0) {
        P[D + 1] += qs;
        P[0] += ql;
      }
      __syncthreads();
    }
  }
  const float ll = -0.5f * (acc + quad + (float)nobs * CI_LOG_2PI);

  // ---- reverse sweep -----------------------------------------------------------------------------------
  float Rb = 0.f, qlb = 0.f, qsb = 0.f, qdb = 0.f, qd2b = 0.f;     // thread 0
  float cur_g = col_ok ? tp[(long)(T - 1) * (D + 1) + tx] : 0.f, cur_v = tp[(long)(T - 1) * (D + 1) + D];
  walk = SeasonWalk<TWO>(L, so, so2, T - 1);
  for (int t = T - 1; t >= 0; --t, walk.prev()) {
    const BigSeason z = walk.get();
    const int J = z.J, J2 = z.J2;
    float nxt_g = 0.f, nxt_v = 0.f;
    if (t > 0) {
      nxt_g = col_ok ? tp[(long)(t - 1) * (D + 1) + tx] : 0.f;
      nxt_v = tp[(long)(t - 1) * (D + 1) + D];
    }
    const float v = cur_v;
    const bool obs = (v == v);
    const float gx = obs ? cur_g : 0.f;
    if (ty == 0 && col_ok) g[tx] = gx;
    if (tid == 0) qlb += Pb[0];
    if (so == 2) {   // adjoint of the LocalLinearTrend transition
      if (tid == 0) qsb += Pb[D + 1];
      __syncthreads();
      if (ty == 0 && col_ok) Pb[D + tx] += Pb[tx];
      __syncthreads();
      if (ty == 0 && col_ok) Pb[tx * D + 1] += Pb[tx * D];
      if (tid == 0) mb[1] += mb[0];
    }
    __syncthreads();
    // partial column sums over this thread's rows: Pbar g and Pbar w (drift adjoints)
    {
      float c0 = 0.f, c1 = 0.f, c2 = 0.f;
      if (col_ok && (obs || z.change || z.change2)) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int i = ty + RG * q;
          if (i < D) {
            const float pb = Pb[i * D + tx];
            c0 = fmaf(pb, g[i], c0);
            c1 = fmaf(pb, i == J ? wJ : in1[q], c1);
            if (TWO) c2 = fmaf(pb, i == J2 ? wJ2 : in2[q], c2);
          }
        }
      }
      part[(0 * RG + ty) * CTA_COLS + tx] = c0;
      part[(1 * RG + ty) * CTA_COLS + tx] = c1;
      part[(2 * RG + ty) * CTA_COLS + tx] = c2;
    }
    __syncthreads();
    float col = 0.f;
    if (ty == 0) {    // warps 0 and 1: finish the column sums and reduce the four scalars
      float nac = 0.f, nac2 = 0.f;
#pragma unroll
      for (int q = 0; q < RG; ++q) {
        col += part[(0 * RG + q) * CTA_COLS + tx];
        nac += part[(1 * RG + q) * CTA_COLS + tx];
        nac2 += part[(2 * RG + q) * CTA_COLS + tx];
      }
      const float wx = col_ok ? (tx == J ? wJ : in1x) : 0.f;
      const float w2x = (TWO && col_ok) ? (tx == J2 ? wJ2 : in2x) : 0.f;
      const float pa = col_ok ? mb[tx] * gx : 0.f;
      const float pg = gx * col;
      const float s0 = warp_sum(pa), s1 = warp_sum(pg), s2 = warp_sum(wx * nac), s3 = warp_sum(w2x * nac2);
      if ((tid & 31) == 0) {
        float* rr = red + (tid >> 5) * 4;
        rr[0] = s0; rr[1] = s1; rr[2] = s2; rr[3] = s3;
      }
    }
    __syncthreads();
    const float a = red[0] + red[4], gPg = red[1] + red[5];
    if (tid == 0) {
      if (z.change) qdb += red[2] + red[6];
      if (TWO && z.change2) qd2b += red[3] + red[7];
    }
    float gb = 0.f, vb = 0.f;
    if (obs) {
      const float s = g[0] + (S > 1 ? g[J] : 0.f) + (TWO ? g[J2] : 0.f) + R;
      const float is = rcp(s);
      const float vis = v * is;
      vb = (a - v) * is;
      const float sbar = (gPg - a * v) * is * is - 0.5f * (is - vis * vis);
      if (tid == 0) Rb += sbar;
      if (ty == 0 && col_ok) {
        const bool inH = tx == 0 || (S > 1 && tx == J) || (TWO && tx == J2);
        gb = fmaf(mb[tx], vis, fmaf(-2.f * is, col, inH ? sbar : 0.f));
        const float h = 0.5f * gb;     // Pbar += sym(gb H^T): level row / column first
        Pb[tx * D] += h;
        Pb[tx] += h;
      }
    }
    if (tid == 0) rs[T + t] = obs ? vb : 0.f;
    __syncthreads();
    if (obs) {
      if (S > 1 && ty == 0 && col_ok) {
        const float h = 0.5f * gb;
        Pb[tx * D + J] += h;
        Pb[J * D + tx] += h;
      }
      if (tid == 0) {
        mb[0] -= vb;
        if (S > 1) mb[J] -= vb;
        if (TWO) mb[J2] -= vb;
      }
    }
    if (TWO) {
      __syncthreads();
      if (obs && ty == 0 && col_ok) {
        const float h = 0.5f * gb;
        Pb[tx * D + J2] += h;
        Pb[J2 * D + tx] += h;
      }
    }
    cur_g = nxt_g;
    cur_v = nxt_v;
    __syncthreads();
  }
  // ---- post-pass: beta_bar_j = -sum_t rbar_t X_tj  (threads = covariates x time slices) -----------------
  if (k > 0) {
    const int slices = NT / kk > 0 ? NT / kk : 1;
    float* pb = part;   // [slices][k] partial sums (3 RG 64 floats >= NT available)
    const int j = tid % kk, sl = tid / kk;
    if (sl < slices) {
      float accb = 0.f;
      for (int t = sl; t < T; t += slices) {
        const float* xp = XY + ((long)(t / CI_TC) * L.N + n) * stride + (t % CI_TC);
        accb = fmaf(-rs[T + t], __ldg(xp + j * CI_TC), accb);
      }
      pb[sl * kk + j] = accb;
    }
    __syncthreads();
    if (tid < k) {
      float accb = 0.f;
      for (int q = 0; q < slices; ++q) accb += pb[q * kk + tid];
      beta[tid] = accb;
    }
    __syncthreads();
  }
  if (tid == 0) {
    if (!big_gen(L)) params_backward_lane(u + b, B, k, S, sconst + n, L.N, beta, 1, Rb, qlb, qdb, grad + b);
    else params_backward_gen(u + b, B, k, S, L.flags, sconst + n, L.N, beta, 1, Rb, qlb, qsb, qdb, grad + b, L.S2, qd2b);
    logp[b] = ll + lp0;
  }
}

template <int RG>
static size_t cta_smem_bytes(const ci_layout* L) {
  const int D = big_D(*L);
  return (size_t)(2 * D * D + 3 * D + 3 * RG * CTA_COLS + 16 + (L->k > 0 ? L->k : 1) + 8) * sizeof(float);
}

// CTA-per-lane pays when the lanes cannot fill the machine with warps
bool big_use_cta(const ci_layout* L) {
  const long B = (long)L->N * L->G;
  return big_D(*L) <= CTA_COLS && big_D(*L) >= 8 && B <= 2 * 148 && L->k <= 256 &&
         cta_smem_bytes<8>(L) <= 227 * 1024;
}

size_t big_smem_bytes(const ci_layout* L) {
  const int D = big_D(*L);
  return (size_t)BIG_WARPS * (2 * D * D + 4 * D + (L->k > 0 ? L->k : 1) + 8) * sizeof(float);
}

bool big_supported(const ci_layout* L) {
  const int D = big_D(*L);
  return D <= CI_BIG_MAXD && big_smem_bytes(L) <= 227 * 1024;
}

size_t big_tape_floats(const ci_layout* L) {
  const size_t D = big_D(*L);
  return (size_t)L->N * L->G * L->T * (D + 1 + 2);   // + residuals / rbar scratch of the CTA-per-lane kernel
}

size_t big_fstate_floats(const ci_layout* L) {
  const size_t D = big_D(*L);
  return D + 2 * D * D;
}

int g_eval_rw = 1;   // ci_set_option("eval_rw", 0 / 1): register-resident 4-warp kernel (ci_eval_rw.cu) on / off

template <class K>
static int big_attr(K kern, size_t smem) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(CI_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  return CI_OK;
}

int launch_eval_big(const ci_layout* L, const ci_data* D, const float* d_u, float* d_logp, float* d_grad,
                    float* tape, cudaStream_t st) {
  if (!big_supported(L)) return fail(CI_ERR_UNSUPPORTED, "nseasons=%d exceeds the shared-memory latent budget", L->S);
  if (g_eval_rw && rw_eligible(L)) {
    const long B = (long)L->N * L->G;
    return launch_eval_rw(L, D, d_u, d_logp, d_grad, tape, tape + (size_t)B * L->T * (big_D(*L) + 1), st);
  }
  if (big_use_cta(L)) {
#ifndef CI_CTA_RG
#define CI_CTA_RG 8
#endif
    constexpr int RG = CI_CTA_RG;
    const size_t smem = cta_smem_bytes<RG>(L);
    const long B = (long)L->N * L->G;
    float* rscr = tape + (size_t)B * L->T * (big_D(*L) + 1);
    if (L->S2 > 1) {
      if (int e = big_attr(k_eval_cta<RG, true>, smem)) return e;
      k_eval_cta<RG, true><<<(unsigned)B, CTA_COLS * RG, smem, st>>>(*L, D->d_X, D->d_sconst, d_u, d_logp, d_grad, tape, rscr);
    } else {
      if (int e = big_attr(k_eval_cta<RG, false>, smem)) return e;
      k_eval_cta<RG, false><<<(unsigned)B, CTA_COLS * RG, smem, st>>>(*L, D->d_X, D->d_sconst, d_u, d_logp, d_grad, tape, rscr);
    }
    return CI_OK;
  }
  const size_t smem = big_smem_bytes(L);
  const long B = (long)L->N * L->G;
  const unsigned grid = (unsigned)((B + BIG_WARPS - 1) / BIG_WARPS);
  const int Dl = big_D(*L);
#define CI_BIG_EVAL(NC)                                                                                           \
  {                                                                                                               \
    if (L->S2 > 1) {                                                                                              \
      if (int e = big_attr(k_eval_big<NC, true>, smem)) return e;                                                 \
      k_eval_big<NC, true><<<grid, BIG_WARPS * 32, smem, st>>>(*L, D->d_X, D->d_sconst, d_u, d_logp, d_grad, tape);  \
    } else {                                                                                                      \
      if (int e = big_attr(k_eval_big<NC, false>, smem)) return e;                                                \
      k_eval_big<NC, false><<<grid, BIG_WARPS * 32, smem, st>>>(*L, D->d_X, D->d_sconst, d_u, d_logp, d_grad, tape); \
    }                                                                                                             \
  }
  if (Dl <= 32) CI_BIG_EVAL(1) else if (Dl <= 64) CI_BIG_EVAL(2) else CI_BIG_EVAL(3)
#undef CI_BIG_EVAL
  return CI_OK;
}

int launch_filter_big(const ci_layout* L, const ci_data* D, const float* d_u, float* pm, float* pv, float* fs,
                      cudaStream_t st) {
  if (!big_supported(L)) return fail(CI_ERR_UNSUPPORTED, "nseasons=%d exceeds the shared-memory latent budget", L->S);
  const size_t smem = big_smem_bytes(L);
  const long B = (long)L->N * L->G;
  const unsigned grid = (unsigned)((B + BIG_WARPS - 1) / BIG_WARPS);
  const int Dl = big_D(*L);
#define CI_BIG_FILTER(NC)                                                                                  \
  {                                                                                                        \
    if (L->S2 > 1) {                                                                                       \
      if (int e = big_attr(k_filter_big<NC, true>, smem)) return e;                                        \
      k_filter_big<NC, true><<<grid, BIG_WARPS * 32, smem, st>>>(*L, D->d_X, D->d_sconst, d_u, pm, pv, fs);  \
    } else {                                                                                               \
      if (int e = big_attr(k_filter_big<NC, false>, smem)) return e;                                       \
      k_filter_big<NC, false><<<grid, BIG_WARPS * 32, smem, st>>>(*L, D->d_X, D->d_sconst, d_u, pm, pv, fs); \
    }                                                                                                      \
  }
  if (Dl <= 32) CI_BIG_FILTER(1) else if (Dl <= 64) CI_BIG_FILTER(2) else CI_BIG_FILTER(3)
#undef CI_BIG_FILTER
  return CI_OK;
}

void launch_forecast_big(const ci_layout* L, const float* fs, const float* scal, const float* off, float* pm,
                         const float* mu_sig, int nsamp, uint64_t seed, float* sims, cudaStream_t st) {
  const long B = (long)L->N * L->G;
  const long tot = B * L->Tf;
  k_forecast_mean_big<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(*L, fs, off, pm);
  if (sims)
    k_forecast_sample_big<<<dim3((nsamp + 127) / 128, L->N), 128, 0, st>>>(*L, nsamp, seed, fs, scal, off, mu_sig, sims);
}

int launch_smooth_big(const ci_layout* L, const ci_data* D, const float* d_u, float* hist, float* cmean, float* cvar,
                      cudaStream_t st) {
  if (!big_supported(L)) return fail(CI_ERR_UNSUPPORTED, "nseasons=%d exceeds the shared-memory latent budget", L->S);
  const size_t smem = big_smooth_smem_bytes(L);
  const long B = (long)L->N * L->G;
  const unsigned grid = (unsigned)((B + BIG_WARPS - 1) / BIG_WARPS);
  const int Dl = big_D(*L);
#define CI_BIG_SMOOTH(NC)                                                                                    \
  {                                                                                                          \
    if (L->S2 > 1) {                                                                                         \
      if (int e = big_attr(k_smooth_big<NC, true>, smem)) return e;                                          \
      k_smooth_big<NC, true><<<grid, BIG_WARPS * 32, smem, st>>>(*L, D->d_X, D->d_sconst, d_u, hist, cmean, cvar);  \
    } else {                                                                                                 \
      if (int e = big_attr(k_smooth_big<NC, false>, smem)) return e;                                         \
      k_smooth_big<NC, false><<<grid, BIG_WARPS * 32, smem, st>>>(*L, D->d_X, D->d_sconst, d_u, hist, cmean, cvar); \
    }                                                                                                        \
  }
  if (Dl <= 32) CI_BIG_SMOOTH(1) else if (Dl <= 64) CI_BIG_SMOOTH(2) else CI_BIG_SMOOTH(3)
#undef CI_BIG_SMOOTH
  return CI_OK;
}

int launch_forecast_moments_big(const ci_layout* L, const float* fs, const float* scal, float* cmean, float* cvar,
                                cudaStream_t st) {
  if (!big_supported(L)) return fail(CI_ERR_UNSUPPORTED, "nseasons=%d exceeds the shared-memory latent budget", L->S);
  const size_t smem = big_smooth_smem_bytes(L);
  const long B = (long)L->N * L->G;
  const unsigned grid = (unsigned)((B + BIG_WARPS - 1) / BIG_WARPS);
  const int Dl = big_D(*L);
#define CI_BIG_FMOM(NC)                                                                                  \
  {                                                                                                      \
    if (L->S2 > 1) {                                                                                     \
      if (int e = big_attr(k_forecast_moments_big<NC, true>, smem)) return e;                            \
      k_forecast_moments_big<NC, true><<<grid, BIG_WARPS * 32, smem, st>>>(*L, fs, scal, cmean, cvar);     \
    } else {                                                                                             \
      if (int e = big_attr(k_forecast_moments_big<NC, false>, smem)) return e;                           \
      k_forecast_moments_big<NC, false><<<grid, BIG_WARPS * 32, smem, st>>>(*L, fs, scal, cmean, cvar);    \
    }                                                                                                    \
  }
  if (Dl <= 32) CI_BIG_FMOM(1) else if (Dl <= 64) CI_BIG_FMOM(2) else CI_BIG_FMOM(3)
#undef CI_BIG_FMOM
  return CI_OK;
}

}  // namespace ci
