// One complete joint log-prob + gradient evaluation for one (series, chain[, sample]) lane:
//   u -> (sigma^2's, beta, prior + log-det)                                     [K3 forward]
//   r_t = y_t - mean - X_t . beta, Kalman filter, tape (g_t, v_t)                [K2 + K1 forward]
//   reverse-time adjoint sweep -> dL/dr_t -> beta_bar = -sum_t rbar_t X_t        [K1 + K2 reverse]
//   VJP through horseshoe / softplus / sqrt + prior gradient                     [K3 reverse]
// in ONE pass per direction over the series.  Nothing but the tape leaves the thread: beta /
// beta_bar live in a per-lane scratch column (shared memory on the device), the CI_TC residuals of
// the current time chunk in a second one, mean / covariance / adjoints in registers.
//
// Design matrix layout ("chunk-major"): XY[tc][n][j][CI_TC], tc = t / CI_TC, j = 0..k with row k
// holding the centred observed series y - mean(y) (NaN = missing) -- the k covariates and y of CI_TC consecutive
// steps of one series are a contiguous 16(k+1)-byte run and consecutive series follow each other, so
// the run a warp needs for a chunk (its 32/G series) is ONE contiguous block, which the CUDA kernel
// stages into shared memory with one TMA bulk copy while the previous chunk's filter steps run.
//
// Tape entry per step: g_0..g_{S-1}, v  (s = g_0 + g_1 + R is recomputed; v = NaN marks a masked
// step).
//
// The function is templated on an X-source policy (`acquire` / `chunk` / `release`) so that the
// TEST-ONLY host shim (tests/host_shim) runs exactly this arithmetic on the CPU against the oracle,
// and the kernel can choose between TMA-staged shared memory and direct global loads.
// Math: SURVEY.md Appendix A.2-A.5.
#pragma once
#include <type_traits>
#include "ci_core.cuh"
#include "ci_kalman.cuh"
#include "ci_params.cuh"

#define CI_TC 4          // time steps per chunk
#define CI_TAPE_AHEAD 8  // reverse sweep: L2 prefetch distance of the tape, in steps

namespace ci {

struct f4 {
  float x, y, z, w;
};

// Tape reads: every entry is read exactly once, after an L2 prefetch CI_TAPE_AHEAD steps earlier -> streaming
// (evict-first) loads, so the consumed lines leave L2 ahead of the design chunks and the tape still to come: measured
// 0.8233 -> 0.8152 ms per evaluation and 13.21 -> 13.06 ms per HMC transition at configs[3] (scripts/ab_eval.py, three
// rounds; the same hint on the tape STORES changed nothing and is not used).
template <bool STREAM>
CI_HD float tape_load(const float* p) {
#if defined(__CUDA_ARCH__)
  if constexpr (STREAM) return __ldcs(p);   // ld.global.cs: the pointer must be a global address
  else return *p;
#else
  return *p;
#endif
}
// an X-source policy that declares `using tape_in_shared = void;` hands the reverse sweep tape entries staged in shared
// memory (ci_eval_ws.cu): plain generic loads there
template <class XS, class = void>
struct tape_streams { static constexpr bool value = true; };
template <class XS>
struct tape_streams<XS, std::void_t<typename XS::tape_in_shared>> { static constexpr bool value = false; };

CI_HD void prefetch_l2(const float* p) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}


// Direct policy: the lane reads its series' chunks straight from (global) memory.
struct XDirect {
  const float* base;  // XY[0][n]
  long cstride;       // floats between consecutive chunks of a series = N * (k + 1) * CI_TC
  CI_HD void acquire() {}
  CI_HD void release(int /*next_tc*/) {}
  CI_HD constexpr int rs() const { return CI_TC; }     // floats between consecutive rows of a chunk
  CI_HD const float* chunk(int tc) const { return base + (long)tc * cstride; }
  CI_HD f4 load(const float* p) const {
#if defined(__CUDA_ARCH__)
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    return f4{v.x, v.y, v.z, v.w};
#else
    return f4{p[0], p[1], p[2], p[3]};
#endif
  }
};

// sb   : per-lane scratch column for beta / beta_bar, element j at sb[j * sbld]
// sr   : per-lane scratch for the CI_TC residuals of a chunk (sr[0..3], 16-byte aligned on device)
// tape : [T][S+1] with element stride ldt
// u, grad : [P] with element stride ld
template <int S, class XS>
CI_HD void eval_lane(const float* u, long ld, const float* __restrict__ sc, long scld, XS& xs, int T,
                     int k, int r_dur, float* __restrict__ sb, int sbld, float* __restrict__ sr,
                     float* __restrict__ tape, long ldt, float* __restrict__ logp_out, float* __restrict__ grad) {
  constexpr int TW = S + 1;
  const LaneScalars ls = params_forward_lane(u, ld, k, S, sc, scld, sb, sbld);
  const float R = ls.R, ql = ls.ql, qd = ls.qd;
  const int nchunk = (T + CI_TC - 1) / CI_TC;

  // ---- forward: residual chunk (regression GEMV) + CI_TC filter steps --------------------------
  KState<S> ks;
  kf_init<S>(ks, sc[SC_M0 * scld], sc[SC_P0 * scld]);
  float acc = 0.f, comp = 0.f;  // compensated sum of log s + v^2/s
  int nobs = 0;
  int phase = 0;
  for (int tc = 0; tc < nchunk; ++tc) {
    const int t0 = tc * CI_TC;
    const int nstep = (T - t0) < CI_TC ? (T - t0) : CI_TC;
    xs.acquire();
    {
      const float* __restrict__ xp = xs.chunk(tc);
      const f4 yv = xs.load(xp + k * xs.rs());
      float r0 = yv.x, r1 = yv.y, r2 = yv.z, r3 = yv.w;   // row k holds the centred series
#pragma unroll 5
      for (int j = 0; j < k; ++j) {
        const float bj = sb[j * sbld];
        const f4 x = xs.load(xp + j * xs.rs());
        r0 = fmaf(-bj, x.x, r0); r1 = fmaf(-bj, x.y, r1);
        r2 = fmaf(-bj, x.z, r2); r3 = fmaf(-bj, x.w, r3);
      }
      sr[0] = r0; sr[1] = r1; sr[2] = r2; sr[3] = r3;
    }
    xs.release(tc + 1 < nchunk ? tc + 1 : -1);   // chunk tc+1 streams in while the steps below run
    for (int i = 0; i < nstep; ++i) {
      const float r = sr[i];
      float* tp = tape + (long)(t0 + i) * TW * ldt;
      if (r == r) {
        float g[S], s, v, is;
        kf_update<S>(ks, r, R, g, s, v, is);
#pragma unroll
        for (int q = 0; q < S; ++q) tp[(long)q * ldt] = g[q];
        tp[(long)S * ldt] = v;
        const float term = fast_log(s) + v * v * is;
        const float yk = term - comp;
        const float tk = acc + yk;
        comp = (tk - acc) - yk;
        acc = tk;
        ++nobs;
      } else {
#pragma unroll
        for (int q = 0; q < S; ++q) tp[(long)q * ldt] = 0.f;
        tp[(long)S * ldt] = r;  // NaN marks the masked step
      }
      const bool change = (phase == r_dur - 1);
      kf_predict<S>(ks, ql, qd, change);
      phase = change ? 0 : phase + 1;
    }
  }
  const float ll = -0.5f * (acc + (float)nobs * CI_LOG_2PI);

  // ---- reverse: CI_TC adjoint steps + beta_bar accumulation per chunk --------------------------
  for (int j = 0; j < k; ++j) sb[j * sbld] = 0.f;
  KState<S> b;
#pragma unroll
  for (int i = 0; i < S; ++i) b.m[i] = 0.f;
#pragma unroll
  for (int i = 0; i < KState<S>::NP; ++i) b.P[i] = 0.f;
  float Rb = 0.f, qlb = 0.f, qdb = 0.f;
  phase = (T - 1) % r_dur;
  sr[0] = 0.f; sr[1] = 0.f; sr[2] = 0.f; sr[3] = 0.f;
  // software-pipelined tape reads: the entry of step t-1 is requested before step t is processed
  float cur[TW], nxt[TW] = {};
#pragma unroll
  for (int q = 0; q < TW; ++q) cur[q] = tape[((long)(T - 1) * TW + q) * ldt];
  for (int t = T - 1; t >= 0; --t) {
    const int i = t & (CI_TC - 1);
    if (t > 0) {
#pragma unroll
      for (int q = 0; q < TW; ++q) nxt[q] = tape[((long)(t - 1) * TW + q) * ldt];
    }
    const bool change = (phase == r_dur - 1);
    kb_predict_adj<S>(b, qlb, qdb, change);
    phase = (phase == 0) ? r_dur - 1 : phase - 1;
    float rb = 0.f;
    const float v = cur[S];
    if (v == v) {
      float g[S];
#pragma unroll
      for (int q = 0; q < S; ++q) g[q] = cur[q];
      const float s = g[0] + (S > 1 ? g[1] : 0.f) + R;
      rb = kb_update_adj<S>(b, g, s, v, Rb);
    }
    sr[i] = rb;
    if (i == 0) {
      const int tc = t / CI_TC;
      if (tc != nchunk - 1) xs.acquire();       // the last forward chunk is still resident
      if (k > 0) {
        const float rb0 = sr[0], rb1 = sr[1], rb2 = sr[2], rb3 = sr[3];
        const float* __restrict__ xp = xs.chunk(tc);
#pragma unroll 5
        for (int j = 0; j < k; ++j) {
          const f4 x = xs.load(xp + j * xs.rs());
          float a0 = sb[j * sbld], a1;
          a0 = fmaf(-rb0, x.x, a0); a1 = -rb1 * x.y;
          a0 = fmaf(-rb2, x.z, a0); a1 = fmaf(-rb3, x.w, a1);
          sb[j * sbld] = a0 + a1;
        }
      }
      xs.release(tc - 1);                        // chunk tc-1 streams in during the next CI_TC steps
    }
#pragma unroll
    for (int q = 0; q < TW; ++q) cur[q] = nxt[q];
  }
  params_backward_lane(u, ld, k, S, sc, scld, sb, sbld, Rb, qlb, qdb, grad);
  *logp_out = ll + ls.lp0;
}

// Barrier-only twin of eval_lane for the padding lanes of the last CTA: the same sequence of
// acquire / release calls, no arithmetic and no stores.
template <class XS>
CI_HD void idle_lane(XS& xs, int T) {
  const int nchunk = (T + CI_TC - 1) / CI_TC;
  for (int tc = 0; tc < nchunk; ++tc) {
    xs.acquire();
    xs.release(tc + 1 < nchunk ? tc + 1 : -1);
  }
  for (int tc = nchunk - 1; tc >= 0; --tc) {
    if (tc != nchunk - 1) xs.acquire();
    xs.release(tc - 1);
  }
}

}  // namespace ci

// ================================================================================================
// Seasonal-effects variant (ci_kalman_eff.cuh): season_duration == 1, 2 <= S <= CI_MAX_EFF_S.
// Time is walked in blocks of S steps whose bodies are fully unrolled, so the season index of every
// step is a compile-time constant and the state never moves between registers.
//
// Tape entry per step: g_0 .. g_{S-1}, v at tp[q * XS::TQ]  (g_S = -(g_1 + .. + g_{S-1}) because the
// seasonal part of every covariance column sums to zero; s is recomputed; v = NaN = masked step).
// Consecutive steps are `tstep` floats apart.
// ================================================================================================
#include "ci_kalman_eff.cuh"

namespace ci {

template <int S, class XS, int TQ, int SBLD>
struct EffLane {
  static constexpr int D = S + 1;
  XS& xs;
  const int T, k;
  float* __restrict__ sb;   // beta_bar / MINUS beta column, element j at sb[j * SBLD]
  float* __restrict__ sr;
  float* tp;            // this lane's column of the tape entry being written / read
  const int tstep;
  const int pfoff;      // reverse sweep: offset (floats) from tp to the line this lane prefetches, 0 = none
  const float R, ql, q11;
  EState<S> st;
  float acc = 0.f, quad = 0.f;   // sum of log2 s_t and of v_t^2 / s_t over the observed steps
  int nmask = 0;                 // masked (missing) steps; the common observed path carries no counter
  float Rb = 0.f, qlb = 0.f, qdb = 0.f;
  float sig = 0.f;   // 1^T Pbar_seas 1, carried incrementally through the reverse sweep

  CI_HD EffLane(XS& xs_, int T_, int k_, float* sb_, float* sr_, float* tape, int tstep_, float R_, float ql_,
                float qd_, int pfrow_)
      : xs(xs_), T(T_), k(k_), sb(sb_), sr(sr_), tp(tape), tstep(tstep_),
        pfoff(pfrow_ >= 0 ? pfrow_ * TQ - CI_TAPE_AHEAD * tstep_ : 0), R(R_), ql(ql_), q11(qd_) {}
  CI_HD int nchunk() const { return (T + CI_TC - 1) / CI_TC; }

  // r_i = (y_i - mean) + sum_j (-beta_j) X_ji for the CI_TC steps of chunk tc (sb holds -beta, row k
  // of the chunk holds the centred series).  Loads are issued in batches of 5 covariates and the sums
  // run in 4 independent FFMA2 chains so that shared-memory latency overlaps the arithmetic.
  CI_HD void regress_fwd(int tc) {
    xs.acquire();
    const float* __restrict__ xp = xs.chunk(tc);
    const f4 yv = xs.load(xp + k * xs.rs());
    pf2 r01 = mk2(yv.x, yv.y), r23 = mk2(yv.z, yv.w), s01 = mk2(0.f, 0.f), s23 = mk2(0.f, 0.f);
    const float* __restrict__ bp = sb;
    int j = 0;
    for (; j + 10 <= k; j += 10) {
#pragma unroll
      for (int h = 0; h < 10; h += 5) {
        float nb[5];
        f4 x[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          nb[q] = bp[(h + q) * SBLD];
          x[q] = xs.load(xp + (h + q) * xs.rs());
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          if ((h + q) & 1) {
            s01 = ffma2(mk2(nb[q], nb[q]), mk2(x[q].x, x[q].y), s01);
            s23 = ffma2(mk2(nb[q], nb[q]), mk2(x[q].z, x[q].w), s23);
          } else {
            r01 = ffma2(mk2(nb[q], nb[q]), mk2(x[q].x, x[q].y), r01);
            r23 = ffma2(mk2(nb[q], nb[q]), mk2(x[q].z, x[q].w), r23);
          }
        }
      }
      bp += 10 * SBLD;
      xp += 10 * xs.rs();
    }
    for (; j < k; ++j) {
      const float nb = bp[0];
      const f4 x = xs.load(xp);
      r01 = ffma2(mk2(nb, nb), mk2(x.x, x.y), r01);
      r23 = ffma2(mk2(nb, nb), mk2(x.z, x.w), r23);
      bp += SBLD;
      xp += xs.rs();
    }
    r01 = fadd2(r01, s01);
    r23 = fadd2(r23, s23);
    sr[0] = r01.x; sr[1] = r01.y; sr[2] = r23.x; sr[3] = r23.y;
    xs.release(tc + 1 < nchunk() ? tc + 1 : -1);
  }

  // beta_bar_j -= sum_i rbar_i X_ji for the CI_TC steps of chunk tc.
  CI_HD void regress_bwd(int tc) {
    if (tc != nchunk() - 1) xs.acquire();
    if (k > 0) {
      const pf2 n01 = mk2(-sr[0], -sr[1]), n23 = mk2(-sr[2], -sr[3]);
      const float* __restrict__ xp = xs.chunk(tc);
      float* __restrict__ bp = sb;
      int j = 0;
      for (; j + 10 <= k; j += 10) {
#pragma unroll
        for (int h = 0; h < 10; h += 5) {
          f4 x[5];
          pf2 a[5];
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            x[q] = xs.load(xp + (h + q) * xs.rs());
            a[q] = mk2(bp[(h + q) * SBLD], 0.f);
          }
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            a[q] = ffma2(mk2(x[q].x, x[q].y), n01, a[q]);
            a[q] = ffma2(mk2(x[q].z, x[q].w), n23, a[q]);
            bp[(h + q) * SBLD] = a[q].x + a[q].y;
          }
        }
        bp += 10 * SBLD;
        xp += 10 * xs.rs();
      }
      for (; j < k; ++j) {
        const f4 x = xs.load(xp);
        pf2 a = mk2(bp[0], 0.f);
        a = ffma2(mk2(x.x, x.y), n01, a);
        a = ffma2(mk2(x.z, x.w), n23, a);
        bp[0] = a.x + a.y;
        bp += SBLD;
        xp += xs.rs();
      }
    }
    xs.release(tc - 1);
  }

  template <int C>
  CI_HD void fwd_step(float r) {
    if (r == r) {
      float g[D], s, v, is;
      ef_update<S, C>(st, r, R, g, s, v, is);
#ifndef CI_NO_TAPE   // -DCI_NO_TAPE: timing-only build without any tape traffic (results are wrong), scripts/ab_eval.py
#pragma unroll
      for (int q = 0; q < S; ++q) tp[q * TQ] = g[q];
      tp[S * TQ] = v;
#endif
      // two plain sums: log2 of the innovation variances (scaled by ln 2 once at the end) and v^2 / s
      acc += fast_log2(s);
      quad = fmaf(v * is, v, quad);
    } else {
      ++nmask;
#ifndef CI_NO_TAPE
#pragma unroll
      for (int q = 0; q < S; ++q) tp[q * TQ] = 0.f;
      tp[S * TQ] = r;
#endif
    }
    ef_predict<S, C>(st, ql, q11, -(float)(S - 1) * q11, (float)((S - 1) * (S - 1)) * q11);
    tp += tstep;
  }

  // The entry of step t was pulled into L2 CI_TAPE_AHEAD steps ago: lane q < S + 1 of the warp prefetches
  // row q (one 128-byte line per row in the warp-blocked layout), i.e. ONE prefetch instruction per step.
  template <int C>
  CI_HD float bwd_step(bool far) {
    float cur[D];
#ifndef CI_NO_TAPE
#pragma unroll
    for (int q = 0; q < D; ++q) cur[q] = tape_load<tape_streams<XS>::value>(tp + q * TQ);
    if (far && pfoff != 0) prefetch_l2(tp + pfoff);
#else
#pragma unroll
    for (int q = 0; q < D; ++q) cur[q] = R * (0.01f * (float)(q + 1));   // stand-in values, no memory traffic
#endif
    eb_predict_adj<S, C>(st, qlb, qdb, sig);
    float rb = 0.f;
    const float v = cur[S];
    if (v == v) {
      float g[D];
      float gs = cur[1];
#pragma unroll
      for (int q = 0; q < S; ++q) {
        g[q] = cur[q];
        if (q >= 2) gs += cur[q];
      }
      g[S] = -gs;
      const float s = g[0] + g[1 + C] + R;
      rb = eb_update_adj<S, C>(st, g, s, v, Rb, sig);
    }
    tp -= tstep;
    return rb;
  }

  // uniform dispatch on the (runtime) season index: every case has static register indices
  template <int C>
  CI_HD void disp_fwd(int c, float r) {
    if (C == S - 1 || c == C) fwd_step<C>(r);
    else disp_fwd<(C + 1 < S ? C + 1 : C)>(c, r);
  }
  template <int C>
  CI_HD float disp_bwd(int c, bool far) {
    if (C == S - 1 || c == C) return bwd_step<C>(far);
    else return disp_bwd<(C + 1 < S ? C + 1 : C)>(c, far);
  }

  CI_HD void forward() {
    int c = 0;
    for (int t = 0; t < T; ++t) {
      const int i = t & (CI_TC - 1);
      if (i == 0) regress_fwd(t / CI_TC);
      disp_fwd<0>(c, sr[i]);
      c = (c + 1 == S) ? 0 : c + 1;
    }
  }

  CI_HD void backward() {
    int c = (T - 1) % S;
    for (int t = T - 1; t >= 0; --t) {
      const int i = t & (CI_TC - 1);
      sr[i] = disp_bwd<0>(c, t >= CI_TAPE_AHEAD);
      if (i == 0) regress_bwd(t / CI_TC);
      c = (c == 0) ? S - 1 : c - 1;
    }
  }
};

// Same contract as eval_lane; `tape` is this lane's base pointer, entries TQ floats apart, steps
// `tstep` floats apart (device: warp-blocked [T][warps][S+1][32], TQ = 32; host shim: TQ = 1); the
// scratch column stride SBLD is a compile-time constant so the regression loops address shared
// memory with immediates.
// The lane's global pointers (u, constants, outputs) are obtained from `ptrs()` twice -- before the sweeps and again
// after them -- instead of being passed in: on the device the functor re-derives them from the kernel parameters
// and the (volatile-read) block / thread index, so none of them occupies a register during the T-step sweeps.
struct LanePtrs {
  const float* u;
  const float* sc;
  float* logp;
  float* grad;
};
struct FixedLanePtrs {
  LanePtrs p;
  CI_HD LanePtrs operator()() const { return p; }
};

template <int S, int TQ, int SBLD, class XS, class PF>
CI_HD void eval_lane_eff(const PF& ptrs, long ld, long scld, XS& xs, int T, int k, float* __restrict__ sb,
                         float* __restrict__ sr, float* __restrict__ tape, int tstep, int pfrow = -1) {
  using LaneT = EffLane<S, XS, TQ, SBLD>;
  float lp0;
  const LanePtrs p0 = ptrs();
  const LaneScalars ls = params_forward_lane(p0.u, ld, k, S, p0.sc, scld, sb, SBLD, -1.0f);   // sb <- -beta
  lp0 = ls.lp0;
  LaneT L(xs, T, k, sb, sr, tape, tstep, ls.R, ls.ql, ls.qd, pfrow);
  ef_init<S>(L.st, p0.sc[SC_M0 * scld], p0.sc[SC_P0 * scld]);
  L.forward();
  const float ll = -0.5f * (fmaf(L.acc, 0.6931471805599453f, L.quad) + (float)(T - L.nmask) * CI_LOG_2PI);

  for (int j = 0; j < k; ++j) sb[j * SBLD] = 0.f;
#pragma unroll
  for (int i = 0; i < S + 1; ++i) L.st.m[i] = 0.f;
#pragma unroll
  for (int i = 0; i < EState<S>::NP; ++i) L.st.P[i] = 0.f;
  sr[0] = 0.f; sr[1] = 0.f; sr[2] = 0.f; sr[3] = 0.f;
  L.tp -= tstep;  // entry of step T-1
  L.backward();
  const LanePtrs p1 = ptrs();
  params_backward_lane(p1.u, ld, k, S, p1.sc, scld, sb, SBLD, L.Rb, L.qlb, L.qdb, p1.grad);
  *p1.logp = ll + lp0;
}

// pointer-argument form (host shim, tests)
template <int S, int TQ, int SBLD, class XS>
CI_HD void eval_lane_eff(const float* u, long ld, const float* __restrict__ sc, long scld, XS& xs,
                         int T, int k, float* __restrict__ sb, float* __restrict__ sr, float* __restrict__ tape,
                         int tstep, float* __restrict__ logp_out, float* __restrict__ grad, int pfrow = -1) {
  const FixedLanePtrs pf{LanePtrs{u, sc, logp_out, grad}};
  eval_lane_eff<S, TQ, SBLD>(pf, ld, scld, xs, T, k, sb, sr, tape, tstep, pfrow);
}

}  // namespace ci
