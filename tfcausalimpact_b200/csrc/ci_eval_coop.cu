// K1-K3 for SUB-WAVE batches: lane-cooperative joint log-prob + gradient evaluation.
//
// The thread-per-lane kernel (ci_eval.cu) fills the machine only when there are >= ~57k lanes; below that a lane
// is one thread walking 2 T dependent steps and the SMs idle (10k lanes: 2 warps per SM).  Here TPL threads
// (8, 4 or 2) share ONE (series, chain) lane of the seasonal-effects formulation (ci_kalman_eff.cuh: x = [level |
// e_0 .. e_{S-1}], D = S + 1 <= 8, H_t = e_0 + e_{1+c}, F = I):
//   * row-distributed covariance: sub-thread `sub` owns rows sub, sub + TPL, .. of the FULL D x D matrix P (and of
//     its adjoint W); the mean m and its adjoint are replicated in every sub-thread
//   * forward step: g_i = P[i][0] + P[i][J] is local, the D values are all-gathered with D width-TPL shuffles, the
//     rank-1 update P[i][:] -= g_i g / s and the time update are row-local
//   * reverse step: W g is row-local, the D values are all-gathered, everything else (g^T W g, sbar, gbar) is
//     replicated arithmetic; the symmetric update W += sym(gbar H^T) touches columns 0 / J of every row and the
//     whole rows 0 / J.  The quadratic forms of the Q adjoint are carried per row and reduced once at the end.
//   * masked (missing) steps are branch-free (gain forced to 0), so a warp never diverges and every shuffle uses
//     the full mask
//   * regression: the k covariates are split over the TPL sub-threads (-beta / beta_bar slices in shared memory,
//     X chunk TMA-staged per warp exactly as in ci_eval.cu), partial residuals combined by a butterfly
//   * parameter transforms and their VJP (ci_params.cuh arithmetic) are split the same way; the HMC leapfrog
//     kick / drift is applied by the sub-thread that produced the gradient component.
// Tape: [T][warps][lanes per warp][8] = g_0 .. g_{S-1}, v (NaN = masked) per lane and step -- one 128 B line per
// warp and step for TPL = 8.
// Same math and the same parity tests as the thread-per-lane kernel; replaces the
// arithmetic behind /root/reference/causalimpact/model.py:361-364,375-377 for small batches.
#include "ci_abi_common.cuh"
#include "ci_xstage.cuh"
#include "ci_kalman_eff.cuh"

namespace ci {

// lanes a warp of the cooperative kernel can touch -> series it can touch (lanes of a series are contiguous)
__host__ __device__ inline int coop_series_max(int G, int lpw) {
  const int a = (lpw + G - 2) / G + 1;
  return a < lpw ? a : lpw;
}

template <int TPL>
__device__ __forceinline__ float group_sum(float x) {
#pragma unroll
  for (int d = 1; d < TPL; d <<= 1) x += __shfl_xor_sync(0xffffffffu, x, d, TPL);
  return x;
}

template <int S, int TPL>
struct CoopLane {
  static constexpr int D = S + 1;
  static constexpr int RPT = 8 / TPL;   // row slots per sub-thread (row = sub + TPL * q; rows >= D are empty)
  static constexpr int LPW = 32 / TPL;
  const int sub;
  const float R, ql, q11, q1c, qcc;
  float m[D];          // forward: mean (replicated);  reverse: its adjoint (replicated)
  float ml[RPT];       // reverse: the adjoint mean at this sub-thread's own rows
  float P[RPT][D];     // forward: own rows of P;  reverse: own rows of W = Pbar
  float rs[RPT];       // reverse: sum_{j >= 1} W[row][j]
  float acc = 0.f, quad = 0.f;
  int nmask = 0;
  float Rb = 0.f, qlacc = 0.f, qdacc = 0.f;

  __device__ __forceinline__ CoopLane(int sub_, float R_, float ql_, float qd_)
      : sub(sub_), R(R_), ql(ql_), q11(qd_), q1c(-(float)(S - 1) * qd_), qcc((float)((S - 1) * (S - 1)) * qd_) {}

  __device__ __forceinline__ void init_forward(float m0, float p0) {
#pragma unroll
    for (int j = 0; j < D; ++j) m[j] = 0.f;
    m[0] = m0;
    const float off = -p0 * (1.0f / S);
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int row = sub + TPL * q;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        float v = 0.f;
        if (row == 0) v = (j == 0) ? p0 : 0.f;
        else if (row < D) v = (j == 0) ? 0.f : ((j == row) ? p0 + off : off);
        P[q][j] = v;
      }
    }
  }
  __device__ __forceinline__ void init_reverse() {
#pragma unroll
    for (int j = 0; j < D; ++j) m[j] = 0.f;
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      ml[q] = 0.f;
      rs[q] = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) P[q][j] = 0.f;
    }
  }

  // measurement + time update at season C; writes this lane's tape entry (tp = the lane's 8-float slot)
  template <int C>
  __device__ __forceinline__ void fwd(float r, float* __restrict__ tp) {
    constexpr int J = 1 + C;
    float gl[RPT];
#pragma unroll
    for (int q = 0; q < RPT; ++q) gl[q] = P[q][0] + P[q][J];
    float g[D];
#pragma unroll
    for (int j = 0; j < D; ++j) g[j] = __shfl_sync(0xffffffffu, gl[j / TPL], j % TPL, TPL);
    const bool obs = (r == r);
    const float s = g[0] + g[J] + R;
    const float v = obs ? r - m[0] - m[J] : 0.f;
    const float is = obs ? rcp(s) : 0.f;
    const float vis = v * is;
#pragma unroll
    for (int j = 0; j < D; ++j) m[j] = fmaf(g[j], vis, m[j]);
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int row = sub + TPL * q;
      const float ki = gl[q] * is;
#pragma unroll
      for (int j = 0; j < D; ++j) P[q][j] = fmaf(-ki, g[j], P[q][j]);
      // tape slot `row`: g_row for row < S, v (NaN marks a masked step) at slot S
      if (row <= S) tp[row] = (row == S) ? (obs ? v : r) : gl[q];
      // time update P += diag(ql, qd w w^T), w = 1 - S e_C  (row-local)
      const bool seas = row >= 1 && row < D;
      const bool isj = row == J;
      const float qa = seas ? (isj ? q1c : q11) : 0.f;
      const float qb = seas ? (isj ? qcc : q1c) : 0.f;
      P[q][0] += (row == 0) ? ql : 0.f;
#pragma unroll
      for (int j = 1; j < D; ++j) P[q][j] += (j == J) ? qb : qa;
    }
    acc += obs ? fast_log2(s) : 0.f;
    quad = fmaf(vis, v, quad);
    nmask += obs ? 0 : 1;
  }

  // adjoint of the step at season C given the taped entry `e` (8 floats); returns rbar = dL/dr_t
  template <int C>
  __device__ __forceinline__ float bwd(const float (&e)[8]) {
    constexpr int J = 1 + C;
    float g[D];
    float gs = 0.f;
#pragma unroll
    for (int j = 0; j < S; ++j) {
      g[j] = e[j];
      if (j >= 1) gs += e[j];
    }
    g[S] = -gs;   // the seasonal part of every covariance column sums to zero
    const float vraw = e[S];
    const bool obs = (vraw == vraw);
    const float v = obs ? vraw : 0.f;
    // adjoint of the time update (W before this step's measurement adjoint)
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int row = sub + TPL * q;
      qlacc += (row == 0) ? P[q][0] : 0.f;
      const bool seas = row >= 1 && row < D;
      qdacc += seas ? rs[q] : 0.f;
      qdacc += (row == J) ? fmaf((float)(S * S), P[q][J], -(2.f * S) * rs[q]) : 0.f;
    }
    const float s = g[0] + g[J] + R;
    const float is = obs ? rcp(s) : 0.f;
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) a = fmaf(m[j], g[j], a);
    float Pgl[RPT];
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) t = fmaf(P[q][j], g[j], t);
      Pgl[q] = t;
    }
    float Pg[D];
#pragma unroll
    for (int j = 0; j < D; ++j) Pg[j] = __shfl_sync(0xffffffffu, Pgl[j / TPL], j % TPL, TPL);
    float gPg = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) gPg = fmaf(g[j], Pg[j], gPg);
    const float vis = v * is;
    const float vb = (a - v) * is;
    const float sb = (gPg - a * v) * is * is - 0.5f * (is - vis * vis);
    Rb += sb;
    float gb[D];
    float gsum = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      gb[j] = fmaf(m[j], vis, fmaf(-2.f * is, Pg[j], (j == 0 || j == J) ? sb : 0.f));
      if (j >= 1) gsum += gb[j];
    }
    // W += sym(gb H^T), H = e_0 + e_J: columns 0 / J of every row, the whole rows 0 / J
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int row = sub + TPL * q;
      const bool hrow = (row == 0) || (row == J);
      const float gbl = fmaf(ml[q], vis, fmaf(-2.f * is, Pgl[q], hrow ? sb : 0.f));
      const float h = (row < D) ? 0.5f * gbl : 0.f;
      const float hs = hrow ? 0.5f : 0.f;
      P[q][0] += h;
      P[q][J] += h;
#pragma unroll
      for (int j = 0; j < D; ++j) P[q][j] = fmaf(hs, gb[j], P[q][j]);
      rs[q] += h + ((row == J) ? 0.5f * gsum : 0.f);
      ml[q] -= hrow ? vb : 0.f;
    }
    m[0] -= vb;
    m[J] -= vb;
    return vb;
  }

  template <int C>
  __device__ __forceinline__ void disp_fwd(int c, float r, float* __restrict__ tp) {
    if (C == S - 1 || c == C) fwd<C>(r, tp);
    else disp_fwd<(C + 1 < S ? C + 1 : C)>(c, r, tp);
  }
  template <int C>
  __device__ __forceinline__ float disp_bwd(int c, const float (&e)[8]) {
    if (C == S - 1 || c == C) return bwd<C>(e);
    else return disp_bwd<(C + 1 < S ? C + 1 : C)>(c, e);
  }
};

template <int S, int TPL>
__global__ void __launch_bounds__(32)
k_eval_coop(ci_layout L, const float* __restrict__ XY, const float* __restrict__ sconst, const float* u,
            float* logp, float* grad, float* __restrict__ tape, LeapArgs lf) {
  extern __shared__ __align__(16) float smem[];
  using Lane = CoopLane<S, TPL>;
  constexpr int D = S + 1, LPW = 32 / TPL;
  const int tid = threadIdx.x;
  const int sub = tid % TPL, lw = tid / TPL;
  const long B = (long)L.N * L.G;
  const long bw0 = (long)blockIdx.x * LPW;       // first lane of this warp (< B by construction of the grid)
  const bool valid = bw0 + lw < B;
  const long b = valid ? bw0 + lw : B - 1;       // padding lanes shadow the last lane (stores suppressed)
  const int N = L.N, k = L.k, T = L.T;
  const int n = (int)(b / L.G);
  const int stride = (k + 1) * CI_TC;
  const int KQ = (k + TPL - 1) / TPL;
  const int nchunk = (T + CI_TC - 1) / CI_TC;
  float* sb = smem + tid;                        // [KQ][32]: -beta / beta_bar of covariate sub + TPL q
  float* sx = smem + KQ * 32;                    // staging buffer [smax][stride]
  const int smax = coop_series_max(L.G, LPW);
  unsigned char* wbase = reinterpret_cast<unsigned char*>(sx + (long)smax * stride);

  // ---- TMA staging of the warp's run of series (same protocol as ci_eval.cu, XM = 1) --------------------------
  const int n_lo = (int)(bw0 / L.G);
  const long b_last = (bw0 + LPW - 1 < B ? bw0 + LPW - 1 : B - 1);
  const int ns = (int)(b_last / L.G) - n_lo + 1;
  XStaged xs;
  xs.lane_x = sx + (long)(n - n_lo) * stride;
  xs.bar = smem_u32(wbase);
  xs.parity = 0u;
  if (tid == 0) {
    *reinterpret_cast<uint64_t*>(wbase + 16) = reinterpret_cast<uint64_t>(XY + (long)n_lo * stride);
    *reinterpret_cast<uint32_t*>(wbase + 24) = (uint32_t)ns * (uint32_t)stride * 4u;
    *reinterpret_cast<uint32_t*>(wbase + 28) = smem_u32(sx);
    *reinterpret_cast<uint64_t*>(wbase + 32) = (uint64_t)((long)N * stride) * 4u;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xs.bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (tid == 0) xs.issue(0);

  // ---- u -> sigma^2's, -beta slices, prior + log-det (ci_params.cuh arithmetic, covariates split over sub-threads)
  const float* ub = u + b;
  const float* sc = sconst + n;
  const float sd = sc[SC_SD * N];
  const int Pn = num_params(k, S);
  float R, ql, qd, lp0;
  {
    const SoftplusPack p0 = softplus_pack(ub[0]), p1 = softplus_pack(ub[B]);
    const float so = sd * p0.sp, sl = sd * p1.sp;
    R = so * so;
    ql = sl * sl;
    float lps = sc[SC_PRIOR_CONST * N];
    lps += -(2.f * CI_PRIOR_A + 1.f) * logf(so) - sc[SC_OBS_B * N] / R + p0.lsig;
    lps += -(2.f * CI_PRIOR_A + 1.f) * logf(sl) - sc[SC_LEVEL_B * N] / ql + p1.lsig;
    float lpp = 0.f;
    if (k > 0) {
      const SoftplusPack pg = softplus_pack(ub[2 * B]), pn = softplus_pack(ub[3 * B]);
      const float gsv = pg.sp, gsn = pn.sp;
      lps += -1.5f * logf(gsv) - 0.5f / gsv - 0.5f * gsn * gsn + pg.lsig + pn.lsig;
      const float Gs = -gsn * sqrtf(gsv) * CI_WEIGHTS_PRIOR_SCALE;    // sb <- MINUS beta
      for (int q = 0; q < KQ; ++q) {
        const int j = sub + TPL * q;
        float nb = 0.f;
        if (j < k) {
          const float a = ub[(long)(4 + j) * B], bb = ub[(long)(4 + k + j) * B], wn = ub[(long)(4 + 2 * k + j) * B];
          const SoftplusPack pa = softplus_pack(a), pb = softplus_pack(bb);
          const float lsv = pa.sp, lsn = pb.sp;
          lpp += -1.5f * logf(lsv) - 0.5f * rcp(lsv) - 0.5f * lsn * lsn - 0.5f * wn * wn + pa.lsig + pb.lsig;
          nb = wn * lsn * sqrtf(lsv) * Gs;
        }
        sb[q * 32] = nb;
      }
    }
    {
      const SoftplusPack pd = softplus_pack(ub[(long)(Pn - 1) * B]);
      const float sdr = sd * pd.sp;
      const float lg = logf(sdr);
      const float z = (lg - sc[SC_DRIFT_MU * N]) * (1.0f / 3.0f);
      lps += -lg - 0.5f * z * z + pd.lsig;
      const float qq = sdr * (1.0f / S);
      qd = qq * qq;
    }
    lp0 = lps + group_sum<TPL>(lpp);
  }

  Lane ln(sub, R, ql, qd);
  ln.init_forward(sc[SC_M0 * N], sc[SC_P0 * N]);
  const int tstep = (int)gridDim.x * LPW * 8;                       // floats between consecutive steps
  float* tp = tape + ((long)blockIdx.x * LPW + lw) * 8;

  // ---- forward sweep ----------------------------------------------------------------------------------------
  {
    int c = 0;
    for (int tc = 0; tc < nchunk; ++tc) {
      xs.acquire();
      const float* __restrict__ xp = xs.lane_x;
      pf2 r01 = mk2(0.f, 0.f), r23 = mk2(0.f, 0.f);
      for (int q = 0; q < KQ; ++q) {
        const int j = sub + TPL * q;
        const float nb = sb[q * 32];                                // 0 past k
        const float4 x = *reinterpret_cast<const float4*>(xp + (j < k ? j : k) * CI_TC);
        r01 = ffma2(mk2(nb, nb), mk2(x.x, x.y), r01);
        r23 = ffma2(mk2(nb, nb), mk2(x.z, x.w), r23);
      }
      const float4 yv = *reinterpret_cast<const float4*>(xp + k * CI_TC);   // row k: centred series
      xs.release(tc + 1 < nchunk ? tc + 1 : -1);
      float rr[CI_TC] = {r01.x, r01.y, r23.x, r23.y};
#pragma unroll
      for (int i = 0; i < CI_TC; ++i) rr[i] = group_sum<TPL>(rr[i]);
      rr[0] += yv.x; rr[1] += yv.y; rr[2] += yv.z; rr[3] += yv.w;
      const int nstep = (T - tc * CI_TC) < CI_TC ? (T - tc * CI_TC) : CI_TC;
      for (int i = 0; i < nstep; ++i) {
        const float r = (i == 0) ? rr[0] : (i == 1) ? rr[1] : (i == 2) ? rr[2] : rr[3];
        ln.template disp_fwd<0>(c, r, tp);
        tp += tstep;
        c = (c + 1 == S) ? 0 : c + 1;
      }
    }
  }
  const float ll = -0.5f * (fmaf(ln.acc, 0.6931471805599453f, ln.quad) + (float)(T - ln.nmask) * CI_LOG_2PI);

  // ---- reverse sweep ----------------------------------------------------------------------------------------
  for (int q = 0; q < KQ; ++q) sb[q * 32] = 0.f;
  ln.init_reverse();
  {
    tp -= tstep;                                                    // entry of step T - 1
    float cur[8], nxt[8];
    {
      const float4 a = *reinterpret_cast<const float4*>(tp), bq = *reinterpret_cast<const float4*>(tp + 4);
      cur[0] = a.x; cur[1] = a.y; cur[2] = a.z; cur[3] = a.w;
      cur[4] = bq.x; cur[5] = bq.y; cur[6] = bq.z; cur[7] = bq.w;
    }
    float rb0 = 0.f, rb1 = 0.f, rb2 = 0.f, rb3 = 0.f;
    int c = (T - 1) % S;
    for (int t = T - 1; t >= 0; --t) {
      const int i = t & (CI_TC - 1);
      if (t > 0) {
        const float* tq = tp - tstep;
        const float4 a = *reinterpret_cast<const float4*>(tq), bq = *reinterpret_cast<const float4*>(tq + 4);
        nxt[0] = a.x; nxt[1] = a.y; nxt[2] = a.z; nxt[3] = a.w;
        nxt[4] = bq.x; nxt[5] = bq.y; nxt[6] = bq.z; nxt[7] = bq.w;
        if (sub == 0 && t > CI_TAPE_AHEAD) prefetch_l2(tp - (long)CI_TAPE_AHEAD * tstep);
      }
      const float rb = ln.template disp_bwd<0>(c, cur);
      rb0 = (i == 0) ? rb : rb0;
      rb1 = (i == 1) ? rb : rb1;
      rb2 = (i == 2) ? rb : rb2;
      rb3 = (i == 3) ? rb : rb3;
      if (i == 0) {
        const int tc = t / CI_TC;
        if (tc != nchunk - 1) xs.acquire();                         // the last forward chunk is still resident
        if (k > 0) {
          const pf2 n01 = mk2(-rb0, -rb1), n23 = mk2(-rb2, -rb3);
          const float* __restrict__ xp = xs.lane_x;
          for (int q = 0; q < KQ; ++q) {
            const int j = sub + TPL * q;
            const float4 x = *reinterpret_cast<const float4*>(xp + (j < k ? j : k) * CI_TC);
            pf2 a = mk2(sb[q * 32], 0.f);
            a = ffma2(mk2(x.x, x.y), n01, a);
            a = ffma2(mk2(x.z, x.w), n23, a);
            sb[q * 32] = a.x + a.y;
          }
        }
        xs.release(tc - 1);
        rb0 = rb1 = rb2 = rb3 = 0.f;
      }
      tp -= tstep;
      c = (c == 0) ? S - 1 : c - 1;
#pragma unroll
      for (int q = 0; q < 8; ++q) cur[q] = nxt[q];
    }
  }
  const float Rb = ln.Rb;
  const float qlb = __shfl_sync(0xffffffffu, ln.qlacc, 0, TPL);     // row 0 lives in sub-thread 0
  const float qdb = group_sum<TPL>(ln.qdacc);

  // ---- VJP through the parameter transforms + priors; leapfrog kick / drift on the fresh component -----------
  // every scalar of u that more than one sub-thread reads is read BEFORE any sub-thread may drift it
  const float u0 = ub[0], u1 = ub[B];
  const float ug = k > 0 ? ub[2 * B] : 0.f, un = k > 0 ? ub[3 * B] : 0.f;
  const float ud = ub[(long)(Pn - 1) * B];
  __syncwarp();
  const bool leap = lf.pp != nullptr;
  const float lfac = leap ? expf(lf.logfac[b]) : 0.f;
  const float lc = lf.last ? 0.5f : 1.0f;
  auto emit = [&](int p, float gv) {
    if (!valid) return;
    const long o = (long)p * B + b;
    grad[o] = gv;
    if (leap) {
      const float eps = lf.step0[o] * lfac;
      const float ph = fmaf(lc * eps, gv, lf.pp[o]);
      lf.pp[o] = ph;
      if (!lf.last) lf.uu[o] = fmaf(eps, ph, lf.uu[o]);
    }
  };
  float dG = 0.f;
  float gsv = 0.f, gsn = 0.f, sq_gsv = 0.f;
  SoftplusPack pg{}, pn{};
  if (k > 0) {
    pg = softplus_pack(ug);
    pn = softplus_pack(un);
    gsv = pg.sp;
    gsn = pn.sp;
    sq_gsv = sqrtf(gsv);
    const float Gs = gsn * sq_gsv * CI_WEIGHTS_PRIOR_SCALE;
    for (int q = 0; q < KQ; ++q) {
      const int j = sub + TPL * q;
      if (j < k) {
        const float a = ub[(long)(4 + j) * B], bq = ub[(long)(4 + k + j) * B], wn = ub[(long)(4 + 2 * k + j) * B];
        const float bb = sb[q * 32];
        const SoftplusPack pa = softplus_pack(a), pb = softplus_pack(bq);
        const float lsv = pa.sp, lsn = pb.sp;
        const float sq = sqrtf(lsv);
        const float ilsv = rcp(lsv);
        const float l = lsn * sq;
        dG = fmaf(bb, wn * l, dG);
        const float dl = bb * wn * Gs;
        emit(4 + 2 * k + j, bb * l * Gs - wn);
        emit(4 + k + j, (dl * sq - lsn) * pb.sig + pb.nsig);
        emit(4 + j, (dl * lsn * 0.5f * sq * ilsv + ilsv * (0.5f * ilsv - 1.5f)) * pa.sig + pa.nsig);
      }
    }
    dG = group_sum<TPL>(dG);
  }
  if (sub == 0) {
    {
      const SoftplusPack p = softplus_pack(u0);
      const float so = sd * p.sp;
      const float dth = -(2.f * CI_PRIOR_A + 1.f) / so + 2.f * sc[SC_OBS_B * N] / (so * so * so) + 2.f * so * Rb;
      emit(0, dth * sd * p.sig + p.nsig);
    }
    {
      const SoftplusPack p = softplus_pack(u1);
      const float sl = sd * p.sp;
      const float dth = -(2.f * CI_PRIOR_A + 1.f) / sl + 2.f * sc[SC_LEVEL_B * N] / (sl * sl * sl) + 2.f * sl * qlb;
      emit(1, dth * sd * p.sig + p.nsig);
    }
    if (k > 0) {
      emit(3, (dG * sq_gsv * CI_WEIGHTS_PRIOR_SCALE - gsn) * pn.sig + pn.nsig);
      emit(2, (dG * gsn * CI_WEIGHTS_PRIOR_SCALE * 0.5f / sq_gsv - 1.5f / gsv + 0.5f / (gsv * gsv)) * pg.sig + pg.nsig);
    }
    {
      const SoftplusPack pd = softplus_pack(ud);
      const float sdr = sd * pd.sp;
      const float lg = logf(sdr);
      const float dth = -1.0f / sdr - (lg - sc[SC_DRIFT_MU * N]) / (9.0f * sdr) + 2.f * sdr * qdb * (1.0f / (S * S));
      emit(Pn - 1, dth * sd * pd.sig + pd.nsig);
    }
    if (valid) logp[b] = ll + lp0;
  }
}

// ---- host side ------------------------------------------------------------------------------------------------
static int g_coop_tpl = 8;              // sub-threads per lane (8, 4, 2); 0 disables the cooperative kernel
static long g_coop_max_lanes = -1;      // -1 = one resident wave of the chosen TPL

void coop_set_option(int tpl, long max_lanes) {
  g_coop_tpl = tpl;
  g_coop_max_lanes = max_lanes;
}

bool coop_eligible(const ci_layout* L) {
  if (g_coop_tpl == 0) return false;
  if (!ci_reg_path(L) || L->r != 1 || L->S < 2 || L->S > CI_MAX_EFF_S) return false;
  const long B = (long)L->N * L->G;
  const long lim = g_coop_max_lanes >= 0 ? g_coop_max_lanes : (long)148 * 32 * (32 / g_coop_tpl);
  return B <= lim;
}

size_t coop_tape_floats(const ci_layout* L) {
  // any TPL: lanes padded to a whole warp of the widest lane count (16 lanes at TPL = 2)
  const size_t Bp = ((size_t)L->N * L->G + 15) / 16 * 16;
  return (size_t)L->T * Bp * 8;
}

template <int S, int TPL>
static int launch_coop_t(const ci_layout* L, const ci_data* D, const float* d_u, float* d_logp, float* d_grad,
                         float* tape, cudaStream_t st, const LeapArgs& lf) {
  constexpr int LPW = 32 / TPL;
  const long B = (long)L->N * L->G;
  const int KQ = (L->k + TPL - 1) / TPL;
  const size_t smem = (size_t)KQ * 32 * 4 + (size_t)coop_series_max(L->G, LPW) * (L->k + 1) * CI_TC * 4 + 48;
  auto kern = k_eval_coop<S, TPL>;
  if (smem > 48 * 1024) {
    if (smem > 227 * 1024) return fail(CI_ERR_UNSUPPORTED, "k=%d covariates exceed the shared-memory budget", L->k);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(CI_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  kern<<<(unsigned)((B + LPW - 1) / LPW), 32, smem, st>>>(*L, D->d_X, D->d_sconst, d_u, d_logp, d_grad, tape, lf);
  return CI_OK;
}

int launch_eval_coop(const ci_layout* L, const ci_data* D, const float* d_u, float* d_logp, float* d_grad,
                     float* tape, cudaStream_t st, const LeapArgs& lf) {
  int rc = CI_OK;
#define CI_COOP_S(TPL_)                                                                              \
  switch (L->S) {                                                                                    \
    case 2: rc = launch_coop_t<2, TPL_>(L, D, d_u, d_logp, d_grad, tape, st, lf); break;             \
    case 3: rc = launch_coop_t<3, TPL_>(L, D, d_u, d_logp, d_grad, tape, st, lf); break;             \
    case 4: rc = launch_coop_t<4, TPL_>(L, D, d_u, d_logp, d_grad, tape, st, lf); break;             \
    case 5: rc = launch_coop_t<5, TPL_>(L, D, d_u, d_logp, d_grad, tape, st, lf); break;             \
    case 6: rc = launch_coop_t<6, TPL_>(L, D, d_u, d_logp, d_grad, tape, st, lf); break;             \
    case 7: rc = launch_coop_t<7, TPL_>(L, D, d_u, d_logp, d_grad, tape, st, lf); break;             \
    default: return fail(CI_ERR_UNSUPPORTED, "cooperative evaluation needs 2 <= nseasons <= 7");     \
  }
  if (g_coop_tpl == 8) CI_COOP_S(8)
  else if (g_coop_tpl == 4) CI_COOP_S(4)
  else CI_COOP_S(2)
#undef CI_COOP_S
  if (rc) return rc;
  CI_CHECK_LAUNCH();
  return CI_OK;
}

}  // namespace ci
