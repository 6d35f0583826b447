// K1-K3 for SUB-WAVE batches: lane-cooperative joint log-prob + gradient evaluation.
//
// The thread-per-lane kernel (ci_eval.cu) fills the machine only when there are >= ~57k lanes; below that a lane
// is one thread walking 2 T dependent steps and the SMs idle (10k lanes: 2 warps per SM).  Here TPL threads (8 or
// 4) share ONE (series, chain) lane of the seasonal-effects formulation (ci_kalman_eff.cuh: x = [level |
// e_0 .. e_{S-1}], D = S + 1 <= 8, H_t = e_0 + e_{1+c}, F = I).  One warp = one CTA = 32 / TPL lanes.
//
// The evaluation is FOUR passes over the series, so that the dependent Kalman chains contain nothing else:
//   1. residual pass   r_t = y_t - mean - X_t . beta for all t -> shared memory (covariates split over the TPL
//                      sub-threads, -beta slices in registers, X chunks TMA-staged two deep, partial sums combined
//                      by a recursive-halving butterfly: 4 shuffles per 4 steps at TPL = 8)
//   2. forward filter  sub-thread `sub` owns rows sub * RPT .. of the FULL D x D covariance P as FP32x2 pairs; the
//                      mean is replicated.  Per step: g_i = P[i][0] + P[i][J] is row-local, the 8 values are
//                      all-gathered (shared-memory exchange: 1 store + 2 x 16-byte loads), the rank-1
//                      update and the time update are row-local FFMA2 / FADD2.  The time loop is unrolled over the
//                      season cycle, so every register index is static and there is no per-step dispatch.
//                      Tape: g_0 .. g_7 per lane and step = exactly the row of the exchange buffer.  The rows of one
//                      season cycle form a block of a shared-memory ring that ONE bulk (TMA) store per cycle writes to
//                      the warp's contiguous tape [warp][T][8 x lanes]; the innovation v_t replaces r_t in shared memory
//                      (NaN = masked step).
//   3. reverse sweep   rows of the symmetric adjoint W = Pbar distributed the same way, mean adjoint replicated;
//                      W g is row-local, all-gathered; g^T W g, sbar, gbar are replicated pair arithmetic; the
//                      update W += sym(gbar H^T) touches columns 0 / J of every row and the whole rows 0 / J.
//                      The tape streams back by bulk (TMA) loads three season cycles ahead of the sweep, so no global
//                      load sits in the dependent chain.  dL/dr_t replaces v_t in shared memory.  The Q adjoint is
//                      accumulated per row, reduced once.
//   4. beta_bar pass   beta_bar_j = -sum_t rbar_t X_tj, second streaming pass over the TMA-staged chunks with
//                      register accumulators, then the VJP through the parameter transforms (ci_params.cuh
//                      arithmetic, covariates split over the sub-threads) and the fused HMC leapfrog kick / drift.
// Masked (missing) steps are branch-free (gain forced to 0), so a warp never diverges.
// Same math and the same parity tests as the thread-per-lane kernel; replaces the
// arithmetic behind /root/reference/causalimpact/model.py:361-364,375-377 for small batches.
#include <type_traits>
#include "ci_abi_common.cuh"
#include "ci_xstage.cuh"
#include "ci_kalman_eff.cuh"

// lane counts served by the cooperative kernel (measured crossovers, scripts/ab_coop.py; B200, k = 50, S = 7, 8 chains)
// eval ms at lanes =          1000   2000   3000   4000   5000   8000   9472   10000  11800  15000  (profiles/r2_coop_ab_k50.jsonl)
//   thread per lane           0.447  0.445  0.446  0.394  0.475  0.432  --     0.499  0.506  0.502   (single-warp kernel)
//   thread per lane + helper  0.262  0.272  0.269  0.285  0.290  0.293  0.295  0.295  0.295  0.295   (ci_eval_ws.cu)
//   8 sub-threads             0.126  0.128  0.159  0.161  0.209  0.266  0.267  0.324  0.394  0.461
//   4 sub-threads, FFMA2      0.173  0.174  0.173  0.193  0.217  0.222  0.235  0.288  0.295  0.481
//   4 sub-threads, tensor c.  0.142  0.143  0.142  0.161  0.187  0.191  0.204  0.261  0.261  0.445
// The steps are scheduler granularity: 592 schedulers x 2 warps x 8 lanes = 9,472 lanes at 4 sub-threads (10,000 lanes
// put a third warp on 66 schedulers: +39 % time for +25 % lanes), 592 x 2 x 4 = 4,736 lanes at 8 sub-threads.
// The step is sharp -- 9,472 lanes 0.200 ms, 9,600 lanes 0.258 ms -- because the kernel ends with its slowest scheduler.
// (Running the overflow past 9,472 lanes as a second launch of 8-sub-thread warps on a side stream, so that the third
// warp leaves its scheduler half way through, was measured: 10,000 lanes 0.259 -> 0.262 ms.  Not kept.)
#define CI_COOP_TPL8_LANES 2500L        // up to here 8 sub-threads per lane (tensor-core passes available above)
#define CI_COOP_TPL8_LANES_FFMA 6000L   // ... when the warp's lanes span several series (FFMA2 passes at 4 sub-threads)
#define CI_COOP_MAX_LANES 13000L    // one resident wave at 4 sub-threads; above: thread-per-lane kernel
#define CI_COOP_MAX_STAGES 16       // design-matrix chunks in flight per warp (TMA stages)

namespace ci {

// lanes a warp of the cooperative kernel can touch -> series it can touch (lanes of a series are contiguous)
__host__ __device__ inline int coop_series_max(int G, int lpw) {
  if (G % lpw == 0) return 1;                  // a warp's lanes never straddle a series
  const int a = (lpw + G - 2) / G + 1;
  return a < lpw ? a : lpw;
}
// floats per lane of the shared-memory residual row: whole chunks, rows of a warp's lanes on distinct banks
__host__ __device__ inline int coop_row_floats(int T, int lpw) {
  const int nchunk = (T + CI_TC - 1) / CI_TC;
  return (nchunk * CI_TC + 31) / 32 * 32 + 32 / lpw;
}

template <int TPL>
__device__ __forceinline__ float group_sum(float x) {
#pragma unroll
  for (int d = 1; d < TPL; d <<= 1) x += __shfl_xor_sync(0xffffffffu, x, d, TPL);
  return x;
}

template <int J>
__device__ __forceinline__ float& el(pf2 (&a)[4]) {
  if constexpr (J & 1) return a[J >> 1].y;
  else return a[J >> 1].x;
}
template <int J>
__device__ __forceinline__ float elc(const pf2 (&a)[4]) {
  if constexpr (J & 1) return a[J >> 1].y;
  else return a[J >> 1].x;
}
__device__ __forceinline__ float dot8(const pf2 (&a)[4], const pf2 (&b)[4]) {
  pf2 t0 = fmul2(a[0], b[0]), t1 = fmul2(a[1], b[1]);
  t0 = ffma2(a[2], b[2], t0);
  t1 = ffma2(a[3], b[3], t1);
  t0 = fadd2(t0, t1);
  return t0.x + t0.y;
}

template <int S, int TPL>
struct CoopLane {
  static constexpr int D = S + 1;
  static constexpr int RPT = 8 / TPL;   // rows per sub-thread: row = sub * RPT + q (rows >= D stay zero)
  static constexpr int LPW = 32 / TPL;
  const int sub;
  const float R;
  pf2 m[4];            // forward: mean (replicated);  reverse: its adjoint (replicated)
  pf2 P[RPT][4];       // forward: own rows of P;  reverse: own rows of W = Pbar
  float ml[RPT];       // reverse: the adjoint mean at this sub-thread's own rows
  // per-row constants of the time update / its adjoint
  float ql0[RPT], c0[RPT], c1[RPT], c2[RPT], r0m[RPT], sm1[RPT], smJ[RPT];
  float acc = 0.f, quad = 0.f;
  int nmask = 0;
  float Rb = 0.f, qlacc = 0.f, qdacc = 0.f;

  __device__ __forceinline__ CoopLane(int sub_, float R_, float ql, float qd) : sub(sub_), R(R_) {
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int row = sub * RPT + q;
      const bool seas = row >= 1 && row < D;
      ql0[q] = row == 0 ? ql : 0.f;
      c0[q] = seas ? qd : 0.f;                                   // Q[i][j], i, j != J
      c1[q] = seas ? -(float)(S - 1) * qd : 0.f;                 // Q[i][J] / Q[J][j]
      c2[q] = seas ? (float)((S - 1) * (S - 1)) * qd : 0.f;      // Q[J][J]
      r0m[q] = row == 0 ? 1.f : 0.f;
      sm1[q] = seas ? 1.f : 0.f;
      smJ[q] = seas ? -(float)(S - 1) : 0.f;
    }
  }

  __device__ __forceinline__ void init_forward(float m0, float p0) {
#pragma unroll
    for (int p = 0; p < 4; ++p) m[p] = mk2(0.f, 0.f);
    m[0].x = m0;
    const float off = -p0 * (1.0f / S);
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int row = sub * RPT + q;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float x = 0.f;
        if (j < D) {
          if (row == 0) x = (j == 0) ? p0 : 0.f;
          else if (row < D) x = (j == 0) ? 0.f : ((j == row) ? p0 + off : off);
        }
        v[j] = x;
      }
#pragma unroll
      for (int p = 0; p < 4; ++p) P[q][p] = mk2(v[2 * p], v[2 * p + 1]);
    }
  }
  __device__ __forceinline__ void init_reverse() {
#pragma unroll
    for (int p = 0; p < 4; ++p) m[p] = mk2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      ml[q] = 0.f;
#pragma unroll
      for (int p = 0; p < 4; ++p) P[q][p] = mk2(0.f, 0.f);
    }
  }

  // all-gather of one value per row through the lane's 8-float exchange row xb: out[j] = the value of row j.
  // The caller alternates rows (or separates reuse by a __syncwarp), so one warp barrier per exchange suffices.
  __device__ __forceinline__ void gather(const float (&loc)[RPT], pf2 (&out)[4], float* __restrict__ xb) {
    if constexpr (RPT == 1) xb[sub] = loc[0];
    else if constexpr (RPT == 2) *reinterpret_cast<float2*>(xb + 2 * sub) = make_float2(loc[0], loc[1]);
    else *reinterpret_cast<float4*>(xb + 4 * sub) = make_float4(loc[0], loc[1], loc[2], loc[3]);
    __syncwarp();
    const float4 a = *reinterpret_cast<const float4*>(xb), b = *reinterpret_cast<const float4*>(xb + 4);
    out[0] = mk2(a.x, a.y);
    out[1] = mk2(a.z, a.w);
    out[2] = mk2(b.x, b.y);
    out[3] = mk2(b.z, b.w);
  }

  // measurement + time update at season C.  r = residual (NaN = masked); xb = the lane's row of the tape ring (the
  // all-gathered gains ARE the tape entry); returns what replaces r_t in shared memory (v_t, NaN kept if masked)
  template <int C>
  __device__ __forceinline__ float fwd(float r, float* __restrict__ xb) {
    constexpr int J = 1 + C;
    float gl[RPT];
#pragma unroll
    for (int q = 0; q < RPT; ++q) gl[q] = elc<0>(P[q]) + elc<J>(P[q]);
    pf2 g[4];
    gather(gl, g, xb);
    const bool obs = (r == r);
    const float s = elc<0>(g) + elc<J>(g) + R;
    const float v = obs ? r - elc<0>(m) - elc<J>(m) : 0.f;
    const float is = obs ? rcp(s) : 0.f;
    const float vis = v * is;
    const pf2 vv = mk2(vis, vis);
#pragma unroll
    for (int p = 0; p < 4; ++p) m[p] = ffma2(g[p], vv, m[p]);
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int row = sub * RPT + q;
      const float ki = -gl[q] * is;
      const pf2 kk = mk2(ki, ki);
      // time update P += diag(ql, qd w w^T), w = 1 - S e_C, folded into the rank-1 update's addend
      const bool isj = row == J;
      const float qa = isj ? c1[q] : c0[q];
      const float qb = isj ? c2[q] : c1[q];
      float a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = (j == 0) ? ql0[q] : (j >= D ? 0.f : (j == J ? qb : qa));
#pragma unroll
      for (int p = 0; p < 4; ++p) P[q][p] = fadd2(ffma2(kk, g[p], P[q][p]), mk2(a[2 * p], a[2 * p + 1]));
    }
    acc += obs ? fast_log2(s) : 0.f;
    quad = fmaf(vis, v, quad);
    nmask += obs ? 0 : 1;
    return obs ? v : r;
  }

  // adjoint of the step at season C given the taped gains g (8 floats) and innovation; returns rbar = dL/dr_t
  template <int C>
  __device__ __forceinline__ float bwd(const float4& ea, const float4& eb, float vraw, float* __restrict__ xb) {
    constexpr int J = 1 + C;
    pf2 g[4] = {mk2(ea.x, ea.y), mk2(ea.z, ea.w), mk2(eb.x, eb.y), mk2(eb.z, eb.w)};
    const bool obs = (vraw == vraw);
    const float v = obs ? vraw : 0.f;
    // adjoint of the time update (W before this step's measurement adjoint): ql += W[0][0], qd += w^T W w
    pf2 wc[4];
    {
      float w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = (j == 0 || j >= D) ? 0.f : (j == J ? -(float)(S - 1) : 1.f);
#pragma unroll
      for (int p = 0; p < 4; ++p) wc[p] = mk2(w[2 * p], w[2 * p + 1]);
    }
    float Wgl[RPT];
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int row = sub * RPT + q;
      qlacc = fmaf(r0m[q], elc<0>(P[q]), qlacc);
      const float t = dot8(P[q], wc);
      qdacc = fmaf(row == J ? smJ[q] : sm1[q], t, qdacc);
      Wgl[q] = dot8(P[q], g);
    }
    const float s = elc<0>(g) + elc<J>(g) + R;
    const float is = obs ? rcp(s) : 0.f;
    const float a = dot8(m, g);
    pf2 Wg[4];
    gather(Wgl, Wg, xb);
    const float gPg = dot8(g, Wg);
    const float vis = v * is;
    const float vb = (a - v) * is;
    const float sb = (gPg - a * v) * is * is - 0.5f * (is - vis * vis);
    Rb += sb;
    const float m2is = -2.f * is;
    const pf2 vv = mk2(vis, vis), nn = mk2(m2is, m2is);
    pf2 gb[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) gb[p] = ffma2(m[p], vv, fmul2(Wg[p], nn));
    el<0>(gb) += sb;
    el<J>(gb) += sb;
    // W += sym(gb H^T), H = e_0 + e_J: columns 0 / J of every row, the whole rows 0 / J
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int row = sub * RPT + q;
      const bool hrow = (row == 0) || (row == J);
      const float gbl = fmaf(ml[q], vis, Wgl[q] * m2is) + (hrow ? sb : 0.f);   // == gb[row], same rounding
      const float h = 0.5f * gbl;
      const float hs = hrow ? 0.5f : 0.f;
      const pf2 hh = mk2(hs, hs);
      float a8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a8[j] = (j == 0 || j == J) ? h : 0.f;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        pf2 t = ffma2(hh, gb[p], P[q][p]);
        if (2 * p == 0 || 2 * p == J || 2 * p + 1 == J) t = fadd2(t, mk2(a8[2 * p], a8[2 * p + 1]));
        P[q][p] = t;
      }
      ml[q] -= hrow ? vb : 0.f;
    }
    el<0>(m) -= vb;
    el<J>(m) -= vb;
    return vb;
  }
};

// compile-time loops over the season index (ascending / descending)
template <int C, int N, class F>
__device__ __forceinline__ void static_up(F&& f) {
  if constexpr (C < N) {
    f(std::integral_constant<int, C>{});
    static_up<C + 1, N>(f);
  }
}
template <int C, class F>
__device__ __forceinline__ void static_down(F&& f) {
  if constexpr (C >= 0) {
    f(std::integral_constant<int, C>{});
    static_down<C - 1>(f);
  }
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// bulk (TMA) copies global -> shared behind an mbarrier; called by ONE thread after a warp barrier.  The buffer being
// overwritten was last READ through the generic proxy by threads that have passed that barrier, which orders the
// reuse (the consumer-release / producer-acquire pattern of TMA pipelines; no cross-proxy fence on this direction).
__device__ __forceinline__ void bulk_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_in(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  bulk_expect(bar, bytes);
  bulk_copy_in(dst, src, bytes, bar);
}
// one bulk (TMA) copy shared -> global as its own bulk group; called by ONE thread after the writers' proxy fence
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}


// ---- regression passes on the tensor cores (TPL = 4, the warp's 8 lanes = 8 chains of ONE series) -------------------
// The residual pass is the product  R[t][chain] = sum_j X[t][j] (-beta[j][chain])  and the beta_bar pass is
// beta_bar[j][chain] = -sum_t X[t][j] rbar[chain][t]: per series a [T x k] x [k x 8] and a [k x T] x [T x 8] matrix
// product.  mma.sync m16n8k8 (TF32 inputs, FP32 accumulation) with the 8 CHAINS on the N axis and full 16-row tiles on
// M (16 steps / 16 covariates): thread (lane lw, sub-thread sub) is fragment position (groupID = lw, threadID_in_group
// = sub), so the B fragments are the thread's OWN -beta slices (pass 1) / OWN rbar row entries (pass 4); the A
// fragments are scalar shared-memory loads of the staged design chunk; the C fragment of pass 1 goes to the rows of
// chains 2 sub, 2 sub + 1 (conflict-free scattered stores), that of pass 4 through a small exchange to the epilogue's
// ownership.  (A first version put the chains on M -- rows 8..15 zero padding, every fragment the thread's own data --
// and needed twice the tensor-core instructions: 0.272 against 0.261 ms at 10,000 lanes.)  Ragged ends (fewer than four
// chunks left in a half) still use that 8-row form.  FP32 accuracy comes from the 3xTF32 split (a = hi + lo, both
// TF32: hi*hi + lo*hi + hi*lo, the dropped lo*lo term is 2^-22 relative), which the parity tests hold to the same 1e-4
// as the FFMA path.
// Measured: HMMA.1688.F32.TF32 issues every 8 cycles per scheduler with 21 cycles latency on B200
// (scripts/microbench/hmma_tf32_rate.cu) = 4x the multiply-add rate of FFMA2, 1.3x after the three-way split.  The
// staged stream is not what the passes wait for: their mbarrier waits add up to 1.3 us of 83 us, and a bulk L2 prefetch
// of the whole design block ahead of each pass changed nothing (0.2915 -> 0.2938 ms at 10,000 lanes; removed).
// hi = x truncated to TF32 (the tensor core reads the upper 19 bits of an operand), lo = x - hi exactly (13 significant
// bits, of which the tensor core keeps 11): two instructions.  (cvt.rna.tf32.f32 has no hardware form on sm_100a -- it
// compiled to a six-instruction integer sequence per value and made the split cost more than the FFMA2 it replaced.)
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
// d += A (16x8, rows 8..15 zero) * B (8x8); a0 / a2 = A[groupID][tig], A[groupID][tig + 4]; b0 / b1 = B[tig][groupID], B[tig + 4][groupID]
__device__ __forceinline__ void mma_tf32_m8(float (&d)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}
// d += A (16x8) * B (8x8), all four A registers: a0 / a1 = A[groupID (+ 8)][tig], a2 / a3 = A[groupID (+ 8)][tig + 4]
__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma3_tf32_m8(float (&d)[4], uint32_t ah0, uint32_t al0, uint32_t ah2, uint32_t al2, float b0,
                                             float b1) {
  uint32_t bh0, bl0, bh1, bl1;
  tf32_split(b0, bh0, bl0);
  tf32_split(b1, bh1, bl1);
  mma_tf32_m8(d, al0, al2, bh0, bh1);      // small terms first
  mma_tf32_m8(d, ah0, ah2, bl0, bl1);
  mma_tf32_m8(d, ah0, ah2, bh0, bh1);
}

// (A two-warp build of the 4-sub-thread kernel -- regression passes in a helper warp, concurrent with the sweeps: the
// residual pass ahead of the filter, the beta_bar pass trailing the reverse sweep in descending time order, 112
// registers, nine 64-thread CTAs per SM -- was built and measured: one CTA alone 0.148 -> 0.120 ms, but 5,000 lanes
// 0.195 -> 0.218, 8,000 lanes 0.198 -> 0.283, 10,000 lanes 0.272 -> 0.367 ms: with two main warps per scheduler the
// helper warps take issue slots from the sweeps, which is where the time is.  Not kept; the thread-per-lane kernel,
// whose lanes are whole threads, does profit from a helper warp between 13k and 28k lanes: ci_eval_ws.cu.)
// register budget: each SM sub-partition owns 16 K registers, so 96 registers = 5 warps per sub-partition = 20 per SM
// (10k lanes at TPL = 8 are 2500 warps = 17 per SM: one resident wave); 128 registers = 16 warps per SM at TPL = 4
template <int S, int TPL>
__global__ void __maxnreg__(TPL == 8 ? 96 : 128)
k_eval_coop(ci_layout L, const float* __restrict__ XY, const float* __restrict__ sconst, const float* u,
            float* logp, float* grad, float* __restrict__ tape, LeapArgs lf, int nst, int use_mma) {
  extern __shared__ __align__(16) float smem[];
  using Lane = CoopLane<S, TPL>;
  constexpr int LPW = 32 / TPL, KQM = 64 / TPL;
  constexpr int RW = LPW * 8;                                    // floats per warp and step of the tape / an exchange row
  constexpr int NRB = 3;                                         // tape ring: blocks of S steps (one season cycle)
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
  const int RS = coop_row_floats(T, LPW);
  const int smax = coop_series_max(L.G, LPW);
  // shared memory: residual rows | region U | mbarriers.  U holds the nst design-matrix stages during passes 1 and 4
  // and the exchange rows + tape ring during passes 2 and 3 (the passes never overlap).
  float* rbl = smem + lw * RS;                                  // residual / innovation / rbar row of this lane
  float* sx = smem + LPW * RS;                                  // design staging buffers [nst][smax][stride]
  float* xch = sx;                                              // reverse-sweep exchange rows [2][RW]
  float* ring = xch + 2 * RW;                                   // tape ring [NRB][S][RW]
  const int stage_floats = smax * stride;
  const int ufl = max(nst * stage_floats, (2 + NRB * S) * RW);
  unsigned char* wbase = reinterpret_cast<unsigned char*>(sx + (ufl + 3) / 4 * 4);
  const uint32_t bar0 = smem_u32(wbase);                        // mbarriers: 2 design-stage halves, NRB tape blocks

  // ---- TMA staging of the warp's run of series: chunk tc -> stage tc & 1 ------------------------------------------
  const int n_lo = (int)(bw0 / L.G);
  const long b_last = (bw0 + LPW - 1 < B ? bw0 + LPW - 1 : B - 1);
  const int ns = (int)(b_last / L.G) - n_lo + 1;
  const float* lane_x = sx + (long)(n - n_lo) * stride;
  const float* xsrc = XY + (long)n_lo * stride;                 // chunk 0 of the warp's series run
  const long xcs = (long)N * stride;                            // floats between consecutive chunks
  const uint32_t xbytes = (uint32_t)ns * (uint32_t)stride * 4u;
  const uint32_t sx_u32 = smem_u32(sx);
  // the nst = 2 h stages form two halves of h chunks; one mbarrier and one issue block per half
  const int xh = nst >> 1;
  const int nhalf = (nchunk + xh - 1) / xh;
  uint32_t xpar = 0u;                                           // bit p: parity to wait for on half p
  auto x_issue_half = [&](int hb, int p) {                      // thread 0
    const int tc0 = hb * xh;
    const int cnt = (nchunk - tc0) < xh ? (nchunk - tc0) : xh;
    const uint32_t bar = bar0 + 8u * (uint32_t)p;
    bulk_expect(bar, xbytes * (uint32_t)cnt);
    for (int i = 0; i < cnt; ++i)
      bulk_copy_in(sx_u32 + (uint32_t)((p * xh + i) * stage_floats) * 4u, xsrc + (long)(tc0 + i) * xcs, xbytes, bar);
  };
#ifdef CI_COOP_TIMING
  unsigned long long xwait_ = 0;
#endif
  auto x_wait_half = [&](int p) {
#ifdef CI_COOP_TIMING
    unsigned long long w0_, w1_;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(w0_));
#endif
    mbar_wait(bar0 + 8u * (uint32_t)p, (xpar >> p) & 1u);
    xpar ^= 1u << p;
#ifdef CI_COOP_TIMING
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(w1_));
    xwait_ += w1_ - w0_;
#endif
  };
  if (tid == 0) {
#pragma unroll
    for (int q = 0; q < 2 + NRB; ++q) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * q) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (tid == 0) {
    x_issue_half(0, 0);
    if (nhalf > 1) x_issue_half(1, 1);
  }
  // shared-memory offsets of this sub-thread's covariate rows inside a staged chunk (rows past k alias row 0)
  int xoff[KQM];
#pragma unroll
  for (int q = 0; q < KQM; ++q) xoff[q] = ((sub + TPL * q) < k ? (sub + TPL * q) : 0) * CI_TC;

  // ---- u -> sigma^2's, -beta slices, prior + log-det (ci_params.cuh arithmetic, covariates split over sub-threads)
  const float* ub = u + b;
  const float* sc = sconst + n;
  const float sd = sc[SC_SD * N];
  const int Pn = num_params(k, S);
  float R, ql, qd, lp0;
  float nbv[KQM];
  {
    const SoftplusPack p0 = softplus_pack(ub[0]), p1 = softplus_pack(ub[B]);
    const float so = sd * p0.sp, sl = sd * p1.sp;
    R = so * so;
    ql = sl * sl;
    float lps = sc[SC_PRIOR_CONST * N];
    lps += -(2.f * CI_PRIOR_A + 1.f) * logf(so) - sc[SC_OBS_B * N] / R + p0.lsig;
    lps += -(2.f * CI_PRIOR_A + 1.f) * logf(sl) - sc[SC_LEVEL_B * N] / ql + p1.lsig;
    float lpp = 0.f;
    float Gs = 0.f;
    if (k > 0) {
      const SoftplusPack pg = softplus_pack(ub[2 * B]), pn = softplus_pack(ub[3 * B]);
      const float gsv = pg.sp, gsn = pn.sp;
      lps += -1.5f * logf(gsv) - 0.5f / gsv - 0.5f * gsn * gsn + pg.lsig + pn.lsig;
      Gs = -gsn * sqrtf(gsv) * CI_WEIGHTS_PRIOR_SCALE;                // registers hold MINUS beta
    }
#pragma unroll
    for (int q = 0; q < KQM; ++q) {
      const int j = sub + TPL * q;
      float nb = 0.f;
      if (j < k) {
        const float a = ub[(long)(4 + j) * B], bb = ub[(long)(4 + k + j) * B], wn = ub[(long)(4 + 2 * k + j) * B];
        const SoftplusPack pa = softplus_pack(a), pb = softplus_pack(bb);
        const float lsv = pa.sp, lsn = pb.sp;
        lpp += -1.5f * logf(lsv) - 0.5f * rcp(lsv) - 0.5f * lsn * lsn - 0.5f * wn * wn + pa.lsig + pb.lsig;
        nb = wn * lsn * sqrtf(lsv) * Gs;
      }
      nbv[q] = nb;
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

#ifdef CI_COOP_TIMING
  unsigned long long tm_[6];
#define CI_TM(i) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm_[i]))
#else
#define CI_TM(i)
#endif
  CI_TM(0);
  // ---- pass 1: residuals of every step -> shared memory -------------------------------------------------------
  // tensor-core path: TPL = 4, one series per warp, halves of an even number of chunks (host: coop_use_mma)
  const bool mma = (TPL == 4) && (use_mma & 1) != 0;
  if constexpr (TPL == 4) {
   if (mma) {
    const int KK = (k + 7) >> 3;                                       // k-steps of 8 covariates
    // NG groups of 8 steps at a time, three accumulators per group (one per term of the split): the tensor-core
    // instructions issue in order, so consecutive ones must not depend on each other
    auto groups = [&](auto ng_tag, int tc, const float* __restrict__ st0) {
      constexpr int NG = decltype(ng_tag)::value;
      float d[NG][3][4];
#pragma unroll
      for (int g = 0; g < NG; ++g)
#pragma unroll
        for (int e = 0; e < 3; ++e) d[g][e][0] = d[g][e][1] = d[g][e][2] = d[g][e][3] = 0.f;
      // B[j][t]: t = 8-step column lw -> chunk tc + 2 g + (lw >> 2), position lw & 3
      const float* __restrict__ xb_ = st0 + (lw >> 2) * stage_floats + (lw & 3);
#pragma unroll
      for (int kk = 0; kk < KQM / 2; ++kk) {
        if (kk < KK) {
          uint32_t ah0, al0, ah2, al2, bh0[NG], bl0[NG], bh1[NG], bl1[NG];
          tf32_split(nbv[2 * kk], ah0, al0);
          tf32_split(nbv[2 * kk + 1], ah2, al2);
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            tf32_split(xb_[2 * g * stage_floats + xoff[2 * kk]], bh0[g], bl0[g]);
            tf32_split(xb_[2 * g * stage_floats + xoff[2 * kk + 1]], bh1[g], bl1[g]);
          }
#pragma unroll
          for (int g = 0; g < NG; ++g) mma_tf32_m8(d[g][0], al0, al2, bh0[g], bh1[g]);
#pragma unroll
          for (int g = 0; g < NG; ++g) mma_tf32_m8(d[g][1], ah0, ah2, bl0[g], bl1[g]);
#pragma unroll
          for (int g = 0; g < NG; ++g) mma_tf32_m8(d[g][2], ah0, ah2, bh0[g], bh1[g]);
        }
      }
      // C[lw][2 sub], C[lw][2 sub + 1] = -X beta at steps (tc + 2 g) * 4 + 2 sub (+ 1) of this thread's own lane
      const int tl = 2 * sub;
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const int tcg = tc + 2 * g;
        const float2 yv = *reinterpret_cast<const float2*>(st0 + (2 * g + (tl >> 2)) * stage_floats + k * CI_TC + (tl & 3));
        if (tcg + (tl >> 2) < nchunk)
          *reinterpret_cast<float2*>(rbl + tcg * CI_TC + tl) =
              make_float2(yv.x + ((d[g][0][0] + d[g][1][0]) + d[g][2][0]), yv.y + ((d[g][0][1] + d[g][1][1]) + d[g][2][1]));
      }
    };
    // 16 steps at a time with the STEPS on the M axis (no padding rows: half the tensor-core instructions of two
    // 8-step groups): A[t][j] = X (rows t = lw, lw + 8), B[j][chain] = the thread's own -beta slices, and
    // C[t][chain] lands in ANOTHER lane's row -- conflict-free scattered stores (bank = 8 sub + lw)
    auto group16 = [&](int tc, const float* __restrict__ st0) {
      float d[3][4];
#pragma unroll
      for (int e = 0; e < 3; ++e) d[e][0] = d[e][1] = d[e][2] = d[e][3] = 0.f;
      const float* __restrict__ xa = st0 + (lw >> 2) * stage_floats + (lw & 3);   // step lw; step lw + 8 is two chunks on
      const int s2 = 2 * stage_floats;
#pragma unroll
      for (int kk = 0; kk < KQM / 2; ++kk) {
        if (kk < KK) {
          uint32_t bh0, bl0, bh1, bl1, ah[4], al[4];
          tf32_split(nbv[2 * kk], bh0, bl0);
          tf32_split(nbv[2 * kk + 1], bh1, bl1);
          tf32_split(xa[xoff[2 * kk]], ah[0], al[0]);
          tf32_split(xa[s2 + xoff[2 * kk]], ah[1], al[1]);
          tf32_split(xa[xoff[2 * kk + 1]], ah[2], al[2]);
          tf32_split(xa[s2 + xoff[2 * kk + 1]], ah[3], al[3]);
          mma_tf32(d[0], al[0], al[1], al[2], al[3], bh0, bh1);
          mma_tf32(d[1], ah[0], ah[1], ah[2], ah[3], bl0, bl1);
          mma_tf32(d[2], ah[0], ah[1], ah[2], ah[3], bh0, bh1);
        }
      }
      const float y0 = xa[k * CI_TC], y1 = xa[s2 + k * CI_TC];          // centred series at steps lw, lw + 8
      float* r0 = smem + (2 * sub) * RS + tc * CI_TC + lw;              // row of chain 2 sub, step lw of the group
      if (tc + (lw >> 2) < nchunk) {
        r0[0] = y0 + ((d[0][0] + d[1][0]) + d[2][0]);
        r0[RS] = y0 + ((d[0][1] + d[1][1]) + d[2][1]);
      }
      if (tc + 2 + (lw >> 2) < nchunk) {
        r0[8] = y1 + ((d[0][2] + d[1][2]) + d[2][2]);
        r0[RS + 8] = y1 + ((d[0][3] + d[1][3]) + d[2][3]);
      }
    };
    for (int hb = 0; hb < nhalf; ++hb) {
      const int hp = hb & 1;
      x_wait_half(hp);
      const int tce = (hb + 1) * xh < nchunk ? (hb + 1) * xh : nchunk;
      int tc = hb * xh;
      for (; tc + 3 < tce; tc += 4)                                    // 16 steps = chunks tc .. tc + 3
        group16(tc, sx + (hp * xh + (tc - hb * xh)) * stage_floats);
      for (; tc < tce; tc += 2)                                        // 8 steps = chunks tc, tc + 1
        groups(std::integral_constant<int, 1>{}, tc, sx + (hp * xh + (tc - hb * xh)) * stage_floats);
      __syncwarp();                                                    // the staged half is consumed
      if (tid == 0 && hb + 2 < nhalf) x_issue_half(hb + 2, hp);
    }
   }
  }
  if (!mma)
  for (int hb = 0; hb < nhalf; ++hb) {
   const int hp = hb & 1;
   x_wait_half(hp);
   const int tce = (hb + 1) * xh < nchunk ? (hb + 1) * xh : nchunk;
   for (int tc = hb * xh; tc < tce; ++tc) {
    const float* __restrict__ xp = lane_x + (hp * xh + (tc - hb * xh)) * stage_floats;
    pf2 r01 = mk2(0.f, 0.f), r23 = mk2(0.f, 0.f), s01 = mk2(0.f, 0.f), s23 = mk2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < KQM; q += 2) {
      if (q < KQ) {                                                   // pairs of covariates: half the uniform branches
        const float4 x0 = *reinterpret_cast<const float4*>(xp + xoff[q]);
        const float4 x1 = *reinterpret_cast<const float4*>(xp + xoff[q + 1]);
        const float n0 = nbv[q], n1 = nbv[q + 1];                     // 0 past k
        r01 = ffma2(mk2(n0, n0), mk2(x0.x, x0.y), r01);
        r23 = ffma2(mk2(n0, n0), mk2(x0.z, x0.w), r23);
        s01 = ffma2(mk2(n1, n1), mk2(x1.x, x1.y), s01);
        s23 = ffma2(mk2(n1, n1), mk2(x1.z, x1.w), s23);
      }
    }
    r01 = fadd2(r01, s01);
    r23 = fadd2(r23, s23);
    const float4 yv = *reinterpret_cast<const float4*>(xp + k * CI_TC);   // row k: centred series
    // recursive halving over the TPL sub-threads: every sub-thread ends with one of the 4 step sums
    float val;
    int idx;
    if constexpr (TPL == 8) {
      const bool h4 = sub & 4, h2 = sub & 2;
      pf2 keep = h4 ? r23 : r01;
      const pf2 send = h4 ? r01 : r23;
      keep.x += __shfl_xor_sync(0xffffffffu, send.x, 4, 8);
      keep.y += __shfl_xor_sync(0xffffffffu, send.y, 4, 8);
      val = h2 ? keep.y : keep.x;
      val += __shfl_xor_sync(0xffffffffu, h2 ? keep.x : keep.y, 2, 8);
      val += __shfl_xor_sync(0xffffffffu, val, 1, 8);
      idx = (h4 ? 2 : 0) + (h2 ? 1 : 0);
      val += h4 ? (h2 ? yv.w : yv.z) : (h2 ? yv.y : yv.x);
    } else {
      const bool h2 = sub & 2, h1 = sub & 1;
      pf2 keep = h2 ? r23 : r01;
      const pf2 send = h2 ? r01 : r23;
      keep.x += __shfl_xor_sync(0xffffffffu, send.x, 2, 4);
      keep.y += __shfl_xor_sync(0xffffffffu, send.y, 2, 4);
      val = h1 ? keep.y : keep.x;
      val += __shfl_xor_sync(0xffffffffu, h1 ? keep.x : keep.y, 1, 4);
      idx = (h2 ? 2 : 0) + (h1 ? 1 : 0);
      val += h2 ? (h1 ? yv.w : yv.z) : (h1 ? yv.y : yv.x);
    }
    rbl[tc * CI_TC + idx] = val;
   }
   __syncwarp();                                                      // the staged half is consumed
   if (tid == 0 && hb + 2 < nhalf) x_issue_half(hb + 2, hp);
  }

  CI_TM(1);
  Lane ln(sub, R, ql, qd);
  ln.init_forward(sc[SC_M0 * N], sc[SC_P0 * N]);
  // tape [warps][T][RW]: a warp's entries of consecutive steps are contiguous, S steps = one bulk copy
  float* wtape = tape + (long)blockIdx.x * T * RW;
  const int nblk = (T + S - 1) / S;
  const uint32_t ring_u32 = smem_u32(ring);

  // ---- pass 2: forward filter, one season cycle per block, gains -> tape ring -> bulk store ----------------------
  {
    auto fwd_block = [&](int blk, auto full_tag) {
      constexpr bool FULL = decltype(full_tag)::value;
      const int t0 = blk * S;
      const int nrow = FULL ? S : T - t0;
      const int rbuf = blk & 1;
      if (blk >= 2) {                                                 // the store of block blk - 2 has read its rows
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
      }
      float* rows = ring + rbuf * (S * RW) + lw * 8;
      float rr[S];
#pragma unroll
      for (int c = 0; c < S; ++c) rr[c] = (FULL || c < nrow) ? rbl[t0 + c] : 0.f;
      static_up<0, S>([&](auto cc) {
        constexpr int C = decltype(cc)::value;
        if (FULL || C < nrow) {
          const float vout = ln.template fwd<C>(rr[C], rows + C * RW);
          if (sub == 0) rbl[t0 + C] = vout;
        }
      });
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // ring rows -> visible to the bulk store
      __syncwarp();
      if (tid == 0)
        bulk_store(wtape + (long)t0 * RW, ring_u32 + (uint32_t)(rbuf * S * RW) * 4u, (uint32_t)(nrow * RW) * 4u);
    };
    const int nfull = T / S;
    for (int blk = 0; blk < nfull; ++blk) fwd_block(blk, std::true_type{});
    if (nfull < nblk) fwd_block(nfull, std::false_type{});
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // tape complete before it is re-read
    __syncwarp();
  }
  const float ll = -0.5f * (fmaf(ln.acc, 0.6931471805599453f, ln.quad) + (float)(T - ln.nmask) * CI_LOG_2PI);
  if (sub == 0)
    for (int t = T; t < nchunk * CI_TC; ++t) rbl[t] = 0.f;            // steps past T carry no adjoint

  CI_TM(2);
  // ---- pass 3: reverse sweep, tape blocks stream back NRB cycles ahead ---------------------------------------
  ln.init_reverse();
  {
    uint32_t tpar = 0u;                                               // bit q: parity to wait for on ring buffer q
    auto t_issue = [&](int blk, int q) {                              // thread 0
      const int t0 = blk * S;
      const int nrow = (T - t0) < S ? (T - t0) : S;
      bulk_load(ring_u32 + (uint32_t)(q * S * RW) * 4u, wtape + (long)t0 * RW, (uint32_t)(nrow * RW) * 4u,
                bar0 + 8u * (uint32_t)(2 + q));
    };
    if (tid == 0) {
#pragma unroll
      for (int q = 0; q < NRB; ++q)
        if (nblk - 1 - q >= 0) t_issue(nblk - 1 - q, q);
    }
    int q = 0;
    auto bwd_block = [&](int blk, auto full_tag) {
      constexpr bool FULL = decltype(full_tag)::value;
      const int t0 = blk * S;
      const int nrow = FULL ? S : T - t0;
      mbar_wait(bar0 + 8u * (uint32_t)(2 + q), (tpar >> q) & 1u);
      tpar ^= 1u << q;
      const float* rows = ring + q * (S * RW) + lw * 8;
      float vr[S];
#pragma unroll
      for (int c = 0; c < S; ++c) vr[c] = (FULL || c < nrow) ? rbl[t0 + c] : 0.f;
      static_down<S - 1>([&](auto cc) {
        constexpr int C = decltype(cc)::value;
        if (FULL || C < nrow) {
          const float4 ea = *reinterpret_cast<const float4*>(rows + C * RW);
          const float4 eb = *reinterpret_cast<const float4*>(rows + C * RW + 4);
          const float vb = ln.template bwd<C>(ea, eb, vr[C], xch + (C & 1) * RW + lw * 8);
          if (sub == 0) rbl[t0 + C] = vb;
        }
      });
      __syncwarp();                                                   // ring buffer q and both exchange rows are free
      if (tid == 0 && blk - NRB >= 0) t_issue(blk - NRB, q);
      q = (q == NRB - 1) ? 0 : q + 1;
    };
    int blk = nblk - 1;
    if (T % S != 0) bwd_block(blk--, std::false_type{});
    for (; blk >= 0; --blk) bwd_block(blk, std::true_type{});
  }
  const float Rb = ln.Rb;
  const float qlb = __shfl_sync(0xffffffffu, ln.qlacc, 0, TPL);       // row 0 lives in sub-thread 0
  const float qdb = group_sum<TPL>(ln.qdacc);
  __syncwarp();                                                       // rbar rows complete

  CI_TM(3);
  // ---- pass 4: beta_bar_j = -sum_t rbar_t X_tj ---------------------------------------------------------------------
  pf2 bacc[KQM];
#pragma unroll
  for (int q = 0; q < KQM; ++q) bacc[q] = mk2(0.f, 0.f);
  if (k > 0) {
    if (tid == 0) {
      x_issue_half(0, 0);
      if (nhalf > 1) x_issue_half(1, 1);
    }
    bool done = false;
    if constexpr (TPL == 4) {
     if (mma) {
      // beta_bar[j][chain] = -sum_t X[t][j] rbar[chain][t] with the COVARIATES on the M axis (16 per tile, no padding
      // rows): A[j][t] = X (j = jt * 16 + lw, + 8; t = sub, sub + 4), B[t][chain] = the thread's own rbar row entries,
      // accumulators C[j][chain 2 sub (+ 1)] -- 3 x ceil(k / 16) tensor-core instructions per 8 steps instead of
      // 3 x ceil(k / 8)
      constexpr int NJT = KQM / 4;                                     // tiles of 16 covariates (k <= 64)
      const int JT = (k + 15) >> 4;
      float acc4[NJT][4];
#pragma unroll
      for (int jt = 0; jt < NJT; ++jt) acc4[jt][0] = acc4[jt][1] = acc4[jt][2] = acc4[jt][3] = 0.f;
      // rows past k alias row 0 (finite values; their accumulators are never read)
      int aoff0[NJT], aoff1[NJT];
#pragma unroll
      for (int jt = 0; jt < NJT; ++jt) {
        aoff0[jt] = ((jt * 16 + lw) < k ? (jt * 16 + lw) : 0) * CI_TC + sub;
        aoff1[jt] = ((jt * 16 + 8 + lw) < k ? (jt * 16 + 8 + lw) : 0) * CI_TC + sub;
      }
      for (int hb = 0; hb < nhalf; ++hb) {
        const int hp = hb & 1;
        x_wait_half(hp);
        const int tce = (hb + 1) * xh < nchunk ? (hb + 1) * xh : nchunk;
        for (int tc = hb * xh; tc < tce; tc += 2) {                    // 8 steps = chunks tc, tc + 1
          const float* __restrict__ st0 = sx + (hp * xh + (tc - hb * xh)) * stage_floats;
          const bool two = tc + 1 < nchunk;                            // a ragged last group has one chunk only
          uint32_t bh0, bl0, bh1, bl1, ah[NJT][4], al[NJT][4];
          tf32_split(rbl[tc * CI_TC + sub], bh0, bl0);
          tf32_split(two ? rbl[tc * CI_TC + 4 + sub] : 0.f, bh1, bl1);
#pragma unroll
          for (int jt = 0; jt < NJT; ++jt) {
            if (jt < JT) {
              tf32_split(st0[aoff0[jt]], ah[jt][0], al[jt][0]);
              tf32_split(st0[aoff1[jt]], ah[jt][1], al[jt][1]);
              tf32_split(two ? st0[stage_floats + aoff0[jt]] : 0.f, ah[jt][2], al[jt][2]);
              tf32_split(two ? st0[stage_floats + aoff1[jt]] : 0.f, ah[jt][3], al[jt][3]);
            }
          }
          // term-major order: consecutive tensor-core instructions belong to different covariate tiles
#pragma unroll
          for (int jt = 0; jt < NJT; ++jt)
            if (jt < JT) mma_tf32(acc4[jt], al[jt][0], al[jt][1], al[jt][2], al[jt][3], bh0, bh1);
#pragma unroll
          for (int jt = 0; jt < NJT; ++jt)
            if (jt < JT) mma_tf32(acc4[jt], ah[jt][0], ah[jt][1], ah[jt][2], ah[jt][3], bl0, bl1);
#pragma unroll
          for (int jt = 0; jt < NJT; ++jt)
            if (jt < JT) mma_tf32(acc4[jt], ah[jt][0], ah[jt][1], ah[jt][2], ah[jt][3], bh0, bh1);
        }
        __syncwarp();
        if (tid == 0 && hb + 2 < nhalf) x_issue_half(hb + 2, hp);
      }
      // C[jt * 16 + lw (+ 8)][chain 2 sub (+ 1)] -> the [lane][sub + 4 q] ownership of the epilogue, through region U
      float* ex = sx;                                                  // [8 lanes][64] floats (the stages are consumed)
#pragma unroll
      for (int jt = 0; jt < NJT; ++jt) {
        const int j0 = jt * 16 + lw;
        ex[(2 * sub) * 64 + j0] = acc4[jt][0];
        ex[(2 * sub + 1) * 64 + j0] = acc4[jt][1];
        ex[(2 * sub) * 64 + j0 + 8] = acc4[jt][2];
        ex[(2 * sub + 1) * 64 + j0 + 8] = acc4[jt][3];
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < KQM; ++q) bacc[q] = mk2(ex[lw * 64 + sub + TPL * q], 0.f);
      done = true;
     }
    }
    if (!done)
    for (int hb = 0; hb < nhalf; ++hb) {
      const int hp = hb & 1;
      x_wait_half(hp);
      const int tce = (hb + 1) * xh < nchunk ? (hb + 1) * xh : nchunk;
      for (int tc = hb * xh; tc < tce; ++tc) {
        const float* __restrict__ xp = lane_x + (hp * xh + (tc - hb * xh)) * stage_floats;
        const float4 rb = *reinterpret_cast<const float4*>(rbl + tc * CI_TC);
        const pf2 n01 = mk2(rb.x, rb.y), n23 = mk2(rb.z, rb.w);
#pragma unroll
        for (int q = 0; q < KQM; q += 2) {
          if (q < KQ) {
            const float4 x0 = *reinterpret_cast<const float4*>(xp + xoff[q]);
            const float4 x1 = *reinterpret_cast<const float4*>(xp + xoff[q + 1]);
            bacc[q] = ffma2(mk2(x0.x, x0.y), n01, bacc[q]);
            bacc[q + 1] = ffma2(mk2(x1.x, x1.y), n01, bacc[q + 1]);
            bacc[q] = ffma2(mk2(x0.z, x0.w), n23, bacc[q]);
            bacc[q + 1] = ffma2(mk2(x1.z, x1.w), n23, bacc[q + 1]);
          }
        }
      }
      __syncwarp();
      if (tid == 0 && hb + 2 < nhalf) x_issue_half(hb + 2, hp);
    }
  }

  CI_TM(4);
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
    if constexpr (TPL == 4) {
      // Three phases, so that the strided global loads of a phase are all in flight together: grad / the leapfrog arrays
      // may alias u as far as the compiler knows (lf.uu IS u), and a load-compute-store loop per covariate serialised
      // one memory round trip per covariate (the epilogue took 19-35 us of a 280 us evaluation at 10,000 lanes).
      // (A) the raw parameters of every covariate of this sub-thread
      float pa_[KQM], pb_[KQM], pw_[KQM];
  #pragma unroll
      for (int q = 0; q < KQM; ++q) {
        const int j = (sub + TPL * q) < k ? (sub + TPL * q) : 0;       // past k: a valid address, the value is not used
        pa_[q] = ub[(long)(4 + j) * B];
        pb_[q] = ub[(long)(4 + k + j) * B];
        pw_[q] = ub[(long)(4 + 2 * k + j) * B];
      }
      // (B) the three gradient components per covariate (in place of the raw parameters)
  #pragma unroll
      for (int q = 0; q < KQM; ++q) {
        if (sub + TPL * q < k) {
          const float wn = pw_[q];
          const float bb = -(bacc[q].x + bacc[q].y);
          const SoftplusPack pa = softplus_pack(pa_[q]), pb = softplus_pack(pb_[q]);
          const float lsv = pa.sp, lsn = pb.sp;
          const float sq = sqrtf(lsv);
          const float ilsv = rcp(lsv);
          const float l = lsn * sq;
          dG = fmaf(bb, wn * l, dG);
          const float dl = bb * wn * Gs;
          pw_[q] = bb * l * Gs - wn;
          pb_[q] = (dl * sq - lsn) * pb.sig + pb.nsig;
          pa_[q] = (dl * lsn * 0.5f * sq * ilsv + ilsv * (0.5f * ilsv - 1.5f)) * pa.sig + pa.nsig;
        }
      }
      // (C) stores + leapfrog, one parameter block at a time, loads of a half block batched ahead of its stores
      auto emit_block = [&](int p0, const float (&gv)[KQM]) {
  #pragma unroll
        for (int h0 = 0; h0 < KQM; h0 += KQM / 2) {
          float e_[KQM / 2], pp_[KQM / 2], uu_[KQM / 2];
          if (leap) {
  #pragma unroll
            for (int i = 0; i < KQM / 2; ++i) {
              const int j = (sub + TPL * (h0 + i)) < k ? (sub + TPL * (h0 + i)) : 0;
              const long o = (long)(p0 + j) * B + b;
              e_[i] = lf.step0[o];
              pp_[i] = lf.pp[o];
              uu_[i] = lf.last ? 0.f : lf.uu[o];
            }
          }
  #pragma unroll
          for (int i = 0; i < KQM / 2; ++i) {
            const int j = sub + TPL * (h0 + i);
            if (j < k && valid) {
              const long o = (long)(p0 + j) * B + b;
              const float g1 = gv[h0 + i];
              grad[o] = g1;
              if (leap) {
                const float eps = e_[i] * lfac;
                const float ph = fmaf(lc * eps, g1, pp_[i]);
                lf.pp[o] = ph;
                if (!lf.last) lf.uu[o] = fmaf(eps, ph, uu_[i]);
              }
            }
          }
        }
      };
      emit_block(4 + 2 * k, pw_);
      emit_block(4 + k, pb_);
      emit_block(4, pa_);
    } else {   // 8 sub-threads (96-register budget): the per-covariate loop measured faster (0.126 vs 0.130 ms at 1,000 lanes)
  #pragma unroll
      for (int q = 0; q < KQM; ++q) {
        const int j = sub + TPL * q;
        if (j < k) {
          const float a = ub[(long)(4 + j) * B], bq = ub[(long)(4 + k + j) * B], wn = ub[(long)(4 + 2 * k + j) * B];
          const float bb = -(bacc[q].x + bacc[q].y);
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
#ifdef CI_COOP_TIMING
  CI_TM(5);
  if (tid == 0 && (blockIdx.x % CI_COOP_TIMING) == 0)
    { unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid)); printf("blk %d sm %u start %llu xwait(p1+p4) %llu p1 %llu fwd %llu rev %llu p4 %llu epi %llu\n", blockIdx.x, smid, tm_[0] % 100000000ull, xwait_, tm_[1] - tm_[0], tm_[2] - tm_[1], tm_[3] - tm_[2],
           tm_[4] - tm_[3], tm_[5] - tm_[4]); }
#endif
}

// ---- host side ------------------------------------------------------------------------------------------------
static int g_coop_tpl = -1;             // sub-threads per lane: 8, 4; 0 disables the cooperative kernel; -1 = by lane count
int g_coop_mma = 1;                     // ci_set_option("coop_mma", 0 / 1): regression passes on FFMA2 / on the tensor cores
static long g_coop_max_lanes = -1;      // -1 = default crossover to the thread-per-lane kernel

void coop_set_option(int tpl, long max_lanes) {
  g_coop_tpl = tpl;
  g_coop_max_lanes = max_lanes;
}

static size_t coop_smem_bytes(const ci_layout* L, int tpl);
// one resident wave? (CTAs per SM the grid needs) x (shared memory of a CTA + the 1 KB the driver reserves) must fit an SM
static bool coop_one_wave(const ci_layout* L, int tpl) {
  const int lpw = 32 / tpl;
  const long nwarps = ((long)L->N * L->G + lpw - 1) / lpw;
  const long per_sm = (nwarps + 147) / 148;
  return per_sm <= (tpl == 8 ? 20 : 16) && per_sm * (long)(coop_smem_bytes(L, tpl) + 1024) <= 227L * 1024;
}
static int coop_pick_tpl(const ci_layout* L) {
  if (g_coop_tpl >= 0) return g_coop_tpl;
  const long B = (long)L->N * L->G;
  const bool tc = g_coop_mma && L->k > 0 && coop_series_max(L->G, 8) == 1;   // tensor-core passes at 4 sub-threads
  if (B <= (tc ? CI_COOP_TPL8_LANES : CI_COOP_TPL8_LANES_FFMA)) return 8;
  // 4 sub-threads stage 8 lanes' series per chunk: with one lane per series (the reference's default settings) that is
  // 8 x 16 (k + 1) bytes per stage and a second wave from 9,472 lanes on (measured, k = 50, one lane per series, 10,000
  // lanes: 0.598 ms against 0.464 ms at 8 sub-threads; 8,000 lanes: 0.308 against 0.338)
  return coop_one_wave(L, 4) ? 4 : 8;
}

// design-matrix stages that fit beside the residual rows when the SM holds `warps` CTAs (at least 2, at most 16).
// The regression passes are bound by the latency of the staged stream (a warp has half its stages in flight), so a
// grid that leaves SMs partly empty spends the free shared memory on deeper staging: warps = the CTAs an SM really gets.
static int coop_stages(const ci_layout* L, int tpl) {
  const int lpw = 32 / tpl;
  const long nwarps = ((long)L->N * L->G + lpw - 1) / lpw;
  int warps = (int)((nwarps + 147) / 148);
  const int wmax = tpl == 8 ? 17 : 16;
  warps = warps > wmax ? wmax : (warps < 1 ? 1 : warps);
  const long budget = (227L * 1024 - warps * 1024) / warps - (long)lpw * coop_row_floats(L->T, lpw) * 4 - 256;
  const long stage = (long)coop_series_max(L->G, lpw) * (L->k + 1) * CI_TC * 4;
  const long ring = (2 + 3 * (long)L->S) * lpw * 8 * 4;           // the stages alias the tape ring: that much is free
  long room = budget > ring ? budget : ring;
  if (room > 48L * 1024) room = 48L * 1024;                         // the CTA stays far below the 96 KB it may ask for
  long nst = room / stage;
  nst = nst < 2 ? 2 : (nst > CI_COOP_MAX_STAGES ? CI_COOP_MAX_STAGES : nst);
  nst &= ~1L;                                                       // two halves of nst / 2 chunks
  if (tpl == 4 && g_coop_mma && nst >= 4) nst &= ~3L;               // tensor-core passes: halves of whole 8-step groups
  return (int)nst;
}
// tensor-core regression passes: 4 sub-threads per lane, the 8 lanes of every warp are chains of ONE series
static bool coop_use_mma(const ci_layout* L, int tpl) {
  return g_coop_mma && tpl == 4 && L->k > 0 && coop_series_max(L->G, 8) == 1 && coop_stages(L, tpl) % 4 == 0;
}
static size_t coop_smem_bytes(const ci_layout* L, int tpl) {
  const int lpw = 32 / tpl;
  const size_t stages = (size_t)coop_stages(L, tpl) * coop_series_max(L->G, lpw) * (L->k + 1) * CI_TC;
  const size_t ring = (2 + 3 * (size_t)L->S) * lpw * 8;
  return ((size_t)lpw * coop_row_floats(L->T, lpw) + (stages > ring ? stages : ring) + 4) * 4 + 8 * (CI_COOP_MAX_STAGES + 3) + 16;
}

bool coop_eligible(const ci_layout* L) {
  if (g_coop_tpl == 0) return false;
  if (!ci_reg_path(L) || L->r != 1 || L->S < 2 || L->S > CI_MAX_EFF_S || L->k > 64) return false;
  const long B = (long)L->N * L->G;
  const long lim = g_coop_max_lanes >= 0 ? g_coop_max_lanes : CI_COOP_MAX_LANES;
  if (B > lim) return false;
  return coop_smem_bytes(L, coop_pick_tpl(L)) <= 96 * 1024;
}

size_t coop_tape_floats(const ci_layout* L) {
  // [warps][T][lanes per warp * 8]: lanes padded to a whole warp of the widest lane count (8 lanes at TPL = 4)
  const size_t Bp = ((size_t)L->N * L->G + 7) / 8 * 8;
  return (size_t)L->T * Bp * 8;
}

template <int S, int TPL>
static int launch_coop_t(const ci_layout* L, const ci_data* D, const float* d_u, float* d_logp, float* d_grad,
                         float* tape, cudaStream_t st, const LeapArgs& lf) {
  constexpr int LPW = 32 / TPL;
  const long B = (long)L->N * L->G;
  const size_t smem = coop_smem_bytes(L, TPL);
  auto kern = k_eval_coop<S, TPL>;
  static bool configured = false;   // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(96 * 1024));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return fail(CI_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  kern<<<(unsigned)((B + LPW - 1) / LPW), 32, smem, st>>>(*L, D->d_X, D->d_sconst, d_u, d_logp, d_grad, tape, lf,
                                                          coop_stages(L, TPL), coop_use_mma(L, TPL) ? 1 : 0);
  return CI_OK;
}

template <int TPL>
static int launch_coop_s(const ci_layout* L, const ci_data* D, const float* d_u, float* d_logp, float* d_grad,
                         float* tape, cudaStream_t st, const LeapArgs& lf) {
  switch (L->S) {
    case 2: return launch_coop_t<2, TPL>(L, D, d_u, d_logp, d_grad, tape, st, lf);
    case 3: return launch_coop_t<3, TPL>(L, D, d_u, d_logp, d_grad, tape, st, lf);
    case 4: return launch_coop_t<4, TPL>(L, D, d_u, d_logp, d_grad, tape, st, lf);
    case 5: return launch_coop_t<5, TPL>(L, D, d_u, d_logp, d_grad, tape, st, lf);
    case 6: return launch_coop_t<6, TPL>(L, D, d_u, d_logp, d_grad, tape, st, lf);
    case 7: return launch_coop_t<7, TPL>(L, D, d_u, d_logp, d_grad, tape, st, lf);
    default: return fail(CI_ERR_UNSUPPORTED, "cooperative evaluation needs 2 <= nseasons <= 7");
  }
}

int launch_eval_coop(const ci_layout* L, const ci_data* D, const float* d_u, float* d_logp, float* d_grad,
                     float* tape, cudaStream_t st, const LeapArgs& lf) {
  const int tpl = coop_pick_tpl(L);
  const int rc = tpl == 8 ? launch_coop_s<8>(L, D, d_u, d_logp, d_grad, tape, st, lf)
                          : launch_coop_s<4>(L, D, d_u, d_logp, d_grad, tape, st, lf);
  if (rc) return rc;
  CI_CHECK_LAUNCH();
  return CI_OK;
}

}  // namespace ci
