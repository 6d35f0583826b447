// K1-K3 for FEW lanes with a LARGE latent (BASELINE configs[4]: 100 series x 10,000 steps, nseasons = 52, D = 53):
// one CTA of FOUR WARPS per (series, chain) lane with the covariance and its adjoint RESIDENT IN REGISTERS.
//
// With B << number of SMs a lane must be spread over an SM, and every step is a dense rank-1 update of a D x D
// matrix (~ D^2 FMAs) on a dependent chain T long, so what matters is the latency of ONE step.  The previous
// CTA-per-lane kernel (k_eval_cta, ci_bigd.cu: 512 threads, P in shared memory) spent it on shared-memory round
// trips and 7 block barriers per step pair (17.1 ms per evaluation = 3250 cycles per step pair).  Here:
//   * thread (w, l) -- warp w < 4, lane l -- owns rows i = w + 4 q (q < NR) x columns j = l + 32 c (c < 2) of the
//     FULL symmetric matrix: NR x 2 registers with static indices, nothing of P / W in shared memory
//   * the two columns a step needs (H_t = e_0 + e_J) are broadcast INSIDE each warp with shuffles from the owner
//     lanes (lane 0 and lane J mod 32): g_i for the warp's rows without any memory traffic; the D values are
//     published through a 64-float shared-memory row so that every thread can read g_j of its columns:
//     ONE 128-thread barrier per forward step
//   * reverse sweep: partial column sums of W g over the warp's rows -> 4 partials per column in shared memory
//     (barrier 1), the scalars m.g and g^T W g by warp all-reduce (every warp redundantly: no cross-warp reduction),
//     gbar published by rows (barrier 2), the symmetric update W += sym(gbar H^T) as two predicated FMAs per entry
//   * the drift adjoint w^T W w is accumulated per thread and reduced ONCE after the sweep
//   * seasonal-effects coordinates, any season_duration, default-model and tfp.sts-default priors; LocalLinearTrend
//     and two seasonal components stay on k_eval_cta.
// Tape layout and the hoisted regression passes (residual pre-pass, beta_bar post-pass) are those of k_eval_cta.
// Same math and the same parity tests (tests/test_logp_grad_gpu.py, S = 52 cases).
#include "ci_abi_common.cuh"
#include "ci_eval_lane.cuh"
#include "ci_xstage.cuh"
#include "ci_params_gen.cuh"

namespace ci {

#define RW_RD 8      // reverse sweep: steps of the taped-gain ring in shared memory
#define RW_PF 6      // ... and how many steps ahead of the sweep the ring is filled

__device__ __forceinline__ float rw_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// NW warps per lane; warp w owns the NR = 2 NRP consecutive rows w * NR .. (NW * NR >= D, NR <= SW = 64 / NW), lane l
// the columns l, l + 32.  Registers pair ROWS: P2[p][c] = (P[i0 + 2p][col c], P[i0 + 2p + 1][col c]), so every per-row
// operand of the updates (gains, drift weights, adjoint gains) is loaded from shared memory AS a pair and only the two
// per-column operands need duplicating.  Vectors indexed by state component live in shared memory in SLOT order,
// slot(i) = (i / NR) * SW + i % NR: a warp's rows are one 16-byte aligned run.
// The drift direction w(J) = 1 - S e_J (zero on the level row and the padding rows) comes from a [D][64] table in
// shared memory: no per-row compare / select in the sweeps.
template <int NW, int NRP>
__global__ void __launch_bounds__(NW * 32)
k_eval_rw(ci_layout L, const float* __restrict__ XY, const float* __restrict__ sconst, const float* __restrict__ u,
          float* __restrict__ logp, float* __restrict__ grad, float* __restrict__ tape, float* __restrict__ rscr) {
  constexpr int NT = NW * 32, NR = 2 * NRP, SW = 64 / NW;
  static_assert(NR <= SW, "rows per warp exceed the slot width");
  __shared__ __align__(16) float gbuf[2][64];   // forward: g of the step by slot (double buffered); reverse: taped g by slot
  __shared__ __align__(16) float colp[NW][64];  // reverse: partial column sums of W g per warp (by index)
  __shared__ __align__(16) float gbb[64];       // reverse: gbar by slot
  __shared__ __align__(16) float rg[RW_RD][64]; // reverse: ring of taped gains by slot, filled RW_PF steps ahead by cp.async
  __shared__ float rvv[RW_RD];                  // reverse: ring of taped innovations (NaN = masked step)
  __shared__ __align__(16) float wt[64 * 64];   // wt[J][slot]: drift direction of season index J - 1 (row 0 unused)
  __shared__ float scal[8];
  __shared__ float red[NW];
  __shared__ __align__(16) float gpp[8];                     // reverse: g . (partial column sums) per warp
  extern __shared__ float beta[];               // [max(k, 1)]
  const int S = L.S, D = S + 1, k = L.k, T = L.T, r = L.r;
  const long B = (long)L.N * L.G;
  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  const long b = blockIdx.x;
  const int n = (int)(b / L.G);
  const bool gen = L.flags != 0;
  if (tid == 0) {
    if (!gen) {
      const LaneScalars ls = params_forward_lane(u + b, B, k, S, sconst + n, L.N, beta, 1);
      scal[0] = ls.R; scal[1] = ls.ql; scal[2] = ls.qd; scal[3] = ls.lp0;
    } else {
      const GenScalars ls = params_forward_gen(u + b, B, k, S, L.flags, sconst + n, L.N, beta, 1, L.S2);
      scal[0] = ls.R; scal[1] = ls.ql; scal[2] = ls.qd; scal[3] = ls.lp0;
    }
  }
  const float wJ = (float)(1 - S);
  if (tid < 64) gbuf[0][tid] = gbuf[1][tid] = gbb[tid] = 0.f;   // unused and padding slots stay zero
  for (int e = tid; e < RW_RD * 64; e += NT) (&rg[0][0])[e] = 0.f;
  for (int e = tid; e < D * 64; e += NT) {
    const int J = e >> 6, sl = e & 63;
    const int i = (sl / SW) * NR + (sl % SW);                      // state component of this slot
    const bool real = (sl % SW) < NR && i >= 1 && i < D;
    wt[e] = real ? (i == J ? wJ : 1.f) : 0.f;
  }
  __syncthreads();
  const float R = scal[0], ql = scal[1], qd = scal[2], lp0 = scal[3];
  const int stride = (k + 1) * CI_TC;
  float* rs = rscr + b * (long)(2 * T);       // [T] residuals, then [T] rbar
  // ---- pre-pass: residuals -----------------------------------------------------------------------------
  for (int t = tid; t < T; t += NT) {
    const float* xp = XY + ((long)(t / CI_TC) * L.N + n) * stride + (t % CI_TC);
    float acc = 0.f;
    for (int j = 0; j < k; ++j) acc = fmaf(beta[j], __ldg(xp + j * CI_TC), acc);
    rs[t] = __ldg(xp + k * CI_TC) - acc;
  }
  // ---- ownership: rows i0 + q (q < NR), columns (l, l + 32) ------------------------------------------------
  const int i0 = w * NR;
  const int j0 = l, j1 = l + 32;
  const bool ok0 = j0 < D, ok1 = j1 < D;
  // slots of the own columns (columns past D: slot 63, never written, never used)
  const int sj0 = ok0 ? (j0 / NR) * SW + j0 % NR : 63, sj1 = ok1 ? (j1 / NR) * SW + j1 % NR : 63;
  const float seas0 = (j0 >= 1 && ok0) ? 1.f : 0.f, seas1 = ok1 ? 1.f : 0.f;
  const float* wrow = wt + w * SW;              // + J * 64: the drift weights of the warp's rows
  pf2 P2[NRP][2];
  {
    const float p0 = sconst[SC_P0 * L.N + n];
    const float off = -p0 / (float)S;
#pragma unroll
    for (int p = 0; p < NRP; ++p) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int j = l + 32 * c;
        float v[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = i0 + 2 * p + h;
          float x = 0.f;
          if (i < D && j < D) {
            if (i == 0 || j == 0) x = (i == j) ? p0 : 0.f;
            else x = (i == j) ? p0 + off : off;
          }
          v[h] = x;
        }
        P2[p][c] = mk2(v[0], v[1]);
      }
    }
  }
  pf2 mr = mk2((j0 == 0) ? sconst[SC_M0 * L.N + n] : 0.f, 0.f);        // mean of the own columns (every warp)
  __syncthreads();                                                       // residuals visible
  float* tp = tape + b * (long)T * (D + 1);
  float acc = 0.f, quad = 0.f;
  int nobs = 0;
  // loads of NR consecutive floats starting at a 16-byte aligned shared-memory address, as NRP pairs
  auto load_pairs = [&](const float* src, pf2 (&out)[NRP]) {
#pragma unroll
    for (int p = 0; p + 1 < NRP; p += 2) {
      const float4 x = *reinterpret_cast<const float4*>(src + 2 * p);
      out[p] = mk2(x.x, x.y);
      out[p + 1] = mk2(x.z, x.w);
    }
    if constexpr (NRP & 1) {
      const float2 x = *reinterpret_cast<const float2*>(src + 2 * (NRP - 1));
      out[NRP - 1] = mk2(x.x, x.y);
    }
  };
  // ---- forward sweep -----------------------------------------------------------------------------------
  {
    int c = 0, ph = 0;                         // season index, position inside the season
    float r_next = rs[0];
    float* tq = tp;
    for (int t = 0; t < T; ++t) {
      const int J = 1 + c, Jl = J & 31;
      const bool Jhi = J >= 32;
      const int sJ = (J / NR) * SW + J % NR;
      const bool change = ph == r - 1;
      const float resid = r_next;
      if (t + 1 < T) r_next = rs[t + 1];
      const bool obs = (resid == resid);
      pf2 wi[NRP];
      load_pairs(wrow + J * 64, wi);
      // g of the warp's rows: column J comes from its owner lane, lane 0 adds its column 0 and publishes the run
      pf2 gi[NRP];
      if (!Jhi) {
#pragma unroll
        for (int p = 0; p < NRP; ++p)
          gi[p] = mk2(__shfl_sync(0xffffffffu, P2[p][0].x, Jl), __shfl_sync(0xffffffffu, P2[p][0].y, Jl));
      } else {
#pragma unroll
        for (int p = 0; p < NRP; ++p)
          gi[p] = mk2(__shfl_sync(0xffffffffu, P2[p][1].x, Jl), __shfl_sync(0xffffffffu, P2[p][1].y, Jl));
      }
      float* gb_ = gbuf[t & 1];
      if (l == 0) {
        float* dst = gb_ + w * SW;
#pragma unroll
        for (int p = 0; p + 1 < NRP; p += 2) {
          const pf2 a = fadd2(gi[p], P2[p][0]), bq = fadd2(gi[p + 1], P2[p + 1][0]);
          *reinterpret_cast<float4*>(dst + 2 * p) = make_float4(a.x, a.y, bq.x, bq.y);
        }
        if constexpr (NRP & 1) {
          const pf2 a = fadd2(gi[NRP - 1], P2[NRP - 1][0]);
          *reinterpret_cast<float2*>(dst + 2 * (NRP - 1)) = make_float2(a.x, a.y);
        }
      }
      const float mh0 = __shfl_sync(0xffffffffu, mr.x, 0);
      const float mhJ = __shfl_sync(0xffffffffu, Jhi ? mr.y : mr.x, Jl);
      __syncthreads();
      load_pairs(gb_ + w * SW, gi);
      const float g0 = gb_[sj0], g1 = gb_[sj1];
      const float s = gb_[0] + gb_[sJ] + R;
      const float v = resid - mh0 - mhJ;
      const float is = rcp(s);
      const float nis = obs ? -is : 0.f;
      const float vis = obs ? v * is : 0.f;
      const pf2 gj = mk2(ok0 ? g0 : 0.f, ok1 ? g1 : 0.f);
      mr = ffma2(gj, mk2(vis, vis), mr);
      // covariance: P -= g g^T / s  (+ qd w w^T on season changes), w = 1 - S e_J on the seasonal block
      const float qq = change ? qd : 0.f;
      const float wq0 = qq * (j0 == J ? wJ : seas0), wq1 = qq * (j1 == J ? wJ : seas1);
      const pf2 ga = mk2(gj.x, gj.x), gc = mk2(gj.y, gj.y), wa = mk2(wq0, wq0), wc = mk2(wq1, wq1);
      const pf2 nn = mk2(nis, nis);
#pragma unroll
      for (int p = 0; p < NRP; ++p) {
        const pf2 kg = fmul2(gi[p], nn);
        P2[p][0] = ffma2(wi[p], wa, ffma2(kg, ga, P2[p][0]));
        P2[p][1] = ffma2(wi[p], wc, ffma2(kg, gc, P2[p][1]));
      }
      if (tid == 0) P2[0][0].x += ql;
      if (w == 0) {
        if (ok0) tq[j0] = gj.x;
        if (ok1) tq[j1] = gj.y;
      }
      if (tid == NT - 1) tq[D] = obs ? v : __int_as_float(0x7fc00000);
      tq += D + 1;
      if (obs) {
        acc += fast_log2(s);
        quad = fmaf(vis, v, quad);
        ++nobs;
      }
      if (++ph == r) {
        ph = 0;
        c = (c + 1 == S) ? 0 : c + 1;
      }
    }
  }
  const float ll = -0.5f * (fmaf(acc, 0.6931471805599453f, quad) + (float)nobs * CI_LOG_2PI);

  // ---- reverse sweep -----------------------------------------------------------------------------------
  float Rb = 0.f, qlb = 0.f, qdacc = 0.f;
  pf2 mb = mk2(0.f, 0.f);                      // mean adjoint of the own columns (every warp)
#pragma unroll
  for (int p = 0; p < NRP; ++p) P2[p][0] = P2[p][1] = mk2(0.f, 0.f);   // P2 now holds W = Pbar
  // rf_i = 0.5 on the rows of H = e_0 + e_J, else 0:  rf = crf - (0.5 / S) w_i(J),  crf = 0.5 [i == 0] + (0.5 / S) [i seasonal]
  const float hS = 0.5f / (float)S;
  pf2 crf[NRP];
#pragma unroll
  for (int p = 0; p < NRP; ++p) {
    const int ia = i0 + 2 * p, ib = ia + 1;
    crf[p] = mk2(ia == 0 ? 0.5f : (ia < D ? hS : 0.f), ib < D ? hS : 0.f);
  }
  __syncthreads();
  {
    int c = ((T - 1) / r) % S, ph = (T - 1) % r;
    // The tape (up to ~2 MB per lane) streams back from DRAM: warp 0 copies the gains and the innovation of step
    // t - RW_PF straight into a shared-memory ring while step t runs (cp.async, one group per step), so no global
    // load latency sits inside the two-barrier step.
    const uint32_t rg_u32 = smem_u32(&rg[0][0]), rv_u32 = smem_u32(rvv);
    auto tape_fetch = [&](int ts) {                              // warp 0
      if (ts >= 0) {
        const float* src = tp + (long)ts * (D + 1);
        const uint32_t dst = rg_u32 + (uint32_t)((ts & (RW_RD - 1)) * 64) * 4u;
        if (ok0) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + (uint32_t)sj0 * 4u), "l"(src + j0) : "memory");
        if (ok1) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + (uint32_t)sj1 * 4u), "l"(src + j1) : "memory");
        if (l == 0) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(rv_u32 + (uint32_t)(ts & (RW_RD - 1)) * 4u), "l"(src + D) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (w == 0) {
#pragma unroll
      for (int d = 0; d < RW_PF; ++d) tape_fetch(T - 1 - d);
      asm volatile("cp.async.wait_group %0;" ::"n"(RW_PF - 1) : "memory");
    }
    __syncthreads();
    for (int t = T - 1; t >= 0; --t) {
      const int J = 1 + c;
      const int sJ = (J / NR) * SW + J % NR;
      const bool change = ph == r - 1;
      const float* gT = rg[t & (RW_RD - 1)];
      const float v = rvv[t & (RW_RD - 1)];
      const bool obs = (v == v);
      const pf2 gj = mk2(ok0 ? gT[sj0] : 0.f, ok1 ? gT[sj1] : 0.f);
      if (w == 0) tape_fetch(t - RW_PF);
      const float g0 = gT[0], gJ = gT[sJ];
      pf2 gi[NRP], wi[NRP];
      load_pairs(gT + w * SW, gi);
      load_pairs(wrow + J * 64, wi);
      // the scalar m . g needs nothing of this step: its warp reduction overlaps the column sums
      float a = fmaf(mb.x, gj.x, mb.y * gj.y);
      // adjoint of the time update: ql += W[0][0], qd += w^T W w (per-thread partial, reduced after the sweep)
      if (tid == 0) qlb += P2[0][0].x;
      pf2 ca = mk2(0.f, 0.f), cb = mk2(0.f, 0.f), qa = mk2(0.f, 0.f), qb = mk2(0.f, 0.f);
#pragma unroll
      for (int p = 0; p < NRP; ++p) {
        ca = ffma2(P2[p][0], gi[p], ca);
        cb = ffma2(P2[p][1], gi[p], cb);
        qa = ffma2(P2[p][0], wi[p], qa);
        qb = ffma2(P2[p][1], wi[p], qb);
      }
      const float cp0 = ca.x + ca.y, cp1 = cb.x + cb.y;
      colp[w][j0] = cp0;
      colp[w][j1] = cp1;
      // g^T W g = sum over warps of (g . partial column sums): one warp reduction BEFORE the barrier
      float gp = fmaf(gj.x, cp0, gj.y * cp1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        gp += __shfl_xor_sync(0xffffffffu, gp, o);
      }
      if (l == 0) gpp[w] = gp;
      if (change) {
        const float wj0 = j0 == J ? wJ : seas0, wj1 = j1 == J ? wJ : seas1;
        qdacc += fmaf(qa.x + qa.y, wj0, (qb.x + qb.y) * wj1);
      }
      __syncthreads();                                           // barrier 1
      pf2 col = mk2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < NW; ++q) col = fadd2(col, mk2(colp[q][j0], colp[q][j1]));
      float gPg;
      {
        const float4 x = *reinterpret_cast<const float4*>(gpp);
        gPg = (x.x + x.y) + (x.z + x.w);
        if constexpr (NW == 8) {
          const float4 y = *reinterpret_cast<const float4*>(gpp + 4);
          gPg += (y.x + y.y) + (y.z + y.w);
        }
      }
      const float s = g0 + gJ + R;
      const float is = obs ? rcp(s) : 0.f;
      const float vz = obs ? v : 0.f;
      const float vis = vz * is;
      const float vb = (a - vz) * is;
      const float sbar = obs ? (gPg - a * vz) * is * is - 0.5f * (is - vis * vis) : 0.f;
      Rb += sbar;
      const bool h0 = (j0 == 0) || (j0 == J), h1 = (j1 == J);
      pf2 gb = ffma2(mb, mk2(vis, vis), fmul2(col, mk2(-2.f * is, -2.f * is)));
      gb.x += h0 ? sbar : 0.f;
      gb.y += h1 ? sbar : 0.f;
      if (w == 0) {
        if (ok0) gbb[sj0] = gb.x;
        if (ok1) gbb[sj1] = gb.y;
        if (l == 0) rs[T + t] = vb;
        asm volatile("cp.async.wait_group %0;" ::"n"(RW_PF - 1) : "memory");   // step t - 1 of the ring has landed
      }
      __syncthreads();                                           // barrier 2
      // W += sym(gbar H^T), H = e_0 + e_J: columns 0 / J of every row, the whole rows 0 / J
      pf2 gbi[NRP];
      load_pairs(gbb + w * SW, gbi);
      const float cf0 = h0 ? 0.5f : 0.f, cf1 = h1 ? 0.5f : 0.f;
      const pf2 fa = mk2(cf0, cf0), fc = mk2(cf1, cf1), ba = mk2(gb.x, gb.x), bc = mk2(gb.y, gb.y);
      const pf2 nh = mk2(-hS, -hS);
#pragma unroll
      for (int p = 0; p < NRP; ++p) {
        const pf2 rf = ffma2(wi[p], nh, crf[p]);
        P2[p][0] = ffma2(fa, gbi[p], ffma2(rf, ba, P2[p][0]));
        P2[p][1] = ffma2(fc, gbi[p], ffma2(rf, bc, P2[p][1]));
      }
      mb.x -= h0 ? vb : 0.f;
      mb.y -= h1 ? vb : 0.f;
      if (--ph < 0) {
        ph = r - 1;
        c = (c == 0) ? S - 1 : c - 1;
      }
    }
  }
  // drift adjoint: one block reduction
  {
    const float ws = rw_warp_sum(qdacc);
    if (l == 0) red[w] = ws;
  }
  __syncthreads();
  float qdb = 0.f;
#pragma unroll
  for (int q = 0; q < NW; ++q) qdb += red[q];
  // ---- post-pass: beta_bar_j = -sum_t rbar_t X_tj  (threads = covariates x time slices) -----------------
  if (k > 0) {
    float* pb = &colp[0][0];                  // [slices][k] partial sums (NW * 64 >= NT floats)
    if (k <= NT) {
      const int slices = NT / k;
      const int j = tid % k, sl = tid / k;
      if (sl < slices) {
        float accb = 0.f;
        for (int t = sl; t < T; t += slices) {
          const float* xp = XY + ((long)(t / CI_TC) * L.N + n) * stride + (t % CI_TC);
          accb = fmaf(-rs[T + t], __ldg(xp + j * CI_TC), accb);
        }
        pb[sl * k + j] = accb;
      }
      __syncthreads();
      if (tid < k) {
        float a2 = 0.f;
        for (int q = 0; q < slices; ++q) a2 += pb[q * k + tid];
        beta[tid] = a2;
      }
    } else {
      for (int j = tid; j < k; j += NT) {
        float accb = 0.f;
        for (int t = 0; t < T; ++t) {
          const float* xp = XY + ((long)(t / CI_TC) * L.N + n) * stride + (t % CI_TC);
          accb = fmaf(-rs[T + t], __ldg(xp + j * CI_TC), accb);
        }
        beta[j] = accb;
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    if (!gen) params_backward_lane(u + b, B, k, S, sconst + n, L.N, beta, 1, Rb, qlb, qdb, grad + b);
    else params_backward_gen(u + b, B, k, S, L.flags, sconst + n, L.N, beta, 1, Rb, qlb, 0.f, qdb, grad + b, L.S2, 0.f);
    logp[b] = ll + lp0;
  }
}

// ---- host side ----------------------------------------------------------------------------------------------
bool rw_eligible(const ci_layout* L) {
  const int D = L->S + 1;
  const long B = (long)L->N * L->G;
  return L->S > 1 && L->S2 <= 1 && !(L->flags & CI_FLAG_LLT) && D >= 8 && D <= 64 && B <= 2 * 148 && L->k <= 4096;
}

int g_rw_warps = 4;    // ci_set_option("eval_rw_warps", 4 / 8)

int launch_eval_rw(const ci_layout* L, const ci_data* D, const float* d_u, float* d_logp, float* d_grad, float* tape,
                   float* rscr, cudaStream_t st) {
  const unsigned B = (unsigned)((long)L->N * L->G);
  const int Dl = L->S + 1;
  const size_t smem = (size_t)(L->k > 0 ? L->k : 1) * sizeof(float);
#define CI_RW(NW_, NR_) k_eval_rw<NW_, NR_><<<B, NW_ * 32, smem, st>>>(*L, D->d_X, D->d_sconst, d_u, d_logp, d_grad, tape, rscr)
  if (g_rw_warps == 8) {
    const int nrp = (Dl + 15) / 16;            // row pairs per warp, 8 warps
    if (nrp <= 1) CI_RW(8, 1); else if (nrp == 2) CI_RW(8, 2); else if (nrp == 3) CI_RW(8, 3); else CI_RW(8, 4);
  } else {
    const int nrp = (Dl + 7) / 8;              // row pairs per warp, 4 warps
    if (nrp <= 2) CI_RW(4, 2); else if (nrp <= 4) CI_RW(4, 4); else if (nrp == 5) CI_RW(4, 5); else if (nrp == 6) CI_RW(4, 6);
    else if (nrp == 7) CI_RW(4, 7); else CI_RW(4, 8);
  }
#undef CI_RW
  return CI_OK;
}

}  // namespace ci
