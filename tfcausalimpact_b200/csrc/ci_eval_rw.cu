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
#include "ci_params_gen.cuh"

namespace ci {

__device__ __forceinline__ float rw_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// NW warps per lane; warp w owns the NR consecutive rows w * NR .. (NW * NR >= D, <= 64), lane l the columns l, l + 32
// as one FP32x2 pair.
template <int NW, int NR>
__global__ void __launch_bounds__(NW * 32)
k_eval_rw(ci_layout L, const float* __restrict__ XY, const float* __restrict__ sconst, const float* __restrict__ u,
          float* __restrict__ logp, float* __restrict__ grad, float* __restrict__ tape, float* __restrict__ rscr) {
  constexpr int NT = NW * 32;
  __shared__ __align__(16) float gbuf[2][64];   // forward: g of the step (double buffered); reverse: taped g by index
  __shared__ __align__(16) float colp[NW][64];  // reverse: partial column sums of W g per warp
  __shared__ __align__(16) float gbb[64];       // reverse: gbar by index
  __shared__ float scal[8];
  __shared__ float red[NW];
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
  // ---- ownership: rows i = w * NR + q, columns (l, l + 32); seasonal membership as 0/1 weights -----------
  const int i0 = w * NR;
  float seas_i[NR];
#pragma unroll
  for (int q = 0; q < NR; ++q) seas_i[q] = (i0 + q >= 1 && i0 + q < D) ? 1.f : 0.f;
  const int j0 = l, j1 = l + 32;
  const bool ok0 = j0 < D, ok1 = j1 < D;
  const pf2 seas_j = mk2((j0 >= 1 && ok0) ? 1.f : 0.f, ok1 ? 1.f : 0.f);
  const float wJ = (float)(1 - S);
  pf2 P[NR];
  {
    const float p0 = sconst[SC_P0 * L.N + n];
    const float off = -p0 / (float)S;
#pragma unroll
    for (int q = 0; q < NR; ++q) {
      const int i = i0 + q;
      float v[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int j = l + 32 * c;
        float x = 0.f;
        if (i < D && j < D) {
          if (i == 0 || j == 0) x = (i == j) ? p0 : 0.f;
          else x = (i == j) ? p0 + off : off;
        }
        v[c] = x;
      }
      P[q] = mk2(v[0], v[1]);
    }
  }
  pf2 mr = mk2((j0 == 0) ? sconst[SC_M0 * L.N + n] : 0.f, 0.f);        // mean of the own columns (every warp)
  __syncthreads();                                                       // residuals visible
  float* tp = tape + b * (long)T * (D + 1);
  float acc = 0.f, quad = 0.f;
  int nobs = 0;
  // ---- forward sweep -----------------------------------------------------------------------------------
  {
    int c = 0, ph = 0;                         // season index, position inside the season
    float r_next = rs[0];
    float* tq = tp;
    for (int t = 0; t < T; ++t) {
      const int J = 1 + c, Jl = J & 31;
      const bool Jhi = J >= 32;
      const bool change = ph == r - 1;
      const float resid = r_next;
      if (t + 1 < T) r_next = rs[t + 1];
      const bool obs = (resid == resid);
      // g of the warp's rows: columns 0 and J, broadcast from their owner lanes
      float gi[NR];
#pragma unroll
      for (int q = 0; q < NR; ++q) {
        const float a = __shfl_sync(0xffffffffu, P[q].x, 0);
        const float bq = __shfl_sync(0xffffffffu, Jhi ? P[q].y : P[q].x, Jl);
        gi[q] = a + bq;
      }
      float* gb_ = gbuf[t & 1];
      if (l == 0) {
#pragma unroll
        for (int q = 0; q < NR; ++q) gb_[i0 + q] = gi[q];          // rows past D carry zeros
      }
      const float mh0 = __shfl_sync(0xffffffffu, mr.x, 0);
      const float mhJ = __shfl_sync(0xffffffffu, Jhi ? mr.y : mr.x, Jl);
      __syncthreads();
      const pf2 gj = mk2(gb_[j0], gb_[j1]);
      const float s = gb_[0] + gb_[J] + R;
      const float v = resid - mh0 - mhJ;
      const float is = rcp(s);
      const float nis = obs ? -is : 0.f;
      const float vis = obs ? v * is : 0.f;
      mr = ffma2(gj, mk2(vis, vis), mr);
      // covariance: P -= g g^T / s  (+ qd w w^T on season changes), w = 1 - S e_J on the seasonal block
      const float qq = change ? qd : 0.f;
      const pf2 wj = mk2(qq * (j0 == J ? wJ : seas_j.x), qq * (j1 == J ? wJ : seas_j.y));
      const int qJ = J - i0;                   // row slot of J in this warp (out of range otherwise)
#pragma unroll
      for (int q = 0; q < NR; ++q) {
        const float wi = (q == qJ) ? wJ : seas_i[q];
        const float kg = gi[q] * nis;
        P[q] = ffma2(mk2(wi, wi), wj, ffma2(mk2(kg, kg), gj, P[q]));
      }
      if (tid == 0) P[0].x += ql;
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
  for (int q = 0; q < NR; ++q) P[q] = mk2(0.f, 0.f);           // P now holds W = Pbar
  __syncthreads();
  {
    int c = ((T - 1) / r) % S, ph = (T - 1) % r;
    const float* tq = tp + (long)(T - 1) * (D + 1);
    pf2 cg = mk2(ok0 ? tq[j0] : 0.f, ok1 ? tq[j1] : 0.f);
    float cv = tq[D];
    if (w == 0) {
      gbuf[(T - 1) & 1][j0] = (cv == cv) ? cg.x : 0.f;
      gbuf[(T - 1) & 1][j1] = (cv == cv) ? cg.y : 0.f;
    }
    __syncthreads();
    for (int t = T - 1; t >= 0; --t) {
      const int J = 1 + c;
      const bool change = ph == r - 1;
      pf2 ng = mk2(0.f, 0.f);
      float nv = 0.f;
      if (t > 0) {
        tq -= D + 1;
        ng = mk2(ok0 ? tq[j0] : 0.f, ok1 ? tq[j1] : 0.f);
        nv = tq[D];
      }
      const float v = cv;
      const bool obs = (v == v);
      const pf2 gj = obs ? cg : mk2(0.f, 0.f);
      const float* gT = gbuf[t & 1];
      const float g0 = gT[0], gJ = gT[J];
      float gi[NR];
      if constexpr (NR % 4 == 0) {
#pragma unroll
        for (int q = 0; q < NR; q += 4) {
          const float4 x = *reinterpret_cast<const float4*>(gT + i0 + q);
          gi[q] = x.x; gi[q + 1] = x.y; gi[q + 2] = x.z; gi[q + 3] = x.w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < NR; ++q) gi[q] = gT[i0 + q];
      }
      // adjoint of the time update: ql += W[0][0], qd += w^T W w (per-thread partial, reduced after the sweep)
      if (tid == 0) qlb += P[0].x;
      const pf2 wj = mk2(j0 == J ? wJ : seas_j.x, j1 == J ? wJ : seas_j.y);
      const int qJ = J - i0;
      pf2 cp = mk2(0.f, 0.f), qp = mk2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < NR; ++q) {
        const float wi = (q == qJ) ? wJ : seas_i[q];
        cp = ffma2(P[q], mk2(gi[q], gi[q]), cp);
        qp = ffma2(P[q], mk2(wi, wi), qp);
      }
      if (change) qdacc += fmaf(qp.x, wj.x, qp.y * wj.y);
      colp[w][j0] = cp.x;
      colp[w][j1] = cp.y;
      __syncthreads();                                           // barrier 1
      pf2 col = mk2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < NW; ++q) col = fadd2(col, mk2(colp[q][j0], colp[q][j1]));
      const float a = rw_warp_sum(fmaf(mb.x, gj.x, mb.y * gj.y));
      const float gPg = rw_warp_sum(fmaf(gj.x, col.x, gj.y * col.y));
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
        gbb[j0] = gb.x;
        gbb[j1] = gb.y;
        // the next step's taped gains by index (read after barrier 2)
        float* gN = gbuf[(t - 1) & 1];
        gN[j0] = (nv == nv) ? ng.x : 0.f;
        gN[j1] = (nv == nv) ? ng.y : 0.f;
        if (l == 0) rs[T + t] = vb;
      }
      __syncthreads();                                           // barrier 2
      // W += sym(gbar H^T), H = e_0 + e_J: columns 0 / J of every row, the whole rows 0 / J
      const pf2 cf = mk2(h0 ? 0.5f : 0.f, h1 ? 0.5f : 0.f);
      float gbi[NR];
      if constexpr (NR % 4 == 0) {
#pragma unroll
        for (int q = 0; q < NR; q += 4) {
          const float4 x = *reinterpret_cast<const float4*>(gbb + i0 + q);
          gbi[q] = x.x; gbi[q + 1] = x.y; gbi[q + 2] = x.z; gbi[q + 3] = x.w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < NR; ++q) gbi[q] = gbb[i0 + q];
      }
      const int q0 = -i0;                      // row slot of row 0 in this warp (0 for warp 0, negative otherwise)
#pragma unroll
      for (int q = 0; q < NR; ++q) {
        const float rf = (q == q0 || q == qJ) ? 0.5f : 0.f;
        P[q] = ffma2(cf, mk2(gbi[q], gbi[q]), ffma2(mk2(rf, rf), gb, P[q]));
      }
      mb.x -= h0 ? vb : 0.f;
      mb.y -= h1 ? vb : 0.f;
      cg = ng;
      cv = nv;
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
    const int nr = (Dl + 7) / 8;
    if (nr <= 2) CI_RW(8, 2); else if (nr <= 4) CI_RW(8, 4); else if (nr <= 6) CI_RW(8, 6); else if (nr <= 7) CI_RW(8, 7); else CI_RW(8, 8);
  } else {
    const int nr = (Dl + 3) / 4;
    if (nr <= 4) CI_RW(4, 4); else if (nr <= 8) CI_RW(4, 8); else if (nr <= 12) CI_RW(4, 12); else if (nr <= 14) CI_RW(4, 14); else CI_RW(4, 16);
  }
#undef CI_RW
  return CI_OK;
}

}  // namespace ci
