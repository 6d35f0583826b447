// K6 / K7: predictive distributions per posterior draw and Philox-seeded posterior-predictive
// sampling.
//
//   ci_filter_predict   tfp.sts.one_step_predictive  (/root/reference/causalimpact/model.py:414-418)
//                       + the forward_filter half of tfp.sts.forecast (model.py:446-451)
//   ci_onestep_sample   one_step_dist.sample(n) / .mean()     (causalimpact/inferences.py:108,119)
//   ci_forecast_sample  posterior_dist.sample(n) / .mean()    (inferences.py:113,134; main.py:377)
//
// Sampling semantics: a sample first picks one posterior draw j uniformly (mixture
// over draws), then -- one-step: y_t ~ N(mean_jt, var_jt) independently over t; forecast:
// x_0 ~ N(m_j, P_j), then the state-space model is rolled forward Tf steps with transition and
// observation noise (one joint trajectory).  Outputs are written in ORIGINAL units
// (y * sigma + mu, causalimpact/misc.py:75-77) when d_mu_sig is given.
#include "ci_abi_common.cuh"
#include "ci_kalman.cuh"
#include "ci_xstage.cuh"

namespace ci {

struct PredWs {
  float* beta;  // [max(k,1)][B]
  float* scal;  // [8][B]
  float* off;   // [T][B] offsets, then residuals in place
};
static size_t pred_ws_floats(const ci_layout* L, PredWs* ws, Carver* c) {
  const size_t B = (size_t)L->N * L->G;
  const size_t nb = (size_t)(L->k > 0 ? L->k : 1) * B, ns = 8 * B, no = (size_t)L->T * B;
  if (ws && c) {
    ws->beta = c->floats(nb);
    ws->scal = c->floats(ns);
    ws->off = c->floats(no);
  }
  return align_up(nb * 4, 256) + align_up(ns * 4, 256) + align_up(no * 4, 256);
}

struct FcWs {
  float* beta;  // [max(k,1)][B]
  float* scal;  // [8][B]
  float* off;   // [Tf][B] forecast-period offsets per draw
  float* pm;    // [Tf][B] per-draw forecast mean path
};
static size_t fc_ws_floats(const ci_layout* L, FcWs* ws, Carver* c) {
  const size_t B = (size_t)L->N * L->G;
  const size_t nb = (size_t)(L->k > 0 ? L->k : 1) * B, ns = 8 * B, no = (size_t)(L->Tf > 0 ? L->Tf : 1) * B;
  if (ws && c) {
    ws->beta = c->floats(nb);
    ws->scal = c->floats(ns);
    ws->off = c->floats(no);
    ws->pm = c->floats(no);
  }
  return align_up(nb * 4, 256) + align_up(ns * 4, 256) + 2 * align_up(no * 4, 256);
}

// Forward filter per draw lane.  off_resid: in = offset_t (mean + X_t beta), used as scratch.
// pred_mean[t][b] = H m_t|t-1 + offset_t ; pred_var[t][b] = s_t ; fstate = m, P (packed upper),
// lower Cholesky factor of P (packed, row-major lower) of the predicted state for step T.
template <int S>
__global__ void __launch_bounds__(128)
k_filter(int N, int G, int T, int r, const float* __restrict__ y, const float* __restrict__ sconst,
         const float* __restrict__ scal, const float* __restrict__ off, float* __restrict__ pred_mean,
         float* __restrict__ pred_var, float* __restrict__ fstate) {
  constexpr int NP = S * (S + 1) / 2;
  const long B = (long)N * G;
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = (int)(b / G);
  const float R = scal[0 * B + b], ql = scal[1 * B + b], qd = scal[2 * B + b];
  const float* __restrict__ yn = y + (long)n * T;
  KState<S> k;
  kf_init<S>(k, sconst[SC_M0 * N + n], sconst[SC_P0 * N + n]);
  int phase = 0;
  for (int t = 0; t < T; ++t) {
    const float o = off[(long)t * B + b];
    const float yt = __ldg(yn + t);
    float hm = k.m[0], s = k.P[sym<S>(0, 0)] + R;
    if (S > 1) {
      hm += k.m[1];
      s += 2.f * k.P[sym<S>(0, 1)] + k.P[sym<S>(1, 1)];
    }
    pred_mean[(long)t * B + b] = hm + o;
    pred_var[(long)t * B + b] = s;
    if (yt == yt) {
      float g[S], s2, v, is;
      kf_update<S>(k, yt - o, R, g, s2, v, is);
    }
    const bool change = (phase == r - 1);
    kf_predict<S>(k, ql, qd, change);
    phase = change ? 0 : phase + 1;
  }
#pragma unroll
  for (int i = 0; i < S; ++i) fstate[(long)i * B + b] = k.m[i];
#pragma unroll
  for (int i = 0; i < NP; ++i) fstate[(long)(S + i) * B + b] = k.P[i];
  // guarded Cholesky (a non-positive pivot zeroes its column), row-major packed lower factor
  float Lf[NP];
#pragma unroll
  for (int i = 0; i < S; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      float a = k.P[sym<S>(i, j)];
#pragma unroll
      for (int q = 0; q < j; ++q) a = fmaf(-Lf[i * (i + 1) / 2 + q], Lf[j * (j + 1) / 2 + q], a);
      if (i == j) {
        Lf[i * (i + 1) / 2 + j] = a > 0.f ? sqrtf(a) : 0.f;
      } else {
        const float dj = Lf[j * (j + 1) / 2 + j];
        Lf[i * (i + 1) / 2 + j] = dj > 0.f ? a / dj : 0.f;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) fstate[(long)(S + NP + i) * B + b] = Lf[i];
}

// Fused forward filter (K6): ONE kernel per call.  A lane = one (series, posterior draw); the parameter
// transform, the regression offsets and the filter run in the same thread, exactly as the forward half of the
// evaluation kernel (ci_eval.cu): -beta lives in a shared-memory column, the (X, y) chunk of the warp's series
// run is TMA-bulk-copied into shared memory behind a per-warp mbarrier while the previous chunk's steps run, the
// state stays in registers.  Nothing but the outputs goes to HBM (the unfused path wrote and re-read a [T][B]
// offset array and re-derived it in a separate LSU-bound kernel).  Padding lanes shadow the last lane.
__host__ __device__ inline int filter_series_max(int G) {
  const int a = (32 + G - 2) / G + 1;
  return a < 32 ? a : 32;
}
// (The seasonal-effects frame of the evaluation kernel was measured here too: with its 8 x 8 covariance and the
// per-step season dispatch the forward-only filter took 4.16 ms against 3.45 ms for 1M lanes x 365 -- not kept.)
template <int S>
__global__ void __launch_bounds__(32, 16)
k_filter_fused(ci_layout L, const float* __restrict__ XY, const float* __restrict__ sconst, const float* __restrict__ u,
               float* __restrict__ pred_mean, float* __restrict__ pred_var, float* __restrict__ fstate) {
  extern __shared__ __align__(16) float smem[];
  constexpr int NP = S * (S + 1) / 2;
  const int N = L.N, G = L.G, T = L.T, k = L.k, r = L.r;
  const long B = (long)N * G;
  const long bw0 = (long)blockIdx.x * 32;
  const bool valid = bw0 + threadIdx.x < B;
  const long b = valid ? bw0 + threadIdx.x : B - 1;
  const int n = (int)(b / G);
  const int stride = (k + 1) * CI_TC;
  float* sb = smem + threadIdx.x;                         // [k][32]: MINUS beta
  const int smax = filter_series_max(G);
  float* sx = smem + (long)k * 32;                        // [smax][stride]
  unsigned char* wbase = reinterpret_cast<unsigned char*>(sx + (long)smax * stride);
  const int n_lo = (int)(bw0 / G);
  const long b_last = (bw0 + 31 < B ? bw0 + 31 : B - 1);
  const int ns = (int)(b_last / G) - n_lo + 1;
  XStaged xs;
  xs.lane_x = sx + (long)(n - n_lo) * stride;
  xs.bar = smem_u32(wbase);
  xs.parity = 0u;
  if (threadIdx.x == 0) {
    *reinterpret_cast<uint64_t*>(wbase + 16) = reinterpret_cast<uint64_t>(XY + (long)n_lo * stride);
    *reinterpret_cast<uint32_t*>(wbase + 24) = (uint32_t)ns * (uint32_t)stride * 4u;
    *reinterpret_cast<uint32_t*>(wbase + 28) = smem_u32(sx);
    *reinterpret_cast<uint64_t*>(wbase + 32) = (uint64_t)((long)N * stride) * 4u;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xs.bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (threadIdx.x == 0) xs.issue(0);
  const float* sc = sconst + n;
  const LaneScalars ls = params_forward_lane(u + b, B, k, L.S, sc, N, sb, 32, -1.0f);
  const float R = ls.R, ql = ls.ql, qd = ls.qd;
  const float mean = sc[SC_MEAN * N];
  KState<S> ks;
  kf_init<S>(ks, sc[SC_M0 * N], sc[SC_P0 * N]);
  const int nchunk = (T + CI_TC - 1) / CI_TC;
  int phase = 0;
  float* pmp = pred_mean + b;
  float* pvp = pred_var + b;
  for (int tc = 0; tc < nchunk; ++tc) {
    xs.acquire();
    const float* __restrict__ xp = xs.lane_x;
    const float4 yv = *reinterpret_cast<const float4*>(xp + k * CI_TC);          // row k: centred series
    pf2 a01 = mk2(0.f, 0.f), a23 = mk2(0.f, 0.f), c01 = mk2(0.f, 0.f), c23 = mk2(0.f, 0.f);
    int j = 0;
    for (; j + 2 <= k; j += 2) {
      const float n0 = sb[j * 32], n1 = sb[(j + 1) * 32];
      const float4 x0 = *reinterpret_cast<const float4*>(xp + j * CI_TC);
      const float4 x1 = *reinterpret_cast<const float4*>(xp + (j + 1) * CI_TC);
      a01 = ffma2(mk2(n0, n0), mk2(x0.x, x0.y), a01);
      a23 = ffma2(mk2(n0, n0), mk2(x0.z, x0.w), a23);
      c01 = ffma2(mk2(n1, n1), mk2(x1.x, x1.y), c01);
      c23 = ffma2(mk2(n1, n1), mk2(x1.z, x1.w), c23);
    }
    if (j < k) {
      const float n0 = sb[j * 32];
      const float4 x0 = *reinterpret_cast<const float4*>(xp + j * CI_TC);
      a01 = ffma2(mk2(n0, n0), mk2(x0.x, x0.y), a01);
      a23 = ffma2(mk2(n0, n0), mk2(x0.z, x0.w), a23);
    }
    a01 = fadd2(a01, c01);
    a23 = fadd2(a23, c23);
    xs.release(tc + 1 < nchunk ? tc + 1 : -1);
    const float acc[CI_TC] = {a01.x, a01.y, a23.x, a23.y};                       // -X_t . beta
    const float yc[CI_TC] = {yv.x, yv.y, yv.z, yv.w};
    const int nstep = (T - tc * CI_TC) < CI_TC ? (T - tc * CI_TC) : CI_TC;
#pragma unroll
    for (int i = 0; i < CI_TC; ++i) {
      if (i < nstep) {
        const float rr = yc[i] + acc[i];                                         // NaN = missing step
        float hm = ks.m[0], s = ks.P[sym<S>(0, 0)] + R;
        if (S > 1) {
          hm += ks.m[1];
          s += 2.f * ks.P[sym<S>(0, 1)] + ks.P[sym<S>(1, 1)];
        }
        if (valid) {
          *pmp = hm + mean - acc[i];
          *pvp = s;
        }
        pmp += B;
        pvp += B;
        if (rr == rr) {
          float g[S], s2, v, is;
          kf_update<S>(ks, rr, R, g, s2, v, is);
        }
        const bool change = (phase == r - 1);
        kf_predict<S>(ks, ql, qd, change);
        phase = change ? 0 : phase + 1;
      }
    }
  }
  if (!valid) return;
#pragma unroll
  for (int i = 0; i < S; ++i) fstate[(long)i * B + b] = ks.m[i];
#pragma unroll
  for (int i = 0; i < NP; ++i) fstate[(long)(S + i) * B + b] = ks.P[i];
  // guarded Cholesky (a non-positive pivot zeroes its column), row-major packed lower factor
  float Lf[NP];
#pragma unroll
  for (int i = 0; i < S; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      float a = ks.P[sym<S>(i, j)];
#pragma unroll
      for (int q = 0; q < j; ++q) a = fmaf(-Lf[i * (i + 1) / 2 + q], Lf[j * (j + 1) / 2 + q], a);
      if (i == j) {
        Lf[i * (i + 1) / 2 + j] = a > 0.f ? sqrtf(a) : 0.f;
      } else {
        const float dj = Lf[j * (j + 1) / 2 + j];
        Lf[i * (i + 1) / 2 + j] = dj > 0.f ? a / dj : 0.f;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) fstate[(long)(S + NP + i) * B + b] = Lf[i];
}

__device__ __forceinline__ uint32_t pick_draw(uint32_t sample_id, uint32_t stream, int G, uint64_t seed) {
  const float un = philox_uniform4(sample_id, 0u, stream, TAG_PICK, seed).x;
  const int j = (int)(un * (float)G);
  return (uint32_t)(j < G ? j : G - 1);
}

// pre_sims[n][i][t]: 128 samples of one series per CTA, time tiled by 32 through shared memory.  The
// lanes of a warp read the (mean, variance) of their picked draws of the SAME series and step (<= G
// adjacent floats: a few sectors, L1-resident) and the [nsamp][T] rows are written as full 128 B lines.
__global__ void __launch_bounds__(128)
k_onestep_sample(int N, int G, int T, int nsamp, uint64_t seed, const float* __restrict__ pm,
                 const float* __restrict__ pv, const float* __restrict__ mu_sig, float* __restrict__ sims) {
  __shared__ float tile[128][33];
  const int n = blockIdx.y;
  const int i = blockIdx.x * 128 + threadIdx.x;
  const long B = (long)N * G;
  const long si = (long)n * nsamp + (i < nsamp ? i : 0);
  const long b = (long)n * G + pick_draw((uint32_t)si, 0u, G, seed);
  const float mu = mu_sig ? mu_sig[n] : 0.f, sg = mu_sig ? mu_sig[N + n] : 1.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int tn = min(32, T - t0);
#pragma unroll
    for (int q4 = 0; q4 < 8; ++q4) {
      // counter = (sample, time quad): the same stream as one Philox call per 4 consecutive steps
      const f32x4 z = philox_normal4_fast((uint32_t)si, (uint32_t)((t0 >> 2) + q4), 0u, TAG_ONESTEP, seed);
      const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int q = q4 * 4 + e;
        if (q < tn) {
          const long o = (long)(t0 + q) * B + b;
          tile[threadIdx.x][q] = fmaf(fmaf(sqrtf(pv[o]), zz[e], pm[o]), sg, mu);
        }
      }
    }
    __syncthreads();
    for (int row = warp; row < 128; row += 4) {
      const int ii = blockIdx.x * 128 + row;
      if (ii < nsamp && lane < tn) sims[((long)n * nsamp + ii) * T + t0 + lane] = tile[row][lane];
    }
    __syncthreads();
  }
}

// pre_sims[n][i][t], tiled version (G <= ONESTEP_TILED_MAX_G draws).  One CTA = one series x a slab of samples.  The
// (mean, sd) of ALL G draws of the series are staged per block of 32 steps in shared memory, already unstandardised
// (a staged value serves slab / G samples); a warp walks the slab four samples at a time with LANE = TIME STEP: one
// Philox call = four normals, one per sample of the quad; the picked draw is warp-uniform (conflict-free shared-memory
// reads at odd row pitch) and every store instruction writes one 128-byte run of a [nsamp][T] row.
//   * ALIGNED RUNS.  Rows are T floats long, so a run that starts at a multiple of 32 steps is not 128-byte aligned
//     in memory and every store leaves two partially written sectors behind: measured 2.79 ms (T = 365) against
//     1.37 ms (T = 384) for 2000 series x 1000 samples.  Each sample therefore shifts its runs by a_i = (-row start)
//     mod 32: iteration m of the CTA covers t in [a_i + 32 (m - 1), a_i + 32 m), i.e. whole 128-byte lines of the
//     output; the staged window holds the two blocks of 32 steps such a run can touch.
//   * The raw (mean, variance) block of the NEXT 32 steps lands in a second pair of shared-memory arrays by cp.async
//     while the current runs are sampled (the moments are not L2-resident: 584 MB for 2000 series).
// (A first staged variant with 128-sample CTAs re-used a staged value only 1.3 times and lost to the L1-served gathers
// of k_onestep_sample: 5.50 against 4.15 ms for 2000 series x 1000 samples x 365 steps.)
#define ONESTEP_TILED_MAX_G 128
#define ONESTEP_NT 512          // threads per CTA of the tiled sampler
__global__ void __launch_bounds__(ONESTEP_NT)
k_onestep_sample_tiled(int N, int G, int T, int nsamp, int slab, uint64_t seed, const float* __restrict__ pm,
                       const float* __restrict__ pv, const float* __restrict__ mu_sig, float* __restrict__ sims,
                       float* __restrict__ pre_mean) {
  extern __shared__ __align__(16) float smem[];
  const int GP = G | 1;
  float2* sms = reinterpret_cast<float2*>(smem);                 // [64][GP] (mean, sd) of step t at row t & 63
  float* raw = smem + 128 * GP;                                  // [2][32][G] raw mean / variance of the next block
  unsigned short* sdraw = reinterpret_cast<unsigned short*>(raw + 64 * G);   // [slab] picked draw | run shift << 8
  const int n = blockIdx.x;
  const int i_lo = blockIdx.y * slab;
  const int nloc = min(slab, nsamp - i_lo);
  const long B = (long)N * G;
  const float mu = mu_sig ? mu_sig[n] : 0.f, sg = mu_sig ? mu_sig[N + n] : 1.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* pmn = pm + (long)n * G;
  const float* pvn = pv + (long)n * G;
  const uint32_t raw_u32 = smem_u32(raw);
  auto fetch = [&](int t0) {                                     // raw block of the 32 steps starting at t0 (may be empty)
    const int tn = min(32, T - t0);
    for (int tt = warp; tt < tn; tt += ONESTEP_NT / 32) {
      const long o = (long)(t0 + tt) * B;
      for (int j = lane; j < G; j += 32) {
        const uint32_t d = raw_u32 + (uint32_t)(tt * G + j) * 4u;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(pmn + o + j) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + (uint32_t)(32 * G) * 4u), "l"(pvn + o + j) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  fetch(0);
  const long row0 = (long)n * nsamp + i_lo;                      // first output row of the slab
  for (int i = threadIdx.x; i < slab; i += ONESTEP_NT) {
    unsigned v = 0u;
    if (i < nloc) {
      const unsigned shift = (unsigned)(32 - (int)(((row0 + i) * T) & 31)) & 31u;
      v = pick_draw((uint32_t)(row0 + i), 0u, G, seed) | (shift << 8);
    }
    sdraw[i] = (unsigned short)v;
  }
  float* out0 = sims + row0 * T;
  const uint32_t sid0 = (uint32_t)row0;
  const uint32_t sms_u32 = smem_u32(sms);
  const int nfull = nloc & ~3;
  const int niter = (T - 1) / 32 + 2;                            // runs of a sample: a partial head, whole lines, a tail
  for (int m = 0; m < niter; ++m) {
    const int tb = 32 * m;                                       // block of steps staged in this iteration
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                             // raw block landed; the previous iteration is done
    if (tb < T) {
      const int tn = min(32, T - tb);
      for (int tt = warp; tt < tn; tt += ONESTEP_NT / 32) {
        float2* dst = sms + ((tb + tt) & 63) * GP;
        float acc = 0.f;
        for (int j = lane; j < G; j += 32) {
          const float mj = raw[tt * G + j];
          acc += mj;
          dst[j] = make_float2(fmaf(mj, sg, mu), sqrtf(raw[32 * G + tt * G + j]) * sg);
        }
        if (pre_mean && blockIdx.y == 0) {                       // mixture mean over the draws (one_step_dist.mean())
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
          if (lane == 0) pre_mean[(long)n * T + tb + tt] = fmaf(acc / (float)G, sg, mu);
        }
      }
    }
    __syncthreads();
    if (tb + 32 < T) fetch(tb + 32);
    const int tl = tb - 32 + lane;                               // + shift = this lane's step
    // branch-free: the staged row (t & 63) exists for every t, only the store is predicated on 0 <= t < T;
    // 32-bit element index inside the slab (slab * T < 2^32, checked by the host), one wide multiply-add per store
    auto emit = [&](unsigned e16, float z, uint32_t row_idx) {
      const int t = tl + (int)(e16 >> 8);
      float2 v;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];"
                   : "=f"(v.x), "=f"(v.y)
                   : "r"(sms_u32 + (uint32_t)((t & 63) * GP + (int)(e16 & 0xffu)) * 8u));
      const float val = fmaf(v.y, z, v.x);
      if ((unsigned)t < (unsigned)T) out0[row_idx + (uint32_t)t] = val;
    };
    const uint32_t uT = (uint32_t)T;
    uint32_t r0 = (uint32_t)(warp * 4) * uT;
    int q = warp * 4;
    for (; q < nfull; q += ONESTEP_NT / 8) {
      // counter = (first sample of the quad, iteration, lane): four normals, one per sample of the quad
      const f32x4 z = philox_normal4_fast(sid0 + (uint32_t)q, (uint32_t)(tb + lane), 1u, TAG_ONESTEP, seed);
      const uint2 jj = *reinterpret_cast<const uint2*>(sdraw + q);
      emit(jj.x & 0xffffu, z.x, r0);
      emit(jj.x >> 16, z.y, r0 + uT);
      emit(jj.y & 0xffffu, z.z, r0 + 2u * uT);
      emit(jj.y >> 16, z.w, r0 + 3u * uT);
      r0 += (uint32_t)(ONESTEP_NT / 8) * uT;
    }
    if (q < nloc) {                                              // the slab's last, partial quad
      const f32x4 z = philox_normal4_fast(sid0 + (uint32_t)q, (uint32_t)(tb + lane), 1u, TAG_ONESTEP, seed);
      const uint2 jj = *reinterpret_cast<const uint2*>(sdraw + q);
      emit(jj.x & 0xffffu, z.x, r0);
      if (q + 1 < nloc) emit(jj.x >> 16, z.y, r0 + uT);
      if (q + 2 < nloc) emit(jj.y & 0xffffu, z.z, r0 + 2u * uT);
    }
  }
}

// out[n][t] = mean over the G draws of a[t][n*G + j], unstandardised.
__global__ void k_mean_over_draws(int N, int G, int W, const float* __restrict__ a,
                                  const float* __restrict__ mu_sig, float* __restrict__ out) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)N * W) return;
  const int t = (int)(idx % W);
  const int n = (int)(idx / W);
  const long B = (long)N * G;
  float acc = 0.f;
  for (int j = 0; j < G; ++j) acc += a[(long)t * B + (long)n * G + j];
  acc /= (float)G;
  const float mu = mu_sig ? mu_sig[n] : 0.f, sg = mu_sig ? mu_sig[N + n] : 1.f;
  out[idx] = fmaf(acc, sg, mu);
}

// Deterministic forecast mean path per draw: pm[t][b] = H F^t m_b + offset_{T+t}.
template <int S>
__global__ void k_forecast_mean(int N, int G, int T, int Tf, int r, const float* __restrict__ fstate,
                                const float* __restrict__ off, float* __restrict__ pm) {
  const long B = (long)N * G;
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  KState<S> k;
#pragma unroll
  for (int i = 0; i < S; ++i) k.m[i] = fstate[(long)i * B + b];
#pragma unroll
  for (int i = 0; i < KState<S>::NP; ++i) k.P[i] = 0.f;
  int phase = T % r;
  for (int t = 0; t < Tf; ++t) {
    pm[(long)t * B + b] = k.m[0] + (S > 1 ? k.m[1] : 0.f) + off[(long)t * B + b];
    const bool change = (phase == r - 1);
    if (S > 1 && change) {
      float zs = 0.f;
#pragma unroll
      for (int i = 1; i < S; ++i) zs += k.m[i];
#pragma unroll
      for (int i = 1; i < S - 1; ++i) k.m[i] = k.m[i + 1];
      k.m[S - 1] = -zs;
    }
    phase = change ? 0 : phase + 1;
  }
}

// One joint forecast trajectory per thread; 128 samples of one series per CTA, time tiled by 32
// through shared memory so the [N][nsamp][Tf] rows are written with full 128-byte lines.
template <int S>
__global__ void __launch_bounds__(128)
k_forecast_sample(int N, int G, int T, int Tf, int r, int nsamp, uint64_t seed,
                  const float* __restrict__ fstate, const float* __restrict__ scal,
                  const float* __restrict__ off, const float* __restrict__ mu_sig, float* __restrict__ sims) {
  constexpr int NP = S * (S + 1) / 2;
  __shared__ float tile[128][33];
  const int n = blockIdx.y;
  const int i = blockIdx.x * 128 + threadIdx.x;
  const bool live = i < nsamp;
  const long B = (long)N * G;
  const long si = (long)n * nsamp + (live ? i : 0);
  const long b = (long)n * G + pick_draw((uint32_t)si, 1u, G, seed);
  const float so = sqrtf(scal[0 * B + b]), sl = sqrtf(scal[1 * B + b]), sdr = sqrtf(scal[2 * B + b]);
  const float mu = mu_sig ? mu_sig[n] : 0.f, sg = mu_sig ? mu_sig[N + n] : 1.f;
  // x0 = m + L z
  float x[S];
  {
    float z[S];
#pragma unroll
    for (int q = 0; q < (S + 3) / 4; ++q) {
      const f32x4 zz = philox_normal4_fast((uint32_t)si, (uint32_t)q, 0u, TAG_FORECAST_X0, seed);
      if (4 * q + 0 < S) z[4 * q + 0] = zz.x;
      if (4 * q + 1 < S) z[4 * q + 1] = zz.y;
      if (4 * q + 2 < S) z[4 * q + 2] = zz.z;
      if (4 * q + 3 < S) z[4 * q + 3] = zz.w;
    }
#pragma unroll
    for (int a = 0; a < S; ++a) {
      float acc = fstate[(long)a * B + b];
#pragma unroll
      for (int c = 0; c <= a; ++c) acc = fmaf(fstate[(long)(S + NP + a * (a + 1) / 2 + c) * B + b], z[c], acc);
      x[a] = acc;
    }
  }
  int phase = T % r;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t0 = 0; t0 < Tf; t0 += 32) {
    const int tn = min(32, Tf - t0);
    for (int q0 = 0; q0 < tn; q0 += 4) {
      // 12 normals per 4 steps (observation, level and drift noise of each) from THREE generator calls
      const uint32_t grp = (uint32_t)((t0 + q0) >> 2) * 3u;
      const f32x4 na = philox_normal4_fast((uint32_t)si, grp, 0u, TAG_FORECAST, seed);
      const f32x4 nb = philox_normal4_fast((uint32_t)si, grp + 1u, 0u, TAG_FORECAST, seed);
      const f32x4 nc = philox_normal4_fast((uint32_t)si, grp + 2u, 0u, TAG_FORECAST, seed);
      const float nz[12] = {na.x, na.y, na.z, na.w, nb.x, nb.y, nb.z, nb.w, nc.x, nc.y, nc.z, nc.w};
#pragma unroll
      for (int e4 = 0; e4 < 4; ++e4) {
        const int q = q0 + e4;
        if (q < tn) {
          const int t = t0 + q;
          const float yv = x[0] + (S > 1 ? x[1] : 0.f) + off[(long)t * B + b] + so * nz[3 * e4];
          tile[threadIdx.x][q] = fmaf(yv, sg, mu);
          x[0] = fmaf(sl, nz[3 * e4 + 1], x[0]);
          const bool change = (phase == r - 1);
          if (S > 1 && change) {
            float zs = 0.f;
#pragma unroll
            for (int a = 1; a < S; ++a) zs += x[a];
            const float xi = sdr * nz[3 * e4 + 2];
#pragma unroll
            for (int a = 1; a < S - 1; ++a) x[a] = x[a + 1] + xi;
            x[S - 1] = -zs + xi;
          }
          phase = change ? 0 : phase + 1;
        }
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

template <int S>
static void launch_filter(const ci_layout* L, const ci_data* D, const PredWs& ws, float* pm, float* pv, float* fs,
                          cudaStream_t st) {
  const long B = (long)L->N * L->G;
  k_filter<S><<<(unsigned)((B + 63) / 64), 64, 0, st>>>(L->N, L->G, L->T, L->r, D->d_y, D->d_sconst, ws.scal, ws.off,
                                                         pm, pv, fs);
}
static size_t filter_fused_smem(const ci_layout* L) {
  return ((size_t)L->k * 32 + (size_t)filter_series_max(L->G) * (L->k + 1) * CI_TC) * 4 + 48;
}
template <int S>
static int launch_filter_fused(const ci_layout* L, const ci_data* D, const float* d_u, float* pm, float* pv, float* fs,
                               cudaStream_t st) {
  const long B = (long)L->N * L->G;
  const size_t smem = filter_fused_smem(L);
  auto kern = k_filter_fused<S>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return fail(CI_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  kern<<<(unsigned)((B + 31) / 32), 32, smem, st>>>(*L, D->d_X, D->d_sconst, d_u, pm, pv, fs);
  return CI_OK;
}
template <int S>
static void launch_forecast(const ci_layout* L, const FcWs& ws, const float* fs, const float* mu_sig, int nsamp,
                            uint64_t seed, float* sims, cudaStream_t st) {
  const long B = (long)L->N * L->G;
  k_forecast_mean<S><<<(unsigned)((B + 127) / 128), 128, 0, st>>>(L->N, L->G, L->T, L->Tf, L->r, fs, ws.off, ws.pm);
  if (sims)
    k_forecast_sample<S><<<dim3((nsamp + 127) / 128, L->N), 128, 0, st>>>(L->N, L->G, L->T, L->Tf, L->r, nsamp, seed,
                                                                         fs, ws.scal, ws.off, mu_sig, sims);
}

int g_filter_fused = 1;   // ci_set_option("filter_fused", 0 / 1): three kernels / one fused kernel
int g_onestep_tiled = 1;  // ci_set_option("onestep_tiled", 0 / 1): sample-per-thread gather kernel / tiled lane-per-step kernel

}  // namespace ci

using namespace ci;

extern "C" {

size_t ci_predict_workspace_bytes(const ci_layout* L) {
  if (check_layout(L)) return 0;
  const size_t a = pred_ws_floats(L, nullptr, nullptr), b = fc_ws_floats(L, nullptr, nullptr);
  return a > b ? a : b;
}

int ci_filter_predict(const ci_layout* L, const ci_data* D, const float* d_u, float* d_pred_mean,
                      float* d_pred_var, float* d_fstate, void* d_workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_layout(L)) return e;
  CI_CHECK_ARG(D && D->d_y && D->d_sconst && D->d_X, "NULL data pointer");
  CI_CHECK_ARG(d_u && d_pred_mean && d_pred_var && d_fstate && d_workspace, "NULL pointer");
  if (!ci_reg_path(L) && !big_supported(L))
    return fail(CI_ERR_UNSUPPORTED, "nseasons=%d not supported by this build", L->S);
  if (workspace_bytes < pred_ws_floats(L, nullptr, nullptr))
    return fail(CI_ERR_WORKSPACE_TOO_SMALL, "workspace needs %zu bytes", pred_ws_floats(L, nullptr, nullptr));
  cudaStream_t st = (cudaStream_t)stream;
  Carver c(d_workspace);
  PredWs ws;
  pred_ws_floats(L, &ws, &c);
  if (!ci_reg_path(L)) {
    if (int e = launch_filter_big(L, D, d_u, d_pred_mean, d_pred_var, d_fstate, st)) return e;
    CI_CHECK_LAUNCH();
    return CI_OK;
  }
  if (g_filter_fused && filter_fused_smem(L) <= 100 * 1024) {
    int rc = CI_OK;
    CI_DISPATCH_S(L->S, (rc = launch_filter_fused<S_>(L, D, d_u, d_pred_mean, d_pred_var, d_fstate, st)));
    if (rc) return rc;
    CI_CHECK_LAUNCH();
    return CI_OK;
  }
  launch_params_forward(L, D, d_u, ws.beta, ws.scal, st);
  launch_offsets(L, D, ws.beta, ws.off, 0, L->T, 0, st);
  CI_DISPATCH_S(L->S, (launch_filter<S_>(L, D, ws, d_pred_mean, d_pred_var, d_fstate, st)));
  CI_CHECK_LAUNCH();
  return CI_OK;
}

int ci_onestep_sample(const ci_layout* L, const float* d_pred_mean, const float* d_pred_var, const float* d_mu_sig,
                      int32_t nsamp, uint64_t seed, float* d_pre_sims, float* d_pre_mean, void* stream) {
  if (int e = check_layout(L)) return e;
  CI_CHECK_ARG(d_pred_mean && d_pred_var, "NULL pointer");
  CI_CHECK_ARG(nsamp >= 0 && (nsamp == 0 || d_pre_sims), "d_pre_sims is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (nsamp > 0) {
    if (g_onestep_tiled && L->G <= ONESTEP_TILED_MAX_G && (long)L->T * 2112 < (1L << 32)) {   // 32-bit index inside a slab
      // slab of samples per CTA: a multiple of 4, at most 2048, small enough that the grid covers ~4 CTAs per SM
      int nsl = (int)((4L * 148 + L->N - 1) / L->N);
      if (nsl < (nsamp + 2047) / 2048) nsl = (nsamp + 2047) / 2048;
      int slab = ((nsamp + nsl - 1) / nsl + 3) / 4 * 4;
      if (slab < 32) slab = 32;
      nsl = (nsamp + slab - 1) / slab;
      const size_t smem = ((size_t)128 * (L->G | 1) + (size_t)64 * L->G) * 4 + (size_t)slab * 2 + 16;
      static bool configured = false;
      if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_onestep_sample_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
        if (e != cudaSuccess) return fail(CI_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
      }
      k_onestep_sample_tiled<<<dim3(L->N, nsl), ONESTEP_NT, smem, st>>>(L->N, L->G, L->T, nsamp, slab, seed, d_pred_mean,
                                                                 d_pred_var, d_mu_sig, d_pre_sims, d_pre_mean);
      d_pre_mean = nullptr;                    // written by the sampler's staging pass
    } else {
      k_onestep_sample<<<dim3((nsamp + 127) / 128, L->N), 128, 0, st>>>(L->N, L->G, L->T, nsamp, seed, d_pred_mean,
                                                                        d_pred_var, d_mu_sig, d_pre_sims);
    }
  }
  if (d_pre_mean) {
    const long total = (long)L->N * L->T;
    k_mean_over_draws<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(L->N, L->G, L->T, d_pred_mean, d_mu_sig,
                                                                       d_pre_mean);
  }
  CI_CHECK_LAUNCH();
  return CI_OK;
}

int ci_forecast_sample(const ci_layout* L, const ci_data* D, const float* d_u, const float* d_fstate,
                       const float* d_mu_sig, int32_t nsamp, uint64_t seed, float* d_post_sims,
                       float* d_post_mean, void* d_workspace, size_t workspace_bytes, void* stream) {
  if (int e = check_layout(L)) return e;
  CI_CHECK_ARG(D && D->d_y && D->d_sconst && D->d_X, "NULL data pointer");
  CI_CHECK_ARG(d_u && d_fstate && d_workspace, "NULL pointer");
  CI_CHECK_ARG(L->Tf >= 1, "forecast horizon Tf must be >= 1");
  CI_CHECK_ARG(nsamp >= 0 && (nsamp == 0 || d_post_sims), "d_post_sims is NULL");
  if (!ci_reg_path(L) && !big_supported(L))
    return fail(CI_ERR_UNSUPPORTED, "nseasons=%d not supported by this build", L->S);
  if (workspace_bytes < fc_ws_floats(L, nullptr, nullptr))
    return fail(CI_ERR_WORKSPACE_TOO_SMALL, "workspace needs %zu bytes", fc_ws_floats(L, nullptr, nullptr));
  cudaStream_t st = (cudaStream_t)stream;
  Carver c(d_workspace);
  FcWs ws;
  fc_ws_floats(L, &ws, &c);
  launch_params_forward(L, D, d_u, ws.beta, ws.scal, st);
  launch_offsets(L, D, ws.beta, ws.off, L->T, L->T + L->Tf, 0, st);
  if (!ci_reg_path(L)) {
    launch_forecast_big(L, d_fstate, ws.scal, ws.off, ws.pm, d_mu_sig, nsamp, seed, nsamp > 0 ? d_post_sims : nullptr, st);
  } else {
    CI_DISPATCH_S(L->S, (launch_forecast<S_>(L, ws, d_fstate, d_mu_sig, nsamp, seed, nsamp > 0 ? d_post_sims : nullptr, st)));
  }
  if (d_post_mean) {
    const long total = (long)L->N * L->Tf;
    k_mean_over_draws<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(L->N, L->G, L->Tf, ws.pm, d_mu_sig, d_post_mean);
  }
  CI_CHECK_LAUNCH();
  return CI_OK;
}

}  // extern "C"

/* Floats per lane of the forecast-state array ci_filter_predict writes (layout is path-specific:
 * [nfs][B] lane-fastest for the register-resident kernels, lane-contiguous [B][nfs] otherwise). */
extern "C" size_t ci_fstate_floats(const ci_layout* L) {
  if (check_layout(L)) return 0;
  if (ci_reg_path(L)) return (size_t)L->S + (size_t)L->S * (L->S + 1);
  return big_fstate_floats(L);
}
