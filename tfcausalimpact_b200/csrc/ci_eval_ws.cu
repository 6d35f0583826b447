// K1-K3 for MID-SIZE batches (above the lane-cooperative kernel's one-wave limit, below a full machine: ~13k .. 28k
// lanes -- the named configs[3] split at 4 GPUs): thread-per-lane evaluation with the regression taken OFF the
// dependent Kalman chain by warp specialisation.
//
// In the thread-per-lane kernel (ci_eval.cu) a lane alternates  residuals of a 4-step chunk (50 x shared-memory load +
// FFMA2)  ->  4 filter steps  ->  next chunk: with one or two warps per scheduler nothing hides the regression, and it
// is HALF of a lane's latency (one warp alone: 0.223 ms at k = 0, 0.446 ms at k = 50).  Here a CTA is TWO warps for the
// same 32 lanes:
//   * the MAIN warp runs only the filter / adjoint steps (EffLane::fwd_step / bwd_step of ci_eval_lane.cuh, state in
//     registers, tape as in ci_eval.cu) and the parameter transforms;
//   * the HELPER warp runs the regression of the same lanes one chunk ahead (forward: residuals; reverse: beta_bar
//     accumulation), from TMA-staged design chunks it double-buffers itself.
// They meet through a two-slot shared-memory exchange ([2][32][4] residuals forward, rbar backward) behind mbarriers
// (full / empty per slot): no CTA-wide barrier inside the sweeps, the helper is never more than two chunks ahead.
// With 13k-28k lanes the SMs have idle schedulers, so the helper warps cost nothing the main warps could have used.
// The helper also stages the TAPE for the reverse sweep: a step's entry of the warp is one contiguous (S + 1) x 128-byte
// run, which it bulk-copies (TMA) into a two-chunk shared-memory ring while the main warp works on the previous chunk,
// so the reverse step starts from shared memory instead of an L2 round trip (885 -> ~600 cycles per reverse step).
// Measured (scripts/ab_ws.py, k = 50, 8 chains, ms per evaluation, single-warp kernel -> this one; profiles/
// r2_ws_ab_k50.jsonl): one CTA alone 0.446 -> 0.263; 15,000 lanes 0.509 -> 0.296; 20,000 0.521 -> 0.352; 28,000
// 0.522 -> 0.355 (without the tape staging: 0.379 / 0.435 / 0.440).  Beyond one wave at 6 CTAs per SM (28,416 lanes) a
// 128-register build gave 32,000: 0.522 -> 0.504, 35,200: 0.526 -> 0.566, 40,000: 0.607 -> 0.717 and a 96-register build
// 40,000: 0.607 -> 0.663 (helper warps then compete with main warps for issue slots): the single-warp kernel stays there.
// Same math and tests as the other evaluation kernels (tests/test_ws_gpu.py).
#include "ci_abi_common.cuh"
#include "ci_eval_lane.cuh"
#include "ci_xstage.cuh"
#include "ci_kalman_eff.cuh"
#include "ci_leap.cuh"

namespace ci {

struct XNone {     // EffLane's own regression members are never instantiated here
  using tape_in_shared = void;   // the reverse steps read tape entries the helper staged in shared memory
};

__host__ __device__ inline int ws_series_max(int G) {   // series a warp's 32 lanes can touch
  const int a = (32 + G - 2) / G + 1;
  return a < 32 ? a : 32;
}

__device__ __forceinline__ void ws_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void ws_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ws_tma(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads of the slot precede the async write
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// barrier slots (8 bytes each): forward xfull[2] rfull[2] rempty[2], reverse xfull[2] rfull[2] rempty[2], then
// tape-ready (forward tape complete) and tape-full[2] (a chunk's tape entries staged for the reverse steps)
enum { WS_XF = 0, WS_RF = 2, WS_RE = 4, WS_REV = 6, WS_TR = 12, WS_TF = 13, WS_NBAR = 15 };

// MINB = resident CTAs per SM the register budget must allow: 6 -> up to 170 registers (one wave up to 28,416 lanes); the
// main warp's filter steps take 148.  Builds for 8 / 9 CTAs per SM (128 / 96 registers) were measured and lost to the
// single-warp kernel above 28k lanes (header).
template <int S, int MINB>
__global__ void __launch_bounds__(64, MINB)
k_eval_ws(ci_layout L, const float* __restrict__ XY, const float* __restrict__ sconst, const float* u,
          float* __restrict__ logp, float* __restrict__ grad, float* __restrict__ tape, LeapArgs lf) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = L.N, k = L.k, T = L.T;
  const long B = (long)N * L.G;
  const long bw0 = (long)blockIdx.x * 32;
  const bool valid = bw0 + lane < B;
  const long b = valid ? bw0 + lane : B - 1;            // padding lanes shadow the last lane (outputs suppressed)
  const int n = (int)(b / L.G);
  const int stride = (k + 1) * CI_TC;
  const int nchunk = (T + CI_TC - 1) / CI_TC;
  const int smax = ws_series_max(L.G);
  float* sb = smem + lane;                                // [k][32]  -beta, then beta_bar
  float* sr = smem + k * 32;                              // [2][32][4] exchange slots
  float* sx = sr + 2 * 32 * 4;                            // [2][smax][stride] staged design chunks
  const int stage_floats = smax * stride;
  constexpr int TE = (S + 1) * 32;                        // floats of one step's tape entry of the warp
  float* sring = sx + 2 * stage_floats;                   // [2][CI_TC][TE] tape entries staged for the reverse steps
  const uint32_t bar0 = smem_u32(sring + 2 * CI_TC * TE);
  const int n_lo = (int)(bw0 / L.G);
  const long b_last = (bw0 + 31 < B ? bw0 + 31 : B - 1);
  const int ns = (int)(b_last / L.G) - n_lo + 1;
  const float* xsrc = XY + (long)n_lo * stride;           // chunk 0 of the warp's series run
  const long xcs = (long)N * stride;                      // floats between consecutive chunks
  const uint32_t xbytes = (uint32_t)ns * (uint32_t)stride * 4u;
  const uint32_t sx_u32 = smem_u32(sx);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < WS_NBAR; ++q) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * q) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1 && lane == 0) {                           // the first two chunks are on their way during the transforms
    ws_tma(sx_u32, xsrc, xbytes, bar0 + 8u * (WS_XF + 0));
    if (nchunk > 1) ws_tma(sx_u32 + (uint32_t)stage_floats * 4u, xsrc + xcs, xbytes, bar0 + 8u * (WS_XF + 1));
  }
  constexpr int ES = S;
  const long nwarp = (B + 31) / 32;
  float lp0 = 0.f, ll = 0.f, Rb = 0.f, qlb = 0.f, qdb = 0.f;
  XNone xn;
  using LaneT = EffLane<ES, XNone, 32, 32>;
  if (warp == 0) {
    // ================================ MAIN: parameters, filter, adjoint =========================================
    const LaneScalars ls = params_forward_lane(u + b, B, k, L.S, sconst + n, N, sb, 32, -1.0f);   // sb <- -beta
    lp0 = ls.lp0;
    __syncthreads();                                      // (A) -beta column visible to the helper
    float* tape_eff = tape + (bw0 >> 5) * ((ES + 1) * 32) + lane;
    const int tstep = (int)(nwarp * ((ES + 1) * 32));
    LaneT Ln(xn, T, k, sb, sr, tape_eff, tstep, ls.R, ls.ql, ls.qd, lane < ES + 1 ? lane : -1);
    ef_init<ES>(Ln.st, sconst[SC_M0 * N + n], sconst[SC_P0 * N + n]);
    int c = 0;                                            // season index
    for (int tc = 0; tc < nchunk; ++tc) {
      const int slot = tc & 1;
      ws_wait(bar0 + 8u * (WS_RF + slot), (uint32_t)(tc >> 1) & 1u);
      const float4 r4 = *reinterpret_cast<const float4*>(sr + (slot * 32 + lane) * 4);
      __syncwarp();
      if (lane == 0) ws_arrive(bar0 + 8u * (WS_RE + slot));           // the slot is free again
      const float rr[4] = {r4.x, r4.y, r4.z, r4.w};
      const int nstep = (T - tc * CI_TC) < CI_TC ? (T - tc * CI_TC) : CI_TC;
#pragma unroll
      for (int i = 0; i < CI_TC; ++i) {
        if (i < nstep) {
          Ln.template disp_fwd<0>(c, rr[i]);
          c = (c + 1 == ES) ? 0 : c + 1;
        }
      }
    }
    ll = -0.5f * (fmaf(Ln.acc, 0.6931471805599453f, Ln.quad) + (float)(T - Ln.nmask) * CI_LOG_2PI);
#pragma unroll
    for (int i = 0; i < ES + 1; ++i) Ln.st.m[i] = 0.f;
#pragma unroll
    for (int i = 0; i < EState<ES>::NP; ++i) Ln.st.P[i] = 0.f;
    // the forward tape (generic-proxy stores of all 32 lanes) must be visible to the helper's bulk (async-proxy) loads
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncwarp();
    if (lane == 0) ws_arrive(bar0 + 8u * WS_TR);
    c = (T - 1) % ES;
    for (int tc = nchunk - 1, it = 0; tc >= 0; --tc, ++it) {
      const int slot = it & 1;
      float rb[4] = {0.f, 0.f, 0.f, 0.f};
      const int nstep = (T - tc * CI_TC) < CI_TC ? (T - tc * CI_TC) : CI_TC;
      ws_wait(bar0 + 8u * (WS_TF + slot), (uint32_t)(it >> 1) & 1u);   // the chunk's tape entries are in shared memory
#pragma unroll
      for (int i = CI_TC - 1; i >= 0; --i) {
        if (i < nstep) {
          Ln.tp = sring + (slot * CI_TC + i) * TE + lane;
          rb[i] = Ln.template disp_bwd<0>(c, false);
          c = (c == 0) ? ES - 1 : c - 1;
        }
      }
      if (it >= 2) ws_wait(bar0 + 8u * (WS_REV + WS_RE + slot), (uint32_t)((it >> 1) - 1) & 1u);
      *reinterpret_cast<float4*>(sr + (slot * 32 + lane) * 4) = make_float4(rb[0], rb[1], rb[2], rb[3]);
      __syncwarp();
      if (lane == 0) ws_arrive(bar0 + 8u * (WS_REV + WS_RF + slot));
    }
    Rb = Ln.Rb; qlb = Ln.qlb; qdb = Ln.qdb;
  } else {
    // ================================ HELPER: regression one chunk ahead =========================================
    __syncthreads();                                      // (A)
    const float* lane_x = sx + (long)(n - n_lo) * stride;
    for (int tc = 0; tc < nchunk; ++tc) {
      const int slot = tc & 1;
      if (tc >= 2) ws_wait(bar0 + 8u * (WS_RE + slot), (uint32_t)((tc >> 1) - 1) & 1u);
      ws_wait(bar0 + 8u * (WS_XF + slot), (uint32_t)(tc >> 1) & 1u);
      const float* __restrict__ xp = lane_x + slot * stage_floats;
      const float4 yv = *reinterpret_cast<const float4*>(xp + k * CI_TC);          // row k: centred series
      pf2 r01 = mk2(yv.x, yv.y), r23 = mk2(yv.z, yv.w), s01 = mk2(0.f, 0.f), s23 = mk2(0.f, 0.f);
      int j = 0;
      for (; j + 4 <= k; j += 4) {
        float nb[4];
        float4 x[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          nb[q] = sb[(j + q) * 32];
          x[q] = *reinterpret_cast<const float4*>(xp + (j + q) * CI_TC);
        }
        r01 = ffma2(mk2(nb[0], nb[0]), mk2(x[0].x, x[0].y), r01);
        r23 = ffma2(mk2(nb[0], nb[0]), mk2(x[0].z, x[0].w), r23);
        s01 = ffma2(mk2(nb[1], nb[1]), mk2(x[1].x, x[1].y), s01);
        s23 = ffma2(mk2(nb[1], nb[1]), mk2(x[1].z, x[1].w), s23);
        r01 = ffma2(mk2(nb[2], nb[2]), mk2(x[2].x, x[2].y), r01);
        r23 = ffma2(mk2(nb[2], nb[2]), mk2(x[2].z, x[2].w), r23);
        s01 = ffma2(mk2(nb[3], nb[3]), mk2(x[3].x, x[3].y), s01);
        s23 = ffma2(mk2(nb[3], nb[3]), mk2(x[3].z, x[3].w), s23);
      }
      for (; j < k; ++j) {
        const float nb = sb[j * 32];
        const float4 x = *reinterpret_cast<const float4*>(xp + j * CI_TC);
        r01 = ffma2(mk2(nb, nb), mk2(x.x, x.y), r01);
        r23 = ffma2(mk2(nb, nb), mk2(x.z, x.w), r23);
      }
      r01 = fadd2(r01, s01);
      r23 = fadd2(r23, s23);
      *reinterpret_cast<float4*>(sr + (slot * 32 + lane) * 4) = make_float4(r01.x, r01.y, r23.x, r23.y);
      __syncwarp();                                       // every lane has consumed the staged chunk and written its slot
      if (lane == 0) {
        if (tc + 2 < nchunk)
          ws_tma(sx_u32 + (uint32_t)(slot * stage_floats) * 4u, xsrc + (long)(tc + 2) * xcs, xbytes, bar0 + 8u * (WS_XF + slot));
        ws_arrive(bar0 + 8u * (WS_RF + slot));
      }
    }
    // reverse: beta_bar_j = -sum_t rbar_t X_tj, chunks in descending order; the accumulators replace -beta in sb
    for (int j = 0; j < k; ++j) sb[j * 32] = 0.f;
    __syncwarp();
    if (lane == 0) {
      ws_tma(sx_u32, xsrc + (long)(nchunk - 1) * xcs, xbytes, bar0 + 8u * (WS_REV + WS_XF + 0));
      if (nchunk > 1)
        ws_tma(sx_u32 + (uint32_t)stage_floats * 4u, xsrc + (long)(nchunk - 2) * xcs, xbytes, bar0 + 8u * (WS_REV + WS_XF + 1));
    }
    // tape entries of a chunk (one contiguous (S + 1) x 128-byte run per step in the warp-blocked layout) -> ring slot
    const float* wtape = tape + (bw0 >> 5) * TE;
    const long tstep_g = (long)((B + 31) / 32) * TE;
    const uint32_t ring_u32 = smem_u32(sring);
    auto tape_issue = [&](int tc, int slot) {                                     // lane 0
      const int nstep = (T - tc * CI_TC) < CI_TC ? (T - tc * CI_TC) : CI_TC;
      const uint32_t bar = bar0 + 8u * (WS_TF + slot);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(nstep * TE) * 4u) : "memory");
      for (int i = 0; i < nstep; ++i)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         ring_u32 + (uint32_t)((slot * CI_TC + i) * TE) * 4u),
                     "l"(wtape + (long)(tc * CI_TC + i) * tstep_g), "r"((uint32_t)TE * 4u), "r"(bar)
                     : "memory");
    };
    if (lane == 0) {
      ws_wait(bar0 + 8u * WS_TR, 0u);                      // the main warp has finished (and fenced) the forward tape
      asm volatile("fence.proxy.async;" ::: "memory");
      tape_issue(nchunk - 1, 0);
      if (nchunk > 1) tape_issue(nchunk - 2, 1);
    }
    for (int tc = nchunk - 1, it = 0; tc >= 0; --tc, ++it) {
      const int slot = it & 1;
      ws_wait(bar0 + 8u * (WS_REV + WS_RF + slot), (uint32_t)(it >> 1) & 1u);
      if (lane == 0 && tc - 2 >= 0) {                      // the main warp is done with the slot's tape entries too
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tape_issue(tc - 2, slot);
      }
      ws_wait(bar0 + 8u * (WS_REV + WS_XF + slot), (uint32_t)(it >> 1) & 1u);
      const float4 rb = *reinterpret_cast<const float4*>(sr + (slot * 32 + lane) * 4);
      const pf2 n01 = mk2(-rb.x, -rb.y), n23 = mk2(-rb.z, -rb.w);
      const float* __restrict__ xp = lane_x + slot * stage_floats;
      int j = 0;
      for (; j + 4 <= k; j += 4) {
        float4 x[4];
        pf2 a[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          x[q] = *reinterpret_cast<const float4*>(xp + (j + q) * CI_TC);
          a[q] = mk2(sb[(j + q) * 32], 0.f);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          a[q] = ffma2(mk2(x[q].x, x[q].y), n01, a[q]);
          a[q] = ffma2(mk2(x[q].z, x[q].w), n23, a[q]);
          sb[(j + q) * 32] = a[q].x + a[q].y;
        }
      }
      for (; j < k; ++j) {
        const float4 x = *reinterpret_cast<const float4*>(xp + j * CI_TC);
        pf2 a = mk2(sb[j * 32], 0.f);
        a = ffma2(mk2(x.x, x.y), n01, a);
        a = ffma2(mk2(x.z, x.w), n23, a);
        sb[j * 32] = a.x + a.y;
      }
      __syncwarp();
      if (lane == 0) {
        if (tc - 2 >= 0)
          ws_tma(sx_u32 + (uint32_t)(slot * stage_floats) * 4u, xsrc + (long)(tc - 2) * xcs, xbytes,
                 bar0 + 8u * (WS_REV + WS_XF + slot));
        ws_arrive(bar0 + 8u * (WS_REV + WS_RE + slot));
      }
    }
  }
  __syncthreads();                                        // (B) beta_bar complete
  if (warp == 0) {
    // VJP through the parameter transforms + priors; padding lanes write into a scratch gradient nobody reads
    if (valid) {
      params_backward_lane(u + b, B, k, L.S, sconst + n, N, sb, 32, Rb, qlb, qdb, grad + b);
      logp[b] = ll + lp0;
      leap_epilogue(lf, num_params(k, L.S), B, b, grad);
    }
  }
}

// ---- host side ------------------------------------------------------------------------------------------------
int g_eval_ws = 1;                       // ci_set_option("eval_ws", 0 / 1)
#define CI_WS_LANES_6 (148L * 6 * 32)    // one wave at 6 CTAs (= 6 main warps) per SM: 28,416 lanes, 170 registers
#define CI_WS_MAX_LANES CI_WS_LANES_6     // beyond one wave at 6 CTAs per SM the 128-register build wins little or loses (below)
long g_ws_max_lanes = -1;                // ci_set_option("eval_ws_max_lanes"): -1 = default

static size_t ws_smem_bytes(const ci_layout* L) {
  return ((size_t)L->k * 32 + 2 * 32 * 4 + 2 * (size_t)ws_series_max(L->G) * (L->k + 1) * CI_TC +
          2 * CI_TC * (size_t)(L->S + 1) * 32) * 4 + WS_NBAR * 8 + 16;
}

bool ws_eligible(const ci_layout* L) {
  if (!g_eval_ws || !ci_reg_path(L) || L->flags != 0 || L->S2 > 1) return false;
  if (L->r != 1 || L->S < 2 || L->S > CI_MAX_EFF_S || L->G < 4 || L->k < 4) return false;
  const long B = (long)L->N * L->G;
  if (B > (g_ws_max_lanes >= 0 ? g_ws_max_lanes : CI_WS_MAX_LANES)) return false;
  return ws_smem_bytes(L) * 6 <= (size_t)220 * 1024;
}

template <int S, int MINB>
static int launch_ws_t(const ci_layout* L, const ci_data* D, const float* d_u, float* d_logp, float* d_grad, float* tape,
                       cudaStream_t st, const LeapArgs& lf) {
  const long B = (long)L->N * L->G;
  const size_t smem = ws_smem_bytes(L);
  auto kern = k_eval_ws<S, MINB>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return fail(CI_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  kern<<<(unsigned)((B + 31) / 32), 64, smem, st>>>(*L, D->d_X, D->d_sconst, d_u, d_logp, d_grad, tape, lf);
  return CI_OK;
}

int launch_eval_ws(const ci_layout* L, const ci_data* D, const float* d_u, float* d_logp, float* d_grad, float* tape,
                   cudaStream_t st, const LeapArgs& lf) {
  int rc = CI_OK;
#define CI_WS_CASE(S_) case S_: rc = launch_ws_t<S_, 6>(L, D, d_u, d_logp, d_grad, tape, st, lf); break;
  switch (L->S) {
    CI_WS_CASE(2) CI_WS_CASE(3) CI_WS_CASE(4) CI_WS_CASE(5) CI_WS_CASE(6) CI_WS_CASE(7)
    default: return fail(CI_ERR_UNSUPPORTED, "warp-specialised evaluation needs 2 <= nseasons <= 7");
  }
#undef CI_WS_CASE
  if (rc) return rc;
  CI_CHECK_LAUNCH();
  return CI_OK;
}

}  // namespace ci
