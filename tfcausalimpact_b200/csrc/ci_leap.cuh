// Leapfrog kick / drift fused behind an evaluation (shared by the thread-per-lane kernels, ci_eval.cu / ci_eval_ws.cu):
// with the fresh gradient,  p += (last ? 1/2 : 1) eps g ;  if (!last) u += eps p ,  eps = step0 * exp(logfac[lane]).
#pragma once
#include "ci_abi_common.cuh"

namespace ci {

__device__ __forceinline__ void leap_epilogue(const LeapArgs& lf, int P, long B, long b,
                                              const float* __restrict__ grad) {
  if (lf.pp == nullptr) return;
  float* __restrict__ pp = lf.pp + b;
  float* __restrict__ uu = lf.uu + b;
  const float* __restrict__ st = lf.step0 + b;
  const float* __restrict__ gr = grad + b;
  const float fac = expf(lf.logfac[b]);
  const float c = lf.last ? 0.5f : 1.0f;
  constexpr int U = 8;   // 8 parameters per round: 32 independent loads in flight
  for (int i0 = 0; i0 < P; i0 += U) {
    float e[U], g[U], p[U], x[U];
#pragma unroll
    for (int q = 0; q < U; ++q) {
      const long o = (long)(i0 + q < P ? i0 + q : P - 1) * B;
      e[q] = st[o]; g[q] = gr[o]; p[q] = pp[o]; x[q] = uu[o];
    }
#pragma unroll
    for (int q = 0; q < U; ++q) {
      if (i0 + q < P) {
        const long o = (long)(i0 + q) * B;
        const float eps = e[q] * fac;
        const float ph = fmaf(c * eps, g[q], p[q]);
        pp[o] = ph;
        if (!lf.last) uu[o] = fmaf(eps, ph, x[q]);
      }
    }
  }
}

}  // namespace ci
