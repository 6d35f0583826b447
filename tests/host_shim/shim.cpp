// TEST-ONLY host shim: runs the per-lane __host__ __device__ arithmetic of the CUDA kernels on the
// CPU so `pytest -m "not gpu"` can check the hand-derived adjoint, the parameter VJP and the Philox
// stream against the oracle without a GPU.  Never linked into, loaded by, or shipped with the
// product library (libci_b200.so); the product path has no CPU fallback.
#include <cstring>
#include <vector>
#include "ci_core.cuh"
#include "ci_eval_lane.cuh"
#include "ci_dispatch.cuh"

using namespace ci;

template <int S>
static void lane_eval(const float* u, long B, const float* sc, int N, const float* XYn, int T, int k, int r,
                      float* sb, float* sr, float* tape, float* logp, float* grad) {
  XDirect xs{XYn, (long)N * (k + 1) * CI_TC};
  // same dispatch rule as launch_eval (ci_eval.cu)
  if (r == 1 && S >= 2 && S <= CI_MAX_EFF_S) {
    constexpr int ES = (S >= 2 && S <= CI_MAX_EFF_S) ? S : 2;
    eval_lane_eff<ES, 1, 1>(u, B, sc, N, xs, T, k, sb, sr, tape, ES + 1, logp, grad);
  } else {
    eval_lane<S>(u, B, sc, N, xs, T, k, r, sb, 1, sr, tape, 1, logp, grad);
  }
}

// Runs ci::eval_lane -- the exact per-lane function the CUDA kernel k_eval executes -- for every lane.
// XY: chunk-major [Tpad/4][N][k+1][4] (row k = y).
extern "C" int shim_logp_grad(int N, int G, int T, int k, int S, int r, const float* XY, int Tpad,
                              const float* sconst, const float* u, float* logp, float* grad) {
  const long B = (long)N * G;
  std::vector<float> sb((size_t)(k > 0 ? k : 1)), sr(4), tape((size_t)T * (S + 1));
  for (long b = 0; b < B; ++b) {
    const int n = (int)(b / G);
    CI_DISPATCH_S(S, (lane_eval<S_>(u + b, B, sconst + n, N, XY + (long)n * (k + 1) * CI_TC, T, k, r, sb.data(),
                                    sr.data(), tape.data(), logp + b, grad + b)));
  }
  return 0;
}

extern "C" void shim_philox4x32(int n, const uint32_t* c0, const uint32_t* c1, const uint32_t* c2,
                                const uint32_t* c3, uint64_t seed, uint32_t* out) {
  for (int i = 0; i < n; ++i) {
    u32x4 r = philox4x32(c0[i], c1[i], c2[i], c3[i], seed);
    out[4 * i + 0] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
  }
}

extern "C" void shim_philox_normal4(int n, const uint32_t* c0, const uint32_t* c1, const uint32_t* c2,
                                    uint32_t tag, uint64_t seed, float* out) {
  for (int i = 0; i < n; ++i) {
    f32x4 r = philox_normal4(c0[i], c1[i], c2[i], tag, seed);
    out[4 * i + 0] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
  }
}
