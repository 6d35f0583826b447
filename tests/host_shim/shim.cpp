// TEST-ONLY host shim: runs the per-lane __host__ __device__ arithmetic of the CUDA kernels on the
// CPU so `pytest -m "not gpu"` can check the hand-derived adjoint, the parameter VJP and the Philox
// stream against the oracle without a GPU.  Never linked into, loaded by, or shipped with the
// product library (libci_b200.so); the product path has no CPU fallback.
#include <cstring>
#include <vector>
#include "ci_core.cuh"
#include "ci_kalman.cuh"
#include "ci_params.cuh"
#include "ci_dispatch.cuh"

using namespace ci;

template <int S>
static void lane_eval(float* resid, float* tape, long ld, int T, int r, LaneScalars ls, float m0, float p0,
                      float& ll, float& Rb, float& qlb, float& qdb) {
  kalman_logp_grad_lane<S>(resid, tape, ld, T, r, ls.R, ls.ql, ls.qd, m0, p0, ll, Rb, qlb, qdb);
}

extern "C" int shim_logp_grad(int N, int G, int T, int k, int S, int r, const float* y, const float* X,
                              int Tpad, const float* sconst, const float* u, float* logp, float* grad) {
  const long B = (long)N * G;
  const int P = num_params(k, S);
  (void)P;
  std::vector<float> beta((size_t)(k > 0 ? k : 1) * B), resid((size_t)T * B), tape((size_t)T * (S + 2) * B);
  for (long b = 0; b < B; ++b) {
    const int n = (int)(b / G);
    const float* sc = sconst + n;
    LaneScalars ls = params_forward_lane(u + b, B, k, S, sc, N, beta.data() + b);
    for (int t = 0; t < T; ++t) {
      float acc = y[(long)n * T + t] - sc[SC_MEAN * N];
      for (int j = 0; j < k; ++j) acc -= beta[(size_t)j * B + b] * X[((long)n * k + j) * Tpad + t];
      resid[(size_t)t * B + b] = acc;
    }
    float ll = 0, Rb = 0, qlb = 0, qdb = 0;
    const float m0 = sc[SC_M0 * N], p0 = sc[SC_P0 * N];
    CI_DISPATCH_S(S, (lane_eval<S_>(resid.data() + b, tape.data() + b, B, T, r, ls, m0, p0, ll, Rb, qlb, qdb)));
    for (int j = 0; j < k; ++j) {
      float acc = 0.f;
      for (int t = 0; t < T; ++t) acc -= resid[(size_t)t * B + b] * X[((long)n * k + j) * Tpad + t];
      beta[(size_t)j * B + b] = acc;
    }
    params_backward_lane(u + b, B, k, S, sc, N, beta.data() + b, Rb, qlb, qdb, grad + b);
    logp[b] = ll + ls.lp0;
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
