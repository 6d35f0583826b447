// Shared scalar helpers for the sm_100a kernels: packed-symmetric indexing, stable softplus,
// Philox4x32-10 counter RNG.  Everything here is `__host__ __device__` so the test-only host shim
// (tests/host_shim) can run the exact per-lane arithmetic on the CPU against the oracle.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CI_HD __host__ __device__ __forceinline__
#else
#define CI_HD inline
#endif

#define CI_LOG_2PI 1.8378770664093453f

namespace ci {

// reciprocal: one MUFU.RCP (rcp.approx, <= 1 ulp) on device, IEEE divide on host.  Arguments on this
// path are innovation variances / softplus values, far from the denormal and overflow ranges.
CI_HD float rcp(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / x;
#endif
}

// log of an innovation variance for the log-likelihood sum: MUFU.LG2 based on device (absolute error
// ~4e-7 per term, i.e. ~1e-7 relative on the summed log-likelihood), libm on host.
CI_HD float fast_log(float x) {
#if defined(__CUDA_ARCH__)
  return __logf(x);
#else
  return logf(x);
#endif
}

// log2 of an innovation variance (never denormal: s >= sigma_obs^2 > 0): a single MUFU.LG2
CI_HD float fast_log2(float x) {
#if defined(__CUDA_ARCH__) && defined(CI_LOG2_BUILTIN)
  return __log2f(x);
#elif defined(__CUDA_ARCH__)
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return log2f(x);
#endif
}

CI_HD float softplus(float u) {
  // log(1 + e^u), stable on both tails
  return (u > 0.f ? u : 0.f) + log1pf(expf(-fabsf(u)));
}
CI_HD float sigmoid(float u) {
  float e = expf(-fabsf(u));
  float s = 1.0f / (1.0f + e);
  return u >= 0.f ? s : e * s;
}
CI_HD float log_sigmoid(float u) {
  return (u < 0.f ? u : 0.f) - log1pf(expf(-fabsf(u)));
}

// softplus(u), log_sigmoid(u), sigmoid(u), sigmoid(-u) from ONE exp / log1p / reciprocal.
struct SoftplusPack {
  float sp, lsig, sig, nsig;
};
CI_HD SoftplusPack softplus_pack(float u) {
  const float e = expf(-fabsf(u));
  const float l = log1pf(e);
  const float s = rcp(1.0f + e);
  SoftplusPack r;
  r.sp = (u > 0.f ? u : 0.f) + l;
  r.lsig = (u < 0.f ? u : 0.f) - l;
  r.sig = u >= 0.f ? s : e * s;
  r.nsig = u >= 0.f ? e * s : s;
  return r;
}

// Packed pair of floats: on sm_100a the pair arithmetic below compiles to the FP32x2 instructions
// FFMA2 / FADD2 / FMUL2 (one issue slot for two IEEE-rounded operations, bit-identical to the scalar
// forms); the host build (test shim) runs the same arithmetic as two scalar operations.
#if defined(__CUDACC__)
typedef float2 pf2;
#else
struct pf2 {
  float x, y;
};
#endif
CI_HD pf2 mk2(float a, float b) {
  pf2 r;
  r.x = a;
  r.y = b;
  return r;
}
CI_HD pf2 ffma2(pf2 a, pf2 b, pf2 c) {
#if defined(__CUDA_ARCH__)
  return __ffma2_rn(a, b, c);
#else
  return mk2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
CI_HD pf2 fadd2(pf2 a, pf2 b) {
#if defined(__CUDA_ARCH__)
  return __fadd2_rn(a, b);
#else
  return mk2(a.x + b.x, a.y + b.y);
#endif
}
CI_HD pf2 fmul2(pf2 a, pf2 b) {
#if defined(__CUDA_ARCH__)
  return __fmul2_rn(a, b);
#else
  return mk2(a.x * b.x, a.y * b.y);
#endif
}

// Packed upper-triangular index of a symmetric SxS matrix (i <= j).
template <int S>
CI_HD constexpr int tri(int i, int j) {
  return i * S - (i * (i - 1)) / 2 + (j - i);
}
template <int S>
CI_HD constexpr int sym(int i, int j) {
  return i <= j ? tri<S>(i, j) : tri<S>(j, i);
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  Counter words: (c0, c1, c2, tag); key = 64-bit seed.
// The oracle (oracle/sts_oracle.py::philox4x32) implements the same function bit for bit.
// ---------------------------------------------------------------------------------------------
struct u32x4 {
  uint32_t x, y, z, w;
};

CI_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
  hi = __umulhi(a, b);
  lo = a * b;
#else
  uint64_t p = (uint64_t)a * (uint64_t)b;
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
#endif
}

template <int ROUNDS>
CI_HD u32x4 philox4x32_r(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < ROUNDS; ++i) {
    uint32_t hi0, lo0, hi1, lo1;
    mulhilo(0xD2511F53u, c0, hi0, lo0);
    mulhilo(0xCD9E8D57u, c2, hi1, lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return u32x4{c0, c1, c2, c3};
}
CI_HD u32x4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed) {
  return philox4x32_r<10>(c0, c1, c2, c3, seed);
}

// (0,1) uniform with 24 random bits: ((x >> 8) + 0.5) * 2^-24 -- exact in fp32.
CI_HD float u32_to_unit(uint32_t x) { return ((float)(x >> 8) + 0.5f) * 5.9604644775390625e-08f; }

struct f32x4 {
  float x, y, z, w;
};

CI_HD f32x4 philox_uniform4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t tag, uint64_t seed) {
  u32x4 r = philox4x32(c0, c1, c2, tag, seed);
  return f32x4{u32_to_unit(r.x), u32_to_unit(r.y), u32_to_unit(r.z), u32_to_unit(r.w)};
}

// Four standard normals (two Box-Muller pairs).
CI_HD f32x4 philox_normal4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t tag, uint64_t seed) {
  f32x4 u = philox_uniform4(c0, c1, c2, tag, seed);
  float r0 = sqrtf(-2.0f * logf(u.x));
  float r1 = sqrtf(-2.0f * logf(u.z));
  float s0, c0f, s1, c1f;
  const float two_pi = 6.283185307179586f;
#if defined(__CUDA_ARCH__)
  sincosf(two_pi * u.y, &s0, &c0f);
  sincosf(two_pi * u.w, &s1, &c1f);
#else
  s0 = sinf(two_pi * u.y); c0f = cosf(two_pi * u.y);
  s1 = sinf(two_pi * u.w); c1f = cosf(two_pi * u.w);
#endif
  return f32x4{r0 * c0f, r0 * s0, r1 * c1f, r1 * s1};
}

// Philox4x32-7 (the smallest round count that passes BigCrush, Salmon et al. 2011, table 2) + Box-Muller on the
// special-function unit (MUFU lg2 / sin / cos / sqrt, angle in [-pi, pi)): for the posterior-predictive SAMPLERS
// only, whose contract is distributional (absolute error of a draw ~1e-6) and whose cost is dominated by the
// generator (a 10-round call is ~100 of the ~170 instructions of a forecast step); the fit drivers keep the
// 10-round philox_normal4, which the oracle reproduces bit for bit.
CI_HD f32x4 philox_normal4_fast(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t tag, uint64_t seed) {
#if defined(__CUDA_ARCH__)
  const u32x4 rr = philox4x32_r<7>(c0, c1, c2, tag, seed);
  // radius: u = ((x >> 8) + 0.5) 2^-24 in (0, 1), never denormal -> lg2.approx.ftz; sqrt(-2 ln u) = sqrt(-2 ln2 lg2 u)
  // angle:  2 pi (u - 0.5) in [-pi, pi) straight from the integer: one conversion + one FMA
  const float ux = fmaf((float)(rr.x >> 8), 5.9604644775390625e-08f, 2.98023223876953125e-08f);
  const float uz = fmaf((float)(rr.z >> 8), 5.9604644775390625e-08f, 2.98023223876953125e-08f);
  const float a0 = fmaf((float)(rr.y >> 8), 3.7450702829239664e-07f, -3.1415924662916617f);
  const float a1 = fmaf((float)(rr.w >> 8), 3.7450702829239664e-07f, -3.1415924662916617f);
  float l0, l1, r0, r1;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l0) : "f"(ux));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l1) : "f"(uz));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(-1.3862943611198906f * l0));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(-1.3862943611198906f * l1));
  return f32x4{r0 * __cosf(a0), r0 * __sinf(a0), r1 * __cosf(a1), r1 * __sinf(a1)};
#else
  return philox_normal4(c0, c1, c2, tag, seed);
#endif
}

// Stream tags (counter word 3) -- keep in sync with oracle/sts_oracle.py TAG_*.
enum : uint32_t {
  TAG_VI_INIT = 1,
  TAG_VI_EPS = 2,
  TAG_VI_DRAW = 3,
  TAG_HMC_MOM = 4,
  TAG_HMC_ACC = 5,
  TAG_ONESTEP = 6,
  TAG_FORECAST = 7,
  TAG_FORECAST_X0 = 8,
  TAG_PICK = 9,
};

}  // namespace ci
