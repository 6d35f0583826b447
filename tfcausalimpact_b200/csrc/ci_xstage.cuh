// TMA-staged design-matrix chunks: per-warp staging buffer + mbarrier (shared by the thread-per-lane evaluation
// kernel, ci_eval.cu, and the lane-cooperative one, ci_eval_coop.cu).
#pragma once
#include <stdint.h>
#include "ci_eval_lane.cuh"

namespace ci {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// TMA-staged policy: every WARP owns a staging buffer + mbarrier.  The series its 32 lanes belong to
// are adjacent, so with the [chunk][series][k+1][4] layout the run it needs for one chunk is a single
// contiguous block: ONE cp.async.bulk per (warp, chunk), issued by lane 0 as soon as the warp has
// consumed the previous chunk (__syncwarp -- no CTA-wide barrier, warps never wait for each other).
struct XStaged {
  const float* lane_x;   // this lane's series slice of the warp's staging buffer (shared memory)
  uint32_t bar;          // warp mbarrier, shared-space address; the copy descriptor sits 16 bytes behind it
  uint32_t parity;
  // Copy descriptor in shared memory (written once per warp, read by lane 0 when it issues a copy), so that the
  // warp-uniform copy parameters do not occupy registers in every lane for the whole kernel:
  //   [bar + 16] global source of chunk 0 (8 bytes), [bar + 24] bytes per chunk, [bar + 28] staging buffer address,
  //   [bar + 32] bytes between consecutive chunks (8 bytes)
  __device__ __forceinline__ void issue(int tc) const {   // lane 0 of the warp
    uint64_t src, cbytes;
    uint32_t bytes, sx_addr;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(src) : "r"(bar + 16));
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(bytes), "=r"(sx_addr) : "r"(bar + 24));
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(cbytes) : "r"(bar + 32));
    src += (uint64_t)tc * cbytes;
    // (measured: dropping this proxy fence gains 0.2 % -- kept, the cross-proxy write-after-read stays explicit)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     sx_addr),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
  }
  __device__ __forceinline__ void acquire() {
    uint32_t ok;
    do {
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
          : "=r"(ok)
          : "r"(bar), "r"(parity)
          : "memory");
    } while (!ok);
    parity ^= 1u;
  }
  __device__ __forceinline__ void release(int next_tc) const {
    __syncwarp();   // every lane of the warp (padding lanes included) has consumed the staged chunk
    if ((threadIdx.x & 31) == 0 && next_tc >= 0) issue(next_tc);
  }
  __device__ __forceinline__ const float* chunk(int) const { return lane_x; }
  __device__ __forceinline__ constexpr int rs() const { return CI_TC; }
  __device__ __forceinline__ f4 load(const float* p) const {
    const float4 v = *reinterpret_cast<const float4*>(p);
    return f4{v.x, v.y, v.z, v.w};
  }
};

}  // namespace ci
