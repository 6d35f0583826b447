// Microbenchmark: issue interval of mma.sync.m16n8k8 TF32 (HMMA.1688.F32.TF32) on sm_100a for ONE warp per scheduler
// with 1 / 2 / 4 / 8 independent accumulator chains, against FFMA2 -- decides whether the 3xTF32 regression passes of
// k_eval_coop can beat the FFMA2 passes.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_tf32_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k_mma(float* out, int iters, long long* cyc) {
  float d[CH][4];
  for (int c = 0; c < CH; ++c) d[c][0] = d[c][1] = d[c][2] = d[c][3] = threadIdx.x * 1e-3f;
  unsigned a0 = __float_as_uint(1.0f + threadIdx.x), a2 = __float_as_uint(0.5f), b0 = __float_as_uint(0.25f), b1 = __float_as_uint(2.f);
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int c = 0; c < CH; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH>
__global__ void k_ffma2(float* out, int iters, long long* cyc) {
  float2 d[CH];
  for (int c = 0; c < CH; ++c) d[c] = make_float2(threadIdx.x * 1e-3f, 1.f);
  const float2 a = make_float2(1.0001f, 0.9999f), b = make_float2(1e-6f, 2e-6f);
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) d[c] = __ffma2_rn(d[c], a, b);
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int c = 0; c < CH; ++c) s += d[c].x + d[c].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <class K>
void run(const char* name, K kern, int ch, int threads) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  kern<<<148, threads>>>(out, iters, cyc); kern<<<148, threads>>>(out, iters, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("{\"kernel\": \"%s\", \"chains\": %d, \"warps_per_sm\": %d, \"cycles_per_instruction_per_warp\": %.2f}\n", name, ch, threads / 32,
         (double)h / iters / ch);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run("hmma.1688.f32.tf32", k_mma<1>, 1, 128); run("hmma.1688.f32.tf32", k_mma<2>, 2, 128);
  run("hmma.1688.f32.tf32", k_mma<4>, 4, 128); run("hmma.1688.f32.tf32", k_mma<8>, 8, 128);
  run("hmma.1688.f32.tf32", k_mma<8>, 8, 256); run("hmma.1688.f32.tf32", k_mma<8>, 8, 512);
  run("ffma2", k_ffma2<1>, 1, 128); run("ffma2", k_ffma2<8>, 8, 128); run("ffma2", k_ffma2<8>, 8, 512);
  return 0;
}
