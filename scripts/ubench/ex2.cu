// MUFU micro-benchmark: ex2.approx.ftz.f32 vs ex2.approx.ftz.f16x2 throughput per SM (elements / clk), to decide whether the
// attention softmax should exponentiate packed halves.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ex2 ex2.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned ex2h2(unsigned x) { unsigned y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
template <int MODE> __global__ void k(float* out, int iters) {
  float a[8]; unsigned h[8];
  for (int i = 0; i < 8; i++) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0xb800b800u + threadIdx.x + i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) a[i] = ex2f(a[i]) - 1.0f;                       // keep the argument bounded: 1 MUFU + 1 FADD
      else h[i] = ex2h2(h[i]) ^ 0x80008000u;                          // 1 MUFU(f16x2) + 1 LOP
    }
  }
  float s = 0; for (int i = 0; i < 8; i++) s += a[i] + (float)h[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; mode++) {
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148, 1024>>>(d, iters); else k<1><<<148, 1024>>>(d, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep) {
        const double ops = 148.0 * 1024 * 8.0 * iters;                // MUFU instructions (per thread op)
        printf("%s: %.3f ms, %.2f MUFU lane-ops / clk / SM at 1.9 GHz (elements / clk / SM: %.2f)\n", mode ? "ex2.f16x2" : "ex2.f32", ms,
               ops / 148.0 / (ms * 1e-3 * 1.9e9), (mode ? 2.0 : 1.0) * ops / 148.0 / (ms * 1e-3 * 1.9e9));
      }
    }
  }
  return 0;
}
