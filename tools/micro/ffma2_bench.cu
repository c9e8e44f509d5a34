// Microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int iters) {
  float2 acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) { acc[i].x = fmaf(acc[i].x, aa.x, bb.x); acc[i].y = fmaf(acc[i].y, aa.y, bb.y); }
      else acc[i] = __ffma2_rn(acc[i], aa, bb);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 4, 256>>>(out, 1.0001f, 0.5f, iters); else k<1><<<148 * 4, 256>>>(out, 1.0001f, 0.5f, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma = 148.0 * 4 * 256 * 32.0 * iters;
      printf("%s: %.3f ms  %.1f TFLOP/s\n", mode == 0 ? "FFMA " : "FFMA2", ms, 2 * fma / ms * 1e-9);
    }
  }
  return 0;
}
