// Microbenchmark: does the parity of the two fresh register operands of an FFMA decide its issue rate on sm_100a?
// Emulates the mat-vec tile loop of decode_ws_kernel (4-row micro-batch: one LDS.128 of x, a float4 of weights in
// registers, 16 accumulators).  MODE 0: plain C++ (ptxas places the accumulators).  MODE 1: the accumulators live in
// 64-bit pairs (lo = even register, hi = odd register) holding columns (1, 0) and (3, 2), so that every FFMA pairs an
// accumulator with a weight register of the opposite parity; row order.  MODE 2: same with a snake order through the
// 4 x 4 tile, so that consecutive FFMAs always share x or w (one operand from the reuse cache).
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NW = 16;
__device__ __forceinline__ void fma2s(unsigned long long& p, float wlo, float whi, float x) {
  asm volatile("{\n\t.reg .f32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tfma.rn.f32 lo, %1, %3, lo;\n\tfma.rn.f32 hi, %2, %3, hi;\n\tmov.b64 %0, {lo, hi};\n\t}"
               : "+l"(p) : "f"(wlo), "f"(whi), "f"(x));
}
__device__ __forceinline__ void fma_lo(unsigned long long& p, float w, float x) {
  asm volatile("{\n\t.reg .f32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tfma.rn.f32 lo, %1, %2, lo;\n\tmov.b64 %0, {lo, hi};\n\t}" : "+l"(p) : "f"(w), "f"(x));
}
__device__ __forceinline__ void fma_hi(unsigned long long& p, float w, float x) {
  asm volatile("{\n\t.reg .f32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tfma.rn.f32 hi, %1, %2, hi;\n\tmov.b64 %0, {lo, hi};\n\t}" : "+l"(p) : "f"(w), "f"(x));
}
template <int MODE>
__global__ void __launch_bounds__(384, 1) k(float* out, const float4* wsrc, int iters, long long* cyc) {
  __shared__ float4 xs[NW * 24];
  for (int i = threadIdx.x; i < NW * 24; i += blockDim.x) xs[i] = make_float4(i * 1e-3f, 1.f - i * 1e-3f, 0.5f, 0.25f + i * 1e-4f);
  float4 w[NW];
#pragma unroll
  for (int j = 0; j < NW; ++j) w[j] = wsrc[j * 384 + threadIdx.x];
  __syncthreads();
  const float4* xp = xs + (threadIdx.x / 16);
  float acc[4][4];
  unsigned long long p[4][2];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    p[r][0] = p[r][1] = 0ull;
  }
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < NW; ++j) {
      const float4 v = xp[j * 24];
      const float x[4] = {v.x, v.y, v.z, v.w};
      if (MODE == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          acc[r][0] = fmaf(x[r], w[j].x, acc[r][0]);
          acc[r][1] = fmaf(x[r], w[j].y, acc[r][1]);
          acc[r][2] = fmaf(x[r], w[j].z, acc[r][2]);
          acc[r][3] = fmaf(x[r], w[j].w, acc[r][3]);
        }
      } else if (MODE == 1) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          fma2s(p[r][0], w[j].y, w[j].x, x[r]);      // lo (even) = column 1 with w.y (odd), hi (odd) = column 0 with w.x (even)
          fma2s(p[r][1], w[j].w, w[j].z, x[r]);
        }
      } else {
        // snake: row 0 columns 0..3, row 1 columns 3..0, row 2 columns 0..3, row 3 columns 3..0
        fma_hi(p[0][0], w[j].x, x[0]); fma_lo(p[0][0], w[j].y, x[0]); fma_hi(p[0][1], w[j].z, x[0]); fma_lo(p[0][1], w[j].w, x[0]);
        fma_lo(p[1][1], w[j].w, x[1]); fma_hi(p[1][1], w[j].z, x[1]); fma_lo(p[1][0], w[j].y, x[1]); fma_hi(p[1][0], w[j].x, x[1]);
        fma_hi(p[2][0], w[j].x, x[2]); fma_lo(p[2][0], w[j].y, x[2]); fma_hi(p[2][1], w[j].z, x[2]); fma_lo(p[2][1], w[j].w, x[2]);
        fma_lo(p[3][1], w[j].w, x[3]); fma_hi(p[3][1], w[j].z, x[3]); fma_lo(p[3][0], w[j].y, x[3]); fma_hi(p[3][0], w[j].x, x[3]);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int c = 0; c < 4; ++c) s += acc[r][c];
    s += __uint_as_float((unsigned)p[r][0]) + __uint_as_float((unsigned)(p[r][0] >> 32)) + __uint_as_float((unsigned)p[r][1]) + __uint_as_float((unsigned)(p[r][1] >> 32));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[MODE] = t1 - t0;
}
int main() {
  float* out; float4* w; long long* cyc;
  cudaMalloc(&out, 148 * 384 * sizeof(float));
  cudaMalloc(&w, NW * 384 * sizeof(float4));
  cudaMemset(w, 0x3c, NW * 384 * sizeof(float4));
  cudaMallocManaged(&cyc, 3 * sizeof(long long));
  const int iters = 4000;
  for (int rep = 0; rep < 2; ++rep) {
    k<0><<<148, 384>>>(out, w, iters, cyc);
    k<1><<<148, 384>>>(out, w, iters, cyc);
    k<2><<<148, 384>>>(out, w, iters, cyc);
    cudaDeviceSynchronize();
    for (int m = 0; m < 3; ++m) {
      const double ffma_per_sched = 3.0 * NW * 16.0 * iters;      // 12 warps on 4 schedulers
      printf("mode %d: %lld cycles, %.3f FFMA / cycle / scheduler\n", m, cyc[m], ffma_per_sched / (double)cyc[m]);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
