// Backward of one highwayConv (models/TTSModel.py:63-84) in FP32 -- the first building block of the training step
// (train/adversarial_wasserstein_gp.py:277-300 back-propagates through 38 of these per Text2Mel step).
//
//   H = conv(X) + b;  h1 = LN1(H[:, :d]);  h2 = LN2(H[:, d:]);  g = sigmoid(h1);  Y = g h2 + (1 - g) X
//
// Given dY (rows = (b, t), channels last):
//   1. H is recomputed with the forward's fused conv kernel (EPI_NONE);
//   2. hwy_bwd_rows_kernel: one warp per row redoes both LayerNorms and the gate in registers, forms
//        dh2 = dY g,  dg = dY (h2 - X),  dXres = dY (1 - g),  dh1 = dg g (1 - g),
//      runs both LayerNorm backward passes (dH = rstd (dxhat - mean(dxhat) - xhat mean(dxhat xhat))), writes dH and
//      dXres, and accumulates the per-column parameter sums (d gamma, d beta of both LayerNorms, d bias of the conv)
//      over the rows of its block; the block partials are summed by a second kernel in a fixed order (no atomics);
//   3. dgrad: dX = conv(dH; W flipped in time, transposed) + dXres -- the forward's conv kernel again, with the taps
//      mirrored (a causal forward has an anti-causal backward);
//   4. wgrad: dW[co][ci][j] = sum over rows of dH[row][co] X[row shifted by tap j][ci]: tiled outer-product kernel
//      over row chunks, chunk partials summed in a fixed order.
#include "common.cuh"

namespace ssv {

namespace {

constexpr unsigned FULLM = 0xffffffffu;

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLM, v, o);
  return v;
}

constexpr int BWD_RPB = 64;          // rows per block
constexpr int BWD_WARPS = 8;

// H: [M][2d] raw conv output; X, dY: [M][d]; g*/b*: LayerNorm parameters.
// dH: [M][2d]; dXres: [M][d]; partial: [gridDim.x][6 d] = (d g1 | d b1 | d g2 | d b2 | d bias (2d)).
template <int CH>    // channels per lane = d / 32
__global__ void __launch_bounds__(32 * BWD_WARPS) hwy_bwd_rows_kernel(
    const float* __restrict__ H, const float* __restrict__ X, const float* __restrict__ dY, int M,
    const float* __restrict__ g1, const float* __restrict__ b1, const float* __restrict__ g2,
    const float* __restrict__ b2, float* __restrict__ dH, float* __restrict__ dXres, float* __restrict__ partial) {
  constexpr int d = 32 * CH;
  extern __shared__ float red[];                     // [BWD_WARPS][6 d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float ag1[CH], ab1[CH], ag2[CH], ab2[CH], abh1[CH], abh2[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) ag1[i] = ab1[i] = ag2[i] = ab2[i] = abh1[i] = abh2[i] = 0.f;
  const float inv_d = 1.0f / (float)d;
  const int row_end = min(M, (int)(blockIdx.x + 1) * BWD_RPB);
  for (int row = blockIdx.x * BWD_RPB + warp; row < row_end; row += BWD_WARPS) {
    const float* h = H + (size_t)row * 2 * d;
    float h1[CH], h2[CH], x[CH], dy[CH];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = lane + 32 * i;
      h1[i] = h[c]; h2[i] = h[d + c];
      x[i] = X[(size_t)row * d + c]; dy[i] = dY[(size_t)row * d + c];
      s1 += h1[i]; s2 += h2[i];
    }
    const float m1 = wsum(s1) * inv_d, m2 = wsum(s2) * inv_d;
    float q1 = 0.f, q2 = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const float e1 = h1[i] - m1, e2 = h2[i] - m2;
      q1 = fmaf(e1, e1, q1); q2 = fmaf(e2, e2, q2);
    }
    const float r1 = 1.0f / sqrtf(wsum(q1) * inv_d + 1e-5f), r2 = 1.0f / sqrtf(wsum(q2) * inv_d + 1e-5f);
    float dx1[CH], dx2[CH];           // d xhat of the two LayerNorms
    float t1 = 0.f, u1 = 0.f, t2 = 0.f, u2 = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = lane + 32 * i;
      const float xh1 = (h1[i] - m1) * r1, xh2 = (h2[i] - m2) * r2;
      const float ga1 = g1[c], ga2 = g2[c];
      const float a = xh1 * ga1 + b1[c], cc = xh2 * ga2 + b2[c];
      const float g = 1.0f / (1.0f + expf(-a));
      const float dh2 = dy[i] * g;
      const float dg = dy[i] * (cc - x[i]);
      const float dh1 = dg * g * (1.0f - g);
      dXres[(size_t)row * d + c] = dy[i] * (1.0f - g);
      ag1[i] = fmaf(dh1, xh1, ag1[i]); ab1[i] += dh1;
      ag2[i] = fmaf(dh2, xh2, ag2[i]); ab2[i] += dh2;
      dx1[i] = dh1 * ga1; dx2[i] = dh2 * ga2;
      t1 += dx1[i]; u1 = fmaf(dx1[i], xh1, u1);
      t2 += dx2[i]; u2 = fmaf(dx2[i], xh2, u2);
      h1[i] = xh1; h2[i] = xh2;        // keep xhat
    }
    t1 = wsum(t1) * inv_d; u1 = wsum(u1) * inv_d; t2 = wsum(t2) * inv_d; u2 = wsum(u2) * inv_d;
    float* dh = dH + (size_t)row * 2 * d;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = lane + 32 * i;
      const float o1 = r1 * (dx1[i] - t1 - h1[i] * u1), o2 = r2 * (dx2[i] - t2 - h2[i] * u2);
      dh[c] = o1; dh[d + c] = o2;
      abh1[i] += o1; abh2[i] += o2;
    }
  }
  float* mine = red + (size_t)warp * 6 * d;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = lane + 32 * i;
    mine[c] = ag1[i]; mine[d + c] = ab1[i]; mine[2 * d + c] = ag2[i]; mine[3 * d + c] = ab2[i];
    mine[4 * d + c] = abh1[i]; mine[5 * d + c] = abh2[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 6 * d; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < BWD_WARPS; ++w) s += red[(size_t)w * 6 * d + c];
    partial[(size_t)blockIdx.x * 6 * d + c] = s;
  }
}

// Forward epilogue as a separate pass over a saved H (training keeps H for the backward pass instead of
// recomputing the conv): Y = sigmoid(LN1(H1)) * LN2(H2) + (1 - sigmoid(LN1(H1))) * X, one warp per row.
template <int CH>
__global__ void __launch_bounds__(256) hwy_fwd_rows_kernel(const float* __restrict__ H, const float* __restrict__ X, int M,
                                                           const float* __restrict__ g1, const float* __restrict__ b1,
                                                           const float* __restrict__ g2, const float* __restrict__ b2,
                                                           float* __restrict__ Y) {
  constexpr int d = 32 * CH;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* h = H + (size_t)row * 2 * d;
  float h1[CH], h2[CH];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    h1[i] = h[lane + 32 * i]; h2[i] = h[d + lane + 32 * i];
    s1 += h1[i]; s2 += h2[i];
  }
  const float inv_d = 1.0f / (float)d;
  const float m1 = wsum(s1) * inv_d, m2 = wsum(s2) * inv_d;
  float q1 = 0.f, q2 = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const float e1 = h1[i] - m1, e2 = h2[i] - m2;
    q1 = fmaf(e1, e1, q1); q2 = fmaf(e2, e2, q2);
  }
  const float r1 = 1.0f / sqrtf(wsum(q1) * inv_d + 1e-5f), r2 = 1.0f / sqrtf(wsum(q2) * inv_d + 1e-5f);
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = lane + 32 * i;
    const float a = (h1[i] - m1) * r1 * g1[c] + b1[c], cc = (h2[i] - m2) * r2 * g2[c] + b2[c];
    const float g = 1.0f / (1.0f + expf(-a));
    Y[(size_t)row * d + c] = g * cc + (1.0f - g) * X[(size_t)row * d + c];
  }
}

// out[c] = sum_blk partial[blk][c], blocks in order
__global__ void colsum_partials_kernel(const float* __restrict__ partial, int nblk, int ncol, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol) return;
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += partial[(size_t)b * ncol + c];
  out[c] = s;
}

// the same, with the column range cut into up to 6 segments that go to separate destinations (the parameter gradients
// of one layer): saves a device-to-device copy per gradient on a host-bound path
__global__ void colsum_scatter_kernel(const float* __restrict__ partial, int nblk, int ncol, ColSegs segs) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol) return;
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += partial[(size_t)b * ncol + c];
#pragma unroll
  for (int i = 0; i < 6; ++i)
    if (i < segs.n && c >= segs.begin[i] && c < segs.begin[i] + segs.len[i]) segs.dst[i][c - segs.begin[i]] = s;
}

// dgrad weight: Wd[(j' * 2d + co)][ci] = conv_w[co][ci][k - 1 - j']   (conv_w is (2d, d, k))
__global__ void pack_dgrad_w_kernel(const float* __restrict__ w, int d, int k, float* __restrict__ dst) {
  const long total = (long)k * 2 * d * d;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % d);
    const int co = (int)((i / d) % (2 * d));
    const int jp = (int)(i / ((long)d * 2 * d));
    dst[i] = w[((long)co * d + ci) * k + (k - 1 - jp)];
  }
}

__global__ void add_inplace_kernel(float* __restrict__ a, const float* __restrict__ b, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) a[i] += b[i];
}

// wgrad partials: P[chunk][j][co][ci] = sum over the chunk's rows of dH[row][co] * X[(b, t + off_j)][ci].
// grid (2d / 64, d / 64, k * chunks), 256 threads = 16 x 16, 4 x 4 outputs per thread, 32 rows per smem tile.
constexpr int WG_ROWS = 32;
__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ dH, const float* __restrict__ X, int M, int T,
                                                     int d, int k, int dil, int tap_base, int chunks, int rows_per_chunk,
                                                     float* __restrict__ P) {
  __shared__ __align__(16) float As[WG_ROWS][64 + 4];       // dH tile  [row][co]
  __shared__ __align__(16) float Bs[WG_ROWS][64 + 4];       // X tile   [row][ci]
  const int co0 = blockIdx.x * 64, ci0 = blockIdx.y * 64;
  const int j = blockIdx.z % k, chunk = blockIdx.z / k;
  const int off = (tap_base + j) * dil;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const int r_lo = chunk * rows_per_chunk, r_hi = min(M, r_lo + rows_per_chunk);
  for (int r0 = r_lo; r0 < r_hi; r0 += WG_ROWS) {
    for (int i = threadIdx.x; i < WG_ROWS * 16; i += 256) {          // float4 loads: 16 per row and tile
      const int rr = i >> 4, c4 = (i & 15) * 4;
      const int row = r0 + rr;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (row < r_hi) {
        va = *reinterpret_cast<const float4*>(dH + (size_t)row * 2 * d + co0 + c4);
        const int b = row / T, t = row % T + off;
        if (t >= 0 && t < T) vb = *reinterpret_cast<const float4*>(X + ((size_t)b * T + t) * d + ci0 + c4);
      }
      *reinterpret_cast<float4*>(&As[rr][c4]) = va;
      *reinterpret_cast<float4*>(&Bs[rr][c4]) = vb;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < WG_ROWS; ++rr) {
      const float4 av = *reinterpret_cast<const float4*>(&As[rr][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[rr][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int ia = 0; ia < 4; ++ia)
#pragma unroll
        for (int ib = 0; ib < 4; ++ib) acc[ia][ib] = fmaf(a[ia], b[ib], acc[ia][ib]);
    }
    __syncthreads();
  }
  float* out = P + (((size_t)chunk * k + j) * 2 * d + co0) * d + ci0;
#pragma unroll
  for (int ia = 0; ia < 4; ++ia)
#pragma unroll
    for (int ib = 0; ib < 4; ++ib) out[(size_t)(ty * 4 + ia) * d + tx * 4 + ib] = acc[ia][ib];
}

// dW[co][ci][j] = sum_chunk P[chunk][j][co][ci]
__global__ void wgrad_reduce_kernel(const float* __restrict__ P, int chunks, int d, int k, float* __restrict__ dW) {
  const long per = (long)k * 2 * d * d;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < per; i += (long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < chunks; ++c) s += P[(size_t)c * per + i];
    const int ci = (int)(i % d);
    const int co = (int)((i / d) % (2 * d));
    const int j = (int)(i / ((long)d * 2 * d));
    dW[((long)co * d + ci) * k + j] = s;
  }
}

}  // namespace

int launch_hwy_bwd_rows(const float* H, const float* X, const float* dY, int M, int d, const float* g1, const float* b1,
                        const float* g2, const float* b2, float* dH, float* dXres, float* partial, int* nblk_out,
                        cudaStream_t s) {
  const int nblk = (M + BWD_RPB - 1) / BWD_RPB;
  const size_t smem = (size_t)BWD_WARPS * 6 * d * sizeof(float);
  if (d == 256) {
    SSV_CUDA(cudaFuncSetAttribute(hwy_bwd_rows_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hwy_bwd_rows_kernel<8><<<nblk, 32 * BWD_WARPS, smem, s>>>(H, X, dY, M, g1, b1, g2, b2, dH, dXres, partial);
  } else if (d == 512) {
    SSV_CUDA(cudaFuncSetAttribute(hwy_bwd_rows_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hwy_bwd_rows_kernel<16><<<nblk, 32 * BWD_WARPS, smem, s>>>(H, X, dY, M, g1, b1, g2, b2, dH, dXres, partial);
  } else {
    set_error("highway_conv_bwd: dimension %d unsupported (256 or 512)", d);
    return kInval;
  }
  SSV_CUDA(cudaGetLastError());
  ++g_launches;
  *nblk_out = nblk;
  return kOk;
}

int launch_hwy_fwd_rows(const float* H, const float* X, int M, int d, const float* g1, const float* b1, const float* g2,
                        const float* b2, float* Y, cudaStream_t s) {
  const int grid = (M + 7) / 8;
  if (d == 256) hwy_fwd_rows_kernel<8><<<grid, 256, 0, s>>>(H, X, M, g1, b1, g2, b2, Y);
  else if (d == 512) hwy_fwd_rows_kernel<16><<<grid, 256, 0, s>>>(H, X, M, g1, b1, g2, b2, Y);
  else { set_error("highway_conv: dimension %d unsupported (256 or 512)", d); return kInval; }
  SSV_CUDA(cudaGetLastError());
  ++g_launches;
  return kOk;
}

int hwy_bwd_row_blocks(int M) { return (M + BWD_RPB - 1) / BWD_RPB; }

int launch_colsum_partials(const float* partial, int nblk, int ncol, float* out, cudaStream_t s) {
  colsum_partials_kernel<<<(ncol + 255) / 256, 256, 0, s>>>(partial, nblk, ncol, out);
  SSV_CUDA(cudaGetLastError());
  ++g_launches;
  return kOk;
}

int launch_colsum_scatter(const float* partial, int nblk, int ncol, const ColSegs& segs, cudaStream_t s) {
  colsum_scatter_kernel<<<(ncol + 255) / 256, 256, 0, s>>>(partial, nblk, ncol, segs);
  SSV_CUDA(cudaGetLastError());
  ++g_launches;
  return kOk;
}

int launch_pack_dgrad_w(const float* conv_w, int d, int k, float* dst, cudaStream_t s) {
  pack_dgrad_w_kernel<<<1024, 256, 0, s>>>(conv_w, d, k, dst);
  SSV_CUDA(cudaGetLastError());
  ++g_launches;
  return kOk;
}

int launch_add_inplace(float* a, const float* b, long n, cudaStream_t s) {
  long g = (n + 255) / 256;
  if (g > 4096) g = 4096;
  add_inplace_kernel<<<(int)g, 256, 0, s>>>(a, b, n);
  SSV_CUDA(cudaGetLastError());
  ++g_launches;
  return kOk;
}

int wgrad_chunks(int M) {
  int c = (M + 511) / 512;
  return c < 1 ? 1 : (c > 32 ? 32 : c);
}

int launch_wgrad(const float* dH, const float* X, int M, int T, int d, int k, int dil, int causal, float* P, float* dW,
                 cudaStream_t s) {
  const int chunks = wgrad_chunks(M);
  const int rpc = ((M + chunks - 1) / chunks + WG_ROWS - 1) / WG_ROWS * WG_ROWS;
  const int tap_base = causal ? -(k - 1) : -((k - 1) / 2);
  wgrad_kernel<<<dim3(2 * d / 64, d / 64, k * chunks), 256, 0, s>>>(dH, X, M, T, d, k, dil, tap_base, chunks, rpc, P);
  SSV_CUDA(cudaGetLastError());
  wgrad_reduce_kernel<<<2048, 256, 0, s>>>(P, chunks, d, k, dW);
  SSV_CUDA(cudaGetLastError());
  g_launches += 2;
  return kOk;
}

}  // namespace ssv
