// The small layers of the Text2Mel training graph in FP32 (train/adversarial_wasserstein_gp.py:277-300 back-propagates
// through them every generator iteration): the eleven 1x1 conv + LayerNorm layers (models/TTSModel.py:128-131, 173-180,
// 218-230), the unmasked softmax attention of the train branch (:268-272), the text embedding (:25-35) and the two
// speaker projections (:172-173).  The highway convs -- 90+ % of the FLOPs -- are in backward.cu; with these the
// generator's forward / backward graph no longer calls a library GEMM.
//
// All kernels work on channels-last rows (row = (b, t)); fixed-order reductions, no atomics.
#include "common.cuh"

namespace ssv {

namespace {

constexpr unsigned FULLM = 0xffffffffu;
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLM, v, o);
  return v;
}

// ---- LayerNorm over the n real columns of H rows, one warp per row; CH = ceil(n / 32) channels per lane
template <int CH>
__global__ void __launch_bounds__(256) ln_rows_fwd_kernel(const float* __restrict__ H, int ldh, int M, int n,
                                                          const float* __restrict__ g, const float* __restrict__ b,
                                                          float* __restrict__ Y, int ldy) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* h = H + (size_t)row * ldh;
  float v[CH];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < n ? h[c] : 0.f;
    s += v[i];
  }
  const float mean = wsum(s) / (float)n;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const float e = lane + 32 * i < n ? v[i] - mean : 0.f;
    q = fmaf(e, e, q);
  }
  const float rstd = 1.0f / sqrtf(wsum(q) / (float)n + 1e-5f);
  float* y = Y + (size_t)row * ldy;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = lane + 32 * i;
    if (c < ldy) y[c] = c < n ? (v[i] - mean) * rstd * g[c] + b[c] : 0.f;
  }
}

constexpr int LB_RPB = 64, LB_WARPS = 8;
// dH = rstd (g dY - mean(g dY) - xhat mean(g dY xhat)); partial[blk] = (d gamma | d beta) summed over the block's rows.
template <int CH>
__global__ void __launch_bounds__(32 * LB_WARPS) ln_rows_bwd_kernel(const float* __restrict__ H, int ldh, const float* __restrict__ dY,
                                                                    int ldy, int M, int n, const float* __restrict__ g,
                                                                    float* __restrict__ dH, float* __restrict__ partial) {
  extern __shared__ float red[];                     // [LB_WARPS][2 * 32 * CH]
  constexpr int NP = 32 * CH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float ag[CH], ab[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) ag[i] = ab[i] = 0.f;
  const float inv_n = 1.0f / (float)n;
  const int row_end = min(M, (int)(blockIdx.x + 1) * LB_RPB);
  for (int row = blockIdx.x * LB_RPB + warp; row < row_end; row += LB_WARPS) {
    const float* h = H + (size_t)row * ldh;
    const float* dy = dY + (size_t)row * ldy;
    float v[CH], d[CH];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < n ? h[c] : 0.f;
      d[i] = c < n ? dy[c] : 0.f;
      s += v[i];
    }
    const float mean = wsum(s) * inv_n;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const float e = lane + 32 * i < n ? v[i] - mean : 0.f;
      q = fmaf(e, e, q);
    }
    const float rstd = 1.0f / sqrtf(wsum(q) * inv_n + 1e-5f);
    float t = 0.f, u = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = lane + 32 * i;
      const float xh = c < n ? (v[i] - mean) * rstd : 0.f;
      ag[i] = fmaf(d[i], xh, ag[i]);
      ab[i] += d[i];
      const float dx = c < n ? d[i] * g[c] : 0.f;
      t += dx;
      u = fmaf(dx, xh, u);
      v[i] = xh;
      d[i] = dx;
    }
    t = wsum(t) * inv_n;
    u = wsum(u) * inv_n;
    float* o = dH + (size_t)row * ldh;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = lane + 32 * i;
      if (c < ldh) o[c] = c < n ? rstd * (d[i] - t - v[i] * u) : 0.f;
    }
  }
  float* mine = red + (size_t)warp * 2 * NP;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    mine[lane + 32 * i] = ag[i];
    mine[NP + lane + 32 * i] = ab[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * NP; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LB_WARPS; ++w) s += red[(size_t)w * 2 * NP + c];
    partial[(size_t)blockIdx.x * 2 * NP + c] = s;
  }
}

// U[b][c] = sum over the T rows of utterance b of X[(b, t)][c]
__global__ void utt_colsum_kernel(const float* __restrict__ X, int ld, int T, int n, float* __restrict__ U) {
  const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  float s = 0.f;
  for (int t = 0; t < T; ++t) s += X[((size_t)b * T + t) * ld + c];
  U[(size_t)b * n + c] = s;
}

__global__ void relu_rows_kernel(float* __restrict__ x, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) x[i] = fmaxf(x[i], 0.f);
}
// dx *= (x > 0)
__global__ void relu_mask_kernel(float* __restrict__ dx, const float* __restrict__ x, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    if (!(x[i] > 0.f)) dx[i] = 0.f;
}
// dst[r][c] (rows x ld, zero padded) = src[r][c] (r < n_rows, c < n_cols)
__global__ void pad_matrix_kernel(const float* __restrict__ src, int n_rows, int n_cols, int rows, int ld, float* __restrict__ dst) {
  const long total = (long)rows * ld;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ld), c = (int)(i % ld);
    dst[i] = (r < n_rows && c < n_cols) ? src[(size_t)r * n_cols + c] : 0.f;
  }
}

// wgrad of a 1x1 layer: P[chunk][co][ci] = sum over the chunk's rows of dH[row][co] X[row][ci]; 64 x 64 outputs per block.
constexpr int WG_ROWS = 32;
__global__ void __launch_bounds__(256) wgrad_plain_kernel(const float* __restrict__ dH, int ldh, const float* __restrict__ X, int ldx,
                                                           int M, int n, int cin, int rows_per_chunk, float* __restrict__ P) {
  __shared__ __align__(16) float As[WG_ROWS][64 + 4];
  __shared__ __align__(16) float Bs[WG_ROWS][64 + 4];
  const int co0 = blockIdx.x * 64, ci0 = blockIdx.y * 64, chunk = blockIdx.z;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const int r_lo = chunk * rows_per_chunk, r_hi = min(M, r_lo + rows_per_chunk);
  for (int r0 = r_lo; r0 < r_hi; r0 += WG_ROWS) {
    for (int i = threadIdx.x; i < WG_ROWS * 64; i += 256) {
      const int rr = i >> 6, c = i & 63;
      const int row = r0 + rr;
      As[rr][c] = (row < r_hi && co0 + c < n) ? dH[(size_t)row * ldh + co0 + c] : 0.f;
      Bs[rr][c] = (row < r_hi && ci0 + c < cin) ? X[(size_t)row * ldx + ci0 + c] : 0.f;
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
#pragma unroll
  for (int ia = 0; ia < 4; ++ia)
#pragma unroll
    for (int ib = 0; ib < 4; ++ib) {
      const int co = co0 + ty * 4 + ia, ci = ci0 + tx * 4 + ib;
      if (co < n && ci < cin) P[((size_t)chunk * n + co) * cin + ci] = acc[ia][ib];
    }
}
__global__ void sum_chunks_kernel(const float* __restrict__ P, int chunks, long per, float* __restrict__ out) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < per; i += (long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < chunks; ++c) s += P[(size_t)c * per + i];
    out[i] = s;
  }
}
// float4 version for the split-K partial outputs of the conv kernels (per % 4 == 0)
__global__ void sum_chunks4_kernel(const float4* __restrict__ P, int chunks, long per4, float4* __restrict__ out) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < per4; i += (long)gridDim.x * blockDim.x) {
    float4 s = P[i];
    for (int c = 1; c < chunks; ++c) {
      const float4 v = P[(size_t)c * per4 + i];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    out[i] = s;
  }
}

// ---- attention backward, train branch: A = softmax_n(K^T q / 16), R = V A.
// One warp per (b, t): g[n] = dA[n, t] + V[n, :] . dR[t, :];  dS[n] = A[n, t] (g[n] - sum_n' A[n', t] g[n']);
// dq[t, :] = (1 / 16) sum_n dS[n] K[n, :].  Kx is (B, N, 512) = [K | V] rows, dR / dq (B, T, 256), A / dA / dS (B, N, T).
constexpr int AB_W = 8, AB_NMAX = 192;
__global__ void __launch_bounds__(AB_W * 32) att_bwd_q_kernel(const float* __restrict__ Kx, const float* __restrict__ dR, int ldr,
                                                              const float* __restrict__ dq_add, const float* __restrict__ A,
                                                              const float* __restrict__ dA, int N, int T, float* __restrict__ dS,
                                                              float* __restrict__ dq) {
  __shared__ __align__(16) float rs[AB_W][256];
  __shared__ float ss[AB_W][AB_NMAX];
  const int b = blockIdx.y, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * AB_W + w;
  if (t >= T) return;
  const float* dr = dR + ((size_t)b * T + t) * ldr;
  *reinterpret_cast<float4*>(&rs[w][lane * 8]) = *reinterpret_cast<const float4*>(dr + lane * 8);
  *reinterpret_cast<float4*>(&rs[w][lane * 8 + 4]) = *reinterpret_cast<const float4*>(dr + lane * 8 + 4);
  __syncwarp();
  float gv[AB_NMAX / 32], av[AB_NMAX / 32];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < AB_NMAX / 32; ++i) {
    const int n = lane + 32 * i;
    gv[i] = av[i] = 0.f;
    if (n < N) {
      const float4* vr = reinterpret_cast<const float4*>(Kx + ((size_t)b * N + n) * 512 + 256);
      const float4* rv = reinterpret_cast<const float4*>(rs[w]);
      float acc = 0.f;
#pragma unroll 8
      for (int c = 0; c < 64; ++c) {
        const float4 vv = vr[c], rr = rv[c];
        acc = fmaf(vv.x, rr.x, acc); acc = fmaf(vv.y, rr.y, acc); acc = fmaf(vv.z, rr.z, acc); acc = fmaf(vv.w, rr.w, acc);
      }
      const size_t ai = ((size_t)b * N + n) * T + t;
      av[i] = A[ai];
      gv[i] = acc + (dA ? dA[ai] : 0.f);
      dot = fmaf(av[i], gv[i], dot);
    }
  }
  dot = wsum(dot);
#pragma unroll
  for (int i = 0; i < AB_NMAX / 32; ++i) {
    const int n = lane + 32 * i;
    if (n < N) {
      const float s = av[i] * (gv[i] - dot);
      dS[((size_t)b * N + n) * T + t] = s;
      ss[w][n] = s;
    }
  }
  __syncwarp();
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int n = 0; n < N; ++n) {
    const float s = ss[w][n];
    const float* kr = Kx + ((size_t)b * N + n) * 512 + lane * 8;
    const float4 k0 = *reinterpret_cast<const float4*>(kr), k1 = *reinterpret_cast<const float4*>(kr + 4);
    acc[0] = fmaf(s, k0.x, acc[0]); acc[1] = fmaf(s, k0.y, acc[1]); acc[2] = fmaf(s, k0.z, acc[2]); acc[3] = fmaf(s, k0.w, acc[3]);
    acc[4] = fmaf(s, k1.x, acc[4]); acc[5] = fmaf(s, k1.y, acc[5]); acc[6] = fmaf(s, k1.z, acc[6]); acc[7] = fmaf(s, k1.w, acc[7]);
  }
  float add[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};       // gradient that reaches q directly (the [R ; q] concatenation)
  if (dq_add) {
    const float* ap = dq_add + ((size_t)b * T + t) * ldr + lane * 8;
    const float4 a0 = *reinterpret_cast<const float4*>(ap), a1 = *reinterpret_cast<const float4*>(ap + 4);
    add[0] = a0.x; add[1] = a0.y; add[2] = a0.z; add[3] = a0.w; add[4] = a1.x; add[5] = a1.y; add[6] = a1.z; add[7] = a1.w;
  }
  float* o = dq + ((size_t)b * T + t) * 256 + lane * 8;
  *reinterpret_cast<float4*>(o) = make_float4(fmaf(acc[0], 0.0625f, add[0]), fmaf(acc[1], 0.0625f, add[1]), fmaf(acc[2], 0.0625f, add[2]), fmaf(acc[3], 0.0625f, add[3]));
  *reinterpret_cast<float4*>(o + 4) = make_float4(fmaf(acc[4], 0.0625f, add[4]), fmaf(acc[5], 0.0625f, add[5]), fmaf(acc[6], 0.0625f, add[6]), fmaf(acc[7], 0.0625f, add[7]));
}
// One warp per (b, n): dK[n, :] = (1 / 16) sum_t dS[n, t] q[t, :],  dV[n, :] = sum_t A[n, t] dR[t, :]  -> dKx (B, N, 512)
__global__ void __launch_bounds__(AB_W * 32) att_bwd_kv_kernel(const float* __restrict__ Q, const float* __restrict__ dR, int ldr,
                                                               const float* __restrict__ A, const float* __restrict__ dS, int N, int T,
                                                               float* __restrict__ dKx) {
  const int b = blockIdx.y, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * AB_W + w;
  if (n >= N) return;
  float ak[8], avv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) ak[i] = avv[i] = 0.f;
  const float* arow = A + ((size_t)b * N + n) * T;
  const float* srow = dS + ((size_t)b * N + n) * T;
  for (int t = 0; t < T; ++t) {
    const float a = arow[t], s = srow[t];
    const float* q = Q + ((size_t)b * T + t) * 256 + lane * 8;
    const float* r = dR + ((size_t)b * T + t) * ldr + lane * 8;
    const float4 q0 = *reinterpret_cast<const float4*>(q), q1 = *reinterpret_cast<const float4*>(q + 4);
    const float4 r0 = *reinterpret_cast<const float4*>(r), r1 = *reinterpret_cast<const float4*>(r + 4);
    ak[0] = fmaf(s, q0.x, ak[0]); ak[1] = fmaf(s, q0.y, ak[1]); ak[2] = fmaf(s, q0.z, ak[2]); ak[3] = fmaf(s, q0.w, ak[3]);
    ak[4] = fmaf(s, q1.x, ak[4]); ak[5] = fmaf(s, q1.y, ak[5]); ak[6] = fmaf(s, q1.z, ak[6]); ak[7] = fmaf(s, q1.w, ak[7]);
    avv[0] = fmaf(a, r0.x, avv[0]); avv[1] = fmaf(a, r0.y, avv[1]); avv[2] = fmaf(a, r0.z, avv[2]); avv[3] = fmaf(a, r0.w, avv[3]);
    avv[4] = fmaf(a, r1.x, avv[4]); avv[5] = fmaf(a, r1.y, avv[5]); avv[6] = fmaf(a, r1.z, avv[6]); avv[7] = fmaf(a, r1.w, avv[7]);
  }
  float* o = dKx + ((size_t)b * N + n) * 512 + lane * 8;
  *reinterpret_cast<float4*>(o) = make_float4(ak[0] * 0.0625f, ak[1] * 0.0625f, ak[2] * 0.0625f, ak[3] * 0.0625f);
  *reinterpret_cast<float4*>(o + 4) = make_float4(ak[4] * 0.0625f, ak[5] * 0.0625f, ak[6] * 0.0625f, ak[7] * 0.0625f);
  *reinterpret_cast<float4*>(o + 256) = make_float4(avv[0], avv[1], avv[2], avv[3]);
  *reinterpret_cast<float4*>(o + 260) = make_float4(avv[4], avv[5], avv[6], avv[7]);
}

// h[(b, t)][c] += sb[b][c]: the per-utterance speaker term of AudioEnc's conv1 / conv3 after a tensor-core conv
__global__ void add_utt_bias_kernel(float* __restrict__ h, int ld, int B, int T, int n, const float* __restrict__ sb) {
  const long total = (long)B * T * n;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % n);
    const long r = i / n;
    h[r * ld + c] += sb[(r / T) * n + c];
  }
}

// ---- embedding backward: dWt[v][c] = sum over the positions with id v of dX[pos][c];  block v < vocab, block vocab: d bias.
// Block (v, chunk): 8 row groups x 128 channel lanes over the chunk's rows, fixed-order sums (deterministic); the chunks'
// partial sums are added by embed_bwd_sum_kernel.  (One block per v walking all rows serially took 470 us at 2048 rows.)
constexpr int EB_CHUNKS = 8;
__global__ void __launch_bounds__(1024) embed_bwd_kernel(const int64_t* __restrict__ ids, int rows, const float* __restrict__ dX, int ld,
                                                         int vocab, int E, float* __restrict__ P) {
  __shared__ float red[8][128];
  const int v = blockIdx.x, chunk = blockIdx.y;
  const int tx = threadIdx.x & 127, ty = threadIdx.x >> 7;
  const int rpc = (rows + EB_CHUNKS - 1) / EB_CHUNKS;
  const int r_end = min(rows, (chunk + 1) * rpc);
  for (int c0 = 0; c0 < E; c0 += 128) {
    const int c = c0 + tx;
    float s = 0.f;
    if (c < E)
      for (int r = chunk * rpc + ty; r < r_end; r += 8)
        if (v == vocab || ids[r] == v) s += dX[(size_t)r * ld + c];
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && c < E) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i][tx];
      P[((size_t)chunk * (vocab + 1) + v) * E + c] = t;
    }
    __syncthreads();
  }
}
__global__ void embed_bwd_sum_kernel(const float* __restrict__ P, int vocab, int E, float* __restrict__ dWt, float* __restrict__ dbias) {
  const int n = (vocab + 1) * E;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < EB_CHUNKS; ++c) s += P[(size_t)c * n + i];
    if (i < vocab * E) dWt[i] = s;
    else dbias[i - vocab * E] = s;
  }
}

// ---- speaker projection backward: dW[o][i] = sum_b dS[b][o] x[b][i],  db[o] = sum_b dS[b][o]
__global__ void linear_small_bwd_kernel(const float* __restrict__ dS, int ds_ld, const float* __restrict__ x, long x_ld, int B,
                                        int in_f, int out_f, float* __restrict__ dW, float* __restrict__ db) {
  const int o = blockIdx.x;
  for (int i = threadIdx.x; i <= in_f; i += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(dS[(size_t)b * ds_ld + o], i < in_f ? x[(size_t)b * x_ld + i] : 1.f, s);
    if (i < in_f) dW[(size_t)o * in_f + i] = s;
    else db[o] = s;
  }
}

inline int grid_for(long n) {
  long g = (n + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 4096 ? 4096 : g));
}

}  // namespace

int launch_ln_rows_fwd(const float* H, int ldh, int M, int n, const float* g, const float* b, float* Y, int ldy, cudaStream_t s) {
  const int grid = (M + 7) / 8;
  SSV_CHECK(n <= 512 && ldy <= (n <= 128 ? 128 : n <= 256 ? 256 : 512), "conv_ln: %d channels / row stride %d unsupported", n, ldy);
  if (n <= 128) ln_rows_fwd_kernel<4><<<grid, 256, 0, s>>>(H, ldh, M, n, g, b, Y, ldy);
  else if (n <= 256) ln_rows_fwd_kernel<8><<<grid, 256, 0, s>>>(H, ldh, M, n, g, b, Y, ldy);
  else ln_rows_fwd_kernel<16><<<grid, 256, 0, s>>>(H, ldh, M, n, g, b, Y, ldy);
  SSV_CUDA(cudaGetLastError());
  ++g_launches;
  return kOk;
}

int ln_bwd_row_blocks(int M) { return (M + LB_RPB - 1) / LB_RPB; }
int ln_bwd_partial_cols(int n) { return 2 * 32 * (n <= 128 ? 4 : n <= 256 ? 8 : 16); }

int launch_ln_rows_bwd(const float* H, int ldh, const float* dY, int ldy, int M, int n, const float* g, float* dH, float* partial,
                       cudaStream_t s) {
  const int nblk = ln_bwd_row_blocks(M);
  const int ch = n <= 128 ? 4 : n <= 256 ? 8 : 16;
  SSV_CHECK(n <= 512 && ldh <= 32 * ch, "conv_ln backward: %d channels / row stride %d unsupported", n, ldh);
  const size_t smem = (size_t)LB_WARPS * 2 * 32 * ch * sizeof(float);
  if (ch == 4) ln_rows_bwd_kernel<4><<<nblk, 32 * LB_WARPS, smem, s>>>(H, ldh, dY, ldy, M, n, g, dH, partial);
  else if (ch == 8) ln_rows_bwd_kernel<8><<<nblk, 32 * LB_WARPS, smem, s>>>(H, ldh, dY, ldy, M, n, g, dH, partial);
  else ln_rows_bwd_kernel<16><<<nblk, 32 * LB_WARPS, smem, s>>>(H, ldh, dY, ldy, M, n, g, dH, partial);
  SSV_CUDA(cudaGetLastError());
  ++g_launches;
  return kOk;
}

int launch_utt_colsum(const float* X, int ld, int B, int T, int n, float* U, cudaStream_t s) {
  utt_colsum_kernel<<<dim3((n + 127) / 128, B), 128, 0, s>>>(X, ld, T, n, U);
  SSV_CUDA(cudaGetLastError());
  ++g_launches;
  return kOk;
}
int launch_relu_rows(float* x, long n, cudaStream_t s) {
  relu_rows_kernel<<<grid_for(n), 256, 0, s>>>(x, n);
  SSV_CUDA(cudaGetLastError());
  ++g_launches;
  return kOk;
}
int launch_relu_mask(float* dx, const float* x, long n, cudaStream_t s) {
  relu_mask_kernel<<<grid_for(n), 256, 0, s>>>(dx, x, n);
  SSV_CUDA(cudaGetLastError());
  ++g_launches;
  return kOk;
}
int launch_pad_matrix(const float* src, int n_rows, int n_cols, int rows, int ld, float* dst, cudaStream_t s) {
  pad_matrix_kernel<<<grid_for((long)rows * ld), 256, 0, s>>>(src, n_rows, n_cols, rows, ld, dst);
  SSV_CUDA(cudaGetLastError());
  ++g_launches;
  return kOk;
}
int wgrad_plain_chunks(int M) {
  int c = (M + 511) / 512;
  return c < 1 ? 1 : (c > 32 ? 32 : c);
}
int launch_sum_chunks(const float* P, int chunks, long per, float* out, cudaStream_t s) {
  SSV_CHECK(per % 4 == 0, "sum_chunks: length must be a multiple of 4");
  sum_chunks4_kernel<<<grid_for(per / 4), 256, 0, s>>>(reinterpret_cast<const float4*>(P), chunks, per / 4, reinterpret_cast<float4*>(out));
  ++g_launches;
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

int launch_wgrad_plain(const float* dH, int ldh, const float* X, int ldx, int M, int n, int cin, float* P, float* dW, cudaStream_t s) {
  const int chunks = wgrad_plain_chunks(M);
  const int rpc = ((M + chunks - 1) / chunks + WG_ROWS - 1) / WG_ROWS * WG_ROWS;
  wgrad_plain_kernel<<<dim3((n + 63) / 64, (cin + 63) / 64, chunks), 256, 0, s>>>(dH, ldh, X, ldx, M, n, cin, rpc, P);
  SSV_CUDA(cudaGetLastError());
  sum_chunks_kernel<<<grid_for((long)n * cin), 256, 0, s>>>(P, chunks, (long)n * cin, dW);
  SSV_CUDA(cudaGetLastError());
  g_launches += 2;
  return kOk;
}
int launch_att_bwd(const float* Kx, const float* Q, const float* dR, int ldr, const float* dq_add, const float* A, const float* dA, int B,
                   int N, int T, float* dS, float* dq, float* dKx, cudaStream_t s) {
  SSV_CHECK(N >= 1 && N <= AB_NMAX, "attention backward: text length %d outside [1, %d]", N, AB_NMAX);
  att_bwd_q_kernel<<<dim3((T + AB_W - 1) / AB_W, B), AB_W * 32, 0, s>>>(Kx, dR, ldr, dq_add, A, dA, N, T, dS, dq);
  SSV_CUDA(cudaGetLastError());
  att_bwd_kv_kernel<<<dim3((N + AB_W - 1) / AB_W, B), AB_W * 32, 0, s>>>(Q, dR, ldr, A, dS, N, T, dKx);
  SSV_CUDA(cudaGetLastError());
  g_launches += 2;
  return kOk;
}
int launch_add_utt_bias(float* h, int ld, int B, int T, int n, const float* sb, cudaStream_t s) {
  add_utt_bias_kernel<<<grid_for((long)B * T * n), 256, 0, s>>>(h, ld, B, T, n, sb);
  SSV_CUDA(cudaGetLastError());
  ++g_launches;
  return kOk;
}
int embed_bwd_scratch_floats(int vocab, int E) { return EB_CHUNKS * (vocab + 1) * E; }
int launch_embed_bwd(const int64_t* ids, int rows, const float* dX, int ld, int vocab, int E, float* scratch, float* dWt, float* dbias,
                     cudaStream_t s) {
  embed_bwd_kernel<<<dim3((unsigned)(vocab + 1), EB_CHUNKS), 1024, 0, s>>>(ids, rows, dX, ld, vocab, E, scratch);
  SSV_CUDA(cudaGetLastError());
  embed_bwd_sum_kernel<<<grid_for((long)(vocab + 1) * E), 256, 0, s>>>(scratch, vocab, E, dWt, dbias);
  SSV_CUDA(cudaGetLastError());
  g_launches += 2;
  return kOk;
}
int launch_linear_small_bwd(const float* dS, int ds_ld, const float* x, long x_ld, int B, int in_f, int out_f, float* dW, float* db,
                            cudaStream_t s) {
  linear_small_bwd_kernel<<<out_f, 128, 0, s>>>(dS, ds_ld, x, x_ld, B, in_f, out_f, dW, db);
  SSV_CUDA(cudaGetLastError());
  ++g_launches;
  return kOk;
}

}  // namespace ssv
