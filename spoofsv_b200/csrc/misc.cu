// Layout / packing helpers: module-boundary transposes (reference tensors are
// (B, C, T) channels-first, models/TTSModel.py; kernels work on channels-last rows),
// the text-embedding gather (models/TTSModel.py:25-35), weight repacking and the
// hoisted speaker projections (models/TTSModel.py:174,179).
#include "common.cuh"

namespace ssv {

namespace {

// (B, C, T) with arbitrary element strides -> (B, T, ld) channels-last, zero padded to ld.
__global__ void transpose_in_kernel(const float* __restrict__ src, long sb, long sc, long st, int C, int T,
                                    float* __restrict__ dst, int ld) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    tile[i][tx] = (c < C && t < T) ? src[b * sb + c * sc + t * st] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    if (t < T && c < ld) dst[((long)b * T + t) * ld + c] = tile[tx][i];
  }
}

// (B, T, ld) channels-last, channels [c_off, c_off + C) -> (B, C, T) contiguous.
__global__ void transpose_out_kernel(const float* __restrict__ src, int ld, int c_off, int C, int T,
                                     float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    tile[i][tx] = (t < T && c < C) ? src[((long)b * T + t) * ld + c_off + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    if (c < C && t < T) dst[((long)b * C + c) * T + t] = tile[tx][i];
  }
}

__global__ void embed_kernel(const int64_t* __restrict__ ids, int rows, const float* __restrict__ Wt,
                             const float* __restrict__ bias, int vocab, int E, float* __restrict__ dst, int ld,
                             int* err_flag) {
  const int r = blockIdx.x;
  if (r >= rows) return;
  long id = ids[r];
  if (id < 0 || id >= vocab) {
    if (threadIdx.x == 0) atomicExch(err_flag, 1);
    id = 0;
  }
  for (int c = threadIdx.x; c < ld; c += blockDim.x) dst[(long)r * ld + c] = c < E ? Wt[id * E + c] + bias[c] : 0.f;
}

// w [n][cin][k] -> dst [k*cin_p][n_pad] (tap-major rows), zero padded.
__global__ void pack_conv_w_kernel(const float* __restrict__ w, int n, int cin, int k, int cin_p, int n_pad,
                                   float* __restrict__ dst) {
  const long total = (long)k * cin_p * n_pad;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int col = (int)(i % n_pad);
    const int kk = (int)(i / n_pad);
    const int j = kk / cin_p, ci = kk % cin_p;
    dst[i] = (col < n && ci < cin) ? w[((long)col * cin + ci) * k + j] : 0.f;
  }
}

// ConvTranspose1d weight [cin][cout][2] -> [cin][2*cout] with column j*cout + co; bias duplicated.
__global__ void pack_deconv_w_kernel(const float* __restrict__ w, const float* __restrict__ b, int cin, int cout,
                                     float* __restrict__ dst, float* __restrict__ bias_dst) {
  const long total = (long)cin * 2 * cout;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int col = (int)(i % (2 * cout));
    const int ci = (int)(i / (2 * cout));
    const int j = col / cout, co = col % cout;
    dst[i] = w[((long)ci * cout + co) * 2 + j];
    if (ci == 0) bias_dst[col] = b[co];
  }
}

// w [n][cin][k] -> dst [n][k*cin] (per-output-column contiguous K, tap-major) for the decode GEMVs.
__global__ void pack_rowmajor_w_kernel(const float* __restrict__ w, int n, int cin, int k, float* __restrict__ dst) {
  const long total = (long)n * cin * k;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int kk = (int)(i % ((long)cin * k));
    const int col = (int)(i / ((long)cin * k));
    const int j = kk / cin, ci = kk % cin;
    dst[i] = w[((long)col * cin + ci) * k + j];
  }
}

__global__ void pad_vec_kernel(const float* __restrict__ src, int n, int n_pad, float* __restrict__ dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
    dst[i] = i < n ? src[i] : 0.f;
}

// y[b][o] = sum_i w[o][i] x[b][i] + bias[o]; one warp per (b, o).
__global__ void linear_small_kernel(const float* __restrict__ x, long x_ld, const float* __restrict__ w,
                                    const float* __restrict__ bias, int B, int in_f, int out_f,
                                    float* __restrict__ y, int y_ld) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= B * out_f) return;
  const int b = gw / out_f, o = gw % out_f;
  float acc = 0.f;
  for (int i = lane; i < in_f; i += 32) acc = fmaf(w[(long)o * in_f + i], x[b * x_ld + i], acc);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) y[(long)b * y_ld + o] = acc + bias[o];
}

// ---- feature front-end (data/dataset.py:97-118): |STFT| -> mel projection -> normalise -> frame reduction ----------
// Pass 1, one block per 32 spectrogram frames: magnitudes of the block's (F x 32) tile go to shared memory and to the
// workspace, the mel projection (n_mels x F filter bank, L2-resident) is taken from the shared tile, and the maxima
// of both spectrograms are folded into two device words (the values are >= 0, so their bit patterns order like
// unsigned integers).  Pass 2 normalises element-wise and keeps every `reduction`-th mel frame.
constexpr int FEAT_T = 256;
constexpr int FEAT_COLS = 32;
constexpr int FEAT_MAX_MELS = 128;
__global__ void __launch_bounds__(FEAT_T) spec_mag_mel_kernel(const float2* __restrict__ stft, int F, int T,
                                                              const float* __restrict__ melfb, int n_mels,
                                                              float* __restrict__ lin_raw, float* __restrict__ mel_raw,
                                                              unsigned* __restrict__ maxima) {
  extern __shared__ float tile[];                       // [F][32]
  __shared__ float red[2][FEAT_T / 32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int t = blockIdx.x * FEAT_COLS + tx;
  float mx_lin = 0.f, mx_mel = 0.f;
  for (int f = ty; f < F; f += FEAT_T / 32) {
    float m = 0.f;
    if (t < T) {
      const float2 v = stft[(long)f * T + t];
      m = hypotf(v.x, v.y);
      lin_raw[(long)f * T + t] = m;
    }
    tile[f * FEAT_COLS + tx] = m;
    mx_lin = fmaxf(mx_lin, m);
  }
  __syncthreads();
  float acc[FEAT_MAX_MELS / 8];
#pragma unroll
  for (int i = 0; i < FEAT_MAX_MELS / 8; ++i) acc[i] = 0.f;
  for (int f = 0; f < F; ++f) {
    const float sv = tile[f * FEAT_COLS + tx];
#pragma unroll
    for (int i = 0; i < FEAT_MAX_MELS / 8; ++i) {
      const int m = ty + 8 * i;
      if (m < n_mels) acc[i] = fmaf(__ldg(melfb + (long)m * F + f), sv, acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < FEAT_MAX_MELS / 8; ++i) {
    const int m = ty + 8 * i;
    if (m < n_mels && t < T) {
      mel_raw[(long)m * T + t] = acc[i];
      mx_mel = fmaxf(mx_mel, acc[i]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx_lin = fmaxf(mx_lin, __shfl_xor_sync(0xffffffffu, mx_lin, o));
    mx_mel = fmaxf(mx_mel, __shfl_xor_sync(0xffffffffu, mx_mel, o));
  }
  if (tx == 0) { red[0][ty] = mx_lin; red[1][ty] = mx_mel; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float m = 0.f;
    for (int i = 0; i < FEAT_T / 32; ++i) m = fmaxf(m, red[threadIdx.x][i]);
    atomicMax(maxima + threadIdx.x, __float_as_uint(m));
  }
}

__device__ __forceinline__ float feat_norm(float x, float mx, int log_feature, float p, float ref_db, float max_db) {
  if (log_feature) {
    const float db = 20.0f * log10f(fmaxf(1e-5f, x));
    return fminf(fmaxf((db - ref_db + max_db) / max_db, 1e-8f), 1.0f);
  }
  return powf(x / mx, p);
}

__global__ void spec_norm_kernel(const float* __restrict__ lin_raw, const float* __restrict__ mel_raw,
                                 const unsigned* __restrict__ maxima, int F, int T, int n_mels, int reduction, int T4,
                                 int log_feature, float p, float ref_db, float max_db, float* __restrict__ lin_norm,
                                 float* __restrict__ mel_red) {
  const float mx_lin = __uint_as_float(maxima[0]), mx_mel = __uint_as_float(maxima[1]);
  const int Tc = reduction * T4;
  const long n_lin = (long)F * Tc, n_mel = (long)n_mels * T4;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n_lin + n_mel; i += (long)gridDim.x * blockDim.x) {
    if (i < n_lin) {
      const int f = (int)(i / Tc), t = (int)(i % Tc);
      lin_norm[i] = feat_norm(lin_raw[(long)f * T + t], mx_lin, log_feature, p, ref_db, max_db);
    } else {
      const long j = i - n_lin;
      const int m = (int)(j / T4), k = (int)(j % T4);
      mel_red[j] = feat_norm(mel_raw[(long)m * T + (long)reduction * k], mx_mel, log_feature, p, ref_db, max_db);
    }
  }
}

// De-emphasis y[n] = x[n] + c y[n-1] (scipy.signal.lfilter([1], [1, -c]) at generate_test_utterances.py:136) of one
// waveform per block: 1024 threads filter contiguous chunks from zero state, one warp-free serial pass chains the
// 1024 chunk ends (carry_t = end_{t-1} + c^L carry_{t-1}), and a second pass adds c^(i+1) * carry to every sample.
constexpr int DEEMPH_T = 1024;
__global__ void __launch_bounds__(DEEMPH_T) deemphasis_kernel(const float* x, float* y, long n, float c) {
  __shared__ float ends[DEEMPH_T];
  __shared__ float carry[DEEMPH_T];
  const float* xr = x + (long)blockIdx.x * n;
  float* yr = y + (long)blockIdx.x * n;
  const long L = (n + DEEMPH_T - 1) / DEEMPH_T;
  const long lo = (long)threadIdx.x * L, hi = lo + L < n ? lo + L : n;
  float acc = 0.f;
  for (long i = lo; i < hi; ++i) {
    acc = fmaf(c, acc, xr[i]);
    yr[i] = acc;
  }
  ends[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float cL = powf(c, (float)L);
    float cin = 0.f;
    for (int t = 0; t < DEEMPH_T; ++t) {
      carry[t] = cin;
      cin = fmaf(cL, cin, ends[t]);        // a short last chunk only matters for threads past the end
    }
  }
  __syncthreads();
  const float cin = carry[threadIdx.x];
  if (cin != 0.f) {
    float p = c;
    for (long i = lo; i < hi; ++i) {
      yr[i] = fmaf(p, cin, yr[i]);
      p *= c;
    }
  }
}

// Unmasked attention of the train branch (models/TTSModel.py:266-270): A = softmax_n(K^T Q / 16), R = V A, output
// [R ; Q].  One warp per (utterance, time step): lane n-strided dot products against the query held in shared
// memory, warp softmax, then a coalesced weighted sum of the V rows.  Kx: (B, N, 512) channels-last text-encoder
// output (K = channels 0..255, V = 256..511); Q: (B, T, 256); A: (B, N, T) reference layout; RQ: (B, T, 512).
constexpr int ATT_W = 8;            // warps (time steps) per block
constexpr int ATT_NMAX = 192;       // MAX_TEXT_LEN of the reference config is 186
__global__ void __launch_bounds__(ATT_W * 32) train_attention_kernel(const float* __restrict__ Kx, const float* __restrict__ Q,
                                                                     int N, int T, float* __restrict__ A, float* __restrict__ RQ) {
  __shared__ __align__(16) float qs[ATT_W][256];
  __shared__ float ps[ATT_W][ATT_NMAX];
  const int b = blockIdx.y, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * ATT_W + w;
  if (t >= T) return;
  const float* q = Q + ((long)b * T + t) * 256;
  float* rq = RQ + ((long)b * T + t) * 512;
  {
    const float4 a = *reinterpret_cast<const float4*>(q + lane * 8), c = *reinterpret_cast<const float4*>(q + lane * 8 + 4);
    *reinterpret_cast<float4*>(&qs[w][lane * 8]) = a;
    *reinterpret_cast<float4*>(&qs[w][lane * 8 + 4]) = c;
    *reinterpret_cast<float4*>(rq + 256 + lane * 8) = a;
    *reinterpret_cast<float4*>(rq + 256 + lane * 8 + 4) = c;
  }
  __syncwarp();
  float lg[ATT_NMAX / 32];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < ATT_NMAX / 32; ++i) {
    const int n = lane + 32 * i;
    float d = -INFINITY;
    if (n < N) {
      const float4* kr = reinterpret_cast<const float4*>(Kx + ((long)b * N + n) * 512);
      const float4* qv = reinterpret_cast<const float4*>(qs[w]);
      float acc = 0.f;
#pragma unroll 8
      for (int c = 0; c < 64; ++c) {
        const float4 kk = kr[c], qq = qv[c];
        acc = fmaf(kk.x, qq.x, acc); acc = fmaf(kk.y, qq.y, acc); acc = fmaf(kk.z, qq.z, acc); acc = fmaf(kk.w, qq.w, acc);
      }
      d = acc * 0.0625f;                       // 1 / sqrt(256)
    }
    lg[i] = d;
    mx = fmaxf(mx, d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < ATT_NMAX / 32; ++i) {
    lg[i] = lane + 32 * i < N ? expf(lg[i] - mx) : 0.f;
    sum += lg[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
#pragma unroll
  for (int i = 0; i < ATT_NMAX / 32; ++i) {
    const int n = lane + 32 * i;
    if (n < N) {
      const float pv = lg[i] / sum;
      ps[w][n] = pv;
      A[((long)b * N + n) * T + t] = pv;
    }
  }
  __syncwarp();
  float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
  for (int n = 0; n < N; ++n) {
    const float pv = ps[w][n];
    const float* vr = Kx + ((long)b * N + n) * 512 + 256 + lane * 8;
    const float4 v0 = *reinterpret_cast<const float4*>(vr), v1 = *reinterpret_cast<const float4*>(vr + 4);
    r0.x = fmaf(pv, v0.x, r0.x); r0.y = fmaf(pv, v0.y, r0.y); r0.z = fmaf(pv, v0.z, r0.z); r0.w = fmaf(pv, v0.w, r0.w);
    r1.x = fmaf(pv, v1.x, r1.x); r1.y = fmaf(pv, v1.y, r1.y); r1.z = fmaf(pv, v1.z, r1.z); r1.w = fmaf(pv, v1.w, r1.w);
  }
  *reinterpret_cast<float4*>(rq + lane * 8) = r0;
  *reinterpret_cast<float4*>(rq + lane * 8 + 4) = r1;
}

inline int grid_for(long total, int block = 256) {
  long g = (total + block - 1) / block;
  return (int)(g > 4096 ? 4096 : (g < 1 ? 1 : g));
}

}  // namespace

int launch_transpose_in(const float* src, long sb, long sc, long st, int B, int C, int T, float* dst, int ld,
                        cudaStream_t s) {
  dim3 grid((T + 31) / 32, (ld + 31) / 32, B), block(32, 8);
  transpose_in_kernel<<<grid, block, 0, s>>>(src, sb, sc, st, C, T, dst, ld);
  ++g_launches;
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

int launch_transpose_out2(const float* src, int ld, int c_off, int B, int C, int T, float* dst, cudaStream_t s) {
  dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
  transpose_out_kernel<<<grid, block, 0, s>>>(src, ld, c_off, C, T, dst);
  ++g_launches;
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

int launch_transpose_out(const float* src, int ld, int B, int C, int T, float* dst, cudaStream_t s) {
  return launch_transpose_out2(src, ld, 0, B, C, T, dst, s);
}

int launch_embed(const int64_t* ids, int B, int N, const float* Wt, const float* bias, int vocab, int E, float* dst,
                 int ld, int* err_flag, cudaStream_t s) {
  embed_kernel<<<B * N, 128, 0, s>>>(ids, B * N, Wt, bias, vocab, E, dst, ld, err_flag);
  ++g_launches;
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

int launch_pack_conv_w(const float* w, int n, int cin, int k, int cin_p, int n_pad, float* dst, cudaStream_t s) {
  pack_conv_w_kernel<<<grid_for((long)k * cin_p * n_pad), 256, 0, s>>>(w, n, cin, k, cin_p, n_pad, dst);
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

int launch_pack_deconv_w(const float* w, const float* b, int cin, int cout, float* dst, float* bias_dst,
                         cudaStream_t s) {
  pack_deconv_w_kernel<<<grid_for((long)cin * 2 * cout), 256, 0, s>>>(w, b, cin, cout, dst, bias_dst);
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

int launch_pack_rowmajor_w(const float* w, int n, int cin, int k, float* dst, cudaStream_t s) {
  pack_rowmajor_w_kernel<<<grid_for((long)n * cin * k), 256, 0, s>>>(w, n, cin, k, dst);
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

int launch_pad_vec(const float* src, int n, int n_pad, float* dst, cudaStream_t s) {
  pad_vec_kernel<<<grid_for(n_pad), 256, 0, s>>>(src, n, n_pad, dst);
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

int launch_linear_small(const float* x, long x_ld, const float* w, const float* b, int B, int in_f, int out_f,
                        float* y, int y_ld, cudaStream_t s) {
  const long threads = (long)B * out_f * 32;
  linear_small_kernel<<<(int)((threads + 255) / 256), 256, 0, s>>>(x, x_ld, w, b, B, in_f, out_f, y, y_ld);
  ++g_launches;
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

int launch_spec_features(const float* stft_ri, int F, int T, const float* melfb, int n_mels, int log_feature,
                         float norm_power, float ref_db, float max_db, int reduction, float* lin_norm,
                         float* mel_red, float* workspace, cudaStream_t s) {
  float* lin_raw = workspace;
  float* mel_raw = workspace + (size_t)F * T;
  unsigned* maxima = reinterpret_cast<unsigned*>(mel_raw + (size_t)n_mels * T);
  SSV_CUDA(cudaMemsetAsync(maxima, 0, 2 * sizeof(unsigned), s));
  const size_t smem = (size_t)F * FEAT_COLS * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    SSV_CUDA(cudaFuncSetAttribute(spec_mag_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  spec_mag_mel_kernel<<<(T + FEAT_COLS - 1) / FEAT_COLS, FEAT_T, smem, s>>>(reinterpret_cast<const float2*>(stft_ri), F, T,
                                                                            melfb, n_mels, lin_raw, mel_raw, maxima);
  SSV_CUDA(cudaGetLastError());
  const int T4 = T / reduction;
  const long total = (long)F * reduction * T4 + (long)n_mels * T4;
  if (total > 0) {
    long g = (total + 255) / 256;
    if (g > 4096) g = 4096;
    spec_norm_kernel<<<(int)g, 256, 0, s>>>(lin_raw, mel_raw, maxima, F, T, n_mels, reduction, T4, log_feature, norm_power,
                                            ref_db, max_db, lin_norm, mel_red);
    SSV_CUDA(cudaGetLastError());
  }
  g_launches += 2;
  return kOk;
}

int launch_deemphasis(const float* x, float* y, int B, long n, float coeff, cudaStream_t s) {
  deemphasis_kernel<<<B, DEEMPH_T, 0, s>>>(x, y, n, coeff);
  ++g_launches;
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

int launch_train_attention(const float* Kx, const float* Q, int B, int N, int T, float* A, float* RQ, cudaStream_t s) {
  SSV_CHECK(N >= 1 && N <= ATT_NMAX, "train attention: text length %d outside [1, %d]", N, ATT_NMAX);
  dim3 grid((T + ATT_W - 1) / ATT_W, B);
  train_attention_kernel<<<grid, ATT_W * 32, 0, s>>>(Kx, Q, N, T, A, RQ);
  ++g_launches;
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace ssv
