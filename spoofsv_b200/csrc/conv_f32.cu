// FP32 (CUDA-core FFMA) fused conv1d + LayerNorm / highway-gate kernel.
//
// This is the FP32-exact arm of the hot path (north_star: 1e-4 max-abs vs the
// reference): every highwayConv (models/TTSModel.py:63-84), every 1x1 conv +
// LayerNorm (+ReLU / +sigmoid) and the ConvTranspose1d (as a 1x1 GEMM whose 2*C
// output columns are the two interleaved output frames) of TextEnc and SSRN run
// through this one kernel.  Implicit GEMM: M = (b,t) rows, K = taps*Cin, N = Cout;
// a CTA owns BM complete rows so the channel LayerNorm and the gate are done in
// registers before anything is written.
//
// Tiling: 256 threads = 64 column-threads x 4 row-groups; thread tile RPT x CPT.
// Wide layers (CPT % 8 == 0) give a thread 4 consecutive columns per 256-column
// block (4 tx + 256 j'), so the weight fragment is read with LDS.128 (5.75 shared
// loads per 112 FMAs instead of 17.75); narrow ones keep columns tx + 64 j.  Either
// way H1[c] and H2[c] of a highway layer live in the same thread.  K is streamed in chunks of 16 through a cp.async ring (3 stages):
// the weight chunk is one contiguous KC*N block (one bulk copy, cp.async.bulk ->
// mbarrier complete_tx), the activation chunk is a row gather whose tap shift /
// zero padding is resolved per row (zero-fill cp.async arriving on the same mbarrier).
#include "common.cuh"

namespace ssv {

namespace {

constexpr int KC = 16;
constexpr int NT = 256;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// Sum v[0..CNT) over the 64 column-threads that share a row group (two warps).
template <int CNT>
__device__ __forceinline__ void rowgroup_sum(float (&v)[CNT], float* red /*[8][CNT]*/, int tid) {
  const int warp = tid >> 5, lane = tid & 31, ty = tid >> 6;
#pragma unroll
  for (int i = 0; i < CNT; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < CNT; ++i) red[warp * CNT + i] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < CNT; ++i) v[i] = red[(2 * ty) * CNT + i] + red[(2 * ty + 1) * CNT + i];
  __syncthreads();
}

// column of accumulator j of column-thread tx
template <int CPT>
__device__ __forceinline__ int col_of(int tx, int j) {
  return CPT % 8 == 0 ? 4 * tx + (j & 3) + 256 * (j >> 2) : tx + 64 * j;
}

template <int BM, int CPT, int STAGES>
__global__ void __launch_bounds__(NT) conv_f32_kernel(const ConvArgs a) {
  constexpr int NP = 64 * CPT;
  constexpr int RPT = BM / 4;
  constexpr bool VEC = CPT % 8 == 0;
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;                        // [STAGES][KC][NP]
  float* As = smem + STAGES * KC * NP;     // [STAGES][BM][KC]
  __shared__ float red[8 * 2 * RPT];
  __shared__ int row_b[BM], row_t[BM];

  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
  const int row0 = blockIdx.x * BM;
  if (tid < BM) {
    int r = row0 + tid;
    if (r < a.M) {
      row_b[tid] = r / a.t_rows;
      row_t[tid] = a.t0 + r % a.t_rows;
    } else {
      row_b[tid] = -1;
      row_t[tid] = 0;
    }
  }
  __syncthreads();

  const int chunks_per_tap = a.cin_p / KC;
  const int nchunks = a.ktaps * chunks_per_tap;
  const int tap_base = a.causal == 1 ? -(a.ktaps - 1) : a.causal == 2 ? 0 : -((a.ktaps - 1) / 2);   // 2: anti-causal (dgrad of a causal conv)

  // One mbarrier per ring stage collects both halves of a chunk: the weight block (KC x NP floats, contiguous) comes
  // as one bulk copy issued by thread 0 (complete_tx), the activation rows as per-thread zero-fill cp.async whose
  // completion arrives on the same barrier -- no per-thread address arithmetic for 16 weight copies, no wait_group.
  __shared__ __align__(8) unsigned long long full[STAGES];
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      const unsigned bar = static_cast<unsigned>(__cvta_generic_to_shared(&full[s]));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1 + BM * 4));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // loader state of this thread's activation segment: (tap, channel offset) of the next chunk to fetch
  int ld_j = 0, ld_c0 = 0;
  const int ld_r = tid >> 2, ld_seg = tid & 3;
  int ld_b = -1, ld_t = 0;
  if (tid < BM * 4) { ld_b = row_b[ld_r]; ld_t = row_t[ld_r]; }
  const float* ld_row = a.X + (long)(ld_b < 0 ? 0 : ld_b) * a.x_sb + ld_seg * 4;

  auto load_stage = [&](int q, int st) {
    const unsigned bar = static_cast<unsigned>(__cvta_generic_to_shared(&full[st]));
    if (tid == 0) {
      constexpr unsigned bytes = KC * NP * sizeof(float);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(Ws + st * KC * NP));
      const float* src = a.W + (size_t)q * KC * NP;
      constexpr unsigned piece = bytes / 4;        // 4 bulk copies of 2-16 KB
#pragma unroll
      for (int i = 0; i < 4; ++i)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst + i * piece), "l"(reinterpret_cast<const char*>(src) + i * piece), "r"(piece), "r"(bar) : "memory");
    }
    if (tid < BM * 4) {
      const int t = ld_t + (tap_base + ld_j) * a.dil;
      const bool ok = (ld_b >= 0) && (t >= 0) && (t < a.t_in);
      const float* p = ok ? ld_row + (long)t * a.x_st + ld_c0 : a.X;
      cp_async16(As + (st * BM + ld_r) * KC + ld_seg * 4, p, ok);
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
      ld_c0 += KC;
      if (ld_c0 == a.cin_p) { ld_c0 = 0; ++ld_j; }
    }
  };
  auto wait_stage = [&](int st, unsigned parity) {
    const unsigned bar = static_cast<unsigned>(__cvta_generic_to_shared(&full[st]));
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
          "selp.u32 %0, 1, 0, p;\n"
          "}\n"
          : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
  };

  float acc[RPT][CPT];
#pragma unroll
  for (int i = 0; i < RPT; ++i)
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[i][j] = 0.f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s)
    if (s < nchunks) load_stage(s, s);

  for (int q = 0; q < nchunks; ++q) {
    wait_stage(q % STAGES, (unsigned)(q / STAGES) & 1u);
    __syncthreads();                       // every warp is done with chunk q - 1: its stage may be refilled
    {
      const int qn = q + STAGES - 1;
      if (qn < nchunks) load_stage(qn, qn % STAGES);
    }
    const int st = q % STAGES;
    const float* Wst = Ws + st * KC * NP + (VEC ? 4 * tx : tx);
    const float* Ast = As + (st * BM + ty * RPT) * KC;
    // two of the four k-quads per trip: the fully unrolled 16-k body (38 KB of SASS) missed the instruction cache
    // (ncu: stall_no_instruction 0.24 per issue); measured TextEnc 4.78 -> 4.34 ms, SSRN fp32 19.8 -> 17.9 ms
#pragma unroll 2
    for (int k4 = 0; k4 < KC / 4; ++k4) {
      float4 av[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) av[i] = *reinterpret_cast<const float4*>(Ast + i * KC + k4 * 4);
#pragma unroll
      for (int kq = 0; kq < 4; ++kq) {
        float w[CPT];
        if constexpr (VEC) {
#pragma unroll
          for (int j = 0; j < CPT / 4; ++j) {
            const float4 wv = *reinterpret_cast<const float4*>(Wst + (k4 * 4 + kq) * NP + 256 * j);
            w[4 * j] = wv.x; w[4 * j + 1] = wv.y; w[4 * j + 2] = wv.z; w[4 * j + 3] = wv.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < CPT; ++j) w[j] = Wst[(k4 * 4 + kq) * NP + 64 * j];
        }
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const float x = kq == 0 ? av[i].x : kq == 1 ? av[i].y : kq == 2 ? av[i].z : av[i].w;
#pragma unroll
          for (int j = 0; j < CPT; ++j) acc[i][j] = fmaf(x, w[j], acc[i][j]);
        }
      }
    }
  }

  // ---------------- epilogue ----------------
#pragma unroll
  for (int j = 0; j < CPT; ++j) {
    const float bv = a.bias[col_of<CPT>(tx, j)];
#pragma unroll
    for (int i = 0; i < RPT; ++i) acc[i][j] += bv;
  }
  if (a.bias_b != nullptr) {
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int b = row_b[ty * RPT + i];
      if (b >= 0) {
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc[i][j] += a.bias_b[(long)b * a.bias_b_ld + col_of<CPT>(tx, j)];
      }
    }
  }

  if (a.epi == EPI_NONE) {
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int b = row_b[ty * RPT + i];
      if (b < 0) continue;
      float* y = a.Y + (long)b * a.y_sb + (long)row_t[ty * RPT + i] * a.y_st;
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int c = col_of<CPT>(tx, j);
        if (c < a.y_cols) y[c] = c < a.n ? acc[i][j] : 0.f;
      }
    }
    return;
  }

  if (a.epi == EPI_HIGHWAY) {
    if constexpr (CPT % 2 == 0) {
      constexpr int H = CPT / 2;
      const float inv_d = 1.0f / (float)(a.n / 2);
      float s[2 * RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < H; ++j) { s1 += acc[i][j]; s2 += acc[i][j + H]; }
        s[2 * i] = s1; s[2 * i + 1] = s2;
      }
      rowgroup_sum<2 * RPT>(s, red, tid);
      float mean[2 * RPT];
#pragma unroll
      for (int i = 0; i < 2 * RPT; ++i) mean[i] = s[i] * inv_d;
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        float q1 = 0.f, q2 = 0.f;
#pragma unroll
        for (int j = 0; j < H; ++j) {
          const float d1 = acc[i][j] - mean[2 * i], d2 = acc[i][j + H] - mean[2 * i + 1];
          q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2);
        }
        s[2 * i] = q1; s[2 * i + 1] = q2;
      }
      rowgroup_sum<2 * RPT>(s, red, tid);
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int b = row_b[ty * RPT + i];
        if (b < 0) continue;
        const int t = row_t[ty * RPT + i];
        const float r1 = 1.0f / sqrtf(s[2 * i] * inv_d + 1e-5f);
        const float r2 = 1.0f / sqrtf(s[2 * i + 1] * inv_d + 1e-5f);
        const float* xr = a.X + (long)b * a.x_sb + (long)t * a.x_st;
        float* y = a.Y + (long)b * a.y_sb + (long)t * a.y_st;
#pragma unroll
        for (int j = 0; j < H; ++j) {
          const int c = col_of<CPT>(tx, j);
          const float h1 = (acc[i][j] - mean[2 * i]) * r1 * a.g1[c] + a.b1[c];
          const float h2 = (acc[i][j + H] - mean[2 * i + 1]) * r2 * a.g2[c] + a.b2[c];
          const float g = sigmoidf_(h1);
          y[c] = g * h2 + (1.0f - g) * xr[c];
        }
      }
    }
    return;
  }

  // EPI_LN / EPI_LN_RELU / EPI_LN_SIGMOID over the first a.n columns
  {
    const float inv_n = 1.0f / (float)a.n;
    float s[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < CPT; ++j) v += (col_of<CPT>(tx, j) < a.n) ? acc[i][j] : 0.f;
      s[i] = v;
    }
    rowgroup_sum<RPT>(s, red, tid);
    float mean[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) mean[i] = s[i] * inv_n;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const float d = acc[i][j] - mean[i];
        v += (col_of<CPT>(tx, j) < a.n) ? d * d : 0.f;
      }
      s[i] = v;
    }
    rowgroup_sum<RPT>(s, red, tid);
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int b = row_b[ty * RPT + i];
      if (b < 0) continue;
      const float rstd = 1.0f / sqrtf(s[i] * inv_n + 1e-5f);
      float* y = a.Y + (long)b * a.y_sb + (long)row_t[ty * RPT + i] * a.y_st;
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int c = col_of<CPT>(tx, j);
        if (c >= a.y_cols) continue;
        float o = 0.f;
        if (c < a.n) {
          o = (acc[i][j] - mean[i]) * rstd * a.g1[c] + a.b1[c];
          if (a.epi == EPI_LN_RELU) o = fmaxf(o, 0.f);
          else if (a.epi == EPI_LN_SIGMOID) o = sigmoidf_(o);
        }
        y[c] = o;
      }
    }
  }
}

// ---- few rows (batch-1 TextEnc: 58 rows): split K and N over the whole GPU ----------------------------------------
// With one CTA per row tile a 58-row problem runs on 4 SMs and every one of them streams the whole weight matrix
// (TextEnc at B = 1: 3.0 ms, 25 % of a batch-1 synthesis).  Here the weight matrix is cut into (64-column, K / ks)
// slabs, one per CTA (~128 CTAs); each CTA multiplies its slab with all rows and writes a partial product, and a
// second kernel (one warp per row) adds the ks partials in a fixed order and applies bias / LayerNorm / gate.
constexpr int SPL_MAXM = 128;
constexpr int SPL_T = 256;

__global__ void __launch_bounds__(SPL_T) conv_f32_split_kernel(const ConvArgs a, int klen, int n_pad, float* __restrict__ part) {
  extern __shared__ __align__(16) float xs[];                 // [klen][Mp] rows contiguous, Mp = M rounded up to 16 / 32 / 64 / 128
  const int Mp = a.M <= 16 ? 16 : a.M <= 32 ? 32 : a.M <= 64 ? 64 : 128;
  const int col0 = blockIdx.x * 64, ks = blockIdx.y;
  const int K = a.ktaps * a.cin_p;
  const int k0 = ks * klen, k1 = min(K, k0 + klen);
  const int tap_base = a.causal == 1 ? -(a.ktaps - 1) : a.causal == 2 ? 0 : -((a.ktaps - 1) / 2);   // 2: anti-causal (dgrad of a causal conv)
  // stage my K range of every row: thread (r, q) copies the float4s q, q + T/Mp, ... of row r (channels are contiguous
  // in X; k0 and cin_p are multiples of 16, so a float4 never straddles a tap), transposed into xs[k][r]
  {
    const int r = threadIdx.x % Mp, q0 = threadIdx.x / Mp, qs = SPL_T / Mp;       // Mp in {16, 32, 64, 128} divides 256
    const int b = r < a.M ? r / a.t_rows : 0, tr = r < a.M ? a.t0 + r % a.t_rows : 0;
    for (int k4 = q0; k4 < (k1 - k0) / 4; k4 += qs) {
      const int k = k0 + 4 * k4, j = k / a.cin_p, c = k - j * a.cin_p;
      const int t = tr + (tap_base + j) * a.dil;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < a.M && t >= 0 && t < a.t_in) v = *reinterpret_cast<const float4*>(a.X + (long)b * a.x_sb + (long)t * a.x_st + c);
      float* d = xs + (4 * k4) * Mp + r;
      d[0] = v.x; d[Mp] = v.y; d[2 * Mp] = v.z; d[3 * Mp] = v.w;
    }
  }
  __syncthreads();
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;     // column, row quarter
  const int rq = Mp / 4;                                      // rows per thread (multiple of 4, <= 32)
  float acc[SPL_MAXM / 4];
#pragma unroll
  for (int i = 0; i < SPL_MAXM / 4; ++i) acc[i] = 0.f;
  const float* wp = a.W + (size_t)k0 * n_pad + col0 + tx;
  // weights in batches of 16 rows: the 16 loads are in flight together (klen is a multiple of 16; rows past K read
  // as zero), otherwise every k step would wait a full L2 round trip
  for (int kb = 0; kb < klen; kb += 16) {
    float w[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) w[u] = k0 + kb + u < K ? __ldg(wp + (size_t)(kb + u) * n_pad) : 0.f;
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (k0 + kb + u >= k1) break;
      const float4* xr = reinterpret_cast<const float4*>(xs + (kb + u) * Mp + ty * rq);
#pragma unroll
      for (int i = 0; i < SPL_MAXM / 16; ++i) {
        if (4 * i < rq) {
          const float4 x = xr[i];
          acc[4 * i] = fmaf(x.x, w[u], acc[4 * i]);
          acc[4 * i + 1] = fmaf(x.y, w[u], acc[4 * i + 1]);
          acc[4 * i + 2] = fmaf(x.z, w[u], acc[4 * i + 2]);
          acc[4 * i + 3] = fmaf(x.w, w[u], acc[4 * i + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < SPL_MAXM / 4; ++i) {
    const int r = ty * rq + i;
    if (i < rq && r < a.M) part[((size_t)ks * a.M + r) * n_pad + col0 + tx] = acc[i];
  }
}

// One 256-thread block per row: sum the ks partials (fixed order), bias, then the epilogue of the fused kernels.
// Thread tid owns columns tid + 256 i, so H1[c] and H2[c] of a highway layer (d = 256 or 512) sit in one thread.
template <int CNT>
__device__ __forceinline__ void block_sum(float (&v)[CNT], float* red /*[8][CNT]*/) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < CNT; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < CNT; ++i) red[warp * CNT + i] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < CNT; ++i) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w * CNT + i];
    v[i] = t;
  }
}

template <int NPT>      // columns per thread = n_pad / 256
__global__ void __launch_bounds__(256) conv_f32_split_epilogue(const ConvArgs a, int nks, const float* __restrict__ part) {
  constexpr int n_pad = 256 * NPT;
  __shared__ float red[8 * 2];
  const int r = blockIdx.x, tid = threadIdx.x;
  const int b = r / a.t_rows, t = a.t0 + r % a.t_rows;
  float* y = a.Y + (long)b * a.y_sb + (long)t * a.y_st;
  float v[NPT];
#pragma unroll
  for (int i = 0; i < NPT; ++i) v[i] = 0.f;
  for (int ks = 0; ks < nks; ++ks) {                          // fixed order; the loads of a step are independent
    const float* pr = part + ((size_t)ks * a.M + r) * n_pad + tid;
#pragma unroll
    for (int i = 0; i < NPT; ++i) v[i] += pr[256 * i];
  }
#pragma unroll
  for (int i = 0; i < NPT; ++i) {
    const int c = tid + 256 * i;
    v[i] += a.bias[c];
    if (a.bias_b != nullptr && c < a.n) v[i] += a.bias_b[(long)b * a.bias_b_ld + c];
  }
  if (a.epi == EPI_NONE) {
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      const int c = tid + 256 * i;
      if (c < a.y_cols) y[c] = c < a.n ? v[i] : 0.f;
    }
    return;
  }
  if (a.epi == EPI_HIGHWAY) {
    if constexpr (NPT >= 2) {
      constexpr int H = NPT / 2;
      const int d = a.n / 2;
      float s[2] = {0.f, 0.f};
#pragma unroll
      for (int i = 0; i < H; ++i) { s[0] += v[i]; s[1] += v[i + H]; }
      block_sum<2>(s, red);
      const float m1 = s[0] / (float)d, m2 = s[1] / (float)d;
      float q[2] = {0.f, 0.f};
#pragma unroll
      for (int i = 0; i < H; ++i) {
        const float d1 = v[i] - m1, d2 = v[i + H] - m2;
        q[0] = fmaf(d1, d1, q[0]); q[1] = fmaf(d2, d2, q[1]);
      }
      block_sum<2>(q, red);
      const float r1 = 1.0f / sqrtf(q[0] / (float)d + 1e-5f), r2 = 1.0f / sqrtf(q[1] / (float)d + 1e-5f);
      const float* xr = a.X + (long)b * a.x_sb + (long)t * a.x_st;
#pragma unroll
      for (int i = 0; i < H; ++i) {
        const int c = tid + 256 * i;
        const float h1 = (v[i] - m1) * r1 * a.g1[c] + a.b1[c];
        const float h2 = (v[i + H] - m2) * r2 * a.g2[c] + a.b2[c];
        const float g = sigmoidf_(h1);
        y[c] = g * h2 + (1.0f - g) * xr[c];
      }
    }
    return;
  }
  float s1[1] = {0.f};
#pragma unroll
  for (int i = 0; i < NPT; ++i)
    if (tid + 256 * i < a.n) s1[0] += v[i];
  block_sum<1>(s1, red);
  const float mean = s1[0] / (float)a.n;
  float q1[1] = {0.f};
#pragma unroll
  for (int i = 0; i < NPT; ++i)
    if (tid + 256 * i < a.n) { const float dd = v[i] - mean; q1[0] = fmaf(dd, dd, q1[0]); }
  block_sum<1>(q1, red);
  const float rstd = 1.0f / sqrtf(q1[0] / (float)a.n + 1e-5f);
#pragma unroll
  for (int i = 0; i < NPT; ++i) {
    const int c = tid + 256 * i;
    if (c >= a.y_cols) continue;
    float o = 0.f;
    if (c < a.n) {
      o = (v[i] - mean) * rstd * a.g1[c] + a.b1[c];
      if (a.epi == EPI_LN_RELU) o = fmaxf(o, 0.f);
      else if (a.epi == EPI_LN_SIGMOID) o = sigmoidf_(o);
    }
    y[c] = o;
  }
}

int launch_split(const ConvArgs& a, int n_pad, cudaStream_t s) {
  static float* part = nullptr;
  static size_t part_floats = 0;
  const int K = a.ktaps * a.cin_p;
  const int slabs = n_pad / 64;
  int nks = 128 / slabs;
  if (nks < 1) nks = 1;
  int klen = (K + nks - 1) / nks;
  klen = (klen + 15) / 16 * 16;
  nks = (K + klen - 1) / klen;
  const size_t need = (size_t)nks * a.M * n_pad;
  if (need > part_floats) {
    if (part) cudaFree(part);
    part = nullptr;
    if (cudaMalloc((void**)&part, need * sizeof(float)) != cudaSuccess) {
      cudaGetLastError();
      part_floats = 0;
      set_error("conv_f32: cannot allocate %zu bytes of split-K partials", need * sizeof(float));
      return kNoMem;
    }
    part_floats = need;
  }
  const int Mp = a.M <= 16 ? 16 : a.M <= 32 ? 32 : a.M <= 64 ? 64 : 128;
  const size_t smem = (size_t)klen * Mp * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    SSV_CUDA(cudaFuncSetAttribute(conv_f32_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
    configured = 200 * 1024;
  }
  SSV_CHECK(smem <= 200 * 1024, "conv_f32: split tile does not fit shared memory");
  conv_f32_split_kernel<<<dim3(slabs, nks), SPL_T, smem, s>>>(a, klen, n_pad, part);
  SSV_CUDA(cudaGetLastError());
  if (n_pad == 256) conv_f32_split_epilogue<1><<<a.M, 256, 0, s>>>(a, nks, part);
  else if (n_pad == 512) conv_f32_split_epilogue<2><<<a.M, 256, 0, s>>>(a, nks, part);
  else conv_f32_split_epilogue<4><<<a.M, 256, 0, s>>>(a, nks, part);
  SSV_CUDA(cudaGetLastError());
  g_launches += 2;
  return kOk;
}

template <int BM, int CPT, int STAGES>
int launch_inst(const ConvArgs& a, cudaStream_t s) {
  constexpr int NP = 64 * CPT;
  constexpr size_t smem = (size_t)STAGES * (KC * NP + BM * KC) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    SSV_CUDA(cudaFuncSetAttribute(conv_f32_kernel<BM, CPT, STAGES>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int grid = (a.M + BM - 1) / BM;
  conv_f32_kernel<BM, CPT, STAGES><<<grid, NT, smem, s>>>(a);
  ++g_launches;
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace

int launch_conv_f32(const ConvArgs& a, cudaStream_t s) {
  SSV_CHECK(a.M > 0 && a.cin_p % KC == 0 && a.ktaps >= 1 && a.ktaps <= 3, "conv_f32: bad shape");
  const int n_pad = round_up(a.n, 64);
  if (a.epi == EPI_HIGHWAY) SSV_CHECK(a.n == n_pad && (a.n / 64) % 2 == 0, "conv_f32: highway needs n %% 128 == 0");
  // few rows: cut the weight matrix over the whole GPU instead of streaming all of it through a handful of CTAs
  if (a.M <= SPL_MAXM && (n_pad == 256 || n_pad == 512 || n_pad == 1024) && (a.epi != EPI_HIGHWAY || (a.n == n_pad && n_pad >= 512)))
    return launch_split(a, n_pad, s);
  switch (n_pad / 64) {
    case 2:  return launch_inst<32, 2, 3>(a, s);
    case 4:  return launch_inst<32, 4, 3>(a, s);
    case 8:  return launch_inst<32, 8, 3>(a, s);
    case 9:  return launch_inst<32, 9, 3>(a, s);
    case 16:
      // d = 512 layers stream a 6.3 MB weight matrix per CTA from L2: 32-row tiles halve that traffic and make the
      // inner loop FMA-bound (16 weight loads per 128 FMAs); small problems keep 16-row tiles for more CTAs
      if (a.M < 32 * 96) return launch_inst<16, 16, 3>(a, s);
      {
        // one CTA per SM: pick the tile height that wastes the least of the last wave (TextEnc at B = 64 is
        // 3712 rows: 116 CTAs of 32 rows leave 32 of 148 SMs idle, 133 CTAs of 28 rows do not)
        int sms = 148, dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        auto cost = [&](int bm) { const long ctas = (a.M + bm - 1) / bm; return ((ctas + sms - 1) / sms) * bm; };
        return cost(28) < cost(32) ? launch_inst<28, 16, 3>(a, s) : launch_inst<32, 16, 3>(a, s);
      }
    default: break;
  }
  set_error("conv_f32: unsupported output width %d (padded %d)", a.n, n_pad);
  return kInval;
}

}  // namespace ssv
