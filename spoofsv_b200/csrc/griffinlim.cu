// Fused fast Griffin-Lim for the waveform stage (generate_test_utterances.py:133, synthesize.py:141:
// librosa.core.griffinlim(S, n_iter=64, hop_length=256, win_length=1024) of librosa 0.7, momentum 0.99).
//
//   for it in range(n_iter):  tprev = rebuilt;  rebuilt = stft(istft(S * angles));
//                             angles = rebuilt - momentum / (1 + momentum) * tprev;  angles /= |angles| + 1e-16
//   return istft(S * angles)
//
// As a chain of library calls every iteration walks the (B, 513, T) complex spectrogram about a dozen times
// (complex multiply, irfft, window, overlap-add, envelope division, padding, framing, rfft, momentum update,
// abs, divide): 2.4 ms per iteration for 64 utterances x 1304 frames.  Here an iteration is two kernels:
//
//   gl_istft_ola_kernel  a CTA owns 24 consecutive output hops of one utterance; the frames touching them go through
//                        the FFT as PAIRS (Z = A + iB: two Hermitian spectra packed into one complex signal, so one
//                        1024-point complex FFT inverts both; shared-memory Stockham radix 4), and their windowed
//                        segments are overlap-added in shared memory in frame order; the window sum-square division
//                        is applied on the way out.  Three halo frames per chunk are recomputed instead of exchanged.
//   gl_stft_kernel       frame pair a + ib gathered from the signal with reflect padding, window, FFT, unpack the two
//                        spectra, momentum update and phase normalisation fused into the epilogue.
//
// Spectra live in a time-major layout (B, T, 513) inside the loop (a frame's bins are contiguous), S and the initial
// phases are transposed once.  FFT size and hop are the reference's (n_fft = win_length = 1024, hop = 256).
#include "common.cuh"

namespace ssv {

namespace {

constexpr int GL_N = 1024;          // n_fft = win_length
constexpr int GL_HOP = 256;
constexpr int GL_F = GL_N / 2 + 1;  // 513 bins
constexpr int GL_T = 256;           // threads: one radix-4 butterfly each per pass

// Shared-memory FFT buffers are padded by one element per 16 (index i lives at i + i / 16): the radix-4 stores of the
// first two passes (stride 4 and 16 elements between neighbouring lanes) would otherwise be 4-way bank conflicts.
constexpr int GL_PAD = GL_N + GL_N / 16;
__device__ __forceinline__ int pidx(int i) { return i + (i >> 4); }

struct GlSmem {
  float2 a[GL_PAD];
  float2 b[GL_PAD];
  float2 tw[4 + 16 + 64 + 256];     // per pass p = 1..4 (Ns = 4^p): e^{-2 pi i k / (4 Ns)}, k < Ns, stored contiguously
  float win[GL_N];                  // periodic Hann
};

__device__ __forceinline__ float2 cmul(float2 x, float2 y) { return make_float2(x.x * y.x - x.y * y.y, x.x * y.y + x.y * y.x); }

__device__ __forceinline__ void gl_tables(GlSmem& sm, int tid) {
  for (int k = tid; k < GL_N; k += GL_T) sm.win[k] = 0.5f - 0.5f * cospif(2.0f * (float)k / (float)GL_N);
  for (int i = tid; i < 4 + 16 + 64 + 256; i += GL_T) {
    const int Ns = i < 4 ? 4 : i < 20 ? 16 : i < 84 ? 64 : 256;
    const int k = i - (Ns == 4 ? 0 : Ns == 16 ? 4 : Ns == 64 ? 20 : 84);
    float s, c;
    sincospif(-2.0f * (float)k / (float)(4 * Ns), &s, &c);
    sm.tw[i] = make_float2(c, s);
  }
}

// 1024-point complex FFT, Stockham autosort radix 4, five passes; input in sm.a, result in sm.b (padded indexing).
// INV: conjugate twiddles and butterfly (no 1/N scaling here).
template <bool INV>
__device__ __forceinline__ void fft1024(GlSmem& sm, int tid) {
  float2* in = sm.a;
  float2* out = sm.b;
#pragma unroll
  for (int pass = 0; pass < 5; ++pass) {
    const int Ns = 1 << (2 * pass);
    const int k = tid & (Ns - 1);
    float2 v0 = in[pidx(tid)], v1 = in[pidx(tid + 256)], v2 = in[pidx(tid + 512)], v3 = in[pidx(tid + 768)];
    if (pass > 0) {
      const int base = pass == 1 ? 0 : pass == 2 ? 4 : pass == 3 ? 20 : 84;
      float2 w1 = sm.tw[base + k];
      if (INV) w1.y = -w1.y;
      const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
      v1 = cmul(v1, w1); v2 = cmul(v2, w2); v3 = cmul(v3, w3);
    }
    const float2 t0 = make_float2(v0.x + v2.x, v0.y + v2.y), t1 = make_float2(v0.x - v2.x, v0.y - v2.y);
    const float2 t2 = make_float2(v1.x + v3.x, v1.y + v3.y);
    const float2 d = make_float2(v1.x - v3.x, v1.y - v3.y);
    const float2 t3 = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);      // +-i (v1 - v3)
    const int j0 = ((tid - k) << 2) + k;
    out[pidx(j0)] = make_float2(t0.x + t2.x, t0.y + t2.y);
    out[pidx(j0 + Ns)] = make_float2(t1.x + t3.x, t1.y + t3.y);
    out[pidx(j0 + 2 * Ns)] = make_float2(t0.x - t2.x, t0.y - t2.y);
    out[pidx(j0 + 3 * Ns)] = make_float2(t1.x - t3.x, t1.y - t3.y);
    __syncthreads();
    float2* tmp = in; in = out; out = tmp;
  }
  // five passes: the result sits in the buffer that was `out` in the last pass = sm.b
}

// (B, F, T) -> (B, T, F): magnitudes and initial phases, once.
__global__ void gl_setup_kernel(const float* __restrict__ S, const float2* __restrict__ ang0, int B, int T,
                                float* __restrict__ St, float2* __restrict__ ang) {
  __shared__ float ts[32][33];
  __shared__ float2 ta[32][33];
  const int b = blockIdx.z, f0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int f = f0 + i, t = t0 + threadIdx.x;
    if (f < GL_F && t < T) {
      ts[i][threadIdx.x] = S[((long)b * GL_F + f) * T + t];
      ta[i][threadIdx.x] = ang0[((long)b * GL_F + f) * T + t];
    }
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, f = f0 + threadIdx.x;
    if (f < GL_F && t < T) {
      St[((long)b * T + t) * GL_F + f] = ts[threadIdx.x][i];
      ang[((long)b * T + t) * GL_F + f] = ta[threadIdx.x][i];
    }
  }
}

// Operands of one frame pair for the inverse STFT (S * angles of frames t and t + 1).
// The spectra of the next pair are fetched into registers before the FFT passes of the current one.
struct IstftRegs {
  float2 A[3], Bv[3];      // bins tid, tid + 256, (tid == 0: 512) of the two frames, already scaled by S
};

__device__ __forceinline__ void istft_fetch_bt(IstftRegs& r, const float* __restrict__ St, const float2* __restrict__ ang,
                                               int T, int b, int t, int tid) {
  const bool second = t + 1 < T;
  const float* Sa = St + ((long)b * T + t) * GL_F;
  const float2* Aa = ang + ((long)b * T + t) * GL_F;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int k = tid + GL_T * i;
    r.A[i] = r.Bv[i] = make_float2(0.f, 0.f);
    if (k < GL_F && (i < 2 || tid == 0)) {
      const float2 a = Aa[k];
      const float sa = Sa[k];
      r.A[i] = make_float2(a.x * sa, a.y * sa);
      if (second) {
        const float2 bb = Aa[GL_F + k];
        const float sb = Sa[GL_F + k];
        r.Bv[i] = make_float2(bb.x * sb, bb.y * sb);
      }
    }
  }
}

// Inverse STFT with the overlap-add fused in: a CTA owns GL_CH consecutive output hops of one utterance, runs the
// (up to GL_CH + 3) frames that touch them -- three frames of halo are recomputed instead of exchanged -- through the
// paired FFT, and accumulates their windowed segments in shared memory in frame order (deterministic, no atomics).
// The windowed frames never go to global memory (35 % of the loop's DRAM traffic), and there is no separate
// overlap-add pass.
constexpr int GL_CH = 24;
__global__ void __launch_bounds__(GL_T) gl_istft_ola_kernel(const float* __restrict__ St, const float2* __restrict__ ang,
                                                            int T, int chunks_per_utt, float* __restrict__ y) {
  __shared__ GlSmem sm;
  extern __shared__ float accs[];                       // [GL_CH][256]
  const int tid = threadIdx.x;
  const int b = blockIdx.x / chunks_per_utt, c = blockIdx.x % chunks_per_utt;
  const int o0 = c * GL_CH;                              // first output hop of this chunk
  const int nh = min(GL_CH, (T - 1) - o0);               // output hops here (the signal has T - 1 hops)
  const int t_lo = max(0, o0 - 1), t_hi = min(T - 1, o0 + nh + 1);      // frames touching padded hops o0 + 2 .. o0 + nh + 1
  gl_tables(sm, tid);
  for (int i = tid; i < GL_CH * GL_HOP; i += GL_T) accs[i] = 0.f;
  IstftRegs cur;
  istft_fetch_bt(cur, St, ang, T, b, t_lo, tid);
  __syncthreads();
  const float sc = 1.0f / (float)GL_N;
  for (int t = t_lo; t <= t_hi; t += 2) {
    const bool second = t + 1 <= t_hi;                  // a frame past t_hi contributes nothing here: drop it
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int k = tid + GL_T * i;
      if (i < 2 || tid == 0) {
        float2 A = cur.A[i], Bv = second ? cur.Bv[i] : make_float2(0.f, 0.f);
        if (k == 0 || k == GL_N / 2) { A.y = 0.f; Bv.y = 0.f; }
        sm.a[pidx(k)] = make_float2(A.x - Bv.y, A.y + Bv.x);
        if (k > 0 && k < GL_N / 2) sm.a[pidx(GL_N - k)] = make_float2(A.x + Bv.y, Bv.x - A.y);
      }
    }
    __syncthreads();
    if (t + 2 <= t_hi) istft_fetch_bt(cur, St, ang, T, b, t + 2, tid);      // lands under the FFT
    fft1024<true>(sm, tid);
#pragma unroll
    for (int q = 0; q < 4; ++q) {                        // segment q of frame t lands in padded hop t + q = output hop t + q - 2
      const int n = tid + GL_HOP * q;
      const float2 z = sm.b[pidx(n)];
      const float w = sm.win[n] * sc;
      const int ia = t + q - 2 - o0, ib = ia + 1;
      if (ia >= 0 && ia < nh) accs[ia * GL_HOP + tid] += z.x * w;
      if (second && ib >= 0 && ib < nh) accs[ib * GL_HOP + tid] += z.y * w;
    }
    __syncthreads();
  }
  float* yb = y + (long)b * GL_HOP * (T - 1) + (long)o0 * GL_HOP;
  for (int i = tid; i < nh * GL_HOP; i += GL_T) {
    const int h = o0 + i / GL_HOP + 2, r = i % GL_HOP;   // padded hop, offset in it
    float wss = 1.5f;                                    // Hann^2 at 75 % overlap, all four frames present
    if (h < 3 || h > T - 1) {
      wss = 0.f;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const int t = h - d;
        if (t >= 0 && t < T) {
          const float w = sm.win[r + GL_HOP * d];
          wss = fmaf(w, w, wss);
        }
      }
    }
    yb[i] = wss > 1.17549435e-38f ? accs[i] / wss : accs[i];
  }
}

// STFT of frame pairs + momentum phase update.  rebuilt -> tprev, angles updated in place.
__global__ void __launch_bounds__(GL_T) gl_stft_kernel(const float* __restrict__ y, int T, long L, long n_pairs,
                                                       int pairs_per_utt, float c, int first,
                                                       float2* __restrict__ ang, float2* __restrict__ tprev) {
  __shared__ GlSmem sm;
  const int tid = threadIdx.x;
  gl_tables(sm, tid);
  __syncthreads();
  auto fetch = [&](long pr, float (&ya)[4], float (&yb2)[4]) {
    const int b = (int)(pr / pairs_per_utt), t = 2 * (int)(pr % pairs_per_utt);
    const bool second = t + 1 < T;
    const float* yb = y + (long)b * L;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = tid + GL_T * i;
      long ma = (long)t * GL_HOP + n - GL_N / 2;           // centre = True, reflect padding
      long mb = ma + GL_HOP;
      ma = ma < 0 ? -ma : (ma >= L ? 2 * (L - 1) - ma : ma);
      mb = mb < 0 ? -mb : (mb >= L ? 2 * (L - 1) - mb : mb);
      ya[i] = yb[ma];
      yb2[i] = second ? yb[mb] : 0.f;
    }
  };
  float ca[4], cb[4];
  if ((long)blockIdx.x < n_pairs) fetch(blockIdx.x, ca, cb);
  for (long pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
    const int b = (int)(pr / pairs_per_utt), t = 2 * (int)(pr % pairs_per_utt);
    const bool second = t + 1 < T;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = tid + GL_T * i;
      const float w = sm.win[n];
      sm.a[pidx(n)] = make_float2(w * ca[i], w * cb[i]);
    }
    __syncthreads();
    if (pr + gridDim.x < n_pairs) fetch(pr + gridDim.x, ca, cb);      // lands under the FFT
    float2* Aa = ang + ((long)b * T + t) * GL_F;
    float2* Pa = tprev + ((long)b * T + t) * GL_F;
    float2 pva[3], pvb[3];                                           // previous spectra of my bins: fetched under the FFT too
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int k = tid + GL_T * i;
      pva[i] = pvb[i] = make_float2(0.f, 0.f);
      if (!first && (i < 2 || tid == 0)) {
        pva[i] = Pa[k];
        if (second) pvb[i] = Pa[GL_F + k];
      }
    }
    fft1024<false>(sm, tid);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int k = tid + GL_T * i;
      if (i < 2 || tid == 0) {
        const float2 z = sm.b[pidx(k)];
        const float2 zc = sm.b[pidx((GL_N - k) & (GL_N - 1))];       // Z[N - k], conjugated below
        // A = (Z[k] + conj(Z[N-k])) / 2,  B = (Z[k] - conj(Z[N-k])) / (2i)
        const float2 A = make_float2(0.5f * (z.x + zc.x), 0.5f * (z.y - zc.y));
        const float2 Bv = make_float2(0.5f * (z.y + zc.y), 0.5f * (zc.x - z.x));
        {
          const float2 n_ = make_float2(A.x - c * pva[i].x, A.y - c * pva[i].y);
          const float inv = 1.0f / (sqrtf(n_.x * n_.x + n_.y * n_.y) + 1e-16f);
          Aa[k] = make_float2(n_.x * inv, n_.y * inv);
          Pa[k] = A;
        }
        if (second) {
          const float2 n_ = make_float2(Bv.x - c * pvb[i].x, Bv.y - c * pvb[i].y);
          const float inv = 1.0f / (sqrtf(n_.x * n_.x + n_.y * n_.y) + 1e-16f);
          Aa[GL_F + k] = make_float2(n_.x * inv, n_.y * inv);
          Pa[GL_F + k] = Bv;
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace

// St is padded to a multiple of 4 floats so that the complex arrays behind it stay 16-byte aligned
static size_t st_floats(size_t bt) { return (bt * GL_F + 3) / 4 * 4; }

size_t griffin_lim_workspace_floats(int B, int T) {
  const size_t bt = (size_t)B * T;
  return st_floats(bt) /*St*/ + 2 * bt * GL_F /*ang*/ + 2 * bt * GL_F /*tprev*/;
}

int launch_griffin_lim(const float* S, const float* angles0_ri, int B, int T, int n_iter, float momentum, float* y,
                       float* workspace, cudaStream_t s) {
  const size_t bt = (size_t)B * T;
  float* St = workspace;
  float2* ang = reinterpret_cast<float2*>(St + st_floats(bt));
  float2* tprev = ang + bt * GL_F;
  const long L = (long)GL_HOP * (T - 1);
  const int pairs_per_utt = (T + 1) / 2;
  const long n_pairs = (long)B * pairs_per_utt;
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int fft_grid = (int)(n_pairs < (long)sms * 6 ? n_pairs : (long)sms * 6);
  const float c = momentum / (1.0f + momentum);

  gl_setup_kernel<<<dim3((T + 31) / 32, (GL_F + 31) / 32, B), dim3(32, 8), 0, s>>>(
      S, reinterpret_cast<const float2*>(angles0_ri), B, T, St, ang);
  SSV_CUDA(cudaGetLastError());
  ++g_launches;
  const int chunks_per_utt = (T - 1 + GL_CH - 1) / GL_CH;
  const size_t acc_bytes = (size_t)GL_CH * GL_HOP * sizeof(float);
  static bool configured = false;
  if (!configured) {
    SSV_CUDA(cudaFuncSetAttribute(gl_istft_ola_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)acc_bytes));
    configured = true;
  }
  for (int it = 0; it <= n_iter; ++it) {
    gl_istft_ola_kernel<<<B * chunks_per_utt, GL_T, acc_bytes, s>>>(St, ang, T, chunks_per_utt, y);
    ++g_launches;
    if (it == n_iter) break;
    gl_stft_kernel<<<fft_grid, GL_T, 0, s>>>(y, T, L, n_pairs, pairs_per_utt, c, it == 0 ? 1 : 0, ang, tprev);
    ++g_launches;
  }
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace ssv
