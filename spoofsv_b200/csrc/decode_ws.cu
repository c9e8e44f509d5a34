// Weight-stationary, layer-pipelined incremental Text2Mel decode (the default decode kernel).
//
// Replaces the reference AR loop (generate_test_utterances.py:105-116, synthesize.py:103-109 driving the
// eval branch of melSyn.forward, models/TTSModel.py:275-300).  Algorithm as in decode.cu: every causal
// highwayConv keeps the history of its own input, so frame t costs O(1) instead of O(t).
//
// What is different here is WHERE the work lives.  The 24 mat-vec stages of a frame (AudioEnc 13,
// attention + AudioDec 11) own 6.8 M fp32 weights = 27.3 MB; the 148 SMs of a B200 have 33 MB of shared
// memory between them.  So every stage gets a fixed group of CTAs (8 per highway layer, 1-4 per 1x1 conv,
// 144 in total), each CTA loads its [K][ncol] weight slice into shared memory ONCE per launch, and the
// utterances flow through the stage groups as micro-batches of RT rows:
//
//     stage s, micro-batch g, frame t   needs   stage s-1, g, t      (stage 0: stage 23, g, t-1)
//
// which is a software pipeline over (frame, micro-batch): up to 24 micro-batches are in flight, HBM/L2 see
// only activations, and there is no grid-wide barrier anywhere.
//
// Cross-CTA hand-off: every value a CTA publishes is an 8-byte word {float value, int tag} written with one
// 64-bit store (single-copy atomic); tag = seq_base + frame + 1.  A consumer first polls one sentinel word
// per producer CTA (cheap: 32 B per round), then loads the row and checks every tag, re-loading until all
// match -- so no fences and no release/acquire chains sit on the critical path, and a word that has not
// landed yet can never be mistaken for data.  Everything else a CTA touches is private to it: its weight
// slice, its LayerNorm parameters, and its own ring of past stage inputs (taps t-d, t-2d), so the only
// cross-CTA traffic is the tagged words.  Consumers redo the cheap LayerNorm / highway gate / windowed
// attention of their input row redundantly (as decode.cu does), which is what lets a stage boundary be a
// single hand-off.
#include "decode.cuh"

#include <cstdlib>

namespace ssv {

namespace {

constexpr int NT = 512;
constexpr int HD = 256;
constexpr long long SPIN_LIMIT = 4000000000LL;   // ~2 s of SM clocks
constexpr unsigned FULL = 0xffffffffu;

// fixed shared-memory carve-up (floats)
constexpr int SM_W = 768 * 64;             // weight slice, at most [768][64]
constexpr int SM_SCRATCH_MAX = 4096;       // highway CTA: X [768][RT<=4] (3072) | partials [16][RT][64] (4096)
constexpr int SM_LN = SM_W + SM_SCRATCH_MAX;   // [4][256] LayerNorm parameters of my prologue
constexpr int SM_BIAS = SM_LN + 4 * HD;    // [256] bias of my columns
constexpr int SM_PMA = SM_BIAS + 256;      // [WS_MAX_BATCH] ints (attention stage only)
constexpr int SM_TOTAL = SM_PMA + WS_MAX_BATCH;

struct __align__(8) Word { float v; int tag; };

__device__ __forceinline__ void st_word(Word* p, float v, int tag) {
  asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_int(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ void st_word2(Word* p, float v0, float v1, int tag) {      // p 16-byte aligned
  asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1, %2, %3, %2};" ::"l"(p), "r"(__float_as_int(v0)), "r"(tag),
               "r"(__float_as_int(v1))
               : "memory");
}
__device__ __forceinline__ void ld_word(const Word* p, float& v, int& tag) {
  int a, b;
  asm volatile("ld.relaxed.gpu.global.v2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "l"(p) : "memory");
  v = __int_as_float(a);
  tag = b;
}
__device__ __forceinline__ void ld_word2(const Word* p, float& v0, int& t0, float& v1, int& t1) {   // 16-byte aligned
  int a, b, c, d;
  asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
  v0 = __int_as_float(a); t0 = b; v1 = __int_as_float(c); t1 = d;
}
__device__ __forceinline__ int ld_relaxed_s32(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_s32(int* p, int v) {
  asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ void warp_sum2(float& a, float& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(FULL, a, o);
    b += __shfl_xor_sync(FULL, b, o);
  }
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int RT> struct XVec;
template <> struct XVec<1> {
  static __device__ __forceinline__ void ld(const float* p, float (&x)[1]) { x[0] = *p; }
};
template <> struct XVec<2> {
  static __device__ __forceinline__ void ld(const float* p, float (&x)[2]) {
    const float2 v = *reinterpret_cast<const float2*>(p);
    x[0] = v.x; x[1] = v.y;
  }
};
template <> struct XVec<4> {
  static __device__ __forceinline__ void ld(const float* p, float (&x)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
  }
};

// Partial products of the CTA's [K][4*CG] weight slice with the RT input rows X[K][RT].
// Thread (cg, ks) owns 4 columns and the k indices ks + KS*i; partial sums of every k-slice go to
// part[slice][r][col] (aliases X: the caller's data in X is dead after the barrier inside).
template <int RT, int CG>
__device__ __forceinline__ void gemv_partials(const float* __restrict__ Ws, float* __restrict__ Xs, int K, int tid) {
  constexpr int NCOL = 4 * CG;
  constexpr int KS = NT / CG;
  const int cg = tid % CG, ks = tid / CG;
  const int iters = K / KS;
  float acc[RT][4];
#pragma unroll
  for (int r = 0; r < RT; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
  const float* wp = Ws + ks * NCOL + cg * 4;
  const float* xp = Xs + ks * RT;
#pragma unroll 8
  for (int i = 0; i < iters; ++i) {
    const float4 w = *reinterpret_cast<const float4*>(wp + (size_t)i * KS * NCOL);
    float x[RT];
    XVec<RT>::ld(xp + i * KS * RT, x);
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      acc[r][0] = fmaf(x[r], w.x, acc[r][0]);
      acc[r][1] = fmaf(x[r], w.y, acc[r][1]);
      acc[r][2] = fmaf(x[r], w.z, acc[r][2]);
      acc[r][3] = fmaf(x[r], w.w, acc[r][3]);
    }
  }
  int slice;
  bool writer = true;
  if (CG == 16) {               // the two half-warps hold adjacent k-slices of the same columns
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] += __shfl_xor_sync(FULL, acc[r][c], 16);
    slice = tid >> 5;
    writer = (tid & 16) == 0;
  } else {
    slice = ks;
  }
  __syncthreads();              // every thread is done reading X
  if (writer) {
#pragma unroll
    for (int r = 0; r < RT; ++r)
      *reinterpret_cast<float4*>(Xs + ((size_t)slice * RT + r) * NCOL + cg * 4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  }
  __syncthreads();
}

template <int RT>
__global__ void __launch_bounds__(NT, 1) decode_ws_kernel(const DecParams p) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int s_bad;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- which stage / column slice am I?
  int s = -1;
#pragma unroll 1
  for (int i = 0; i < DEC_STAGES; ++i) {
    const int c0 = p.ws_stages[i].cta0;
    if ((int)blockIdx.x >= c0 && (int)blockIdx.x < c0 + p.ws_stages[i].parts) s = i;
  }
  if (s < 0) return;
  const int prev = (s + DEC_STAGES - 1) % DEC_STAGES;
  const WsStage st = p.ws_stages[s];
  const int prev_parts = p.ws_stages[prev].parts;
  const int part = (int)blockIdx.x - st.cta0;
  const bool designated = part == 0;
  const int K = st.K, ncol = st.ncol, half = ncol / 2;
  const int G = p.G, B = p.B;

  float* Ws = smem;
  float* Xs = smem + (size_t)K * ncol;     // scratch: X [K][RT], later partials [slices][RT][ncol]
  float* lnp = smem + SM_LN;
  float* bias_s = smem + SM_BIAS;
  int* pma_s = reinterpret_cast<int*>(smem + SM_PMA);
  const float* g1 = lnp;
  const float* b1 = lnp + HD;
  const float* g2 = lnp + 2 * HD;
  const float* b2 = lnp + 3 * HD;

  // global column (= tagged word index) of local column lc; highway stages own matching H1 / H2 slices
  auto gcol = [&](int lc) { return st.hwy ? (lc < half ? part * half + lc : HD + part * half + (lc - half)) : part * ncol + lc; };

  // ---- one-time loads: weight slice, LayerNorm parameters, bias, alignment state
  {
    const float* img = st.img + (size_t)part * K * ncol;
    const int n4 = K * ncol / 4;
    for (int i = tid; i < n4; i += NT) cp_async16(Ws + (size_t)i * 4, img + (size_t)i * 4, true);
    for (int i = tid; i < 4 * HD; i += NT) {
      const int which = i / HD, c = i % HD;
      const float* src = which == 0 ? st.g1 : which == 1 ? st.b1 : which == 2 ? st.g2 : st.b2;
      const int len = st.pro == PRO_X ? p.F : HD;
      lnp[i] = (src != nullptr && c < len) ? src[c] : 0.f;
    }
    for (int i = tid; i < 256; i += NT) {
      float v = 0.f;
      if (i < ncol) { const int gc = gcol(i); if (gc < st.n) v = st.bias[gc]; }
      bias_s[i] = v;
    }
    if (st.pro == PRO_ATT) {
      for (int i = tid; i < B; i += NT) {
        const int v = p.pma_in ? (int)p.pma_in[i] : p.pma_state[i];
        pma_s[i] = max(0, min(v, p.N - 1));
      }
    }
    if (tid == 0) s_bad = 0;
    cp_async_wait_all();
    __syncthreads();
  }

  long long prof_last = 0;
  long long prof_acc[7] = {0, 0, 0, 0, 0, 0, 0};
  const bool prof_on = p.prof != nullptr && tid == 0;
  if (prof_on) prof_last = clock64();
#define PROF_T(i)                                  \
  if (prof_on) {                                   \
    const long long now_ = clock64();              \
    prof_acc[i] += now_ - prof_last;               \
    prof_last = now_;                              \
  }

  const Word* raw_in = reinterpret_cast<const Word*>(p.ws_raw) + (size_t)prev * B * WS_WORDS;
  Word* raw_out = reinterpret_cast<Word*>(p.ws_raw) + (size_t)s * B * WS_WORDS;
  const int koff = (st.ntaps - 1) * st.k_seg;               // X row of the current tap
  const int n_visits = p.n_steps + (s == 0 ? 1 : 0);        // stage 0 also finishes the last frame (y = sigmoid(LN5))
  long visits = 0;

  for (int step = 0; step < n_visits; ++step) {
    const int t = p.t_start + step;
    const bool final_visit = s == 0 && step == p.n_steps;
    const int tag = p.seq_base + t + 1;                     // tag of everything produced for frame t
    const int tag_in = s == 0 ? tag - 1 : tag;              // stage 0 consumes frame t-1 of stage 23
    const bool need_wait = !(s == 0 && step == 0);
    for (int g = 0; g < G; ++g, ++visits) {
      const int row0 = g * RT;
      const int nrows = min(RT, B - row0);
      PROF_T(0);

      // ---- 1. old taps t-2d, t-d of this micro-batch from my private ring -> X[0 .. 2*256)  (independent of the wait)
      if (st.ntaps == 3) {
        const int chunks = HD * RT / 4;                      // 16-byte chunks per tap block
        for (int i = tid; i < 2 * chunks; i += NT) {
          const int j = i / chunks, c4 = i - j * chunks;
          const int tt = t - (2 - j) * st.dil;
          const bool ok = tt >= 0;
          const int slot = ok ? tt % st.hist_depth : 0;
          const float* src = p.ws_hist + (((size_t)(st.hist_blk0 + part * st.hist_depth + slot)) * G + g) * (HD * RT) + c4 * 4;
          cp_async16(Xs + (size_t)j * HD * RT + c4 * 4, src, ok);
        }
      }
      PROF_T(1);

      // ---- 2./3. wait for the producers of my input row, then the prologue: u_t -> X[koff ..][r]
      if (warp < RT) {
        const int r = warp, b = row0 + r;
        float* xcur = Xs + (size_t)koff * RT + r;            // channel c at xcur[c * RT]
        if (r >= nrows) {
          if (!final_visit)
            for (int c = lane; c < st.k_seg; c += 32) xcur[(size_t)c * RT] = 0.f;
        } else {
          bool bad = false;
          long long t0 = 0;
          unsigned spins = 0;
          auto spin_check = [&]() {        // bounded spinning: flag the abort and leave
            if ((++spins & 255u) == 0) {
              if (t0 == 0) t0 = clock64();
              else if (clock64() - t0 > SPIN_LIMIT) atomicExch(p.abort_flag, 8);
              if (*reinterpret_cast<volatile int*>(p.abort_flag) != 0) bad = true;
            }
          };
          if (need_wait) {
            const int* sp = p.ws_sent + ((size_t)prev * G + g) * WS_MAX_PARTS;
            for (;;) {
              const int v = lane < prev_parts ? ld_relaxed_s32(sp + lane) : tag_in;
              if (__all_sync(FULL, v - tag_in >= 0)) break;
              spin_check();
              if (__any_sync(FULL, bad)) { bad = true; break; }
            }
          }
          const Word* R = raw_in + (size_t)b * WS_WORDS;
          const int pro = st.pro;
          if (pro == PRO_X) {
            float y[3];
            if (!need_wait) {
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                const int f = lane + 32 * i;
                float v = 0.f;
                if (f < p.F) {
                  if (p.x_ext) v = p.x_ext[(long)b * p.x_sb + (long)f * p.x_sf];
                  else if (t > 0) v = __ldcg(p.Y + ((size_t)b * p.F + f) * p.t_cap + (t - 1));
                }
                y[i] = v;
              }
            } else {
              float v[3] = {0.f, 0.f, 0.f};
              while (!bad) {
                bool ok = true;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                  const int f = lane + 32 * i;
                  v[i] = 0.f;
                  if (f < p.F) { int tg; ld_word(R + f, v[i], tg); ok &= tg == tag_in; }
                }
                if (__all_sync(FULL, ok)) break;
                spin_check();
                if (__any_sync(FULL, bad)) bad = true;
              }
              float sm = v[0] + v[1] + v[2];
              sm = warp_sum(sm);
              const float mean = sm / (float)p.F;
              float q = 0.f;
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                const float d = v[i] - mean;
                q += lane + 32 * i < p.F ? d * d : 0.f;
              }
              q = warp_sum(q);
              const float rstd = 1.0f / sqrtf(q / (float)p.F + 1e-5f);
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                const int f = lane + 32 * i;
                y[i] = f < p.F ? sigmoidf_((v[i] - mean) * rstd * g1[f] + b1[f]) : 0.f;
                if (f < p.F && !bad) p.Y[((size_t)b * p.F + f) * p.t_cap + (t - 1)] = y[i];
              }
            }
            if (!final_visit) {
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                const int f = lane + 32 * i;
                if (f < p.F) xcur[(size_t)f * RT] = y[i];
              }
            }
          } else if (pro == PRO_LN || pro == PRO_LN_RELU) {
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            while (!bad) {
              bool ok = true;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                int ta, tb;
                ld_word2(R + 2 * lane + 64 * i, v[2 * i], ta, v[2 * i + 1], tb);
                ok &= ta == tag_in && tb == tag_in;
              }
              if (__all_sync(FULL, ok)) break;
              spin_check();
              if (__any_sync(FULL, bad)) bad = true;
            }
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) sum += v[i];
            sum = warp_sum(sum);
            const float mean = sum / (float)HD;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
            q = warp_sum(q);
            const float rstd = 1.0f / sqrtf(q / (float)HD + 1e-5f);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int c = 2 * lane + 64 * (i >> 1) + (i & 1);
              float o = (v[i] - mean) * rstd * g1[c] + b1[c];
              if (pro == PRO_LN_RELU) o = fmaxf(o, 0.f);
              v[i] = o;
              xcur[(size_t)c * RT] = o;
            }
            if (st.hwy && (lane >> 4) == (part & 1)) {       // my residual slice travels with my outputs
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (i == (part >> 1))
                  st_word2(raw_out + (size_t)b * WS_WORDS + 2 * HD + 2 * lane + 64 * i, v[2 * i], v[2 * i + 1], tag);
            }
          } else {   // PRO_HWY / PRO_ATT: the producer is a highway layer: H1 | H2 | its input (my residual)
            float h1[8] = {}, h2[8] = {}, xr[8] = {};
            while (!bad) {
              bool ok = true;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                int ta, tb;
                const Word* w = R + 2 * lane + 64 * i;
                ld_word2(w, h1[2 * i], ta, h1[2 * i + 1], tb);
                ok &= ta == tag_in && tb == tag_in;
                ld_word2(w + HD, h2[2 * i], ta, h2[2 * i + 1], tb);
                ok &= ta == tag_in && tb == tag_in;
                ld_word2(w + 2 * HD, xr[2 * i], ta, xr[2 * i + 1], tb);
                ok &= ta == tag_in && tb == tag_in;
              }
              if (__all_sync(FULL, ok)) break;
              spin_check();
              if (__any_sync(FULL, bad)) bad = true;
            }
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { s1 += h1[i]; s2 += h2[i]; }
            warp_sum2(s1, s2);
            const float m1 = s1 / (float)HD, m2 = s2 / (float)HD;
            float q1 = 0.f, q2 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float d1 = h1[i] - m1, d2 = h2[i] - m2;
              q1 = fmaf(d1, d1, q1);
              q2 = fmaf(d2, d2, q2);
            }
            warp_sum2(q1, q2);
            const float r1 = 1.0f / sqrtf(q1 / (float)HD + 1e-5f);
            const float r2 = 1.0f / sqrtf(q2 / (float)HD + 1e-5f);
            float u[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int c = 2 * lane + 64 * (i >> 1) + (i & 1);
              const float a = (h1[i] - m1) * r1 * g1[c] + b1[c];
              const float bb = (h2[i] - m2) * r2 * g2[c] + b2[c];
              const float gt = sigmoidf_(a);
              u[i] = gt * bb + (1.0f - gt) * xr[i];
            }
            if (pro == PRO_HWY) {
#pragma unroll
              for (int i = 0; i < 8; ++i) xcur[(size_t)(2 * lane + 64 * (i >> 1) + (i & 1)) * RT] = u[i];
              if (st.hwy && (lane >> 4) == (part & 1)) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (i == (part >> 1))
                    st_word2(raw_out + (size_t)b * WS_WORDS + 2 * HD + 2 * lane + 64 * i, u[2 * i], u[2 * i + 1], tag);
              }
            } else {
              // windowed attention, models/TTSModel.py:281-295: logits over [pma, min(pma+2, N-1)];
              // every other character is masked to -2^32 and gets softmax weight exactly 0.
              const int p0 = pma_s[b];
              const int cnt = min(p0 + 2, p.N - 1) - p0 + 1;
              const float* kp = p.Kt + ((size_t)b * p.N + p0) * HD;
              const float* vp = p.Vt + ((size_t)b * p.N + p0) * HD;
              float l0 = 0.f, l1 = 0.f, l2 = 0.f;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int c = 2 * lane + 64 * i;
                const float2 k0 = __ldg(reinterpret_cast<const float2*>(kp + c));
                l0 = fmaf(k0.x, u[2 * i], l0); l0 = fmaf(k0.y, u[2 * i + 1], l0);
                if (cnt > 1) {
                  const float2 k1 = __ldg(reinterpret_cast<const float2*>(kp + HD + c));
                  l1 = fmaf(k1.x, u[2 * i], l1); l1 = fmaf(k1.y, u[2 * i + 1], l1);
                }
                if (cnt > 2) {
                  const float2 k2 = __ldg(reinterpret_cast<const float2*>(kp + 2 * HD + c));
                  l2 = fmaf(k2.x, u[2 * i], l2); l2 = fmaf(k2.y, u[2 * i + 1], l2);
                }
              }
              warp_sum2(l0, l1);
              l2 = warp_sum(l2);
              l0 *= 0.0625f; l1 *= 0.0625f; l2 *= 0.0625f;    // 1/sqrt(256)
              float m = l0;
              if (cnt > 1) m = fmaxf(m, l1);
              if (cnt > 2) m = fmaxf(m, l2);
              const float e0 = expf(l0 - m);
              const float e1 = cnt > 1 ? expf(l1 - m) : 0.f;
              const float e2 = cnt > 2 ? expf(l2 - m) : 0.f;
              const float den = e0 + e1 + e2;
              const float a0 = e0 / den, a1 = e1 / den, a2 = e2 / den;
              int best = 0;
              float bv = a0;
              if (cnt > 1 && a1 > bv) { best = 1; bv = a1; }
              if (cnt > 2 && a2 > bv) { best = 2; bv = a2; }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int c = 2 * lane + 64 * i;
                const float2 v0 = __ldg(reinterpret_cast<const float2*>(vp + c));
                float rx = a0 * v0.x, ry = a0 * v0.y;
                if (cnt > 1) {
                  const float2 v1 = __ldg(reinterpret_cast<const float2*>(vp + HD + c));
                  rx = fmaf(a1, v1.x, rx); ry = fmaf(a1, v1.y, ry);
                }
                if (cnt > 2) {
                  const float2 v2 = __ldg(reinterpret_cast<const float2*>(vp + 2 * HD + c));
                  rx = fmaf(a2, v2.x, rx); ry = fmaf(a2, v2.y, ry);
                }
                xcur[(size_t)c * RT] = rx;                    // R
                xcur[(size_t)(c + 1) * RT] = ry;
                xcur[(size_t)(HD + c) * RT] = u[2 * i];       // Q
                xcur[(size_t)(HD + c + 1) * RT] = u[2 * i + 1];
              }
              if (lane == 0 && !bad) {
                pma_s[b] = p0 + best;
                if (designated) {
                  float* Ab = p.A + ((size_t)b * p.N + p0) * p.t_cap + t;
                  Ab[0] = a0;
                  if (cnt > 1) Ab[p.t_cap] = a1;
                  if (cnt > 2) Ab[2 * (size_t)p.t_cap] = a2;
                  p.pma_traj[(size_t)t * B + b] = p0 + best;
                  p.pma_state[b] = p0 + best;
                }
              }
            }
          }
          if (bad && lane == 0) s_bad = 1;
        }
      }
      if (final_visit) continue;                 // stage 0 after the last frame: prologue only (uniform per CTA)
      PROF_T(2);
      cp_async_wait_all();
      __syncthreads();
      PROF_T(3);
      if (s_bad) break;

      // ---- 4. my stage input of frame t joins my private ring (taps of later frames)
      if (st.ntaps == 3) {
        const int slot = t % st.hist_depth;
        float* dst = p.ws_hist + (((size_t)(st.hist_blk0 + part * st.hist_depth + slot)) * G + g) * (HD * RT);
        for (int i = tid; i < HD * RT / 4; i += NT)
          __stcg(reinterpret_cast<float4*>(dst) + i, *reinterpret_cast<const float4*>(Xs + (size_t)koff * RT + (size_t)i * 4));
      }

      // ---- 5. mat-vec on the resident weight slice
      if (st.cg == 16) gemv_partials<RT, 16>(Ws, Xs, K, tid);
      else if (st.cg == 32) gemv_partials<RT, 32>(Ws, Xs, K, tid);
      else gemv_partials<RT, 64>(Ws, Xs, K, tid);
      PROF_T(4);

      // ---- 6. reduce the k-slices, add bias (+ hoisted speaker projection), publish tagged words
      {
        const int slices = st.cg == 64 ? 8 : 16;
        for (int o = tid; o < RT * ncol; o += NT) {
          const int r = o / ncol, lc = o - r * ncol;
          float v = 0.f;
          for (int sl = 0; sl < slices; ++sl) v += Xs[((size_t)sl * RT + r) * ncol + lc];
          const int gc = gcol(lc);
          if (r < nrows && gc < st.n) {
            const int b = row0 + r;
            v += bias_s[lc];
            if (st.bias_b == 1) v += __ldg(p.s1 + (size_t)b * HD + gc);
            else if (st.bias_b == 2) v += __ldg(p.s2 + (size_t)b * HD + gc);
            st_word(raw_out + (size_t)b * WS_WORDS + gc, v, tag);
          }
        }
      }
      __syncthreads();
      if (tid == 0) st_relaxed_s32(p.ws_sent + ((size_t)s * G + g) * WS_MAX_PARTS + part, tag);
      PROF_T(5);
    }
    if (s_bad) break;
  }
  if (prof_on) {
#pragma unroll
    for (int i = 0; i < 7; ++i) p.prof[(size_t)blockIdx.x * 8 + i] = prof_acc[i];
    p.prof[(size_t)blockIdx.x * 8 + 7] = visits > 0 ? visits : 1;
  }
#undef PROF_T
}

// dst[part][kk][lc] = W[gcol(lc)][kk] (W row-major [n][K]), zero for padded columns.
__global__ void ws_pack_image_kernel(const float* __restrict__ W, int n, int K, int parts, int ncol, int hwy,
                                     float* __restrict__ dst) {
  const long total = (long)parts * K * ncol;
  const int half = ncol / 2;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int lc = (int)(i % ncol);
    const int kk = (int)((i / ncol) % K);
    const int part = (int)(i / ((long)ncol * K));
    const int gc = hwy ? (lc < half ? part * half + lc : HD + part * half + (lc - half)) : part * ncol + lc;
    dst[i] = gc < n ? W[(long)gc * K + kk] : 0.f;
  }
}

template <int RT>
int launch_rt(const DecParams& p, cudaStream_t s) {
  constexpr size_t smem = (size_t)SM_TOTAL * sizeof(float);
  static bool configured = false;
  if (!configured) {
    SSV_CUDA(cudaFuncSetAttribute(decode_ws_kernel<RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  DecParams pl = p;
  void* args[] = {&pl};
  SSV_CUDA(cudaLaunchCooperativeKernel((void*)decode_ws_kernel<RT>, dim3(WS_GRID), dim3(NT), args, smem, s));
  ++g_launches;
  return kOk;
}

}  // namespace

// Stage -> CTA-group layout (static).  Highway layers: 8 CTAs x 64 columns (32 H1 + the matching 32 H2).
void ws_stage_layout(int s, const DecStage& d, WsStage* w) {
  w->n = d.n; w->k_seg = d.k_seg; w->ntaps = d.ntaps; w->dil = d.dil; w->pro = d.pro; w->bias_b = d.bias_b;
  w->K = d.ntaps * d.k_seg;
  w->bias = d.bias; w->g1 = d.g1; w->b1 = d.b1; w->g2 = d.g2; w->b2 = d.b2;
  w->hwy = d.n == 2 * HD ? 1 : 0;
  if (w->hwy) { w->parts = 8; w->ncol = 64; }
  else if (w->K > 256) { w->parts = 4; w->ncol = 64; }              // AudioDec conv1 (512 -> 256)
  else if (w->K < 256) { w->parts = 1; w->ncol = 256; }             // AudioEnc conv1 (80 -> 256)
  else if (d.n < 256) { w->parts = 1; w->ncol = 128; }              // AudioDec conv5 (256 -> 80), columns padded
  else { w->parts = 2; w->ncol = 128; }                             // 256 -> 256
  w->cg = w->ncol / 4;
  w->hist_depth = d.ntaps == 3 ? 2 * d.dil + 1 : 0;
  (void)s;
}

int ws_pack_image(const float* W_rowmajor, const WsStage& w, float* dst, cudaStream_t s) {
  const long total = (long)w.parts * w.K * w.ncol;
  long g = (total + 255) / 256;
  if (g > 4096) g = 4096;
  ws_pack_image_kernel<<<(int)g, 256, 0, s>>>(W_rowmajor, w.n, w.K, w.parts, w.ncol, w.hwy, dst);
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

bool decode_ws_supported(int sm_count) {
  static int ok = -1;
  if (ok < 0) {
    ok = 0;
    int dev = 0, coop = 0, smem_max = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess)
      ok = coop && smem_max >= (int)(SM_TOTAL * sizeof(float)) ? 1 : 0;
    cudaGetLastError();
  }
  return ok == 1 && sm_count >= WS_GRID;
}

int launch_decode_ws(const DecParams& p, cudaStream_t s) {
  SSV_CHECK(p.H == HD, "decode: hidden_dim must be %d", HD);
  SSV_CHECK(p.F <= 96 && p.F % 4 == 0, "decode: freq_bins must be <= 96 and a multiple of 4");
  SSV_CHECK(p.B >= 1 && p.n_steps >= 1, "decode: empty launch");
  SSV_CHECK(p.B <= WS_MAX_BATCH, "decode: batch %d exceeds %d", p.B, WS_MAX_BATCH);
  SSV_CHECK(p.ws_stages && p.ws_raw && p.ws_sent && p.ws_hist, "decode: weight-stationary buffers missing");
  if (p.R == 1) return launch_rt<1>(p, s);
  if (p.R == 2) return launch_rt<2>(p, s);
  SSV_CHECK(p.R == 4, "decode: micro-batch rows must be 1, 2 or 4");
  return launch_rt<4>(p, s);
}

}  // namespace ssv
