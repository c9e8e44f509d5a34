// Persistent incremental Text2Mel decode kernel (AudioEnc -> windowed attention -> AudioDec).
//
// Replaces the reference AR loop (generate_test_utterances.py:105-116, synthesize.py:103-109),
// which re-runs audio_encoder/audio_decoder over the whole mel prefix every frame
// (models/TTSModel.py:280-295, O(T^2)), by an O(T) recurrence: every causal highwayConv keeps the
// history of its own input, so frame t needs only rows t, t-d, t-2d of each layer.
//
// One cooperative launch runs n_steps frames.  A frame is 24 dependent mat-vec stages; a stage is
//   prologue : every CTA rebuilds the stage input u_t for its row group from the previous stage's
//              raw (pre-LayerNorm) GEMV output -- LayerNorm(+ReLU), the highway gate, or the
//              3-character windowed attention (models/TTSModel.py:281-295) -- redundantly, so that
//              one grid barrier per stage suffices (no separate "normalise" phase);
//   GEMV     : output columns are split over all warps of the grid (one warp = two columns,
//              H1[c]/H2[c] of a highway layer), weights prefetched into registers BEFORE the barrier,
//              older taps prefetched into shared memory before the barrier as well;
//   barrier  : one monotonically increasing global counter, release/acquire, bounded spin.
// Batch rows are split into row groups of <=16 utterances; CTAs are dealt round-robin to row groups.
#include "decode.cuh"

#include <cstdlib>

namespace ssv {

namespace {

constexpr int NT = 512;
constexpr int NW = NT / 32;
constexpr int HD = 256;              // hidden width (validated on the host)
constexpr int NE = HD / 32;          // elements per lane of a hidden vector
constexpr int KMAX = DEC_XS_LD;      // longest GEMV reduction (3 taps x 256)
constexpr long long SPIN_LIMIT = 4000000000LL;   // ~2 s of SM clocks

__device__ __forceinline__ int ld_acquire_s32(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_s32(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void warp_sum2(float& a, float& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  return fmaf(a.w, b.w, acc);
}

// Barrier among the `members` CTAs of one row group: every CTA publishes the generation it has
// finished in its own flag word (plain release store -- no atomics to serialise in L2) and ONE warp
// polls all member flags with coalesced relaxed loads (<= 5 cache lines per round; a thread-per-flag
// poll issues ~150x the L2 requests and delays the very stores it waits for).
__device__ __forceinline__ int ld_relaxed_s32(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void group_arrive(int* flags, int my_index, int gen) {
  // caller has just executed __syncthreads(): every store of this CTA precedes the release
  if (threadIdx.x == 0) st_release_s32(flags + my_index, gen);
}
__device__ __forceinline__ bool group_wait(int* flags, int members, int gen, int* abort_flag) {
  int bad = 0;
  if (threadIdx.x < 32) {
    const long long t0 = clock64();
    unsigned spins = 0;
    for (;;) {
      int ok = 1;
      for (int i = threadIdx.x; i < members; i += 32) ok &= ld_relaxed_s32(flags + i) >= gen;
      if (__all_sync(0xffffffffu, ok)) break;
      if ((++spins & 1023u) == 0) {
        if (clock64() - t0 > SPIN_LIMIT) atomicExch(abort_flag, 2);
        if (*reinterpret_cast<volatile int*>(abort_flag) != 0) { bad = 1; break; }
      }
    }
    // acquire: re-read the (now final) flags with acquire semantics so the data loads that follow
    // the CTA barrier are ordered after them; no MEMBAR that would drain the in-flight prefetch
    for (int i = threadIdx.x; i < members; i += 32) (void)ld_acquire_s32(flags + i);
  }
  return __syncthreads_or(bad) == 0;
}

// Sum V per-lane values across the warp; lane L ends up holding the total of value index
// fold_index<V>(L) (a transposing butterfly: V + log2 shuffles instead of 5 V).
template <int V>
__device__ __forceinline__ float fold_reduce(float (&v)[V], int lane) {
  int off = 16;
#pragma unroll
  for (int n = V; n > 1; n >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = hi ? v[i] : v[i + n / 2];
      const float keep = hi ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    off >>= 1;
  }
  for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
  return v[0];
}
template <int V>
__device__ __forceinline__ int fold_index(int lane) {
  int idx = 0, off = 16;
#pragma unroll
  for (int n = V; n > 1; n >>= 1) {
    if (lane & off) idx += n / 2;
    off >>= 1;
  }
  return idx;
}
template <int V>
__device__ __forceinline__ bool fold_writer(int lane) {
  constexpr int rep = 32 / V;      // after log2(V) folding levels the low lane bits hold replicas
  return (lane & (rep - 1)) == 0;
}

// y = sigmoid(LN5(raw[0:F])) for one row (F <= 96); lane handles f = lane + 32 i.
__device__ __forceinline__ void final_row(const DecParams& p, const float* raw_row, int lane, float (&y)[3]) {
  float v[3];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int f = lane + 32 * i;
    v[i] = f < p.F ? __ldcg(raw_row + f) : 0.f;
    s += v[i];
  }
  s = warp_sum(s);
  const float mean = s / (float)p.F;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int f = lane + 32 * i;
    const float d = v[i] - mean;
    q += f < p.F ? d * d : 0.f;
  }
  q = warp_sum(q);
  const float rstd = 1.0f / sqrtf(q / (float)p.F + 1e-5f);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int f = lane + 32 * i;
    y[i] = f < p.F ? sigmoidf_((v[i] - mean) * rstd * __ldg(p.fin_g + f) + __ldg(p.fin_b + f)) : 0.f;
  }
}

// ROWS: utterances per row group.  RC: rows per GEMV task (ROWS / RC warps share a column pair).
template <int ROWS, int RC>
__global__ void __launch_bounds__(NT, 1) decode_kernel(const DecParams p) {
  constexpr int CHUNKS = ROWS / RC;           // row chunks per column pair
  constexpr int PAIRS = NW / CHUNKS;          // column pairs per CTA per stage
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                            // [ROWS][DEC_XS_LD]   stage input rows (taps | current)
  float* ws = smem + ROWS * DEC_XS_LD;         // [2][PAIRS][KMAX]    weights of this CTA's column pairs
  float* ps = ws + 2 * PAIRS * KMAX;           // [4][HD]: g1 b1 g2 b2 of the prologue
  __shared__ int pma_s[ROWS];
  __shared__ DecStage stab[DEC_STAGES];        // stage table (global copies sit behind an L1 that every
                                               // acquire load invalidates)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x;
  const int rg = blockIdx.x % p.RG;
  const int slot_cta = blockIdx.x / p.RG;
  const int Gr = G / p.RG;
  if (slot_cta >= Gr) return;                  // leftover CTAs take no part (barriers are per row group)
  const int row0 = rg * ROWS;
  const int nrows = min(ROWS, p.B - row0);
  const bool designated = slot_cta == 0;
  int* flags = reinterpret_cast<int*>(p.bar_counter) + rg * Gr;

  for (int i = tid; i < ROWS * DEC_XS_LD; i += NT) xs[i] = 0.f;
  for (int i = tid; i < (int)(DEC_STAGES * sizeof(DecStage) / sizeof(int)); i += NT)
    reinterpret_cast<int*>(stab)[i] = reinterpret_cast<const int*>(p.stages)[i];
  if (tid < ROWS) {
    int v = 0;
    if (tid < nrows) v = p.pma_in ? (int)p.pma_in[row0 + tid] : p.pma_state[row0 + tid];
    pma_s[tid] = max(0, min(v, p.N - 1));
  }
  __syncthreads();

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
  const int total = p.n_steps * DEC_STAGES + 1;
  const size_t raw_stage = (size_t)p.B * DEC_RAW_LD;
  const size_t hist_buf = (size_t)p.B * p.t_cap * HD;
  const int pair_local = warp / CHUNKS;
  const int chunk = warp % CHUNKS;

  for (int gs = 0; gs < total; ++gs) {
    const bool final_stage = gs == total - 1;
    const int step = gs / DEC_STAGES;
    const int s = final_stage ? 0 : gs - step * DEC_STAGES;
    const int t = p.t_start + step;           // for the final stage: t == t_last + 1
    const DecStage* sp = stab + s;
    const int n = sp->n, k_seg = sp->k_seg, ntaps = sp->ntaps, dil = sp->dil;
    const int K = ntaps * k_seg;
    const int half = n / 2;
    const int ppc = min(PAIRS, (half + Gr - 1) / Gr);     // column pairs this stage gives each CTA
    const int pair0 = slot_cta * ppc;
    PROF_T(0);

    // ---- 0. publish completion of the previous stage BEFORE starting new memory traffic
    if (gs > 0) group_arrive(flags, slot_cta, gs);

    // ---- 1. prefetch (independent of the previous stage): weights and old taps -> smem (cp.async)
    if (!final_stage) {
      const float* Wg = sp->W;
      const int k4 = K / 4;
      for (int row = warp; row < 2 * ppc; row += NW) {       // row = h * ppc + pr
        const int h = row >= ppc ? 1 : 0;
        const int pr = row - h * ppc;
        const int col = pair0 + pr;
        const bool ok = col < half;
        const float* src = ok ? Wg + (size_t)(col + h * half) * K : Wg;
        float* dst = ws + (h * PAIRS + pr) * KMAX;
        for (int c4 = lane; c4 < k4; c4 += 32) cp_async16(dst + c4 * 4, src + (ok ? c4 * 4 : 0), ok);
      }
      if (ntaps == 3) {
        const float* hb = p.hist + (size_t)sp->hist_in * hist_buf;
        for (int rj = warp; rj < nrows * 2; rj += NW) {
          const int r = rj >> 1, j = rj & 1;
          const int tt = t - (2 - j) * dil;
          const bool ok = tt >= 0;
          const float* src = ok ? hb + ((size_t)(row0 + r) * p.t_cap + tt) * HD : hb;
          float* dst = xs + r * DEC_XS_LD + j * HD;
          for (int c4 = lane; c4 < HD / 4; c4 += 32) cp_async16(dst + c4 * 4, src + (ok ? c4 * 4 : 0), ok);
        }
      }
      // LayerNorm affine parameters the prologue of this stage applies
      if (tid < 4 * (HD / 4)) {
        const int which = tid / (HD / 4), c4 = tid % (HD / 4);
        const float* src = which == 0 ? sp->g1 : which == 1 ? sp->b1 : which == 2 ? sp->g2 : sp->b2;
        if (src != nullptr) cp_async16(ps + which * HD + c4 * 4, src + c4 * 4, true);
      }
    }

    PROF_T(1);
    if (gs > 0) {
      if (!group_wait(flags, Gr, gs, p.abort_flag)) return;
    }
    PROF_T(2);

    // ---- 3. prologue: u_t for my rows -> xs[r][(ntaps-1)*k_seg ...]
    const int prev = (s + DEC_STAGES - 1) % DEC_STAGES;
    const float* rawp = p.raw + (size_t)prev * raw_stage;
    const int pro = final_stage ? PRO_X : sp->pro;
    const float* g1 = ps;
    const float* b1 = ps + HD;
    const float* g2 = ps + 2 * HD;
    const float* b2 = ps + 3 * HD;
    if (pro != PRO_X) {            // the prologue reads the cp.async'ed LayerNorm parameters
      cp_async_wait_all();
      __syncthreads();
    }
    for (int r = warp; r < nrows; r += NW) {
      const int b = row0 + r;
      float* xrow = xs + r * DEC_XS_LD + (ntaps - 1) * k_seg;
      if (pro == PRO_X) {
        float y[3];
        if (gs == 0) {
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
          final_row(p, rawp + (size_t)b * DEC_RAW_LD, lane, y);
          if (designated) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              const int f = lane + 32 * i;
              if (f < p.F) p.Y[((size_t)b * p.F + f) * p.t_cap + (t - 1)] = y[i];
            }
          }
        }
        if (!final_stage) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int f = lane + 32 * i;
            if (f < p.F) xrow[f] = y[i];
          }
        }
      } else if (pro == PRO_LN || pro == PRO_LN_RELU) {
        const float* R = rawp + (size_t)b * DEC_RAW_LD;
        float v[NE];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NE; ++i) { v[i] = __ldcg(R + lane + 32 * i); sum += v[i]; }
        sum = warp_sum(sum);
        const float mean = sum / (float)HD;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NE; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
        q = warp_sum(q);
        const float rstd = 1.0f / sqrtf(q / (float)HD + 1e-5f);
#pragma unroll
        for (int i = 0; i < NE; ++i) {
          const int c = lane + 32 * i;
          float o = (v[i] - mean) * rstd * g1[c] + b1[c];
          if (pro == PRO_LN_RELU) o = fmaxf(o, 0.f);
          xrow[c] = o;
        }
      } else {   // PRO_HWY / PRO_ATT
        const float* R = rawp + (size_t)b * DEC_RAW_LD;
        const float* res = p.hist + (size_t)sp->res_hist * hist_buf + ((size_t)b * p.t_cap + t) * HD;
        float h1[NE], h2[NE], xr[NE];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NE; ++i) {
          const int c = lane + 32 * i;
          h1[i] = __ldcg(R + c);
          h2[i] = __ldcg(R + HD + c);
          xr[i] = __ldcg(res + c);
          s1 += h1[i];
          s2 += h2[i];
        }
        warp_sum2(s1, s2);
        const float m1 = s1 / (float)HD, m2 = s2 / (float)HD;
        float q1 = 0.f, q2 = 0.f;
#pragma unroll
        for (int i = 0; i < NE; ++i) {
          const float d1 = h1[i] - m1, d2 = h2[i] - m2;
          q1 = fmaf(d1, d1, q1);
          q2 = fmaf(d2, d2, q2);
        }
        warp_sum2(q1, q2);
        const float r1 = 1.0f / sqrtf(q1 / (float)HD + 1e-5f);
        const float r2 = 1.0f / sqrtf(q2 / (float)HD + 1e-5f);
        float u[NE];
#pragma unroll
        for (int i = 0; i < NE; ++i) {
          const int c = lane + 32 * i;
          const float a = (h1[i] - m1) * r1 * g1[c] + b1[c];
          const float bb = (h2[i] - m2) * r2 * g2[c] + b2[c];
          const float g = sigmoidf_(a);
          u[i] = g * bb + (1.0f - g) * xr[i];
        }
        if (pro == PRO_HWY) {
#pragma unroll
          for (int i = 0; i < NE; ++i) xrow[lane + 32 * i] = u[i];
        } else {
          // windowed attention, models/TTSModel.py:281-295: logits over [pma, min(pma+2, N-1)];
          // every other character is masked to -2^32 and gets softmax weight exactly 0.
          const int p0 = pma_s[r];
          const int cnt = min(p0 + 2, p.N - 1) - p0 + 1;
          const float* kp = p.Kt + ((size_t)b * p.N + p0) * HD;
          const float* vp = p.Vt + ((size_t)b * p.N + p0) * HD;
          float l0 = 0.f, l1 = 0.f, l2 = 0.f;
#pragma unroll
          for (int i = 0; i < NE; ++i) {
            const int c = lane + 32 * i;
            l0 = fmaf(__ldg(kp + c), u[i], l0);
            if (cnt > 1) l1 = fmaf(__ldg(kp + HD + c), u[i], l1);
            if (cnt > 2) l2 = fmaf(__ldg(kp + 2 * HD + c), u[i], l2);
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
          for (int i = 0; i < NE; ++i) {
            const int c = lane + 32 * i;
            float rr = a0 * __ldg(vp + c);
            if (cnt > 1) rr = fmaf(a1, __ldg(vp + HD + c), rr);
            if (cnt > 2) rr = fmaf(a2, __ldg(vp + 2 * HD + c), rr);
            xrow[c] = rr;             // R
            xrow[HD + c] = u[i];      // Q
          }
          __syncwarp();
          if (lane == 0) {
            pma_s[r] = p0 + best;
            if (designated) {
              float* Ab = p.A + ((size_t)b * p.N + p0) * p.t_cap + t;
              Ab[0] = a0;
              if (cnt > 1) Ab[p.t_cap] = a1;
              if (cnt > 2) Ab[2 * (size_t)p.t_cap] = a2;
              p.pma_traj[(size_t)t * p.B + b] = p0 + best;
              p.pma_state[b] = p0 + best;
            }
          }
        }
      }
      // the designated CTA of the row group publishes the stage input to its history buffer
      if (!final_stage && designated && sp->hist_in >= 0) {
        __syncwarp();
        float* hrow = p.hist + (size_t)sp->hist_in * hist_buf + ((size_t)b * p.t_cap + t) * HD;
#pragma unroll
        for (int i = 0; i < NE; ++i) hrow[lane + 32 * i] = xrow[lane + 32 * i];
      }
    }
    if (final_stage) break;
    PROF_T(3);
    cp_async_wait_all();
    __syncthreads();
    PROF_T(4);

    // ---- 4. GEMV: one column pair x RC rows per warp
    const int col = pair0 + pair_local;
    if (pair_local < ppc && col < half) {
      float4 w0[6], w1[6];
      const float* wa = ws + pair_local * KMAX;
      const float* wb = ws + (PAIRS + pair_local) * KMAX;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int k = lane * 4 + 128 * i;
        if (k < K) {
          w0[i] = *reinterpret_cast<const float4*>(wa + k);
          w1[i] = *reinterpret_cast<const float4*>(wb + k);
        } else {
          w0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          w1[i] = w0[i];
        }
      }
      // the lane that will own output (row r, column oc) after the fold fetches its bias now,
      // so the load's L2 latency hides behind the FMA loop
      const int vi = fold_index<2 * RC>(lane);
      const int r_out = chunk * RC + vi % RC, hsel = vi / RC;
      const int oc = col + hsel * half;
      const bool writer = fold_writer<2 * RC>(lane) && r_out < nrows;
      float bias_v = 0.f;
      if (writer) {
        bias_v = __ldg(sp->bias + oc);
        if (sp->bias_b == 1) bias_v += __ldg(p.s1 + (size_t)(row0 + r_out) * HD + oc);
        else if (sp->bias_b == 2) bias_v += __ldg(p.s2 + (size_t)(row0 + r_out) * HD + oc);
      }
      float acc[2 * RC];
#pragma unroll
      for (int i = 0; i < 2 * RC; ++i) acc[i] = 0.f;
#pragma unroll
      for (int r = 0; r < RC; ++r) {
        const float* xr = xs + (chunk * RC + r) * DEC_XS_LD;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int k = lane * 4 + 128 * i;
          if (k < K) {
            const float4 x = *reinterpret_cast<const float4*>(xr + k);
            acc[r] = dot4(w0[i], x, acc[r]);
            acc[RC + r] = dot4(w1[i], x, acc[RC + r]);
          }
        }
      }
      const float tot = fold_reduce<2 * RC>(acc, lane);
      if (writer)
        __stcg(p.raw + (size_t)s * raw_stage + (size_t)(row0 + r_out) * DEC_RAW_LD + oc, tot + bias_v);
    }
    // xs / ws are rewritten by the next stage's prefetch only after every warp passed this point
    PROF_T(5);
    __syncthreads();
    PROF_T(6);
  }
  if (prof_on) {
#pragma unroll
    for (int i = 0; i < 7; ++i) p.prof[(size_t)blockIdx.x * 8 + i] = prof_acc[i];
    p.prof[(size_t)blockIdx.x * 8 + 7] = total;
  }
#undef PROF_T
}

template <int ROWS, int RC>
int launch_rows(const DecParams& p, int grid, cudaStream_t s) {
  constexpr int PAIRS = NW / (ROWS / RC);
  constexpr size_t smem = ((size_t)ROWS * DEC_XS_LD + 2 * (size_t)PAIRS * KMAX + 5 * HD) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    SSV_CUDA(cudaFuncSetAttribute(decode_kernel<ROWS, RC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  DecParams pl = p;
  void* args[] = {&pl};
  SSV_CUDA(cudaLaunchCooperativeKernel((void*)decode_kernel<ROWS, RC>, dim3(grid), dim3(NT), args, smem, s));
  ++g_launches;
  return kOk;
}

}  // namespace

int decode_select_impl(int sm_count) {
  const char* e = getenv("SSV_DECODE_IMPL");
  if (e && e[0] == 'g') return DEC_IMPL_GRID;
  if (e && e[0] == 'c') return decode_cluster_capacity() > 0 ? DEC_IMPL_CLUSTER : DEC_IMPL_GRID;
  return decode_ws_supported(sm_count) ? DEC_IMPL_WS : DEC_IMPL_GRID;
}

int launch_decode(const DecParams& p, int sm_count, int impl, cudaStream_t s) {
  if (impl == DEC_IMPL_WS) return launch_decode_ws(p, s);
  if (impl == DEC_IMPL_CLUSTER) return launch_decode_cluster(p, s);
  SSV_CHECK(p.H == HD, "decode: hidden_dim must be %d", HD);
  SSV_CHECK(p.F <= 96 && p.F % 4 == 0, "decode: freq_bins must be <= 96 and a multiple of 4");
  SSV_CHECK(p.B >= 1 && p.n_steps >= 1, "decode: empty launch");
  SSV_CHECK(sm_count <= DEC_MAX_GRID, "decode: more than %d SMs", DEC_MAX_GRID);
  // per-CTA barrier flags (generation numbers restart at 1 every launch)
  SSV_CUDA(cudaMemsetAsync(p.bar_counter, 0, sizeof(int) * DEC_MAX_GRID, s));
  DecParams q = p;
  if (p.B == 1) { q.RG = 1; return launch_rows<1, 1>(q, sm_count, s); }
  if (p.B <= 4) { q.RG = 1; return launch_rows<4, 4>(q, sm_count, s); }
  q.RG = (p.B + 15) / 16;
  SSV_CHECK(q.RG <= sm_count, "decode: batch too large for one launch");
  return launch_rows<16, 8>(q, sm_count, s);
}

}  // namespace ssv
