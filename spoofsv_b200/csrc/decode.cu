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

namespace ssv {

namespace {

constexpr int NT = 256;
constexpr int NW = NT / 32;
constexpr int HD = 256;              // hidden width (validated on the host)
constexpr int NE = HD / 32;          // elements per lane of a hidden vector
constexpr long long SPIN_LIMIT = 4000000000LL;   // ~2 s of SM clocks

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_u32(unsigned* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
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

// Grid-wide barrier on a monotonically increasing counter.  Returns false on timeout/abort.
__device__ __forceinline__ bool grid_barrier(unsigned* counter, unsigned target, int* abort_flag, int* s_abort) {
  __syncthreads();
  if (threadIdx.x == 0) {
    red_release_u32(counter);
    const long long t0 = clock64();
    int bad = 0;
    while (ld_acquire_u32(counter) < target) {
      if (*reinterpret_cast<volatile int*>(abort_flag) != 0) { bad = 1; break; }
      if (clock64() - t0 > SPIN_LIMIT) {
        atomicExch(abort_flag, 2);
        bad = 1;
        break;
      }
    }
    *s_abort = bad;
  }
  __syncthreads();
  return *s_abort == 0;
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
  // after folding log2(V) levels the remaining low bits are replicas
  constexpr int rep = 32 / V;
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
    y[i] = f < p.F ? sigmoidf_((v[i] - mean) * rstd * p.fin_g[f] + p.fin_b[f]) : 0.f;
  }
}

template <int ROWS>
__global__ void __launch_bounds__(NT, 1) decode_kernel(const DecParams p) {
  extern __shared__ __align__(16) float xs[];   // [ROWS][DEC_XS_LD]
  __shared__ int pma_s[ROWS];
  __shared__ int s_abort;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x;
  const int rg = blockIdx.x % p.RG;
  const int slot_cta = blockIdx.x / p.RG;
  const int Gr = G / p.RG;
  const bool active = slot_cta < Gr;
  const int row0 = rg * ROWS;
  const int nrows = min(ROWS, p.B - row0);
  const bool designated = active && slot_cta == 0;

  for (int i = tid; i < ROWS * DEC_XS_LD; i += NT) xs[i] = 0.f;
  if (tid < ROWS) {
    int v = 0;
    if (tid < nrows) v = p.pma_in ? (int)p.pma_in[row0 + tid] : p.pma_state[row0 + tid];
    pma_s[tid] = max(0, min(v, p.N - 1));
  }
  if (tid == 0) s_abort = 0;
  __syncthreads();

  const int total = p.n_steps * DEC_STAGES + 1;
  const size_t raw_stage = (size_t)p.B * DEC_RAW_LD;
  const size_t hist_buf = (size_t)p.B * p.t_cap * HD;

  for (int gs = 0; gs < total; ++gs) {
    const bool final_stage = gs == total - 1;
    const int step = gs / DEC_STAGES;
    const int s = final_stage ? 0 : gs - step * DEC_STAGES;
    const int t = p.t_start + step;           // for the final stage: t == t_last + 1
    const DecStage st = p.stages[s];
    const int K = st.ntaps * st.k_seg;

    // ---- 1. prefetch (independent of the previous stage): weights -> registers, old taps -> smem
    const int task = slot_cta * NW + warp;
    const bool has_task = active && !final_stage && task < st.n / 2;
    float4 w0[6], w1[6];
    if (has_task) {
      const float* wa = st.W + (size_t)task * K;
      const float* wb = st.W + (size_t)(task + st.n / 2) * K;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int k = lane * 4 + 128 * i;
        if (k < K) {
          w0[i] = __ldg(reinterpret_cast<const float4*>(wa + k));
          w1[i] = __ldg(reinterpret_cast<const float4*>(wb + k));
        } else {
          w0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          w1[i] = w0[i];
        }
      }
    }
    if (active && !final_stage && st.ntaps == 3) {
      const float* hb = p.hist + (size_t)st.hist_in * hist_buf;
      for (int i = tid; i < nrows * 2 * (HD / 4); i += NT) {
        const int c4 = i % (HD / 4);
        const int j = (i / (HD / 4)) & 1;
        const int r = i / (2 * (HD / 4));
        const int tt = t - (2 - j) * st.dil;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tt >= 0) v = __ldcg(reinterpret_cast<const float4*>(hb + ((size_t)(row0 + r) * p.t_cap + tt) * HD) + c4);
        *reinterpret_cast<float4*>(xs + r * DEC_XS_LD + j * HD + c4 * 4) = v;
      }
    }

    // ---- 2. wait for the previous stage of every CTA
    if (gs > 0) {
      if (!grid_barrier(p.bar_counter, (unsigned)G * (unsigned)gs, p.abort_flag, &s_abort)) return;
    }
    if (!active) continue;

    // ---- 3. prologue: u_t for my rows -> xs[r][(ntaps-1)*k_seg ...]
    const int prev = (s + DEC_STAGES - 1) % DEC_STAGES;
    const float* rawp = p.raw + (size_t)prev * raw_stage;
    const int pro = final_stage ? PRO_X : st.pro;
    for (int r = warp; r < nrows; r += NW) {
      const int b = row0 + r;
      float* xrow = xs + r * DEC_XS_LD + (st.ntaps - 1) * st.k_seg;
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
          float o = (v[i] - mean) * rstd * st.g1[c] + st.b1[c];
          if (pro == PRO_LN_RELU) o = fmaxf(o, 0.f);
          xrow[c] = o;
        }
      } else {   // PRO_HWY / PRO_ATT
        const float* R = rawp + (size_t)b * DEC_RAW_LD;
        const float* res = p.hist + (size_t)st.res_hist * hist_buf + ((size_t)b * p.t_cap + t) * HD;
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
          const float a = (h1[i] - m1) * r1 * st.g1[c] + st.b1[c];
          const float bb = (h2[i] - m2) * r2 * st.g2[c] + st.b2[c];
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
      if (!final_stage && designated && st.hist_in >= 0) {
        __syncwarp();
        float* hrow = p.hist + (size_t)st.hist_in * hist_buf + ((size_t)b * p.t_cap + t) * HD;
#pragma unroll
        for (int i = 0; i < NE; ++i) hrow[lane + 32 * i] = xrow[lane + 32 * i];
      }
    }
    if (final_stage) break;
    __syncthreads();

    // ---- 4. GEMV: two output columns per warp over all rows of the group
    if (has_task) {
      float acc[2 * ROWS];
#pragma unroll
      for (int i = 0; i < 2 * ROWS; ++i) acc[i] = 0.f;
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const float* xr = xs + r * DEC_XS_LD;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int k = lane * 4 + 128 * i;
          if (k < K) {
            const float4 x = *reinterpret_cast<const float4*>(xr + k);
            acc[r] = dot4(w0[i], x, acc[r]);
            acc[ROWS + r] = dot4(w1[i], x, acc[ROWS + r]);
          }
        }
      }
      const float tot = fold_reduce<2 * ROWS>(acc, lane);
      const int vi = fold_index<2 * ROWS>(lane);
      const int r = vi % ROWS, half = vi / ROWS;
      if (fold_writer<2 * ROWS>(lane) && r < nrows) {
        const int b = row0 + r;
        const int col = task + half * (st.n / 2);
        float v = tot + st.bias[col];
        if (st.bias_b == 1) v += p.s1[(size_t)b * HD + col];
        else if (st.bias_b == 2) v += p.s2[(size_t)b * HD + col];
        __stcg(p.raw + (size_t)s * raw_stage + (size_t)b * DEC_RAW_LD + col, v);
      }
    }
    // xs is rewritten by the next stage's tap prefetch only after every warp passed this point
    __syncthreads();
  }
}

template <int ROWS>
int launch_rows(const DecParams& p, int grid, cudaStream_t s) {
  constexpr size_t smem = (size_t)ROWS * DEC_XS_LD * sizeof(float);
  static bool configured = false;
  if (!configured) {
    SSV_CUDA(cudaFuncSetAttribute(decode_kernel<ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  DecParams pl = p;
  void* args[] = {&pl};
  SSV_CUDA(cudaLaunchCooperativeKernel((void*)decode_kernel<ROWS>, dim3(grid), dim3(NT), args, smem, s));
  ++g_launches;
  return kOk;
}

}  // namespace

int launch_decode(const DecParams& p, int sm_count, cudaStream_t s) {
  SSV_CHECK(p.H == HD, "decode: hidden_dim must be %d", HD);
  SSV_CHECK(p.F <= 96 && p.F % 4 == 0, "decode: freq_bins must be <= 96 and a multiple of 4");
  SSV_CHECK(p.B >= 1 && p.n_steps >= 1, "decode: empty launch");
  SSV_CUDA(cudaMemsetAsync(p.bar_counter, 0, sizeof(unsigned), s));
  DecParams q = p;
  if (p.B == 1) { q.RG = 1; return launch_rows<1>(q, sm_count, s); }
  if (p.B <= 4) { q.RG = 1; return launch_rows<4>(q, sm_count, s); }
  q.RG = (p.B + 15) / 16;
  SSV_CHECK(q.RG <= sm_count, "decode: batch too large for one launch");
  return launch_rows<16>(q, sm_count, s);
}

}  // namespace ssv
