// Cluster version of the incremental Text2Mel decode kernel (see decode.cu for the algorithm and the
// reference anchors: generate_test_utterances.py:105-116, models/TTSModel.py:275-300).
//
// One 16-CTA thread-block cluster owns up to 8 utterances; clusters are independent, so there is no
// grid-wide barrier and no co-residency requirement.  Per stage every CTA computes a 1/16 column
// slice of the GEMV for the cluster's rows and then pushes that slice into the shared memory of all
// 16 CTAs with cp.async.bulk (DSMEM): the copy completes on the RECEIVER's mbarrier, so data and
// synchronisation arrive together and the next stage's prologue (LayerNorm / highway gate /
// windowed attention) reads only local shared memory -- no L2 round trip, no polling.
//   stage gs:  arm xbar[gs&1]  ->  cp.async prefetch (weights, old taps, LN params)
//              -> wait xbar[(gs-1)&1]  ->  prologue from raw_s[(gs-1)&1]  ->  GEMV -> slice_s[gs&1]
//              -> 16 bulk copies slice_s -> peer.raw_s[gs&1][my rank]
// Buffers are double-buffered by stage parity; a peer cannot produce stage gs+2 before it has received
// my stage gs+1 slice, which I send after I have finished reading raw_s[gs&1].
#include "decode.cuh"

#include <cstdlib>

namespace ssv {

namespace {

constexpr int NT = 512;
constexpr int NW = NT / 32;
constexpr int CL = 16;               // CTAs per cluster (non-portable size; one cluster per GPC)
constexpr int HD = 256;
constexpr int NE = HD / 32;
constexpr int KMAX = DEC_XS_LD;
constexpr int SLICE = 32;            // floats per row in a CTA's output slice (2 * 16 column pairs)
constexpr long long SPIN_LIMIT = 4000000000LL;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz));
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
  constexpr int rep = 32 / V;
  return (lane & (rep - 1)) == 0;
}

// ---- cluster / mbarrier / DSMEM bulk copy ----
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arm(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// dst / mbar are shared::cluster addresses of the receiving CTA, src is local shared memory.
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t mbar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(mbar_cluster)
               : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// global -> own shared memory through the TMA engine; completes (bytes) on a local mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// Bounded wait on a local mbarrier phase; on timeout flags the abort and returns false.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* abort_flag) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > SPIN_LIMIT) {
      atomicExch(abort_flag, 4);
      return false;
    }
  }
  return true;
}

template <int ROWS>
__global__ void __launch_bounds__(NT, 1) decode_cluster_kernel(const DecParams p) {
  constexpr int RCH = ROWS < 8 ? ROWS : 8;                 // rows per GEMV register tile
  constexpr int NCH = ROWS / RCH;                          // row chunks (2 for the 16-row variant)
  constexpr int V = 4 * RCH;                               // values a warp reduces per chunk (4 columns x RCH rows)
  constexpr int SLICE_FLOATS = ROWS * SLICE;               // one CTA's slice for all rows
  constexpr int RAW_BUF = CL * SLICE_FLOATS;               // all 16 slices of one stage
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                                        // [ROWS][768] stage input rows (taps | current)
  float* ws = xs + ROWS * DEC_XS_LD;                       // [32][768] weights of my 2*ppc local columns
  float* ps = ws + 2 * CL * KMAX;                          // [4][256] LayerNorm parameters of the prologue
  float* raw_s = ps + 4 * HD;                              // [2][16][ROWS][32] raw GEMV outputs of the whole cluster
  float* slice_s = raw_s + 2 * RAW_BUF;                    // [2][ROWS][32] my own slice (bulk-copy source)
  float* part_s = slice_s + 2 * SLICE_FLOATS;              // [2 K-halves][NCH][8 column groups][32] GEMV partials
  __shared__ __align__(8) uint64_t xbar[2];                // all-gather of the previous stage's raw outputs
  __shared__ __align__(8) uint64_t pbar;                   // LayerNorm parameters + old taps of this stage
  __shared__ __align__(8) uint64_t wbar;                   // weights of this stage
  __shared__ int pma_s[ROWS];
  __shared__ int s_bad;
  __shared__ DecStage stab[DEC_STAGES];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank();
  const int cidx = blockIdx.x / CL;
  const int row0 = cidx * ROWS;
  const int nrows = min(ROWS, p.B - row0);
  const bool designated = rank == 0;

  for (int i = tid; i < ROWS * DEC_XS_LD; i += NT) xs[i] = 0.f;
  for (int i = tid; i < (int)(DEC_STAGES * sizeof(DecStage) / sizeof(int)); i += NT)
    reinterpret_cast<int*>(stab)[i] = reinterpret_cast<const int*>(p.stages)[i];
  if (tid < ROWS) {
    int v = 0;
    if (tid < nrows) v = p.pma_in ? (int)p.pma_in[row0 + tid] : p.pma_state[row0 + tid];
    pma_s[tid] = max(0, min(v, p.N - 1));
  }
  if (tid == 0) {
    mbar_init(&xbar[0], 1);
    mbar_init(&xbar[1], 1);
    mbar_init(&pbar, 1);
    mbar_init(&wbar, 1);
    s_bad = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();            // the zero fill above precedes async-proxy writes into xs
  __syncthreads();
  cluster_sync_all();            // every CTA's barriers exist before any peer copy can target them

  const int total = p.n_steps * DEC_STAGES + 1;
  const size_t hist_buf = (size_t)p.B * p.t_cap * HD;
  const uint32_t slice_bytes = SLICE_FLOATS * sizeof(float);

  // raw value of output column `col` of the previous stage (columns split in pairs over the CTAs)
  auto raw_at = [&](const float* rawb, int r, int col, int half_prev, int ppc_prev) -> float {
    const int h = col >= half_prev ? 1 : 0;
    const int c = col - h * half_prev;
    const int owner = c / ppc_prev;
    return rawb[(owner * ROWS + r) * SLICE + (c - owner * ppc_prev) + h * ppc_prev];
  };

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
  int n_prev = 0;                // output width of the previous stage
  for (int gs = 0; gs < total; ++gs) {
    const bool final_stage = gs == total - 1;
    const int step = gs / DEC_STAGES;
    const int s = final_stage ? 0 : gs - step * DEC_STAGES;
    const int t = p.t_start + step;
    const DecStage* sp = stab + s;
    const int n = sp->n, k_seg = sp->k_seg, ntaps = sp->ntaps, dil = sp->dil;
    const int K = ntaps * k_seg;
    const int half = n / 2;
    const int ppc = (half + CL - 1) / CL;             // column pairs per CTA (16 / 8 / 3)
    const int pair0 = (int)rank * ppc;
    const int pb = gs & 1;
    const uint32_t par = (uint32_t)gs & 1u;           // pbar / wbar complete one phase per stage
    // my local column lc in [0, 2*ppc): H1 columns first, then the matching H2 columns
    auto global_col = [&](int lc) { return lc < ppc ? pair0 + lc : half + pair0 + (lc - ppc); };
    auto col_valid = [&](int lc) { return (lc < ppc ? pair0 + lc : pair0 + lc - ppc) < half; };
    PROF_T(0);

    // ---- 0/1. arm the barriers, then let the TMA engine fetch everything that does not depend on the
    //           previous stage: weights -> wbar; LayerNorm parameters + old taps -> pbar
    if (!final_stage) {
      if (tid == 0) {
        int wrows = 0;
        for (int lc = 0; lc < 2 * ppc; ++lc) wrows += col_valid(lc) ? 1 : 0;
        int prm = (sp->g1 ? 1 : 0) + (sp->b1 ? 1 : 0) + (sp->g2 ? 1 : 0) + (sp->b2 ? 1 : 0);
        int taps = 0;
        if (ntaps == 3) {
          if (t - dil >= 0) taps += nrows;
          if (t - 2 * dil >= 0) taps += nrows;
        }
        mbar_arm(&xbar[pb], CL * slice_bytes);
        mbar_arm(&wbar, (uint32_t)wrows * (uint32_t)K * 4u);
        mbar_arm(&pbar, (uint32_t)(prm + taps) * (uint32_t)(HD * 4));
      }
      if (lane == 0) {
        const float* Wg = sp->W;
        for (int lc = warp; lc < 2 * ppc; lc += NW)
          if (col_valid(lc)) bulk_g2s(ws + lc * KMAX, Wg + (size_t)global_col(lc) * K, (uint32_t)K * 4u, &wbar);
        if (warp < 4) {
          const float* src = warp == 0 ? sp->g1 : warp == 1 ? sp->b1 : warp == 2 ? sp->g2 : sp->b2;
          if (src != nullptr) bulk_g2s(ps + warp * HD, src, HD * 4, &pbar);
        }
        if (ntaps == 3) {
          const float* hb = p.hist + (size_t)sp->hist_in * hist_buf;
          for (int rj = warp; rj < nrows * 2; rj += NW) {
            const int r = rj >> 1, j = rj & 1;
            const int tt = t - (2 - j) * dil;
            if (tt >= 0) bulk_g2s(xs + r * DEC_XS_LD + j * HD, hb + ((size_t)(row0 + r) * p.t_cap + tt) * HD, HD * 4, &pbar);
          }
        }
      }
      if (ntaps == 3 && t - 2 * dil < 0) {            // taps before the first frame are zeros
        for (int rj = warp; rj < nrows * 2; rj += NW) {
          const int r = rj >> 1, j = rj & 1;
          if (t - (2 - j) * dil < 0)
            for (int c = lane; c < HD; c += 32) xs[r * DEC_XS_LD + j * HD + c] = 0.f;
        }
      }
    }
    PROF_T(1);

    // ---- 2. wait until all 16 slices of the previous stage have landed in my raw_s
    if (gs > 0) {
      if (!mbar_wait(&xbar[(gs - 1) & 1], (uint32_t)((gs - 1) >> 1) & 1u, p.abort_flag)) s_bad = 1;
    }
    PROF_T(2);
    const float* rawb = raw_s + ((gs - 1) & 1) * RAW_BUF;
    const int half_prev = n_prev / 2;
    const int ppc_prev = (half_prev + CL - 1) / CL;

    // ---- 3. prologue: u_t for my rows -> xs[r][(ntaps-1)*k_seg ...]
    const int pro = final_stage ? PRO_X : sp->pro;
    const float* g1 = ps;
    const float* b1 = ps + HD;
    const float* g2 = ps + 2 * HD;
    const float* b2 = ps + 3 * HD;
    if (!final_stage) {
      if (!mbar_wait(&pbar, par, p.abort_flag)) s_bad = 1;
    }
    if (!final_stage || designated) {
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
            // y_{t-1} = sigmoid(LN5(raw)) of the previous frame's last stage
            float v[3];
            float sm = 0.f;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              const int f = lane + 32 * i;
              v[i] = f < p.F ? raw_at(rawb, r, f, half_prev, ppc_prev) : 0.f;
              sm += v[i];
            }
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
              y[i] = f < p.F ? sigmoidf_((v[i] - mean) * rstd * __ldg(p.fin_g + f) + __ldg(p.fin_b + f)) : 0.f;
              if (designated && f < p.F) p.Y[((size_t)b * p.F + f) * p.t_cap + (t - 1)] = y[i];
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
          float v[NE];
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < NE; ++i) { v[i] = raw_at(rawb, r, lane + 32 * i, half_prev, ppc_prev); sum += v[i]; }
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
        } else {   // PRO_HWY / PRO_ATT: the previous stage is a 3-tap highway conv (512 raw columns, 16 pairs per CTA);
                   // its input u_{t} -- this stage's residual -- still sits in the current-tap slot of xs
          const float* res = xs + r * DEC_XS_LD + 2 * HD;
          float h1[NE], h2[NE], xr[NE];
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int i = 0; i < NE; ++i) {
            const int c = lane + 32 * i;
            const float* cell = rawb + ((c / ppc_prev) * ROWS + r) * SLICE + (c % ppc_prev);
            h1[i] = cell[0];
            h2[i] = cell[ppc_prev];
            xr[i] = res[c];
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
            // windowed attention, models/TTSModel.py:281-295
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
            l0 *= 0.0625f; l1 *= 0.0625f; l2 *= 0.0625f;
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
            __syncwarp();                // every lane has read its residual before the row is overwritten
#pragma unroll
            for (int i = 0; i < NE; ++i) {
              const int c = lane + 32 * i;
              float rr = a0 * __ldg(vp + c);
              if (cnt > 1) rr = fmaf(a1, __ldg(vp + HD + c), rr);
              if (cnt > 2) rr = fmaf(a2, __ldg(vp + 2 * HD + c), rr);
              xrow[c] = rr;
              xrow[HD + c] = u[i];
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
        // rank 0 publishes the stage input to the layer's history in HBM (taps t-d, t-2d of later frames)
        if (!final_stage && designated && sp->hist_in >= 0) {
          __syncwarp();
          float* hrow = p.hist + (size_t)sp->hist_in * hist_buf + ((size_t)b * p.t_cap + t) * HD;
#pragma unroll
          for (int i = 0; i < NE; ++i) hrow[lane + 32 * i] = xrow[lane + 32 * i];
          __threadfence();
        }
      }
    }
    if (final_stage) break;
    PROF_T(3);
    if (!mbar_wait(&wbar, par, p.abort_flag)) s_bad = 1;
    __syncthreads();
    PROF_T(4);
    if (s_bad) break;

    // ---- 4. GEMV: warp (cg, kh) = 4 local columns x half of K x all rows (register tile 4 x RCH)
    const int nkh = (K % 256 == 0) ? 2 : 1;
    const int ncg = (2 * ppc + 3) / 4;
    const int kh = warp % nkh, cg = warp / nkh;
    if (cg < ncg) {
      const int kpart = K / nkh;
      const int kbeg = kh * kpart;
      float4 w[4][3];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int lc = cg * 4 + c;
        const bool okc = lc < 2 * ppc && col_valid(lc);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int k = lane * 4 + 128 * i;
          w[c][i] = (okc && k < kpart) ? *reinterpret_cast<const float4*>(ws + lc * KMAX + kbeg + k)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        float acc[V];
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = 0.f;
#pragma unroll
        for (int r = 0; r < RCH; ++r) {
          const float* xr = xs + (ch * RCH + r) * DEC_XS_LD + kbeg;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int k = lane * 4 + 128 * i;
            if (k < kpart) {
              const float4 x = *reinterpret_cast<const float4*>(xr + k);
#pragma unroll
              for (int c = 0; c < 4; ++c) acc[c * RCH + r] = dot4(w[c][i], x, acc[c * RCH + r]);
            }
          }
        }
        const float tot = fold_reduce<V>(acc, lane);
        if (fold_writer<V>(lane)) part_s[((kh * NCH + ch) * 8 + cg) * 32 + fold_index<V>(lane)] = tot;
      }
    }
    __syncthreads();
    // combine the K halves, add bias (+ hoisted speaker projection) -> my slice [row][lc]
    float* my_slice = slice_s + pb * SLICE_FLOATS;
    for (int idx = tid; idx < ROWS * 2 * ppc; idx += NT) {
      const int row = idx / (2 * ppc), lc = idx - row * (2 * ppc);
      if (!col_valid(lc)) continue;
      const int cgi = lc >> 2, c = lc & 3;
      const int ch = row / RCH, r = row - ch * RCH;
      const int vi = c * RCH + r;
      float v = part_s[((0 * NCH + ch) * 8 + cgi) * 32 + vi];
      if (nkh == 2) v += part_s[((1 * NCH + ch) * 8 + cgi) * 32 + vi];
      const int oc = global_col(lc);
      v += __ldg(sp->bias + oc);
      if (row < nrows) {
        if (sp->bias_b == 1) v += __ldg(p.s1 + (size_t)(row0 + row) * HD + oc);
        else if (sp->bias_b == 2) v += __ldg(p.s2 + (size_t)(row0 + row) * HD + oc);
      }
      my_slice[row * SLICE + lc] = v;
    }
    fence_async_smem();          // my generic-proxy smem writes -> visible to the bulk-copy (async) proxy
    __syncthreads();
    PROF_T(5);

    // ---- 5. all-gather: my slice -> raw_s[pb][rank] of every CTA of the cluster (including me);
    //         warp w issues the copy to CTA w so the 16 issues proceed in parallel
    if (lane == 0) {
      const uint32_t dst_local = smem_u32(raw_s + pb * RAW_BUF + (int)rank * SLICE_FLOATS);
      const uint32_t bar_local = smem_u32(&xbar[pb]);
      dsmem_bulk_copy(map_to_cta(dst_local, (uint32_t)warp), smem_u32(my_slice), slice_bytes,
                      map_to_cta(bar_local, (uint32_t)warp));
    }
    n_prev = n;
    PROF_T(6);
  }
  if (prof_on) {
#pragma unroll
    for (int i = 0; i < 7; ++i) p.prof[(size_t)blockIdx.x * 8 + i] = prof_acc[i];
    p.prof[(size_t)blockIdx.x * 8 + 7] = total;
  }
#undef PROF_T
  // no CTA may exit while a peer's copy still reads its slice or targets its smem
  cluster_sync_all();
}

template <int ROWS>
size_t smem_bytes() {
  constexpr int NCH = ROWS < 8 ? 1 : ROWS / 8;
  return ((size_t)ROWS * DEC_XS_LD + 2 * (size_t)CL * KMAX + 4 * HD + 2 * (size_t)CL * ROWS * SLICE + 2 * (size_t)ROWS * SLICE +
          (size_t)2 * NCH * 8 * 32) * sizeof(float);
}

template <int ROWS>
int configure(int* max_clusters) {
  static int cached = -1;
  if (cached < 0) {
    auto k = decode_cluster_kernel<ROWS>;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<ROWS>()) != cudaSuccess) {
      cudaGetLastError();
      cached = 0;
    } else {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(CL * 8);
      cfg.blockDim = dim3(NT);
      cfg.dynamicSmemBytes = smem_bytes<ROWS>();
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, k, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
      cached = n;
    }
  }
  *max_clusters = cached;
  return kOk;
}

template <int ROWS>
int launch_rows(const DecParams& p, cudaStream_t s) {
  int maxc = 0;
  SSV_TRY(configure<ROWS>(&maxc));
  SSV_CHECK(maxc >= 1, "decode (cluster): a 16-CTA cluster does not fit this device");
  const int nclusters = (p.B + ROWS - 1) / ROWS;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(nclusters * CL));
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem_bytes<ROWS>();
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  SSV_CUDA(cudaLaunchKernelEx(&cfg, decode_cluster_kernel<ROWS>, p));
  ++g_launches;
  return kOk;
}

}  // namespace

// Returns the number of 16-CTA clusters that can be resident at once (0: unsupported).
int decode_cluster_capacity() {
  int n = 0;
  if (configure<8>(&n) != kOk) return 0;
  if (getenv("SSV_DECODE_PROF")) fprintf(stderr, "[decode] resident 16-CTA clusters: %d\n", n);
  return n;
}

int launch_decode_cluster(const DecParams& p, cudaStream_t s) {
  SSV_CHECK(p.H == HD, "decode: hidden_dim must be %d", HD);
  SSV_CHECK(p.F <= 96 && p.F % 4 == 0, "decode: freq_bins must be <= 96 and a multiple of 4");
  SSV_CHECK(p.B >= 1 && p.n_steps >= 1, "decode: empty launch");
  if (p.B == 1) return launch_rows<1>(p, s);
  if (p.B == 2) return launch_rows<2>(p, s);
  if (p.B <= 4) return launch_rows<4>(p, s);
  // 8 rows per cluster while one wave of resident clusters covers the batch, else 16 rows per cluster
  if (p.B <= 8 * decode_cluster_capacity()) return launch_rows<8>(p, s);
  return launch_rows<16>(p, s);
}

}  // namespace ssv
