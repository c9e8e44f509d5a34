// tcgen05 implicit-GEMM conv1d + LayerNorm / highway-gate kernel, BF16 operands, second generation.
//
// The BF16 arm of the SSRN conv stack (models/TTSModel.py:342-362: highwayConv :63-84, 1x1 conv + LayerNorm,
// ConvTranspose1d as a 1x1 GEMM).  Same GEMM view as conv_tc.cu (M = 128 time steps of one utterance, K = taps * Cin
// tap-major, N = output channels, TMA tap loads with out-of-bounds zero fill), different occupancy plan:
//   * a CTA owns 256 accumulator columns (one 128 x 256 fp32 tile = HALF of TMEM) and 90 KB of shared memory, so TWO
//     CTAs are resident per SM: while one runs its epilogue (TMEM -> LayerNorm -> gate -> global), the other one's
//     MMAs keep the tensor pipe busy -- the overlap a persistent kernel gets from TMEM double buffering, without
//     a tile scheduler (conv_tc.cu: 512 columns, one CTA per SM, epilogue serial with the mainloop: 36 % tensor-pipe
//     activity on the d = 512 layers);
//   * wider layers split N over a cluster of 2 (d = 256 highway, 512-column plain layers) or 4 CTAs (d = 512
//     highway), each CTA with matching H1 / H2 column slices; the per-row LayerNorm sums cross the cluster through
//     distributed shared memory;
//   * 4 epilogue warps, one thread per row; the residual is read from global memory one 16-channel chunk ahead.
// What bounds it (ncu, profiles/): the d = 512 layers stream 2.1 GB of operand tiles from L2 per layer (a 128 x 256 tile
// has 43 MAC per operand byte), 7.6 TB/s -- the L2 ceiling -- at 36 % tensor-pipe activity, in this kernel as in
// conv_tc.cu.  Tried: TMA multicast of the activation tile over the N slices and of the weight tile over the M tiles
// of a 2 x 4 / 4 x 2 cluster (every CTA fetches 1/nn + 1/nm of the stage): SSRN 64 x 217 1.65 -> 1.87 / 1.94 ms, the
// cluster-wide stage hand-shake and the 2 KB boxes cost more than the L2 traffic saved (git history).  The lever that
// is left is cta_group::2 (a CTA pair shares the weight tile: 64 MAC per byte).
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

#include <cudaTypedefs.h>

#include <cstdlib>
#include <cstring>
#include <vector>

namespace ssv {

namespace {

using namespace tcx;

constexpr int NT = 384;                   // warp 0: TMA, warp 1: MMA, warp 2: TMEM alloc, warps 4-11: epilogue
constexpr int BM = 128, NL = 256;
constexpr int BKE = 32;                   // bf16 elements per k-block: 64-byte rows, SWIZZLE_64B, two K = 16 MMA steps
constexpr int NSTAGES = 3;
constexpr uint32_t A_BYTES = BM * BKE * 2;            // 8 KB
constexpr uint32_t B_BYTES = NL * BKE * 2;            // 16 KB
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;   // 24 KB
constexpr int HW_LD = 136;                // bf16 per row of the staged 128-column output tile (272 B: odd number of 16-byte units)
constexpr int PL_LD = 264;                // same for 256 columns (528 B)
constexpr int XK_MAX = 576;               // longest K of a layer with an extra column (the 513-bin heads)
constexpr size_t SMEM_BYTES = 1024 + (size_t)NSTAGES * STAGE_BYTES + 256 + (size_t)(NL + 4) * 16 + 4 * BM * 16 + 2 * BM * 16 +
                              2 * BM * 4 + XK_MAX * 2;

struct alignas(64) Args {
  CUtensorMap tmA;        // activations, 3-D (C, T, B) bf16, box (32, 128, 1), SWIZZLE_64B
  CUtensorMap tmB;        // weights, 2-D (K, rows) bf16, box (32, 128)
  int T, B, tiles_per_b;
  int kb_per_tap, ktaps, dil, causal;
  int cluster_n;          // CTAs splitting N
  int w0_base, w0_rank, w1_base, w1_rank;
  int n_real, epi;
  int xcol;               // 1: the layer has n_real = 512 + 1 columns; column 512 is a dot product on the CUDA cores (rank 0)
  int out_fp32;           // plain layers: fp32 output straight from the registers (the last SSRN layer): 1 = channels-last rows,
                          // 2 = channels-first (B, n_real, T) -- the module's output layout: for every channel the 32 lanes
                          // of a warp (32 consecutive time steps) store one 128-byte segment, no transpose kernel afterwards
  const float* bias; const float* g1; const float* b1; const float* g2; const float* b2;
  const __nv_bfloat16* Wx;                        // weight row of the extra column, [K]
  const __nv_bfloat16* Xres; long x_sb, x_st;
  void* Y; long y_sb, y_st;
  long long* prof;        // optional per-CTA cycle counters (SSV_TC_PROF=1)
};

__device__ __forceinline__ uint32_t idesc_bf16() {   // kind::f16: D = F32, A = B = BF16, K-major, N = 256, M = 128
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NL >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive1(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

struct __align__(16) bf16x8 { __nv_bfloat162 v[4]; };
__device__ __forceinline__ float dot8(const bf16x8& a, const bf16x8& b, float acc) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 x = __bfloat1622float2(a.v[i]), y = __bfloat1622float2(b.v[i]);
    acc = fmaf(x.x, y.x, acc);
    acc = fmaf(x.y, y.y, acc);
  }
  return acc;
}

// Epilogue: 8 warps; warp = TMEM lane quadrant (warp % 4) x column half; a thread owns one row and half of the CTA's
// columns (highway: 64 output channels = 64 H1 + the matching 64 H2 columns; plain: 128 columns).  The halves meet
// through shared memory for the LayerNorm sums, the CTAs of a cluster through distributed shared memory.
__global__ void __launch_bounds__(NT, 2) conv_bf16_v2_kernel(const __grid_constant__ Args a, int* err) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* tiles = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)NSTAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + NSTAGES;
  uint64_t* accum_bar = empty_bar + NSTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  float4* prm_s = reinterpret_cast<float4*>(smem + (((size_t)NSTAGES * STAGE_BYTES + (2 * NSTAGES + 1) * 8 + 8 + 15) & ~size_t(15)));
  float4* stat_s = prm_s + (NL + 4);                              // [4 ranks][128] the cluster's sums
  float4* half_s = stat_s + 4 * BM;                               // [2 halves][128] the CTA's column halves
  float* xdot_s = reinterpret_cast<float*>(half_s + 2 * BM);      // [2 halves][128] extra column: partial dot products
  __nv_bfloat16* wx_s = reinterpret_cast<__nv_bfloat16*>(xdot_s + 2 * BM);   // [K] extra column: weight row
  __nv_bfloat16* out_s = reinterpret_cast<__nv_bfloat16*>(tiles); // output staging tile (aliases the drained stages)

  const long long t_begin = a.prof ? clock64() : 0;
  long long* prof = a.prof ? a.prof + (size_t)blockIdx.x * 8 : nullptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = a.cluster_n > 1 ? cluster_rank() : 0u;
  const int tile = blockIdx.x / a.cluster_n;
  const int b = tile / a.tiles_per_b;
  const int t0 = (tile - b * a.tiles_per_b) * BM;
  const int w0 = a.w0_base + (int)rank * a.w0_rank;
  const int w1 = a.w1_base + (int)rank * a.w1_rank;
  const int nk = a.ktaps * a.kb_per_tap;
  const bool hwy = a.epi == EPI_HIGHWAY;
  const bool xcol = a.xcol != 0 && rank == 0;

  if (warp == 0 && lane == 0) { prefetch_tmap(&a.tmA); prefetch_tmap(&a.tmB); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NSTAGES; ++i) {
      mbar_init(full_bar + i, 1);
      mbar_init(empty_bar + i, xcol ? 9 : 1);        // the MMA's commit (+ the 8 epilogue warps that read the extra column's operand)
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(tmem_slot, NL);
  for (int j = threadIdx.x; j < NL + 1; j += NT) {
    if (j == NL) {
      prm_s[NL] = a.xcol ? make_float4(a.bias[512], a.g1[512], a.b1[512], 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    const int gc = j < 128 ? w0 + j : w1 + (j - 128);
    float g = 1.f, be = 0.f;
    if (hwy) {
      const int c = j < 128 ? gc : gc - a.n_real;
      g = j < 128 ? a.g1[c] : a.g2[c];
      be = j < 128 ? a.b1[c] : a.b2[c];
    } else if (a.epi != EPI_NONE) {
      g = a.g1[gc];
      be = a.b1[gc];
    }
    prm_s[j] = make_float4(a.bias[gc], g, be, 0.f);
  }
  if (xcol)
    for (int j = threadIdx.x; j < nk * BKE; j += NT) wx_s[j] = a.Wx[j];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (a.cluster_n > 1) cluster_sync_all();
  const long long t_setup = prof ? clock64() : 0;
  if (prof && threadIdx.x == 0) prof[0] = t_setup - t_begin;

  if (warp == 0) {
    if (lane == 0) {
      const int tap_base = a.causal ? -(a.ktaps - 1) : -((a.ktaps - 1) / 2);
      for (int kb = 0; kb < nk; ++kb) {
        const int st = kb % NSTAGES;
        const uint32_t ph = (uint32_t)(kb / NSTAGES) & 1u;
        if (!mbar_wait(empty_bar + st, ph ^ 1u, err)) break;
        uint8_t* S = tiles + (size_t)st * STAGE_BYTES;
        mbar_expect_tx(full_bar + st, STAGE_BYTES);
        const int j = kb / a.kb_per_tap;
        const int c0 = (kb - j * a.kb_per_tap) * BKE;
        tma_load_3d(S, &a.tmA, full_bar + st, c0, t0 + (tap_base + j) * a.dil, b);
        tma_load_2d(S + A_BYTES, &a.tmB, full_bar + st, kb * BKE, w0);
        tma_load_2d(S + A_BYTES + B_BYTES / 2, &a.tmB, full_bar + st, kb * BKE, w1);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16();
      bool ok = true;
      for (int kb = 0; kb < nk && ok; ++kb) {
        const int st = kb % NSTAGES;
        const uint32_t ph = (uint32_t)(kb / NSTAGES) & 1u;
        ok = mbar_wait(full_bar + st, ph, err);
        if (!ok) break;
        tc_fence_after();
        const uint32_t As = smem_u32(tiles + (size_t)st * STAGE_BYTES);
        const uint32_t Bs = As + A_BYTES;
#pragma unroll
        for (int k = 0; k < BKE / 16; ++k)
          umma_bf16(tmem_base, umma_desc64(As + k * 32), umma_desc64(Bs + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(empty_bar + st);
      }
      umma_commit(accum_bar);
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3, ew = warp - 4, half = ew >> 2;
    const int row = q * 32 + lane;
    const int t = t0 + row;
    const bool row_in = t < a.T;
    const int cw = hwy ? 64 : 128;            // columns of one LayerNorm pass of this thread
    const int cbase = half * cw;              // first TMEM column (highway: of H1; H2 is 128 further)
    const __nv_bfloat16* xres = a.Xres + (long)b * a.x_sb + (long)t * a.x_st + w0 + cbase;
    // the thread's 64 residual channels travel while the mainloop runs.  (One chunk ahead inside the gate pass, as the
    // first version did, stalls every chunk: tcgen05.wait::ld also waits for the thread's outstanding global loads.)
    bf16x8 xrs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xrs[i].v[0] = xrs[i].v[1] = xrs[i].v[2] = xrs[i].v[3] = __floats2bfloat162_rn(0.f, 0.f);
    if (hwy && row_in) {
#pragma unroll
      for (int i = 0; i < 8; ++i) xrs[i] = *reinterpret_cast<const bf16x8*>(xres + 8 * i);
    }
    // extra column (the 513th bin of the SSRN heads): x[row] . w over the operand tiles as they pass through the pipeline
    // -- the epilogue warps are idle during the mainloop, the activations are in shared memory anyway, and the
    // stage is released only after my read (empty_bar counts 9 arrivals on this rank).  A 64-byte row of the
    // SWIZZLE_64B tile keeps its 16-byte chunk c at position c ^ ((row >> 1) & 3); this half takes two of the four.
    float xd = 0.f;
    if (xcol) {
      const int sw = (row >> 1) & 3;
      for (int kb = 0; kb < nk; ++kb) {
        const int st = kb % NSTAGES;
        const uint32_t ph = (uint32_t)(kb / NSTAGES) & 1u;
        if (!mbar_wait(full_bar + st, ph, err)) break;
        const uint8_t* arow = tiles + (size_t)st * STAGE_BYTES + (size_t)row * 64;
        const bf16x8* wk = reinterpret_cast<const bf16x8*>(wx_s + kb * BKE);
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = 2 * half + cc;
          const bf16x8 xa = *reinterpret_cast<const bf16x8*>(arow + ((c ^ sw) << 4));
          xd = dot8(xa, wk[c], xd);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive1(empty_bar + st);
      }
    }
    const bool got = mbar_wait(accum_bar, 0u, err);
    tc_fence_after();
    const bool ptime = prof != nullptr && warp == 4 && lane == 0;
    long long t_acc = 0, t_p1 = 0, t_ex = 0, t_p2 = 0;
    if (ptime) { t_acc = clock64(); prof[1] = t_acc - t_setup; }
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cbase;
    uint32_t r0[16], r1[16];

    float s1 = 0.f, q1 = 0.f, s2 = 0.f, q2 = 0.f, x512 = 0.f;
    if (a.epi != EPI_NONE) {
      for (int c = 0; c < cw; c += 16) {
        tmem_ld16_issue(tq + c, r0);
        if (hwy) tmem_ld16_issue(tq + 128 + c, r1);
        tmem_wait16(r0);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float x = __uint_as_float(r0[i]) + prm_s[cbase + c + i].x;
          s1 += x; q1 = fmaf(x, x, q1);
        }
        if (hwy) {
          tmem_wait16(r1);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float x = __uint_as_float(r1[i]) + prm_s[128 + cbase + c + i].x;
            s2 += x; q2 = fmaf(x, x, q2);
          }
        }
      }
      if (ptime) t_p1 = clock64();
      // the two column halves of the CTA
      half_s[half * BM + row] = make_float4(s1, q1, s2, q2);
      if (xcol) xdot_s[half * BM + row] = xd;
      epi_sync();
      {
        const float4 o = half_s[(half ^ 1) * BM + row];
        s1 += o.x; q1 += o.y; s2 += o.z; q2 += o.w;
      }
      if (xcol) {
        x512 = xdot_s[row] + xdot_s[BM + row] + prm_s[NL].x;
        s1 += x512; q1 = fmaf(x512, x512, q1);
      }
      if (a.cluster_n > 1) {
        const float4 mine = make_float4(s1, q1, s2, q2);
        if (half == 0) {
          stat_s[rank * BM + row] = mine;
          for (uint32_t p = 0; p < (uint32_t)a.cluster_n; ++p)
            if (p != rank) st_peer_f32x4(stat_s + rank * BM + row, p, mine);
        }
        cluster_sync_all();
        s1 = q1 = s2 = q2 = 0.f;
        for (int p = 0; p < a.cluster_n; ++p) {
          const float4 v = stat_s[p * BM + row];
          s1 += v.x; q1 += v.y; s2 += v.z; q2 += v.w;
        }
      }
    }
    if (ptime) t_ex = clock64();
    const float inv_n = 1.0f / (float)a.n_real;
    const float m1 = s1 * inv_n, m2 = s2 * inv_n;
    const float rs1 = rsqrtf(fmaxf(q1 * inv_n - m1 * m1, 0.f) + 1e-5f);
    const float rs2 = rsqrtf(fmaxf(q2 * inv_n - m2 * m2, 0.f) + 1e-5f);

    const int o_ld = hwy ? HW_LD : PL_LD;
    const bool relu = a.epi == EPI_LN_RELU, none = a.epi == EPI_NONE, sigm = a.epi == EPI_LN_SIGMOID;
    float* yf = reinterpret_cast<float*>(a.Y) + (long)b * a.y_sb + (long)t * a.y_st;
#pragma unroll
    for (int c = 0; c < 128; c += 16) {
      if (c >= cw) break;
      tmem_ld16_issue(tq + c, r0);
      if (hwy) tmem_ld16_issue(tq + 128 + c, r1);
      tmem_wait16(r0);
      float o[16];
      if (hwy) {
        tmem_wait16(r1);
        float xr[16];
        const bf16x8 xr0 = xrs[(c >> 3) & 7], xr1 = xrs[((c >> 3) + 1) & 7];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 fa = __bfloat1622float2(xr0.v[i]), fb = __bfloat1622float2(xr1.v[i]);
          xr[2 * i] = fa.x; xr[2 * i + 1] = fa.y; xr[8 + 2 * i] = fb.x; xr[8 + 2 * i + 1] = fb.y;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 p1 = prm_s[cbase + c + i], p2 = prm_s[128 + cbase + c + i];
          const float A1 = rs1 * p1.y, A2 = rs2 * p2.y;
          const float h1 = fmaf(__uint_as_float(r0[i]), A1, fmaf(p1.x - m1, A1, p1.z));
          const float h2 = fmaf(__uint_as_float(r1[i]), A2, fmaf(p2.x - m2, A2, p2.z));
          const float g = sigmoid_fast(h1);
          o[i] = fmaf(g, h2 - xr[i], xr[i]);
        }
      } else if (none) {
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = __uint_as_float(r0[i]) + prm_s[cbase + c + i].x;
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 p1 = prm_s[cbase + c + i];
          const float A1 = rs1 * p1.y;
          const float h1 = fmaf(__uint_as_float(r0[i]), A1, fmaf(p1.x - m1, A1, p1.z));
          o[i] = relu ? fmaxf(h1, 0.f) : (sigm ? sigmoid_fast(h1) : h1);
        }
      }
      if (a.out_fp32 == 2) {
        if (row_in && got) {
          const int j = cbase + c;
          float* dst = reinterpret_cast<float*>(a.Y) + ((long)b * a.n_real + (j < 128 ? w0 + j : w1 + (j - 128))) * a.T + t;
#pragma unroll
          for (int i = 0; i < 16; ++i) dst[(long)i * a.T] = o[i];
        }
      } else if (a.out_fp32) {          // 64 contiguous bytes per thread and chunk: whole sectors
        if (row_in && got) {
          const int j = cbase + c;
          float4* dst = reinterpret_cast<float4*>(yf + (j < 128 ? w0 + j : w1 + (j - 128)));
#pragma unroll
          for (int i = 0; i < 4; ++i) dst[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
        }
      } else {
        bf16x8 oa, ob;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          oa.v[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
          ob.v[i] = __floats2bfloat162_rn(o[8 + 2 * i], o[8 + 2 * i + 1]);
        }
        bf16x8* dst = reinterpret_cast<bf16x8*>(out_s + (size_t)row * o_ld + cbase + c);
        dst[0] = oa;
        dst[1] = ob;
      }
    }
    // the extra column: normalised like the others; bf16 rows also get their zero tail (columns 513..575 are the next
    // layer's K padding)
    if (xcol && half == 0 && row_in && got) {
      const float4 pp = prm_s[NL];
      const float h = (x512 - m1) * rs1 * pp.y + pp.z;
      const float y = relu ? fmaxf(h, 0.f) : (sigm ? sigmoid_fast(h) : h);
      if (a.out_fp32 == 2) {
        reinterpret_cast<float*>(a.Y)[((long)b * a.n_real + 512) * a.T + t] = y;
      } else if (a.out_fp32) {
        yf[512] = y;
      } else {
        __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(a.Y) + (long)b * a.y_sb + (long)t * a.y_st + 512;
        bf16x8 z;
        z.v[0] = z.v[1] = z.v[2] = z.v[3] = __floats2bfloat162_rn(0.f, 0.f);
        bf16x8 first = z;
        first.v[0] = __floats2bfloat162_rn(y, 0.f);
        const int ntail = (int)((a.y_st - 512) / 8);
        for (int i = 0; i < ntail; ++i) reinterpret_cast<bf16x8*>(yb)[i] = i == 0 ? first : z;
      }
    }
    epi_sync();
    if (ptime) t_p2 = clock64();
    // coalesced copy-out: highway: columns [w0, w0 + 128) (16 lanes x 16 B per row, two rows per warp pass);
    // plain: [w0, w0 + 128) | [w1, w1 + 128) (32 lanes per row)
    __nv_bfloat16* Yb = reinterpret_cast<__nv_bfloat16*>(a.Y);
    if (a.out_fp32) {
    } else if (hwy) {
      const int hl = lane >> 4, l16 = lane & 15;
      for (int i = 0; i < BM / 8; i += 2) {
        const int r = ew * (BM / 8) + i + hl;
        const int tr = t0 + r;
        if (tr < a.T && got) {
          const uint4 v = *reinterpret_cast<const uint4*>(out_s + (size_t)r * o_ld + l16 * 8);
          *reinterpret_cast<uint4*>(Yb + (long)b * a.y_sb + (long)tr * a.y_st + w0 + l16 * 8) = v;
        }
      }
    } else {
      for (int i = 0; i < BM / 8; ++i) {
        const int r = ew * (BM / 8) + i;
        const int tr = t0 + r;
        if (tr < a.T && got) {
          const uint4 v = *reinterpret_cast<const uint4*>(out_s + (size_t)r * o_ld + lane * 8);
          const int gc = lane < 16 ? w0 + lane * 8 : w1 + (lane - 16) * 8;
          *reinterpret_cast<uint4*>(Yb + (long)b * a.y_sb + (long)tr * a.y_st + gc) = v;
        }
      }
    }

    if (ptime) {
      const long long t_end = clock64();
      prof[2] = t_p1 - t_acc; prof[3] = t_ex - t_p1; prof[4] = t_p2 - t_ex; prof[5] = t_end - t_p2; prof[6] = t_end - t_begin;
    }
  }

  if (a.cluster_n > 1 && a.epi != EPI_NONE && warp < 4) cluster_sync_all();

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, NL);
  }
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn2() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

int make_map_bf16(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box) {
  auto fn = encode_fn2();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return kCuda;
  }
  cuuint32_t ones[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, ones,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (bf16) failed with CUresult %d", (int)r);
    return kCuda;
  }
  return kOk;
}

int* v2_err_flag() {
  static int* flag = nullptr;
  if (!flag) {
    if (cudaMalloc((void**)&flag, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(flag, 0, sizeof(int));
  }
  return flag;
}

}  // namespace

struct Tc2LaunchImpl {
  Args args;
};
static_assert(sizeof(Tc2LaunchImpl) <= sizeof(((Tc2Launch*)nullptr)->storage), "Tc2Launch::storage too small");

bool tc2_supported(const TcLayer& L, int epi) {
  if (epi == EPI_HIGHWAY) return L.n_real == 256 || L.n_real == 512;
  if (epi == EPI_LN || epi == EPI_LN_RELU || epi == EPI_NONE) return L.rows == 256 || L.rows == 512 || (epi != EPI_NONE && L.rows == 513);
  if (epi == EPI_LN_SIGMOID) return L.rows == 513 || L.rows == 256 || L.rows == 512;
  return false;
}

int tc2_prepare(const TcLayer& L, int epi, int dil, int causal, const __nv_bfloat16* X, int x_ld, int T, int B,
                void* Y, int y_ld, int out_fp32, Tc2Launch* out) {
  SSV_CHECK(tc2_supported(L, epi), "conv_tc2: layer shape (rows %d, epilogue %d) not built", L.rows, epi);
  SSV_CHECK(L.cin_p % BKE == 0 && x_ld >= L.cin_p && x_ld % 8 == 0 && y_ld % 8 == 0, "conv_tc2: bad padding (cin_p %d, ld %d)", L.cin_p, x_ld);
  const bool xcol = L.rows == 513;                 // 512 tensor-core columns + one on the CUDA cores
  SSV_CHECK(!xcol || (L.k == 1 && L.cin_p <= XK_MAX && y_ld >= 520 && L.rows_pad > 512), "conv_tc2: extra-column layer needs k = 1, K <= %d", XK_MAX);
  SSV_CHECK(!out_fp32 || epi != EPI_HIGHWAY, "conv_tc2: fp32 output is built for the plain layers only");
  Args& a = reinterpret_cast<Tc2LaunchImpl*>(out->storage)->args;
  memset(&a, 0, sizeof(a));
  a.T = T; a.B = B;
  a.tiles_per_b = (T + BM - 1) / BM;
  {
    cuuint64_t dims[3] = {(cuuint64_t)x_ld, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)x_ld * 2, (cuuint64_t)T * x_ld * 2};
    cuuint32_t box[3] = {(cuuint32_t)BKE, BM, 1};
    SSV_TRY(make_map_bf16(&a.tmA, X, 3, dims, strides, box));
  }
  const int kp = L.k * L.cin_p;
  {
    cuuint64_t dims[2] = {(cuuint64_t)kp, (cuuint64_t)L.rows_pad};
    cuuint64_t strides[1] = {(cuuint64_t)kp * 2};
    cuuint32_t box[2] = {(cuuint32_t)BKE, 128};
    SSV_TRY(make_map_bf16(&a.tmB, L.W, 2, dims, strides, box));
  }
  a.kb_per_tap = xcol ? (L.cin + BKE - 1) / BKE : L.cin_p / BKE;      // the zero padding past cin is skipped where it is whole k-blocks
  a.ktaps = L.k; a.dil = dil; a.causal = causal;
  if (epi == EPI_HIGHWAY) {
    a.cluster_n = L.n_real / 128;
    a.w0_base = 0; a.w0_rank = 128; a.w1_base = L.n_real; a.w1_rank = 128;
    a.n_real = L.n_real;
  } else {
    a.cluster_n = xcol ? 2 : L.rows / 256;
    a.w0_base = 0; a.w0_rank = 256; a.w1_base = 128; a.w1_rank = 256;
    a.n_real = L.rows;
  }
  a.xcol = xcol ? 1 : 0;
  a.Wx = xcol ? L.W + (size_t)512 * kp : nullptr;
  a.out_fp32 = out_fp32;
  a.epi = epi;
  a.bias = L.bias;
  a.g1 = L.g1; a.b1 = L.b1; a.g2 = L.g2; a.b2 = L.b2;
  a.Xres = X; a.x_sb = (long)T * x_ld; a.x_st = x_ld;
  a.Y = Y; a.y_sb = (long)T * y_ld; a.y_st = y_ld;
  out->n_ctas = B * a.tiles_per_b * a.cluster_n;
  out->cluster_n = a.cluster_n;
  return kOk;
}

static long long* v2_prof_buf() {       // SSV_TC_PROF=1: [4096 CTAs][8] cycle counters of the last launch
  static long long* buf = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    if (getenv("SSV_TC_PROF") && cudaMalloc((void**)&buf, sizeof(long long) * 4096 * 8) != cudaSuccess) buf = nullptr;
  }
  return buf;
}

void tc2_set_output(Tc2Launch* L, void* Y) { reinterpret_cast<Tc2LaunchImpl*>(L->storage)->args.Y = Y; }

int tc2_run(const Tc2Launch& L, cudaStream_t s) {
  int* err = v2_err_flag();
  SSV_CHECK(err != nullptr, "conv_tc2: cannot allocate the error flag");
  long long* prof = v2_prof_buf();
  if (prof) {                           // development aid: phase cycles, mean over the CTAs
    Args a = reinterpret_cast<const Tc2LaunchImpl*>(L.storage)->args;
    a.prof = L.n_ctas <= 4096 ? prof : nullptr;
    SSV_CUDA(cudaMemsetAsync(prof, 0, sizeof(long long) * 4096 * 8, s));
    const size_t smem_p = SMEM_BYTES;
    SSV_CUDA(cudaFuncSetAttribute(conv_bf16_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)L.n_ctas); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem_p; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)L.cluster_n; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, s);
    SSV_CUDA(cudaLaunchKernelEx(&cfg, conv_bf16_v2_kernel, a, err));
    cudaEventRecord(e1, s);
    ++g_launches;
    SSV_CUDA(cudaStreamSynchronize(s));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (a.prof) {
      std::vector<long long> h((size_t)L.n_ctas * 8);
      SSV_CUDA(cudaMemcpy(h.data(), prof, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
      double m[7] = {0};
      for (int c = 0; c < L.n_ctas; ++c)
        for (int i = 0; i < 7; ++i) m[i] += (double)h[(size_t)c * 8 + i] / L.n_ctas;
      fprintf(stderr, "[tc2 prof] grid=%d cluster=%d K=%d taps epi=%d T=%d B=%d %.1f us: setup=%.0f accum-ready=%.0f stats=%.0f exchange=%.0f pass2=%.0f copy-out=%.0f total=%.0f\n",
              L.n_ctas, L.cluster_n, a.ktaps * a.kb_per_tap * BKE, a.epi, a.T, a.B, 1e3 * ms, m[0], m[1], m[2], m[3], m[4], m[5], m[6]);
    }
    return kOk;
  }
  const size_t smem = SMEM_BYTES;
  static bool configured = false;
  if (!configured) {
    SSV_CUDA(cudaFuncSetAttribute(conv_bf16_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)L.n_ctas);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)L.cluster_n;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SSV_CUDA(cudaLaunchKernelEx(&cfg, conv_bf16_v2_kernel, reinterpret_cast<const Tc2LaunchImpl*>(L.storage)->args, err));
  ++g_launches;
  return kOk;
}

int* tc2_err_flag_dev() { return v2_err_flag(); }

int tc2_check_error() {
  int* flag = v2_err_flag();
  if (!flag) return kOk;
  int h = 0;
  SSV_CUDA(cudaMemcpy(&h, flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (h != 0) {
    cudaMemset(flag, 0, sizeof(int));
    set_error("tcgen05 conv kernel (v2): pipeline wait timed out (code %d)", h);
    return kState;
  }
  return kOk;
}

}  // namespace ssv
