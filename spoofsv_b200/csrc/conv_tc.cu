// tcgen05 implicit-GEMM conv1d + LayerNorm / highway-gate kernel for sm_100a.
//
// The BF16 arm of the batched conv stacks (SSRN, models/TTSModel.py:342-362; every highwayConv
// :63-84, 1x1 conv + LayerNorm, ConvTranspose1d as a 1x1 GEMM).  GEMM view: M = 128 time steps of
// one utterance, K = taps * Cin (tap-major), N = output channels.
//   * TMA (cp.async.bulk.tensor) stages both operands in 128B-swizzled shared memory.  The
//     activation tensor map is 3-D (C, T, B): a conv tap is the same box at a shifted T
//     coordinate, and the "same"/causal zero padding is TMA out-of-bounds zero fill, so taps
//     never bleed across utterances.
//   * one elected thread issues tcgen05.mma (kind::f16, M=128, N<=256 per instruction) into a
//     TMEM accumulator holding the CTA's complete output rows (<= 512 fp32 columns);
//   * 4 epilogue warps read the accumulator with tcgen05.ld (one thread = one row), so the channel
//     LayerNorm is a per-thread reduction; sigmoid gate + residual are fused; nothing but the
//     finished activation is written.
//   * layers wider than 512 accumulator columns (d=512 highway: N=1024; the 513-bin heads) split N
//     over a 2-CTA cluster: each CTA owns matching H1/H2 column slices and the per-row LayerNorm
//     partial sums are exchanged through distributed shared memory.
#include "conv_tc.cuh"

#include <cudaTypedefs.h>

#include <cstring>

namespace ssv {

namespace {

constexpr int NT = 256;                   // warp 0: TMA, warp 1: MMA, warp 2: TMEM alloc, warps 4-7: epilogue
constexpr int A_STAGE_BYTES = TC_BM * TC_BK * 2;      // 16 KiB
constexpr int TMEM_COLS = 512;
constexpr long long WAIT_LIMIT = 2000000000LL;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must not hang the GPU.  Returns false after ~1 s and flags the error.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* err) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > WAIT_LIMIT) {
      atomicExch(err, 3);
      return false;
    }
  }
  return true;
}

// ---- TMA ----
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ---- tcgen05 ----
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, both operands K-major bf16.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// K-major, SWIZZLE_128B operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start>>4 | LBO=1 | SBO=64 | version=1 | layout_type=2).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor, kind::f16: D=F32, A=B=BF16, both K-major, M=128.
__device__ __forceinline__ uint32_t umma_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
// 32 lanes x 16 consecutive fp32 columns; thread i of the warp gets lane (base_lane + i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- cluster ----
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_peer_f32(float* local_ptr, uint32_t peer, float v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_ptr)), "r"(peer));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

struct __align__(16) bf16x8 { __nv_bfloat162 v[4]; };

// Store 16 consecutive output columns of one row.
__device__ __forceinline__ void store16(void* Y, bool fp32, long row_off, int col, const float (&o)[16], int y_cols) {
  if (col + 16 <= y_cols) {
    if (fp32) {
      float4* p = reinterpret_cast<float4*>(static_cast<float*>(Y) + row_off + col);
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
    } else {
      bf16x8 a, b;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a.v[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
        b.v[i] = __floats2bfloat162_rn(o[8 + 2 * i], o[8 + 2 * i + 1]);
      }
      bf16x8* p = reinterpret_cast<bf16x8*>(static_cast<__nv_bfloat16*>(Y) + row_off + col);
      p[0] = a;
      p[1] = b;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (col + i < y_cols) {
        if (fp32) static_cast<float*>(Y)[row_off + col + i] = o[i];
        else static_cast<__nv_bfloat16*>(Y)[row_off + col + i] = __float2bfloat16_rn(o[i]);
      }
    }
  }
}

__global__ void __launch_bounds__(NT, 1) conv_tc_kernel(const __grid_constant__ ConvTcArgs a, int* err) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NL = a.n0 + a.n1;
  const uint32_t b_bytes = (uint32_t)NL * (TC_BK * 2);
  const uint32_t stage_bytes = A_STAGE_BYTES + b_bytes;
  uint8_t* tiles = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)a.nstages * stage_bytes);
  uint64_t* empty_bar = full_bar + a.nstages;
  uint64_t* accum_bar = empty_bar + a.nstages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 2);        // [NL] bias of my local columns
  float* gam_s = bias_s + NL;                                     // [NL] LayerNorm weight per local column
  float* bet_s = gam_s + NL;                                      // [NL] LayerNorm bias per local column
  float* stat_s = bet_s + NL;                                     // [128][4] peer partial sums (cluster_n == 2)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = a.cluster_n > 1 ? cluster_rank() : 0u;
  const int tile = blockIdx.x / a.cluster_n;
  const int b = tile / a.tiles_per_b;
  const int t0 = (tile - b * a.tiles_per_b) * TC_BM;
  const int w0 = a.w0_base + (int)rank * a.w0_rank;
  const int w1 = a.w1_base + (int)rank * a.w1_rank;
  const int nk = a.ktaps * a.kb_per_tap;

  // ---- one-time setup
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&a.tmA);
    prefetch_tmap(&a.tmB0);
    if (a.n1) prefetch_tmap(&a.tmB1);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < a.nstages; ++i) {
      mbar_init(full_bar + i, 1);
      mbar_init(empty_bar + i, 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  // per-column parameters of my local columns -> smem (broadcast reads in the epilogue)
  for (int j = threadIdx.x; j < NL; j += NT) {
    const int gc = j < a.n0 ? w0 + j : w1 + (j - a.n0);
    bias_s[j] = a.bias[gc];
    float g = 1.f, be = 0.f;
    if (a.epi == EPI_HIGHWAY) {
      const int c = j < a.n0 ? gc : gc - a.n_real;        // channel index inside LN1 / LN2
      g = j < a.n0 ? a.g1[c] : a.g2[c];
      be = j < a.n0 ? a.b1[c] : a.b2[c];
    } else if (a.epi != EPI_NONE && gc < a.n_real) {
      g = a.g1[gc];
      be = a.b1[gc];
    }
    gam_s[j] = g;
    bet_s[j] = be;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (a.cluster_n > 1) cluster_sync_all();     // peer CTA is running before any DSMEM store targets it

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int tap_base = a.causal ? -(a.ktaps - 1) : -((a.ktaps - 1) / 2);
      for (int kb = 0; kb < nk; ++kb) {
        const int st = kb % a.nstages;
        const uint32_t ph = (uint32_t)(kb / a.nstages) & 1u;
        if (!mbar_wait(empty_bar + st, ph ^ 1u, err)) break;
        uint8_t* As = tiles + (size_t)st * stage_bytes;
        uint8_t* Bs = As + A_STAGE_BYTES;
        mbar_expect_tx(full_bar + st, stage_bytes);
        const int j = kb / a.kb_per_tap;
        const int c0 = (kb - j * a.kb_per_tap) * TC_BK;
        tma_load_3d(As, &a.tmA, full_bar + st, c0, t0 + (tap_base + j) * a.dil, b);
        tma_load_2d(Bs, &a.tmB0, full_bar + st, kb * TC_BK, w0);
        if (a.n1) tma_load_2d(Bs + (size_t)a.n0 * (TC_BK * 2), &a.tmB1, full_bar + st, kb * TC_BK, w1);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t id0 = umma_idesc(a.n0);
      const uint32_t id1 = umma_idesc(a.n1 ? a.n1 : 16);
      bool ok = true;
      for (int kb = 0; kb < nk && ok; ++kb) {
        const int st = kb % a.nstages;
        const uint32_t ph = (uint32_t)(kb / a.nstages) & 1u;
        ok = mbar_wait(full_bar + st, ph, err);
        if (!ok) break;
        tc_fence_after();
        const uint32_t As = smem_u32(tiles + (size_t)st * stage_bytes);
        const uint32_t Bs = As + A_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
          const uint64_t ad = umma_desc(As + k * 32);
          umma_bf16(tmem_base, ad, umma_desc(Bs + k * 32), id0, acc);
          if (a.n1) umma_bf16(tmem_base + (uint32_t)a.n0, ad, umma_desc(Bs + (uint32_t)a.n0 * (TC_BK * 2) + k * 32), id1, acc);
        }
        umma_commit(empty_bar + st);          // frees the smem slot once these MMAs retire
      }
      umma_commit(accum_bar);                 // accumulator complete
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> LN / gate -> global =====================
    const int q = warp & 3;                    // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;
    const int t = t0 + row;
    const bool row_ok = t < a.T;
    const bool got = mbar_wait(accum_bar, 0u, err);
    tc_fence_after();
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    float v[16], v2[16], o[16];
    const long yoff = (long)b * a.y_sb + (long)t * a.y_st;
    const bool f32 = a.out_fp32 != 0;

    if (a.epi == EPI_NONE) {
      for (int c = 0; c < NL; c += 16) {
        tmem_ld16(tq + c, v);
        const int gc = c < a.n0 ? w0 + c : w1 + (c - a.n0);
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = v[i] + bias_s[c + i];
        if (row_ok && got) store16(a.Y, f32, yoff, gc, o, a.y_cols);
      }
    } else {
      // pass 1: per-row sums (single pass: sum and sum of squares, fp32)
      float s1 = 0.f, q1 = 0.f, s2 = 0.f, q2 = 0.f;
      for (int c = 0; c < NL; c += 16) {
        tmem_ld16(tq + c, v);
        const bool second = a.epi == EPI_HIGHWAY && c >= a.n0;
        const int gc = c < a.n0 ? w0 + c : w1 + (c - a.n0);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x = v[i] + bias_s[c + i];
          if (a.epi != EPI_HIGHWAY && gc + i >= a.n_real) x = 0.f;     // padded columns of the 513-bin heads
          if (second) { s2 += x; q2 = fmaf(x, x, q2); }
          else { s1 += x; q1 = fmaf(x, x, q1); }
        }
      }
      if (a.cluster_n > 1) {
        // exchange partial sums with the CTA holding the other half of the channels (DSMEM)
        float* mine = stat_s + row * 4;
        st_peer_f32(mine + 0, rank ^ 1u, s1);
        st_peer_f32(mine + 1, rank ^ 1u, q1);
        st_peer_f32(mine + 2, rank ^ 1u, s2);
        st_peer_f32(mine + 3, rank ^ 1u, q2);
        cluster_sync_all();
        s1 += mine[0]; q1 += mine[1]; s2 += mine[2]; q2 += mine[3];
      }
      const float inv_n = 1.0f / (float)a.n_real;
      const float m1 = s1 * inv_n, m2 = s2 * inv_n;
      const float r1 = rsqrtf(fmaxf(q1 * inv_n - m1 * m1, 0.f) + 1e-5f);
      const float r2 = rsqrtf(fmaxf(q2 * inv_n - m2 * m2, 0.f) + 1e-5f);

      if (a.epi == EPI_HIGHWAY) {
        const __nv_bfloat16* xres = a.Xres + (long)b * a.x_sb + (long)t * a.x_st;
        for (int c = 0; c < a.n0; c += 16) {
          tmem_ld16(tq + c, v);
          tmem_ld16(tq + a.n0 + c, v2);
          const int gc = w0 + c;
          float xr[16];
          if (row_ok) {
            const bf16x8 xa = *reinterpret_cast<const bf16x8*>(xres + gc);
            const bf16x8 xb = *reinterpret_cast<const bf16x8*>(xres + gc + 8);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 fa = __bfloat1622float2(xa.v[i]), fb = __bfloat1622float2(xb.v[i]);
              xr[2 * i] = fa.x; xr[2 * i + 1] = fa.y; xr[8 + 2 * i] = fb.x; xr[8 + 2 * i + 1] = fb.y;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) xr[i] = 0.f;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float h1 = (v[i] + bias_s[c + i] - m1) * r1 * gam_s[c + i] + bet_s[c + i];
            const float h2 = (v2[i] + bias_s[a.n0 + c + i] - m2) * r2 * gam_s[a.n0 + c + i] + bet_s[a.n0 + c + i];
            const float g = sigmoidf_(h1);
            o[i] = g * h2 + (1.0f - g) * xr[i];
          }
          if (row_ok && got) store16(a.Y, f32, yoff, gc, o, a.y_cols);
        }
      } else {
        for (int c = 0; c < NL; c += 16) {
          tmem_ld16(tq + c, v);
          const int gc = c < a.n0 ? w0 + c : w1 + (c - a.n0);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float y = (v[i] + bias_s[c + i] - m1) * r1 * gam_s[c + i] + bet_s[c + i];
            if (a.epi == EPI_LN_RELU) y = fmaxf(y, 0.f);
            else if (a.epi == EPI_LN_SIGMOID) y = sigmoidf_(y);
            o[i] = gc + i < a.n_real ? y : 0.f;
          }
          if (row_ok && got) store16(a.Y, f32, yoff, gc, o, a.y_cols);
        }
        // zero the tail of the padded row that no CTA's columns cover (e.g. 544..575 of a 576-wide row)
        if (row_ok && got && rank == (uint32_t)(a.cluster_n - 1)) {
          const int covered = a.cluster_n * NL;
          for (int c = covered + 0; c < a.y_cols; ++c) {
            if (f32) static_cast<float*>(a.Y)[yoff + c] = 0.f;
            else static_cast<__nv_bfloat16*>(a.Y)[yoff + c] = __float2bfloat16_rn(0.f);
          }
        }
      }
    }
  }

  // non-epilogue warps of a cluster-split layer still have to take part in the cluster barrier
  if (a.cluster_n > 1 && a.epi != EPI_NONE && warp < 4) cluster_sync_all();

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
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

int make_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
             const cuuint32_t* box) {
  auto fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return kCuda;
  }
  cuuint32_t ones[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                  strides_bytes, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return kCuda;
  }
  return kOk;
}

__global__ void pack_w_bf16_kernel(const float* __restrict__ w, int n, int cin, int k, int cin_p, int rows_pad,
                                   __nv_bfloat16* __restrict__ dst) {
  const long kp = (long)k * cin_p;
  const long total = (long)rows_pad * kp;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int kk = (int)(i % kp);
    const int row = (int)(i / kp);
    const int j = kk / cin_p, ci = kk % cin_p;
    const float v = (row < n && ci < cin) ? w[((long)row * cin + ci) * k + j] : 0.f;
    dst[i] = __float2bfloat16_rn(v);
  }
}

// ConvTranspose1d weight [cin][cout][2] -> rows j*cout + co, K = cin contiguous.
__global__ void pack_deconv_bf16_kernel(const float* __restrict__ w, int cin, int cout, __nv_bfloat16* __restrict__ dst) {
  const long total = (long)2 * cout * cin;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin);
    const int row = (int)(i / cin);
    const int j = row / cout, co = row % cout;
    dst[i] = __float2bfloat16_rn(w[((long)ci * cout + co) * 2 + j]);
  }
}

__global__ void transpose_in_bf16_kernel(const float* __restrict__ src, long sb, long sc, long st, int C, int T,
                                         __nv_bfloat16* __restrict__ dst, int ld) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    tile[i][tx] = (c < C && t < T) ? src[b * sb + c * sc + t * st] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    if (t < T && c < ld) dst[((long)b * T + t) * ld + c] = __float2bfloat16_rn(tile[tx][i]);
  }
}

__global__ void cast_f2b_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ d, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    d[i] = __float2bfloat16_rn(s[i]);
}
__global__ void cast_b2f_kernel(const __nv_bfloat16* __restrict__ s, float* __restrict__ d, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    d[i] = __bfloat162float(s[i]);
}

int* tc_err_flag() {
  static int* flag = nullptr;
  if (!flag) {
    if (cudaMalloc((void**)&flag, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(flag, 0, sizeof(int));
  }
  return flag;
}

}  // namespace

int tc_pack_weights(const float* w, int n, int cin, int k, int cin_p, int rows_pad, __nv_bfloat16* dst, cudaStream_t s) {
  pack_w_bf16_kernel<<<1024, 256, 0, s>>>(w, n, cin, k, cin_p, rows_pad, dst);
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

int tc_pack_deconv(const float* w, int cin, int cout, __nv_bfloat16* dst, cudaStream_t s) {
  pack_deconv_bf16_kernel<<<512, 256, 0, s>>>(w, cin, cout, dst);
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

int launch_transpose_in_bf16(const float* src, long sb, long sc, long st, int B, int C, int T, __nv_bfloat16* dst,
                             int ld, cudaStream_t s) {
  dim3 grid((T + 31) / 32, (ld + 31) / 32, B), block(32, 8);
  transpose_in_bf16_kernel<<<grid, block, 0, s>>>(src, sb, sc, st, C, T, dst, ld);
  ++g_launches;
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

int launch_cast_f32_to_bf16(const float* src, __nv_bfloat16* dst, size_t n, cudaStream_t s) {
  cast_f2b_kernel<<<1024, 256, 0, s>>>(src, dst, n);
  ++g_launches;
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

int launch_cast_bf16_to_f32(const __nv_bfloat16* src, float* dst, size_t n, cudaStream_t s) {
  cast_b2f_kernel<<<1024, 256, 0, s>>>(src, dst, n);
  ++g_launches;
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

int tc_check_error() {
  int* flag = tc_err_flag();
  if (!flag) return kOk;
  int h = 0;
  SSV_CUDA(cudaMemcpy(&h, flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (h != 0) {
    cudaMemset(flag, 0, sizeof(int));
    set_error("tcgen05 conv kernel: pipeline wait timed out (code %d)", h);
    return kState;
  }
  return kOk;
}

int tc_launch(const TcLayer& L, int epi, int dil, int causal, const __nv_bfloat16* X, int x_ld, int T, int B, void* Y,
              int y_ld, bool out_fp32, cudaStream_t s) {
  SSV_CHECK(L.cin_p % TC_BK == 0 && x_ld >= L.cin_p && x_ld % 8 == 0, "conv_tc: bad K padding (cin_p %d, ld %d)", L.cin_p, x_ld);
  SSV_CHECK(L.n0 > 0 && L.n0 <= 256 && L.n0 % 16 == 0 && L.n1 >= 0 && L.n1 <= 256 && L.n1 % 16 == 0 && L.n0 + L.n1 <= 512,
            "conv_tc: bad accumulator blocks (%d, %d)", L.n0, L.n1);
  SSV_CHECK(L.cluster_n == 1 || L.cluster_n == 2, "conv_tc: cluster_n must be 1 or 2");
  if (epi == EPI_HIGHWAY) SSV_CHECK(L.n0 == L.n1, "conv_tc: highway needs matching H1/H2 blocks");
  int* err = tc_err_flag();
  SSV_CHECK(err != nullptr, "conv_tc: cannot allocate the error flag");

  ConvTcArgs a;
  memset(&a, 0, sizeof(a));
  {
    cuuint64_t dims[3] = {(cuuint64_t)x_ld, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)x_ld * 2, (cuuint64_t)T * x_ld * 2};
    cuuint32_t box[3] = {TC_BK, TC_BM, 1};
    SSV_TRY(make_map(&a.tmA, X, 3, dims, strides, box));
  }
  const int kp = L.k * L.cin_p;
  {
    cuuint64_t dims[2] = {(cuuint64_t)kp, (cuuint64_t)L.rows_pad};
    cuuint64_t strides[1] = {(cuuint64_t)kp * 2};
    cuuint32_t box0[2] = {TC_BK, (cuuint32_t)L.n0};
    SSV_TRY(make_map(&a.tmB0, L.W, 2, dims, strides, box0));
    if (L.n1) {
      cuuint32_t box1[2] = {TC_BK, (cuuint32_t)L.n1};
      SSV_TRY(make_map(&a.tmB1, L.W, 2, dims, strides, box1));
    }
  }
  a.T = T; a.B = B;
  a.tiles_per_b = (T + TC_BM - 1) / TC_BM;
  a.kb_per_tap = L.cin_p / TC_BK;
  a.ktaps = L.k; a.dil = dil; a.causal = causal;
  a.n0 = L.n0; a.n1 = L.n1; a.cluster_n = L.cluster_n;
  a.w0_base = L.w0_base; a.w0_rank = L.w0_rank; a.w1_base = L.w1_base; a.w1_rank = L.w1_rank;
  a.n_real = L.n_real;
  a.epi = epi;
  a.bias = L.bias;
  a.g1 = L.g1; a.b1 = L.b1; a.g2 = L.g2; a.b2 = L.b2;
  a.Xres = X; a.x_sb = (long)T * x_ld; a.x_st = x_ld;
  a.Y = Y; a.y_sb = (long)T * y_ld; a.y_st = y_ld;
  a.y_cols = y_ld;
  a.out_fp32 = out_fp32 ? 1 : 0;
  const int NL = L.n0 + L.n1;
  const size_t stage = A_STAGE_BYTES + (size_t)NL * TC_BK * 2;
  const size_t fixed = 1024 /*align*/ + 256 /*barriers, tmem slot*/ + (size_t)NL * 12 + TC_BM * 16;
  int nstages = (int)((200 * 1024 - fixed) / stage);
  if (nstages > 6) nstages = 6;
  SSV_CHECK(nstages >= 2, "conv_tc: tile does not fit shared memory");
  a.nstages = nstages;
  const size_t smem = fixed + (size_t)nstages * stage;
  static size_t configured = 0;
  if (smem > configured) {
    SSV_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
    configured = 220 * 1024;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(B * a.tiles_per_b * L.cluster_n));
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
  SSV_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel, a, err));
  ++g_launches;
  return kOk;
}

}  // namespace ssv
