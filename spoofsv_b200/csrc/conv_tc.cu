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

#include <cstdlib>
#include <cstring>
#include <vector>

namespace ssv {

namespace {

constexpr int NT = 384;                   // warp 0: TMA, warp 1: MMA, warp 2: TMEM alloc, warps 4-11: epilogue
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
// With 64-byte rows (32-element k-blocks) the same with SWIZZLE_64B: SBO = 512 B, layout_type = 4.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, bool sw64) {
  const uint64_t hi = sw64 ? ((32ull << 32) | (4ull << 61)) : ((64ull << 32) | (2ull << 61));
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (1ull << 46) | hi;
}
// cute::UMMA::InstrDescriptor, kind::f16: D=F32, A=B=BF16, both K-major, M=128.
__device__ __forceinline__ uint32_t umma_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
// 32 lanes x 16 consecutive fp32 columns; thread i of the warp gets lane (base_lane + i).
// Issue only: the registers are valid after tmem_wait16 (which names them, so no use can be hoisted above it).
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

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
__device__ __forceinline__ void st_peer_f32x4(float4* local_ptr, uint32_t peer, float4 v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_ptr)), "r"(peer));
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(remote), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

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

// Write 16 consecutive output columns of one row into the padded smem staging tile (row stride rs bytes;
// rs/16 is odd, so the 8 lanes of a quarter-warp hit 8 different 16-byte bank groups).
__device__ __forceinline__ void stage16(uint8_t* out_s, uint32_t rs, int row, int col_local, bool fp32, const float (&o)[16]) {
  if (fp32) {
    float4* p = reinterpret_cast<float4*>(out_s + (size_t)row * rs + (size_t)col_local * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
  } else {
    bf16x8 a, b;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a.v[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
      b.v[i] = __floats2bfloat162_rn(o[8 + 2 * i], o[8 + 2 * i + 1]);
    }
    bf16x8* p = reinterpret_cast<bf16x8*>(out_s + (size_t)row * rs + (size_t)col_local * 2);
    p[0] = a;
    p[1] = b;
  }
}

__global__ void __launch_bounds__(NT, 1) conv_tc_kernel(const __grid_constant__ ConvTcArgs a, int* err) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  // (offset arithmetic on the shared array, not a uintptr_t round trip, so loads stay LDS/STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int NL = a.n0 + a.n1;
  const int BK = a.bk;                               // 64 (SWIZZLE_128B rows) or 32 (SWIZZLE_64B rows)
  const uint32_t A_STAGE_BYTES = TC_BM * BK * 2;
  const uint32_t b_bytes = (uint32_t)NL * (BK * 2);
  const uint32_t stage_bytes = A_STAGE_BYTES + b_bytes;
  uint8_t* tiles = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)a.nstages * stage_bytes);
  uint64_t* empty_bar = full_bar + a.nstages;
  uint64_t* accum_bar = empty_bar + a.nstages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  // [NL] {bias, LN weight, LN bias, 0} per local column (16-byte aligned)
  float4* prm_s = reinterpret_cast<float4*>(smem + (((size_t)a.nstages * stage_bytes + (2 * a.nstages + 1) * 8 + 8 + 15) & ~size_t(15)));
  float4* part_s = prm_s + NL;                                    // [2][128] per-row partial sums of the two column halves
  float4* stat_s = part_s + 2 * TC_BM;                            // [128] the peer CTA's sums (cluster_n == 2)
  uint8_t* res_s = reinterpret_cast<uint8_t*>(stat_s + TC_BM);    // [128][n0*2 + 16] residual tile (highway, staged)
  uint8_t* out_s = tiles;                                         // output staging aliases the drained pipeline stages

  const long long t_begin = clock64();
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
    float g = 1.f, be = 0.f;
    if (a.epi == EPI_HIGHWAY) {
      const int c = j < a.n0 ? gc : gc - a.n_real;        // channel index inside LN1 / LN2
      g = j < a.n0 ? a.g1[c] : a.g2[c];
      be = j < a.n0 ? a.b1[c] : a.b2[c];
    } else if (a.epi != EPI_NONE && gc < a.n_real) {
      g = a.g1[gc];
      be = a.b1[gc];
    }
    prm_s[j] = make_float4(a.bias[gc], g, be, 0.f);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (a.cluster_n > 1) cluster_sync_all();     // peer CTA is running before any DSMEM store targets it
  long long* prof = a.prof ? a.prof + (size_t)blockIdx.x * 8 : nullptr;
  const long long t_setup = clock64();
  if (prof && threadIdx.x == 0) prof[0] = t_setup - t_begin;

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
        const int c0 = (kb - j * a.kb_per_tap) * BK;
        tma_load_3d(As, &a.tmA, full_bar + st, c0, t0 + (tap_base + j) * a.dil, b);
        tma_load_2d(Bs, &a.tmB0, full_bar + st, kb * BK, w0);
        if (a.n1) tma_load_2d(Bs + (size_t)a.n0 * (BK * 2), &a.tmB1, full_bar + st, kb * BK, w1);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t id0 = umma_idesc(a.n0);
      const uint32_t id1 = umma_idesc(a.n1 ? a.n1 : 16);
      bool ok = true;
      long long mma_wait_cycles = 0;
      for (int kb = 0; kb < nk && ok; ++kb) {
        const int st = kb % a.nstages;
        const uint32_t ph = (uint32_t)(kb / a.nstages) & 1u;
        const long long tw = clock64();
        ok = mbar_wait(full_bar + st, ph, err);
        mma_wait_cycles += clock64() - tw;
        if (!ok) break;
        tc_fence_after();
        const uint32_t As = smem_u32(tiles + (size_t)st * stage_bytes);
        const uint32_t Bs = As + A_STAGE_BYTES;
        const bool sw64 = BK == 32;
        for (int k = 0; k < BK / 16; ++k) {
          const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
          const uint64_t ad = umma_desc(As + k * 32, sw64);
          umma_bf16(tmem_base, ad, umma_desc(Bs + k * 32, sw64), id0, acc);
          if (a.n1) umma_bf16(tmem_base + (uint32_t)a.n0, ad, umma_desc(Bs + (uint32_t)a.n0 * (BK * 2) + k * 32, sw64), id1, acc);
        }
        umma_commit(empty_bar + st);          // frees the smem slot once these MMAs retire
      }
      umma_commit(accum_bar);                 // accumulator complete
      if (prof) { prof[1] = mma_wait_cycles; prof[2] = clock64() - t_setup; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> LN / gate -> global =====================
    // 8 warps: warp (q, hsel) reads TMEM lane quadrant q (rows q*32..) and column half hsel.
    const int q = warp & 3, hsel = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int t = t0 + row;
    const int ew = warp - 4;
    const uint32_t res_rs = (uint32_t)a.n0 * 2 + 16;
    if (a.stage_res) {
      // while the mainloop runs: coalesced copy of the residual rows (my channel slice) into smem
      const int chunks = a.n0 / 8;                              // 16-byte chunks per row
      for (int i = 0; i < TC_BM / 8; ++i) {
        const int r = ew * (TC_BM / 8) + i;
        const int tr = t0 + r;
        const __nv_bfloat16* src = a.Xres + (long)b * a.x_sb + (long)tr * a.x_st + w0;
        for (int ch = lane; ch < chunks; ch += 32) {
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (tr < a.T) v = *reinterpret_cast<const uint4*>(src + ch * 8);
          *reinterpret_cast<uint4*>(res_s + (size_t)r * res_rs + ch * 16) = v;
        }
      }
      epi_bar_sync();
    }
    const bool got = mbar_wait(accum_bar, 0u, err);
    const bool row_ok = t < a.T && got;
    const bool ptime = prof != nullptr && warp == 4 && lane == 0;
    long long t_acc = clock64(), t_p1 = t_acc, t_ex = t_acc;
    if (ptime) prof[3] = t_acc - t_setup;
    tc_fence_after();
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    const long yoff = (long)b * a.y_sb + (long)t * a.y_st;
    const bool f32 = a.out_fp32 != 0;
    const bool hwy = a.epi == EPI_HIGHWAY;
    const bool staged = a.stage_out != 0;
    const uint32_t out_rs = (uint32_t)(hwy ? a.n0 : NL) * (f32 ? 4u : 2u) + 16u;
    // my local column range [c_lo, c_hi): for a highway layer it indexes H1 (H2 is the same range + n0)
    int c_lo, c_hi;
    if (hwy) {
      c_lo = hsel * (a.n0 / 2);
      c_hi = c_lo + a.n0 / 2;
    } else {
      const int split = ((NL / 2 + 15) / 16) * 16;
      c_lo = hsel ? split : 0;
      c_hi = hsel ? NL : split;
    }
    uint32_t ra[16], rb[16], ra2[16], rb2[16];
    float o[16];

    if (a.epi == EPI_NONE) {
      auto emit = [&](int c, const uint32_t (&r)[16]) {
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = __uint_as_float(r[i]) + prm_s[c + i].x;
        const int gc = c < a.n0 ? w0 + c : w1 + (c - a.n0);
        if (staged) stage16(out_s, out_rs, row, c, f32, o);
        else if (row_ok) store16(a.Y, f32, yoff, gc, o, a.y_cols);
      };
      for (int c = c_lo; c < c_hi; c += 32) {
        const bool two = c + 16 < c_hi;
        tmem_ld16_issue(tq + c, ra);
        if (two) tmem_ld16_issue(tq + c + 16, ra2);
        tmem_wait16(ra);
        emit(c, ra);
        if (two) { tmem_wait16(ra2); emit(c + 16, ra2); }
      }
    } else {
      // pass 1: per-row sum and sum of squares over my columns (fp32); 2-4 TMEM loads in flight
      float s1 = 0.f, q1 = 0.f, s2 = 0.f, q2 = 0.f;
      auto acc_hwy = [&](int c, const uint32_t (&r1_)[16], const uint32_t (&r2_)[16]) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float x1 = __uint_as_float(r1_[i]) + prm_s[c + i].x;
          const float x2 = __uint_as_float(r2_[i]) + prm_s[a.n0 + c + i].x;
          s1 += x1; q1 = fmaf(x1, x1, q1);
          s2 += x2; q2 = fmaf(x2, x2, q2);
        }
      };
      auto acc_ln = [&](int c, const uint32_t (&r1_)[16]) {
        const int gc = c < a.n0 ? w0 + c : w1 + (c - a.n0);
        const int nv = a.n_real - gc;                 // padded columns of the 513-bin heads do not count
        if (nv >= 16) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float x1 = __uint_as_float(r1_[i]) + prm_s[c + i].x;
            s1 += x1; q1 = fmaf(x1, x1, q1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float x1 = i < nv ? __uint_as_float(r1_[i]) + prm_s[c + i].x : 0.f;
            s1 += x1; q1 = fmaf(x1, x1, q1);
          }
        }
      };
      for (int c = c_lo; c < c_hi; c += 32) {
        const bool two = c + 16 < c_hi;
        tmem_ld16_issue(tq + c, ra);
        if (hwy) tmem_ld16_issue(tq + a.n0 + c, rb);
        if (two) {
          tmem_ld16_issue(tq + c + 16, ra2);
          if (hwy) tmem_ld16_issue(tq + a.n0 + c + 16, rb2);
        }
        tmem_wait16(ra);
        if (hwy) { tmem_wait16(rb); acc_hwy(c, ra, rb); } else acc_ln(c, ra);
        if (two) {
          tmem_wait16(ra2);
          if (hwy) { tmem_wait16(rb2); acc_hwy(c + 16, ra2, rb2); } else acc_ln(c + 16, ra2);
        }
      }
      // combine the two column halves of this CTA, then (cluster split) the two CTAs through DSMEM
      t_p1 = clock64();
      part_s[hsel * TC_BM + row] = make_float4(s1, q1, s2, q2);
      epi_bar_sync();
      {
        const float4 p4 = part_s[(hsel ^ 1) * TC_BM + row];
        s1 += p4.x; q1 += p4.y; s2 += p4.z; q2 += p4.w;
      }
      if (a.cluster_n > 1) {
        if (hsel == 0) st_peer_f32x4(stat_s + row, rank ^ 1u, make_float4(s1, q1, s2, q2));
        cluster_sync_all();
        const float4 p4 = stat_s[row];
        s1 += p4.x; q1 += p4.y; s2 += p4.z; q2 += p4.w;
      }
      t_ex = clock64();
      const float inv_n = 1.0f / (float)a.n_real;
      const float m1 = s1 * inv_n, m2 = s2 * inv_n;
      const float r1 = rsqrtf(fmaxf(q1 * inv_n - m1 * m1, 0.f) + 1e-5f);
      const float r2 = rsqrtf(fmaxf(q2 * inv_n - m2 * m2, 0.f) + 1e-5f);

      if (hwy) {
        const __nv_bfloat16* xres = a.Xres + (long)b * a.x_sb + (long)t * a.x_st;
        auto gate = [&](int c, const uint32_t (&r1_)[16], const uint32_t (&r2_)[16]) {
          const int gc = w0 + c;
          float xr[16];
          if (a.stage_res || row_ok) {
            const bf16x8* xp = a.stage_res ? reinterpret_cast<const bf16x8*>(res_s + (size_t)row * res_rs + (size_t)c * 2)
                                           : reinterpret_cast<const bf16x8*>(xres + gc);
            const bf16x8 xa = xp[0];
            const bf16x8 xb = xp[1];
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
            const float4 p1 = prm_s[c + i], p2 = prm_s[a.n0 + c + i];
            const float A1 = r1 * p1.y, A2 = r2 * p2.y;
            const float h1 = fmaf(__uint_as_float(r1_[i]), A1, fmaf(p1.x - m1, A1, p1.z));
            const float h2 = fmaf(__uint_as_float(r2_[i]), A2, fmaf(p2.x - m2, A2, p2.z));
            const float g = sigmoidf_(h1);
            o[i] = fmaf(g, h2 - xr[i], xr[i]);
          }
          if (staged) stage16(out_s, out_rs, row, c, f32, o);
          else if (row_ok) store16(a.Y, f32, yoff, gc, o, a.y_cols);
        };
        for (int c = c_lo; c < c_hi; c += 32) {
          const bool two = c + 16 < c_hi;
          tmem_ld16_issue(tq + c, ra);
          tmem_ld16_issue(tq + a.n0 + c, rb);
          if (two) {
            tmem_ld16_issue(tq + c + 16, ra2);
            tmem_ld16_issue(tq + a.n0 + c + 16, rb2);
          }
          tmem_wait16(ra);
          tmem_wait16(rb);
          gate(c, ra, rb);
          if (two) {
            tmem_wait16(ra2);
            tmem_wait16(rb2);
            gate(c + 16, ra2, rb2);
          }
        }
      } else {
        const bool relu = a.epi == EPI_LN_RELU, sig = a.epi == EPI_LN_SIGMOID;
        const float lo = relu ? 0.f : -INFINITY;
        auto norm = [&](int c, const uint32_t (&r1_)[16]) {
          const int gc = c < a.n0 ? w0 + c : w1 + (c - a.n0);
          const int nv = a.n_real - gc;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 p1 = prm_s[c + i];
            const float A1 = r1 * p1.y;
            o[i] = fmaxf(fmaf(__uint_as_float(r1_[i]), A1, fmaf(p1.x - m1, A1, p1.z)), lo);
          }
          if (sig) {
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = sigmoidf_(o[i]);
          }
          if (nv < 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = i < nv ? o[i] : 0.f;
          }
          if (staged) stage16(out_s, out_rs, row, c, f32, o);
          else if (row_ok) store16(a.Y, f32, yoff, gc, o, a.y_cols);
        };
        for (int c = c_lo; c < c_hi; c += 32) {
          const bool two = c + 16 < c_hi;
          tmem_ld16_issue(tq + c, ra);
          if (two) tmem_ld16_issue(tq + c + 16, ra2);
          tmem_wait16(ra);
          norm(c, ra);
          if (two) { tmem_wait16(ra2); norm(c + 16, ra2); }
        }
        // zero the tail of the padded row that no CTA's columns cover (e.g. 544..575 of a 576-wide row); with a
        // staged output the coalesced copy-out below does it
        if (!staged && row_ok && hsel == 1 && rank == (uint32_t)(a.cluster_n - 1)) {
          for (int c = a.cluster_n * NL; c < a.y_cols; ++c) {
            if (f32) static_cast<float*>(a.Y)[yoff + c] = 0.f;
            else static_cast<__nv_bfloat16*>(a.Y)[yoff + c] = __float2bfloat16_rn(0.f);
          }
        }
      }
    }
    if (staged) {
      // coalesced copy-out: the CTA's columns [w0, w0 + cols) of every valid row
      epi_bar_sync();
      const uint32_t esz = f32 ? 4u : 2u;
      const int chunks = (int)((out_rs - 16u) / 16u);
      for (int i = 0; i < TC_BM / 8; ++i) {
        const int r = ew * (TC_BM / 8) + i;
        const int tr = t0 + r;
        if (tr >= a.T || !got) continue;
        uint8_t* dst = static_cast<uint8_t*>(a.Y) + ((size_t)((long)b * a.y_sb + (long)tr * a.y_st + w0)) * esz;
        const uint8_t* src = out_s + (size_t)r * out_rs;
        for (int ch = lane; ch < chunks; ch += 32)
          *reinterpret_cast<uint4*>(dst + ch * 16) = *reinterpret_cast<const uint4*>(src + ch * 16);
        // last CTA of the row: zero the padded tail no CTA's columns cover (e.g. 544..575 of a 576-wide row)
        if (!hwy && a.epi != EPI_NONE && rank == (uint32_t)(a.cluster_n - 1)) {
          const int tail_bytes = (a.y_cols - a.cluster_n * NL) * (int)esz;
          for (int o = lane * 16; o < tail_bytes; o += 32 * 16)
            *reinterpret_cast<uint4*>(dst + (size_t)chunks * 16 + o) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
    if (ptime) {
      const long long t_end = clock64();
      prof[4] = t_p1 - t_acc; prof[5] = t_ex - t_p1; prof[6] = t_end - t_ex; prof[7] = t_end - t_begin;
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
             const cuuint32_t* box, bool sw64) {
  auto fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return kCuda;
  }
  cuuint32_t ones[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                  strides_bytes, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
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

long long* tc_prof_buf() {       // SSV_TC_PROF=1: [4096 CTAs][8] cycle counters of the last launch
  static long long* buf = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    if (getenv("SSV_TC_PROF") && cudaMalloc((void**)&buf, sizeof(long long) * 4096 * 8) != cudaSuccess) buf = nullptr;
  }
  return buf;
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

int* tc_err_flag_dev() { return tc_err_flag(); }

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
  // wide tiles (N_local > 272) use 32-element k-blocks so that 5 stages of operands are in flight
  const int BK = (L.n0 + L.n1) > 272 ? 32 : 64;
  const bool sw64 = BK == 32;
  a.bk = BK;
  {
    cuuint64_t dims[3] = {(cuuint64_t)x_ld, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)x_ld * 2, (cuuint64_t)T * x_ld * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, TC_BM, 1};
    SSV_TRY(make_map(&a.tmA, X, 3, dims, strides, box, sw64));
  }
  const int kp = L.k * L.cin_p;
  {
    cuuint64_t dims[2] = {(cuuint64_t)kp, (cuuint64_t)L.rows_pad};
    cuuint64_t strides[1] = {(cuuint64_t)kp * 2};
    cuuint32_t box0[2] = {(cuuint32_t)BK, (cuuint32_t)L.n0};
    SSV_TRY(make_map(&a.tmB0, L.W, 2, dims, strides, box0, sw64));
    if (L.n1) {
      cuuint32_t box1[2] = {(cuuint32_t)BK, (cuuint32_t)L.n1};
      SSV_TRY(make_map(&a.tmB1, L.W, 2, dims, strides, box1, sw64));
    }
  }
  a.T = T; a.B = B;
  a.tiles_per_b = (T + TC_BM - 1) / TC_BM;
  a.kb_per_tap = L.cin_p / BK;
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
  const size_t stage = (size_t)(TC_BM + NL) * BK * 2;
  const bool hwy = epi == EPI_HIGHWAY;
  const size_t res_bytes = hwy ? (size_t)TC_BM * ((size_t)L.n0 * 2 + 16) : 0;      // staged residual tile
  const size_t fixed = 1024 /*align*/ + 256 /*barriers, tmem slot*/ + (size_t)NL * 16 + 3 * TC_BM * 16 + res_bytes;
  int nstages = (int)((222 * 1024 - fixed) / stage);
  if (nstages > 8) nstages = 8;
  SSV_CHECK(nstages >= 2, "conv_tc: tile does not fit shared memory");
  a.nstages = nstages;
  a.stage_res = hwy ? 1 : 0;
  // the output tile is staged in the drained pipeline buffers when it fits there
  const size_t out_bytes = (size_t)TC_BM * ((size_t)(hwy ? L.n0 : NL) * (out_fp32 ? 4 : 2) + 16);
  a.stage_out = out_bytes <= (size_t)nstages * stage ? 1 : 0;
  const size_t smem = fixed + (size_t)nstages * stage;
  static size_t configured = 0;
  if (smem > configured) {
    SSV_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(226 * 1024)));
    configured = 226 * 1024;
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
  long long* prof = tc_prof_buf();
  const int ncta = (int)cfg.gridDim.x;
  a.prof = (prof && ncta <= 4096) ? prof : nullptr;
  if (a.prof) SSV_CUDA(cudaMemsetAsync(prof, 0, sizeof(long long) * 4096 * 8, s));
  SSV_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel, a, err));
  ++g_launches;
  if (a.prof) {
    SSV_CUDA(cudaStreamSynchronize(s));
    std::vector<long long> h((size_t)ncta * 8);
    SSV_CUDA(cudaMemcpy(h.data(), prof, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    const char* nm[8] = {"setup", "mma-wait-full", "mma-loop", "epi-wait-accum", "pass1", "exchange", "pass2", "total"};
    fprintf(stderr, "[tc prof] grid=%d cluster=%d NL=%d K=%d BK=%d stages=%d epi=%d T=%d B=%d :", ncta, L.cluster_n, NL, kp, BK,
            nstages, epi, T, B);
    for (int i = 0; i < 8; ++i) {
      double sum = 0;
      for (int c = 0; c < ncta; ++c) sum += (double)h[(size_t)c * 8 + i];
      fprintf(stderr, " %s=%.0f", nm[i], sum / ncta);
    }
    fprintf(stderr, "\n");
  }
  return kOk;
}

}  // namespace ssv
