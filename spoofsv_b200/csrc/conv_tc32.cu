// FP32-accurate tensor-core arm of the batched conv stacks: tcgen05 (kind::tf32) implicit-GEMM conv1d with split
// operands ("3xTF32") + the fused LayerNorm / highway-gate epilogue.
//
// TextEnc (models/TTSModel.py:126-140) has to stay inside the FP32 bar (1e-4 max-abs on the spectrograms, identical
// alignments): its K / V feed 217 argmax decisions per utterance.  Plain TF32 / BF16 operands are 1e-3 off.
// Here every fp32 operand x is carried as two TF32 numbers, hi = tf32(x) and lo = tf32(x - hi) (22 mantissa bits
// together), and a product is three MMAs,
//      D0 += A_hi B_hi            D1 += A_hi B_lo + A_lo B_hi            (A_lo B_lo ~ 2^-22: dropped)
// with the large term and the two corrections in SEPARATE fp32 TMEM accumulators, so the corrections are not
// truncated against the 2^11 times larger running sum; the epilogue adds D0 + D1 in fp32.
//
// GEMM view as in conv_tc.cu: M = 128 rows of time steps (T <= 64: two utterances of 64 rows, so the 58-character
// TextEnc batch fills 91 % of a tile instead of 45 %), K = taps * Cin (tap-major), N = output channels.  TMA stages
// both halves of both operands in 64-byte-swizzled shared memory (k-blocks of 16 fp32; a conv tap is the same box at
// a shifted T coordinate, zero padding is TMA out-of-bounds fill).  A CTA owns 256 accumulator columns (D0 and D1
// fill the 512 TMEM columns); wider layers split N over a cluster of 2 or 4 CTAs, each with matching H1 / H2 column
// slices, and the per-row LayerNorm partial sums cross the cluster through distributed shared memory.
// The output is written as the (hi, lo) pair the next layer's TMA reads, or as plain fp32 for the last layer.
#include "conv_tc32.cuh"
#include "tc_ptx.cuh"

#include <cudaTypedefs.h>

#include <cstdlib>
#include <cstring>
#include <vector>

namespace ssv {

namespace {

constexpr int NT = 384;                   // warp 0: TMA, warp 1: MMA, warp 2: TMEM alloc, warps 4-11: epilogue
constexpr int TMEM_COLS = 512;
constexpr uint32_t A_HALF_BYTES = T32_BM * T32_BK * 4;          // 8 KB: one of (hi, lo) of the activation tile
constexpr uint32_t B_HALF_BYTES = T32_NL * T32_BK * 4;          // 16 KB: one of (hi, lo) of the weight tile
constexpr uint32_t STAGE_BYTES = 2 * A_HALF_BYTES + 2 * B_HALF_BYTES;   // 48 KB
constexpr int RES_LD = 132;               // floats per row of the staged residual / highway output tile (odd number of 16-byte units)
constexpr int OUT_LD = 260;               // floats per row of the staged 256-column output tile

using namespace tcx;

// cute::UMMA::InstrDescriptor, kind::tf32: D = F32 (bit 4), A = B = TF32 (2 at bits 7 and 10), both K-major, N = 256, M = 128.
__device__ __forceinline__ uint32_t umma_idesc_tf32() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(T32_NL >> 3) << 17) | ((uint32_t)(T32_BM >> 4) << 24);
}

// MUFU-based (2 ulp), as the decode kernel's gate: the epilogue is instruction-bound on the IEEE slow paths otherwise
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
// round to nearest TF32 (10 explicit mantissa bits), kept in an fp32 container
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
  hi = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
  lo = make_float4(to_tf32(v.x - hi.x), to_tf32(v.y - hi.y), to_tf32(v.z - hi.z), to_tf32(v.w - hi.w));
}

__global__ void __launch_bounds__(NT, 1) conv_tf32x3_kernel(const __grid_constant__ ConvTf32Args a, int* err) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* tiles = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)a.nstages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + a.nstages;
  uint64_t* accum_bar = empty_bar + a.nstages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);
  float4* prm_s = reinterpret_cast<float4*>(smem + (((size_t)a.nstages * STAGE_BYTES + (2 * a.nstages + 1) * 8 + 8 + 15) & ~size_t(15)));
  float4* part_s = prm_s + T32_NL;                                // [2][128] per-row partial sums of the two column halves
  float4* stat_s = part_s + 2 * T32_BM;                           // [4 ranks][128] the cluster's sums
  float* out_s = reinterpret_cast<float*>(tiles);                 // output staging tile (aliases the drained pipeline stages)

  const long long t_begin = clock64();
  long long* prof = a.prof ? a.prof + (size_t)blockIdx.x * 8 : nullptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = a.cluster_n > 1 ? cluster_rank() : 0u;
  int tile = blockIdx.x / a.cluster_n;
  int chunk = 0;
  if (a.ksplit > 1) { chunk = tile % a.ksplit; tile /= a.ksplit; }
  int b0, t0;
  if (a.utt_per_tile > 1) { b0 = tile * a.utt_per_tile; t0 = 0; }
  else { b0 = tile / a.tiles_per_b; t0 = (tile - b0 * a.tiles_per_b) * T32_BM; }
  const int w0 = a.w0_base + (int)rank * a.w0_rank;
  const int w1 = a.w1_base + (int)rank * a.w1_rank;
  const int nk = a.ktaps * a.kb_per_tap;
  const bool hwy = a.epi == EPI_HIGHWAY, none = a.epi == EPI_NONE;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&a.tmAh); prefetch_tmap(&a.tmAl); prefetch_tmap(&a.tmBh); prefetch_tmap(&a.tmBl);
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
  // per-column parameters of my 256 local columns -> smem: {bias, LN weight, LN bias, 0}
  for (int j = threadIdx.x; j < T32_NL; j += NT) {
    const int gc = j < 128 ? w0 + j : w1 + (j - 128);
    float g, be;
    if (hwy) {
      const int c = j < 128 ? gc : gc - a.n_real;          // channel inside LN1 / LN2
      g = j < 128 ? a.g1[c] : a.g2[c];
      be = j < 128 ? a.b1[c] : a.b2[c];
    } else if (none) {
      g = 1.f;
      be = 0.f;
    } else if (gc < a.n_cols) {
      g = a.g1[gc];
      be = a.b1[gc];
    } else {                                               // padding column of a layer whose width does not fill the cluster
      g = 0.f;
      be = 0.f;
    }
    prm_s[j] = make_float4(gc < a.n_cols && chunk == 0 ? a.bias[gc] : 0.f, g, be, 0.f);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (a.cluster_n > 1) cluster_sync_all();     // every CTA of the cluster is running before a DSMEM store targets it
  const long long t_setup = clock64();
  if (prof && threadIdx.x == 0) prof[0] = t_setup - t_begin;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // causal: 0 = centred taps, 1 = taps t-(k-1)d .. t, 2 = taps t .. t+(k-1)d (the dgrad of a causal conv)
      const int tap_base = a.causal == 1 ? -(a.ktaps - 1) : (a.causal == 2 ? 0 : -((a.ktaps - 1) / 2));
      for (int kb = 0; kb < nk; ++kb) {
        const int st = kb % a.nstages;
        const uint32_t ph = (uint32_t)(kb / a.nstages) & 1u;
        if (!mbar_wait(empty_bar + st, ph ^ 1u, err)) break;
        uint8_t* S = tiles + (size_t)st * STAGE_BYTES;
        mbar_expect_tx(full_bar + st, STAGE_BYTES);
        const int j = kb / a.kb_per_tap;
        const int c0 = (chunk * a.kb_per_tap + kb - j * a.kb_per_tap) * T32_BK;
        const int tc = t0 + (tap_base + j) * a.dil;
        tma_load_3d(S, &a.tmAh, full_bar + st, c0, tc, b0);
        tma_load_3d(S + A_HALF_BYTES, &a.tmAl, full_bar + st, c0, tc, b0);
        uint8_t* Bh = S + 2 * A_HALF_BYTES;
        uint8_t* Bl = Bh + B_HALF_BYTES;
        const int wk = j * a.w_tap_stride + c0 + b0 * a.w_k_per_b;
        tma_load_2d(Bh, &a.tmBh, full_bar + st, wk, w0);
        tma_load_2d(Bh + B_HALF_BYTES / 2, &a.tmBh, full_bar + st, wk, w1);
        tma_load_2d(Bl, &a.tmBl, full_bar + st, wk, w0);
        tma_load_2d(Bl + B_HALF_BYTES / 2, &a.tmBl, full_bar + st, wk, w1);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32();
      const uint32_t d_main = tmem_base, d_corr = tmem_base + (uint32_t)T32_NL;
      bool ok = true;
      long long wait_cycles = 0;
      for (int kb = 0; kb < nk && ok; ++kb) {
        const int st = kb % a.nstages;
        const uint32_t ph = (uint32_t)(kb / a.nstages) & 1u;
        const long long tw = prof ? clock64() : 0;
        ok = mbar_wait(full_bar + st, ph, err);
        if (prof) wait_cycles += clock64() - tw;
        if (!ok) break;
        tc_fence_after();
        const uint32_t Ah = smem_u32(tiles + (size_t)st * STAGE_BYTES);
        const uint32_t Al = Ah + A_HALF_BYTES;
        const uint32_t Bh = Al + A_HALF_BYTES;
        const uint32_t Bl = Bh + B_HALF_BYTES;
#pragma unroll
        for (int k = 0; k < T32_BK / 8; ++k) {
          const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
          const uint64_t ah = umma_desc64(Ah + k * 32), al = umma_desc64(Al + k * 32);
          const uint64_t bh = umma_desc64(Bh + k * 32), bl = umma_desc64(Bl + k * 32);
          umma_tf32(d_main, ah, bh, idesc, acc);
          umma_tf32(d_corr, ah, bl, idesc, acc);
          umma_tf32(d_corr, al, bh, idesc, 1u);
        }
        umma_commit(empty_bar + st);          // frees the smem slot once these MMAs retire
      }
      umma_commit(accum_bar);                 // accumulators complete
      if (prof) { prof[1] = wait_cycles; prof[2] = clock64() - t_setup; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> LN / gate -> smem -> global =====================
    // 8 warps: warp (q, hsel) reads TMEM lane quadrant q (rows q*32 ..) and column half hsel of each 128-column block.
    const int q = warp & 3, hsel = (warp - 4) >> 2, ew = warp - 4;
    const int row = q * 32 + lane;
    auto row_coords = [&](int r, int& bb, int& tt) {
      const int u = r / a.rows_per_utt;
      bb = b0 + u;
      tt = t0 + (r - u * a.rows_per_utt);
      return bb < a.B && tt < a.T;
    };
    // highway: my row's residual (my 64 channels, hi + lo) -> registers while the mainloop runs
    float xres[64];
    if (hwy) {
      int bb, tt;
      const bool ok = row_coords(row, bb, tt);
      const long off = ok ? (long)bb * a.x_sb + (long)tt * a.x_st + w0 + hsel * 64 : 0;
      float4 h4[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) h4[i] = ok ? __ldg(reinterpret_cast<const float4*>(a.Xh + off) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float4 l4 = ok ? __ldg(reinterpret_cast<const float4*>(a.Xl + off) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        xres[4 * i] = h4[i].x + l4.x; xres[4 * i + 1] = h4[i].y + l4.y; xres[4 * i + 2] = h4[i].z + l4.z; xres[4 * i + 3] = h4[i].w + l4.w;
      }
    }
    const bool got = mbar_wait(accum_bar, 0u, err);
    tc_fence_after();
    const bool ptime = prof != nullptr && warp == 4 && lane == 0;
    long long t_acc = clock64(), t_p1 = t_acc, t_ex = t_acc, t_p2 = t_acc;
    if (ptime) prof[3] = t_acc - t_setup;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);

    uint32_t r0[16], r1[16], r2[16], r3[16];
    float s1 = 0.f, q1 = 0.f, s2 = 0.f, q2 = 0.f;
    // pass 1: per-row sums over my columns.  Highway: my 64 channels of H1 (block 0) and of H2 (block 1);
    // plain LayerNorm: my 128 columns (block hsel).
    const int c_lo = hwy ? hsel * 64 : hsel * 128;
    const int c_n = hwy ? 64 : 128;
    for (int c = c_lo; c < c_lo + c_n && !none; c += 16) {
      tmem_ld16_issue(tq + c, r0);
      tmem_ld16_issue(tq + T32_NL + c, r1);
      if (hwy) {
        tmem_ld16_issue(tq + 128 + c, r2);
        tmem_ld16_issue(tq + T32_NL + 128 + c, r3);
      }
      tmem_wait16(r0);
      tmem_wait16(r1);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float x = (__uint_as_float(r0[i]) + __uint_as_float(r1[i])) + prm_s[c + i].x;
        s1 += x; q1 = fmaf(x, x, q1);
      }
      if (hwy) {
        tmem_wait16(r2);
        tmem_wait16(r3);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float x = (__uint_as_float(r2[i]) + __uint_as_float(r3[i])) + prm_s[128 + c + i].x;
          s2 += x; q2 = fmaf(x, x, q2);
        }
      }
    }
    t_p1 = clock64();
    if (!none) {
      part_s[hsel * T32_BM + row] = make_float4(s1, q1, s2, q2);
      epi_bar_sync();
      const float4 lo4 = part_s[row], hi4 = part_s[T32_BM + row];        // fixed order: both halves get identical sums
      s1 = lo4.x + hi4.x; q1 = lo4.y + hi4.y; s2 = lo4.z + hi4.z; q2 = lo4.w + hi4.w;
    }
    if (a.cluster_n > 1 && !none) {
      if (hsel == 0) {
        const float4 mine = make_float4(s1, q1, s2, q2);
        stat_s[rank * T32_BM + row] = mine;
        for (uint32_t p = 0; p < (uint32_t)a.cluster_n; ++p)
          if (p != rank) st_peer_f32x4(stat_s + rank * T32_BM + row, p, mine);
      }
      cluster_sync_all();
      s1 = q1 = s2 = q2 = 0.f;
      for (int p = 0; p < a.cluster_n; ++p) {                              // fixed order: every CTA gets identical sums
        const float4 v = stat_s[p * T32_BM + row];
        s1 += v.x; q1 += v.y; s2 += v.z; q2 += v.w;
      }
    }
    t_ex = clock64();
    const float inv_n = 1.0f / (float)a.n_real;
    const float m1 = s1 * inv_n, m2 = s2 * inv_n;
    const float rs1 = 1.0f / sqrtf(fmaxf(q1 * inv_n - m1 * m1, 0.f) + 1e-5f);
    const float rs2 = 1.0f / sqrtf(fmaxf(q2 * inv_n - m2 * m2, 0.f) + 1e-5f);

    // pass 2: normalise (+ gate with the residual), into the staging tile
    const int o_ld = hwy ? RES_LD : OUT_LD;
    const bool relu = a.epi == EPI_LN_RELU, sigm = a.epi == EPI_LN_SIGMOID;
#pragma unroll
    for (int cc = 0; cc < 128; cc += 16) {
      if (cc >= c_n) break;
      const int c = c_lo + cc;
      tmem_ld16_issue(tq + c, r0);
      tmem_ld16_issue(tq + T32_NL + c, r1);
      if (hwy) {
        tmem_ld16_issue(tq + 128 + c, r2);
        tmem_ld16_issue(tq + T32_NL + 128 + c, r3);
      }
      tmem_wait16(r0);
      tmem_wait16(r1);
      float o[16];
      if (hwy) {
        tmem_wait16(r2);
        tmem_wait16(r3);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 p1 = prm_s[c + i], p2 = prm_s[128 + c + i];
          const float x1 = (__uint_as_float(r0[i]) + __uint_as_float(r1[i])) + p1.x;
          const float x2 = (__uint_as_float(r2[i]) + __uint_as_float(r3[i])) + p2.x;
          const float h1 = (x1 - m1) * rs1 * p1.y + p1.z;
          const float h2 = (x2 - m2) * rs2 * p2.y + p2.z;
          const float g = sigmoid_fast(h1);
          const float xr = xres[(cc & 63) + i];
          o[i] = g * h2 + (1.0f - g) * xr;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 p1 = prm_s[c + i];
          const float x1 = (__uint_as_float(r0[i]) + __uint_as_float(r1[i])) + p1.x;
          const float h1 = (x1 - m1) * rs1 * p1.y + p1.z;
          o[i] = none ? x1 : (relu ? fmaxf(h1, 0.f) : (sigm ? sigmoid_fast(h1) : h1));   // EPI_NONE: the raw conv output (training keeps it)
        }
      }
      float4* dst = reinterpret_cast<float4*>(out_s + (size_t)row * o_ld + c);
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
    }
    t_p2 = clock64();
    epi_bar_sync();
    // coalesced copy-out of my CTA's columns: highway [w0, w0 + 128); plain [w0, w0 + 128) and [w1, w1 + 128)
    const int f4_per_row = hwy ? 32 : 64;
    for (int i = 0; i < T32_BM / 8; i += 4) {
      float4 v[4][2];
      long yoff[4];
      bool okr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = ew * (T32_BM / 8) + i + u;
        int bb, tt;
        okr[u] = row_coords(r, bb, tt) && got;
        yoff[u] = (long)bb * a.y_sb + (long)tt * a.y_st;
        v[u][0] = *reinterpret_cast<const float4*>(out_s + (size_t)r * o_ld + lane * 4);
        if (!hwy) v[u][1] = *reinterpret_cast<const float4*>(out_s + (size_t)r * o_ld + (lane + 32) * 4);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (!okr[u]) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h * 32 >= f4_per_row) break;
          const int gc = h == 0 ? w0 + lane * 4 : w1 + lane * 4;
          if (gc >= a.y_cols) continue;
          if (a.Yl != nullptr) {
            float4 hi, lo;
            split4(v[u][h], hi, lo);
            *reinterpret_cast<float4*>(a.Yh + yoff[u] + gc) = hi;
            *reinterpret_cast<float4*>(a.Yl + yoff[u] + gc) = lo;
          } else {
            *reinterpret_cast<float4*>(a.Yh + (size_t)chunk * a.y_chunk_stride + yoff[u] + gc) = v[u][h];
          }
        }
      }
    }
    if (ptime) {
      const long long t_end = clock64();
      prof[4] = t_p1 - t_acc; prof[5] = t_ex - t_p1; prof[6] = t_p2 - t_ex; prof[7] = t_end - t_begin;
    }
  }

  // non-epilogue warps of a cluster-split layer still have to take part in the cluster barrier
  if (a.cluster_n > 1 && warp < 4 && !none) cluster_sync_all();

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
PFN_cuTensorMapEncodeTiled_v12000 encode_fn32() {
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

int make_map_f32(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                 const cuuint32_t* box) {
  auto fn = encode_fn32();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return kCuda;
  }
  cuuint32_t ones[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, ones,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp32) failed with CUresult %d", (int)r);
    return kCuda;
  }
  return kOk;
}

// [n][cin][k] fp32 -> [n][k * cin_p] (K contiguous, tap-major), split into hi / lo
__global__ void pack_w_tf32_kernel(const float* __restrict__ w, int n, int cin, int k, int cin_p, float* __restrict__ hi,
                                   float* __restrict__ lo) {
  const long kp = (long)k * cin_p;
  const long total = (long)n * kp;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int kk = (int)(i % kp);
    const int row = (int)(i / kp);
    const int j = kk / cin_p, ci = kk % cin_p;
    const float v = ci < cin ? w[((long)row * cin + ci) * k + j] : 0.f;
    const float h = to_tf32(v);
    hi[i] = h;
    lo[i] = to_tf32(v - h);
  }
}

__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float h = to_tf32(v);
    hi[i] = h;
    lo[i] = to_tf32(v - h);
  }
}

long long* tf32_prof_buf() {       // SSV_TC_PROF=1: [4096 CTAs][8] cycle counters of the last launch
  static long long* buf = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    if (getenv("SSV_TC_PROF") && cudaMalloc((void**)&buf, sizeof(long long) * 4096 * 8) != cudaSuccess) buf = nullptr;
  }
  return buf;
}

int* tf32_err_flag() {
  static int* flag = nullptr;
  if (!flag) {
    if (cudaMalloc((void**)&flag, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(flag, 0, sizeof(int));
  }
  return flag;
}

}  // namespace

void tf32_shape_highway(Tf32Layer* L, int d) {
  L->n_real = d;
  L->cluster_n = d / 128;
  L->w0_base = 0; L->w0_rank = 128;
  L->w1_base = d; L->w1_rank = 128;
}

void tf32_shape_plain(Tf32Layer* L, int n) {          // n need not fill the cluster: 513 -> three CTAs, 255 padding columns
  L->n_real = n;
  L->cluster_n = (n + 255) / 256;
  L->w0_base = 0; L->w0_rank = 256;
  L->w1_base = 128; L->w1_rank = 256;
}

// dgrad operand of a highwayConv (weight (2d, d, k)): rows = d input channels, K = k taps x 2d conv outputs, taps mirrored:
// dst[ci][j' * 2d + co] = w[co][ci][k - 1 - j'], split into hi / lo
__global__ void pack_dgrad_w_tf32_kernel(const float* __restrict__ w, int d, int k, float* __restrict__ hi, float* __restrict__ lo) {
  const long kp = (long)k * 2 * d;
  const long total = (long)d * kp;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int kk = (int)(i % kp);
    const int ci = (int)(i / kp);
    const int jp = kk / (2 * d), co = kk - jp * 2 * d;
    const float v = w[((long)co * d + ci) * k + (k - 1 - jp)];
    const float h = to_tf32(v);
    hi[i] = h;
    lo[i] = to_tf32(v - h);
  }
}

// ConvTranspose1d(k = 2, stride 2) weight (cin, cout, 2) as the 1x1 operand [2 cout rows (j, co)][cin], split into hi / lo
__global__ void pack_deconv_w_tf32_kernel(const float* __restrict__ w, int cin, int cout, float* __restrict__ hi, float* __restrict__ lo) {
  const long total = (long)2 * cout * cin;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin);
    const int row = (int)(i / cin);
    const int j = row / cout, co = row - j * cout;
    const float v = w[((long)ci * cout + co) * 2 + j];
    const float h = to_tf32(v);
    hi[i] = h;
    lo[i] = to_tf32(v - h);
  }
}

int tf32_pack_deconv_weights(const float* w, int cin, int cout, float* hi, float* lo, cudaStream_t s) {
  pack_deconv_w_tf32_kernel<<<1024, 256, 0, s>>>(w, cin, cout, hi, lo);
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

int tf32_pack_dgrad_weights(const float* w, int d, int k, float* hi, float* lo, cudaStream_t s) {
  pack_dgrad_w_tf32_kernel<<<1024, 256, 0, s>>>(w, d, k, hi, lo);
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

int tf32_pack_weights(const float* w, int n, int cin, int k, int cin_p, float* hi, float* lo, cudaStream_t s) {
  pack_w_tf32_kernel<<<1024, 256, 0, s>>>(w, n, cin, k, cin_p, hi, lo);
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

int launch_split_tf32(const float* x, float* hi, float* lo, size_t n, cudaStream_t s) {
  split_tf32_kernel<<<1024, 256, 0, s>>>(x, hi, lo, n);
  ++g_launches;
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

// ---- split-K wgrad operands (see conv_tc32.cuh) ----
// dH (M rows x ldh, n_out real columns, channels-last) -> [chunk][n_out][kc], hi / lo; rows past M are zero.  32 x 32 tiles;
// kc % 32 == 0.
__global__ void __launch_bounds__(256) wgrad_prep_dh_kernel(const float* __restrict__ dH, int ldh, int M, int n_out, int kc,
                                                            float* __restrict__ Ah, float* __restrict__ Al) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int r = r0 + ty + i;
    tile[ty + i][tx] = r < M && c0 + tx < n_out ? dH[(size_t)r * ldh + c0 + tx] : 0.f;
  }
  __syncthreads();
  const int chunk = r0 / kc, kk = r0 - chunk * kc + tx;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int co = c0 + ty + i;
    if (co >= n_out) continue;
    const float v = tile[tx][ty + i];
    const float h = to_tf32(v);
    const size_t o = ((size_t)chunk * n_out + co) * kc + kk;
    Ah[o] = h;
    Al[o] = to_tf32(v - h);
  }
}
// X (B T rows x x_ld, cin real columns, channels-last) -> [(j, ci)][k_pad] with row r = (b, t) holding
// X[(b, t + (tap_base + j) dil)][ci] (zero outside the utterance and past the last row), hi / lo
__global__ void __launch_bounds__(256) wgrad_prep_x_kernel(const float* __restrict__ X, int x_ld, int M, int T, int cin, int dil, int tap_base,
                                                           int k_pad, float* __restrict__ Wh, float* __restrict__ Wl) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32, j = blockIdx.z;
  const int off = (tap_base + j) * dil;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int r = r0 + ty + i;
    float v = 0.f;
    if (r < M && c0 + tx < cin) {
      const int b = r / T, ts = r - b * T + off;
      if (ts >= 0 && ts < T) v = X[((size_t)b * T + ts) * x_ld + c0 + tx];
    }
    tile[ty + i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int ci = c0 + ty + i;
    if (ci >= cin) continue;
    const float v = tile[tx][ty + i];
    const float h = to_tf32(v);
    const size_t o = ((size_t)j * cin + ci) * k_pad + r0 + tx;
    Wh[o] = h;
    Wl[o] = to_tf32(v - h);
  }
}
// partials [chunk][n_out][k cin] -> dW (n_out, cin, k)
__global__ void wgrad_tc_reduce_kernel(const float* __restrict__ P, int chunks, int n_out, int cin, int k, float* __restrict__ dW) {
  const long total = (long)n_out * k * cin;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin);
    const int j = (int)((i / cin) % k);
    const int co = (int)(i / ((long)cin * k));
    float sum = 0.f;
    for (int c = 0; c < chunks; ++c) sum += P[(size_t)c * total + i];
    dW[((size_t)co * cin + ci) * k + j] = sum;
  }
}
// w (n, cin) -> the dgrad operand of a 1x1 conv, [cin][k_p] = w^T with zero columns past n, hi / lo
__global__ void pack_transposed_tf32_kernel(const float* __restrict__ w, int n, int cin, int k_p, float* __restrict__ hi, float* __restrict__ lo) {
  const long total = (long)cin * k_p;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int kk = (int)(i % k_p);
    const int ci = (int)(i / k_p);
    const float v = kk < n ? w[(size_t)kk * cin + ci] : 0.f;
    const float h = to_tf32(v);
    hi[i] = h;
    lo[i] = to_tf32(v - h);
  }
}

int launch_wgrad_prep_dh(const float* dH, int ldh, int M, int n_out, int kc, int chunks, float* Ah, float* Al, cudaStream_t s) {
  SSV_CHECK(kc % 32 == 0 && (long)chunks * kc >= M, "wgrad_prep_dh: bad chunking (kc %d, chunks %d, M %d)", kc, chunks, M);
  wgrad_prep_dh_kernel<<<dim3((unsigned)(chunks * kc / 32), (unsigned)((n_out + 31) / 32)), 256, 0, s>>>(dH, ldh, M, n_out, kc, Ah, Al);
  ++g_launches;
  SSV_CUDA(cudaGetLastError());
  return kOk;
}
int launch_wgrad_prep_x(const float* X, int x_ld, int B, int T, int cin, int k, int dil, int tap_base, int k_pad, float* Wh, float* Wl,
                        cudaStream_t s) {
  SSV_CHECK(k_pad % 32 == 0 && k_pad >= B * T, "wgrad_prep_x: bad padding");
  wgrad_prep_x_kernel<<<dim3((unsigned)(k_pad / 32), (unsigned)((cin + 31) / 32), (unsigned)k), 256, 0, s>>>(X, x_ld, B * T, T, cin, dil, tap_base, k_pad, Wh, Wl);
  ++g_launches;
  SSV_CUDA(cudaGetLastError());
  return kOk;
}
int launch_wgrad_tc_reduce(const float* P, int chunks, int n_out, int cin, int k, float* dW, cudaStream_t s) {
  wgrad_tc_reduce_kernel<<<1024, 256, 0, s>>>(P, chunks, n_out, cin, k, dW);
  ++g_launches;
  SSV_CUDA(cudaGetLastError());
  return kOk;
}
int tf32_pack_transposed(const float* w, int n, int cin, int k_p, float* hi, float* lo, cudaStream_t s) {
  pack_transposed_tf32_kernel<<<256, 256, 0, s>>>(w, n, cin, k_p, hi, lo);
  ++g_launches;
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

int* tf32_err_flag_dev() { return tf32_err_flag(); }

int tf32_check_error() {
  int* flag = tf32_err_flag();
  if (!flag) return kOk;
  int h = 0;
  SSV_CUDA(cudaMemcpy(&h, flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (h != 0) {
    cudaMemset(flag, 0, sizeof(int));
    set_error("tcgen05 (3xTF32) conv kernel: pipeline wait timed out (code %d)", h);
    return kState;
  }
  return kOk;
}

int tf32_prepare(const Tf32Layer& L, int epi, int dil, int causal, const float* Xh, const float* Xl, int x_ld, int T, int B,
                 float* Yh, float* Yl, int y_ld, Tf32Launch* out) {
  return tf32_prepare_split(L, epi, dil, causal, Xh, Xl, x_ld, T, B, Yh, Yl, y_ld, 1, 0, out);
}

int tf32_max_clusters(int cluster_n) {
  static int cached[5] = {0, 0, 0, 0, 0};
  if (cluster_n < 1 || cluster_n > 4) return 0;
  if (cached[cluster_n] == 0) {
    int sm = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    int n = sm / cluster_n;
    if (cluster_n > 1) {
      cudaFuncSetAttribute(conv_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(226 * 1024));
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)(sm / cluster_n * cluster_n));
      cfg.blockDim = dim3(NT);
      cfg.dynamicSmemBytes = 1024 + 256 + (size_t)T32_NL * 16 + 6 * T32_BM * 16 + 4 * (size_t)STAGE_BYTES;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)cluster_n;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int q = 0;
      if (cudaOccupancyMaxActiveClusters(&q, conv_tf32x3_kernel, &cfg) == cudaSuccess && q > 0) n = q;
      else cudaGetLastError();
    }
    cached[cluster_n] = n > 0 ? n : 1;
  }
  return cached[cluster_n];
}

int tf32_prepare_split(const Tf32Layer& L, int epi, int dil, int causal, const float* Xh, const float* Xl, int x_ld, int T, int B,
                       float* Yh, float* Yl, int y_ld, int ksplit, long y_chunk_stride, Tf32Launch* out) {
  SSV_CHECK(ksplit >= 1 && (ksplit == 1 || (epi == EPI_NONE && Yl == nullptr && L.cin_p % (T32_BK * ksplit) == 0 && L.w_k_per_b == 0)),
            "conv_tf32: split-K needs the no-epilogue mode and Cin a multiple of %d", T32_BK * ksplit);
  SSV_CHECK(L.cin_p % T32_BK == 0 && x_ld >= L.cin_p && x_ld % 4 == 0, "conv_tf32: bad K padding (cin_p %d, ld %d)", L.cin_p, x_ld);
  SSV_CHECK(epi == EPI_HIGHWAY || epi == EPI_LN || epi == EPI_LN_RELU || epi == EPI_LN_SIGMOID || epi == EPI_NONE,
            "conv_tf32: epilogue %d not built", epi);
  SSV_CHECK(L.cluster_n >= 1 && L.cluster_n <= 4, "conv_tf32: cluster_n must be 1..4");
  SSV_CHECK(y_ld % 4 == 0, "conv_tf32: output row stride must be a multiple of 4");
  ConvTf32Args& a = out->args;
  memset(&a, 0, sizeof(a));
  a.T = T; a.B = B;
  a.rows_per_utt = T <= 64 ? 64 : T32_BM;
  a.utt_per_tile = T32_BM / a.rows_per_utt;
  a.tiles_per_b = (T + T32_BM - 1) / T32_BM;
  const int n_tiles = a.utt_per_tile > 1 ? (B + a.utt_per_tile - 1) / a.utt_per_tile : B * a.tiles_per_b;
  {
    cuuint64_t dims[3] = {(cuuint64_t)x_ld, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)x_ld * 4, (cuuint64_t)T * x_ld * 4};
    cuuint32_t box[3] = {(cuuint32_t)T32_BK, (cuuint32_t)a.rows_per_utt, (cuuint32_t)a.utt_per_tile};
    SSV_TRY(make_map_f32(&a.tmAh, Xh, 3, dims, strides, box));
    SSV_TRY(make_map_f32(&a.tmAl, Xl, 3, dims, strides, box));
  }
  const int kp = L.w_cols ? L.w_cols : L.k * L.cin_p;
  SSV_CHECK(L.w_k_per_b == 0 || a.utt_per_tile == 1, "conv_tf32: split-K operands need whole 128-row tiles");
  a.w_k_per_b = L.w_k_per_b;
  {
    cuuint64_t dims[2] = {(cuuint64_t)kp, (cuuint64_t)L.rows};
    cuuint64_t strides[1] = {(cuuint64_t)kp * 4};
    cuuint32_t box[2] = {(cuuint32_t)T32_BK, 128};
    SSV_TRY(make_map_f32(&a.tmBh, L.Wh, 2, dims, strides, box));
    SSV_TRY(make_map_f32(&a.tmBl, L.Wl, 2, dims, strides, box));
  }
  a.kb_per_tap = L.cin_p / T32_BK / ksplit;
  a.ksplit = ksplit;
  a.w_tap_stride = L.cin_p;
  a.y_chunk_stride = y_chunk_stride;
  a.ktaps = L.k; a.dil = dil; a.causal = causal;
  a.cluster_n = L.cluster_n;
  a.w0_base = L.w0_base; a.w0_rank = L.w0_rank; a.w1_base = L.w1_base; a.w1_rank = L.w1_rank;
  a.n_real = L.n_real;
  a.n_cols = L.rows;
  {
    const int wide = epi == EPI_HIGHWAY ? L.n_real : L.cluster_n * T32_NL;
    a.y_cols = y_ld < wide ? y_ld : wide;
  }
  a.epi = epi;
  a.bias = L.bias;
  a.g1 = L.g1; a.b1 = L.b1; a.g2 = L.g2; a.b2 = L.b2;
  a.Xh = Xh; a.Xl = Xl; a.x_sb = (long)T * x_ld; a.x_st = x_ld;
  a.Yh = Yh; a.Yl = Yl; a.y_sb = (long)T * y_ld; a.y_st = y_ld;
  a.nstages = 4;
  out->n_ctas = n_tiles * ksplit * L.cluster_n;
  out->cluster_n = L.cluster_n;
  return kOk;
}

int tf32_run(const Tf32Launch& L, cudaStream_t s) {
  int* err = tf32_err_flag();
  SSV_CHECK(err != nullptr, "conv_tf32: cannot allocate the error flag");
  const size_t fixed = 1024 /*align*/ + 256 /*barriers, tmem slot*/ + (size_t)T32_NL * 16 + 2 * T32_BM * 16 + 4 * T32_BM * 16;
  const size_t smem = fixed + (size_t)L.args.nstages * STAGE_BYTES;
  static bool configured = false;
  if (!configured) {
    SSV_CUDA(cudaFuncSetAttribute(conv_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(226 * 1024)));
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
  long long* prof = tf32_prof_buf();
  const int ncta = L.n_ctas;
  if (prof && ncta <= 4096) {
    Tf32Launch P = L;
    P.args.prof = prof;
    SSV_CUDA(cudaMemsetAsync(prof, 0, sizeof(long long) * 4096 * 8, s));
    SSV_CUDA(cudaLaunchKernelEx(&cfg, conv_tf32x3_kernel, P.args, err));
    ++g_launches;
    SSV_CUDA(cudaStreamSynchronize(s));
    std::vector<long long> h((size_t)ncta * 8);
    SSV_CUDA(cudaMemcpy(h.data(), prof, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    const char* nm[8] = {"setup", "mma-wait-full", "mma-loop", "epi-wait-accum", "pass1", "exchange", "pass2", "total"};
    fprintf(stderr, "[tf32 prof] grid=%d cluster=%d K=%d epi=%d T=%d B=%d :", ncta, L.cluster_n, L.args.ktaps * L.args.kb_per_tap * T32_BK,
            L.args.epi, L.args.T, L.args.B);
    for (int i = 0; i < 8; ++i) {
      double sum = 0;
      int cnt = 0;
      for (int c = 0; c < ncta; ++c)
        if (h[(size_t)c * 8 + i] != 0) { sum += (double)h[(size_t)c * 8 + i]; ++cnt; }
      fprintf(stderr, " %s=%.0f", nm[i], cnt ? sum / cnt : 0.0);
    }
    fprintf(stderr, "\n");
    return kOk;
  }
  SSV_CUDA(cudaLaunchKernelEx(&cfg, conv_tf32x3_kernel, L.args, err));
  ++g_launches;
  return kOk;
}

int tf32_launch(const Tf32Layer& L, int epi, int dil, int causal, const float* Xh, const float* Xl, int x_ld, int T, int B,
                float* Yh, float* Yl, int y_ld, cudaStream_t s) {
  Tf32Launch P;
  SSV_TRY(tf32_prepare(L, epi, dil, causal, Xh, Xl, x_ld, T, B, Yh, Yl, y_ld, &P));
  return tf32_run(P, s);
}

}  // namespace ssv
