// Weight-stationary, layer-pipelined incremental Text2Mel decode (the default decode kernel).
//
// Replaces the reference AR loop (generate_test_utterances.py:105-116, synthesize.py:103-109 driving the
// eval branch of melSyn.forward, models/TTSModel.py:275-300): every causal highwayConv keeps the history of
// its own input, so frame t costs O(1) instead of O(t).
//
// What is different here is WHERE the work lives.  The 24 mat-vec stages of a frame (AudioEnc 13,
// attention + AudioDec 11) own 6.8 M fp32 weights = 27.3 MB; a B200 has 37 MB of registers and 33 MB of
// shared memory spread over its 148 SMs.  So every stage gets a fixed group of CTAs (8 per highway layer,
// 1-4 per 1x1 conv, 145 in total), each CTA loads its weight slice ONCE per launch -- 88 weights per thread
// into REGISTERS, the remaining tap of a highway layer into shared memory -- and the utterances flow
// through the stage groups as micro-batches of RT rows:
//
//     stage s, micro-batch g, frame t   needs   stage s-1, g, t      (stage 0: stage 23, g, t-1)
//
// a software pipeline over (frame, micro-batch): up to 24 micro-batches are in flight, HBM/L2 see only
// activations, and there is no grid-wide barrier anywhere.
//
// Inside a CTA the work is warp-specialised.  Warps 0-3 are the front end: they fetch the taps t-2d, t-d from the
// CTA's private history ring, wait for the producer stage, redo the cheap LayerNorm / highway gate / windowed
// attention of the input row (redundantly per consumer CTA), and fill an X buffer.  How the
// four warps are dealt is a launch-time shape (ws_plan): W warps share one row (each owns 256 / W channels, the
// LayerNorm sums cross warps through shared memory and a named barrier), R rows make a micro-batch, and
// 4 / (R W) micro-batches are in the front end at a time.  Few utterances: W = 4 or 2, the dependent chain of a
// stage is 2-4 channels per lane; many: W = 1 and two micro-batches in flight, so that the wait for one producer
// hides behind the prologue of the other.  A warp only ever reads history it wrote itself (its channel slice of
// its visit slot), so the taps need no cross-warp synchronisation.
// Warps 4-15 hold the weights and do the mat-vec: the two OLD taps (2/3 of a highway layer's work) do not depend
// on the producer stage and run while the front end is still waiting for it; only the current tap sits on the
// frame's critical path.  The two halves hand X buffers back and forth through mbarriers.
//
// Cross-CTA hand-off: every value a CTA publishes is an 8-byte word {float value, int tag} written with one
// 64-bit store (single-copy atomic); tag = seq_base + frame + 1.  A consumer polls the tagged words of its own
// row slice directly -- it loads them, checks every tag and re-loads until all match -- so there is exactly one
// L2 round trip between "the producer's store landed" and "the consumer has the data", no fences and no
// release/acquire chain on the critical path, and a word that has not landed yet can never be mistaken for data.
// (Polling a per-CTA sentinel first and loading the row afterwards, as the first version did, costs a second
// dependent round trip per stage: 43 vs 31 us/frame at B = 1.)  Everything else a CTA touches
// is private to it (weights, LayerNorm parameters, its own ring of past stage inputs), so the tagged words are
// the only cross-CTA traffic.
#include "decode.cuh"

#include <cstdlib>
#include <cstring>

namespace ssv {

namespace {

constexpr int GV_T = WS_GEMV_THREADS;    // mat-vec threads (the 12 warps after the front end)
constexpr int MAX_FEW = 8;               // front-end warps: 4 (512-thread CTA) or 8 (640 threads, setmaxnreg)
constexpr int HD = 256;
constexpr int TAPP = WS_TAP_ROWS;        // rows of one tap in X (256 padded to a multiple of 24)
constexpr long long SPIN_LIMIT = 4000000000LL;   // ~2 s of SM clocks
constexpr unsigned FULL = 0xffffffffu;
// timing-only knock-outs (wrong values): free-running stages, gutted roles
#ifdef SSV_KO_NOPOLL
constexpr bool KO_NOPOLL = true;
#else
constexpr bool KO_NOPOLL = false;
#endif
#ifdef SSV_KO_NOFE
constexpr bool KO_NOFE = true;
#else
constexpr bool KO_NOFE = false;
#endif
#ifdef SSV_KO_NOMV
constexpr bool KO_NOMV = true;
#else
constexpr bool KO_NOMV = false;
#endif

// floats per stored activation in an X buffer: 2 = the pair (x, x), the multiplicand of a packed FMA (fma.rn.f32x2)
// (two-row micro-batches in the packed form, X rows (x0, x0, x1, x1): the mat-vec role alone runs 10 % slower, 45.8 vs
// 41.4 us per frame at B = 64 -- the packed FMA issues at less than half the rate of the scalar one)
template <int RT> struct XLayout { static constexpr int D = RT == 1 ? 2 : 1; };

// shared-memory carve-up (floats)
constexpr int XROWS = 3 * TAPP;                    // rows of one X buffer ([rows][RT] activations; RT = 1 stores each as (x, x))
constexpr int XREGION = 2 * 8 * XROWS;             // floats; ws_nbuf(RT) buffers of XROWS rows x (2 or 4) floats
constexpr int MAXBUF = 8;
// X buffers per shape: 8 (one- and two-row micro-batches, 2 floats per X row) or 4 (four rows)
__host__ __device__ constexpr int ws_nbuf(int rt) { return rt * (rt == 1 ? XLayout<1>::D : rt == 2 ? XLayout<2>::D : XLayout<4>::D) >= 4 ? 4 : 8; }
constexpr int WSM_TILES = 15;                      // k rows of weights a mat-vec thread may keep in shared memory
constexpr int SM_WSM = 0;                          // [<= 15][384] float4: the weight rows that do not fit the registers
constexpr int SM_X = SM_WSM + WSM_TILES * GV_T * 4;       // [8 / RT][XROWS][RT * XLayout<RT>::D]
constexpr int SM_PART = SM_X + XREGION;            // [12][RT][ncol <= 128] k-slice partial sums
constexpr int SM_LN = SM_PART + 12 * 4 * 128;      // [4][256] LayerNorm parameters of my prologue
constexpr int SM_BIAS = SM_LN + 4 * HD;            // [128] bias of my columns
constexpr int SM_PMA = SM_BIAS + 128;              // [WS_MAX_BATCH] ints (attention stage only)
constexpr int SM_KV = SM_PMA + WS_MAX_BATCH;        // [rows in the front end <= 8][K window 3 | V window 3][256]: attention rows prefetched by cp.async
constexpr int SM_RED = SM_KV + MAX_FEW * 6 * HD;    // [4 rounds][8 warps][8]: cross-warp LayerNorm statistics / logit sums (cooperative front end)
constexpr int SM_TOTAL = SM_RED + 4 * MAX_FEW * 8;

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
#ifdef SSV_KO_NOLOAD
  v = 0.25f; tag = 0; return;
#endif
  int a, b;
  asm volatile("ld.relaxed.gpu.global.v2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "l"(p) : "memory");
  v = __int_as_float(a);
  tag = b;
}
__device__ __forceinline__ void ld_word2(const Word* p, float& v0, int& t0, float& v1, int& t1) {   // 16-byte aligned
#ifdef SSV_KO_NOLOAD
  v0 = 0.25f; v1 = 0.5f; t0 = t1 = 0; return;
#endif
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
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz));
}
// ---- bulk copies (TMA, 1-D): the history ring entries move between global memory and the X buffers without
// passing through registers.  Entry = the stage input of one (frame, micro-batch) in X's own layout [256][XS].
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// the issuing thread's bulk stores: all but the newest are complete (written), and the newest has read its source
__device__ __forceinline__ void bulk_wait_prev_written_last_read() {
  asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
#ifdef SSV_KO_NOFENCE
__device__ __forceinline__ void fence_proxy_async_smem() {}
#else
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// ---- mbarriers (CTA-local producer / consumer hand-off of the X buffers)
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrival that fires when all cp.async of the executing thread issued so far have landed
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"     // suspend-time hint: sleep in hardware, do not spin
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(2000u)
      : "memory");
  return ok != 0;
}
// bounded wait; false: the launch was aborted (some CTA timed out), leave
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, unsigned parity, int* abort_flag) {
  unsigned spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > SPIN_LIMIT) atomicExch(abort_flag, 9);
      if (*reinterpret_cast<volatile int*>(abort_flag) != 0) return false;
    }
  }
  return true;
}
__device__ __forceinline__ void named_bar(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

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
// One-pass LayerNorm statistics: every lane holds (sum, M2 = sum of squared deviations from its own mean) of N0
// values; the butterfly merges equal-sized groups with Chan's update M2 = M2a + M2b + (sa - sb)^2 / (2 n), which is
// symmetric, so all lanes end with bit-identical totals -- one dependent shuffle chain instead of two (mean, then
// variance).  Two independent channels (H1 | H2) ride the same chain.
template <int N0>
__device__ __forceinline__ void warp_stats2(float& s1, float& m1, float& s2, float& m2) {
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int off = 16 >> k;
    const float c = 0.5f / (float)(N0 << k);
    const float sp1 = __shfl_xor_sync(FULL, s1, off), mp1 = __shfl_xor_sync(FULL, m1, off);
    const float sp2 = __shfl_xor_sync(FULL, s2, off), mp2 = __shfl_xor_sync(FULL, m2, off);
    const float d1 = s1 - sp1, d2 = s2 - sp2;
    m1 = (m1 + mp1) + d1 * d1 * c;
    m2 = (m2 + mp2) + d2 * d2 * c;
    s1 += sp1;
    s2 += sp2;
  }
}
template <int N0>
__device__ __forceinline__ void warp_stats1(float& s1, float& m1) {
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int off = 16 >> k;
    const float c = 0.5f / (float)(N0 << k);
    const float sp1 = __shfl_xor_sync(FULL, s1, off), mp1 = __shfl_xor_sync(FULL, m1, off);
    const float d1 = s1 - sp1;
    m1 = (m1 + mp1) + d1 * d1 * c;
    s1 += sp1;
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// Highway gate / LayerNorm scale on the frame's critical path: MUFU-based, branch-free (2 ulp), so the eight
// per-lane chains interleave instead of serialising on the slow-path calls of IEEE division / sqrt.
#ifdef SSV_KO_NOGATE
__device__ __forceinline__ float sigmoid_fast(float x) { return 0.5f * x; }
__device__ __forceinline__ float rstd_fast(float var) { return var + 1e-5f; }
#else
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float rstd_fast(float var) { return rsqrtf(var + 1e-5f); }
#endif

// One k row of the mat-vec tile: acc[r][0..3] += x[r] * w[0..3] for the RT rows of the micro-batch.
// RT = 1 (latency mode) uses packed FMAs (fma.rn.f32x2: two independent IEEE fp32 FMAs per instruction,
// bit-identical to fmaf): the accumulators are column pairs, the weights are the halves of the thread's float4 and X
// holds every activation twice, (x, x), so the multiplicand needs no shuffling -- 2 FMA instructions + one LDS.64
// per row instead of 4 + 1.  Measured: B = 1 30.0 -> 28.9 us/frame.  With 2 or 4 rows per micro-batch the mat-vec
// phases are bound by the FMA pipe, where the packed form is ~9 % slower than scalar FFMA (B = 64: 53.7 -> 58
// us/frame), so those shapes keep scalar FMAs and the plain X layout.
// (Tried: the accumulators of multi-row micro-batches in 64-bit register pairs holding columns (1, 0) and (3, 2), so that
// every FFMA reads its accumulator and its weight from registers of opposite parity.  A three-register FFMA whose two
// fresh operands share a register bank issues at half rate -- tools/micro/ffma_bank_bench.cu: 0.88 FFMA per cycle and
// scheduler when only the first FFMA of each reuse group conflicts, and ptxas, with all registers in use here, leaves
// 44 % of this loop's FFMAs conflicting, which is the 0.63 per cycle measured in situ.  The pairs remove those
// conflicts in the SASS, but the launch does not get faster: B = 64 49.4 -> 49.4 us/frame, 128 89.5 -> 88.4,
// 192 126 -> 122, 256 164.8 -> 168-170.  The FMA loops are filler between hand-offs, not the bound -- see
// profiles/r2_decode_ws_knockouts.md: without any FMA the launch is only 14-22 % shorter.)

template <int XS> struct XVecN { using T = float2; };         // one X row: (x, x) | (x0, x1) | (x0, x0, x1, x1) | (x0, x1, x2, x3)
template <> struct XVecN<4> { using T = float4; };
template <int RT> struct XVec { using T = typename XVecN<RT * XLayout<RT>::D>::T; };

template <int RT>
__device__ __forceinline__ void fma_row(float (&acc)[RT][4], const float4& w, const typename XVec<RT>::T& v) {
  if constexpr (RT == 1) {
    float2 a01 = make_float2(acc[0][0], acc[0][1]), a23 = make_float2(acc[0][2], acc[0][3]);
    a01 = __ffma2_rn(v, make_float2(w.x, w.y), a01);
    a23 = __ffma2_rn(v, make_float2(w.z, w.w), a23);
    acc[0][0] = a01.x; acc[0][1] = a01.y; acc[0][2] = a23.x; acc[0][3] = a23.y;
  } else if constexpr (RT == 2 && XLayout<RT>::D == 2) {
    const float2 x0 = make_float2(v.x, v.y), x1 = make_float2(v.z, v.w);
    const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
    float2 a = make_float2(acc[0][0], acc[0][1]), b = make_float2(acc[0][2], acc[0][3]);
    float2 c = make_float2(acc[1][0], acc[1][1]), d = make_float2(acc[1][2], acc[1][3]);
    a = __ffma2_rn(x0, w01, a); b = __ffma2_rn(x0, w23, b);
    c = __ffma2_rn(x1, w01, c); d = __ffma2_rn(x1, w23, d);
    acc[0][0] = a.x; acc[0][1] = a.y; acc[0][2] = b.x; acc[0][3] = b.y;
    acc[1][0] = c.x; acc[1][1] = c.y; acc[1][2] = d.x; acc[1][3] = d.y;
  } else {
    float x[RT];
    if constexpr (RT == 2) { x[0] = v.x; x[1] = v.y; }
    else { x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      acc[r][0] = fmaf(x[r], w.x, acc[r][0]);
      acc[r][1] = fmaf(x[r], w.y, acc[r][1]);
      acc[r][2] = fmaf(x[r], w.z, acc[r][2]);
      acc[r][3] = fmaf(x[r], w.w, acc[r][3]);
    }
  }
}
template <int RT>
__device__ __forceinline__ void fma_tile(float (&acc)[RT][4], const float4& w, const float* xp) {
  fma_row<RT>(acc, w, *reinterpret_cast<const typename XVec<RT>::T*>(xp));
}

// Development build (-DSSV_WS_TOTALS): every role of a CTA records the SM cycles it spent in the launch and those it spent
// waiting (mbarriers of the other role, tagged words of the producer stage); two clock reads per wait, nothing else.
#ifdef SSV_WS_TOTALS
#define TOT_DECL long long tot_wait_ = 0, tot_w0_ = 0
#define TOT_W0 tot_w0_ = clock64()
#define TOT_W1 tot_wait_ += clock64() - tot_w0_
#define TOT_END(role_, cond_)                                                  \
  if (p.totals != nullptr && (cond_)) {                                        \
    p.totals[(size_t)blockIdx.x * 4 + 2 * (role_)] = clock64() - c.t0;         \
    p.totals[(size_t)blockIdx.x * 4 + 2 * (role_) + 1] = tot_wait_;            \
  }
#else
#define TOT_DECL
#define TOT_W0
#define TOT_W1
#define TOT_END(role_, cond_)
#endif

struct Ctx {                 // per-CTA constants shared by both roles
  int s, prev, part, G, B;
  float* smem;
  uint64_t* tapsfull;        // [8] old taps of the X buffer have landed (cp.async)
  uint64_t* curfull;         // [8] current tap (prologue output) is in the X buffer
  uint64_t* empty;           // [8] mat-vec warps are done reading the X buffer
  uint64_t* pfull;           // [2] k-slice partial sums of all 12 mat-vec warps are in shared memory
  uint64_t* pfree;           // [2] the reducer warp(s) are done with the partial-sum buffer
  uint64_t* rdone;           // [2] second reducer warp has published its half (wide tiles)
  int* s_bad;
  volatile long long* t_seen;   // profiling, [2][8]: SM clock at which the front end saw the producer's sentinel / handed X over
  long long t0;                 // SM clock at the start of the roles (DecParams::totals)
};

// global column (= tagged word index) of local column lc; highway CTAs own matching H1 / H2 slices
__device__ __forceinline__ int gcol(const WsStage& st, int part, int lc) {
  const int half = st.ncol / 2;
  return st.hwy ? (lc < half ? part * half + lc : HD + part * half + (lc - half)) : part * st.ncol + lc;
}

// ------------------------------------------------------------------------------------------------------
// Mat-vec role (the 12 warps after the front end).  Thread (cg, ks) owns 4 columns and the k rows ks + KS*j of every
// tap: 33 rows ("tiles") of a highway layer, <= 22 of a 1x1 conv.  The last ws_nreg(RT) of them live in registers, the
// first ones in shared memory: 22 + 11 for one- and two-row micro-batches; four rows need 16 accumulators and 8 fetch
// registers, so 18 + 15 (with 22 the tile loops spilled: 2.95 k cycles of FMA loop per visit against 2 x 1.25 k).
__host__ __device__ constexpr int ws_nreg(int rt) { return rt == 4 ? 18 : 22; }
template <int RT, int CG, bool HWY, bool PROF>
__device__ __forceinline__ void gemv_role(const DecParams& p, const WsStage& st, const Ctx& c, int gtid) {
  constexpr int KS = GV_T / CG;            // k-slices: 24 (64 columns) or 12 (128 columns)
  constexpr int NCOL = 4 * CG;
  constexpr int NBUF = ws_nbuf(RT);
  const int cg = gtid % CG, ks = gtid / CG;
  const int lane = gtid & 31, gwarp = gtid >> 5;
  float* parts = c.smem + SM_PART;
  const float* bias_s = c.smem + SM_BIAS;
  const float4* wsm = reinterpret_cast<const float4*>(c.smem + SM_WSM) + gtid;

  // my weights -> registers (image is thread-major: [tile][384] float4): the last NREG tiles
  constexpr int NREG = ws_nreg(RT);
  constexpr int NSM = (HWY ? 33 : 22) - NREG;       // tiles [0, NSM) are read from shared memory
  float4 w[NREG];
  {
    const float4* img = reinterpret_cast<const float4*>(st.img) + (size_t)c.part * st.njt * GV_T + gtid;
#pragma unroll
    for (int j = 0; j < NREG; ++j) w[j] = NSM + j < st.njt ? __ldg(img + (size_t)(NSM + j) * GV_T) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  auto wt = [&](int t) -> float4 { return t < NSM ? wsm[t * GV_T] : w[t - NSM]; };     // t is a compile-time constant

  long long prof_last = 0, prof_acc[5] = {0, 0, 0, 0, 0}, lat_acc = 0, wake_acc = 0;
  TOT_DECL;
  const bool prof_on = PROF && p.prof != nullptr && gtid == 0;
  if (prof_on) prof_last = clock64();
#define PROF_G(i)                                  \
  if (prof_on) {                                   \
    const long long now_ = clock64();              \
    prof_acc[i] += now_ - prof_last;               \
    prof_last = now_;                              \
  }

  Word* raw_out = reinterpret_cast<Word*>(p.ws_raw) + (size_t)c.s * c.B * WS_WORDS;
  auto ring_store = [&](int step, int g, int q) {
    constexpr int XS = RT * XLayout<RT>::D;
    const int slot = (p.t_start + step) % st.hist_depth;
    float* dst = p.ws_hist + (((size_t)(st.hist_blk0 + c.part * st.hist_depth + slot)) * c.G + g) * (HD * XS);
    bulk_s2g(dst, c.smem + SM_X + q * (XROWS * XS) + (size_t)(2 * TAPP) * XS, HD * XS * (unsigned)sizeof(float));
  };
  int v = 0;                                // visit counter (same sequence as the front end; < 2^31: checked at launch)
  for (int step = 0; step < p.n_steps; ++step) {
    const int tag = p.seq_base + p.t_start + step + 1;
    for (int g = 0; g < c.G; ++g, ++v) {
      const int q = v % NBUF;
      const unsigned par = (unsigned)(v / NBUF) & 1u;
      constexpr int XS = RT * XLayout<RT>::D;          // floats per X row
      const float* X = c.smem + SM_X + q * (XROWS * XS) + ks * XS;
      // publisher threads: the hoisted speaker projection of my output is fetched now, not after the mat-vec
      const int row0 = g * RT;
      float sb[(RT * NCOL + GV_T - 1) / GV_T];
#pragma unroll
      for (int i = 0; i < (RT * NCOL + GV_T - 1) / GV_T; ++i) {
        const int o = gtid + i * GV_T;
        sb[i] = 0.f;
        if (st.bias_b != 0 && o < RT * NCOL) {
          const int r = o / NCOL, lc = o - r * NCOL;
          const int gc = gcol(st, c.part, lc), b = row0 + r;
          if (b < c.B && gc < st.n) sb[i] = __ldg((st.bias_b == 1 ? p.s1 : p.s2) + (size_t)b * HD + gc);
        }
      }
      if (KO_NOMV) {                // timing-only: the mat-vec role just hands the X buffers back
        if (HWY) mbar_wait(&c.tapsfull[q], par, p.abort_flag);
        mbar_wait(&c.curfull[q], par, p.abort_flag);
#ifndef SSV_KO_NORING
        if (HWY && gtid == 0) { ring_store(step, g, q); bulk_wait_prev_written_last_read(); }
#endif
        named_bar(1, GV_T);
        if (gtid == 0) mbar_arrive(&c.empty[q]);
        continue;
      }
      float acc[RT][4];
#pragma unroll
      for (int r = 0; r < RT; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
      if (HWY) {
        // old taps: independent of the producer stage, overlaps the front end's wait
        TOT_W0;
        mbar_wait(&c.tapsfull[q], par, p.abort_flag);     // on abort: fall through, the uniform exit is below
        TOT_W1;
        PROF_G(0);
        // shared-memory-weight tiles (tap t-2d) alternate with register-weight tiles (tap t-d): twice the FMAs
        // between two LDS.128 weight fetches, so their latency hides with the two fragment buffers the register
        // budget allows (back to back, every other tile stalled on its fetch: 18 % of all samples short-scoreboard)
#ifndef SSV_KO_NOFMA
#pragma unroll
        for (int j = 0; j < 11; ++j) {
          fma_tile<RT>(acc, wt(j), X + (size_t)(KS * j) * XS);
          fma_tile<RT>(acc, wt(11 + j), X + (size_t)(TAPP + KS * j) * XS);
        }
#endif
        PROF_G(1);
        TOT_W0;
        mbar_wait(&c.curfull[q], par, p.abort_flag);
        TOT_W1;
        PROF_G(2);
        if (PROF && prof_on) wake_acc += prof_last - c.t_seen[8 + q];
        // The finished current-tap rows of this (frame, micro-batch) join the history ring: one bulk copy (TMA)
        // from the X buffer, in X's own layout, read back the same way when they become the taps t-d, t-2d.  The
        // front-end lanes fenced their stores towards the async proxy before they arrived on curfull.  Multi-row
        // micro-batches: issued ahead of this thread's FMAs, the other warps fill the issue slots meanwhile (issued by
        // the last warp after its FMAs instead, the barrier below waited for it: B = 128 88.8 -> 94.4 us/frame).
        // Latency mode (one row): after the publish, off the frame's critical path.
#ifndef SSV_KO_NORING
        if (RT > 1 && gtid == 0) ring_store(step, g, q);
#endif
#ifdef SSV_KO_NOFMA
        acc[0][0] += w[0].x * X[0];
#else
#pragma unroll
        for (int j = 0; j < 11; ++j) fma_tile<RT>(acc, wt(22 + j), X + (size_t)(2 * TAPP + KS * j) * XS);
#endif
      } else {
        TOT_W0;
        mbar_wait(&c.curfull[q], par, p.abort_flag);
        TOT_W1;
        PROF_G(2);
        // (Tried: the X rows of the 1x1 stages fetched as batches ahead of straight-line FMAs instead of row by row
        // under the row-count predicate: B = 64 49.4 -> 48.3 us/frame, but B = 1 29.4 -> 30.2 and B = 128 89.5 -> 95.3.)
#ifdef SSV_KO_NOFMA
        acc[0][0] += w[0].x * X[0];
#else
#pragma unroll
        for (int j = 0; j < 22; ++j)
          if (j < st.nj) fma_tile<RT>(acc, wt(j), X + (size_t)(KS * j) * XS);
#endif
        PROF_G(1);                  // 1x1 stages: "old taps" column = the tile loop, "current tap" = k-slice store + barrier
      }
      // k-slices -> shared memory
      bool writer = true;
      if (CG == 16) {               // the two half-warps hold adjacent k-slices of the same columns
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) acc[r][cc] += __shfl_xor_sync(FULL, acc[r][cc], 16);
        writer = lane < 16;
      }
      if (writer) {
#pragma unroll
        for (int r = 0; r < RT; ++r)
          *reinterpret_cast<float4*>(parts + ((size_t)gwarp * RT + r) * NCOL + cg * 4) =
              make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      }
      named_bar(1, GV_T);
      if (!HWY && gtid == 0) mbar_arrive(&c.empty[q]);     // X buffer q may be refilled (highway: after the publish, below)
      PROF_G(3);
      // reduce the 12 k-slices, add bias (+ hoisted speaker projection), publish tagged words
#ifdef SSV_KO_NOPUB
      if (false)
#endif
#pragma unroll
      for (int i = 0; i < (RT * NCOL + GV_T - 1) / GV_T; ++i) {
        const int o = gtid + i * GV_T;
        if (o < RT * NCOL) {
          const int r = o / NCOL, lc = o - r * NCOL;
          float sum = 0.f;
#pragma unroll
          for (int sl = 0; sl < 12; ++sl) sum += parts[((size_t)sl * RT + r) * NCOL + lc];
          const int gc = gcol(st, c.part, lc);
          const int b = row0 + r;
          if (b < c.B && gc < st.n) st_word(raw_out + (size_t)b * WS_WORDS + gc, sum + bias_s[lc] + sb[i], tag);
        }
      }
      if (HWY && gtid == 0) {
#ifndef SSV_KO_NORING
        if (RT == 1) ring_store(step, g, q);
        bulk_wait_prev_written_last_read();
#endif        // older ring entries are in memory, this one has left the X buffer
        mbar_arrive(&c.empty[q]);
      }
      // Second CTA-wide barrier: the k-slice buffer may be rewritten, and -- measured -- it keeps the warps without
      // an output from running ahead into the next micro-batch and taking issue slots from the publishing warps
      // (without it: B = 64 54 -> 60 us/frame, B = 128 108 -> 122).
      named_bar(1, GV_T);
      // abort flag: thread 0 looks at it at the end of visits 16k + 15, everybody reads the verdict after this barrier
      // of visit 16k + 16 (written a whole visit earlier, next written 15 visits later: the read is uniform)
      if ((v & 15) == 0 && *reinterpret_cast<volatile int*>(c.s_bad) == 2) return;
      if (gtid == 0 && (v & 15) == 15 && *reinterpret_cast<volatile int*>(p.abort_flag) != 0) *c.s_bad = 2;
      PROF_G(4);
      if (PROF && prof_on) lat_acc += prof_last - c.t_seen[q];
    }
  }
  TOT_END(1, gtid == 0);
  if (prof_on) {
#pragma unroll
    for (int i = 0; i < 5; ++i) p.prof[(size_t)blockIdx.x * 16 + 8 + i] = prof_acc[i];
    p.prof[(size_t)blockIdx.x * 16 + 7] = v > 0 ? v : 1;       // mat-vec visit count (the front end's is slot 15)
    p.prof[(size_t)blockIdx.x * 16 + 13] = lat_acc;
    p.prof[(size_t)blockIdx.x * 16 + 14] = wake_acc;
  }
#undef PROF_G
}

// ------------------------------------------------------------------------------------------------------
// Cooperative front end (warps 0-3): the 4 / RT warps of a row split its 256 channels, so the dependent chain of
// one stage -- tagged-word loads, two LayerNorm reductions, gate, X stores -- is 2 * RT channels per lane instead
// of 8.  One micro-batch is in the front end at a time; a warp reads only taps that it wrote itself (its channel
// slice), so the taps need no cross-warp synchronisation at all.
template <int WPR>
__device__ __forceinline__ void row_reduce(float& a, float& b, float& c3, bool& bad, float* red, int round, int warp, int wbase, int barid, int lane) {
  warp_sum2(a, b);
  c3 = warp_sum(c3);
  if (WPR > 1) {
    float4* slot = reinterpret_cast<float4*>(red) + round * (2 * MAX_FEW);
    if (lane == 0) slot[2 * warp] = make_float4(a, b, c3, bad ? 1.f : 0.f);
    named_bar(barid, WPR * 32);
    float sa = 0.f, sb = 0.f, sc = 0.f, sf = 0.f;
#pragma unroll
    for (int w = 0; w < WPR; ++w) {          // fixed order: every warp of the row gets bit-identical sums
      const float4 x = slot[2 * (wbase + w)];
      sa += x.x; sb += x.y; sc += x.z; sf += x.w;
    }
    a = sa; b = sb; c3 = sc;
    bad = sf != 0.f;                          // row-uniform from here on
  }
}

// (sum, M2) of two channels over the whole row: per-warp butterfly, then the WPR warp totals (CW values each) are
// merged pairwise in a fixed order by every warp.
template <int WPR, int N0>
__device__ __forceinline__ void row_stats2(float& s1, float& m1, float& s2, float& m2, bool& bad, float* red, int round, int warp, int wbase, int barid, int lane) {
  warp_stats2<N0>(s1, m1, s2, m2);
  if (WPR > 1) {
    constexpr int CW = HD / WPR;
    float4* slot = reinterpret_cast<float4*>(red) + round * (2 * MAX_FEW);
    if (lane == 0) {
      slot[2 * warp] = make_float4(s1, m1, s2, m2);
      slot[2 * warp + 1] = make_float4(bad ? 1.f : 0.f, 0.f, 0.f, 0.f);
    }
    named_bar(barid, WPR * 32);
    float4 e[WPR];
    float sf = 0.f;
#pragma unroll
    for (int w = 0; w < WPR; ++w) {
      e[w] = slot[2 * (wbase + w)];
      sf += slot[2 * (wbase + w) + 1].x;
    }
#pragma unroll
    for (int n = CW, cnt = WPR; cnt > 1; cnt >>= 1, n <<= 1) {
      const float c = 0.5f / (float)n;
#pragma unroll
      for (int w = 0; w < cnt / 2; ++w) {
        const float4 x = e[2 * w], y = e[2 * w + 1];
        const float d1 = x.x - y.x, d2 = x.z - y.z;
        e[w] = make_float4(x.x + y.x, (x.y + y.y) + d1 * d1 * c, x.z + y.z, (x.w + y.w) + d2 * d2 * c);
      }
    }
    s1 = e[0].x; m1 = e[0].y; s2 = e[0].z; m2 = e[0].w;
    bad = sf != 0.f;
  }
}

template <int RT, int WPR, int FEW, bool PROF>
__device__ __forceinline__ void front_role(const DecParams& p, const WsStage& st, const Ctx& c, int tid) {
  constexpr int NV = FEW / (RT * WPR);      // micro-batches in flight in the front end (visit slots)
  constexpr int NP = 4 / WPR;               // channel pairs per lane
  constexpr int CW = HD / WPR;              // channels per warp
  constexpr int NBUF = ws_nbuf(RT);
  constexpr int XSF = RT * XLayout<RT>::D;  // floats per X row (one channel of the micro-batch)
  constexpr unsigned ENTRY_BYTES = HD * XSF * sizeof(float);     // one history ring entry = one tap region of X
  const int warp = tid >> 5, lane = tid & 31;
  const int vs = warp / (RT * WPR), r = (warp / WPR) % RT, sub = warp % WPR;
  const int barid = 2 + vs * RT + r, wbase = warp - sub;
  const int cb = sub * CW + 2 * lane;       // pair i: channels cb + 64 i, cb + 64 i + 1
  const int s = c.s, part = c.part, G = c.G, B = c.B;
  const bool designated = part == 0;
  const float* lnp = c.smem + SM_LN;
  const float* g1 = lnp;
  const float* b1 = lnp + HD;
  const float* g2 = lnp + 2 * HD;
  const float* b2 = lnp + 3 * HD;
  int* pma_w = reinterpret_cast<int*>(c.smem + SM_PMA) + sub * B;     // my warp's own copy of the alignment state
  float* red = c.smem + SM_RED;
  float* kvs = c.smem + SM_KV + (vs * RT + r) * (6 * HD);
  const Word* raw_in = reinterpret_cast<const Word*>(p.ws_raw) + (size_t)c.prev * B * WS_WORDS;
  Word* raw_out = reinterpret_cast<Word*>(p.ws_raw) + (size_t)s * B * WS_WORDS;
  const int koff = st.hwy ? 2 * TAPP : 0;
  const int n_visits = p.n_steps + (s == 0 ? 1 : 0);
  const int total_visits = n_visits * G;
  const int pro = st.pro;

  long long prof_last = 0, prof_acc[7] = {0, 0, 0, 0, 0, 0, 0};
  TOT_DECL;
  const bool prof_on = PROF && p.prof != nullptr && tid == 0;
  if (prof_on) prof_last = clock64();
#define PROF_F(i)                                  \
  if (prof_on) {                                   \
    const long long now_ = clock64();              \
    prof_acc[i] += now_ - prof_last;               \
    prof_last = now_;                              \
  }

  // LayerNorm parameters of my channels (fixed for the whole launch): in registers when a lane owns 2 or 4
  // channels; with 8 channels per lane (W = 1) 32 more live registers spill, so they stay in shared memory
  constexpr bool LNREG = NP <= 2 && FEW == 4;      // eight front-end warps run on 56-72 registers (setmaxnreg)
  float2 G1[NP], B1[NP], G2[NP], B2[NP];
  if (LNREG) {
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      G1[i] = *reinterpret_cast<const float2*>(g1 + cb + 64 * i);
      B1[i] = *reinterpret_cast<const float2*>(b1 + cb + 64 * i);
      G2[i] = *reinterpret_cast<const float2*>(g2 + cb + 64 * i);
      B2[i] = *reinterpret_cast<const float2*>(b2 + cb + 64 * i);
    }
  }
  auto lnp1 = [&](int i, float2& gg, float2& bb) {
    if (LNREG) { gg = G1[i]; bb = B1[i]; }
    else { gg = *reinterpret_cast<const float2*>(g1 + cb + 64 * i); bb = *reinterpret_cast<const float2*>(b1 + cb + 64 * i); }
  };
  auto lnp2 = [&](int i, float2& gg, float2& bb) {
    if (LNREG) { gg = G2[i]; bb = B2[i]; }
    else { gg = *reinterpret_cast<const float2*>(g2 + cb + 64 * i); bb = *reinterpret_cast<const float2*>(b2 + cb + 64 * i); }
  };

  const bool issuer = (warp % (RT * WPR)) == 0;      // the visit's first warp issues its bulk copies
  bool bad = false;
  int step = 0, g = vs, vi = 0;              // G is a multiple of NV: slot vs always meets the same micro-batches
  for (int v = vs; v < total_visits; v += NV, g += NV, ++vi) {
    while (g >= G) { g -= G; ++step; }
    const int t = p.t_start + step;
    const bool final_visit = s == 0 && step == p.n_steps;
    const int tag = p.seq_base + t + 1;
    const int tag_in = s == 0 ? tag - 1 : tag;
    const bool need_wait = !(s == 0 && step == 0);
    const int q = v % NBUF;
    const int u = v / NBUF;
    float* X = c.smem + SM_X + q * (XROWS * RT * XLayout<RT>::D);
    const int row0 = g * RT;
    const bool live = row0 + r < B;
    const int b = row0 + r;
    PROF_F(0);
    if ((vi & 15) == 15) {                   // a launch aborted elsewhere: becomes row-uniform in the next reduction
      int ab = 0;
      if (lane == 0) ab = *reinterpret_cast<volatile int*>(p.abort_flag);
      if (__shfl_sync(FULL, ab, 0) != 0) bad = true;
    }

    // ---- 1. the old taps t-2d, t-d -> X rows [0, 256) and [264, 520).  The stage input of (frame, micro-batch) is
    //         kept as a ring ENTRY in X's own layout [256][XS]: the mat-vec side stores the finished current-tap rows
    //         with one bulk copy (TMA) and the taps come back the same way, straight into the X buffer -- no
    //         registers, no per-lane stores.  Only when the tap was produced by one of the last NBUF visits (tiny
    //         batches) it is still in that visit's X buffer, and my warp copies its own channel slice from there
    //         (written by this very warp: the micro-batch count is a multiple of the visit slots).  A ring entry is
    //         therefore read back more than NBUF visits after it was stored: the storing thread has seen that store
    //         complete before it released the X buffer this visit waited for (cp.async.bulk.wait_group 1 there).
    if (st.ntaps == 3) {
      if (u >= 1) {
        TOT_W0;
        if (!mbar_wait(&c.empty[q], (unsigned)(u - 1) & 1u, p.abort_flag)) bad = true;
        TOT_W1;
      }
      unsigned tx = 0;
      const float* bsrc[2] = {nullptr, nullptr};
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int back = (2 - j) * st.dil;
        const int tt = t - back;
        const int dist = back * G;
        if (tt >= p.t_start && tt >= 0 && dist <= NBUF) {      // dist == NBUF: this very buffer, not yet overwritten
          const float* srcX = c.smem + SM_X + ((v - dist) % NBUF) * (XROWS * XSF) + (size_t)koff * XSF;
          float* dstX = X + (size_t)j * TAPP * XSF;
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            const int ch = cb + 64 * i;
            if (RT == 1) {
              *reinterpret_cast<float4*>(dstX + (size_t)ch * 2) = *reinterpret_cast<const float4*>(srcX + (size_t)ch * 2);
            } else if (XLayout<RT>::D == 2) {
              *reinterpret_cast<float2*>(dstX + (size_t)ch * XSF + 2 * r) = *reinterpret_cast<const float2*>(srcX + (size_t)ch * XSF + 2 * r);
              *reinterpret_cast<float2*>(dstX + (size_t)(ch + 1) * XSF + 2 * r) = *reinterpret_cast<const float2*>(srcX + (size_t)(ch + 1) * XSF + 2 * r);
            } else {
              dstX[(size_t)ch * XSF + r] = srcX[(size_t)ch * XSF + r];
              dstX[(size_t)(ch + 1) * XSF + r] = srcX[(size_t)(ch + 1) * XSF + r];
            }
          }
        } else {
          bsrc[j] = tt < 0 ? p.ws_zero
                           : p.ws_hist + (((size_t)(st.hist_blk0 + part * st.hist_depth + tt % st.hist_depth)) * G + g) * (HD * XSF);
          tx += ENTRY_BYTES;
        }
      }
#ifdef SSV_KO_NOTAPS
      tx = 0;
#endif
      __syncwarp();
      if (lane == 0) {
        if (issuer && tx != 0) {
          mbar_arrive_expect_tx(&c.tapsfull[q], tx);
#pragma unroll
          for (int j = 0; j < 2; ++j)
            if (bsrc[j] != nullptr) bulk_g2s(X + (size_t)j * TAPP * XSF, bsrc[j], ENTRY_BYTES, &c.tapsfull[q]);
        } else {
          mbar_arrive(&c.tapsfull[q]);
        }
      }
    } else if (!final_visit && u >= 1) {
      if (!mbar_wait(&c.empty[q], (unsigned)(u - 1) & 1u, p.abort_flag)) bad = true;
    }
    PROF_F(2);

    // ---- 1b. attention stage: fetch my slice of the K / V window rows (cp.async) under the wait for the producer
    int p0 = 0, cnt = 1;
    if (pro == PRO_ATT && live) {
      p0 = pma_w[b];
      cnt = min(p0 + 2, p.N - 1) - p0 + 1;
      const float* kp = p.Kt + ((size_t)b * p.N + p0) * HD + sub * CW;
      const float* vp = p.Vt + ((size_t)b * p.N + p0) * HD + sub * CW;
      float* kd = kvs + sub * CW;
      for (int i = lane; i < cnt * (CW / 4); i += 32) {
        const int row = i / (CW / 4), cc = (i % (CW / 4)) * 4;
        cp_async16(kd + row * HD + cc, kp + (size_t)row * HD + cc, true);
        cp_async16(kd + (3 + row) * HD + cc, vp + (size_t)row * HD + cc, true);
      }
      asm volatile("cp.async.commit_group;\n" ::: "memory");
    }

    // ---- 2./3. wait for the producers of my input row, then the prologue: u_t -> X[koff ..][r]
    float2 o[NP];                                          // my slice of the stage input (joins the rings below)
#pragma unroll
    for (int i = 0; i < NP; ++i) o[i] = make_float2(0.f, 0.f);
    constexpr int XS = RT * XLayout<RT>::D;               // floats per X row
    float* xcur = X + (size_t)koff * XS + r * XLayout<RT>::D;   // channel ch at xcur[ch * XS] (D = 2: as the pair (x, x))
    auto xstore = [&](int ch, float2 val) {
      float* xp = xcur + (size_t)ch * XS;
      if (RT == 1) *reinterpret_cast<float4*>(xp) = make_float4(val.x, val.x, val.y, val.y);
      else if (XLayout<RT>::D == 2) {
        *reinterpret_cast<float2*>(xp) = make_float2(val.x, val.x);
        *reinterpret_cast<float2*>(xp + XS) = make_float2(val.y, val.y);
      } else { xp[0] = val.x; xp[XS] = val.y; }
    };
    auto xstore1 = [&](int ch, float val) {
      float* xp = xcur + (size_t)ch * XS;
      if (XLayout<RT>::D == 2) *reinterpret_cast<float2*>(xp) = make_float2(val, val);
      else xp[0] = val;
    };
    if (KO_NOFE) {
    } else if (!live) {
      if (!final_visit)
        for (int ch = lane + 32 * sub; ch < st.k_seg; ch += 32 * WPR) xstore1(ch, 0.f);
    } else if (pro == PRO_X && sub != 0) {
      // 80 mel channels: the first warp of the row does them alone
    } else {
      long long t0 = 0;
      unsigned spins = 0;
      auto spin_check = [&]() {
        // (Tried: a nanosleep of 40 / 100 / 250 ns between polls: B = 64 50.7 / 51.3 / 52.3 us/frame against 49.4-50.0, B = 256
        // 170-172 against 165.)
        if ((++spins & 255u) == 0) {
          if (t0 == 0) t0 = clock64();
          else if (clock64() - t0 > SPIN_LIMIT) atomicExch(p.abort_flag, 8);
          if (*reinterpret_cast<volatile int*>(p.abort_flag) != 0) bad = true;
        }
      };
      PROF_F(3);
      if (PROF && p.prof != nullptr && tid == 0) c.t_seen[q] = clock64();
      const Word* R = raw_in + (size_t)b * WS_WORDS;
      TOT_W0;                                                // poll time (includes the one L2 round trip of the loads)
      if (pro == PRO_X) {
        float y[3];
        if (!need_wait) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int f = lane + 32 * i;
            float xv = 0.f;
            if (f < p.F) {
              if (p.x_ext) xv = p.x_ext[(long)b * p.x_sb + (long)f * p.x_sf];
              else if (t > 0) xv = __ldcg(p.Y + ((size_t)b * p.F + f) * p.t_cap + (t - 1));
            }
            y[i] = xv;
          }
        } else {
          float vv[3] = {0.f, 0.f, 0.f};
          while (!bad) {
            bool ok = true;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              const int f = lane + 32 * i;
              vv[i] = 0.f;
              if (f < p.F) { int tg; ld_word(R + f, vv[i], tg); ok &= tg == tag_in; }
            }
            if (KO_NOPOLL || __all_sync(FULL, ok)) break;
            spin_check();
            if (__any_sync(FULL, bad)) bad = true;
          }
          TOT_W1;
          float sm = vv[0] + vv[1] + vv[2];
          sm = warp_sum(sm);
          const float mean = sm / (float)p.F;
          float qq = 0.f;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float d = vv[i] - mean;
            qq += lane + 32 * i < p.F ? d * d : 0.f;
          }
          qq = warp_sum(qq);
          const float rstd = 1.0f / sqrtf(qq / (float)p.F + 1e-5f);
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int f = lane + 32 * i;
            y[i] = f < p.F ? sigmoidf_((vv[i] - mean) * rstd * g1[f] + b1[f]) : 0.f;
            if (f < p.F && !bad && designated) p.Y[((size_t)b * p.F + f) * p.t_cap + (t - 1)] = y[i];
          }
        }
        if (!final_visit) {
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int f = lane + 32 * i;
            if (f < p.F) xstore1(f, y[i]);
          }
        }
      } else if (pro == PRO_LN || pro == PRO_LN_RELU) {
        float2 vv[NP];
        while (!bad) {
          bool ok = true;
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            int ta, tb;
            ld_word2(R + cb + 64 * i, vv[i].x, ta, vv[i].y, tb);
            ok &= ta == tag_in && tb == tag_in;
          }
          if (KO_NOPOLL || __all_sync(FULL, ok)) break;
          spin_check();
          if (__any_sync(FULL, bad)) bad = true;
        }
        TOT_W1;
        if (PROF && prof_on) PROF_F(4);
        float sum = 0.f, qq = 0.f, z0 = 0.f, z1 = 0.f;
#pragma unroll
        for (int i = 0; i < NP; ++i) sum += vv[i].x + vv[i].y;
        {
          const float lm = sum * (1.0f / (float)(2 * NP));
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            const float d0 = vv[i].x - lm, d1 = vv[i].y - lm;
            qq = fmaf(d0, d0, qq);
            qq = fmaf(d1, d1, qq);
          }
        }
#ifndef SSV_KO_NOSTATS
        row_stats2<WPR, 2 * NP>(sum, qq, z0, z1, bad, red, 2 * (vi & 1), warp, wbase, barid, lane);
#endif
        const float mean = sum / (float)HD;
        const float rstd = rstd_fast(qq / (float)HD);
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          const int ch = cb + 64 * i;
          float2 ga, ba;
          lnp1(i, ga, ba);
          float o0 = (vv[i].x - mean) * rstd * ga.x + ba.x;
          float o1 = (vv[i].y - mean) * rstd * ga.y + ba.y;
          if (pro == PRO_LN_RELU) { o0 = fmaxf(o0, 0.f); o1 = fmaxf(o1, 0.f); }
          o[i] = make_float2(o0, o1);
          xstore(ch, o[i]);
          if (st.hwy && (ch >> 5) == part)                 // my residual slice travels with my outputs
            st_word2(raw_out + (size_t)b * WS_WORDS + 2 * HD + ch, o0, o1, tag);
        }
      } else {   // PRO_HWY / PRO_ATT: the producer is a highway layer: H1 | H2 | its input (my residual)
        float2 h1[NP], h2[NP], xr[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) h1[i] = h2[i] = xr[i] = make_float2(0.f, 0.f);
        while (!bad) {
          bool ok = true;
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            int ta, tb;
            const Word* wp = R + cb + 64 * i;
            ld_word2(wp, h1[i].x, ta, h1[i].y, tb);
            ok &= ta == tag_in && tb == tag_in;
            ld_word2(wp + HD, h2[i].x, ta, h2[i].y, tb);
            ok &= ta == tag_in && tb == tag_in;
            ld_word2(wp + 2 * HD, xr[i].x, ta, xr[i].y, tb);
            ok &= ta == tag_in && tb == tag_in;
          }
          if (KO_NOPOLL || __all_sync(FULL, ok)) break;
          spin_check();
          if (__any_sync(FULL, bad)) bad = true;
        }
        TOT_W1;
        if (PROF && prof_on) PROF_F(4);
        float s1 = 0.f, s2 = 0.f, q1 = 0.f, q2 = 0.f;
#pragma unroll
        for (int i = 0; i < NP; ++i) { s1 += h1[i].x + h1[i].y; s2 += h2[i].x + h2[i].y; }
        {
          const float l1 = s1 * (1.0f / (float)(2 * NP)), l2 = s2 * (1.0f / (float)(2 * NP));
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            float d = h1[i].x - l1; q1 = fmaf(d, d, q1);
            d = h1[i].y - l1; q1 = fmaf(d, d, q1);
            d = h2[i].x - l2; q2 = fmaf(d, d, q2);
            d = h2[i].y - l2; q2 = fmaf(d, d, q2);
          }
        }
#ifndef SSV_KO_NOSTATS
        row_stats2<WPR, 2 * NP>(s1, q1, s2, q2, bad, red, 2 * (vi & 1), warp, wbase, barid, lane);
#endif
        const float m1 = s1 / (float)HD, m2 = s2 / (float)HD;
        const float r1 = rstd_fast(q1 / (float)HD);
        const float r2 = rstd_fast(q2 / (float)HD);
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          float2 ga, ba, gb, bb;
          lnp1(i, ga, ba);
          lnp2(i, gb, bb);
          const float a0 = (h1[i].x - m1) * r1 * ga.x + ba.x;
          const float a1 = (h1[i].y - m1) * r1 * ga.y + ba.y;
          const float c0 = (h2[i].x - m2) * r2 * gb.x + bb.x;
          const float c1 = (h2[i].y - m2) * r2 * gb.y + bb.y;
          const float gt0 = sigmoid_fast(a0), gt1 = sigmoid_fast(a1);
          o[i] = make_float2(gt0 * c0 + (1.0f - gt0) * xr[i].x, gt1 * c1 + (1.0f - gt1) * xr[i].y);
        }
        if (pro == PRO_HWY) {
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            const int ch = cb + 64 * i;
            xstore(ch, o[i]);
            if (st.hwy && (ch >> 5) == part) st_word2(raw_out + (size_t)b * WS_WORDS + 2 * HD + ch, o[i].x, o[i].y, tag);
          }
        } else {
          // windowed attention, models/TTSModel.py:281-295: logits over [pma, min(pma+2, N-1)];
          // every other character is masked to -2^32 and gets softmax weight exactly 0.
          asm volatile("cp.async.wait_all;\n" ::: "memory");
          __syncwarp();
          const float* kp = kvs;
          const float* vp = kvs + 3 * HD;
          float l0 = 0.f, l1 = 0.f, l2 = 0.f;
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            const int ch = cb + 64 * i;
            const float2 k0 = *reinterpret_cast<const float2*>(kp + ch);
            l0 = fmaf(k0.x, o[i].x, l0); l0 = fmaf(k0.y, o[i].y, l0);
            if (cnt > 1) {
              const float2 k1 = *reinterpret_cast<const float2*>(kp + HD + ch);
              l1 = fmaf(k1.x, o[i].x, l1); l1 = fmaf(k1.y, o[i].y, l1);
            }
            if (cnt > 2) {
              const float2 k2 = *reinterpret_cast<const float2*>(kp + 2 * HD + ch);
              l2 = fmaf(k2.x, o[i].x, l2); l2 = fmaf(k2.y, o[i].y, l2);
            }
          }
          row_reduce<WPR>(l0, l1, l2, bad, red, 2 * (vi & 1) + 1, warp, wbase, barid, lane);
          l0 *= 0.0625f; l1 *= 0.0625f; l2 *= 0.0625f;    // 1/sqrt(256)
          float mx = l0;
          if (cnt > 1) mx = fmaxf(mx, l1);
          if (cnt > 2) mx = fmaxf(mx, l2);
          const float e0 = expf(l0 - mx);
          const float e1 = cnt > 1 ? expf(l1 - mx) : 0.f;
          const float e2 = cnt > 2 ? expf(l2 - mx) : 0.f;
          const float den = e0 + e1 + e2;
          const float a0 = e0 / den, a1 = e1 / den, a2 = e2 / den;
          int best = 0;
          float bv = a0;
          if (cnt > 1 && a1 > bv) { best = 1; bv = a1; }
          if (cnt > 2 && a2 > bv) { best = 2; bv = a2; }
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            const int ch = cb + 64 * i;
            const float2 v0 = *reinterpret_cast<const float2*>(vp + ch);
            float rx = a0 * v0.x, ry = a0 * v0.y;
            if (cnt > 1) {
              const float2 v1 = *reinterpret_cast<const float2*>(vp + HD + ch);
              rx = fmaf(a1, v1.x, rx); ry = fmaf(a1, v1.y, ry);
            }
            if (cnt > 2) {
              const float2 v2 = *reinterpret_cast<const float2*>(vp + 2 * HD + ch);
              rx = fmaf(a2, v2.x, rx); ry = fmaf(a2, v2.y, ry);
            }
            xstore(ch, make_float2(rx, ry));              // R
            xstore(HD + ch, o[i]);                         // Q
          }
          __syncwarp();                                    // every lane has read the window before the next prefetch
          if (lane == 0 && !bad) {
            pma_w[b] = p0 + best;
            if (designated && sub == 0) {
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
    }
    if (final_visit) continue;                 // stage 0 after the last frame: prologue only
    PROF_F(6);
    // the mat-vec side hands the finished current-tap rows to the history ring with a bulk copy (async proxy):
    // order my generic-proxy stores before it
    if (st.ntaps == 3) fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&c.curfull[q]);              // release: my slice of the micro-batch is in X
      if (PROF && p.prof != nullptr && tid == 0) c.t_seen[8 + q] = clock64();
    }
    // leave an aborted launch only where every warp of my row agrees (after a reduction, or where there is none)
    if (bad && (WPR == 1 || pro == PRO_X || live)) return;
    PROF_F(5);
  }
  TOT_END(0, tid == 0);
  if (prof_on) {
#pragma unroll
    for (int i = 0; i < 7; ++i) p.prof[(size_t)blockIdx.x * 16 + i] = prof_acc[i];
    p.prof[(size_t)blockIdx.x * 16 + 15] = vi > 0 ? vi : 1;
  }
#undef PROF_F
}

// Register budget with eight front-end warps: 640 threads x 96 registers at launch is the CTA's pool for good
// (setmaxnreg moves registers between the warpgroups of a CTA, it cannot take more from the SM).  The two front-end
// warpgroups give registers back and the three mat-vec warpgroups, whose threads keep 88 weights each, take them:
// 256 x 72 + 384 x 112 = 61440 (no spills in either role with two or four warps per row; four-row micro-batches keep
// 18 instead of 22 weight rows per thread in registers, ws_nreg).
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
// (four-row shape, B = 256 / 192: 80 + 104 registers 165.3 / 126.1 us per frame, 64 + 112: 189 / 135, 88 + 96: 176 / 139,
// against 164.8 / 129 with 72 + 112)
template <int RT> struct Regs8 { static constexpr int FE = 72, MV = 112; };
static_assert(256 * Regs8<1>::FE + 384 * Regs8<1>::MV <= 640 * 96 && 256 * Regs8<4>::FE + 384 * Regs8<4>::MV <= 640 * 96, "register pool");

template <int RT, int WPR, int FEW, bool PROF>
__global__ void __launch_bounds__(GV_T + 32 * FEW, 1) decode_ws_kernel(const DecParams p) {
  constexpr int NT = GV_T + 32 * FEW;
  constexpr int FE_T = 32 * FEW;
  extern __shared__ __align__(16) float smem[];
  __shared__ __align__(8) uint64_t bars[3 * MAXBUF + 6];
  __shared__ int s_bad;
  __shared__ long long s_t_seen[2 * MAXBUF];
  const int tid = threadIdx.x;

  // ---- which stage / column slice am I?
  int s = -1;
#pragma unroll 1
  for (int i = 0; i < DEC_STAGES; ++i) {
    const int c0 = p.ws_stages[i].cta0;
    if ((int)blockIdx.x >= c0 && (int)blockIdx.x < c0 + p.ws_stages[i].parts) s = i;
  }
  if (s < 0) return;
  const WsStage st = p.ws_stages[s];
  Ctx c;
  c.s = s;
  c.prev = (s + DEC_STAGES - 1) % DEC_STAGES;
  c.part = (int)blockIdx.x - st.cta0;
  c.G = p.G;
  c.B = p.B;
  c.smem = smem;
  c.tapsfull = bars;
  c.curfull = bars + MAXBUF;
  c.empty = bars + 2 * MAXBUF;
  c.pfull = bars + 3 * MAXBUF;
  c.pfree = bars + 3 * MAXBUF + 2;
  c.rdone = bars + 3 * MAXBUF + 4;
  c.s_bad = &s_bad;
  c.t_seen = s_t_seen;

  // ---- one-time loads: tap-0 weights of a highway CTA, LayerNorm parameters, bias, alignment state
  {
    float* lnp = smem + SM_LN;
    {
      const int nsm = min((st.hwy ? 33 : 22) - ws_nreg(RT), st.njt);      // leading tiles of my weight image -> shared memory
      const float4* img = reinterpret_cast<const float4*>(st.img) + (size_t)c.part * st.njt * GV_T;
      float4* dst = reinterpret_cast<float4*>(smem + SM_WSM);
      for (int i = tid; i < nsm * GV_T; i += NT) dst[i] = __ldg(img + i);
    }
    for (int i = tid; i < XREGION; i += NT) smem[SM_X + i] = 0.f;       // padding rows of X stay zero for good
    for (int i = tid; i < 4 * HD; i += NT) {
      const int which = i / HD, ch = i % HD;
      const float* src = which == 0 ? st.g1 : which == 1 ? st.b1 : which == 2 ? st.g2 : st.b2;
      const int len = st.pro == PRO_X ? p.F : HD;
      lnp[i] = (src != nullptr && ch < len) ? src[ch] : 0.f;
    }
    for (int i = tid; i < 128; i += NT) {
      float bv = 0.f;
      if (i < st.ncol) { const int gc = gcol(st, c.part, i); if (gc < st.n) bv = st.bias[gc]; }
      smem[SM_BIAS + i] = bv;
    }
    if (st.pro == PRO_ATT) {
      int* pma_s = reinterpret_cast<int*>(smem + SM_PMA);
      for (int i = tid; i < p.B; i += NT) {
        const int pv = p.pma_in ? (int)p.pma_in[i] : p.pma_state[i];
        for (int w = 0; w < WPR; ++w) pma_s[w * p.B + i]      // one copy per warp of a row
          = max(0, min(pv, p.N - 1));
      }
    }
    if (tid == 0) {
      s_bad = 0;
      for (int i = 0; i < MAXBUF; ++i) {
        mbar_init(&bars[i], RT * WPR);                   // tapsfull: one arrival per front-end warp of the visit
        mbar_init(&bars[MAXBUF + i], RT * WPR);          // curfull: likewise
        mbar_init(&bars[2 * MAXBUF + i], 1);       // empty
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars[3 * MAXBUF + i], 12);                            // pfull
        const int nrw = RT * st.ncol > 128 ? RT * st.ncol / 128 : 1;     // reducer warps (as in gemv_role)
        mbar_init(&bars[3 * MAXBUF + 2 + i], nrw);                       // pfree
        mbar_init(&bars[3 * MAXBUF + 4 + i], nrw > 1 ? nrw - 1 : 1);     // rdone
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }

  // The first FEW warps are the front end, the 12 after them the mat-vec.  (Tried: the front end as the HIGHEST
  // warps, which the issue arbiter is said to prefer -- 3-4 % slower at every batch size: B = 1 31.3 -> 32.4,
  // B = 64 50.1 -> 51.9, B = 128 93.3 -> 97.8 us/frame.)
  // (Tried: the four front-end warps on one scheduler (warps 0, 4, 8, 12), the mat-vec warps on the other three:
  // B = 16 31.1 -> 34.8, B = 64 49.4 -> 51.0, B = 128 89.5 -> 98.1 us/frame.)
  // Eight front-end warps sit ABOVE the mat-vec warps (the arbiter serves the higher warp ids first): B = 256 R=4 W=1
  // 180 -> 165 us/frame.  With four front-end warps the same order is slower (B = 64 49.4 -> 50.8, B = 128 89.5 -> 93.5).
  const bool is_fe = FEW == 8 ? tid >= GV_T : tid < FE_T;
  const int rtid = FEW == 8 ? (is_fe ? tid - GV_T : tid) : (is_fe ? tid : tid - FE_T);
#ifdef SSV_WS_TOTALS
  c.t0 = clock64();
#endif
  if (is_fe) {
    if constexpr (FEW == 8) reg_dec<Regs8<RT>::FE>();
    front_role<RT, WPR, FEW, PROF>(p, st, c, rtid);
  } else {
    if constexpr (FEW == 8) reg_inc<Regs8<RT>::MV>();
    const int gtid = rtid;
    if (st.hwy) gemv_role<RT, 16, true, PROF>(p, st, c, gtid);
    else if (st.cg == 16) gemv_role<RT, 16, false, PROF>(p, st, c, gtid);
    else gemv_role<RT, 32, false, PROF>(p, st, c, gtid);
  }
}

// Weight images, thread-major: dst[part][j][gtid][c] with thread (cg, ks) = (gtid % CG, gtid / CG), column
// lc = 4 cg + c.  Highway layers: j = tap * 11 + jj covers k = tap * 256 + ks + 24 jj (zero past 256);
// plain layers: k = ks + KS j (zero past K).  W is row-major [n][K].
__global__ void ws_pack_image_kernel(const float* __restrict__ W, WsStage w, float* __restrict__ dst) {
  const long total = (long)w.parts * w.njt * GV_T * 4;
  const int KS = GV_T / w.cg;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int cc = (int)(i & 3);
    const int gtid = (int)((i >> 2) % GV_T);
    const int j = (int)(((i >> 2) / GV_T) % w.njt);
    const int part = (int)((i >> 2) / ((long)GV_T * w.njt));
    const int cg = gtid % w.cg, ks = gtid / w.cg;
    const int gc = gcol(w, part, cg * 4 + cc);
    int kk;
    bool ok;
    if (w.hwy) {
      const int tap = j / 11, kin = ks + KS * (j % 11);
      ok = kin < HD;
      kk = tap * HD + kin;
    } else {
      kk = ks + KS * j;
      ok = kk < w.K;
    }
    dst[i] = (ok && gc < w.n) ? W[(long)gc * w.K + kk] : 0.f;
  }
}

template <int RT, int WPR, int FEW, bool PROF>
int launch_cfg(const DecParams& p, cudaStream_t s) {
  constexpr size_t smem = (size_t)SM_TOTAL * sizeof(float);
  static bool configured = false;
  if (!configured) {
    SSV_CUDA(cudaFuncSetAttribute(decode_ws_kernel<RT, WPR, FEW, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  DecParams pl = p;
  void* args[] = {&pl};
  SSV_CUDA(cudaLaunchCooperativeKernel((void*)decode_ws_kernel<RT, WPR, FEW, PROF>, dim3(WS_GRID), dim3(GV_T + 32 * FEW), args, smem, s));
  ++g_launches;
  return kOk;
}

template <int RT, int WPR, int FEW>
int launch_prof(const DecParams& p, cudaStream_t s) {
  return p.prof ? launch_cfg<RT, WPR, FEW, true>(p, s) : launch_cfg<RT, WPR, FEW, false>(p, s);
}

}  // namespace

// Stage -> CTA-group layout (static).  Highway layers: 8 CTAs x 64 columns (32 H1 + the matching 32 H2);
// 1x1 convs: 128 columns per CTA, except AudioDec conv1 (K = 512): 64 columns so that a thread's weights
// still fit its registers.
void ws_stage_layout(const DecStage& d, WsStage* w) {
  w->n = d.n; w->k_seg = d.k_seg; w->ntaps = d.ntaps; w->dil = d.dil; w->pro = d.pro; w->bias_b = d.bias_b;
  w->K = d.ntaps * d.k_seg;
  w->bias = d.bias; w->g1 = d.g1; w->b1 = d.b1; w->g2 = d.g2; w->b2 = d.b2;
  w->hwy = (d.n == 2 * HD && d.ntaps == 3) ? 1 : 0;
  if (w->hwy) { w->parts = 8; w->ncol = 64; }
  else if (w->K > 256) { w->parts = (d.n + 63) / 64; w->ncol = 64; }
  else { w->parts = (d.n + 127) / 128; w->ncol = 128; }
  w->cg = w->ncol / 4;
  const int ks = GV_T / w->cg;
  w->nj = w->hwy ? 22 : (w->K + ks - 1) / ks;       // k iterations held in registers (<= 22)
  w->njt = w->hwy ? 33 : w->nj;
  w->hist_depth = d.ntaps == 3 ? 2 * d.dil + 1 : 0;
}

size_t ws_image_floats(const WsStage& w) { return (size_t)w.parts * w.njt * GV_T * 4; }

int ws_pack_image(const float* W_rowmajor, const WsStage& w, float* dst, cudaStream_t s) {
  SSV_CHECK(w.nj <= 22 && (w.hwy == 0 || w.K == 3 * HD), "decode: stage shape (n=%d, K=%d) does not fit the weight-stationary tiling", w.n, w.K);
  const long total = (long)ws_image_floats(w);
  long g = (total + 255) / 256;
  if (g > 4096) g = 4096;
  ws_pack_image_kernel<<<(int)g, 256, 0, s>>>(W_rowmajor, w, dst);
  SSV_CUDA(cudaGetLastError());
  return kOk;
}

// Shape of the front end for a batch: rows per micro-batch R, warps per row W (4 / (R W) micro-batches in flight in
// the front end), micro-batch count G (padded to a multiple of the in-flight count; padding rows are dead).
// Measured on B200 (us/frame).
void ws_plan(int B, int force_r, int force_w, int force_f, int* R, int* W, int* F, int* G) {
  int r, w, f = 4;
  if (B <= 12) { r = 1; w = 4; }           // B=1 28.3 us/frame, B=12 28.9 (W=2: 30.1)
  else if (B <= 32) { r = 1; w = 2; }      // B=16 30.5 (W=4: 32.7), B=24 31.8, B=32 37.1 (R=2 W=2: 39.3, W=1: 38.8)
  else if (B <= 120) { r = 2; w = 1; }     // B=40 45.1 (R=1: 46.6), B=48 47.4 (R=1: 56.5), B=64 51.7 (R=1 75.7, R=4 59), B=128 104
  else { r = 4; w = 1; f = 8; }            // eight front-end warps above the mat-vec warps: B=128 89.3 (R=2 F=4: 90.7), B=144 94.2 (101.7),
                                           // B=160 101.4 (115.7), B=176 113.2 (130.8), B=256 164.8 (192)
  // Eight front-end warps (round 2): only the four-row shape gains.  B=64: R=2 W=2 F=8 51.2, R=2 W=1 F=8 (four slots)
  // 72.7, R=4 W=1 F=8 88.1, R=4 W=2 F=8 68.9 against 49.4; B=128: 99.9 / 130.8 / 100.4 / 100.2 against 89.5; B=1: 33.1
  // against 29.4; B=16: 34.8 against 31.1 -- the mat-vec warps lose 16 registers each and four more warps compete for
  // the issue slots, and the front end was not the limiter (profiles/r2_decode_ws_fe8.md).
  if (force_f == 4 || force_f == 8) f = force_f;                                                     // ssv_decoder_set_plan
  if (force_r == 1 || force_r == 2 || force_r == 4) { r = force_r; if (r * w > f) w = f / r; }
  if ((force_w == 1 || force_w == 2 || force_w == 4) && r * force_w <= f) w = force_w;
  if (w * B > WS_MAX_BATCH) w = 1;
  if (f == 8 && r * w == 1) w = 2;         // eight one-warp visit slots would need more X buffers than there are
  const int nv = f / (r * w);
  int g = (B + r - 1) / r;
  g = (g + nv - 1) / nv * nv;
  *R = r; *W = w; *F = f; *G = g;
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
  SSV_CHECK(p.ws_stages && p.ws_raw && p.ws_hist, "decode: weight-stationary buffers missing");
  SSV_CHECK(p.R == 1 || p.R == 2 || p.R == 4, "decode: micro-batch rows must be 1, 2 or 4");
  SSV_CHECK(p.W == 1 || p.W == 2 || p.W == 4, "decode: warps per row must be 1, 2 or 4");
  SSV_CHECK(p.FEW == 4 || p.FEW == 8, "decode: front-end warps must be 4 or 8");
  SSV_CHECK(p.R * p.W <= p.FEW && p.G % (p.FEW / (p.R * p.W)) == 0, "decode: micro-batch count %d does not fit %d x %d of %d front-end warps", p.G, p.R, p.W, p.FEW);
  SSV_CHECK(p.W * p.B <= WS_MAX_BATCH, "decode: per-warp alignment state does not fit");
  switch (p.FEW * 64 + p.R * 8 + p.W) {
    case 4 * 64 + 1 * 8 + 1: return launch_prof<1, 1, 4>(p, s);
    case 4 * 64 + 1 * 8 + 2: return launch_prof<1, 2, 4>(p, s);
    case 4 * 64 + 1 * 8 + 4: return launch_prof<1, 4, 4>(p, s);
    case 4 * 64 + 2 * 8 + 1: return launch_prof<2, 1, 4>(p, s);
    case 4 * 64 + 2 * 8 + 2: return launch_prof<2, 2, 4>(p, s);
    case 4 * 64 + 4 * 8 + 1: return launch_prof<4, 1, 4>(p, s);
    case 8 * 64 + 1 * 8 + 2: return launch_prof<1, 2, 8>(p, s);
    case 8 * 64 + 1 * 8 + 4: return launch_prof<1, 4, 8>(p, s);
    case 8 * 64 + 2 * 8 + 1: return launch_prof<2, 1, 8>(p, s);
    case 8 * 64 + 2 * 8 + 2: return launch_prof<2, 2, 8>(p, s);
    case 8 * 64 + 2 * 8 + 4: return launch_prof<2, 4, 8>(p, s);
    case 8 * 64 + 4 * 8 + 1: return launch_prof<4, 1, 8>(p, s);
    case 8 * 64 + 4 * 8 + 2: return launch_prof<4, 2, 8>(p, s);
    default: break;
  }
  set_error("decode: unsupported front-end shape R=%d W=%d F=%d", p.R, p.W, p.FEW);
  return kInval;
}

}  // namespace ssv
