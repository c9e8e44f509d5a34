// tcgen05 / TMEM / TMA implicit-GEMM conv + LayerNorm / highway-gate kernel (BF16 operands, FP32 accumulate).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace ssv {

constexpr int TC_BM = 128;       // rows (time steps of one utterance) per CTA tile == UMMA M
constexpr int TC_BK = 64;        // K padding granule (largest k-block: one 128-byte swizzle row)

// Kernel argument block (passed as a __grid_constant__ parameter; holds the TMA descriptors).
struct alignas(64) ConvTcArgs {
  CUtensorMap tmA;        // activations, 3-D (C, T, B) bf16, box (64, 128, 1), SWIZZLE_128B
  CUtensorMap tmB0;       // weights, 2-D (K, rows) bf16, box (64, n0)
  CUtensorMap tmB1;       // weights, box (64, n1) (unused when n1 == 0)
  int T, B;
  int tiles_per_b;        // ceil(T / 128)
  int kb_per_tap;         // cin_p / 64
  int ktaps, dil, causal;
  int n0, n1;             // local accumulator blocks (each <= 256, multiple of 16); NL = n0 + n1 <= 512
  int cluster_n;          // 1 or 2 CTAs splitting N
  int w0_base, w0_rank;   // weight-row / output-column origin of block 0 = w0_base + rank * w0_rank
  int w1_base, w1_rank;   // same for block 1
  int n_real;             // LayerNorm width (real output columns; d for a highway layer)
  int epi;                // Epilogue
  int nstages;
  int bk;                 // k-block: 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B) bf16 elements
  const float* bias;      // [padded N] fp32, indexed by global column
  const float* g1; const float* b1; const float* g2; const float* b2;
  const __nv_bfloat16* Xres; long x_sb, x_st;     // residual input (highway), channels-last bf16
  void* Y; long y_sb, y_st;                       // output, channels-last (bf16 or fp32)
  int y_cols;             // columns to write per row (>= real columns: the tail is zero-filled)
  int out_fp32;
  int stage_out;          // epilogue writes through a smem staging tile + coalesced copy-out
  int stage_res;          // highway residual rows are copied to smem during the mainloop
  long long* prof;        // optional per-CTA cycle counters (SSV_TC_PROF=1)
};

// Host-side description of one packed layer for the tensor-core path.
struct TcLayer {
  __nv_bfloat16* W = nullptr;   // [rows_pad][k_pad] bf16, K contiguous (tap-major)
  float* bias = nullptr;        // [rows_pad]
  const float *g1 = nullptr, *b1 = nullptr, *g2 = nullptr, *b2 = nullptr;
  int rows = 0, rows_pad = 0;   // output columns (2d for highway)
  int cin = 0, cin_p = 0, k = 1;
  int cluster_n = 1, n0 = 0, n1 = 0, w0_base = 0, w0_rank = 0, w1_base = 0, w1_rank = 0;
  int n_real = 0;
};

int tc_pack_weights(const float* w /*[n][cin][k] fp32*/, int n, int cin, int k, int cin_p, int rows_pad,
                    __nv_bfloat16* dst, cudaStream_t s);
int tc_pack_deconv(const float* w /*[cin][cout][2]*/, int cin, int cout, __nv_bfloat16* dst, cudaStream_t s);
int tc_launch(const TcLayer& L, int epi, int dil, int causal, const __nv_bfloat16* X, int x_ld, int T, int B,
              void* Y, int y_ld, bool out_fp32, cudaStream_t s);
int tc_check_error();   // synchronous: reads and clears the device-side timeout flag
int* tc_err_flag_dev();  // the flag itself (device word), for callers that fetch it with their own asynchronous copy

// ---- second-generation BF16 kernel (conv_tc2.cu): 256 accumulator columns per CTA, two CTAs per SM, N split over a
// cluster of 1 / 2 / 4 CTAs; highway (d = 256, 512), LayerNorm (+ReLU / sigmoid) and plain (ConvTranspose1d) layers with 256
// or 512 output columns, and the 513-column heads (512 + one column on the CUDA cores); bf16 in, bf16 or fp32 out.  A launch is prepared once per (layer, buffers, shape) -- the TMA descriptors are
// encoded there -- and replayed.
struct Tc2Launch {
  alignas(64) unsigned char storage[640];
  int n_ctas = 0, cluster_n = 1;
};
bool tc2_supported(const TcLayer& L, int epi);
int tc2_prepare(const TcLayer& L, int epi, int dil, int causal, const __nv_bfloat16* X, int x_ld, int T, int B,
                void* Y, int y_ld, int out_fp32 /* 0: bf16 rows, 1: fp32 rows, 2: fp32 channels-first (B, n, T) */, Tc2Launch* out);
void tc2_set_output(Tc2Launch* L, void* Y);      // replace the output pointer of a prepared launch
int tc2_run(const Tc2Launch& L, cudaStream_t s);
int tc2_check_error();
int* tc2_err_flag_dev();
int launch_transpose_in_bf16(const float* src, long sb, long sc, long st, int B, int C, int T,
                             __nv_bfloat16* dst, int ld, cudaStream_t s);
int launch_cast_f32_to_bf16(const float* src, __nv_bfloat16* dst, size_t n, cudaStream_t s);
int launch_cast_bf16_to_f32(const __nv_bfloat16* src, float* dst, size_t n, cudaStream_t s);

}  // namespace ssv
