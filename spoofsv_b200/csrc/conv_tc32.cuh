// tcgen05 / TMEM / TMA implicit-GEMM conv + LayerNorm / highway-gate kernel, FP32-accurate: 3xTF32 split operands.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace ssv {

constexpr int T32_BM = 128;   // rows per CTA tile == UMMA M
constexpr int T32_BK = 16;    // fp32 elements per k-block: 64-byte rows, SWIZZLE_64B; two K = 8 MMA steps
constexpr int T32_NL = 256;   // accumulator columns per CTA = two 128-row weight blocks

// Kernel argument block (__grid_constant__; holds the TMA descriptors).
struct alignas(64) ConvTf32Args {
  CUtensorMap tmAh, tmAl;   // activations hi / lo, 3-D (C, T, B) fp32, box (16, rows_per_utt, utt_per_tile)
  CUtensorMap tmBh, tmBl;   // weights hi / lo, 2-D (K, rows) fp32, box (16, 128)
  int T, B;
  int rows_per_utt;         // 64 (two utterances per tile: T <= 64) or 128
  int utt_per_tile;         // 128 / rows_per_utt
  int tiles_per_b;          // ceil(T / 128) when utt_per_tile == 1
  int kb_per_tap;           // cin_p / 16
  int ktaps, dil, causal;
  int cluster_n;            // CTAs splitting N: 1..4
  int w0_base, w0_rank;     // weight-row / output-column origin of block 0 = w0_base + rank * w0_rank
  int w1_base, w1_rank;     // same for block 1
  int n_real;               // LayerNorm width
  int n_cols;               // real weight rows / output columns (columns beyond are padding: zero weights, zero parameters)
  int y_cols;               // columns stored (multiple of 4)
  int w_k_per_b;            // weight K origin advances by this much per "utterance" (split-K wgrad: utterance = K chunk); else 0
  int ksplit;               // EPI_NONE only: the input channels are cut into ksplit chunks, one CTA (cluster) per (tile, chunk);
                            // chunk c writes its partial sums to Yh + c * y_chunk_stride (bias in chunk 0 only)
  int w_tap_stride;         // K distance of two taps in the weight operand = the full padded Cin
  long y_chunk_stride;
  int epi;                  // Epilogue (EPI_LN, EPI_LN_RELU, EPI_LN_SIGMOID, EPI_HIGHWAY; EPI_NONE: raw conv output + bias, plain column layout)
  int nstages;
  const float* bias;        // [N] fp32, indexed by global column
  const float* g1; const float* b1; const float* g2; const float* b2;
  const float* Xh; const float* Xl; long x_sb, x_st;   // residual input of a highway layer = Xh + Xl (channels-last)
  float* Yh; float* Yl;     // output, channels-last: the TF32 split (hi, lo) for the next layer's operands, or,
                            // with Yl == nullptr, the plain fp32 value in Yh
  long y_sb, y_st;
  long long* prof;          // optional per-CTA cycle counters (SSV_TC_PROF=1)
};

// One layer packed for this path: weights [rows][taps * cin_p] fp32 (K contiguous, tap-major) split into hi / lo.
struct Tf32Layer {
  float* Wh = nullptr;
  float* Wl = nullptr;
  const float* bias = nullptr;
  const float *g1 = nullptr, *b1 = nullptr, *g2 = nullptr, *b2 = nullptr;
  int rows = 0, cin = 0, cin_p = 0, k = 1;
  int cluster_n = 1, w0_base = 0, w0_rank = 0, w1_base = 0, w1_rank = 0, n_real = 0;
  int w_cols = 0;           // K extent / row stride of the weight operand when it is not k * cin_p (split-K wgrad)
  int w_k_per_b = 0;        // see ConvTf32Args
};

void tf32_shape_highway(Tf32Layer* L, int d);     // rows = 2 d, d in {256, 512}
void tf32_shape_plain(Tf32Layer* L, int n);       // n <= 1024; three CTAs (255 padding columns) for the 513-bin heads
int tf32_pack_weights(const float* w /*[n][cin][k]*/, int n, int cin, int k, int cin_p, float* hi, float* lo, cudaStream_t s);
// dgrad operand of a highwayConv weight (2d, d, k): [d][k * 2d], taps mirrored
int tf32_pack_dgrad_weights(const float* w, int d, int k, float* hi, float* lo, cudaStream_t s);
// ConvTranspose1d(k = 2, s = 2) weight (cin, cout, 2) as the 1x1 operand [2 cout][cin]
int tf32_pack_deconv_weights(const float* w, int cin, int cout, float* hi, float* lo, cudaStream_t s);
// wgrad of a highwayConv as a split-K GEMM on the same kernel (dW[co][ci][j] = sum_r dH[r][co] X[r shifted by tap j][ci]):
//   A operand  = dH^T in K chunks, [chunk][2d][kc]   (the kernel's "activations": utterance = chunk, time step = co)
//   B operand  = shifted X^T, [(j, ci)][chunks * kc] (the kernel's "weights", K origin advancing with the chunk)
//   partials   = [chunk][2d][k * d], summed into dW (2d, d, k) by the reduce kernel
// (n_out = 2d, cin = d for a highwayConv; any n_out > 64, cin for the 1x1 layers)
int launch_wgrad_prep_dh(const float* dH, int ldh, int M, int n_out, int kc, int chunks, float* Ah, float* Al, cudaStream_t s);
int launch_wgrad_prep_x(const float* X, int x_ld, int B, int T, int cin, int k, int dil, int tap_base, int k_pad, float* Wh, float* Wl,
                        cudaStream_t s);
int launch_wgrad_tc_reduce(const float* P, int chunks, int n_out, int cin, int k, float* dW, cudaStream_t s);
// w (n, cin) -> [cin][k_p] = w^T (zero past n): the dgrad operand of a 1x1 conv
int tf32_pack_transposed(const float* w, int n, int cin, int k_p, float* hi, float* lo, cudaStream_t s);
// x (fp32, any layout, n elements) -> hi = tf32(x), lo = tf32(x - hi)
int launch_split_tf32(const float* x, float* hi, float* lo, size_t n, cudaStream_t s);
// A launch with its TMA descriptors encoded: built once per (layer, buffers, shape) and replayed (encoding four tensor
// maps per launch on the host cost more than the kernel's own prologue: 23 us per layer against 60 us of GPU time).
struct Tf32Launch {
  ConvTf32Args args;
  int n_ctas = 0, cluster_n = 1;
};
// ksplit > 1 (EPI_NONE, Yl == nullptr): split-K over the input channels; Yh then receives ksplit partial outputs,
// y_chunk_stride floats apart
int tf32_prepare_split(const Tf32Layer& L, int epi, int dil, int causal, const float* Xh, const float* Xl, int x_ld, int T, int B,
                       float* Yh, float* Yl, int y_ld, int ksplit, long y_chunk_stride, Tf32Launch* out);
// how many clusters of this size the device runs at once (one CTA per SM)
int tf32_max_clusters(int cluster_n);
int tf32_prepare(const Tf32Layer& L, int epi, int dil, int causal, const float* Xh, const float* Xl, int x_ld, int T, int B,
                 float* Yh, float* Yl, int y_ld, Tf32Launch* out);
int tf32_run(const Tf32Launch& L, cudaStream_t s);
int tf32_launch(const Tf32Layer& L, int epi, int dil, int causal, const float* Xh, const float* Xl, int x_ld, int T, int B,
                float* Yh, float* Yl, int y_ld, cudaStream_t s);    // prepare + run
int tf32_check_error();
int* tf32_err_flag_dev();

}  // namespace ssv
