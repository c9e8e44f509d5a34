// Shared declarations for the spoofsv_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

namespace ssv {

// ---- error plumbing (C-ABI returns an int status; message via ssv_last_error) ----
enum Status { kOk = 0, kInval = 1, kCuda = 2, kNoMem = 3, kState = 4 };
void set_error(const char* fmt, ...);
extern thread_local long g_launches;     // kernels launched by this thread (bench "gpu_launches")

#define SSV_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ssv::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                     __FILE__, __LINE__);                                           \
      return ssv::kCuda;                                                            \
    }                                                                               \
  } while (0)

#define SSV_CHECK(cond, ...)                                                        \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      ssv::set_error(__VA_ARGS__);                                                  \
      return ssv::kInval;                                                           \
    }                                                                               \
  } while (0)

#define SSV_TRY(expr)                                                               \
  do {                                                                              \
    int _s = (expr);                                                                \
    if (_s != ssv::kOk) return _s;                                                   \
  } while (0)

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---- epilogue kinds of the fused conv kernels ----
enum Epilogue {
  EPI_NONE = 0,        // conv + bias                     (ConvTranspose1d as 1x1 GEMM)
  EPI_LN = 1,          // LayerNorm over channels
  EPI_LN_RELU = 2,     // LayerNorm, ReLU (the ReLU the reference applies to the next conv's input)
  EPI_LN_SIGMOID = 3,  // LayerNorm, sigmoid (final layer)
  EPI_HIGHWAY = 4      // split, LN1/LN2, sigma(H1)*H2 + (1-sigma(H1))*x
};

// One fused "conv1d (k taps, dilation, causal|same) -> epilogue" over channels-last rows.
// Rows are (b, t) pairs: row r -> b = r / t_rows, t = t0 + r % t_rows.
struct ConvArgs {
  const float* X;      // input activations, channels-last
  long x_sb, x_st;     // element strides between batch items / time steps
  int t_in;            // taps outside [0, t_in) read as zero
  int t0, t_rows;      // computed time range per batch item
  int M;               // total rows = B * t_rows
  const float* W;      // packed [k*cin_p][n_pad], tap-major
  const float* bias;   // [n_pad]
  const float* bias_b; // optional per-batch additive term [B][bias_b_ld] (speaker projection)
  long bias_b_ld;
  const float* g1; const float* b1;   // LN affine (LN1 of highway / the only LN)
  const float* g2; const float* b2;   // LN2 of highway
  int cin_p;           // input channels rounded up to 16
  int ktaps, dil, causal;   // causal: 0 = centred taps, 1 = taps t-(k-1)d .. t, 2 = taps t .. t+(k-1)d (fp32 kernel only)
  int n;               // real output columns (<= n_pad)
  int epi;
  float* Y;            // output, channels-last
  long y_sb, y_st;
  int y_cols;          // columns written per row (real ones computed, rest zero)
};

int launch_conv_f32(const ConvArgs& a, cudaStream_t s);

// bf16 tensor-core (tcgen05) highway / LN conv.  Same contract; activations bf16.
struct ConvTcArgs;   // defined in conv_tc.cuh

// ---- small utility kernels (misc.cu) ----
int launch_transpose_in(const float* src, long sb, long sc, long st, int B, int C, int T,
                        float* dst, int ld, cudaStream_t s);          // (B,C,T) strided -> (B,T,ld) zero padded
int launch_transpose_out(const float* src, int ld, int B, int C, int T,
                         float* dst, cudaStream_t s);                  // (B,T,ld) -> (B,C,T) contiguous
int launch_transpose_out2(const float* src, int ld, int c_off, int B, int C, int T,
                          float* dst, cudaStream_t s);                 // same, reading channels [c_off, c_off+C)
int launch_embed(const int64_t* ids, int B, int N, const float* Wt /*[vocab][E]*/, const float* bias,
                 int vocab, int E, float* dst, int ld, int* err_flag, cudaStream_t s);
int launch_pack_conv_w(const float* w /*[n][cin][k]*/, int n, int cin, int k, int cin_p, int n_pad,
                       float* dst /*[k*cin_p][n_pad]*/, cudaStream_t s);
int launch_pack_deconv_w(const float* w /*[cin][cout][2]*/, const float* b, int cin, int cout,
                         float* dst /*[cin][2*cout]*/, float* bias_dst, cudaStream_t s);
int launch_pack_rowmajor_w(const float* w /*[n][cin][k]*/, int n, int cin, int k,
                           float* dst /*[n][k*cin]*/, cudaStream_t s);
int launch_pad_vec(const float* src, int n, int n_pad, float* dst, cudaStream_t s);
int launch_linear_small(const float* x, long x_ld, const float* w, const float* b, int B, int in_f, int out_f,
                        float* y, int y_ld, cudaStream_t s);          // y[b] = W x[b] + b (speaker projections)

int launch_train_attention(const float* Kx /*(B,N,512)*/, const float* Q /*(B,T,256)*/, int B, int N, int T,
                           float* A /*(B,N,T)*/, float* RQ /*(B,T,512)*/, cudaStream_t s);
int launch_deemphasis(const float* x, float* y, int B, long n, float coeff, cudaStream_t s);   // y[n] = x[n] + c y[n-1] per row
// ---- backward building blocks (backward.cu) ----
int hwy_bwd_row_blocks(int M);
int launch_hwy_fwd_rows(const float* H, const float* X, int M, int d, const float* g1, const float* b1, const float* g2,
                        const float* b2, float* Y, cudaStream_t s);
int launch_hwy_bwd_rows(const float* H, const float* X, const float* dY, int M, int d, const float* g1, const float* b1,
                        const float* g2, const float* b2, float* dH, float* dXres, float* partial /*[blocks][6d]*/,
                        int* nblk_out, cudaStream_t s);
int launch_colsum_partials(const float* partial, int nblk, int ncol, float* out, cudaStream_t s);
struct ColSegs {          // column ranges [begin, begin + len) of a column-sum row and where each goes
  float* dst[6];
  int begin[6];
  int len[6];
  int n;
};
int launch_colsum_scatter(const float* partial, int nblk, int ncol, const ColSegs& segs, cudaStream_t s);
int launch_pack_dgrad_w(const float* conv_w /*(2d,d,k)*/, int d, int k, float* dst /*[k*2d][d]*/, cudaStream_t s);
int launch_add_inplace(float* a, const float* b, long n, cudaStream_t s);
// out[i] = sum over chunks of P[c * per + i] (split-K partial outputs; per % 4 == 0)
int launch_sum_chunks(const float* P, int chunks, long per, float* out, cudaStream_t s);
int wgrad_chunks(int M);
int launch_wgrad(const float* dH, const float* X, int M, int T, int d, int k, int dil, int causal,
                 float* P /*[chunks][k][2d][d]*/, float* dW /*(2d,d,k)*/, cudaStream_t s);

// ---- the small layers of the training graph (train_small.cu) ----
int launch_ln_rows_fwd(const float* H, int ldh, int M, int n, const float* g, const float* b, float* Y, int ldy, cudaStream_t s);
int ln_bwd_row_blocks(int M);
int ln_bwd_partial_cols(int n);                                     // columns of one block's partial: (d gamma | d beta), padded
int launch_ln_rows_bwd(const float* H, int ldh, const float* dY, int ldy, int M, int n, const float* g, float* dH,
                       float* partial /*[blocks][ln_bwd_partial_cols(n)]*/, cudaStream_t s);
int launch_utt_colsum(const float* X, int ld, int B, int T, int n, float* U /*[B][n]*/, cudaStream_t s);
int launch_relu_rows(float* x, long n, cudaStream_t s);
int launch_relu_mask(float* dx, const float* x, long n, cudaStream_t s);
int launch_pad_matrix(const float* src, int n_rows, int n_cols, int rows, int ld, float* dst, cudaStream_t s);
int wgrad_plain_chunks(int M);
int launch_wgrad_plain(const float* dH, int ldh, const float* X, int ldx, int M, int n, int cin, float* P /*[chunks][n][cin]*/,
                       float* dW /*(n, cin)*/, cudaStream_t s);
int launch_att_bwd(const float* Kx /*(B,N,512)*/, const float* Q /*(B,T,256)*/, const float* dR, int ldr, const float* dq_add,
                   const float* A, const float* dA, int B, int N, int T, float* dS /*(B,N,T)*/, float* dq /*(B,T,256)*/,
                   float* dKx /*(B,N,512)*/, cudaStream_t s);
int launch_add_utt_bias(float* h, int ld, int B, int T, int n, const float* sb /*(B, n)*/, cudaStream_t s);
int embed_bwd_scratch_floats(int vocab, int E);
int launch_embed_bwd(const int64_t* ids, int rows, const float* dX, int ld, int vocab, int E, float* scratch, float* dWt, float* dbias,
                     cudaStream_t s);
int launch_linear_small_bwd(const float* dS, int ds_ld, const float* x, long x_ld, int B, int in_f, int out_f, float* dW, float* db,
                            cudaStream_t s);

size_t griffin_lim_workspace_floats(int B, int T);                 // griffinlim.cu
int launch_griffin_lim(const float* S /*(B,513,T)*/, const float* angles0_ri /*(B,513,T,2)*/, int B, int T, int n_iter,
                       float momentum, float* y /*(B, 256 (T-1))*/, float* workspace, cudaStream_t s);
int launch_spec_features(const float* stft_ri, int F, int T, const float* melfb, int n_mels, int log_feature,
                         float norm_power, float ref_db, float max_db, int reduction, float* lin_norm,
                         float* mel_red, float* workspace, cudaStream_t s);     // |STFT| -> mel -> normalise -> reduce

}  // namespace ssv
