// Incremental Text2Mel decode: device-side stage table and launch parameters.
#pragma once
#include "common.cuh"

namespace ssv {

// How a stage builds its input vector u_t from the previous stage's raw (pre-LayerNorm) output.
enum Prologue {
  PRO_X = 0,        // u = x_t: external frame, zeros (t == 0), Y[:, :, t-1], or sigmoid(LN5(raw)) of the previous step
  PRO_LN = 1,       // u = LN(raw)
  PRO_LN_RELU = 2,  // u = relu(LN(raw))
  PRO_HWY = 3,      // u = sigma(LN1(raw[:d])) * LN2(raw[d:]) + (1 - sigma) * residual
  PRO_ATT = 4       // q = PRO_HWY; windowed attention over K/V; u = [r ; q]
};

struct DecStage {
  const float* W;      // [n][ntaps * k_seg], tap-major, K contiguous
  const float* bias;   // [n]
  const float* g1; const float* b1;   // LayerNorm affine used by the prologue (belongs to the producing layer)
  const float* g2; const float* b2;
  int n;               // outputs of this stage's GEMV (256 / 512 / 80)
  int k_seg;           // input channels per tap
  int ntaps;           // 1 or 3
  int dil;
  int pro;             // Prologue
  int n_prev;          // width of the previous stage's raw output
  int hist_in;         // history buffer holding this stage's input (taps + next stage's residual), -1 if none
  int res_hist;        // history buffer holding the residual for PRO_HWY / PRO_ATT
  int bias_b;          // 0: none, 1: speaker projection fc1, 2: fc2
};

constexpr int DEC_STAGES = 24;
constexpr int DEC_MAX_GRID = 1024;   // rows of the profiling buffer

// ---- weight-stationary pipeline (decode_ws.cu): one CTA group per stage, weights resident in registers / shared memory
struct WsStage {
  const float* img;    // [parts][njt][384] float4: thread-major images of the weight slices
  const float* bias;   // [n]
  const float* g1; const float* b1;   // LayerNorm affine of the prologue (as DecStage)
  const float* g2; const float* b2;
  int n, K, k_seg, ntaps, dil, pro, bias_b;
  int parts;           // CTAs of this stage
  int ncol;            // columns per CTA (64 / 128; padded past n with zero weights)
  int cg;              // ncol / 4: column groups of the thread tiling
  int nj;              // k iterations a mat-vec thread holds in registers (<= 22)
  int njt;             // k iterations in the image (highway: 33 = tap 0 for shared memory + nj)
  int hwy;             // 1: outputs are (H1 | H2) and the CTA owns matching slices; its input row travels along as residual
  int cta0;            // first CTA of the group
  int hist_depth;      // ring depth 2*dil + 1 of the private input history (3-tap stages), else 0
  int hist_blk0;       // first ring block of the stage (block = [G][256][RT] floats)
};
constexpr int WS_WORDS = 768;        // tagged 8-byte words per (stage, utterance): 512 outputs + 256 residual
constexpr int WS_GRID = 145;         // 16 highway layers x 8 + 17 CTAs for the eight 1x1 stages
constexpr int WS_GEMV_THREADS = 384; // warps 4-15 of a CTA
constexpr int WS_TAP_ROWS = 264;     // 256 input channels of a tap, padded to a multiple of 24 k-slices
constexpr int WS_MAX_BATCH = 1024;

struct DecParams {
  const DecStage* stages;       // [DEC_STAGES]
  const float* fin_g; const float* fin_b;   // LN5 of the decoder (80)
  const float* Kt; const float* Vt;   // [B][N][H] channels-last
  const float* s1; const float* s2;   // [B][H] hoisted speaker projections
  float* Y;                     // (B, F, t_cap) caller-owned
  float* A;                     // (B, N, t_cap) caller-owned, zeroed at begin
  long long* pma_traj;          // (t_cap, B) caller-owned
  const long long* pma_in;      // (B,) device, or nullptr -> pma_state
  int* pma_state;               // [B]
  const float* x_ext; long x_sb, x_sf;   // optional external frame for the first step of the launch
  int B, N, t_cap, F, H;
  int t_start, n_steps;
  int* abort_flag;
  long long* prof;              // optional [grid][16] phase cycle counters (SSV_DECODE_PROF=1)
  long long* totals;            // optional [grid][4]: SM cycles each role of a CTA spent in the launch, and the cycles of
                                // those it spent waiting for other CTAs / the other role (SSV_DECODE_TOTALS=1; un-instrumented kernel)
  const WsStage* ws_stages;     // device [DEC_STAGES]
  unsigned long long* ws_raw;   // [DEC_STAGES][B][WS_WORDS] tagged words {float, tag}
  float* ws_hist;               // private input-history rings: entries [256][XS] (X's own layout), [blocks][G] of them
  const float* ws_zero;         // one entry of zeros (taps before frame 0)
  int seq_base;                 // tag of frame t is seq_base + t + 1 (advanced by every begin())
  int R, G, W;                  // rows per micro-batch (1, 2, 4), micro-batches, front-end warps per row (1, 2, 4)
  int FEW;                      // front-end warps of a CTA (4 or 8)
};

int launch_decode_ws(const DecParams& p, cudaStream_t s);          // decode_ws.cu
bool decode_ws_supported(int sm_count);
// front-end shape for a batch; force_r / force_w != 0 override the measured choice (ssv_decoder_set_plan)
void ws_plan(int B, int force_r, int force_w, int force_f, int* R, int* W, int* F, int* G);
void ws_stage_layout(const DecStage& d, WsStage* w);
size_t ws_image_floats(const WsStage& w);
int ws_pack_image(const float* W_rowmajor, const WsStage& w, float* dst, cudaStream_t s);

}  // namespace ssv
