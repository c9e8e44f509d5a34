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
constexpr int DEC_HIST = 16;
constexpr int DEC_RAW_LD = 512;
constexpr int DEC_XS_LD = 768;
constexpr int DEC_MAX_GRID = 1024;   // barrier flag words

struct DecParams {
  const DecStage* stages;       // [DEC_STAGES]
  const float* fin_g; const float* fin_b;   // LN5 of the decoder (80)
  float* raw;                   // [DEC_STAGES][B][DEC_RAW_LD]
  float* hist;                  // [DEC_HIST][B][t_cap][H]
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
  int RG;                       // row groups
  unsigned* bar_counter;      // [DEC_MAX_GRID] per-CTA barrier flags
  int* abort_flag;
  long long* prof;              // optional [grid][8] phase cycle counters (SSV_DECODE_PROF=1), else nullptr
};

int launch_decode(const DecParams& p, int sm_count, cudaStream_t s);
int launch_decode_cluster(const DecParams& p, cudaStream_t s);     // decode_cluster.cu
int decode_cluster_capacity();
int decode_max_grid(int* out);

}  // namespace ssv
