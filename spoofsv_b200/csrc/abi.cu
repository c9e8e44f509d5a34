// C-ABI entry points (include/spoofsv_b200.h): handles, weight repacking, layer schedules.
#include "../../include/spoofsv_b200.h"

#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "conv_tc.cuh"
#include "conv_tc32.cuh"
#include "decode.cuh"

namespace ssv {

thread_local std::string g_err;
thread_local long g_launches = 0;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

namespace {

// Per-stream scratch block behind the stream-ordered arenas: the standalone training ops make 10-20 workspace
// allocations per call and a training step makes ~70 calls; as cudaMallocAsync / cudaFreeAsync pairs that was ~2000
// driver calls per step on a host-bound path.  Calls on one stream are ordered, so consecutive calls can carve their
// workspaces out of the same block; the block grows (stream-ordered free + alloc) when a call needed more.
struct ScratchBlock {
  void* base = nullptr;
  size_t cap = 0;
  bool in_use = false;          // a live Arena owns it (a nested Arena on the same stream falls back to cudaMallocAsync)
};
inline std::mutex& scratch_mutex() { static std::mutex m; return m; }
// keyed by (device, stream): the default stream has the same handle on every device
using ScratchKey = std::pair<int, cudaStream_t>;
inline std::map<ScratchKey, ScratchBlock>& scratch_blocks() { static std::map<ScratchKey, ScratchBlock> m; return m; }
inline ScratchKey scratch_key(cudaStream_t s) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) cudaGetLastError();
  return ScratchKey(dev, s);
}

struct Arena {
  std::vector<void*> ptrs;
  cudaStream_t async_stream = nullptr;
  bool async = false;
  char* blk = nullptr;          // this stream's scratch block
  size_t blk_cap = 0, used = 0, wanted = 0;
  bool owns_blk = false;
  ScratchKey key{0, nullptr};
  Arena() = default;
  // Stream-ordered scratch (a cached block per stream, cudaMallocAsync / cudaFreeAsync on `s` beyond it): for the
  // per-call workspaces of the standalone ops, which would otherwise pay a cudaMalloc / cudaFree (= device sync) per buffer.
  explicit Arena(cudaStream_t s) : async_stream(s), async(true) {
    static bool pool_configured = false;
    if (!pool_configured) {
      int dev = 0;
      cudaMemPool_t pool;
      if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      cudaGetLastError();
      pool_configured = true;
    }
    std::lock_guard<std::mutex> lock(scratch_mutex());
    key = scratch_key(s);
    ScratchBlock& b = scratch_blocks()[key];
    if (!b.in_use) {
      b.in_use = owns_blk = true;
      blk = static_cast<char*>(b.base);
      blk_cap = b.cap;
    }
  }
  ~Arena() {
    for (void* p : ptrs) {
      if (async) cudaFreeAsync(p, async_stream);
      else cudaFree(p);
    }
    if (owns_blk) {
      std::lock_guard<std::mutex> lock(scratch_mutex());
      ScratchBlock& b = scratch_blocks()[key];
      cudaStreamCaptureStatus cap_status = cudaStreamCaptureStatusNone;
      if (wanted > blk_cap && cudaStreamIsCapturing(async_stream, &cap_status) != cudaSuccess) cudaGetLastError();
      // grow this stream's block for the next call -- never inside a stream capture: the block must outlive the graph,
      // and freeing an allocation made outside the capture is not a capturable operation (the fall-back allocations of
      // this call were captured as the graph's own memory nodes, which is fine)
      if (wanted > blk_cap && cap_status == cudaStreamCaptureStatusNone) {
        if (b.base) cudaFreeAsync(b.base, async_stream);
        b.base = nullptr;
        b.cap = 0;
        void* p = nullptr;
        const size_t cap = wanted + wanted / 4;
        if (cudaMallocAsync(&p, cap, async_stream) == cudaSuccess) { b.base = p; b.cap = cap; }
        else cudaGetLastError();
      }
      b.in_use = false;
    }
  }
  template <typename T>
  int alloc(size_t count, T** out) {
    const size_t bytes = (count * sizeof(T) + 256 + 255) & ~size_t(255);
    if (async) {
      wanted += bytes;
      if (used + bytes <= blk_cap) {
        *out = reinterpret_cast<T*>(blk + used);
        used += bytes;
        return kOk;
      }
    }
    void* p = nullptr;
    const cudaError_t e = async ? cudaMallocAsync(&p, bytes, async_stream) : cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("cudaMalloc of %zu bytes failed", count * sizeof(T));
      return kNoMem;
    }
    ptrs.push_back(p);
    *out = static_cast<T*>(p);
    return kOk;
  }
};

struct ParamMap {
  std::map<std::string, std::pair<const float*, int64_t>> m;
  int get(const std::string& name, int64_t numel, const float** out) const {
    auto it = m.find(name);
    if (it == m.end()) {
      set_error("state_dict key '%s' missing", name.c_str());
      return kInval;
    }
    if (it->second.second != numel) {
      set_error("state_dict key '%s' has %lld elements, expected %lld", name.c_str(),
                (long long)it->second.second, (long long)numel);
      return kInval;
    }
    *out = it->second.first;
    return kOk;
  }
};

int build_param_map(const char* const* names, const float* const* ptrs, const int64_t* numels, int n,
                    ParamMap* pm) {
  SSV_CHECK(names && ptrs && numels && n > 0, "empty parameter list");
  for (int i = 0; i < n; ++i) {
    SSV_CHECK(names[i] && ptrs[i], "null parameter entry %d", i);
    pm->m[names[i]] = {ptrs[i], numels[i]};
  }
  return kOk;
}

int copy_vec(Arena& ar, const ParamMap& pm, const std::string& name, int n, float** out, cudaStream_t s) {
  const float* src;
  SSV_TRY(pm.get(name, n, &src));
  SSV_TRY(ar.alloc<float>(n, out));
  SSV_CUDA(cudaMemcpyAsync(*out, src, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
  return kOk;
}

// A conv layer packed for the tiled kernels (fp32 [k*cin_p][n_pad], bf16 [n_pad][k*cin_p]).
struct ConvPack {
  float* W = nullptr;
  float* bias = nullptr;
  float *g1 = nullptr, *b1 = nullptr, *g2 = nullptr, *b2 = nullptr;
  void* Wtc = nullptr;    // bf16 copy for the tensor-core path (optional)
  int cin = 0, cin_p = 0, k = 1, n = 0, n_pad = 0;
};

int pack_conv(Arena& ar, const ParamMap& pm, const std::string& name, int n, int cin, int k, ConvPack* c,
              cudaStream_t s) {
  const float *w, *b;
  SSV_TRY(pm.get(name + ".weight", (int64_t)n * cin * k, &w));
  SSV_TRY(pm.get(name + ".bias", n, &b));
  c->cin = cin;
  c->cin_p = round_up(cin, 16);
  c->k = k;
  c->n = n;
  c->n_pad = round_up(n, 64);
  SSV_TRY(ar.alloc<float>((size_t)k * c->cin_p * c->n_pad, &c->W));
  SSV_TRY(ar.alloc<float>(c->n_pad, &c->bias));
  SSV_TRY(launch_pack_conv_w(w, n, cin, k, c->cin_p, c->n_pad, c->W, s));
  SSV_TRY(launch_pad_vec(b, n, c->n_pad, c->bias, s));
  return kOk;
}

int pack_ln(Arena& ar, const ParamMap& pm, const std::string& name, int n, float** g, float** b, cudaStream_t s) {
  SSV_TRY(copy_vec(ar, pm, name + ".weight", n, g, s));
  SSV_TRY(copy_vec(ar, pm, name + ".bias", n, b, s));
  return kOk;
}

int pack_highway(Arena& ar, const ParamMap& pm, const std::string& prefix, int d, int k, ConvPack* c,
                 cudaStream_t s) {
  SSV_TRY(pack_conv(ar, pm, prefix + ".conv", 2 * d, d, k, c, s));
  SSV_TRY(pack_ln(ar, pm, prefix + ".ln1", d, &c->g1, &c->b1, s));
  SSV_TRY(pack_ln(ar, pm, prefix + ".ln2", d, &c->g2, &c->b2, s));
  return kOk;
}

int pack_conv_ln(Arena& ar, const ParamMap& pm, const std::string& conv, const std::string& ln, int n, int cin,
                 ConvPack* c, cudaStream_t s) {
  SSV_TRY(pack_conv(ar, pm, conv, n, cin, 1, c, s));
  SSV_TRY(pack_ln(ar, pm, ln, n, &c->g1, &c->b1, s));
  return kOk;
}

// Run one packed layer over B x [t0, t0+t_rows) rows of a channels-last activation.
int run_conv(const ConvPack& c, int epi, int dil, int causal, const float* X, int x_ld, int T, int B, float* Y,
             int y_ld, cudaStream_t s, const float* bias_b = nullptr, long bias_b_ld = 0) {
  ConvArgs a{};
  a.X = X;
  a.x_sb = (long)T * x_ld;
  a.x_st = x_ld;
  a.t_in = T;
  a.t0 = 0;
  a.t_rows = T;
  a.M = B * T;
  a.W = c.W;
  a.bias = c.bias;
  a.bias_b = bias_b;
  a.bias_b_ld = bias_b_ld;
  a.g1 = c.g1; a.b1 = c.b1; a.g2 = c.g2; a.b2 = c.b2;
  a.cin_p = c.cin_p;
  a.ktaps = c.k;
  a.dil = dil;
  a.causal = causal;
  a.n = c.n;
  a.epi = epi;
  a.Y = Y;
  a.y_sb = (long)T * y_ld;
  a.y_st = y_ld;
  a.y_cols = y_ld < c.n_pad ? y_ld : c.n_pad;
  SSV_CHECK(x_ld >= c.cin_p, "activation row stride %d smaller than padded Cin %d", x_ld, c.cin_p);
  return launch_conv_f32(a, s);
}


// ---- tensor-core (bf16) packing: reuses the fp32 pack's bias / LayerNorm copies ----
int tc_pack_layer(Arena& ar, const ParamMap& pm, const std::string& wname, const ConvPack& c, int rows, int cin, int k,
                  TcLayer* L, cudaStream_t s) {
  const float* w;
  SSV_TRY(pm.get(wname, (int64_t)rows * cin * k, &w));
  L->rows = rows;
  L->cin = cin;
  L->cin_p = round_up(cin, TC_BK);
  L->k = k;
  L->bias = c.bias;
  L->g1 = c.g1; L->b1 = c.b1; L->g2 = c.g2; L->b2 = c.b2;
  SSV_TRY(ar.alloc<__nv_bfloat16>((size_t)L->rows_pad * k * L->cin_p, &L->W));
  return tc_pack_weights(w, rows, cin, k, L->cin_p, L->rows_pad, L->W, s);
}

// column split of a layer over accumulator blocks / cluster ranks
void tc_shape_highway(TcLayer* L, int d) {          // rows = 2d
  L->rows_pad = 2 * d;
  L->n_real = d;
  L->n0 = L->n1 = 256;
  L->cluster_n = d / 256;
  L->w0_base = 0; L->w0_rank = 256;
  L->w1_base = d; L->w1_rank = 256;
}
void tc_shape_plain(TcLayer* L, int n) {            // LN-only / no-epilogue layer with n output columns
  L->n_real = n;
  if (n <= 256) {
    L->rows_pad = round_up(n, 16); L->cluster_n = 1; L->n0 = L->rows_pad; L->n1 = 0;
    L->w0_base = 0; L->w0_rank = 0; L->w1_base = 0; L->w1_rank = 0;
  } else if (n <= 512) {
    L->rows_pad = round_up(n, 32); L->cluster_n = 1; L->n0 = L->rows_pad / 2; L->n1 = L->rows_pad / 2;
    L->w0_base = 0; L->w0_rank = 0; L->w1_base = L->n0; L->w1_rank = 0;
  } else {                                           // 513..1024: two CTAs, each n_loc columns in blocks (256, rest)
    const int n_loc = round_up((n + 1) / 2, 16);
    L->rows_pad = 2 * n_loc; L->cluster_n = 2;
    L->n0 = n_loc > 256 ? 256 : n_loc; L->n1 = n_loc - L->n0;
    L->w0_base = 0; L->w0_rank = n_loc; L->w1_base = L->n0; L->w1_rank = n_loc;
  }
}

// ---- tensor-core FP32-accurate (3xTF32) packing: reuses the fp32 pack's bias / LayerNorm copies ----
int tf32_pack_layer(Arena& ar, const ParamMap& pm, const std::string& wname, const ConvPack& c, int rows, int cin, int k,
                    Tf32Layer* L, cudaStream_t s) {
  const float* w;
  SSV_TRY(pm.get(wname, (int64_t)rows * cin * k, &w));
  L->rows = rows;
  L->cin = cin;
  L->cin_p = round_up(cin, T32_BK);
  L->k = k;
  L->bias = c.bias;
  L->g1 = c.g1; L->b1 = c.b1; L->g2 = c.g2; L->b2 = c.b2;
  const size_t n = (size_t)rows * k * L->cin_p;
  SSV_TRY(ar.alloc<float>(n, &L->Wh));
  SSV_TRY(ar.alloc<float>(n, &L->Wl));
  return tf32_pack_weights(w, rows, cin, k, L->cin_p, L->Wh, L->Wl, s);
}

cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }

int device_sm_count() {
  static int sm = 0;
  if (sm == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
  }
  return sm;
}

struct Workspace {
  float* buf[2] = {nullptr, nullptr};
  size_t floats = 0;
  int ensure(size_t need) {
    if (need <= floats) return kOk;
    for (int i = 0; i < 2; ++i) {
      if (buf[i]) cudaFree(buf[i]);
      buf[i] = nullptr;
    }
    floats = 0;
    for (int i = 0; i < 2; ++i) {
      if (cudaMalloc((void**)&buf[i], need * sizeof(float) + 256) != cudaSuccess) {
        cudaGetLastError();
        set_error("workspace cudaMalloc of %zu bytes failed", need * sizeof(float));
        return kNoMem;
      }
    }
    floats = need;
    return kOk;
  }
  ~Workspace() {
    for (int i = 0; i < 2; ++i)
      if (buf[i]) cudaFree(buf[i]);
  }
};

}  // namespace
}  // namespace ssv

using namespace ssv;

// ================================================================================================
struct ssv_text2mel {
  Arena arena;
  int vocab, E, temb, F, H;
  // text encoder (tiled kernels)
  float* emb_wt = nullptr;    // [vocab][temb]
  float* emb_b = nullptr;
  ConvPack te_conv1, te_conv2;
  ConvPack te_hc[12];
  int te_dil[12];
  Tf32Layer te32_conv1, te32_conv2, te32_hc[12];     // the same layers for the tensor-core (3xTF32) arm
  float* te32_ws[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t te32_floats = 0;
  Tf32Launch te32_plan[14];                           // encoded launches of the 14 layers for (te32_B, te32_N)
  int te32_B = 0, te32_N = 0;
  ~ssv_text2mel() {
    for (float* p : te32_ws)
      if (p) cudaFree(p);
    for (float* p : tr32_ws)
      if (p) cudaFree(p);
  }
  // speaker projections
  float *fc1_w, *fc1_b, *fc2_w, *fc2_b;
  // AudioEnc / AudioDec packed for the tiled kernels (train-mode full-sequence forward)
  ConvPack ae_conv1, ae_conv2, ae_conv3, ae_hc[10], ad_conv1, ad_hc[6], ad_conv2, ad_conv3, ad_conv4, ad_conv5;
  int ae_dil[10], ad_dil[6];
  Tf32Layer ae32_hc[10], ad32_hc[6];                 // the 16 causal highway layers for the tensor-core (3xTF32) arm
  float* tr32_ws[4] = {nullptr, nullptr, nullptr, nullptr};   // (hi, lo) ping-pong activations of the train-mode forward
  size_t tr32_floats = 0;
  Tf32Launch tr32_plan[16];                          // encoded launches for (tr32_B, tr32_T)
  int tr32_B = 0, tr32_T = 0;
  Workspace tws;      // train-mode activations
  float *ts1 = nullptr, *ts2 = nullptr;
  int ts_cap = 0;
  // decode stage table
  DecStage stages[DEC_STAGES];
  DecStage* stages_dev = nullptr;
  float *fin_g, *fin_b;
  // weight-stationary pipeline (decode_ws.cu): stage -> CTA-group table and the weight-slice images
  WsStage ws_stages[DEC_STAGES];
  WsStage* ws_stages_dev = nullptr;
  int ws_hist_blocks = 0;
  Workspace ws;
  int* err_flag = nullptr;
};

struct ssv_decoder {
  ssv_text2mel* m;
  Arena arena;
  int maxB, maxN, maxT;
  float *Kt, *Vt, *s1, *s2;
  int* pma_state;
  int* abort_flag;
  long long* prof = nullptr;
  long long* totals = nullptr;
  unsigned long long* ws_raw = nullptr;
  float* ws_hist = nullptr;
  float* ws_zero = nullptr;
  int seq_base = 0, R = 1, G = 1, W = 4, F = 4;
  int force_r = 0, force_w = 0, force_f = 0;     // ssv_decoder_set_plan
  // per-batch state
  bool begun = false;
  int B = 0, N = 0, t_cap = 0, t = 0;
  float *Y = nullptr, *A = nullptr;
  long long* traj = nullptr;
  // staging for ssv_synthesize_host
  Arena host_arena;
  size_t hs_key = 0;
  int64_t* h_textid = nullptr;
  float *h_spk = nullptr, *h_K = nullptr, *h_V = nullptr, *h_Y = nullptr, *h_A = nullptr;
  float* h_lin[2] = {nullptr, nullptr};   // one per in-flight batch (the D2H of batch i overlaps batch i + 1)
  __nv_bfloat16* h_lin16[2] = {nullptr, nullptr};   // bf16 copy for the half-size D2H (ssv_decoder_set_lin_output)
  int lin_out_bf16 = 0;
  long long* h_traj = nullptr;
  cudaStream_t copy_stream = nullptr;     // D2H of finished SSRN chunks
  cudaEvent_t copy_ev[2] = {nullptr, nullptr}, done_ev[2] = {nullptr, nullptr}, lin_ev[2] = {nullptr, nullptr};
  int* h_flags = nullptr;                 // pinned: {decode abort, bad text id, three tensor-core pipeline timeouts} per slot
  bool inflight[2] = {false, false}, slot_bf16[2] = {false, false};
  cudaStream_t slot_stream[2] = {nullptr, nullptr};
  ~ssv_decoder() {
    if (copy_stream) cudaStreamDestroy(copy_stream);
    for (int i = 0; i < 2; ++i) {
      if (copy_ev[i]) cudaEventDestroy(copy_ev[i]);
      if (done_ev[i]) cudaEventDestroy(done_ev[i]);
      if (lin_ev[i]) cudaEventDestroy(lin_ev[i]);
    }
    if (h_flags) cudaFreeHost(h_flags);
  }
};

struct ssv_ssrn {
  Arena arena;
  int F, O, D;
  ConvPack conv1, hc1, hc2, dc1, u1h1, u1h2, dc2, u2h1, u2h2, conv2, hc3, hc4, conv3, conv4, conv5, conv6;
  TcLayer t_conv1, t_hc1, t_hc2, t_dc1, t_u1h1, t_u1h2, t_dc2, t_u2h1, t_u2h2, t_conv2, t_hc3, t_hc4, t_conv3, t_conv4,
      t_conv5, t_conv6;
  Workspace ws;
  __nv_bfloat16* tws[2] = {nullptr, nullptr};
  size_t tws_elems = 0;
  Tc2Launch plan2[16];              // prepared launches of the layers the second-generation kernel runs, for (plan_B, plan_T)
  int plan_B = 0, plan_T = 0;
  // FP32-accurate tensor-core arm (3xTF32), all 16 layers (the 513-bin heads on three-CTA clusters, 255 padding columns)
  Tf32Layer f_conv1, f_hc1, f_hc2, f_dc1, f_u1h1, f_u1h2, f_dc2, f_u2h1, f_u2h2, f_conv2, f_hc3, f_hc4, f_conv3, f_conv4, f_conv5, f_conv6;
  float* fws[4] = {nullptr, nullptr, nullptr, nullptr};      // (hi, lo) ping-pong activations
  size_t fws_floats = 0;
  Tf32Launch fplan[16];
  int fplan_B = 0, fplan_T = 0;
  const float* fplan_out = nullptr;                           // ws.buf[0] the last plan stores into (ws may be re-grown)
  ~ssv_ssrn() {
    for (int i = 0; i < 2; ++i)
      if (tws[i]) cudaFree(tws[i]);
    for (float* p : fws)
      if (p) cudaFree(p);
  }
};

extern "C" {

int ssv_version(void) { return 100; }

const char* ssv_last_error(void) { return g_err.c_str(); }

int ssv_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  SSV_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  SSV_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return kOk;
}

long ssv_launch_count(int reset) {
  long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

// ------------------------------------------------------------------------------------------------
int ssv_highway_conv_fwd(const float* x, const float* conv_w, const float* conv_b, const float* ln1_w,
                         const float* ln1_b, const float* ln2_w, const float* ln2_b, int B, int d, int T, int k,
                         int dilation, int causal, float* y, int precision, void* stream) {
  SSV_CHECK(x && conv_w && conv_b && ln1_w && ln1_b && ln2_w && ln2_b && y, "highway_conv: null pointer");
  SSV_CHECK(B > 0 && T > 0, "highway_conv: empty input");
  SSV_CHECK(d == 256 || d == 512, "highway_conv: dimension %d unsupported (256 or 512)", d);
  SSV_CHECK(k == 1 || k == 3, "highway_conv: kernel_size %d unsupported (1 or 3)", k);
  SSV_CHECK(dilation >= 1, "highway_conv: dilation must be >= 1");
  cudaStream_t s = as_stream(stream);
  Arena ar_sync, ar_async(s);
  // CUDA-core arm: stream-ordered scratch, no host synchronisation (it sits inside training steps); the tensor-core
  // arms report asynchronous MMA-pipeline errors, so they keep the synchronous form
  Arena& ar = precision == SSV_PREC_FP32_FFMA ? ar_async : ar_sync;
  ParamMap pm;
  pm.m["conv.weight"] = {conv_w, (int64_t)2 * d * d * k};
  pm.m["conv.bias"] = {conv_b, 2 * d};
  pm.m["ln1.weight"] = {ln1_w, d};
  pm.m["ln1.bias"] = {ln1_b, d};
  pm.m["ln2.weight"] = {ln2_w, d};
  pm.m["ln2.bias"] = {ln2_b, d};
  ConvPack c;
  SSV_TRY(pack_conv(ar, pm, "conv", 2 * d, d, k, &c, s));
  SSV_TRY(pack_ln(ar, pm, "ln1", d, &c.g1, &c.b1, s));
  SSV_TRY(pack_ln(ar, pm, "ln2", d, &c.g2, &c.b2, s));
  float *xin, *yout;
  SSV_TRY(ar.alloc<float>((size_t)B * T * d, &xin));
  SSV_TRY(ar.alloc<float>((size_t)B * T * d, &yout));
  SSV_TRY(launch_transpose_in(x, (long)d * T, T, 1, B, d, T, xin, d, s));
  if (precision == SSV_PREC_BF16) {
    TcLayer L;
    tc_shape_highway(&L, d);
    SSV_TRY(tc_pack_layer(ar, pm, "conv.weight", c, 2 * d, d, k, &L, s));
    __nv_bfloat16* xb;
    SSV_TRY(ar.alloc<__nv_bfloat16>((size_t)B * T * d, &xb));
    SSV_TRY(launch_cast_f32_to_bf16(xin, xb, (size_t)B * T * d, s));
    SSV_TRY(tc_launch(L, EPI_HIGHWAY, dilation, causal, xb, d, T, B, yout, d, true, s));
  } else if (precision == SSV_PREC_FP32) {
    // FP32-accurate tensor-core arm: tcgen05 kind::tf32 with split operands (conv_tc32.cu)
    Tf32Layer L;
    tf32_shape_highway(&L, d);
    SSV_TRY(tf32_pack_layer(ar, pm, "conv.weight", c, 2 * d, d, k, &L, s));
    float *xh, *xl;
    SSV_TRY(ar.alloc<float>((size_t)B * T * d, &xh));
    SSV_TRY(ar.alloc<float>((size_t)B * T * d, &xl));
    SSV_TRY(launch_split_tf32(xin, xh, xl, (size_t)B * T * d, s));
    SSV_TRY(tf32_launch(L, EPI_HIGHWAY, dilation, causal, xh, xl, d, T, B, yout, nullptr, d, s));
  } else {
    SSV_CHECK(precision == SSV_PREC_FP32_FFMA, "highway_conv: unknown precision %d", precision);
    SSV_TRY(run_conv(c, EPI_HIGHWAY, dilation, causal, xin, d, T, B, yout, d, s));
  }
  SSV_TRY(launch_transpose_out(yout, d, B, d, T, y, s));
  if (precision == SSV_PREC_BF16 || precision == SSV_PREC_FP32) {
    SSV_CUDA(cudaStreamSynchronize(s));   // arena is freed on return
    SSV_TRY(precision == SSV_PREC_BF16 ? tc_check_error() : tf32_check_error());
  }
  return kOk;
}

// Raw conv (no epilogue) of the training graph on the tensor cores at FP32 accuracy (conv_tc32.cu: tcgen05 kind::tf32
// with split operands): X (B T x x_ld) channels-last -> Y (B T x y_ld).  Wh / Wl: the [rows][k * cin] operand split by
// tf32_pack_weights / tf32_pack_dgrad_weights.
static int tf32_raw_conv(Arena& ar, float* Wh, float* Wl, const float* bias, int rows, int cin, int k, int dil, int causal,
                         const float* X, int x_ld, int T, int B, float* Y, int y_ld, cudaStream_t s) {
  const size_t n = (size_t)B * T * x_ld;
  float *xh, *xl;
  SSV_TRY(ar.alloc<float>(n, &xh));
  SSV_TRY(ar.alloc<float>(n, &xl));
  SSV_TRY(launch_split_tf32(X, xh, xl, n, s));
  Tf32Layer L;
  L.Wh = Wh; L.Wl = Wl; L.bias = bias;
  L.rows = rows; L.cin = cin; L.cin_p = cin; L.k = k;
  tf32_shape_plain(&L, rows);
  // few row tiles (TextEnc at 64 characters, the dgrad's narrow output): cut the input channels into chunks so that the
  // launch fills the machine, and add the partial outputs up afterwards
  const int n_tiles = T <= 64 ? (B + 1) / 2 : B * ((T + T32_BM - 1) / T32_BM);
  const int room = tf32_max_clusters(L.cluster_n);
  int ksplit = 1;
  while (ksplit < 4 && n_tiles * ksplit * 2 <= room && cin % (T32_BK * ksplit * 2) == 0) ksplit *= 2;
  if (ksplit == 1) return tf32_launch(L, EPI_NONE, dil, causal, xh, xl, x_ld, T, B, Y, nullptr, y_ld, s);
  const long per = (long)B * T * y_ld;
  float* P;
  SSV_TRY(ar.alloc<float>((size_t)ksplit * per, &P));
  Tf32Launch launch;
  SSV_TRY(tf32_prepare_split(L, EPI_NONE, dil, causal, xh, xl, x_ld, T, B, P, nullptr, y_ld, ksplit, per, &launch));
  SSV_TRY(tf32_run(launch, s));
  return launch_sum_chunks(P, ksplit, per, Y, s);
}

// 4096 zero floats for the bias operand of the gradient GEMMs (allocated and cleared once)
static const float* device_zeros() {
  static float* z = nullptr;
  if (!z) {
    if (cudaMalloc((void**)&z, 4096 * sizeof(float)) != cudaSuccess || cudaMemset(z, 0, 4096 * sizeof(float)) != cudaSuccess) {
      cudaGetLastError();
      z = nullptr;
    }
  }
  return z;
}

// wgrad on the tensor cores at FP32 accuracy: a split-K GEMM on conv_tf32x3_kernel (operand layouts in conv_tc32.cuh).
// dH (B T x ldh, n_out columns) and X (B T x ldx, cin columns) channels-last -> dW (n_out, cin, k).  A highwayConv has
// n_out = 2d, cin = d; the 1x1 layers any n_out > 64 and cin.
static int tf32_wgrad(Arena& ar, const float* dH, int ldh, int n_out, const float* X, int ldx, int cin, int B, int T, int k, int dil,
                      int causal, float* dW, cudaStream_t s) {
  SSV_CHECK(n_out > 64 && cin % 4 == 0, "wgrad: shape (n %d, cin %d) not built on the tensor-core arm", n_out, cin);
  const int M = B * T, n_all = k * cin;
  const int taps_per_launch = n_all <= 1024 ? k : 1;            // N of one launch: at most four 256-column CTAs
  const int n_launch = taps_per_launch * cin;
  const int cluster_n = (n_launch + 255) / 256;
  const int tiles = (n_out + T32_BM - 1) / T32_BM;
  int chunks = tf32_max_clusters(cluster_n) / tiles;             // one wave of clusters (a cluster of three does not fit every GPC)
  if (chunks < 1) chunks = 1;
  const int kc = round_up((M + chunks - 1) / chunks, 32);
  chunks = (M + kc - 1) / kc;
  const int k_pad = chunks * kc;
  float *Ah, *Al, *Wh, *Wl, *P;
  const float* zero = device_zeros();
  SSV_CHECK(zero != nullptr, "wgrad: no memory for the zero bias");
  SSV_TRY(ar.alloc<float>((size_t)chunks * n_out * kc, &Ah));
  SSV_TRY(ar.alloc<float>((size_t)chunks * n_out * kc, &Al));
  SSV_TRY(ar.alloc<float>((size_t)n_all * k_pad, &Wh));
  SSV_TRY(ar.alloc<float>((size_t)n_all * k_pad, &Wl));
  SSV_TRY(ar.alloc<float>((size_t)chunks * n_out * n_all, &P));
  const int tap_base = causal ? -(k - 1) : -((k - 1) / 2);
  SSV_TRY(launch_wgrad_prep_dh(dH, ldh, M, n_out, kc, chunks, Ah, Al, s));
  SSV_TRY(launch_wgrad_prep_x(X, ldx, B, T, cin, k, dil, tap_base, k_pad, Wh, Wl, s));
  for (int j0 = 0; j0 < k; j0 += taps_per_launch) {
    Tf32Layer L;
    L.Wh = Wh + (size_t)j0 * cin * k_pad;
    L.Wl = Wl + (size_t)j0 * cin * k_pad;
    L.bias = zero;
    L.rows = n_launch; L.cin = kc; L.cin_p = kc; L.k = 1;
    L.w_cols = k_pad; L.w_k_per_b = kc;
    tf32_shape_plain(&L, n_launch);
    SSV_TRY(tf32_launch(L, EPI_NONE, 1, 0, Ah, Al, kc, n_out, chunks, P + (size_t)j0 * cin, nullptr, n_all, s));
  }
  return launch_wgrad_tc_reduce(P, chunks, n_out, cin, k, dW, s);
}

// Training-time forward of one highwayConv, FP32: also hands back H = conv(x) + b in the library's row layout
// ((B T) x 2d), which ssv_highway_conv_bwd takes instead of recomputing the conv.
int ssv_highway_conv_fwd_save(const float* x, const float* conv_w, const float* conv_b, const float* ln1_w,
                              const float* ln1_b, const float* ln2_w, const float* ln2_b, int B, int d, int T, int k,
                              int dilation, int causal, float* y, float* h_save, int precision, int channels_last, void* stream) {
  SSV_CHECK(x && conv_w && conv_b && ln1_w && ln1_b && ln2_w && ln2_b && y && h_save, "highway_conv_fwd_save: null pointer");
  SSV_CHECK(precision == SSV_PREC_FP32 || precision == SSV_PREC_FP32_FFMA, "highway_conv_fwd_save: precision must be SSV_PREC_FP32 (tensor cores, 3xTF32) or SSV_PREC_FP32_FFMA");
  const bool tc = precision == SSV_PREC_FP32;
  SSV_CHECK(B > 0 && T > 0, "highway_conv_fwd_save: empty input");
  SSV_CHECK(d == 256 || d == 512, "highway_conv_fwd_save: dimension %d unsupported (256 or 512)", d);
  SSV_CHECK(k == 1 || k == 3, "highway_conv_fwd_save: kernel_size %d unsupported (1 or 3)", k);
  SSV_CHECK(dilation >= 1, "highway_conv_fwd_save: dilation must be >= 1");
  cudaStream_t s = as_stream(stream);
  Arena ar(s);
  ParamMap pm;
  pm.m["conv.weight"] = {conv_w, (int64_t)2 * d * d * k};
  pm.m["conv.bias"] = {conv_b, 2 * d};
  ConvPack c;
  if (!tc) SSV_TRY(pack_conv(ar, pm, "conv", 2 * d, d, k, &c, s));
  const int M = B * T;
  // channels_last: x and y are (B, T, d), the library's own row layout -- a stack of layers passes its activations on
  // without the two boundary transposes per call
  const float* xin = x;
  float* yout = y;
  if (!channels_last) {
    float* xt;
    SSV_TRY(ar.alloc<float>((size_t)M * d, &xt));
    SSV_TRY(ar.alloc<float>((size_t)M * d, &yout));
    SSV_TRY(launch_transpose_in(x, (long)d * T, T, 1, B, d, T, xt, d, s));
    xin = xt;
  }
  if (tc) {
    float *wh, *wl;
    SSV_TRY(ar.alloc<float>((size_t)2 * d * k * d, &wh));
    SSV_TRY(ar.alloc<float>((size_t)2 * d * k * d, &wl));
    SSV_TRY(tf32_pack_weights(conv_w, 2 * d, d, k, d, wh, wl, s));
    SSV_TRY(tf32_raw_conv(ar, wh, wl, conv_b, 2 * d, d, k, dilation, causal ? 1 : 0, xin, d, T, B, h_save, 2 * d, s));
  } else {
    SSV_TRY(run_conv(c, EPI_NONE, dilation, causal ? 1 : 0, xin, d, T, B, h_save, 2 * d, s));
  }
  SSV_TRY(launch_hwy_fwd_rows(h_save, xin, M, d, ln1_w, ln1_b, ln2_w, ln2_b, yout, s));
  if (!channels_last) SSV_TRY(launch_transpose_out(yout, d, B, d, T, y, s));
  return kOk;
}

// Backward of one highwayConv, FP32 (see backward.cu).  Standalone like ssv_highway_conv_fwd: weights are packed per call.
int ssv_highway_conv_bwd(const float* x, const float* dy, const float* conv_w, const float* conv_b, const float* ln1_w,
                         const float* ln1_b, const float* ln2_w, const float* ln2_b, int B, int d, int T, int k,
                         int dilation, int causal, const float* h_saved, float* dx, float* dconv_w, float* dconv_b,
                         float* dln1_w, float* dln1_b, float* dln2_w, float* dln2_b, int precision, int channels_last, void* stream) {
  SSV_CHECK(x && dy && conv_w && conv_b && ln1_w && ln1_b && ln2_w && ln2_b, "highway_conv_bwd: null input pointer");
  SSV_CHECK(precision == SSV_PREC_FP32 || precision == SSV_PREC_FP32_FFMA, "highway_conv_bwd: precision must be SSV_PREC_FP32 (tensor cores, 3xTF32) or SSV_PREC_FP32_FFMA");
  const bool tc = precision == SSV_PREC_FP32;
  SSV_CHECK(dx && dconv_w && dconv_b && dln1_w && dln1_b && dln2_w && dln2_b, "highway_conv_bwd: null output pointer");
  SSV_CHECK(B > 0 && T > 0, "highway_conv_bwd: empty input");
  SSV_CHECK(d == 256 || d == 512, "highway_conv_bwd: dimension %d unsupported (256 or 512)", d);
  SSV_CHECK(k == 1 || k == 3, "highway_conv_bwd: kernel_size %d unsupported (1 or 3)", k);
  SSV_CHECK(dilation >= 1, "highway_conv_bwd: dilation must be >= 1");
  cudaStream_t s = as_stream(stream);
  Arena ar(s);                          // stream-ordered scratch: no host synchronisation inside a training step
  ParamMap pm;
  pm.m["conv.weight"] = {conv_w, (int64_t)2 * d * d * k};
  pm.m["conv.bias"] = {conv_b, 2 * d};
  pm.m["ln1.weight"] = {ln1_w, d};
  pm.m["ln1.bias"] = {ln1_b, d};
  pm.m["ln2.weight"] = {ln2_w, d};
  pm.m["ln2.bias"] = {ln2_b, d};
  const int M = B * T;
  float *Hbuf = nullptr, *dH, *dxr, *dxc = dx, *partial, *sums, *Wd, *P;
  const float *xin = x, *dyr = dy;      // channels_last: x, dy and dx are (B, T, d), the library's row layout
  const float* zero_bias = device_zeros();
  SSV_CHECK(zero_bias != nullptr, "highway_conv_bwd: no memory for the zero bias");
  if (!channels_last) {
    float *xt, *dyt;
    SSV_TRY(ar.alloc<float>((size_t)M * d, &xt));
    SSV_TRY(ar.alloc<float>((size_t)M * d, &dyt));
    SSV_TRY(ar.alloc<float>((size_t)M * d, &dxc));
    SSV_TRY(launch_transpose_in(x, (long)d * T, T, 1, B, d, T, xt, d, s));
    SSV_TRY(launch_transpose_in(dy, (long)d * T, T, 1, B, d, T, dyt, d, s));
    xin = xt;
    dyr = dyt;
  }
  if (!h_saved) SSV_TRY(ar.alloc<float>((size_t)M * 2 * d, &Hbuf));
  SSV_TRY(ar.alloc<float>((size_t)M * 2 * d, &dH));
  SSV_TRY(ar.alloc<float>((size_t)M * d, &dxr));
  SSV_TRY(ar.alloc<float>((size_t)hwy_bwd_row_blocks(M) * 6 * d, &partial));
  SSV_TRY(ar.alloc<float>((size_t)6 * d, &sums));
  SSV_TRY(ar.alloc<float>((size_t)k * 2 * d * d, &Wd));
  SSV_TRY(ar.alloc<float>(tc ? (size_t)4 : (size_t)wgrad_chunks(M) * k * 2 * d * d, &P));      // CUDA-core wgrad partials
  // 1. H = conv(X) + b, raw: saved by the training-time forward, or recomputed here
  const float* H = h_saved;
  if (!h_saved) {
    if (tc) {
      float *wh, *wl;
      SSV_TRY(ar.alloc<float>((size_t)2 * d * k * d, &wh));
      SSV_TRY(ar.alloc<float>((size_t)2 * d * k * d, &wl));
      SSV_TRY(tf32_pack_weights(conv_w, 2 * d, d, k, d, wh, wl, s));
      SSV_TRY(tf32_raw_conv(ar, wh, wl, conv_b, 2 * d, d, k, dilation, causal ? 1 : 0, xin, d, T, B, Hbuf, 2 * d, s));
    } else {
      ConvPack c;
      SSV_TRY(pack_conv(ar, pm, "conv", 2 * d, d, k, &c, s));
      SSV_TRY(run_conv(c, EPI_NONE, dilation, causal ? 1 : 0, xin, d, T, B, Hbuf, 2 * d, s));
    }
    H = Hbuf;
  }
  // 2. gate + LayerNorm backward per row, parameter partial sums
  int nblk = 0;
  SSV_TRY(launch_hwy_bwd_rows(H, xin, dyr, M, d, ln1_w, ln1_b, ln2_w, ln2_b, dH, dxr, partial, &nblk, s));
  {
    ColSegs sg{};
    sg.n = 5;
    float* dst[5] = {dln1_w, dln1_b, dln2_w, dln2_b, dconv_b};
    for (int i = 0; i < 5; ++i) { sg.dst[i] = dst[i]; sg.begin[i] = i * d; sg.len[i] = i == 4 ? 2 * d : d; }
    SSV_TRY(launch_colsum_scatter(partial, nblk, 6 * d, sg, s));
  }
  // 3. dgrad: the forward's conv kernel on dH with time-flipped, transposed weights and mirrored taps
  if (tc) {
    float* Wdl;
    SSV_TRY(ar.alloc<float>((size_t)k * 2 * d * d, &Wdl));
    SSV_TRY(tf32_pack_dgrad_weights(conv_w, d, k, Wd, Wdl, s));
    SSV_TRY(tf32_raw_conv(ar, Wd, Wdl, zero_bias, d, 2 * d, k, dilation, causal ? 2 : 0, dH, 2 * d, T, B, dxc, d, s));
  } else {
    SSV_TRY(launch_pack_dgrad_w(conv_w, d, k, Wd, s));
    ConvPack g;
    g.W = Wd; g.bias = const_cast<float*>(zero_bias); g.cin = 2 * d; g.cin_p = 2 * d; g.k = k; g.n = d; g.n_pad = d;
    SSV_TRY(run_conv(g, EPI_NONE, dilation, causal ? 2 : 0, dH, 2 * d, T, B, dxc, d, s));
  }
  SSV_TRY(launch_add_inplace(dxc, dxr, (long)M * d, s));
  if (!channels_last) SSV_TRY(launch_transpose_out(dxc, d, B, d, T, dx, s));
  // 4. wgrad
  if (tc) SSV_TRY(tf32_wgrad(ar, dH, 2 * d, 2 * d, xin, d, d, B, T, k, dilation, causal ? 1 : 0, dconv_w, s));
  else SSV_TRY(launch_wgrad(dH, xin, M, T, d, k, dilation, causal ? 1 : 0, P, dconv_w, s));
  return kOk;
}

// ------------------------------------------------------------------------------------------------
// The small layers of the Text2Mel training graph, FP32 (train_small.cu): standalone like the highway-conv pair above.

// y = LN(W relu?(x) + b (+ sb per utterance)): models/TTSModel.py:128-131, 173-180, 218-230.  h_save receives the raw
// conv output ((B T) x round_up(n, 64)), which ssv_conv_ln_bwd takes.
int ssv_conv_ln_fwd_save(const float* x, const float* w, const float* b, const float* sb, const float* ln_w, const float* ln_b,
                         int B, int cin, int n, int T, int relu_in, float* y, float* h_save, int precision, void* stream) {
  SSV_CHECK(x && w && b && ln_w && ln_b && y && h_save, "conv_ln_fwd_save: null pointer");
  SSV_CHECK(precision == SSV_PREC_FP32 || precision == SSV_PREC_FP32_FFMA, "conv_ln_fwd_save: precision must be SSV_PREC_FP32 (tensor cores, 3xTF32) or SSV_PREC_FP32_FFMA");
  const bool tc = precision == SSV_PREC_FP32 && cin % T32_BK == 0 && n > 64;
  SSV_CHECK(B > 0 && T > 0 && cin > 0 && n > 0 && n <= 512 && cin <= 512, "conv_ln_fwd_save: bad shape (B=%d, cin=%d, n=%d, T=%d)", B, cin, n, T);
  cudaStream_t s = as_stream(stream);
  Arena ar(s);
  ParamMap pm;
  pm.m["conv.weight"] = {w, (int64_t)n * cin};
  pm.m["conv.bias"] = {b, n};
  ConvPack c;
  SSV_TRY(pack_conv(ar, pm, "conv", n, cin, 1, &c, s));
  const int M = B * T, x_ld = round_up(cin, 64), n_pad = round_up(n, 64);
  float *xin, *yr;
  SSV_TRY(ar.alloc<float>((size_t)M * x_ld, &xin));
  SSV_TRY(ar.alloc<float>((size_t)M * n_pad, &yr));
  SSV_TRY(launch_transpose_in(x, (long)cin * T, T, 1, B, cin, T, xin, x_ld, s));
  if (relu_in) SSV_TRY(launch_relu_rows(xin, (long)M * x_ld, s));
  if (tc) {
    float *wh, *wl;
    SSV_TRY(ar.alloc<float>((size_t)n * cin, &wh));
    SSV_TRY(ar.alloc<float>((size_t)n * cin, &wl));
    SSV_TRY(tf32_pack_weights(w, n, cin, 1, cin, wh, wl, s));
    SSV_TRY(tf32_raw_conv(ar, wh, wl, b, n, cin, 1, 1, 0, xin, x_ld, T, B, h_save, n_pad, s));
    if (sb) SSV_TRY(launch_add_utt_bias(h_save, n_pad, B, T, n, sb, s));
  } else {
    SSV_TRY(run_conv(c, EPI_NONE, 1, 0, xin, x_ld, T, B, h_save, n_pad, s, sb, n));
  }
  SSV_TRY(launch_ln_rows_fwd(h_save, n_pad, M, n, ln_w, ln_b, yr, n_pad, s));
  SSV_TRY(launch_transpose_out(yr, n_pad, B, n, T, y, s));
  return kOk;
}

int ssv_conv_ln_bwd(const float* x, const float* dy, const float* w, const float* ln_w, const float* h_saved, int B, int cin, int n,
                    int T, int relu_in, float* dx, float* dw, float* db, float* dsb, float* dln_w, float* dln_b, int precision, void* stream) {
  SSV_CHECK(x && dy && w && ln_w && h_saved && dw && db && dln_w && dln_b, "conv_ln_bwd: null pointer");
  SSV_CHECK(precision == SSV_PREC_FP32 || precision == SSV_PREC_FP32_FFMA, "conv_ln_bwd: precision must be SSV_PREC_FP32 (tensor cores, 3xTF32) or SSV_PREC_FP32_FFMA");
  const bool tc = precision == SSV_PREC_FP32 && cin % T32_BK == 0 && n % T32_BK == 0 && n > 64;
  SSV_CHECK(B > 0 && T > 0 && cin > 0 && n > 0 && n <= 512 && cin <= 512, "conv_ln_bwd: bad shape (B=%d, cin=%d, n=%d, T=%d)", B, cin, n, T);
  cudaStream_t s = as_stream(stream);
  Arena ar(s);
  const int M = B * T, x_ld = round_up(cin, 64), n_pad = round_up(n, 64), pc = ln_bwd_partial_cols(n), nblk = ln_bwd_row_blocks(M);
  float *xin, *dyr, *dH, *partial, *sums, *U, *P;
  SSV_TRY(ar.alloc<float>((size_t)M * x_ld, &xin));
  SSV_TRY(ar.alloc<float>((size_t)M * n_pad, &dyr));
  SSV_TRY(ar.alloc<float>((size_t)M * n_pad, &dH));
  SSV_TRY(ar.alloc<float>((size_t)nblk * pc, &partial));
  SSV_TRY(ar.alloc<float>((size_t)pc, &sums));
  SSV_TRY(ar.alloc<float>((size_t)B * n, &U));
  SSV_TRY(ar.alloc<float>(tc ? (size_t)4 : (size_t)wgrad_plain_chunks(M) * n * cin, &P));
  SSV_TRY(launch_transpose_in(x, (long)cin * T, T, 1, B, cin, T, xin, x_ld, s));
  if (relu_in) SSV_TRY(launch_relu_rows(xin, (long)M * x_ld, s));
  SSV_TRY(launch_transpose_in(dy, (long)n * T, T, 1, B, n, T, dyr, n_pad, s));
  // LayerNorm backward per row, parameter sums
  SSV_TRY(launch_ln_rows_bwd(h_saved, n_pad, dyr, n_pad, M, n, ln_w, dH, partial, s));
  {
    ColSegs sg{};
    sg.n = 2;
    sg.dst[0] = dln_w; sg.begin[0] = 0; sg.len[0] = n;
    sg.dst[1] = dln_b; sg.begin[1] = pc / 2; sg.len[1] = n;
    SSV_TRY(launch_colsum_scatter(partial, nblk, pc, sg, s));
  }
  // bias gradients: per utterance (the speaker projection's), summed over the utterances for the conv bias
  SSV_TRY(launch_utt_colsum(dH, n_pad, B, T, n, U, s));
  SSV_TRY(launch_colsum_partials(U, B, n, db, s));
  if (dsb) SSV_CUDA(cudaMemcpyAsync(dsb, U, sizeof(float) * B * n, cudaMemcpyDeviceToDevice, s));
  // dgrad: dX = dH W through the forward's conv kernel (W as the [K = n][N = cin] operand)
  if (dx) {
    const int k_p = round_up(n, 16);
    float *Wd, *dxr;
    const float* zero_bias = device_zeros();
    SSV_CHECK(zero_bias != nullptr, "conv_ln_bwd: no memory for the zero bias");
    SSV_TRY(ar.alloc<float>((size_t)k_p * x_ld, &Wd));
    SSV_TRY(ar.alloc<float>((size_t)M * x_ld, &dxr));
    if (tc) {
      float *wdh, *wdl;
      SSV_TRY(ar.alloc<float>((size_t)cin * k_p, &wdh));
      SSV_TRY(ar.alloc<float>((size_t)cin * k_p, &wdl));
      SSV_TRY(tf32_pack_transposed(w, n, cin, k_p, wdh, wdl, s));
      SSV_TRY(tf32_raw_conv(ar, wdh, wdl, zero_bias, cin, k_p, 1, 1, 0, dH, n_pad, T, B, dxr, x_ld, s));
    } else {
      SSV_TRY(launch_pad_matrix(w, n, cin, k_p, x_ld, Wd, s));
      ConvPack g;
      g.W = Wd; g.bias = const_cast<float*>(zero_bias); g.cin = n; g.cin_p = k_p; g.k = 1; g.n = cin; g.n_pad = x_ld;
      SSV_TRY(run_conv(g, EPI_NONE, 1, 0, dH, n_pad, T, B, dxr, x_ld, s));
    }
    if (relu_in) SSV_TRY(launch_relu_mask(dxr, xin, (long)M * x_ld, s));
    SSV_TRY(launch_transpose_out(dxr, x_ld, B, cin, T, dx, s));
  }
  if (tc) SSV_TRY(tf32_wgrad(ar, dH, n_pad, n, xin, x_ld, cin, B, T, 1, 1, 0, dw, s));
  else SSV_TRY(launch_wgrad_plain(dH, n_pad, xin, x_ld, M, n, cin, P, dw, s));
  return kOk;
}

// Unmasked softmax attention of the train branch (models/TTSModel.py:268-272): kv (B, 512, N) = [K ; V] as TextEnc emits
// them, q (B, 256, T) -> A (B, N, T), rq (B, 512, T) = [V A ; q].
int ssv_attention_train_fwd(const float* kv, const float* q, int B, int N, int T, float* A, float* rq, void* stream) {
  SSV_CHECK(kv && q && A && rq, "attention_train_fwd: null pointer");
  SSV_CHECK(B > 0 && N > 0 && T > 0, "attention_train_fwd: empty input");
  cudaStream_t s = as_stream(stream);
  Arena ar(s);
  float *Kx, *Qt, *RQ;
  SSV_TRY(ar.alloc<float>((size_t)B * N * 512, &Kx));
  SSV_TRY(ar.alloc<float>((size_t)B * T * 256, &Qt));
  SSV_TRY(ar.alloc<float>((size_t)B * T * 512, &RQ));
  SSV_TRY(launch_transpose_in(kv, (long)512 * N, N, 1, B, 512, N, Kx, 512, s));
  SSV_TRY(launch_transpose_in(q, (long)256 * T, T, 1, B, 256, T, Qt, 256, s));
  SSV_TRY(launch_train_attention(Kx, Qt, B, N, T, A, RQ, s));
  SSV_TRY(launch_transpose_out(RQ, 512, B, 512, T, rq, s));
  return kOk;
}

// dA may be null (no loss term on the alignment).  drq is the gradient of [R ; q]; dq collects both of q's paths.
int ssv_attention_train_bwd(const float* kv, const float* q, const float* A, const float* dA, const float* drq, int B, int N, int T,
                            float* dkv, float* dq, void* stream) {
  SSV_CHECK(kv && q && A && drq && dkv && dq, "attention_train_bwd: null pointer");
  SSV_CHECK(B > 0 && N > 0 && T > 0, "attention_train_bwd: empty input");
  cudaStream_t s = as_stream(stream);
  Arena ar(s);
  float *Kx, *Qt, *dRQ, *dS, *dqr, *dKx;
  SSV_TRY(ar.alloc<float>((size_t)B * N * 512, &Kx));
  SSV_TRY(ar.alloc<float>((size_t)B * T * 256, &Qt));
  SSV_TRY(ar.alloc<float>((size_t)B * T * 512, &dRQ));
  SSV_TRY(ar.alloc<float>((size_t)B * N * T, &dS));
  SSV_TRY(ar.alloc<float>((size_t)B * T * 256, &dqr));
  SSV_TRY(ar.alloc<float>((size_t)B * N * 512, &dKx));
  SSV_TRY(launch_transpose_in(kv, (long)512 * N, N, 1, B, 512, N, Kx, 512, s));
  SSV_TRY(launch_transpose_in(q, (long)256 * T, T, 1, B, 256, T, Qt, 256, s));
  SSV_TRY(launch_transpose_in(drq, (long)512 * T, T, 1, B, 512, T, dRQ, 512, s));
  SSV_TRY(launch_att_bwd(Kx, Qt, dRQ, 512, dRQ + 256, A, dA, B, N, T, dS, dqr, dKx, s));
  SSV_TRY(launch_transpose_out(dKx, 512, B, 512, N, dkv, s));
  SSV_TRY(launch_transpose_out(dqr, 256, B, 256, T, dq, s));
  return kOk;
}

// textEmbedding (models/TTSModel.py:25-35) as a column gather: ids (B, N), weight (E, vocab), bias (E) -> y (B, E, N)
int ssv_text_embedding_fwd(const int64_t* ids, const float* weight, const float* bias, int B, int N, int vocab, int E, float* y,
                           void* stream) {
  SSV_CHECK(ids && weight && bias && y, "text_embedding_fwd: null pointer");
  SSV_CHECK(B > 0 && N > 0 && vocab > 0 && E > 0, "text_embedding_fwd: empty input");
  cudaStream_t s = as_stream(stream);
  Arena ar(s);
  float *Wt, *rows;
  int* flag;
  SSV_TRY(ar.alloc<float>((size_t)vocab * E, &Wt));
  SSV_TRY(ar.alloc<float>((size_t)B * N * E, &rows));
  SSV_TRY(ar.alloc<int>(1, &flag));
  SSV_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), s));
  SSV_TRY(launch_transpose_in(weight, 0, vocab, 1, 1, E, vocab, Wt, E, s));        // (E, vocab) -> (vocab, E)
  SSV_TRY(launch_embed(ids, B, N, Wt, bias, vocab, E, rows, E, flag, s));
  SSV_TRY(launch_transpose_out(rows, E, B, E, N, y, s));
  return kOk;
}

int ssv_text_embedding_bwd(const int64_t* ids, const float* dy, int B, int N, int vocab, int E, float* dweight, float* dbias,
                           void* stream) {
  SSV_CHECK(ids && dy && dweight && dbias, "text_embedding_bwd: null pointer");
  SSV_CHECK(B > 0 && N > 0 && vocab > 0 && E > 0, "text_embedding_bwd: empty input");
  cudaStream_t s = as_stream(stream);
  Arena ar(s);
  float *rows, *dWt, *scratch;
  SSV_TRY(ar.alloc<float>((size_t)B * N * E, &rows));
  SSV_TRY(ar.alloc<float>((size_t)vocab * E, &dWt));
  SSV_TRY(ar.alloc<float>((size_t)embed_bwd_scratch_floats(vocab, E), &scratch));
  SSV_TRY(launch_transpose_in(dy, (long)E * N, N, 1, B, E, N, rows, E, s));
  SSV_TRY(launch_embed_bwd(ids, B * N, rows, E, vocab, E, scratch, dWt, dbias, s));
  SSV_TRY(launch_transpose_out(dWt, E, 1, E, vocab, dweight, s));                   // (vocab, E) -> (E, vocab)
  return kOk;
}

// Speaker projections fc1 / fc2 (models/TTSModel.py:172-173): y (B, out) = x (B, in) W^T + b, and the parameter gradients
int ssv_linear_small_fwd(const float* x, const float* w, const float* b, int B, int in_f, int out_f, float* y, void* stream) {
  SSV_CHECK(x && w && b && y && B > 0 && in_f > 0 && out_f > 0, "linear_small_fwd: bad arguments");
  return launch_linear_small(x, in_f, w, b, B, in_f, out_f, y, out_f, as_stream(stream));
}
int ssv_linear_small_bwd(const float* x, const float* dy, int B, int in_f, int out_f, float* dw, float* db, void* stream) {
  SSV_CHECK(x && dy && dw && db && B > 0 && in_f > 0 && out_f > 0, "linear_small_bwd: bad arguments");
  return launch_linear_small_bwd(dy, out_f, x, in_f, B, in_f, out_f, dw, db, as_stream(stream));
}

// ------------------------------------------------------------------------------------------------
static int fill_stage(ssv_text2mel* m, const ParamMap& pm, int idx, const std::string& conv, int n, int cin, int k,
                      int dil, int pro, const float* g1, const float* b1, const float* g2, const float* b2,
                      int hist_in, int res_hist, int bias_b, cudaStream_t s) {
  const float *w, *b;
  SSV_TRY(pm.get(conv + ".weight", (int64_t)n * cin * k, &w));
  SSV_TRY(pm.get(conv + ".bias", n, &b));
  float *wd, *bd;
  SSV_TRY(m->arena.alloc<float>((size_t)n * cin * k, &wd));
  SSV_TRY(m->arena.alloc<float>(n, &bd));
  SSV_TRY(launch_pack_rowmajor_w(w, n, cin, k, wd, s));
  SSV_CUDA(cudaMemcpyAsync(bd, b, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
  DecStage& st = m->stages[idx];
  st.W = wd; st.bias = bd;
  st.g1 = g1; st.b1 = b1; st.g2 = g2; st.b2 = b2;
  st.n = n; st.k_seg = cin; st.ntaps = k; st.dil = dil; st.pro = pro;
  st.n_prev = 0; st.hist_in = hist_in; st.res_hist = res_hist; st.bias_b = bias_b;
  return kOk;
}

int ssv_text2mel_create(const char* const* names, const float* const* dev_ptrs, const int64_t* numels,
                        int n_params, int vocab_len, int spkemb_dim, int textemb_dim, int freq_bins,
                        int hidden_dim, ssv_text2mel** out) {
  SSV_CHECK(out, "text2mel_create: null out");
  SSV_CHECK(hidden_dim == 256, "text2mel_create: hidden_dim %d unsupported (256)", hidden_dim);
  SSV_CHECK(textemb_dim % 16 == 0 && textemb_dim <= 512, "text2mel_create: textemb_dim must be a multiple of 16");
  SSV_CHECK(freq_bins % 16 == 0 && freq_bins <= 96, "text2mel_create: freq_bins must be a multiple of 16, <= 96");
  SSV_CHECK(vocab_len > 0 && spkemb_dim > 0, "text2mel_create: bad dims");
  ParamMap pm;
  SSV_TRY(build_param_map(names, dev_ptrs, numels, n_params, &pm));
  ssv_text2mel* m = new ssv_text2mel();
  m->vocab = vocab_len; m->E = spkemb_dim; m->temb = textemb_dim; m->F = freq_bins; m->H = hidden_dim;
  cudaStream_t s = nullptr;
  const int H = hidden_dim, D2 = 2 * hidden_dim;
  auto fail = [&](int code) { delete m; return code; };
#define T2M_TRY(expr) do { int _s = (expr); if (_s != kOk) return fail(_s); } while (0)
  // ---- text encoder
  {
    const float *w, *b;
    T2M_TRY(pm.get("text_encoder.textemb_layer.W.weight", (int64_t)textemb_dim * vocab_len, &w));
    T2M_TRY(pm.get("text_encoder.textemb_layer.W.bias", textemb_dim, &b));
    T2M_TRY(m->arena.alloc<float>((size_t)vocab_len * textemb_dim, &m->emb_wt));
    T2M_TRY(m->arena.alloc<float>(textemb_dim, &m->emb_b));
    // (E, vocab) -> [vocab][E]: a k=1 "rowmajor" pack of w viewed as [n=E][cin=vocab] gives [E][vocab];
    // we need the transpose, which pack_conv_w provides: dst[kk=ci][col] with n=E, cin=vocab.
    T2M_TRY(launch_pack_conv_w(w, textemb_dim, vocab_len, 1, vocab_len, textemb_dim, m->emb_wt, s));
    if (cudaMemcpyAsync(m->emb_b, b, sizeof(float) * textemb_dim, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
      return fail(kCuda);
    T2M_TRY(m->arena.alloc<int>(1, &m->err_flag));
    cudaMemsetAsync(m->err_flag, 0, sizeof(int), s);
  }
  T2M_TRY(pack_conv_ln(m->arena, pm, "text_encoder.conv1", "text_encoder.ln1", D2, textemb_dim, &m->te_conv1, s));
  T2M_TRY(pack_conv_ln(m->arena, pm, "text_encoder.conv2", "text_encoder.ln2", D2, D2, &m->te_conv2, s));
  {
    const char* nm[12] = {"hci1.hc1", "hci1.hc2", "hci1.hc3", "hci1.hc4", "hci2.hc1", "hci2.hc2",
                          "hci2.hc3", "hci2.hc4", "hc1", "hc2", "hc3", "hc4"};
    const int dil[12] = {1, 3, 9, 27, 1, 3, 9, 27, 1, 1, 1, 1};
    const int ks[12] = {3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 1, 1};
    for (int i = 0; i < 12; ++i) {
      T2M_TRY(pack_highway(m->arena, pm, std::string("text_encoder.") + nm[i], D2, ks[i], &m->te_hc[i], s));
      m->te_dil[i] = dil[i];
      tf32_shape_highway(&m->te32_hc[i], D2);
      T2M_TRY(tf32_pack_layer(m->arena, pm, std::string("text_encoder.") + nm[i] + ".conv.weight", m->te_hc[i], 2 * D2, D2, ks[i],
                              &m->te32_hc[i], s));
    }
    tf32_shape_plain(&m->te32_conv1, D2);
    T2M_TRY(tf32_pack_layer(m->arena, pm, "text_encoder.conv1.weight", m->te_conv1, D2, textemb_dim, 1, &m->te32_conv1, s));
    tf32_shape_plain(&m->te32_conv2, D2);
    T2M_TRY(tf32_pack_layer(m->arena, pm, "text_encoder.conv2.weight", m->te_conv2, D2, D2, 1, &m->te32_conv2, s));
  }
  // ---- speaker projections
  T2M_TRY(copy_vec(m->arena, pm, "audio_encoder.fc1.weight", H * spkemb_dim, &m->fc1_w, s));
  T2M_TRY(copy_vec(m->arena, pm, "audio_encoder.fc1.bias", H, &m->fc1_b, s));
  T2M_TRY(copy_vec(m->arena, pm, "audio_encoder.fc2.weight", H * spkemb_dim, &m->fc2_w, s));
  T2M_TRY(copy_vec(m->arena, pm, "audio_encoder.fc2.bias", H, &m->fc2_b, s));
  // ---- AudioEnc / AudioDec for the tiled kernels (train branch)
  {
    const std::string ae = "audio_encoder.", ad = "audio_decoder.";
    T2M_TRY(pack_conv_ln(m->arena, pm, ae + "conv1", ae + "ln1", H, freq_bins, &m->ae_conv1, s));
    T2M_TRY(pack_conv_ln(m->arena, pm, ae + "conv2", ae + "ln2", H, H, &m->ae_conv2, s));
    T2M_TRY(pack_conv_ln(m->arena, pm, ae + "conv3", ae + "ln3", H, H, &m->ae_conv3, s));
    const char* eh[10] = {"hci1.hc1", "hci1.hc2", "hci1.hc3", "hci1.hc4", "hci2.hc1", "hci2.hc2", "hci2.hc3", "hci2.hc4", "hc1", "hc2"};
    const int ed[10] = {1, 3, 9, 27, 1, 3, 9, 27, 3, 3};
    for (int i = 0; i < 10; ++i) {
      T2M_TRY(pack_highway(m->arena, pm, ae + eh[i], H, 3, &m->ae_hc[i], s));
      m->ae_dil[i] = ed[i];
      tf32_shape_highway(&m->ae32_hc[i], H);
      T2M_TRY(tf32_pack_layer(m->arena, pm, ae + eh[i] + ".conv.weight", m->ae_hc[i], 2 * H, H, 3, &m->ae32_hc[i], s));
    }
    T2M_TRY(pack_conv_ln(m->arena, pm, ad + "conv1", ad + "ln1", H, D2, &m->ad_conv1, s));
    const char* dh[6] = {"hci.hc1", "hci.hc2", "hci.hc3", "hci.hc4", "hc1", "hc2"};
    const int dd[6] = {1, 3, 9, 27, 1, 1};
    for (int i = 0; i < 6; ++i) {
      T2M_TRY(pack_highway(m->arena, pm, ad + dh[i], H, 3, &m->ad_hc[i], s));
      m->ad_dil[i] = dd[i];
      tf32_shape_highway(&m->ad32_hc[i], H);
      T2M_TRY(tf32_pack_layer(m->arena, pm, ad + dh[i] + ".conv.weight", m->ad_hc[i], 2 * H, H, 3, &m->ad32_hc[i], s));
    }
    T2M_TRY(pack_conv_ln(m->arena, pm, ad + "conv2", ad + "ln2", H, H, &m->ad_conv2, s));
    T2M_TRY(pack_conv_ln(m->arena, pm, ad + "conv3", ad + "ln3", H, H, &m->ad_conv3, s));
    T2M_TRY(pack_conv_ln(m->arena, pm, ad + "conv4", ad + "ln4", H, H, &m->ad_conv4, s));
    T2M_TRY(pack_conv_ln(m->arena, pm, ad + "conv5", ad + "ln5", freq_bins, H, &m->ad_conv5, s));
  }
  // ---- decode stage table
  {
    struct Ln { float *g, *b; };
    auto ln = [&](const std::string& name, int n, Ln* o) { return pack_ln(m->arena, pm, name, n, &o->g, &o->b, s); };
    const std::string ae = "audio_encoder.", ad = "audio_decoder.";
    Ln e1, e2, e3, d1, d2, d3, d4, d5;
    T2M_TRY(ln(ae + "ln1", H, &e1)); T2M_TRY(ln(ae + "ln2", H, &e2)); T2M_TRY(ln(ae + "ln3", H, &e3));
    T2M_TRY(ln(ad + "ln1", H, &d1)); T2M_TRY(ln(ad + "ln2", H, &d2)); T2M_TRY(ln(ad + "ln3", H, &d3));
    T2M_TRY(ln(ad + "ln4", H, &d4)); T2M_TRY(ln(ad + "ln5", freq_bins, &d5));
    m->fin_g = d5.g; m->fin_b = d5.b;
    const char* enc_hc[10] = {"hci1.hc1", "hci1.hc2", "hci1.hc3", "hci1.hc4", "hci2.hc1",
                              "hci2.hc2", "hci2.hc3", "hci2.hc4", "hc1", "hc2"};
    const int enc_dil[10] = {1, 3, 9, 27, 1, 3, 9, 27, 3, 3};
    const char* dec_hc[6] = {"hci.hc1", "hci.hc2", "hci.hc3", "hci.hc4", "hc1", "hc2"};
    const int dec_dil[6] = {1, 3, 9, 27, 1, 1};
    Ln eh1[10], eh2[10], dh1[6], dh2[6];
    for (int i = 0; i < 10; ++i) {
      T2M_TRY(ln(ae + enc_hc[i] + ".ln1", H, &eh1[i]));
      T2M_TRY(ln(ae + enc_hc[i] + ".ln2", H, &eh2[i]));
    }
    for (int i = 0; i < 6; ++i) {
      T2M_TRY(ln(ad + dec_hc[i] + ".ln1", H, &dh1[i]));
      T2M_TRY(ln(ad + dec_hc[i] + ".ln2", H, &dh2[i]));
    }
    int si = 0;
    T2M_TRY(fill_stage(m, pm, si++, ae + "conv1", H, freq_bins, 1, 1, PRO_X, nullptr, nullptr, nullptr, nullptr, -1, -1, 1, s));
    T2M_TRY(fill_stage(m, pm, si++, ae + "conv2", H, H, 1, 1, PRO_LN_RELU, e1.g, e1.b, nullptr, nullptr, -1, -1, 0, s));
    T2M_TRY(fill_stage(m, pm, si++, ae + "conv3", H, H, 1, 1, PRO_LN_RELU, e2.g, e2.b, nullptr, nullptr, -1, -1, 2, s));
    for (int i = 0; i < 10; ++i) {
      if (i == 0)
        T2M_TRY(fill_stage(m, pm, si++, ae + enc_hc[i] + ".conv", D2, H, 3, enc_dil[i], PRO_LN, e3.g, e3.b, nullptr, nullptr, 0, -1, 0, s));
      else
        T2M_TRY(fill_stage(m, pm, si++, ae + enc_hc[i] + ".conv", D2, H, 3, enc_dil[i], PRO_HWY, eh1[i - 1].g, eh1[i - 1].b,
                           eh2[i - 1].g, eh2[i - 1].b, i, i - 1, 0, s));
    }
    T2M_TRY(fill_stage(m, pm, si++, ad + "conv1", H, D2, 1, 1, PRO_ATT, eh1[9].g, eh1[9].b, eh2[9].g, eh2[9].b, -1, 9, 0, s));
    for (int i = 0; i < 6; ++i) {
      if (i == 0)
        T2M_TRY(fill_stage(m, pm, si++, ad + dec_hc[i] + ".conv", D2, H, 3, dec_dil[i], PRO_LN, d1.g, d1.b, nullptr, nullptr, 10, -1, 0, s));
      else
        T2M_TRY(fill_stage(m, pm, si++, ad + dec_hc[i] + ".conv", D2, H, 3, dec_dil[i], PRO_HWY, dh1[i - 1].g, dh1[i - 1].b,
                           dh2[i - 1].g, dh2[i - 1].b, 10 + i, 10 + i - 1, 0, s));
    }
    T2M_TRY(fill_stage(m, pm, si++, ad + "conv2", H, H, 1, 1, PRO_HWY, dh1[5].g, dh1[5].b, dh2[5].g, dh2[5].b, -1, 15, 0, s));
    T2M_TRY(fill_stage(m, pm, si++, ad + "conv3", H, H, 1, 1, PRO_LN_RELU, d2.g, d2.b, nullptr, nullptr, -1, -1, 0, s));
    T2M_TRY(fill_stage(m, pm, si++, ad + "conv4", H, H, 1, 1, PRO_LN_RELU, d3.g, d3.b, nullptr, nullptr, -1, -1, 0, s));
    T2M_TRY(fill_stage(m, pm, si++, ad + "conv5", freq_bins, H, 1, 1, PRO_LN_RELU, d4.g, d4.b, nullptr, nullptr, -1, -1, 0, s));
    if (si != DEC_STAGES) { set_error("internal: stage table has %d entries", si); return fail(kState); }
    T2M_TRY(m->arena.alloc<DecStage>(DEC_STAGES, &m->stages_dev));
    if (cudaMemcpyAsync(m->stages_dev, m->stages, sizeof(DecStage) * DEC_STAGES, cudaMemcpyHostToDevice, s) != cudaSuccess)
      return fail(kCuda);
    // weight-stationary layout: CTA groups, weight-slice images, private-history ring blocks
    int cta = 0, blk = 0;
    for (int i = 0; i < DEC_STAGES; ++i) {
      WsStage& w = m->ws_stages[i];
      ws_stage_layout(m->stages[i], &w);
      if (i == 0) { w.g1 = m->fin_g; w.b1 = m->fin_b; }     // PRO_X finishes the previous frame: sigmoid(LN5(.))
      w.cta0 = cta; cta += w.parts;
      w.hist_blk0 = blk; blk += w.parts * w.hist_depth;
      float* img;
      T2M_TRY(m->arena.alloc<float>(ws_image_floats(w), &img));
      T2M_TRY(ws_pack_image(m->stages[i].W, w, img, s));
      w.img = img;
    }
    if (cta != WS_GRID) { set_error("internal: weight-stationary layout uses %d CTAs, expected %d", cta, WS_GRID); return fail(kState); }
    m->ws_hist_blocks = blk;
    T2M_TRY(m->arena.alloc<WsStage>(DEC_STAGES, &m->ws_stages_dev));
    if (cudaMemcpyAsync(m->ws_stages_dev, m->ws_stages, sizeof(WsStage) * DEC_STAGES, cudaMemcpyHostToDevice, s) != cudaSuccess)
      return fail(kCuda);
  }
#undef T2M_TRY
  if (cudaStreamSynchronize(s) != cudaSuccess) {
    set_error("text2mel_create: %s", cudaGetErrorString(cudaGetLastError()));
    delete m;
    return kCuda;
  }
  *out = m;
  return kOk;
}

int ssv_text2mel_destroy(ssv_text2mel* m) {
  if (m) {
    cudaDeviceSynchronize();
    delete m;
  }
  return kOk;
}

// Text encoder on channels-last buffers; result (B, N, 2H) left in *out_cl.
static int text_encoder_cl(ssv_text2mel* m, const int64_t* textid, int B, int N, float** out_cl, cudaStream_t s) {
  const int D2 = 2 * m->H;
  SSV_TRY(m->ws.ensure((size_t)B * N * D2));
  float* P = m->ws.buf[0];
  float* Q = m->ws.buf[1];
  const int e_ld = round_up(m->temb, 16);
  SSV_TRY(launch_embed(textid, B, N, m->emb_wt, m->emb_b, m->vocab, m->temb, P, e_ld, m->err_flag, s));
  SSV_TRY(run_conv(m->te_conv1, EPI_LN_RELU, 1, 0, P, e_ld, N, B, Q, D2, s));
  SSV_TRY(run_conv(m->te_conv2, EPI_LN, 1, 0, Q, D2, N, B, P, D2, s));
  float* cur = P;
  float* nxt = Q;
  for (int i = 0; i < 12; ++i) {
    SSV_TRY(run_conv(m->te_hc[i], EPI_HIGHWAY, m->te_dil[i], 0, cur, D2, N, B, nxt, D2, s));
    float* t = cur; cur = nxt; nxt = t;
  }
  *out_cl = cur;
  return kOk;
}

// The same on the tensor cores (conv_tc32.cu: tcgen05 kind::tf32 with split operands, FP32-accurate): activations
// travel between layers as the (hi, lo) pair the next layer's TMA reads; the last layer writes plain fp32.
static int text_encoder_cl_tc(ssv_text2mel* m, const int64_t* textid, int B, int N, float** out_cl, cudaStream_t s) {
  const int D2 = 2 * m->H;
  const size_t need = (size_t)B * N * D2;
  if (need > m->te32_floats) {
    for (float*& p : m->te32_ws) {
      if (p) cudaFree(p);
      p = nullptr;
    }
    m->te32_floats = 0;
    for (float*& p : m->te32_ws)
      if (cudaMalloc((void**)&p, need * sizeof(float) + 256) != cudaSuccess) {
        cudaGetLastError();
        set_error("text_encoder: workspace cudaMalloc of %zu bytes failed", need * sizeof(float));
        return kNoMem;
      }
    m->te32_floats = need;
    m->te32_B = m->te32_N = 0;                     // the launches hold the old buffer addresses
  }
  float *Ph = m->te32_ws[0], *Pl = m->te32_ws[1], *Qh = m->te32_ws[2], *Ql = m->te32_ws[3];
  const int e_ld = m->te32_conv1.cin_p;
  if (m->te32_B != B || m->te32_N != N) {          // (re)build the 14 launches: TMA descriptors of this shape
    m->te32_B = m->te32_N = 0;
    SSV_TRY(tf32_prepare(m->te32_conv1, EPI_LN_RELU, 1, 0, Ph, Pl, e_ld, N, B, Qh, Ql, D2, &m->te32_plan[0]));
    SSV_TRY(tf32_prepare(m->te32_conv2, EPI_LN, 1, 0, Qh, Ql, D2, N, B, Ph, Pl, D2, &m->te32_plan[1]));
    float *ch = Ph, *cl = Pl, *nh = Qh, *nl = Ql;
    for (int i = 0; i < 12; ++i) {
      const bool last = i == 11;
      SSV_TRY(tf32_prepare(m->te32_hc[i], EPI_HIGHWAY, m->te_dil[i], 0, ch, cl, D2, N, B, nh, last ? nullptr : nl, D2, &m->te32_plan[2 + i]));
      float* t = ch; ch = nh; nh = t;
      t = cl; cl = nl; nl = t;
    }
    m->te32_B = B; m->te32_N = N;
  }
  SSV_TRY(launch_embed(textid, B, N, m->emb_wt, m->emb_b, m->vocab, m->temb, Qh, e_ld, m->err_flag, s));
  SSV_TRY(launch_split_tf32(Qh, Ph, Pl, (size_t)B * N * e_ld, s));
  for (int i = 0; i < 14; ++i) SSV_TRY(tf32_run(m->te32_plan[i], s));
  float* ch = Ph;                                  // 12 highway layers: the result is back in the first buffer
  *out_cl = ch;
  return kOk;
}

int ssv_text_encoder_fwd(ssv_text2mel* m, const int64_t* textid, int B, int N, float* K, float* V, int precision,
                         void* stream) {
  SSV_CHECK(m && textid && K && V, "text_encoder: null pointer");
  SSV_CHECK(B > 0 && N > 0, "text_encoder: empty input");
  SSV_CHECK(precision == SSV_PREC_FP32 || precision == SSV_PREC_FP32_FFMA,
            "text_encoder: SSV_PREC_FP32 (tensor cores, 3xTF32 split) or SSV_PREC_FP32_FFMA (CUDA cores); a BF16 TextEnc moves "
            "the alignment trajectory and is not built");
  cudaStream_t s = as_stream(stream);
  float* x;
  if (precision == SSV_PREC_FP32 && m->temb % T32_BK == 0)
    SSV_TRY(text_encoder_cl_tc(m, textid, B, N, &x, s));
  else
    SSV_TRY(text_encoder_cl(m, textid, B, N, &x, s));
  SSV_TRY(launch_transpose_out2(x, 2 * m->H, 0, B, m->H, N, K, s));
  SSV_TRY(launch_transpose_out2(x, 2 * m->H, m->H, B, m->H, N, V, s));
  return kOk;
}

// ------------------------------------------------------------------------------------------------
// Train-mode (teacher-forced) forward of melSyn, models/TTSModel.py:263-273: full-sequence AudioEnc, unmasked
// attention over all characters, AudioDec -- the same fused conv kernels as TextEnc / SSRN with causal taps.
int ssv_text2mel_train_fwd(ssv_text2mel* m, const float* melspec, const int64_t* textid, const float* spkemb, int B,
                           int N, int T, float* Y, float* A, int precision, void* stream) {
  SSV_CHECK(m && melspec && textid && spkemb && Y && A, "text2mel_train_fwd: null pointer");
  SSV_CHECK(B > 0 && N > 0 && T > 0, "text2mel_train_fwd: empty input");
  SSV_CHECK(precision == SSV_PREC_FP32 || precision == SSV_PREC_FP32_FFMA,
            "text2mel_train_fwd: SSV_PREC_FP32 (highway stacks on the tensor cores, 3xTF32 split) or SSV_PREC_FP32_FFMA (CUDA cores)");
  cudaStream_t s = as_stream(stream);
  const int H = m->H, D2 = 2 * H, F = m->F;
  const bool tc = precision == SSV_PREC_FP32 && m->temb % T32_BK == 0;
  float* kx;                                                   // (B, N, 512): K | V, channels-last
  if (tc) SSV_TRY(text_encoder_cl_tc(m, textid, B, N, &kx, s));
  else SSV_TRY(text_encoder_cl(m, textid, B, N, &kx, s));
  if (tc) {       // the 16 causal highway layers: (hi, lo) ping-pong buffers and encoded launches for this shape
    const size_t need = (size_t)B * T * H;
    if (need > m->tr32_floats) {
      for (float*& p : m->tr32_ws) {
        if (p) cudaFree(p);
        p = nullptr;
      }
      m->tr32_floats = 0;
      for (float*& p : m->tr32_ws)
        if (cudaMalloc((void**)&p, need * sizeof(float) + 256) != cudaSuccess) {
          cudaGetLastError();
          set_error("text2mel_train_fwd: workspace cudaMalloc of %zu bytes failed", need * sizeof(float));
          return kNoMem;
        }
      m->tr32_floats = need;
      m->tr32_B = m->tr32_T = 0;
    }
    if (m->tr32_B != B || m->tr32_T != T) {
      m->tr32_B = m->tr32_T = 0;
      float** w = m->tr32_ws;                                  // set 0 = (w[0], w[1]), set 1 = (w[2], w[3])
      for (int i = 0; i < 10; ++i) {
        const int a_ = (i & 1) * 2, b_ = ((i + 1) & 1) * 2;
        SSV_TRY(tf32_prepare(m->ae32_hc[i], EPI_HIGHWAY, m->ae_dil[i], 1, w[a_], w[a_ + 1], H, T, B, w[b_], i == 9 ? nullptr : w[b_ + 1], H,
                             &m->tr32_plan[i]));
      }
      for (int i = 0; i < 6; ++i) {
        const int a_ = (i & 1) * 2, b_ = ((i + 1) & 1) * 2;
        SSV_TRY(tf32_prepare(m->ad32_hc[i], EPI_HIGHWAY, m->ad_dil[i], 1, w[a_], w[a_ + 1], H, T, B, w[b_], i == 5 ? nullptr : w[b_ + 1], H,
                             &m->tr32_plan[10 + i]));
      }
      m->tr32_B = B; m->tr32_T = T;
    }
  }
  SSV_TRY(m->tws.ensure((size_t)B * T * D2));
  if (m->ts_cap < B) {
    SSV_TRY(m->arena.alloc<float>((size_t)B * H, &m->ts1));
    SSV_TRY(m->arena.alloc<float>((size_t)B * H, &m->ts2));
    m->ts_cap = B;
  }
  float* P = m->tws.buf[0];
  float* Q = m->tws.buf[1];
  SSV_TRY(launch_linear_small(spkemb, m->E, m->fc1_w, m->fc1_b, B, m->E, H, m->ts1, H, s));
  SSV_TRY(launch_linear_small(spkemb, m->E, m->fc2_w, m->fc2_b, B, m->E, H, m->ts2, H, s));
  const int f_ld = round_up(F, 16);
  SSV_TRY(launch_transpose_in(melspec, (long)F * T, T, 1, B, F, T, P, f_ld, s));
  // AudioEnc (:172-184): conv1 + fc1(e) -> LN -> ReLU -> conv2 -> LN -> ReLU -> conv3 + fc2(e) -> LN -> 10 causal highway convs
  SSV_TRY(run_conv(m->ae_conv1, EPI_LN_RELU, 1, 0, P, f_ld, T, B, Q, H, s, m->ts1, H));
  SSV_TRY(run_conv(m->ae_conv2, EPI_LN_RELU, 1, 0, Q, H, T, B, P, H, s));
  SSV_TRY(run_conv(m->ae_conv3, EPI_LN, 1, 0, P, H, T, B, Q, H, s, m->ts2, H));
  float* cur = Q;
  float* nxt = P;
  if (tc) {       // split, ten layers on the tensor cores, the last one writes plain fp32 (even count: back in set 0)
    SSV_TRY(launch_split_tf32(cur, m->tr32_ws[0], m->tr32_ws[1], (size_t)B * T * H, s));
    for (int i = 0; i < 10; ++i) SSV_TRY(tf32_run(m->tr32_plan[i], s));
    cur = m->tr32_ws[0];
    nxt = P;
  } else {
    for (int i = 0; i < 10; ++i) {
      SSV_TRY(run_conv(m->ae_hc[i], EPI_HIGHWAY, m->ae_dil[i], 1, cur, H, T, B, nxt, H, s));
      float* t_ = cur; cur = nxt; nxt = t_;
    }
  }
  // attention (:266-270) -> A (B, N, T), [R ; Q] (B, T, 512)
  SSV_TRY(launch_train_attention(kx, cur, B, N, T, A, nxt, s));
  // AudioDec (:217-232)
  if (tc) {
    SSV_TRY(run_conv(m->ad_conv1, EPI_LN, 1, 0, nxt, D2, T, B, Q, H, s));
    SSV_TRY(launch_split_tf32(Q, m->tr32_ws[0], m->tr32_ws[1], (size_t)B * T * H, s));
    for (int i = 0; i < 6; ++i) SSV_TRY(tf32_run(m->tr32_plan[10 + i], s));
    cur = m->tr32_ws[0];
    nxt = P;
  } else {
    SSV_TRY(run_conv(m->ad_conv1, EPI_LN, 1, 0, nxt, D2, T, B, cur, H, s));
    for (int i = 0; i < 6; ++i) {
      SSV_TRY(run_conv(m->ad_hc[i], EPI_HIGHWAY, m->ad_dil[i], 1, cur, H, T, B, nxt, H, s));
      float* t_ = cur; cur = nxt; nxt = t_;
    }
  }
  SSV_TRY(run_conv(m->ad_conv2, EPI_LN_RELU, 1, 0, cur, H, T, B, nxt, H, s));
  SSV_TRY(run_conv(m->ad_conv3, EPI_LN_RELU, 1, 0, nxt, H, T, B, cur, H, s));
  SSV_TRY(run_conv(m->ad_conv4, EPI_LN_RELU, 1, 0, cur, H, T, B, nxt, H, s));
  const int y_ld = round_up(F, 64);
  SSV_TRY(run_conv(m->ad_conv5, EPI_LN_SIGMOID, 1, 0, nxt, H, T, B, cur, y_ld, s));
  SSV_TRY(launch_transpose_out(cur, y_ld, B, F, T, Y, s));
  return kOk;
}

// ------------------------------------------------------------------------------------------------
int ssv_decoder_create(ssv_text2mel* m, int max_batch, int max_text, int max_frames, ssv_decoder** out) {
  SSV_CHECK(m && out, "decoder_create: null pointer");
  SSV_CHECK(max_batch >= 1 && max_text >= 1 && max_frames >= 1, "decoder_create: bad capacity");
  ssv_decoder* d = new ssv_decoder();
  d->m = m;
  d->maxB = max_batch; d->maxN = max_text; d->maxT = max_frames;
  const int H = m->H;
  int st = kOk;
  if (max_batch > WS_MAX_BATCH) {
    delete d;
    set_error("decoder_create: max_batch %d exceeds %d", max_batch, WS_MAX_BATCH);
    return kInval;
  }
  if (!decode_ws_supported(device_sm_count())) {
    delete d;
    set_error("decoder_create: the decode kernel needs a cooperative launch of %d co-resident CTAs (B200: 148 SMs)", WS_GRID);
    return kState;
  }
  {
    const size_t bp = (size_t)round_up(max_batch, 8) + 8;      // micro-batch count is padded to the visit slots
    const size_t raw_words = (size_t)DEC_STAGES * max_batch * WS_WORDS;
    if (st == kOk) st = d->arena.alloc<unsigned long long>(raw_words, &d->ws_raw);
    // ring entries are [256][XS] floats per micro-batch, XS = 2 (R = 1: every activation as the pair (x, x); R = 2) or 4
    if (st == kOk) st = d->arena.alloc<float>((size_t)m->ws_hist_blocks * bp * H * 2, &d->ws_hist);
    if (st == kOk) st = d->arena.alloc<float>((size_t)H * 4, &d->ws_zero);
    if (st == kOk && cudaMemset(d->ws_zero, 0, sizeof(float) * H * 4) != cudaSuccess) st = kCuda;
    if (st == kOk && cudaMemset(d->ws_raw, 0, raw_words * sizeof(unsigned long long)) != cudaSuccess) st = kCuda;
  }
  if (st == kOk) st = d->arena.alloc<float>((size_t)max_batch * max_text * H, &d->Kt);
  if (st == kOk) st = d->arena.alloc<float>((size_t)max_batch * max_text * H, &d->Vt);
  if (st == kOk) st = d->arena.alloc<float>((size_t)max_batch * H, &d->s1);
  if (st == kOk) st = d->arena.alloc<float>((size_t)max_batch * H, &d->s2);
  if (st == kOk) st = d->arena.alloc<int>(max_batch, &d->pma_state);
  if (st == kOk) st = d->arena.alloc<int>(1, &d->abort_flag);
  if (st == kOk && getenv("SSV_DECODE_PROF")) st = d->arena.alloc<long long>((size_t)DEC_MAX_GRID * 16, &d->prof);
  if (st == kOk && !d->prof && getenv("SSV_DECODE_TOTALS")) st = d->arena.alloc<long long>((size_t)DEC_MAX_GRID * 4, &d->totals);
  if (st != kOk) { delete d; return st; }
  cudaMemset(d->abort_flag, 0, sizeof(int));
  *out = d;
  return kOk;
}

int ssv_decoder_destroy(ssv_decoder* d) {
  if (d) {
    cudaDeviceSynchronize();
    delete d;
  }
  return kOk;
}

int ssv_decoder_begin(ssv_decoder* d, const float* K, const float* V, const float* spkemb, int B, int N, float* Y,
                      float* A, int64_t* pma_traj, int t_cap, void* stream) {
  SSV_CHECK(d && K && V && spkemb && Y && A && pma_traj, "decoder_begin: null pointer");
  SSV_CHECK(B >= 1 && B <= d->maxB, "decoder_begin: batch %d outside [1, %d]", B, d->maxB);
  SSV_CHECK(N >= 1 && N <= d->maxN, "decoder_begin: text length %d outside [1, %d]", N, d->maxN);
  SSV_CHECK(t_cap >= 1 && t_cap <= d->maxT, "decoder_begin: frame capacity %d outside [1, %d]", t_cap, d->maxT);
  cudaStream_t s = as_stream(stream);
  ssv_text2mel* m = d->m;
  const int H = m->H;
  SSV_TRY(launch_transpose_in(K, (long)H * N, N, 1, B, H, N, d->Kt, H, s));
  SSV_TRY(launch_transpose_in(V, (long)H * N, N, 1, B, H, N, d->Vt, H, s));
  SSV_TRY(launch_linear_small(spkemb, m->E, m->fc1_w, m->fc1_b, B, m->E, H, d->s1, H, s));
  SSV_TRY(launch_linear_small(spkemb, m->E, m->fc2_w, m->fc2_b, B, m->E, H, d->s2, H, s));
  SSV_CUDA(cudaMemsetAsync(A, 0, sizeof(float) * (size_t)B * N * t_cap, s));
  SSV_CUDA(cudaMemsetAsync(d->pma_state, 0, sizeof(int) * B, s));
  // tags of this batch: seq_base + t + 1, strictly above every tag of earlier batches
  d->seq_base += d->t_cap + 2;
  if (d->seq_base > (1 << 30)) {
    SSV_CUDA(cudaMemsetAsync(d->ws_raw, 0, (size_t)DEC_STAGES * d->maxB * WS_WORDS * sizeof(unsigned long long), s));
    d->seq_base = 0;
  }
  ws_plan(B, d->force_r, d->force_w, d->force_f, &d->R, &d->W, &d->F, &d->G);
  d->B = B; d->N = N; d->t_cap = t_cap; d->t = 0;
  d->Y = Y; d->A = A; d->traj = reinterpret_cast<long long*>(pma_traj);
  d->begun = true;
  return kOk;
}

static int decoder_launch(ssv_decoder* d, int n_steps, const float* x_ext, long x_sb, long x_sf,
                          const int64_t* pma_in, cudaStream_t s) {
  SSV_CHECK(d && d->begun, "decoder: begin() has not been called");
  if (d->t + n_steps > d->t_cap) {
    set_error("decoder: %d + %d frames exceed capacity %d", d->t, n_steps, d->t_cap);
    return kState;
  }
  ssv_text2mel* m = d->m;
  DecParams p{};
  p.stages = m->stages_dev;
  p.fin_g = m->fin_g; p.fin_b = m->fin_b;
  p.Kt = d->Kt; p.Vt = d->Vt; p.s1 = d->s1; p.s2 = d->s2;
  p.Y = d->Y; p.A = d->A; p.pma_traj = d->traj;
  p.pma_in = reinterpret_cast<const long long*>(pma_in);
  p.pma_state = d->pma_state;
  p.x_ext = x_ext; p.x_sb = x_sb; p.x_sf = x_sf;
  p.B = d->B; p.N = d->N; p.t_cap = d->t_cap; p.F = m->F; p.H = m->H;
  p.t_start = d->t; p.n_steps = n_steps;
  p.abort_flag = d->abort_flag;
  p.prof = d->prof;
  p.totals = d->totals;
  p.ws_stages = m->ws_stages_dev;
  p.ws_raw = d->ws_raw; p.ws_hist = d->ws_hist; p.ws_zero = d->ws_zero;
  p.seq_base = d->seq_base; p.R = d->R; p.G = d->G; p.W = d->W; p.FEW = d->F;
  if (d->prof) SSV_CUDA(cudaMemsetAsync(d->prof, 0, sizeof(long long) * DEC_MAX_GRID * 16, s));
  SSV_TRY(launch_decode_ws(p, s));
  if (d->totals) {   // development aid: cycles per stage visit of each role, from the un-instrumented kernel
    SSV_CUDA(cudaStreamSynchronize(s));
    std::vector<long long> h((size_t)DEC_MAX_GRID * 4);
    SSV_CUDA(cudaMemcpy(h.data(), d->totals, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    const double visits = (double)n_steps * d->G;
    fprintf(stderr, "[decode totals] B=%d R=%d W=%d F=%d G=%d steps=%d: cycles per visit, mean over the stage's CTAs\n  stage ctas |  front end  (waiting) |  mat-vec  (waiting)\n", d->B, d->R, d->W, d->F, d->G, n_steps);
    for (int sidx = 0; sidx < DEC_STAGES; ++sidx) {
      const WsStage& w = m->ws_stages[sidx];
      double a[4] = {0, 0, 0, 0};
      for (int c = w.cta0; c < w.cta0 + w.parts; ++c)
        for (int i = 0; i < 4; ++i) a[i] += (double)h[(size_t)c * 4 + i] / visits / w.parts;
      fprintf(stderr, "  %5d %4d | %10.0f %10.0f | %8.0f %10.0f\n", sidx, w.parts, a[0], a[1], a[2], a[3]);
    }
  }
  if (d->prof) {   // development aid: per-phase SM cycles, mean / max over CTAs, per stage visit
    SSV_CUDA(cudaStreamSynchronize(s));
    const bool ws = true;
    const int stride = 16, cnt_slot = 15, n_ph = 15;
    std::vector<long long> h((size_t)DEC_MAX_GRID * 16);
    SSV_CUDA(cudaMemcpy(h.data(), d->prof, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    const char* nm_ws[15] = {"FE loop-top", "FE wait X buffer empty", "FE tap prefetch issue", "FE wait for producer sentinel", "FE row poll / load (tags verified)", "FE ring write + tap prefetch", "FE prologue math + X stores", "-",
                             "MV wait taps", "MV old taps (2/3)", "MV wait current tap", "MV current tap + k-slice store", "MV reducer: reduce + publish",
                             "sentinel seen -> my sentinel out", "curfull arrive -> MV awake (highway)"};
    const char* const* nm = nm_ws;
    fprintf(stderr, "[decode prof] B=%d R=%d W=%d G=%d steps=%d (cycles per stage visit: mean / max over CTAs)\n", d->B, d->R, d->W, d->G, n_steps);
    for (int i = 0; i < n_ph; ++i) {
      if (nm[i][0] == '-') continue;
      double sum = 0, mx = 0; int cnt = 0;
      for (int c = 0; c < DEC_MAX_GRID; ++c) {
        if (h[(size_t)c * stride + cnt_slot] == 0) continue;
        const int cs = (ws && i >= 8) ? 7 : cnt_slot;       // mat-vec slots carry their own visit count
        if (h[(size_t)c * stride + cs] == 0) continue;
        const double v = (double)h[(size_t)c * stride + i] / (double)h[(size_t)c * stride + cs];
        sum += v; mx = v > mx ? v : mx; ++cnt;
      }
      fprintf(stderr, "  %-32s %9.0f / %9.0f\n", nm[i], cnt ? sum / cnt : 0.0, mx);
    }
    // per stage (mean over the stage's CTAs): which stage sets the pace of the pipeline
    fprintf(stderr, "  stage  ctas | FE: top  taps  wait  poll  ring  math | MV: wtaps  old  wcur  cur  publish\n");
    for (int sidx = 0; sidx < DEC_STAGES; ++sidx) {
      const WsStage& w = m->ws_stages[sidx];
      double acc[15] = {0};
      for (int c = w.cta0; c < w.cta0 + w.parts; ++c)
        for (int i = 0; i < 15; ++i) {
          const int cs = i >= 8 ? 7 : cnt_slot;
          const double den = (double)h[(size_t)c * stride + cs];
          if (den > 0) acc[i] += (double)h[(size_t)c * stride + i] / den / w.parts;
        }
      fprintf(stderr, "  %5d %5d | %7.0f %5.0f %5.0f %5.0f %5.0f %5.0f | %8.0f %5.0f %5.0f %5.0f %5.0f\n", sidx, w.parts, acc[0], acc[2],
              acc[3], acc[4], acc[5], acc[6], acc[8], acc[9], acc[10], acc[11], acc[12]);
    }
  }
  d->t += n_steps;
  return kOk;
}

int ssv_decoder_step(ssv_decoder* d, const float* x, long x_stride_b, long x_stride_f, const int64_t* pma_in,
                     void* stream) {
  return decoder_launch(d, 1, x, x_stride_b, x_stride_f, pma_in, as_stream(stream));
}

int ssv_decoder_run(ssv_decoder* d, int n_steps, void* stream) {
  SSV_CHECK(n_steps >= 1, "decoder_run: n_steps must be >= 1");
  return decoder_launch(d, n_steps, nullptr, 0, 0, nullptr, as_stream(stream));
}

int ssv_decoder_frames(const ssv_decoder* d) { return d ? d->t : 0; }

int ssv_decoder_set_plan(ssv_decoder* d, int rows_per_microbatch, int warps_per_row, int front_end_warps) {
  SSV_CHECK(d, "decoder_set_plan: null decoder");
  SSV_CHECK(rows_per_microbatch == 0 || rows_per_microbatch == 1 || rows_per_microbatch == 2 || rows_per_microbatch == 4,
            "decoder_set_plan: rows per micro-batch must be 0 (automatic), 1, 2 or 4");
  SSV_CHECK(warps_per_row == 0 || warps_per_row == 1 || warps_per_row == 2 || warps_per_row == 4,
            "decoder_set_plan: warps per row must be 0 (automatic), 1, 2 or 4");
  SSV_CHECK(front_end_warps == 0 || front_end_warps == 4 || front_end_warps == 8,
            "decoder_set_plan: front-end warps must be 0 (automatic), 4 or 8");
  d->force_r = rows_per_microbatch;
  d->force_w = warps_per_row;
  d->force_f = front_end_warps;
  return kOk;
}

int ssv_decoder_set_lin_output(ssv_decoder* d, int bf16) {
  SSV_CHECK(d, "decoder_set_lin_output: null decoder");
  if (d->lin_out_bf16 == (bf16 ? 1 : 0)) return kOk;
  SSV_CHECK(!d->inflight[0] && !d->inflight[1], "decoder_set_lin_output: a batch is in flight");
  d->lin_out_bf16 = bf16 ? 1 : 0;
  return kOk;
}

int ssv_text2mel_check(ssv_text2mel* m, void* stream) {
  SSV_CHECK(m, "text2mel_check: null model");
  SSV_CUDA(cudaStreamSynchronize(as_stream(stream)));
  SSV_TRY(tf32_check_error());
  int eflag = 0;
  SSV_CUDA(cudaMemcpy(&eflag, m->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (eflag != 0) {
    set_error("text id outside [0, vocab_len)");
    cudaMemset(m->err_flag, 0, sizeof(int));
    return kInval;
  }
  return kOk;
}

int ssv_decoder_check(ssv_decoder* d, void* stream) {
  SSV_CHECK(d, "decoder_check: null decoder");
  SSV_CUDA(cudaStreamSynchronize(as_stream(stream)));
  int flag = 0;
  SSV_CUDA(cudaMemcpy(&flag, d->abort_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag != 0) {
    set_error("decode kernel aborted (grid barrier timeout, code %d)", flag);
    cudaMemset(d->abort_flag, 0, sizeof(int));
    return kState;
  }
  SSV_TRY(tf32_check_error());            // TextEnc runs on the split tensor-core arm: its pipeline-timeout flag
  int eflag = 0;
  SSV_CUDA(cudaMemcpy(&eflag, d->m->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (eflag != 0) {
    set_error("text id outside [0, vocab_len)");
    cudaMemset(d->m->err_flag, 0, sizeof(int));
    return kInval;
  }
  return kOk;
}

// ------------------------------------------------------------------------------------------------
int ssv_ssrn_create(const char* const* names, const float* const* dev_ptrs, const int64_t* numels, int n_params,
                    int freq_bins, int output_bins, int ssrn_dim, ssv_ssrn** out) {
  SSV_CHECK(out, "ssrn_create: null out");
  SSV_CHECK(ssrn_dim == 256, "ssrn_create: ssrn_dim %d unsupported (256)", ssrn_dim);
  SSV_CHECK(freq_bins % 16 == 0 && freq_bins <= 128, "ssrn_create: freq_bins must be a multiple of 16, <= 128");
  SSV_CHECK(output_bins > 512 && output_bins <= 576, "ssrn_create: output_bins %d unsupported (513..576)", output_bins);
  ParamMap pm;
  SSV_TRY(build_param_map(names, dev_ptrs, numels, n_params, &pm));
  ssv_ssrn* m = new ssv_ssrn();
  m->F = freq_bins; m->O = output_bins; m->D = ssrn_dim;
  cudaStream_t s = nullptr;
  const int D = ssrn_dim, O = output_bins;
  int st = kOk;
  auto deconv = [&](const std::string& name, ConvPack* c) -> int {
    const float *w, *b;
    SSV_TRY(pm.get(name + ".weight", (int64_t)D * D * 2, &w));
    SSV_TRY(pm.get(name + ".bias", D, &b));
    c->cin = D; c->cin_p = D; c->k = 1; c->n = 2 * D; c->n_pad = 2 * D;
    SSV_TRY(m->arena.alloc<float>((size_t)D * 2 * D, &c->W));
    SSV_TRY(m->arena.alloc<float>(2 * D, &c->bias));
    return launch_pack_deconv_w(w, b, D, D, c->W, c->bias, s);
  };
#define S_TRY(expr) do { if (st == kOk) st = (expr); } while (0)
  S_TRY(pack_conv_ln(m->arena, pm, "conv1", "ln1", D, freq_bins, &m->conv1, s));
  S_TRY(pack_highway(m->arena, pm, "hc1", D, 3, &m->hc1, s));
  S_TRY(pack_highway(m->arena, pm, "hc2", D, 3, &m->hc2, s));
  S_TRY(deconv("ups1.deconv", &m->dc1));
  S_TRY(pack_highway(m->arena, pm, "ups1.hc1", D, 3, &m->u1h1, s));
  S_TRY(pack_highway(m->arena, pm, "ups1.hc2", D, 3, &m->u1h2, s));
  S_TRY(deconv("ups2.deconv", &m->dc2));
  S_TRY(pack_highway(m->arena, pm, "ups2.hc1", D, 3, &m->u2h1, s));
  S_TRY(pack_highway(m->arena, pm, "ups2.hc2", D, 3, &m->u2h2, s));
  S_TRY(pack_conv_ln(m->arena, pm, "conv2", "ln2", 2 * D, D, &m->conv2, s));
  S_TRY(pack_highway(m->arena, pm, "hc3", 2 * D, 3, &m->hc3, s));
  S_TRY(pack_highway(m->arena, pm, "hc4", 2 * D, 3, &m->hc4, s));
  S_TRY(pack_conv_ln(m->arena, pm, "conv3", "ln3", O, 2 * D, &m->conv3, s));
  S_TRY(pack_conv_ln(m->arena, pm, "conv4", "ln4", O, O, &m->conv4, s));
  S_TRY(pack_conv_ln(m->arena, pm, "conv5", "ln5", O, O, &m->conv5, s));
  S_TRY(pack_conv_ln(m->arena, pm, "conv6", "ln6", O, O, &m->conv6, s));
  {  // bf16 copies for the tcgen05 path
    auto hw = [&](const char* name, const ConvPack& c, int d, TcLayer* L) -> int {
      tc_shape_highway(L, d);
      return tc_pack_layer(m->arena, pm, std::string(name) + ".conv.weight", c, 2 * d, d, 3, L, s);
    };
    auto pl = [&](const char* name, const ConvPack& c, int n, int cin, TcLayer* L) -> int {
      tc_shape_plain(L, n);
      return tc_pack_layer(m->arena, pm, std::string(name) + ".weight", c, n, cin, 1, L, s);
    };
    auto dc = [&](const char* name, const ConvPack& c, TcLayer* L) -> int {
      const float* w;
      SSV_TRY(pm.get(std::string(name) + ".weight", (int64_t)D * D * 2, &w));
      tc_shape_plain(L, 2 * D);
      L->rows = 2 * D; L->cin = D; L->cin_p = D; L->k = 1; L->bias = c.bias;
      SSV_TRY(m->arena.alloc<__nv_bfloat16>((size_t)2 * D * D, &L->W));
      return tc_pack_deconv(w, D, D, L->W, s);
    };
    S_TRY(pl("conv1", m->conv1, D, freq_bins, &m->t_conv1));
    S_TRY(hw("hc1", m->hc1, D, &m->t_hc1));
    S_TRY(hw("hc2", m->hc2, D, &m->t_hc2));
    S_TRY(dc("ups1.deconv", m->dc1, &m->t_dc1));
    S_TRY(hw("ups1.hc1", m->u1h1, D, &m->t_u1h1));
    S_TRY(hw("ups1.hc2", m->u1h2, D, &m->t_u1h2));
    S_TRY(dc("ups2.deconv", m->dc2, &m->t_dc2));
    S_TRY(hw("ups2.hc1", m->u2h1, D, &m->t_u2h1));
    S_TRY(hw("ups2.hc2", m->u2h2, D, &m->t_u2h2));
    S_TRY(pl("conv2", m->conv2, 2 * D, D, &m->t_conv2));
    S_TRY(hw("hc3", m->hc3, 2 * D, &m->t_hc3));
    S_TRY(hw("hc4", m->hc4, 2 * D, &m->t_hc4));
    S_TRY(pl("conv3", m->conv3, O, 2 * D, &m->t_conv3));
    S_TRY(pl("conv4", m->conv4, O, O, &m->t_conv4));
    S_TRY(pl("conv5", m->conv5, O, O, &m->t_conv5));
    S_TRY(pl("conv6", m->conv6, O, O, &m->t_conv6));
  }
  {   // FP32-accurate tensor-core arm (conv_tc32.cu)
    auto fh = [&](const char* name, const ConvPack& c, int d, Tf32Layer* L) -> int {
      tf32_shape_highway(L, d);
      return tf32_pack_layer(m->arena, pm, std::string(name) + ".conv.weight", c, 2 * d, d, 3, L, s);
    };
    auto fp = [&](const char* name, const ConvPack& c, int n, int cin, Tf32Layer* L) -> int {
      tf32_shape_plain(L, n);
      return tf32_pack_layer(m->arena, pm, std::string(name) + ".weight", c, n, cin, 1, L, s);
    };
    auto fd = [&](const char* name, const ConvPack& c, Tf32Layer* L) -> int {
      const float* w;
      SSV_TRY(pm.get(std::string(name) + ".weight", (int64_t)D * D * 2, &w));
      tf32_shape_plain(L, 2 * D);
      L->rows = 2 * D; L->cin = D; L->cin_p = D; L->k = 1;
      L->bias = c.bias;
      SSV_TRY(m->arena.alloc<float>((size_t)2 * D * D, &L->Wh));
      SSV_TRY(m->arena.alloc<float>((size_t)2 * D * D, &L->Wl));
      return tf32_pack_deconv_weights(w, D, D, L->Wh, L->Wl, s);
    };
    S_TRY(fp("conv1", m->conv1, D, freq_bins, &m->f_conv1));
    S_TRY(fh("hc1", m->hc1, D, &m->f_hc1));
    S_TRY(fh("hc2", m->hc2, D, &m->f_hc2));
    S_TRY(fd("ups1.deconv", m->dc1, &m->f_dc1));
    S_TRY(fh("ups1.hc1", m->u1h1, D, &m->f_u1h1));
    S_TRY(fh("ups1.hc2", m->u1h2, D, &m->f_u1h2));
    S_TRY(fd("ups2.deconv", m->dc2, &m->f_dc2));
    S_TRY(fh("ups2.hc1", m->u2h1, D, &m->f_u2h1));
    S_TRY(fh("ups2.hc2", m->u2h2, D, &m->f_u2h2));
    S_TRY(fp("conv2", m->conv2, 2 * D, D, &m->f_conv2));
    S_TRY(fh("hc3", m->hc3, 2 * D, &m->f_hc3));
    S_TRY(fh("hc4", m->hc4, 2 * D, &m->f_hc4));
    S_TRY(fp("conv3", m->conv3, O, 2 * D, &m->f_conv3));
    S_TRY(fp("conv4", m->conv4, O, O, &m->f_conv4));
    S_TRY(fp("conv5", m->conv5, O, O, &m->f_conv5));
    S_TRY(fp("conv6", m->conv6, O, O, &m->f_conv6));
  }
#undef S_TRY
  if (st == kOk && cudaStreamSynchronize(s) != cudaSuccess) {
    set_error("ssrn_create: %s", cudaGetErrorString(cudaGetLastError()));
    st = kCuda;
  }
  if (st != kOk) { delete m; return st; }
  *out = m;
  return kOk;
}

int ssv_ssrn_destroy(ssv_ssrn* m) {
  if (m) {
    cudaDeviceSynchronize();
    delete m;
  }
  return kOk;
}

static int ssrn_fwd_f32(ssv_ssrn* m, const float* mel, long sb, long sf, long st_, int B, int T, float* out,
                        cudaStream_t s) {
  const int D = m->D, O = m->O;
  const int o_ld = round_up(O, 64);
  const int f_ld = round_up(m->F, 64);
  SSV_TRY(m->ws.ensure((size_t)B * 4 * T * o_ld));
  float* P = m->ws.buf[0];
  float* Q = m->ws.buf[1];
  SSV_TRY(launch_transpose_in(mel, sb, sf, st_, B, m->F, T, P, f_ld, s));
  SSV_TRY(run_conv(m->conv1, EPI_LN, 1, 0, P, f_ld, T, B, Q, D, s));
  SSV_TRY(run_conv(m->hc1, EPI_HIGHWAY, 1, 0, Q, D, T, B, P, D, s));
  SSV_TRY(run_conv(m->hc2, EPI_HIGHWAY, 3, 0, P, D, T, B, Q, D, s));
  // ConvTranspose1d(k=2, s=2) == 1x1 GEMM to 2*D columns; (B, T, 2D) is (B, 2T, D) in memory.
  SSV_TRY(run_conv(m->dc1, EPI_NONE, 1, 0, Q, D, T, B, P, 2 * D, s));
  SSV_TRY(run_conv(m->u1h1, EPI_HIGHWAY, 1, 0, P, D, 2 * T, B, Q, D, s));
  SSV_TRY(run_conv(m->u1h2, EPI_HIGHWAY, 3, 0, Q, D, 2 * T, B, P, D, s));
  SSV_TRY(run_conv(m->dc2, EPI_NONE, 1, 0, P, D, 2 * T, B, Q, 2 * D, s));
  SSV_TRY(run_conv(m->u2h1, EPI_HIGHWAY, 1, 0, Q, D, 4 * T, B, P, D, s));
  SSV_TRY(run_conv(m->u2h2, EPI_HIGHWAY, 3, 0, P, D, 4 * T, B, Q, D, s));
  SSV_TRY(run_conv(m->conv2, EPI_LN, 1, 0, Q, D, 4 * T, B, P, 2 * D, s));
  SSV_TRY(run_conv(m->hc3, EPI_HIGHWAY, 1, 0, P, 2 * D, 4 * T, B, Q, 2 * D, s));
  SSV_TRY(run_conv(m->hc4, EPI_HIGHWAY, 1, 0, Q, 2 * D, 4 * T, B, P, 2 * D, s));
  SSV_TRY(run_conv(m->conv3, EPI_LN, 1, 0, P, 2 * D, 4 * T, B, Q, o_ld, s));
  SSV_TRY(run_conv(m->conv4, EPI_LN_RELU, 1, 0, Q, o_ld, 4 * T, B, P, o_ld, s));
  SSV_TRY(run_conv(m->conv5, EPI_LN_RELU, 1, 0, P, o_ld, 4 * T, B, Q, o_ld, s));
  SSV_TRY(run_conv(m->conv6, EPI_LN_SIGMOID, 1, 0, Q, o_ld, 4 * T, B, P, o_ld, s));
  SSV_TRY(launch_transpose_out(P, o_ld, B, O, 4 * T, out, s));
  return kOk;
}

// FP32-accurate SSRN on the tensor cores: all 16 layers on conv_tf32x3_kernel with split operands, activations travelling
// as (hi, lo) pairs.  The four 513-bin heads run on three-CTA clusters (768 accumulator columns, 255 of them padding with
// zero weights and zero LayerNorm parameters, so they add nothing to the row statistics).  Same 1e-4 bar as the CUDA-core arm.
static int ssrn_fwd_tf32(ssv_ssrn* m, const float* mel, long sb, long sf, long st_, int B, int T, float* out, cudaStream_t s) {
  const int D = m->D, O = m->O;
  const int o_ld = round_up(O, 64);
  const int f_ld = m->f_conv1.cin_p;
  const size_t need = (size_t)B * 4 * T * (o_ld > 2 * D ? o_ld : 2 * D);
  SSV_TRY(m->ws.ensure((size_t)B * 4 * T * o_ld));
  if (need > m->fws_floats) {
    for (float*& p : m->fws) {
      if (p) cudaFree(p);
      p = nullptr;
    }
    m->fws_floats = 0;
    for (float*& p : m->fws)
      if (cudaMalloc((void**)&p, need * sizeof(float) + 256) != cudaSuccess) {
        cudaGetLastError();
        set_error("ssrn_fwd: workspace cudaMalloc of %zu bytes failed", need * sizeof(float));
        return kNoMem;
      }
    m->fws_floats = need;
    m->fplan_B = m->fplan_T = 0;
  }
  float *Ah = m->fws[0], *Al = m->fws[1], *Bh = m->fws[2], *Bl = m->fws[3];
  if (m->fplan_B != B || m->fplan_T != T || m->fplan_out != m->ws.buf[0]) {
    m->fplan_B = m->fplan_T = 0;
    Tf32Launch* L = m->fplan;
    SSV_TRY(tf32_prepare(m->f_conv1, EPI_LN, 1, 0, Ah, Al, f_ld, T, B, Bh, Bl, D, &L[0]));
    SSV_TRY(tf32_prepare(m->f_hc1, EPI_HIGHWAY, 1, 0, Bh, Bl, D, T, B, Ah, Al, D, &L[1]));
    SSV_TRY(tf32_prepare(m->f_hc2, EPI_HIGHWAY, 3, 0, Ah, Al, D, T, B, Bh, Bl, D, &L[2]));
    SSV_TRY(tf32_prepare(m->f_dc1, EPI_NONE, 1, 0, Bh, Bl, D, T, B, Ah, Al, 2 * D, &L[3]));              // (B,T,2D) == (B,2T,D)
    SSV_TRY(tf32_prepare(m->f_u1h1, EPI_HIGHWAY, 1, 0, Ah, Al, D, 2 * T, B, Bh, Bl, D, &L[4]));
    SSV_TRY(tf32_prepare(m->f_u1h2, EPI_HIGHWAY, 3, 0, Bh, Bl, D, 2 * T, B, Ah, Al, D, &L[5]));
    SSV_TRY(tf32_prepare(m->f_dc2, EPI_NONE, 1, 0, Ah, Al, D, 2 * T, B, Bh, Bl, 2 * D, &L[6]));
    SSV_TRY(tf32_prepare(m->f_u2h1, EPI_HIGHWAY, 1, 0, Bh, Bl, D, 4 * T, B, Ah, Al, D, &L[7]));
    SSV_TRY(tf32_prepare(m->f_u2h2, EPI_HIGHWAY, 3, 0, Ah, Al, D, 4 * T, B, Bh, Bl, D, &L[8]));
    SSV_TRY(tf32_prepare(m->f_conv2, EPI_LN, 1, 0, Bh, Bl, D, 4 * T, B, Ah, Al, 2 * D, &L[9]));
    SSV_TRY(tf32_prepare(m->f_hc3, EPI_HIGHWAY, 1, 0, Ah, Al, 2 * D, 4 * T, B, Bh, Bl, 2 * D, &L[10]));
    SSV_TRY(tf32_prepare(m->f_hc4, EPI_HIGHWAY, 1, 0, Bh, Bl, 2 * D, 4 * T, B, Ah, Al, 2 * D, &L[11]));
    SSV_TRY(tf32_prepare(m->f_conv3, EPI_LN, 1, 0, Ah, Al, 2 * D, 4 * T, B, Bh, Bl, o_ld, &L[12]));
    SSV_TRY(tf32_prepare(m->f_conv4, EPI_LN_RELU, 1, 0, Bh, Bl, o_ld, 4 * T, B, Ah, Al, o_ld, &L[13]));
    SSV_TRY(tf32_prepare(m->f_conv5, EPI_LN_RELU, 1, 0, Ah, Al, o_ld, 4 * T, B, Bh, Bl, o_ld, &L[14]));
    SSV_TRY(tf32_prepare(m->f_conv6, EPI_LN_SIGMOID, 1, 0, Bh, Bl, o_ld, 4 * T, B, m->ws.buf[0], nullptr, o_ld, &L[15]));   // plain fp32
    m->fplan_B = B; m->fplan_T = T; m->fplan_out = m->ws.buf[0];
  }
  SSV_TRY(launch_transpose_in(mel, sb, sf, st_, B, m->F, T, Bh, f_ld, s));
  SSV_TRY(launch_split_tf32(Bh, Ah, Al, (size_t)B * T * f_ld, s));
  for (int i = 0; i < 16; ++i) SSV_TRY(tf32_run(m->fplan[i], s));
  SSV_TRY(launch_transpose_out(m->ws.buf[0], o_ld, B, O, 4 * T, out, s));
  return kOk;
}

static int ssrn_fwd_bf16(ssv_ssrn* m, const float* mel, long sb, long sf, long st_, int B, int T, float* out,
                         cudaStream_t s) {
  const int D = m->D, O = m->O;
  const int o_ld = round_up(O, 64);
  const int f_ld = round_up(m->F, 64);
  const size_t need = (size_t)B * 4 * T * o_ld;
  SSV_TRY(m->ws.ensure(need));
  if (need > m->tws_elems) {
    for (int i = 0; i < 2; ++i) {
      if (m->tws[i]) cudaFree(m->tws[i]);
      m->tws[i] = nullptr;
      if (cudaMalloc((void**)&m->tws[i], need * sizeof(__nv_bfloat16) + 256) != cudaSuccess) {
        cudaGetLastError();
        set_error("bf16 workspace cudaMalloc of %zu bytes failed", need * 2);
        m->tws_elems = 0;
        return kNoMem;
      }
    }
    m->tws_elems = need;
    m->plan_B = m->plan_T = 0;                     // the prepared launches hold the old buffer addresses
  }
  __nv_bfloat16* P = m->tws[0];
  __nv_bfloat16* Q = m->tws[1];
  if (m->plan_B != B || m->plan_T != T) {          // encode the TMA descriptors of this shape once
    m->plan_B = m->plan_T = 0;
    Tc2Launch* L = m->plan2;
    SSV_TRY(tc2_prepare(m->t_conv1, EPI_LN, 1, 0, P, f_ld, T, B, Q, D, 0, &L[0]));
    SSV_TRY(tc2_prepare(m->t_hc1, EPI_HIGHWAY, 1, 0, Q, D, T, B, P, D, 0, &L[1]));
    SSV_TRY(tc2_prepare(m->t_hc2, EPI_HIGHWAY, 3, 0, P, D, T, B, Q, D, 0, &L[2]));
    SSV_TRY(tc2_prepare(m->t_dc1, EPI_NONE, 1, 0, Q, D, T, B, P, 2 * D, 0, &L[3]));           // (B,T,2D) == (B,2T,D)
    SSV_TRY(tc2_prepare(m->t_u1h1, EPI_HIGHWAY, 1, 0, P, D, 2 * T, B, Q, D, 0, &L[4]));
    SSV_TRY(tc2_prepare(m->t_u1h2, EPI_HIGHWAY, 3, 0, Q, D, 2 * T, B, P, D, 0, &L[5]));
    SSV_TRY(tc2_prepare(m->t_dc2, EPI_NONE, 1, 0, P, D, 2 * T, B, Q, 2 * D, 0, &L[6]));
    SSV_TRY(tc2_prepare(m->t_u2h1, EPI_HIGHWAY, 1, 0, Q, D, 4 * T, B, P, D, 0, &L[7]));
    SSV_TRY(tc2_prepare(m->t_u2h2, EPI_HIGHWAY, 3, 0, P, D, 4 * T, B, Q, D, 0, &L[8]));
    SSV_TRY(tc2_prepare(m->t_conv2, EPI_LN, 1, 0, Q, D, 4 * T, B, P, 2 * D, 0, &L[9]));
    SSV_TRY(tc2_prepare(m->t_hc3, EPI_HIGHWAY, 1, 0, P, 2 * D, 4 * T, B, Q, 2 * D, 0, &L[10]));
    SSV_TRY(tc2_prepare(m->t_hc4, EPI_HIGHWAY, 1, 0, Q, 2 * D, 4 * T, B, P, 2 * D, 0, &L[11]));
    // the four 513-bin heads: 512 columns on the tensor cores + the 513th as a dot product in the epilogue warps
    SSV_TRY(tc2_prepare(m->t_conv3, EPI_LN, 1, 0, P, 2 * D, 4 * T, B, Q, o_ld, 0, &L[12]));
    SSV_TRY(tc2_prepare(m->t_conv4, EPI_LN_RELU, 1, 0, Q, o_ld, 4 * T, B, P, o_ld, 0, &L[13]));
    SSV_TRY(tc2_prepare(m->t_conv5, EPI_LN_RELU, 1, 0, P, o_ld, 4 * T, B, Q, o_ld, 0, &L[14]));
    SSV_TRY(tc2_prepare(m->t_conv6, EPI_LN_SIGMOID, 1, 0, Q, o_ld, 4 * T, B, out, o_ld, 2, &L[15]));   // (B, 513, 4T) directly
    m->plan_B = B; m->plan_T = T;
  }
  SSV_TRY(launch_transpose_in_bf16(mel, sb, sf, st_, B, m->F, T, P, f_ld, s));
  tc2_set_output(&m->plan2[15], out);              // the caller's buffer of this call
  for (int i = 0; i < 16; ++i) SSV_TRY(tc2_run(m->plan2[i], s));
  return kOk;
}

int ssv_ssrn_fwd(ssv_ssrn* m, const float* mel, long stride_b, long stride_f, long stride_t, int B, int T,
                 float* out, int precision, void* stream) {
  SSV_CHECK(m && mel && out, "ssrn_fwd: null pointer");
  SSV_CHECK(B > 0 && T > 0, "ssrn_fwd: empty input");
  cudaStream_t s = as_stream(stream);
  if (precision == SSV_PREC_FP32 && m->F % T32_BK == 0) return ssrn_fwd_tf32(m, mel, stride_b, stride_f, stride_t, B, T, out, s);
  if (precision == SSV_PREC_FP32 || precision == SSV_PREC_FP32_FFMA) return ssrn_fwd_f32(m, mel, stride_b, stride_f, stride_t, B, T, out, s);
  if (precision == SSV_PREC_BF16) return ssrn_fwd_bf16(m, mel, stride_b, stride_f, stride_t, B, T, out, s);
  set_error("ssrn_fwd: unknown precision %d", precision);
  return kInval;
}

// ------------------------------------------------------------------------------------------------
// One batch, host buffers in and out.  Enqueues everything on `s` (the lin D2H on the decoder's copy stream, under
// the SSRN of the next chunk) and records the slot's events; nothing here waits for the GPU.
static int synth_enqueue(ssv_text2mel* m, ssv_decoder* d, ssv_ssrn* sr, const int64_t* textid_host,
                         const float* spkemb_host, int B, int N, int n_frames, float* lin_host, float* mel_host,
                         float* A_host, int64_t* pma_traj_host, int t2m_precision, int ssrn_precision, int slot,
                         bool pipelined, cudaStream_t s) {
  SSV_CHECK(m && d && sr && textid_host && spkemb_host && lin_host, "synthesize_host: null pointer");
  SSV_CHECK(d->m == m, "synthesize_host: decoder belongs to another model");
  SSV_CHECK(B >= 1 && B <= d->maxB && N >= 1 && N <= d->maxN && n_frames >= 1 && n_frames <= d->maxT,
            "synthesize_host: shape (B=%d, N=%d, T=%d) exceeds decoder capacity (%d, %d, %d)", B, N, n_frames,
            d->maxB, d->maxN, d->maxT);
  const int H = m->H, F = m->F, O = sr->O, T = n_frames;
  const size_t key = ((size_t)B << 40) ^ ((size_t)N << 20) ^ (size_t)T;
  if (d->hs_key != key) {
    SSV_CHECK(!d->inflight[0] && !d->inflight[1], "synthesize_host: batch shape changed while a batch is in flight");
    SSV_CUDA(cudaStreamSynchronize(s));
    d->host_arena.~Arena();
    new (&d->host_arena) Arena();
    SSV_TRY(d->host_arena.alloc<int64_t>((size_t)B * N, &d->h_textid));
    SSV_TRY(d->host_arena.alloc<float>((size_t)B * m->E, &d->h_spk));
    SSV_TRY(d->host_arena.alloc<float>((size_t)B * H * N, &d->h_K));
    SSV_TRY(d->host_arena.alloc<float>((size_t)B * H * N, &d->h_V));
    SSV_TRY(d->host_arena.alloc<float>((size_t)B * F * T, &d->h_Y));
    SSV_TRY(d->host_arena.alloc<float>((size_t)B * N * T, &d->h_A));
    SSV_TRY(d->host_arena.alloc<long long>((size_t)T * B, &d->h_traj));
    for (int i = 0; i < 2; ++i) SSV_TRY(d->host_arena.alloc<float>((size_t)B * O * 4 * T, &d->h_lin[i]));
    for (int i = 0; i < 2; ++i) SSV_TRY(d->host_arena.alloc<__nv_bfloat16>((size_t)B * O * 4 * T, &d->h_lin16[i]));
    d->hs_key = key;
  }
  if (d->copy_stream == nullptr) {
    SSV_CUDA(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      SSV_CUDA(cudaEventCreateWithFlags(&d->copy_ev[i], cudaEventDisableTiming));
      SSV_CUDA(cudaEventCreateWithFlags(&d->done_ev[i], cudaEventDisableTiming));
      SSV_CUDA(cudaEventCreateWithFlags(&d->lin_ev[i], cudaEventDisableTiming));
    }
    SSV_CUDA(cudaHostAlloc((void**)&d->h_flags, sizeof(int) * 16, cudaHostAllocDefault));
    memset(d->h_flags, 0, sizeof(int) * 16);
  }
  float* lin_dev = d->h_lin[slot];
  SSV_CUDA(cudaMemcpyAsync(d->h_textid, textid_host, sizeof(int64_t) * B * N, cudaMemcpyHostToDevice, s));
  SSV_CUDA(cudaMemcpyAsync(d->h_spk, spkemb_host, sizeof(float) * B * m->E, cudaMemcpyHostToDevice, s));
  SSV_TRY(ssv_text_encoder_fwd(m, d->h_textid, B, N, d->h_K, d->h_V, t2m_precision, s));
  SSV_TRY(ssv_decoder_begin(d, d->h_K, d->h_V, d->h_spk, B, N, d->h_Y, d->h_A,
                            reinterpret_cast<int64_t*>(d->h_traj), T, s));
  SSV_TRY(ssv_decoder_run(d, T, s));
  // SSRN in utterance chunks: the D2H of chunk i (114 MB in all at B = 64) runs on a second stream under the SSRN of
  // chunk i + 1 -- and, with the submit / wait pair, under the next batch's TextEnc and decode
  {
    // one batch at a time: two chunks, so half of the D2H hides under the second chunk's SSRN; pipelined: one chunk --
    // the whole D2H hides under the next batch (measured at B = 64: 22.4 vs 23.0 ms sync, 21.4 vs 22.3 ms pipelined)
    const int nchunks = (!pipelined && B >= 32) ? 2 : 1;
    const int chunk = (B + nchunks - 1) / nchunks;
    const size_t lin_per_utt = (size_t)O * 4 * T;
    int ci = 0;
    for (int b0 = 0; b0 < B; b0 += chunk, ++ci) {
      const int nb = B - b0 < chunk ? B - b0 : chunk;
      SSV_TRY(ssv_ssrn_fwd(sr, d->h_Y + (size_t)b0 * F * T, (long)F * T, T, 1, nb, T, lin_dev + b0 * lin_per_utt,
                           ssrn_precision, s));
      if (d->lin_out_bf16)      // half-size result: the spectrogram leaves the device as bf16 (ssv_decoder_set_lin_output)
        SSV_TRY(launch_cast_f32_to_bf16(lin_dev + b0 * lin_per_utt, d->h_lin16[slot] + b0 * lin_per_utt, nb * lin_per_utt, s));
      SSV_CUDA(cudaEventRecord(d->copy_ev[ci & 1], s));
      SSV_CUDA(cudaStreamWaitEvent(d->copy_stream, d->copy_ev[ci & 1], 0));
      if (b0 + chunk >= B) {
        // Every device->host copy of the batch rides the copy stream, the small ones AHEAD of the last spectrogram
        // chunk.  (They used to follow on the compute stream: behind the 228 MB spectrogram copy in the DMA engine's
        // queue, and the next batch's kernels behind them in stream order -- every other batch completed 4 ms late
        // at 256 utterances.)  The compute stream only waits for the small ones, whose sources the next batch's
        // decoder overwrites.
        cudaStream_t cs = d->copy_stream;
        if (mel_host) SSV_CUDA(cudaMemcpyAsync(mel_host, d->h_Y, sizeof(float) * (size_t)B * F * T, cudaMemcpyDeviceToHost, cs));
        if (A_host) SSV_CUDA(cudaMemcpyAsync(A_host, d->h_A, sizeof(float) * (size_t)B * N * T, cudaMemcpyDeviceToHost, cs));
        if (pma_traj_host)
          SSV_CUDA(cudaMemcpyAsync(pma_traj_host, d->h_traj, sizeof(long long) * (size_t)T * B, cudaMemcpyDeviceToHost, cs));
        // error flags of this batch (decode abort, bad text id, tensor-core pipeline timeouts) -> pinned host words
        SSV_CUDA(cudaMemcpyAsync(d->h_flags + 8 * slot, d->abort_flag, sizeof(int), cudaMemcpyDeviceToHost, cs));
        SSV_CUDA(cudaMemcpyAsync(d->h_flags + 8 * slot + 1, m->err_flag, sizeof(int), cudaMemcpyDeviceToHost, cs));
        int* tcf[3] = {tc_err_flag_dev(), tc2_err_flag_dev(), tf32_err_flag_dev()};
        for (int i = 0; i < 3; ++i) {
          d->h_flags[8 * slot + 2 + i] = 0;
          if (tcf[i]) SSV_CUDA(cudaMemcpyAsync(d->h_flags + 8 * slot + 2 + i, tcf[i], sizeof(int), cudaMemcpyDeviceToHost, cs));
        }
        SSV_CUDA(cudaEventRecord(d->done_ev[slot], cs));
        SSV_CUDA(cudaStreamWaitEvent(s, d->done_ev[slot], 0));
      }
      if (d->lin_out_bf16)
        SSV_CUDA(cudaMemcpyAsync(reinterpret_cast<__nv_bfloat16*>(lin_host) + b0 * lin_per_utt, d->h_lin16[slot] + b0 * lin_per_utt,
                                 sizeof(__nv_bfloat16) * nb * lin_per_utt, cudaMemcpyDeviceToHost, d->copy_stream));
      else
        SSV_CUDA(cudaMemcpyAsync(lin_host + b0 * lin_per_utt, lin_dev + b0 * lin_per_utt, sizeof(float) * nb * lin_per_utt,
                                 cudaMemcpyDeviceToHost, d->copy_stream));
    }
    SSV_CUDA(cudaEventRecord(d->lin_ev[slot], d->copy_stream));
  }
  d->inflight[slot] = true;
  d->slot_bf16[slot] = ssrn_precision == SSV_PREC_BF16;
  d->slot_stream[slot] = s;
  return kOk;
}

static int synth_wait(ssv_decoder* d, int slot) {
  SSV_CHECK(d && (slot == 0 || slot == 1) && d->inflight[slot], "synthesize_wait: no batch in flight under ticket %d", slot);
  d->inflight[slot] = false;
  SSV_CUDA(cudaEventSynchronize(d->done_ev[slot]));
  SSV_CUDA(cudaEventSynchronize(d->lin_ev[slot]));
  const int abort_code = d->h_flags[8 * slot], bad_id = d->h_flags[8 * slot + 1];
  if (abort_code != 0) {
    set_error("decode kernel aborted (hand-off timeout, code %d)", abort_code);
    cudaMemsetAsync(d->abort_flag, 0, sizeof(int), d->slot_stream[slot]);
    return kState;
  }
  if (bad_id != 0) {
    set_error("text id outside [0, vocab_len)");
    cudaMemsetAsync(d->m->err_flag, 0, sizeof(int), d->slot_stream[slot]);
    return kInval;
  }
  // a tensor-core pipeline timed out: the synchronous check reads the flag again, reports and clears it
  if (d->h_flags[8 * slot + 2] != 0) SSV_TRY(tc_check_error());
  if (d->h_flags[8 * slot + 3] != 0) SSV_TRY(tc2_check_error());
  if (d->h_flags[8 * slot + 4] != 0) SSV_TRY(tf32_check_error());
  return kOk;
}

int ssv_synthesize_host(ssv_text2mel* m, ssv_decoder* d, ssv_ssrn* sr, const int64_t* textid_host,
                        const float* spkemb_host, int B, int N, int n_frames, float* lin_host, float* mel_host,
                        float* A_host, int64_t* pma_traj_host, int t2m_precision, int ssrn_precision,
                        void* stream) {
  SSV_CHECK(d && !d->inflight[0] && !d->inflight[1], "synthesize_host: a submitted batch is still in flight (wait for it first)");
  SSV_TRY(synth_enqueue(m, d, sr, textid_host, spkemb_host, B, N, n_frames, lin_host, mel_host, A_host, pma_traj_host,
                        t2m_precision, ssrn_precision, 0, false, as_stream(stream)));
  return synth_wait(d, 0);
}

int ssv_synthesize_host_submit(ssv_text2mel* m, ssv_decoder* d, ssv_ssrn* sr, const int64_t* textid_host,
                               const float* spkemb_host, int B, int N, int n_frames, float* lin_host, float* mel_host,
                               float* A_host, int64_t* pma_traj_host, int t2m_precision, int ssrn_precision,
                               void* stream, int* ticket) {
  SSV_CHECK(d && ticket, "synthesize_host_submit: null pointer");
  const int slot = d->inflight[0] ? 1 : 0;
  SSV_CHECK(!d->inflight[slot], "synthesize_host_submit: two batches are already in flight");
  SSV_TRY(synth_enqueue(m, d, sr, textid_host, spkemb_host, B, N, n_frames, lin_host, mel_host, A_host, pma_traj_host,
                        t2m_precision, ssrn_precision, slot, true, as_stream(stream)));
  *ticket = slot;
  return kOk;
}

int ssv_synthesize_host_wait(ssv_decoder* d, int ticket) { return synth_wait(d, ticket); }

// ------------------------------------------------------------------------------------------------
long long ssv_griffin_lim_workspace(int B, int T) {
  if (B < 1 || T < 2) return 0;
  return (long long)griffin_lim_workspace_floats(B, T);
}

int ssv_griffin_lim(const float* S, const float* angles0_ri, int B, int F, int T, int n_iter, int hop, int win_length,
                    float momentum, float* y, float* workspace, long long workspace_floats, void* stream) {
  SSV_CHECK(S && angles0_ri && y && workspace, "griffin_lim: null pointer");
  SSV_CHECK(F == 513 && hop == 256 && win_length == 1024, "griffin_lim: only n_fft = win_length = 1024, hop = 256 (the reference's STFT) is built, got F=%d hop=%d win=%d", F, hop, win_length);
  // the STFT kernel reflects an index once: the signal (256 (T - 1) samples) must be longer than the 512-sample pad
  SSV_CHECK(B >= 1 && T >= 4, "griffin_lim: need B >= 1 and at least 4 frames (single reflection of the 512-sample padding)");
  SSV_CHECK(n_iter >= 0 && momentum >= 0.f && momentum < 1.f, "griffin_lim: bad n_iter / momentum");
  SSV_CHECK(workspace_floats >= (long long)griffin_lim_workspace_floats(B, T), "griffin_lim: workspace too small");
  return launch_griffin_lim(S, angles0_ri, B, T, n_iter, momentum, y, workspace, as_stream(stream));
}

int ssv_spec_features(const float* stft_ri, int F, int T, const float* melfb, int n_mels, int log_feature,
                      float norm_power, float ref_db, float max_db, int reduction, float* lin_norm, float* mel_red,
                      float* workspace, void* stream) {
  SSV_CHECK(stft_ri && melfb && lin_norm && mel_red && workspace, "spec_features: null pointer");
  SSV_CHECK(F >= 1 && F <= 1025 && T >= 1, "spec_features: bad spectrogram shape (%d, %d)", F, T);
  SSV_CHECK(n_mels >= 1 && n_mels <= 128, "spec_features: n_mels must be in [1, 128]");
  SSV_CHECK(reduction >= 1, "spec_features: reduction must be >= 1");
  SSV_CHECK(log_feature || norm_power > 0.f, "spec_features: norm_power must be positive");
  SSV_CHECK(!log_feature || max_db > 0.f, "spec_features: MAX_DB must be positive");
  return launch_spec_features(stft_ri, F, T, melfb, n_mels, log_feature, norm_power, ref_db, max_db, reduction, lin_norm,
                              mel_red, workspace, as_stream(stream));
}

int ssv_deemphasis(const float* x, float* y, int B, long n, float coeff, void* stream) {
  SSV_CHECK(x && y, "deemphasis: null pointer");
  SSV_CHECK(B >= 0 && n >= 0, "deemphasis: negative size");
  SSV_CHECK(coeff > -1.f && coeff < 1.f, "deemphasis: |coeff| must be < 1");
  if (B == 0 || n == 0) return kOk;
  return launch_deemphasis(x, y, B, n, coeff, as_stream(stream));
}

}  // extern "C"
