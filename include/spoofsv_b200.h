/* spoofsv_b200 -- C ABI of the B200-native Text2Mel + SSRN synthesis hot path.
 *
 * The reference (MingruiYuan/SpoofSV) has no FFI layer: its boundary for this path is the
 * Python nn.Module API of models/TTSModel.py plus the state_dict layout.  The
 * entry points below are what the `forward()` shims of the drop-in `melSyn` / `SSRN` classes
 * (spoofsv_b200/models/TTSModel.py) bind through ctypes; each one names the reference code
 * it replaces.  Plain pointers and sizes only: no torch types, no C++ in the signatures.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error (SSV_E*); ssv_last_error() gives
 *     the message for the calling thread;
 *   - "dev" pointers are CUDA device pointers on the current device, "host" pointers are
 *     ordinary (ideally pinned) host memory; tensors are fp32, contiguous unless strides are
 *     given, in the reference's channels-first layout (B, C, T);
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are
 *     asynchronous on that stream unless stated otherwise;
 *   - parameters are passed by state_dict key: parallel arrays of names, device pointers and
 *     element counts (order irrelevant, every key of SURVEY.md Appendix B required).  The
 *     library keeps repacked private copies; the caller's tensors may be freed afterwards.
 */
#ifndef SPOOFSV_B200_H_
#define SPOOFSV_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSV_OK 0
#define SSV_EINVAL 1   /* bad argument / unsupported shape */
#define SSV_ECUDA 2    /* a CUDA runtime call failed */
#define SSV_ENOMEM 3
#define SSV_ESTATE 4   /* call sequence error, decode kernel aborted */

#define SSV_PREC_FP32 0        /* FP32-accurate, <= 1e-4 max-abs vs the reference: tcgen05 tensor cores with 3xTF32 split
                                * operands where that arm is built (TextEnc, highwayConv), CUDA-core FP32 FMA elsewhere */
#define SSV_PREC_BF16 1        /* tcgen05 BF16 tensor cores, FP32 accumulate, <= 2e-2 rel-L2 */
#define SSV_PREC_FP32_FFMA 2   /* FP32-accurate on the CUDA cores only (cross-check of the split tensor-core arm) */

typedef struct ssv_text2mel ssv_text2mel;   /* packed Text2Mel weights (melSyn) */
typedef struct ssv_decoder ssv_decoder;     /* incremental AR decode state */
typedef struct ssv_ssrn ssv_ssrn;           /* packed SSRN weights + workspace */

int ssv_version(void);
const char* ssv_last_error(void);
/* SM count and compute capability of the current device. */
int ssv_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Kernels launched by the calling thread since the last reset (bench "gpu_launches"). */
long ssv_launch_count(int reset);

/* ---- single highwayConv, replaces highwayConv.forward (models/TTSModel.py:63-84) ----------
 * x, y: dev (B, d, T).  conv_w (2d, d, k), conv_b (2d), ln*_w/b (d): dev.  k in {1,3};
 * d in {256, 512}. */
int ssv_highway_conv_fwd(const float* x, const float* conv_w, const float* conv_b,
                         const float* ln1_w, const float* ln1_b, const float* ln2_w, const float* ln2_b,
                         int B, int d, int T, int k, int dilation, int causal,
                         float* y, int precision, void* stream);

/* ---- Text2Mel ------------------------------------------------------------------------------
 * replaces melSyn.__init__ + load_state_dict (models/TTSModel.py:236-251): 214 tensors. */
int ssv_text2mel_create(const char* const* names, const float* const* dev_ptrs, const int64_t* numels,
                        int n_params, int vocab_len, int spkemb_dim, int textemb_dim, int freq_bins,
                        int hidden_dim, ssv_text2mel** out);
int ssv_text2mel_destroy(ssv_text2mel* m);

/* replaces textEncoder.forward (models/TTSModel.py:126-140).  textid: dev int64 (B, 1, N);
 * K, V: dev (B, hidden, N). */
int ssv_text_encoder_fwd(ssv_text2mel* m, const int64_t* textid, int B, int N, float* K, float* V,
                         int precision, void* stream);

/* Synchronise `stream` and report text ids outside [0, vocab_len) seen by the embedding gather since the last
 * check (textEmbedding.forward, models/TTSModel.py:25-35, would raise from scatter_): SSV_EINVAL.  The id range is
 * verified on the device inside the gather, so the hot path needs no host round trip before launching. */
int ssv_text2mel_check(ssv_text2mel* m, void* stream);

/* replaces the train branch of melSyn.forward (models/TTSModel.py:263-273; teacher-forced, as the training loops
 * call it: train/adversarial_wasserstein_gp.py:277-278): full-sequence forward, unmasked attention.  melspec: dev
 * (B, F, T); textid: dev int64 (B, 1, N); spkemb: dev (B, E, 1); Y: dev (B, F, T); A: dev (B, N, T).  Forward only:
 * the backward pass / optimiser / gradient allreduce of the training step is the next row of the scope table. */
int ssv_text2mel_train_fwd(ssv_text2mel* m, const float* melspec, const int64_t* textid, const float* spkemb,
                           int B, int N, int T, float* Y, float* A, int precision, void* stream);

/* Incremental decoder: replaces the eval branch of melSyn.forward (models/TTSModel.py:275-300)
 * as driven by the AR loops (generate_test_utterances.py:105-116, synthesize.py:103-109). */
int ssv_decoder_create(ssv_text2mel* m, int max_batch, int max_text, int max_frames, ssv_decoder** out);
int ssv_decoder_destroy(ssv_decoder* d);
/* Start a new batch of utterances (the reference's T == 1 call).  K, V: dev (B, hidden, N);
 * spkemb: dev (B, E, 1).  Y (B, F, t_cap), A (B, N, t_cap), pma_traj int64 (t_cap, B) are
 * caller-owned dev buffers the decoder fills column by column; A is zeroed here. */
int ssv_decoder_begin(ssv_decoder* d, const float* K, const float* V, const float* spkemb, int B, int N,
                      float* Y, float* A, int64_t* pma_traj, int t_cap, void* stream);
/* One frame, caller-driven (drop-in protocol): x = the newest column of the caller's melspec,
 * element (b, f) at x[b*x_stride_b + f*x_stride_f]; pma_in: dev int64 (B,) or NULL to chain the
 * decoder's own argmax.  Writes Y[:, :, t], A[:, :, t], pma_traj[t]; advances t. */
int ssv_decoder_step(ssv_decoder* d, const float* x, long x_stride_b, long x_stride_f,
                     const int64_t* pma_in, void* stream);
/* n_steps frames free-running in one launch: x_0 = 0 (or Y[:, :, t-1] when continuing),
 * x_t = y_{t-1}. */
int ssv_decoder_run(ssv_decoder* d, int n_steps, void* stream);
/* Front-end shape of the decode kernel for the batches begun after this call: rows per micro-batch (1, 2, 4),
 * front-end warps per row (1, 2, 4) and front-end warps per CTA (4 or 8); 0 = the measured choice for the batch
 * size.  Tuning and test hook: every shape computes the same values (tests/test_gpu_parity.py runs them all
 * against the oracle). */
int ssv_decoder_set_plan(ssv_decoder* d, int rows_per_microbatch, int warps_per_row, int front_end_warps);
/* Frames decoded since begin. */
int ssv_decoder_frames(const ssv_decoder* d);
/* Synchronise `stream` and report a decode-kernel abort (barrier timeout) if one happened. */
int ssv_decoder_check(ssv_decoder* d, void* stream);

/* ---- SSRN -----------------------------------------------------------------------------------
 * replaces SSRN.__init__ + load_state_dict (models/TTSModel.py:321-340): 76 tensors. */
int ssv_ssrn_create(const char* const* names, const float* const* dev_ptrs, const int64_t* numels,
                    int n_params, int freq_bins, int output_bins, int ssrn_dim, ssv_ssrn** out);
int ssv_ssrn_destroy(ssv_ssrn* m);
/* replaces SSRN.forward (models/TTSModel.py:342-362).  mel: dev (B, F, T) with element strides;
 * out: dev (B, O, 4T) contiguous. */
int ssv_ssrn_fwd(ssv_ssrn* m, const float* mel, long stride_b, long stride_f, long stride_t, int B, int T,
                 float* out, int precision, void* stream);

/* ---- whole utterance batch with HOST buffers -----------------------------------------------
 * replaces generate_test_utterances.py:105-124 (AR loop, SSRN, .cpu()).  Copies textid (B, N)
 * int64 and spkemb (B, E) fp32 from the host, runs TextEnc + n_frames of decode + SSRN, copies
 * lin (B, O, 4*n_frames) and, when non-NULL, mel (B, F, n_frames), A (B, N, n_frames) and
 * pma_traj (n_frames, B) back, and synchronises the stream. */
int ssv_synthesize_host(ssv_text2mel* m, ssv_decoder* d, ssv_ssrn* s, const int64_t* textid_host,
                        const float* spkemb_host, int B, int N, int n_frames, float* lin_host,
                        float* mel_host, float* A_host, int64_t* pma_traj_host,
                        int t2m_precision, int ssrn_precision, void* stream);

/* Pipelined form of the same call for corpus runs: submit enqueues one batch and returns a ticket (0 or 1) without
 * waiting for the GPU; wait blocks until that batch's host buffers are filled and reports its errors.  Two
 * batches may be in flight, so the device->host copy of batch i (114 MB at B = 64) overlaps TextEnc / decode of
 * batch i + 1.  The host buffers of an in-flight batch must stay valid and distinct until its wait returns; all
 * batches in flight must have the same (B, N, n_frames). */
int ssv_synthesize_host_submit(ssv_text2mel* m, ssv_decoder* d, ssv_ssrn* s, const int64_t* textid_host,
                               const float* spkemb_host, int B, int N, int n_frames, float* lin_host,
                               float* mel_host, float* A_host, int64_t* pma_traj_host,
                               int t2m_precision, int ssrn_precision, void* stream, int* ticket);
int ssv_synthesize_host_wait(ssv_decoder* d, int ticket);
/* Element type of `lin_host` for the calls above: 0 = fp32 (default, what the reference's .cpu() returns), 1 = bf16
 * (the same (B, O, 4*n_frames) elements as 2-byte bfloat16: half the device->host bytes -- the copy is what limits
 * the end-to-end rate when eight GPUs write into one host; meant for the BF16 SSRN arm, whose values carry bf16
 * operand rounding already). */
int ssv_decoder_set_lin_output(ssv_decoder* d, int bf16);

/* ---- training step, first building block (next row of the scope table) ---------------------
 * Backward of one highwayConv, FP32: what autograd computes for highwayConv.forward
 * (models/TTSModel.py:63-84) inside loss.backward() of the training step
 * (train/adversarial_wasserstein_gp.py:298-300).  x, dy, dx: dev (B, d, T); dconv_w (2d, d, k), dconv_b (2d),
 * dln*_w/b (d): dev, overwritten (not accumulated). */
int ssv_highway_conv_bwd(const float* x, const float* dy, const float* conv_w, const float* conv_b, const float* ln1_w,
                         const float* ln1_b, const float* ln2_w, const float* ln2_b, int B, int d, int T, int k,
                         int dilation, int causal, const float* h_saved, float* dx, float* dconv_w, float* dconv_b,
                         float* dln1_w, float* dln1_b, float* dln2_w, float* dln2_b,
                         int precision /* as the forward: recomputed H, dgrad and wgrad on the tensor cores or the CUDA cores */,
                         int channels_last /* 1: x, dy and dx are (B, T, d) */, void* stream);

/* ---- the small layers of the Text2Mel training graph, FP32 (reference: the autograd of the nn.Conv1d(kernel 1) +
 * nn.LayerNorm pairs, models/TTSModel.py:128-131, 173-180, 218-230; the unmasked attention of the train branch,
 * :268-272; textEmbedding, :25-35; the speaker projections, :172-173).  All tensors channels-first fp32 on the device. */
/* y (B, n, T) = LN(W relu?(x) + b (+ sb per utterance)); h_save: (B T) x round_up(n, 64) floats for the backward. */
int ssv_conv_ln_fwd_save(const float* x, const float* w, const float* b, const float* sb /* (B, n) or NULL */,
                         const float* ln_w, const float* ln_b, int B, int cin, int n, int T, int relu_in, float* y,
                         float* h_save, int precision /* SSV_PREC_FP32: conv on the tensor cores (3xTF32); SSV_PREC_FP32_FFMA */,
                         void* stream);
int ssv_conv_ln_bwd(const float* x, const float* dy, const float* w, const float* ln_w, const float* h_saved, int B, int cin,
                    int n, int T, int relu_in, float* dx /* or NULL */, float* dw, float* db, float* dsb /* (B, n) or NULL */,
                    float* dln_w, float* dln_b, int precision /* dgrad and wgrad on the tensor cores or the CUDA cores */,
                    void* stream);
/* kv (B, 512, N) = [K ; V], q (B, 256, T) -> A (B, N, T) = softmax_n(K^T q / 16), rq (B, 512, T) = [V A ; q] */
int ssv_attention_train_fwd(const float* kv, const float* q, int B, int N, int T, float* A, float* rq, void* stream);
int ssv_attention_train_bwd(const float* kv, const float* q, const float* A, const float* dA /* or NULL */, const float* drq,
                            int B, int N, int T, float* dkv, float* dq, void* stream);
/* ids (B, N) int64, weight (E, vocab), bias (E) -> y (B, E, N) */
int ssv_text_embedding_fwd(const int64_t* ids, const float* weight, const float* bias, int B, int N, int vocab, int E, float* y,
                           void* stream);
int ssv_text_embedding_bwd(const int64_t* ids, const float* dy, int B, int N, int vocab, int E, float* dweight, float* dbias,
                           void* stream);
/* y (B, out) = x (B, in) W^T + b */
int ssv_linear_small_fwd(const float* x, const float* w, const float* b, int B, int in_f, int out_f, float* y, void* stream);
int ssv_linear_small_bwd(const float* x, const float* dy, int B, int in_f, int out_f, float* dw, float* db, void* stream);
/* Training-time forward: like ssv_highway_conv_fwd (FP32) and also writes H = conv(x) + b, dev ((B T), 2d) in the
 * library's row layout, for h_saved of ssv_highway_conv_bwd (NULL there: the conv is recomputed). */
int ssv_highway_conv_fwd_save(const float* x, const float* conv_w, const float* conv_b, const float* ln1_w,
                              const float* ln1_b, const float* ln2_w, const float* ln2_b, int B, int d, int T, int k,
                              int dilation, int causal, float* y, float* h_save,
                              int precision /* SSV_PREC_FP32: conv on the tensor cores (3xTF32); SSV_PREC_FP32_FFMA */,
                              int channels_last /* 1: x and y are (B, T, d) instead of the reference's (B, d, T) */, void* stream);

/* ---- waveform stage (next row of the scope table) -------------------------------------------
 * De-emphasis of B waveforms of n samples each, y[n] = x[n] + coeff * y[n-1]: replaces
 * scipy.signal.lfilter([1], [1, -PREEMPH], time_signal) (generate_test_utterances.py:136,
 * synthesize.py:144).  x, y: dev (B, n) contiguous; in place (y == x) is allowed. */
int ssv_deemphasis(const float* x, float* y, int B, long n, float coeff, void* stream);

/* Fast Griffin-Lim: replaces librosa.core.griffinlim(S=spec, n_iter=64, hop_length=256, win_length=1024)
 * (generate_test_utterances.py:133, synthesize.py:141; librosa 0.7: momentum 0.99, periodic Hann window, centred
 * reflect-padded frames) for a whole batch.  S: dev (B, 513, T) magnitudes; angles0_ri: dev (B, 513, T) complex64
 * initial phases as interleaved (re, im) -- librosa draws them as exp(2j pi rand); y: dev (B, 256 (T - 1)) samples;
 * workspace: dev, ssv_griffin_lim_workspace(B, T) floats.  Only the reference's STFT shape is built. */
long long ssv_griffin_lim_workspace(int B, int T);
int ssv_griffin_lim(const float* S, const float* angles0_ri, int B, int F, int T, int n_iter, int hop, int win_length,
                    float momentum, float* y, float* workspace, long long workspace_floats, void* stream);

/* ---- feature front-end (next row of the scope table) -------------------------------------------
 * Spectrogram features of one utterance: replaces data/dataset.py:97-118 (np.abs of the STFT, np.dot with
 * librosa.filters.mel, the LOG_FEATURE / NORM_POWER.ANALYSIS normalisation, every `reduction`-th mel frame,
 * the linear spectrogram cut to reduction * (T / reduction) frames).
 * stft_ri: dev (F, T) complex64 as interleaved (re, im) floats -- the layout of the STFT of the trimmed,
 * pre-emphasised signal; melfb: dev (n_mels, F) filter bank; lin_norm: dev (F, reduction * (T / reduction));
 * mel_red: dev (n_mels, T / reduction); workspace: dev, at least F * T + n_mels * T + 2 floats. */
int ssv_spec_features(const float* stft_ri, int F, int T, const float* melfb, int n_mels, int log_feature,
                      float norm_power, float ref_db, float max_db, int reduction, float* lin_norm, float* mel_red,
                      float* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPOOFSV_B200_H_ */
