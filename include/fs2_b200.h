/* fs2_b200.h -- C ABI of libfs2b200.so: the B200 (sm_100a) implementation of the
 * FastSpeech2 acoustic-model inference forward pass.
 *
 * The reference (Napoliee/Expressive-FastSpeech2-Mandarin) has no FFI of its own: its
 * boundary for this path is the Python call `FastSpeech2.forward` (model/fastspeech2.py:73-149).
 * Each entry point below names the reference code it replaces; the ctypes binding a
 * maintainer would add is in INTEGRATION.md and, in full, in
 * expressive-fastspeech2-mandarin_b200/_lib.py.
 *
 * Conventions: every function returns 0 (FS2_OK) or a negative FS2_ERR_* code and never
 * throws; fs2_last_error() gives the message.  All tensor memory is caller-owned and passed
 * as raw device pointers (row-major, contiguous); the library owns only its repacked
 * weights and a workspace that grows monotonically.  Work is enqueued on the caller's
 * stream; the one blocking point is the size read-back at the end of fs2_forward_stage1
 * (output shapes depend on the predicted durations).  A context is bound to one device and
 * is not thread-safe.  There is no CPU implementation behind any of these calls.
 */
#ifndef FS2_B200_H
#define FS2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fs2_ctx fs2_ctx;
typedef void* fs2_stream; /* cudaStream_t */

enum {
  FS2_OK = 0,
  FS2_ERR_INVALID = -1,     /* bad argument / shape / missing weight */
  FS2_ERR_CUDA = -2,        /* a CUDA runtime or driver call failed */
  FS2_ERR_STATE = -3,       /* call order violated (e.g. forward before prepare) */
  FS2_ERR_UNSUPPORTED = -4  /* hyper-parameters the kernels are not compiled for */
};

/* Arithmetic of the tensor-core contractions (accumulation, LayerNorm and softmax are fp32 in all modes).
 * TF32: fp32 activations, operands rounded to TF32 -- the arithmetic class of the reference's own GPU run (cuDNN convs
 * use TF32 by default).  BF16: every Conv1d/Linear takes bf16 operands (weights rounded once, activations written as
 * bf16 by the producing epilogue); the residual stream, LayerNorm, softmax and all outputs stay fp32.
 * TF32X3 ("parity" mode): every Conv1d/Linear runs as a split-operand contraction a_hi*w_hi + a_lo*w_hi + a_hi*w_lo on
 * the same tensor-core kernel (three times the MMAs), which brings the mel to ~1e-4 of the fp64 reference
 * (tests/test_gpu_forward.py states the tolerance); the attention products stay TF32. */
enum { FS2_MATH_TF32 = 0, FS2_MATH_BF16 = 1, FS2_MATH_TF32X3 = 2 };

/* Model dimensions that are data (table sizes); the layer structure is fixed to
 * config/ESD-Chinese-Singing-MFA/model.yaml (d_model 256, 2 heads, 4+6 FFT blocks,
 * conv 9/1 x 1024, predictors 256 k3, 256 bins, 80 mels, PostNet 5 x 512 k5). */
typedef struct {
  int32_t n_src_vocab;  /* transformer/Models.py:40   */
  int32_t n_speaker;    /* model/fastspeech2.py:37-41 */
  int32_t n_emotion;    /* model/fastspeech2.py:52    */
  int32_t n_arousal;    /* model/fastspeech2.py:53    */
  int32_t n_valence;    /* model/fastspeech2.py:54    */
  int32_t max_seq_len;  /* model.yaml:27; position_enc has max_seq_len+1 rows */
  int32_t math_mode;    /* FS2_MATH_*  */
  /* preprocess.yaml preprocessing.{pitch,energy}.feature (model/modules.py:28-35): 0 = phoneme_level (the
   * predictor runs before the LengthRegulator, modules.py:114-125), 1 = frame_level (after it, :139-148). */
  int32_t pitch_frame_level;
  int32_t energy_frame_level;
} fs2_config;

/* Arguments of FastSpeech2.forward (model/fastspeech2.py:73-91). */
typedef struct {
  int32_t batch;
  int32_t max_src_len;     /* == texts.shape[1] */
  const int64_t* speakers; /* device [B] */
  const int64_t* emotions; /* device [B] */
  const int64_t* arousals; /* device [B] */
  const int64_t* valences; /* device [B] */
  const int64_t* texts;    /* device [B, max_src_len] */
  const int64_t* src_lens; /* device [B] */
  const float* p_targets;  /* device [B, max_src_len] or NULL (model/modules.py:82-83); [B, max_mel_len] when the
                              feature is frame_level (must stay valid until fs2_forward_stage2 has been enqueued) */
  const float* e_targets;  /* device [B, max_src_len] or NULL (model/modules.py:93-94); frame_level: as above.  Must stay
                              valid until fs2_forward_stage2 has been enqueued in every configuration: the energy
                              bucketize + embedding add runs inside the length regulator of stage 2 */
  const float* d_targets;  /* device [B, max_src_len] or NULL (model/modules.py:128-130) */
  float p_control;         /* scales pitch AND energy (model/modules.py:118-125) */
  float e_control;         /* accepted and ignored, as in the reference */
  float d_control;         /* model/modules.py:132-135 */
  int32_t max_mel_len;     /* 0 = max over the batch (utils/tools.py:361-364) */
} fs2_inputs;

/* Phoneme-side results; all arrays caller-allocated. */
typedef struct {
  float* pitch;       /* device [B, max_src_len]  output[2] */
  float* energy;      /* device [B, max_src_len]  output[3] */
  float* log_d;       /* device [B, max_src_len]  output[4] */
  float* d_rounded;   /* device [B, max_src_len]  output[5] */
  uint8_t* src_mask;  /* device [B, max_src_len]  output[6], 1 = padding */
  int64_t* mel_lens;  /* device [B]               output[9] */
  /* written on the host before fs2_forward_stage1 returns: */
  int64_t total_frames; /* sum of mel_lens */
  int32_t max_mel_len;  /* T_max the caller must allocate stage-2 outputs with */
} fs2_stage1_out;

typedef struct {
  float* mel;         /* device [B, max_mel_len, 80]  output[0]; padding rows = mel_linear.bias */
  float* postnet;     /* device [B, max_mel_len, 80]  output[1]; padding rows = mel_linear.bias (outside the contract) */
  uint8_t* mel_mask;  /* device [B, max_mel_len]      output[7], 1 = padding */
  /* frame_level features only (NULL otherwise): the predictions live on the frame axis and are written here
   * instead of fs2_stage1_out.pitch / .energy (which are then left zero-filled) */
  float* pitch_frames;   /* device [B, max_mel_len]  output[2] when pitch_frame_level  */
  float* energy_frames;  /* device [B, max_mel_len]  output[3] when energy_frame_level */
} fs2_stage2_io;

/* --- lifetime ------------------------------------------------------------------------ */
int fs2_create(const fs2_config* cfg, int device, fs2_ctx** out);
void fs2_destroy(fs2_ctx* ctx);
const char* fs2_last_error(const fs2_ctx* ctx); /* ctx may be NULL: last create() error */
int fs2_version(void);

/* --- weights: replaces model.load_state_dict(ckpt["model"]) (utils/model.py:16-21) ----- */
/* `key` is the reference state-dict key; `dev_ptr` is fp32 on the context's
 * device (int64 for num_batches_tracked, which is ignored).  The tensor is copied. */
int fs2_set_weight(fs2_ctx* ctx, const char* key, const void* dev_ptr, const int64_t* shape, int ndim);
/* Repack into kernel layouts: QKV concat, conv [Cout,Cin,k] -> [k][Cout][Cin], BatchNorm folded
 * into the PostNet convs (transformer/Layers.py:129-137), operand rounding, TMA descriptors. */
int fs2_prepare(fs2_ctx* ctx, fs2_stream stream);

/* --- forward: replaces FastSpeech2.forward (model/fastspeech2.py:92-149) ---------------- */
/* stage 1 = masks + Encoder + conditioning + VarianceAdaptor up to the durations
 * (fastspeech2.py:92-131, modules.py:102-135); blocks until mel_lens are known. */
int fs2_forward_stage1(fs2_ctx* ctx, fs2_stream stream, const fs2_inputs* in, fs2_stage1_out* out);
/* stage 2 = LengthRegulator + Decoder + mel_linear + PostNet (modules.py:136-137,167-194;
 * fastspeech2.py:133-136); asynchronous on `stream`. */
int fs2_forward_stage2(fs2_ctx* ctx, fs2_stream stream, const fs2_stage2_io* io);
/* Device->host read of the result the callers actually consume (`synth_samples` slices predictions[1][i, :mel_len],
 * utils/tools.py:228-233): the PACKED postnet mel rows, utterance b at rows [starts[b], starts[b] + mel_lens[b]) of
 * `host_rows` ([rows, 80] fp32), separated by 12 reserved rows -- one third of the bytes of the padded [B, T_max, 80]
 * tensor at batch 64.  Enqueued on `stream` after fs2_forward_stage2; `host_rows` / `host_starts` ([batch + 1] int32)
 * should be pinned; `*rows_out` (written before the call returns) = rows to expect; max_rows = capacity of host_rows. */
int fs2_read_packed_postnet(fs2_ctx* ctx, fs2_stream stream, float* host_rows, int64_t max_rows, int32_t* host_starts,
                            int64_t* rows_out);
/* --- hand-over between contexts: one batch sharded over several GPUs, re-balanced by FRAMES ---------------------
 * The shards of a batch can only be balanced by phonemes before stage 1 (the durations are its result); the decoder's
 * cost follows the frames.  A caller that owns one context per GPU therefore runs stage 1 on
 * phoneme-balanced shards, exchanges utterances so that the frames are balanced, and runs stage 2 where the utterance
 * landed (expressive-fastspeech2-mandarin_b200/partition.py does this over torch.distributed / NCCL).
 * fs2_export_stage1: after fs2_forward_stage1, writes what the LengthRegulator consumes (model/modules.py:126-137):
 *   hidden [B, max_src_len, 256] = x + pitch embedding + energy embedding (zero at padding) and reps [B, max_src_len] =
 *   max(int(d), 0), the integer repeat counts.  phoneme_level features only.
 * fs2_import_stage1: takes the place of fs2_forward_stage1 on the receiving context: lays the given utterances out,
 *   computes mel_lens = row sums of reps, and blocks for the sizes exactly like stage 1; fs2_forward_stage2 follows.
 *   max_mel_len: 0 = maximum over the batch.  Results equal the reference's forward on the RECEIVING batch with
 *   d_targets / p_targets / e_targets forced (the PostNet tails follow that batch's T_max). */
int fs2_export_stage1(fs2_ctx* ctx, fs2_stream stream, float* hidden, int32_t* reps);
int fs2_import_stage1(fs2_ctx* ctx, fs2_stream stream, const float* hidden, const int32_t* reps, const int64_t* src_lens,
                      int batch, int max_src_len, int max_mel_len, int64_t* mel_lens, int64_t* total_frames,
                      int32_t* max_mel_len_out);
/* Number of kernels the last stage1+stage2 pair launched. */
int fs2_last_launch_count(const fs2_ctx* ctx);

/* on != 0: fs2_forward_stage1 enqueues all of stage 2 except the final unpack into the caller's tensors before it returns (none
 * of it needs a caller buffer), so the device does not idle while the caller allocates [B, max_mel_len, 80] outputs and calls
 * fs2_forward_stage2, which then only unpacks.  Same kernels, same results.  Ignored with frame_level features and debug taps.
 * A caller that still reads the previous forward's packed rows asynchronously (fs2_read_packed_postnet on another stream) must
 * leave it off for that forward: the stage-2 body overwrites those rows.  (The reference has no such split: its forward is one
 * call, model/fastspeech2.py:73-149.) */
int fs2_set_eager_stage2(fs2_ctx* ctx, int on);
/* event: a recorded cudaEvent_t (or NULL) the eager stage-2 body of the NEXT fs2_forward_stage1 waits for on its stream before it
 * overwrites the frame-side rows -- the completion of an asynchronous fs2_read_packed_postnet of the previous forward.  Consumed
 * (reset to NULL) by that call; the event must stay alive until it returns. */
int fs2_set_stage2_wait_event(fs2_ctx* ctx, void* event);

/* --- introspection for tests -------------------------------------------------------- */
/* When enabled, intermediate activations are copied aside after each stage of the forward
 * ("enc_in", "enc_0".."enc_3", "cond_x", "va_x", "lr_out", "dec_in", "dec_0".."dec_5", ...)
 * in the packed row layout, together with "p_start"/"f_start" (int32 row of each utterance). */
int fs2_debug_enable(fs2_ctx* ctx, int on);
int fs2_debug_fetch(fs2_ctx* ctx, const char* name, void* host_dst, int64_t max_bytes,
                    int64_t* rows, int64_t* cols);

/* Bring-up switches (0 = attention kernel raw-dump mode, 2 = GEMM cluster size, 3 = A-resident GEMM variant on/off,
 * 4 = fused FFN kernel: 0 off, 1 on, 2 automatic (default: on when the row tiles fill the SMs at least four times),
 * 5 = N-split fused-LayerNorm GEMMs for small row counts on/off, 6 = cta_group::2 MMAs on/off,
 * 7 = K-split clusters for single-row-tile contractions on/off, 8 = paired (K/V-multicast) attention kernel:
 * -1 automatic, 0 never, 1 always, 9 = stage-1 fusions (bit 0: duration + pitch predictors share their launches, bit 1:
 * conditioning add inside the last encoder LayerNorm epilogue; default 3), 10 = attention kernel form: 3 persistent with
 * softmax and accumulate warpgroups (default), 2 persistent with one group of row threads, 0 one CTA per work item). */
int fs2_debug_set_flag(int which, int value);
/* which = 1: CTA 0 of the next fs2_op_conv_gemm writes globaltimer stamps; read them back here. */
int fs2_debug_read_trace(int64_t* host_dst, int n);

/* Per-kernel-class timing with CUDA events on the launching stream.  After a forward run with
 * profiling enabled, fs2_profile_read writes lines "label launches total_ms\n" (NUL-terminated). */
int fs2_profile_enable(fs2_ctx* ctx, int on);
int fs2_profile_read(fs2_ctx* ctx, char* buf, int64_t buf_bytes);

/* --- single operators (unit tests, ncu) ------------------------------------------------ */
/* C[r,n] = act( sum_{t<taps} sum_{k<K} A[r+t-pad, k] * W[t][n][k] + bias[n] ) (+ residual[r,n]),
 * rows outside [0,rows) read as zero; then rows with row_vpos[r] >= min(extra,row_room[r]) are
 * zeroed when row_vpos != NULL.  act: 0 none, 1 relu, 2 tanh.  This is every Conv1d / Linear of
 * the path in token-major layout (SubLayers.py:39-41,54,87-88; modules.py:243-247;
 * fastspeech2.py:134; Layers.py:129-137).  math_mode FS2_MATH_TF32: W holds TF32-rounded values;
 * FS2_MATH_TF32X3: W is the plain fp32 weight and both operands are split (hi + lo) inside the call. */
int fs2_op_conv_gemm(fs2_stream stream, int math_mode, const float* A, int lda, int rows,
                     const float* W, const float* bias, int taps, int pad, int K, int N, int act,
                     const float* residual, int ldr, const int32_t* row_vpos, const int32_t* row_room,
                     int extra, float* C, int ldc);
/* The same contraction with the fused post-LayerNorm epilogue
 * (N = 256): y = LayerNorm(act(conv + bias) + residual) * gamma + beta, masked rows -> 0
 * (SubLayers.py:54-55,87-91; modules.py:243-247), optional head dot[r] = y[r,:].head_w + head_b
 * (modules.py:245-246).  C may be NULL when only `head_out` ([rows]) is wanted. */
int fs2_op_conv_gemm_ln(fs2_stream stream, const float* A, int lda, int rows, const float* W,
                        const float* bias, int taps, int pad, int K, int act, const float* residual, int ldr,
                        const float* gamma, const float* beta, const int32_t* row_vpos, const int32_t* row_room,
                        int extra, float* C, int ldc, const float* head_w, const float* head_b, float* head_out);
/* The whole position-wise FFN of an FFT block (transformer/SubLayers.py:85-93) in one kernel:
 * y = LayerNorm(w2 . ReLU(conv9(x) + b1) + b2 + x) * gamma + beta, masked rows -> 0.  x, y [rows,256] (y must not
 * alias x); w1 [9][1024][256] and w2 [256][1024] K-major, TF32 operands.  The 1024-wide hidden rows stay in tensor
 * memory (the second contraction takes its A operand from TMEM).  The forward uses it for large batches and the
 * two-launch form (conv9, then w2 + LayerNorm) otherwise; fs2_debug_set_flag(4, ...) / FS2_FFN_FUSED force either. */
int fs2_op_ffn_fused(fs2_stream stream, const float* x, int rows, const float* w1, const float* b1, const float* w2,
                     const float* b2, const float* gamma, const float* beta, const int32_t* row_vpos,
                     const int32_t* row_room, int extra, float* y);
/* The vocoder's form of the contraction (TF32): dilated taps (tap t reads row
 * r + (t - (taps-1)/2) * dil; hifigan/models.py:27-55), leaky ReLU (act = 3, `slope`) before and/or after (`act2`) the
 * residual add, a residual buffer that holds lrelu(x) and is inverted on the fly (`res_inv_lrelu`), and a row mask
 * looked up at row >> mask_shift (rows of an upsampled stage share the mask of their mel frame).  Small-K multi-tap
 * shapes run in the A-resident variant (activation tile + halo loaded once, one shifted descriptor per tap);
 * fs2_debug_set_flag(3, 0) forces the streaming variant for cross-checks. */
int fs2_op_conv_gemm_ex(fs2_stream stream, const float* A, int lda, int rows, const float* W, const float* bias, int taps,
                        int dil, int K, int N, int act, float slope, const float* residual, int ldr, int res_inv_lrelu,
                        int act2, const int32_t* row_vpos, const int32_t* row_room, int extra, int mask_shift, float* C,
                        int ldc);
/* FS2_MATH_BF16 form of the two contractions above: A [rows,lda] and W [taps][N][K] are bf16 (kind::f16 MMAs,
 * fp32 accumulation); bias / residual / gamma / beta are fp32.  gamma != NULL selects the fused LayerNorm
 * epilogue (N = 256).  The result is written as fp32 to C and/or as bf16 to C2 (either may be NULL): C2 is the
 * A operand of the next contraction of the forward. */
int fs2_op_conv_gemm_bf16(fs2_stream stream, const void* A, int lda, int rows, const void* W, const float* bias,
                          int taps, int pad, int K, int N, int act, const float* residual, int ldr,
                          const float* gamma, const float* beta, const int32_t* row_vpos, const int32_t* row_room,
                          int extra, float* C, int ldc, void* C2, int ldc2);
/* Varlen 2-head self-attention over packed rows (SubLayers.py:42-52, Modules.py:14-25):
 * qkv [rows,768] = [q | k | v], heads are 128-wide halves; utterance b owns rows
 * [starts[b], starts[b]+lens[b]); rows = rows of the qkv buffer.  out [rows,256].  The call builds the
 * longest-utterance-first work list the forward keeps per batch (one entry per 128 queries). */
int fs2_op_attention(fs2_stream stream, const float* qkv, int rows, const int32_t* starts,
                     const int32_t* lens, int batch, int max_len, float* out);
/* The FS2_MATH_BF16 form: qkv is bf16 [rows,768]; kind::f16 MMAs on 128-key tiles, probabilities rounded to bf16 (packed
 * in tensor memory), fp32 softmax statistics and output accumulation.  out [rows,256] fp32. */
int fs2_op_attention_bf16(fs2_stream stream, const void* qkv, int rows, const int32_t* starts, const int32_t* lens,
                          int batch, int max_len, float* out);
/* Durations -> repeat counts -> per-utterance inclusive scan (modules.py:132-135,186-187).
 * d_in: log-durations (is_target=0) or target durations (is_target=1), [B,L].  Writes
 * d_rounded [B,L] (fp32), cum [B,L] (int32 inclusive prefix sums of the repeat counts) and
 * mel_lens [B] (int64). */
int fs2_op_durations(fs2_stream stream, const float* d_in, int is_target, float d_control,
                     const int64_t* src_lens, int batch, int max_src_len,
                     float* d_rounded, int32_t* cum, int64_t* mel_lens);
/* bucketize(v, bins[255], right=False) (modules.py:83,87,94,98): idx[i] = #{bins < v[i]}. */
int fs2_op_bucketize(fs2_stream stream, const float* values, int64_t n, const float* bins, int n_bins,
                     int32_t* idx);
/* LengthRegulator index map (modules.py:182-190): for utterance b and frame t < cum[b,L-1],
 * map[b,t] = first j with cum[b,j] > t; -1 beyond.  map is [B, max_mel_len]. */
int fs2_op_frame_map(fs2_stream stream, const int32_t* cum, int batch, int max_src_len,
                     int max_mel_len, int32_t* map);

/* ======================================================================================
 * HiFi-GAN generator: the vocoder step that follows FastSpeech2.forward in the reference's
 * synthesis scripts (utils/model.py:37-71 `get_vocoder`, :74-92 `vocoder_infer`;
 * hifigan/models.py:112-174 `Generator`; architecture hifigan/config.json, V1).
 * Same conventions as above; one context per device, not thread-safe. */
typedef struct fs2_voc fs2_voc;
/* math_mode: FS2_MATH_TF32 (fp32 activations, TF32 operands) or FS2_MATH_BF16 (bf16 weights and activations -- half the
 * HBM traffic of the audio-rate stages; fp32 accumulation, biases and output waveform). */
int fs2_voc_create(int device, int math_mode, fs2_voc** out);
void fs2_voc_destroy(fs2_voc* voc);
const char* fs2_voc_last_error(const fs2_voc* voc);
/* `key` is a key of Generator.state_dict() after remove_weight_norm() (utils/model.py:66): conv_pre.*, ups.{0..3}.*,
 * resblocks.{0..11}.convs{1,2}.{0..2}.*, conv_post.* (weight, bias); fp32 on the context's device; copied. */
int fs2_voc_set_weight(fs2_voc* voc, const char* key, const void* dev_ptr, const int64_t* shape, int ndim);
/* Repack: Conv1d [Cout,Cin,k] -> [k][Cout][Cin]; ConvTranspose1d [Cin,Cout,2s] -> 3-tap GEMM form [3][s*Cout][Cin]. */
int fs2_voc_prepare(fs2_voc* voc, fs2_stream stream);
/* Generator.forward (hifigan/models.py:148-167): mel [batch, 80, n_frames] (element strides given, so the
 * [B, T, 80] postnet output of fs2_forward_stage2 can be passed transposed without a copy) ->
 * wav [batch, n_frames * 256] fp32 in [-1, 1].  mel_lens == NULL: every frame of the padded batch is data, exactly
 * as `vocoder(mels)` treats it.  mel_lens != NULL (device int64 [batch]): frames beyond mel_lens[b] are skipped,
 * utterance b is synthesised as if it were alone (zero padding at both ends) and wav[b, 256*mel_lens[b]:] = 0 --
 * what `vocoder_infer(..., lengths=)` keeps after trimming, up to the receptive field at the tail. */
int fs2_voc_forward(fs2_voc* voc, fs2_stream stream, const float* mel, int64_t stride_b, int64_t stride_c,
                    int64_t stride_t, int batch, int n_frames, const int64_t* mel_lens, float* wav);
int fs2_voc_last_launch_count(const fs2_voc* voc);

#ifdef __cplusplus
}
#endif
#endif /* FS2_B200_H */
