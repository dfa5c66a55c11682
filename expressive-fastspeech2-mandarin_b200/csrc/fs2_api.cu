// libfs2b200.so -- C ABI (include/fs2_b200.h) and the orchestration of the forward pass.
// One context per device; all work is enqueued on the caller's stream.
#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdlib>
#include <cmath>
#include <cstring>

#include "attention_tc.cuh"
#include "attention_bf16.cuh"
#include "common.cuh"
#include "gemm_tc2.cuh"
#include "attention_tc2.cuh"
#include "attention_tcp.cuh"
#include "attention_tcq.cuh"
#include "ffn_fused.cuh"
#include "rowops.cuh"
#include "vocoder.cuh"

namespace fs2 {

thread_local int g_launches = 0;
static thread_local std::string g_create_error;

struct DevTensor {
  float* ptr = nullptr;
  std::vector<int64_t> shape;
  int64_t numel = 0;
};

struct FFTLayer {
  float *wqkv, *bqkv, *wfc, *bfc, *ln1_g, *ln1_b, *w1, *b1, *w2, *b2, *ln2_g, *ln2_b;
};
struct Predictor {
  float *w1, *b1, *ln1_g, *ln1_b, *w2, *b2, *ln2_g, *ln2_b, *head_w, *head_b;
};
struct PostConv {
  float *w, *b;
  int cin, cout;
};

struct RowSide {  // metadata of one packed row space (phoneme side or frame side)
  int32_t *starts = nullptr, *lens = nullptr;       // [B+1], [B]
  int32_t *utt = nullptr, *vpos = nullptr, *room = nullptr, *slot = nullptr;  // [rows_alloc]
  int64_t* totals = nullptr;                         // device [3]
  uint32_t* work = nullptr;                          // attention work list (rowops.cuh: build_attention_work)
  int32_t* work_count = nullptr;                     // device [1]
  int work_alloc = 0, work_cap = 0;                  // allocated entries / the bound the current batch launches with
  int work_q_rows = 128;                             // query rows per entry: 128, or 256 (paired attention kernel)
  int rows_alloc = 0, batch_alloc = 0;
  RowMeta meta() const { return RowMeta{utt, vpos, room}; }
};

struct Pool {  // activation buffers of one side
  float* act[4] = {nullptr, nullptr, nullptr, nullptr};  // [rows,256]
  float* qkv = nullptr;                                  // [rows,768]
  float* hid = nullptr;                                  // [rows,1024]
  float *mel = nullptr, *post = nullptr;                 // [rows,80]   (frame side only)
  float* pn[2] = {nullptr, nullptr};                     // [rows,512]  (frame side only)
  // FS2_MATH_BF16: bf16 copies that serve as the A operands (the fp32 hid / pn buffers are not allocated)
  __nv_bfloat16* actb[4] = {nullptr, nullptr, nullptr, nullptr};  // mirrors of act[i]
  __nv_bfloat16* qkvb = nullptr;                                   // [rows,768]: Q | K | V as the attention's bf16 operands
  __nv_bfloat16* hidb = nullptr;                                   // [rows,1024]
  __nv_bfloat16* melb = nullptr;                                   // [rows,80]
  __nv_bfloat16* pnb[2] = {nullptr, nullptr};                      // [rows,512]
  int rows = 0;
};

}  // namespace fs2

using namespace fs2;

struct fs2_ctx {
  int device = 0;
  fs2_config cfg{};
  std::string err;
  bool prepared = false;
  bool debug = false;
  std::map<std::string, DevTensor> raw;
  std::vector<void*> owned;  // repacked weights, freed in destroy

  FFTLayer enc[ENC_LAYERS], dec[DEC_LAYERS];
  Predictor pred[3];  // duration, pitch, energy
  PostConv post[PN_LAYERS];
  float *mel_w = nullptr, *mel_b = nullptr;
  float* pe_long = nullptr;  // generated sinusoid table for sequences beyond max_seq_len
  int pe_long_rows = 0;
  float* splitk_ws = nullptr;  // K-split workspace (tc2::SPLITK_WS_BYTES): partial tiles of single-utterance launches
  int32_t* ffn_flags = nullptr;  // fused FFN hand-over flags (grow-only, zero-filled when grown) and the launch counter
  size_t ffn_flags_cap = 0;
  int32_t ffn_epoch = 0;
  float* split_buf = nullptr;  // FS2_MATH_TF32X3: [rows, hi | lo] copy of the activations of the contraction being launched
  size_t split_cap = 0;

  RowSide ps, fs;
  Pool pp, fp;
  int32_t* status = nullptr;
  int32_t* cum = nullptr;       // [B, Lmax]
  int32_t* mel_lens32 = nullptr;
  float *raw_pitch = nullptr, *raw_energy = nullptr;  // [B, Lmax]
  float *cond_spk = nullptr, *cond_emo = nullptr;     // [B,256]
  int64_t scratch_bl = 0, scratch_b = 0;
  int64_t* h_totals = nullptr;  // pinned [8]

  // state carried from stage 1 to stage 2
  bool stage1_done = false;
  bool eager_stage2 = false;        // fs2_set_eager_stage2: stage 1 enqueues stage 2 up to the PostNet before it returns
  bool stage2_body_done = false;    // ... and did so for the current forward: fs2_forward_stage2 only unpacks
  cudaEvent_t stage2_wait = nullptr;   // fs2_set_stage2_wait_event: the eager stage-2 body of the next forward waits for it
  int batch = 0, max_src_len = 0, max_mel_len = 0, phon_rows = 0;
  int64_t frame_rows = 0, total_frames = 0;
  const float* lr_input = nullptr;
  float* va_spare = nullptr;
  bool lr_adds_energy = true;    // false after fs2_import_stage1: the imported rows already carry the energy embedding
  const float *p_targets = nullptr, *e_targets = nullptr;  // frame_level teacher forcing: consumed in stage 2
  float p_control = 1.f;
  int64_t scratch_bt = 0;                                   // frame_level: raw predictions [B, T_max]
  float *raw_pitch_f = nullptr, *raw_energy_f = nullptr;
  int last_launches = 0;

  std::map<std::string, std::pair<void*, std::vector<int64_t>>> taps;  // name -> (device copy, {rows, cols, elt})

  // optional per-kernel-class CUDA-event timing (bench.py's roofline numbers come from here)
  bool profiling = false;
  struct ProfRec { const char* label; cudaEvent_t beg, end; };
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> event_pool;
  std::string prof_text;
};

namespace fs2 {

// Brackets the launches of one kernel class with CUDA events on the launching stream.
struct ProfScope {
  fs2_ctx* c;
  cudaStream_t s;
  cudaEvent_t end = nullptr;
  ProfScope(fs2_ctx* ctx, cudaStream_t stream, const char* label) : c(ctx), s(stream) {
    if (!c->profiling) return;
    cudaEvent_t ev[2];
    for (auto& e : ev) {
      if (c->event_pool.empty()) {
        FS2_CUDA_OK(cudaEventCreate(&e));
      } else {
        e = c->event_pool.back();
        c->event_pool.pop_back();
      }
    }
    FS2_CUDA_OK(cudaEventRecord(ev[0], s));
    end = ev[1];
    c->prof.push_back({label, ev[0], ev[1]});
  }
  ~ProfScope() {
    if (end) cudaEventRecord(end, s);
  }
};

template <typename T>
static T* dalloc(size_t n) {
  void* p = nullptr;
  FS2_CUDA_OK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  return static_cast<T*>(p);
}
// Replace a workspace buffer by a larger, zero-filled one.  cudaFree waits for the device, so no kernel still reads the
// old buffer; the fill is enqueued on the caller's stream, i.e. ordered before every kernel the forward enqueues next
// (the legacy default stream is not ordered against a caller's non-blocking stream).
template <typename T>
static void regrow(T*& p, size_t n, cudaStream_t s) {
  if (p) FS2_CUDA_OK(cudaFree(p));
  p = dalloc<T>(n);
  FS2_CUDA_OK(cudaMemsetAsync(p, 0, std::max<size_t>(n, 1) * sizeof(T), s));
}

static const DevTensor& W(fs2_ctx* c, const std::string& key, std::initializer_list<int64_t> shape) {
  auto it = c->raw.find(key);
  require(it != c->raw.end(), FS2_ERR_INVALID, "missing weight: " + key);
  require(it->second.shape == std::vector<int64_t>(shape), FS2_ERR_INVALID, "unexpected shape for weight: " + key);
  return it->second;
}

static float* keep(fs2_ctx* c, size_t n) {
  float* p = dalloc<float>(n);
  c->owned.push_back(p);
  return p;
}

// [Cout][Cin][k] -> [k][Cout][Cin], optional per-Cout scale, TF32-rounded operands
// In FS2_MATH_BF16 the result is a bf16 array behind the float* (ConvGemmArgs::a_bf16 tells the kernel).
static void repack_into(fs2_ctx* c, const float* w, int cout, int cin, int k, const float* scale, float* out, cudaStream_t s,
                        int64_t lo_off = 0) {
  const int64_t n = (int64_t)cout * cin * k;
  if (c->cfg.math_mode == FS2_MATH_BF16) {
    repack_conv_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(w, cout, cin, k, scale,
                                                                        reinterpret_cast<__nv_bfloat16*>(out));
  } else if (c->cfg.math_mode == FS2_MATH_TF32X3) {   // [hi block ; lo block]
    repack_conv_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(w, cout, cin, k, scale, out, lo_off > 0 ? lo_off : n);
  } else {
    repack_conv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(w, cout, cin, k, scale, 1, out);
  }
  FS2_LAUNCHED();
}
static float* repack_conv(fs2_ctx* c, const float* w, int cout, int cin, int k, const float* scale, cudaStream_t s) {
  float* out = keep(c, (size_t)cout * cin * k * (c->cfg.math_mode == FS2_MATH_TF32X3 ? 2 : 1));
  repack_into(c, w, cout, cin, k, scale, out, s);
  return out;
}

static void prepare_fft(fs2_ctx* c, const std::string& p, FFTLayer& L, cudaStream_t s) {
  const int d = D_MODEL;
  const bool x3 = c->cfg.math_mode == FS2_MATH_TF32X3;
  L.wqkv = keep(c, 3 * d * d * (x3 ? 2 : 1));   // split operands: [Q K V hi ; Q K V lo]
  L.bqkv = keep(c, 3 * d);
  const char* names[3] = {"w_qs", "w_ks", "w_vs"};
  for (int i = 0; i < 3; ++i) {
    const auto& w = W(c, p + ".slf_attn." + names[i] + ".weight", {d, d});
    const auto& b = W(c, p + ".slf_attn." + names[i] + ".bias", {d});
    // Linear weight [N][K] is already K-major: repack as a 1-tap conv to round the operands
    const size_t off = (size_t)i * d * d;   // in elements of the operand type
    repack_into(c, w.ptr, d, d, 1, nullptr,
                c->cfg.math_mode == FS2_MATH_BF16 ? reinterpret_cast<float*>(reinterpret_cast<__nv_bfloat16*>(L.wqkv) + off)
                                                  : L.wqkv + off, s, 3 * d * d);
    FS2_CUDA_OK(cudaMemcpyAsync(L.bqkv + i * d, b.ptr, d * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  L.wfc = repack_conv(c, W(c, p + ".slf_attn.fc.weight", {d, d}).ptr, d, d, 1, nullptr, s);
  L.bfc = W(c, p + ".slf_attn.fc.bias", {d}).ptr;
  L.ln1_g = W(c, p + ".slf_attn.layer_norm.weight", {d}).ptr;
  L.ln1_b = W(c, p + ".slf_attn.layer_norm.bias", {d}).ptr;
  L.w1 = repack_conv(c, W(c, p + ".pos_ffn.w_1.weight", {D_INNER, d, FFN_TAPS}).ptr, D_INNER, d, FFN_TAPS, nullptr, s);
  L.b1 = W(c, p + ".pos_ffn.w_1.bias", {D_INNER}).ptr;
  L.w2 = repack_conv(c, W(c, p + ".pos_ffn.w_2.weight", {d, D_INNER, 1}).ptr, d, D_INNER, 1, nullptr, s);
  L.b2 = W(c, p + ".pos_ffn.w_2.bias", {d}).ptr;
  L.ln2_g = W(c, p + ".pos_ffn.layer_norm.weight", {d}).ptr;
  L.ln2_b = W(c, p + ".pos_ffn.layer_norm.bias", {d}).ptr;
}

static void prepare_predictor(fs2_ctx* c, const std::string& p, Predictor& P, cudaStream_t s) {
  const int d = D_MODEL;
  P.w1 = repack_conv(c, W(c, p + ".conv_layer.conv1d_1.conv.weight", {d, d, VP_TAPS}).ptr, d, d, VP_TAPS, nullptr, s);
  P.b1 = W(c, p + ".conv_layer.conv1d_1.conv.bias", {d}).ptr;
  P.ln1_g = W(c, p + ".conv_layer.layer_norm_1.weight", {d}).ptr;
  P.ln1_b = W(c, p + ".conv_layer.layer_norm_1.bias", {d}).ptr;
  P.w2 = repack_conv(c, W(c, p + ".conv_layer.conv1d_2.conv.weight", {d, d, VP_TAPS}).ptr, d, d, VP_TAPS, nullptr, s);
  P.b2 = W(c, p + ".conv_layer.conv1d_2.conv.bias", {d}).ptr;
  P.ln2_g = W(c, p + ".conv_layer.layer_norm_2.weight", {d}).ptr;
  P.ln2_b = W(c, p + ".conv_layer.layer_norm_2.bias", {d}).ptr;
  P.head_w = W(c, p + ".linear_layer.weight", {1, d}).ptr;
  P.head_b = W(c, p + ".linear_layer.bias", {1}).ptr;
}

__global__ void sinusoid_kernel(float* pe, int rows) {
  // transformer/Models.py:10-30 evaluated in float64, stored as fp32
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * D_MODEL) return;
  const int pos = (int)(i / D_MODEL), j = (int)(i % D_MODEL);
  const double ang = (double)pos / pow(10000.0, 2.0 * (double)(j / 2) / (double)D_MODEL);
  pe[i] = (float)((j & 1) ? cos(ang) : sin(ang));
}

static void tap(fs2_ctx* c, cudaStream_t s, const char* name, const void* ptr, int64_t rows, int64_t cols, int64_t elt);

static const float* position_rows(fs2_ctx* c, const char* key, int n_rows, cudaStream_t s) {
  // Stored table for n_rows <= max_seq_len, regenerated formula beyond (Models.py:82-91,145-162)
  if (n_rows <= c->cfg.max_seq_len) return c->raw.at(key).ptr;
  if (c->pe_long_rows < n_rows) {
    const int rows = round_up(n_rows, 1024);
    FS2_CUDA_OK(cudaStreamSynchronize(s));
    regrow(c->pe_long, (size_t)rows * D_MODEL, s);
    sinusoid_kernel<<<(unsigned)(((int64_t)rows * D_MODEL + 255) / 256), 256, 0, s>>>(c->pe_long, rows);
    FS2_LAUNCHED();
    c->pe_long_rows = rows;
  }
  if (c->debug) tap(c, s, "pe_long", c->pe_long, c->pe_long_rows, D_MODEL, 4);
  return c->pe_long;
}

static void tap(fs2_ctx* c, cudaStream_t s, const char* name, const void* ptr, int64_t rows, int64_t cols, int64_t elt = 4);
static void tap(fs2_ctx* c, cudaStream_t s, const char* name, const void* ptr, int64_t rows, int64_t cols, int64_t elt) {
  if (!c->debug) return;
  auto& slot = c->taps[name];
  if (slot.first) cudaFree(slot.first);
  void* p = nullptr;
  FS2_CUDA_OK(cudaMalloc(&p, std::max<int64_t>(rows * cols * elt, 1)));
  FS2_CUDA_OK(cudaMemcpyAsync(p, ptr, rows * cols * elt, cudaMemcpyDeviceToDevice, s));
  slot = {p, {rows, cols, elt}};
}

// ---------------------------------------------------------------------------- GEMM entry
// FS2_MATH_TF32X3: the activations are split into [rows, hi | lo] (both exactly representable in TF32) in the context's
// scratch buffer and the contraction runs three terms against the [hi ; lo] weight blocks prepared by fs2_prepare.
static void conv_gemm(fs2_ctx* c, ConvGemmArgs a, cudaStream_t s, const ConvGemmArgs* second = nullptr) {
  if (second != nullptr) {   // two contractions of one shape in one launch (never in the split-operand mode)
    a.splitk_ws = c->splitk_ws;
    tc2::launch(a, s, second);
    return;
  }
  if (c->cfg.math_mode == FS2_MATH_TF32X3 && a.rows > 0) {
    const size_t need = (size_t)a.rows * 2 * a.K;
    if (need > c->split_cap) {
      FS2_CUDA_OK(cudaStreamSynchronize(s));
      regrow(c->split_buf, need, s);
      c->split_cap = need;
    }
    const int64_t n4 = (int64_t)a.rows * (a.K / 4);
    split_tf32_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, s>>>(a.A, a.rows, a.K, a.lda, c->split_buf);
    FS2_LAUNCHED();
    a.A = c->split_buf;
    a.lda = 2 * a.K;
    a.terms = 3;
  }
  a.splitk_ws = c->splitk_ws;
  tc2::launch(a, s);
}

static void attention(const float* qkv, int rows, const RowSide& side, int batch, int max_len, float* out, cudaStream_t s,
                      void* out_bf16 = nullptr) {
  (void)batch; (void)max_len;
  if (attn_tc::use_two_sm(side.work_q_rows))
    attn2::launch(qkv, rows, side.starts, side.lens, side.work, side.work_count, side.work_cap, out, s, out_bf16);
  else if (side.work_q_rows == attn_tc::BQ && attn_tc::debug_flag() == 0 && out_bf16 == nullptr &&
           attn_p::use_persistent(side.work_cap, tc2::sm_count())) {
    if (attn_p::enabled_flag() == 3)
      attn_q::launch(qkv, rows, side.starts, side.lens, side.work, side.work_count, side.work_cap, out, s, tc2::sm_count());
    else
      attn_p::launch(qkv, rows, side.starts, side.lens, side.work, side.work_count, side.work_cap, out, s, tc2::sm_count());
  }
  else
    attn_tc::launch(qkv, rows, side.starts, side.lens, side.work, side.work_count, side.work_cap, side.work_q_rows, out, s, out_bf16);
}

// bf16-operand variant: A and W are bf16 behind the float* fields
static ConvGemmArgs gemm_args_b(const void* A, int lda, int rows, const float* Wt, const float* bias, int taps, int K, int N,
                                int act, float* C, int ldc, void* C2, int ldc2);

static ConvGemmArgs gemm_args(const float* A, int lda, int rows, const float* Wt, const float* bias, int taps, int K,
                              int N, int act, float* C, int ldc) {
  ConvGemmArgs a{};
  a.A = A; a.lda = lda; a.rows = rows; a.W = Wt; a.bias = bias; a.taps = taps; a.pad = (taps - 1) / 2;
  a.K = K; a.N = N; a.act = act; a.C = C; a.ldc = ldc;
  return a;
}

static ConvGemmArgs gemm_args_b(const void* A, int lda, int rows, const float* Wt, const float* bias, int taps, int K, int N,
                                int act, float* C, int ldc, void* C2, int ldc2) {
  ConvGemmArgs a = gemm_args(static_cast<const float*>(A), lda, rows, Wt, bias, taps, K, N, act, C, ldc);
  a.a_bf16 = 1; a.C2 = C2; a.ldc2 = ldc2;
  return a;
}

// One FFT block in place on x (transformer/Layers.py:21-30).  t1, t2 are [rows,256] temporaries.
// Launch of a row kernel of the forward (rowops.cuh: the kernels that start with row_pdl_sync()) as a programmatic dependent
// launch: its blocks become resident under the previous kernel's tail and it lets the next kernel start its prologue.
// FS2_ROW_PDL=0 launches them the ordinary way (A/B).
template <typename... KArgs, typename... Args>
static void launch_row(void (*kernel)(KArgs...), unsigned grid, unsigned block, cudaStream_t s, Args&&... args) {
  static const bool pdl = [] { const char* e = std::getenv("FS2_ROW_PDL"); return e == nullptr || std::atoi(e) != 0; }();
  if (pdl) {
    tc::launch_pdl(kernel, dim3(grid), dim3(block), 0, s, 1, std::forward<Args>(args)...);
  } else {
    kernel<<<grid, block, 0, s>>>(std::forward<Args>(args)...);
    FS2_CUDA_OK(cudaGetLastError());
  }
}

// Stage-1 fusions (debug flag 9 / FS2_STAGE1_FUSION; default 3): bit 0 = the duration and pitch predictors share their
// launches, bit 1 = the conditioning add lives in the last encoder layer's LayerNorm epilogue.
static int& stage1_fusion_flag() {
  static int f = [] { const char* e = std::getenv("FS2_STAGE1_FUSION"); return e != nullptr ? std::atoi(e) : 3; }();
  return f;
}

// cond (last encoder layer only): the speaker / emotion vectors added to the block's output inside the FFN's LayerNorm
// epilogue (ConvGemmArgs::post_a / ffn::Args::post_a).
struct CondAdd {
  const float *spk, *emo;
};
static void fft_block(fs2_ctx* c, cudaStream_t s, const FFTLayer& L, const RowSide& side, Pool& pool, int rows, int batch,
                      int max_len, float* x, float* t1, float* t2, bool frame, const CondAdd* cond = nullptr) {
  const int math = c->cfg.math_mode;
  const int32_t* live = reinterpret_cast<const int32_t*>(side.totals);   // low word of totals[0] (little endian)
  if (math == FS2_MATH_BF16) {
    // Same five launches; every A operand is the bf16 mirror written by the producing epilogue (Q, K, V included: the
    // attention runs kind::f16 MMAs on 128-key tiles), the residual stream (x, t2) stays fp32.
    __nv_bfloat16 *xb = pool.actb[0], *t1b = pool.actb[1], *t2b = pool.actb[2];
    ConvGemmArgs a = gemm_args_b(xb, D_MODEL, rows, L.wqkv, L.bqkv, 1, D_MODEL, 3 * D_MODEL, ACT_NONE, nullptr, 0, pool.qkvb,
                                 3 * D_MODEL);
    a.live_rows = live;
    { ProfScope ps(c, s, frame ? "dec.gemm_qkv" : "enc.gemm_qkv"); conv_gemm(c, a, s); }
    { ProfScope ps(c, s, frame ? "dec.attention" : "enc.attention");
      attn_bf::launch(pool.qkvb, rows, side.starts, side.lens, side.work, side.work_count, side.work_cap, nullptr, t1b, s); }
    a = gemm_args_b(t1b, D_MODEL, rows, L.wfc, L.bfc, 1, D_MODEL, D_MODEL, ACT_NONE, t2, D_MODEL, t2b, D_MODEL);
    a.residual = x; a.ldr = D_MODEL; a.live_rows = live;
    a.ln_gamma = L.ln1_g; a.ln_beta = L.ln1_b; a.row_vpos = side.vpos; a.row_room = side.room; a.extra = 0;
    { ProfScope ps(c, s, frame ? "dec.gemm_fc_ln" : "enc.gemm_fc_ln"); conv_gemm(c, a, s); }
    a = gemm_args_b(t2b, D_MODEL, rows, L.w1, L.b1, FFN_TAPS, D_MODEL, D_INNER, ACT_RELU, nullptr, 0, pool.hidb, D_INNER);
    a.live_rows = live;
    { ProfScope ps(c, s, frame ? "dec.gemm_conv9" : "enc.gemm_conv9"); conv_gemm(c, a, s); }
    a = gemm_args_b(pool.hidb, D_INNER, rows, L.w2, L.b2, 1, D_INNER, D_MODEL, ACT_NONE, x, D_MODEL, xb, D_MODEL);
    a.residual = t2; a.ldr = D_MODEL; a.live_rows = live;
    a.ln_gamma = L.ln2_g; a.ln_beta = L.ln2_b; a.row_vpos = side.vpos; a.row_room = side.room; a.extra = 0;
    if (cond != nullptr) { a.post_a = cond->spk; a.post_b = cond->emo; a.post_utt = side.utt; a.post_extra = 2; }
    { ProfScope ps(c, s, frame ? "dec.gemm_w2_ln" : "enc.gemm_w2_ln"); conv_gemm(c, a, s); }
    return;
  }
  ConvGemmArgs a = gemm_args(x, D_MODEL, rows, L.wqkv, L.bqkv, 1, D_MODEL, 3 * D_MODEL, ACT_NONE, pool.qkv, 3 * D_MODEL);
  a.live_rows = live;
  { ProfScope ps(c, s, frame ? "dec.gemm_qkv" : "enc.gemm_qkv"); conv_gemm(c, a, s); }
  { ProfScope ps(c, s, frame ? "dec.attention" : "enc.attention");
    attention(pool.qkv, rows, side, batch, max_len, t1, s); }
  // fused: LayerNorm(fc(ctx) + x) with the row mask, then LayerNorm(w2(relu(conv9(.))) + .) (SubLayers.py:54-55,87-91)
  a = gemm_args(t1, D_MODEL, rows, L.wfc, L.bfc, 1, D_MODEL, D_MODEL, ACT_NONE, t2, D_MODEL);
  a.residual = x; a.ldr = D_MODEL; a.live_rows = live;
  a.ln_gamma = L.ln1_g; a.ln_beta = L.ln1_b; a.row_vpos = side.vpos; a.row_room = side.room; a.extra = 0;
  { ProfScope ps(c, s, frame ? "dec.gemm_fc_ln" : "enc.gemm_fc_ln"); conv_gemm(c, a, s); }
  if (math == FS2_MATH_TF32 && ffn::use_fused(rows)) {
    // conv9 -> ReLU -> w2 -> +residual -> LayerNorm -> mask in one kernel; the hidden tensor stays in tensor memory
    ffn::Args f{};
    f.x = t2; f.rows = rows; f.w1 = L.w1; f.b1 = L.b1; f.w2 = L.w2; f.b2 = L.b2; f.gamma = L.ln2_g; f.beta = L.ln2_b;
    f.row_vpos = side.vpos; f.row_room = side.room; f.extra = 0;
    f.live_rows = live; f.y = x;
    if (cond != nullptr) { f.post_a = cond->spk; f.post_b = cond->emo; f.post_utt = side.utt; f.post_extra = 2; }
    // a row-tile group split between two clusters is handed over through the (otherwise unused) hidden buffer
    if (ffn::flag_count(rows) > c->ffn_flags_cap) {
      FS2_CUDA_OK(cudaStreamSynchronize(s));
      cudaFree(c->ffn_flags);
      c->ffn_flags_cap = ffn::flag_count(rows) * 2;
      c->ffn_flags = dalloc<int32_t>(c->ffn_flags_cap);
      FS2_CUDA_OK(cudaMemsetAsync(c->ffn_flags, 0, c->ffn_flags_cap * sizeof(int32_t), s));
      c->ffn_epoch = 0;
    }
    if (c->ffn_epoch == INT32_MAX) {   // (6 launches per forward: days of continuous serving) restart the epochs behind a clear
      FS2_CUDA_OK(cudaMemsetAsync(c->ffn_flags, 0, c->ffn_flags_cap * sizeof(int32_t), s));
      c->ffn_epoch = 0;
    }
    f.partial = pool.hid; f.flags = c->ffn_flags; f.epoch = ++c->ffn_epoch;
    ProfScope ps(c, s, frame ? "dec.ffn_fused" : "enc.ffn_fused");
    ffn::launch(f, s);
    return;
  }
  a = gemm_args(t2, D_MODEL, rows, L.w1, L.b1, FFN_TAPS, D_MODEL, D_INNER, ACT_RELU, pool.hid, D_INNER);
  a.live_rows = live;
  { ProfScope ps(c, s, frame ? "dec.gemm_conv9" : "enc.gemm_conv9"); conv_gemm(c, a, s); }
  a = gemm_args(pool.hid, D_INNER, rows, L.w2, L.b2, 1, D_INNER, D_MODEL, ACT_NONE, x, D_MODEL);
  a.residual = t2; a.ldr = D_MODEL; a.live_rows = live;
  a.ln_gamma = L.ln2_g; a.ln_beta = L.ln2_b; a.row_vpos = side.vpos; a.row_room = side.room; a.extra = 0;
  if (cond != nullptr) { a.post_a = cond->spk; a.post_b = cond->emo; a.post_utt = side.utt; a.post_extra = 2; }
  { ProfScope ps(c, s, frame ? "dec.gemm_w2_ln" : "enc.gemm_w2_ln"); conv_gemm(c, a, s); }
}

// VariancePredictor (model/modules.py:242-250) over the packed rows; head_out is [B, Lmax], pre-zeroed.
static void predictor(fs2_ctx* c, cudaStream_t s, const Predictor& P, const RowSide& side, int rows, const float* x,
                      float* t1, float* head_out, const __nv_bfloat16* xb = nullptr, __nv_bfloat16* t1b = nullptr) {
  ProfScope ps(c, s, "predictor");
  const int32_t* live = reinterpret_cast<const int32_t*>(side.totals);
  const bool bf = c->cfg.math_mode == FS2_MATH_BF16;
  ConvGemmArgs f = bf ? gemm_args_b(xb, D_MODEL, rows, P.w1, P.b1, VP_TAPS, D_MODEL, D_MODEL, ACT_RELU, nullptr, 0, t1b, D_MODEL)
                      : gemm_args(x, D_MODEL, rows, P.w1, P.b1, VP_TAPS, D_MODEL, D_MODEL, ACT_RELU, t1, D_MODEL);
  f.live_rows = live;
  // the hidden row at t = L_b is live when L_b < L_max (the reference's second conv reads it)
  f.ln_gamma = P.ln1_g; f.ln_beta = P.ln1_b; f.row_vpos = side.vpos; f.row_room = side.room; f.extra = 1;
  conv_gemm(c, f, s);
  f = bf ? gemm_args_b(t1b, D_MODEL, rows, P.w2, P.b2, VP_TAPS, D_MODEL, D_MODEL, ACT_RELU, nullptr, 0, nullptr, 0)
         : gemm_args(t1, D_MODEL, rows, P.w2, P.b2, VP_TAPS, D_MODEL, D_MODEL, ACT_RELU, nullptr, D_MODEL);
  f.live_rows = live;
  f.ln_gamma = P.ln2_g; f.ln_beta = P.ln2_b; f.row_vpos = side.vpos; f.row_room = side.room; f.extra = 0;
  f.head_w = P.head_w; f.head_b = P.head_b; f.head_out = head_out; f.slot = side.slot;
  conv_gemm(c, f, s);
}

// Two predictors that read the SAME rows (duration and pitch, model/modules.py:115-121): each of the two conv + ReLU +
// LayerNorm layers runs both predictors in one launch (grid.y = 2).  h0 / h1 are the two hidden buffers.
static void predictor_pair(fs2_ctx* c, cudaStream_t s, const Predictor& P0, const Predictor& P1, const RowSide& side, int rows,
                           const float* x, float* h0, float* h1, float* head0, float* head1, const __nv_bfloat16* xb,
                           __nv_bfloat16* h0b, __nv_bfloat16* h1b) {
  ProfScope ps(c, s, "predictor");
  const int32_t* live = reinterpret_cast<const int32_t*>(side.totals);
  const bool bf = c->cfg.math_mode == FS2_MATH_BF16;
  auto layer1 = [&](const Predictor& P, float* h, __nv_bfloat16* hb) {
    ConvGemmArgs f = bf ? gemm_args_b(xb, D_MODEL, rows, P.w1, P.b1, VP_TAPS, D_MODEL, D_MODEL, ACT_RELU, nullptr, 0, hb, D_MODEL)
                        : gemm_args(x, D_MODEL, rows, P.w1, P.b1, VP_TAPS, D_MODEL, D_MODEL, ACT_RELU, h, D_MODEL);
    f.live_rows = live;
    f.ln_gamma = P.ln1_g; f.ln_beta = P.ln1_b; f.row_vpos = side.vpos; f.row_room = side.room; f.extra = 1;
    return f;
  };
  auto layer2 = [&](const Predictor& P, const float* h, const __nv_bfloat16* hb, float* head_out) {
    ConvGemmArgs f = bf ? gemm_args_b(hb, D_MODEL, rows, P.w2, P.b2, VP_TAPS, D_MODEL, D_MODEL, ACT_RELU, nullptr, 0, nullptr, 0)
                        : gemm_args(h, D_MODEL, rows, P.w2, P.b2, VP_TAPS, D_MODEL, D_MODEL, ACT_RELU, nullptr, D_MODEL);
    f.live_rows = live;
    f.ln_gamma = P.ln2_g; f.ln_beta = P.ln2_b; f.row_vpos = side.vpos; f.row_room = side.room; f.extra = 0;
    f.head_w = P.head_w; f.head_b = P.head_b; f.head_out = head_out; f.slot = side.slot;
    return f;
  };
  ConvGemmArgs a0 = layer1(P0, h0, h0b), a1 = layer1(P1, h1, h1b);
  conv_gemm(c, a0, s, &a1);
  a0 = layer2(P0, h0, h0b, head0);
  a1 = layer2(P1, h1, h1b, head1);
  conv_gemm(c, a0, s, &a1);
}

static void ensure_side(RowSide& sd, int batch, int rows, cudaStream_t s, int work_cap = 0) {
  if (batch + 1 > sd.batch_alloc) {
    regrow(sd.starts, batch + 1, s);
    regrow(sd.lens, batch, s);
    sd.batch_alloc = batch + 1;
  }
  if (rows > sd.rows_alloc) {
    regrow(sd.utt, rows, s);
    regrow(sd.vpos, rows, s);
    regrow(sd.room, rows, s);
    regrow(sd.slot, rows, s);
    sd.rows_alloc = rows;
  }
  if (work_cap > sd.work_alloc) {
    regrow(sd.work, work_cap, s);
    sd.work_alloc = work_cap;
  }
  if (work_cap > 0) sd.work_cap = work_cap;
  if (!sd.totals) regrow(sd.totals, 3, s);
  if (!sd.work_count) regrow(sd.work_count, 1, s);
}

static void ensure_pool(Pool& p, int rows, bool frame_side, bool bf16, cudaStream_t s) {
  if (rows <= p.rows) return;
  for (auto& a : p.act) regrow(a, (size_t)rows * D_MODEL, s);
  regrow(p.qkv, (size_t)rows * 3 * D_MODEL, s);
  if (bf16) {
    for (auto& a : p.actb) regrow(a, (size_t)rows * D_MODEL, s);
    regrow(p.qkvb, (size_t)rows * 3 * D_MODEL, s);
    regrow(p.hidb, (size_t)rows * D_INNER, s);
  } else {
    regrow(p.hid, (size_t)rows * D_INNER, s);
  }
  if (frame_side) {
    regrow(p.mel, (size_t)rows * N_MEL, s);
    regrow(p.post, (size_t)rows * N_MEL, s);
    if (bf16) {
      regrow(p.melb, (size_t)rows * N_MEL, s);
      regrow(p.pnb[0], (size_t)rows * PN_DIM, s);
      regrow(p.pnb[1], (size_t)rows * PN_DIM, s);
    } else {
      regrow(p.pn[0], (size_t)rows * PN_DIM, s);
      regrow(p.pn[1], (size_t)rows * PN_DIM, s);
    }
  }
  p.rows = rows;
}

static void check_status(fs2_ctx* c, int32_t st) {
  if (st == 0) return;
  std::string m = "invalid input:";
  if (st & ERR_BAD_LEN) m += " src_lens outside [0, max_src_len];";
  if (st & ERR_BAD_ID) m += " phoneme id outside the embedding table;";
  if (st & ERR_BAD_INDEX) m += " speaker/emotion/arousal/valence index outside its table;";
  if (st & ERR_MAXLEN_SMALL) m += " max_mel_len smaller than the longest expanded utterance;";
  throw Error(FS2_ERR_INVALID, m);
}

enum { STAGE2_BODY = 1, STAGE2_UNPACK = 2 };
static void stage2(fs2_ctx* c, cudaStream_t s, const fs2_stage2_io* io, int phases);

static void stage1(fs2_ctx* c, cudaStream_t s, const fs2_inputs* in, fs2_stage1_out* out) {
  require(c->prepared, FS2_ERR_STATE, "fs2_forward_stage1 called before fs2_prepare");
  require(in && out, FS2_ERR_INVALID, "null argument");
  const int B = in->batch, L = in->max_src_len;
  require(B > 0 && B <= 65535 && L > 0, FS2_ERR_INVALID, "batch must be in [1,65535] and max_src_len > 0");
  require(in->speakers && in->emotions && in->arousals && in->valences && in->texts && in->src_lens, FS2_ERR_INVALID,
          "null input tensor");
  require(out->pitch && out->energy && out->log_d && out->d_rounded && out->src_mask && out->mel_lens, FS2_ERR_INVALID,
          "null output tensor");
  FS2_CUDA_OK(cudaSetDevice(c->device));
  const auto t_begin = std::chrono::steady_clock::now();
  g_launches = 0;
  c->stage1_done = false;
  c->stage2_body_done = false;
  for (auto& r : c->prof) { c->event_pool.push_back(r.beg); c->event_pool.push_back(r.end); }
  c->prof.clear();
  if (c->debug) {
    for (auto& kv : c->taps) cudaFree(kv.second.first);
    c->taps.clear();
  }

  const int64_t bound = (int64_t)GAP_PHON + (int64_t)B * (L + GAP_PHON);
  require(bound < (1LL << 30), FS2_ERR_INVALID, "batch * max_src_len too large");
  const int rows = round_up((int)bound, 128);
  const bool bf = c->cfg.math_mode == FS2_MATH_BF16;
  // query rows per attention work item: the paired (K/V-multicast) kernel where utterances span several query tiles
  c->ps.work_q_rows = (bf || L <= attn_tc::BQ) ? attn_tc::BQ : attn_tc::query_rows_per_entry((int64_t)B * L, B, L);
  ensure_side(c->ps, B, rows, s, attn_tc::work_bound((int64_t)B * L, B, L, c->ps.work_q_rows));
  ensure_side(c->fs, B, 0, s);
  ensure_pool(c->pp, rows, false, bf, s);
  const int64_t BL = (int64_t)B * L;
  if (BL > c->scratch_bl) {
    regrow(c->cum, BL, s);
    regrow(c->raw_pitch, BL, s);
    regrow(c->raw_energy, BL, s);
    c->scratch_bl = BL;
  }
  if (B > c->scratch_b) {
    regrow(c->mel_lens32, B, s);
    regrow(c->cond_spk, (size_t)B * D_MODEL, s);
    regrow(c->cond_emo, (size_t)B * D_MODEL, s);
    c->scratch_b = B;
  }

  RowSide& ps = c->ps;
  Pool& pp = c->pp;
  // (the status word is zero on entry: fs2_create clears it and every stage 1 clears it again after reading it back)
  {
    ProfScope pr(c, s, "layout_scan");
    launch_row(layout_scan_kernel<int64_t>, 1, 1024, s, in->src_lens, B, GAP_PHON, L, L, ps.starts, ps.lens, ps.totals, c->status,
               nullptr);
    FS2_LAUNCHED();
  }
  // row metadata + attention work list + zero fill of the [B, L] outputs that are written at real positions only
  // + the source padding mask: one launch
  SlotInit init{};
  init.zero[0] = out->pitch; init.zero[1] = out->energy; init.zero[2] = out->log_d; init.zero[3] = out->d_rounded;
  init.zero[4] = c->raw_pitch; init.zero[5] = c->raw_energy;
  init.n = BL; init.src_mask = out->src_mask; init.src_lens = in->src_lens; init.max_src_len = L;
  {
    ProfScope pr(c, s, "row_meta");
    launch_row(row_meta_kernel, (rows + 255) / 256, 256, s, ps.starts, ps.lens, B, GAP_PHON, L, nullptr, rows, ps.utt, ps.vpos,
               ps.room, ps.slot, ps.work, ps.work_cap, ps.work_count, init, ps.work_q_rows);
    FS2_LAUNCHED();
  }

  // ---- Encoder (transformer/Models.py:73-100)
  float *x = pp.act[0], *t1 = pp.act[1], *t2 = pp.act[2], *t3 = pp.act[3];
  const float* pe = position_rows(c, "encoder.position_enc", L, s);
  {
    ProfScope pr(c, s, "embed_pe");
    launch_row(embed_pe_kernel, (rows + 7) / 8, 256, s, in->texts, L, c->raw.at("encoder.src_word_emb.weight").ptr,
               c->cfg.n_src_vocab, pe, ps.meta(), ps.lens, rows, x, c->status, pp.actb[0]);
    FS2_LAUNCHED();
  }
  tap(c, s, "p_start", ps.starts, 1, B + 1);
  tap(c, s, "enc_in", x, rows, D_MODEL);
  // ---- conditioning vectors (model/fastspeech2.py:101-110): they depend on the inputs only, so they are computed before
  // the encoder and added to its output inside the last layer's LayerNorm epilogue (of the w2 GEMM, or of the fused FFN
  // kernel on large batches).  The stand-alone add remains for the per-layer debug taps (which want the unconditioned
  // encoder output) and for stage-1 fusion flag bit 1 cleared (A/B tests).
  launch_row(cond_kernel, B * COND_PARTS, 256, s, in->speakers, in->emotions, in->arousals, in->valences, c->raw.at("speaker_emb.weight").ptr,
                                c->cfg.n_speaker, c->raw.at("emotion_emb.weight").ptr, c->cfg.n_emotion,
                                c->raw.at("arousal_emb.weight").ptr, c->cfg.n_arousal, c->raw.at("valence_emb.weight").ptr,
                                c->cfg.n_valence, c->raw.at("emotion_linear.0.weight").ptr,
                                c->raw.at("emotion_linear.0.bias").ptr, c->cond_spk, c->cond_emo, c->status);
  FS2_LAUNCHED();
  const bool cond_fused = (stage1_fusion_flag() & 2) != 0 && !c->debug;
  const CondAdd cond{c->cond_spk, c->cond_emo};
  for (int i = 0; i < ENC_LAYERS; ++i) {
    fft_block(c, s, c->enc[i], ps, pp, rows, B, L, x, t1, t2, false, cond_fused && i == ENC_LAYERS - 1 ? &cond : nullptr);
    if (c->debug) tap(c, s, ("enc_" + std::to_string(i)).c_str(), x, rows, D_MODEL);
  }
  float* xc = cond_fused ? x : t3;                       // the conditioned rows
  float* spare = cond_fused ? t3 : x;                    // the other [rows,256] buffer: the pitch add writes there
  __nv_bfloat16* xcb = cond_fused ? pp.actb[0] : pp.actb[3];
  __nv_bfloat16* spareb = cond_fused ? pp.actb[3] : pp.actb[0];
  if (!cond_fused) {
    ProfScope pr(c, s, "add_cond");
    add_cond_kernel<<<(rows + 7) / 8, 256, 0, s>>>(x, ps.meta(), c->cond_spk, c->cond_emo, 2, rows, xc, xcb);
    FS2_LAUNCHED();
  }
  tap(c, s, "cond_x", xc, rows, D_MODEL);

  // ---- VarianceAdaptor (model/modules.py:102-135)
  // phoneme_level features run here (modules.py:114-125); frame_level ones after the LengthRegulator in stage 2
  const bool pitch_here = !c->cfg.pitch_frame_level, energy_here = !c->cfg.energy_frame_level;
  // the duration and pitch predictors read the same rows: one launch per layer for both (FS2_PRED_PAIR=0 / the
  // split-operand mode run them one after the other)
  const bool paired = pitch_here && (stage1_fusion_flag() & 1) != 0 && c->cfg.math_mode != FS2_MATH_TF32X3;
  if (paired) predictor_pair(c, s, c->pred[0], c->pred[1], ps, rows, xc, t1, t2, out->log_d, c->raw_pitch, xcb, pp.actb[1], pp.actb[2]);
  else predictor(c, s, c->pred[0], ps, rows, xc, t1, out->log_d, xcb, pp.actb[1]);
  float* cur = xc;
  __nv_bfloat16* curb = xcb;
  if (pitch_here) {
    if (!paired) predictor(c, s, c->pred[1], ps, rows, cur, t1, c->raw_pitch, curb, pp.actb[1]);
    float* xe = spare;
    ProfScope pr(c, s, "bucket_embed_add");
    launch_row(bucket_embed_add_kernel, (rows + 7) / 8, 256, s,
               cur, ps.meta(), ps.slot, 2, rows, c->raw_pitch, in->p_targets, in->p_control,
               c->raw.at("variance_adaptor.pitch_bins").ptr, N_BINS - 1, c->raw.at("variance_adaptor.pitch_embedding.weight").ptr,
               out->pitch, nullptr, xe, spareb);
    FS2_LAUNCHED();
    cur = xe;
    curb = spareb;
  }
  // phoneme_level energy: only the predictor runs here; its bucketize + embedding add (modules.py:93-100,126) is fused
  // into the length regulator of stage 2, and the returned prediction is written by the durations kernel below
  if (energy_here) predictor(c, s, c->pred[2], ps, rows, cur, t1, c->raw_energy, curb, pp.actb[1]);
  float* xf = cur;
  c->va_spare = cur == xc ? spare : xc;   // free [rows,256] buffer (debug tap of the fused energy add)
  c->p_targets = in->p_targets;
  c->e_targets = in->e_targets;
  c->p_control = in->p_control;
  if (!energy_here) tap(c, s, "va_x", xf, rows, D_MODEL);

  const bool forced = in->d_targets != nullptr;
  launch_row(durations_kernel, (B + 7) / 8, 256, s, forced ? in->d_targets : out->log_d, forced ? 1 : 0, in->d_control,
             in->src_lens, B, L, forced ? nullptr : out->d_rounded, c->cum,
             out->mel_lens, c->mel_lens32, energy_here ? c->raw_energy : nullptr,
             in->e_targets != nullptr ? 1.f : in->p_control /* sic: modules.py:123-125 */,
             energy_here ? out->energy : nullptr);
  FS2_LAUNCHED();
  // (the frame side's sizes and the stage's status word go straight to pinned host memory: see the kernel)
  launch_row(layout_scan_kernel<int32_t>, 1, 1024, s, c->mel_lens32, B, GAP_FRAME, 0, in->max_mel_len, c->fs.starts, c->fs.lens,
             c->fs.totals, c->status, c->h_totals);
  FS2_LAUNCHED();

  static const bool timing = std::getenv("FS2_TIMING") != nullptr;
  const auto t_enq = std::chrono::steady_clock::now();
  // ---- the one blocking point: sizes of the frame side
  FS2_CUDA_OK(cudaStreamSynchronize(s));
  if (timing) {
    const auto t_end = std::chrono::steady_clock::now();
    fprintf(stderr, "[fs2] stage1: enqueue %.1f us (%d launches), wait %.1f us\n",
            std::chrono::duration<double, std::micro>(t_enq - t_begin).count(), g_launches,
            std::chrono::duration<double, std::micro>(t_end - t_enq).count());
  }
  check_status(c, *reinterpret_cast<int32_t*>(c->h_totals + 3));
  c->frame_rows = c->h_totals[0];
  c->max_mel_len = (int)c->h_totals[1];
  out->total_frames = c->total_frames = c->h_totals[2];
  out->max_mel_len = c->max_mel_len;
  require(c->frame_rows < (1LL << 30), FS2_ERR_INVALID, "expanded batch too large");
  c->batch = B;
  c->phon_rows = rows;
  c->max_src_len = L;
  c->lr_input = xf;
  c->lr_adds_energy = true;
  c->stage1_done = true;
  c->last_launches = g_launches;
  // Eager stage 2 (fs2_set_eager_stage2): everything of stage 2 but the final unpack needs no caller buffer, so it is enqueued
  // here, right behind the one host synchronisation -- the device does not wait for the caller to allocate its outputs and
  // come back (tens of microseconds through a Python facade).  Not with frame_level features (they write caller buffers inside
  // stage 2) and not with debug taps.
  // (an asynchronous read of the previous forward's packed rows, which stage 2 overwrites, finishes first; the event is only
  // guaranteed to live through this call, so the wait is enqueued here whether or not the body follows)
  if (c->stage2_wait != nullptr) FS2_CUDA_OK(cudaStreamWaitEvent(s, c->stage2_wait, 0));
  c->stage2_wait = nullptr;
  if (c->eager_stage2 && !c->debug && !c->cfg.pitch_frame_level && !c->cfg.energy_frame_level) {
    stage2(c, s, nullptr, STAGE2_BODY);
    c->stage2_body_done = true;
  }
}

static void stage2(fs2_ctx* c, cudaStream_t s, const fs2_stage2_io* io, int phases) {
  require(c->stage1_done, FS2_ERR_STATE, "fs2_forward_stage2 called without a completed fs2_forward_stage1");
  // an all-zero duration batch has max_mel_len == 0: the outputs are empty tensors (null data pointers) and nothing runs
  require(!(phases & STAGE2_UNPACK) || (io && (c->max_mel_len == 0 || (io->mel && io->postnet && io->mel_mask))), FS2_ERR_INVALID,
          "null stage-2 output");
  const auto t_begin2 = std::chrono::steady_clock::now();
  FS2_CUDA_OK(cudaSetDevice(c->device));
  g_launches = 0;
  const int B = c->batch, L = c->max_src_len, T = c->max_mel_len;
  const int rows = round_up((int)c->frame_rows, 128);
  const int math = c->cfg.math_mode;
  const bool bf = math == FS2_MATH_BF16;
  c->fs.work_q_rows = (bf || T <= attn_tc::BQ) ? attn_tc::BQ : attn_tc::query_rows_per_entry(c->total_frames, B, T);
  ensure_side(c->fs, B, rows, s, attn_tc::work_bound(c->total_frames, B, T, c->fs.work_q_rows));
  ensure_pool(c->fp, rows, true, bf, s);
  RowSide& fsd = c->fs;
  Pool& fp = c->fp;

  if (T > 0 && (phases & STAGE2_BODY)) {
    launch_row(row_meta_kernel, (rows + 255) / 256, 256, s, fsd.starts, fsd.lens, B, GAP_FRAME, T, nullptr, rows, fsd.utt,
               fsd.vpos, fsd.room, fsd.slot, fsd.work, fsd.work_cap, fsd.work_count, SlotInit{}, fsd.work_q_rows);
    FS2_LAUNCHED();
    // ---- LengthRegulator + decoder positional encoding (modules.py:167-194, Models.py:145-162)
    float *x = fp.act[0], *t1 = fp.act[1], *t2 = fp.act[2];
    const float* pe = position_rows(c, "decoder.position_enc", T, s);
    const bool pitch_f = c->cfg.pitch_frame_level != 0, energy_f = c->cfg.energy_frame_level != 0;
    {
      ProfScope ps(c, s, "length_regulator");
      // source-driven expansion: one warp per phoneme row + one per reserved frame row (the frame->phoneme search of
      // length_regulate_kernel / fs2_op_frame_map stays as the cross-check used by the tests)
      const int rows_p = c->phon_rows;                                 // phoneme rows laid out by stage 1
      const int reserved = (B + 1) * GAP_FRAME + 128;                  // gaps + the round-up tail
      const int warps = rows_p + reserved;
      EnergyAdd en{};
      if (!energy_f && c->lr_adds_energy) {   // phoneme_level energy: bucketize + embedding add applied to the row on its way through
        en.raw = c->raw_energy; en.target = c->e_targets; en.control = c->p_control;
        en.bins = c->raw.at("variance_adaptor.energy_bins").ptr; en.n_bins = N_BINS - 1;
        en.table = c->raw.at("variance_adaptor.energy_embedding.weight").ptr;
        if (c->debug) {
          en.va_out = c->va_spare;
          FS2_CUDA_OK(cudaMemsetAsync(en.va_out, 0, (size_t)rows_p * D_MODEL * sizeof(float), s));
        }
      }
      launch_row(length_regulate_scatter_kernel, (warps + 7) / 8, 256, s,
                 c->lr_input, c->ps.meta(), c->ps.lens, rows_p, c->cum, L, fsd.starts, fsd.lens, B, GAP_FRAME, fsd.totals,
                 (pitch_f || energy_f) ? nullptr : pe, rows, x, fp.actb[0], en);
      FS2_LAUNCHED();
      if (en.va_out != nullptr) tap(c, s, "va_x", en.va_out, rows_p, D_MODEL);
    }
    if (pitch_f || energy_f) {
      // frame_level predictors (modules.py:139-148) on the expanded rows: padding rows of the LengthRegulator output
      // are zero (utils/tools.py:360-378), the masked prediction there is 0 and its bucket embedding is still added,
      // so the first two reserved rows are carried exactly as on the phoneme side.  The positional encoding is added
      // afterwards and every reserved row returns to zero for the decoder.
      const int64_t BT = (int64_t)B * T;
      if (BT > c->scratch_bt) {
        regrow(c->raw_pitch_f, BT, s);
        regrow(c->raw_energy_f, BT, s);
        c->scratch_bt = BT;
      }
      float* cur = x;                    // act[0]
      __nv_bfloat16* curb = fp.actb[0];
      if (pitch_f) {
        require(io->pitch_frames != nullptr, FS2_ERR_INVALID, "pitch is frame_level: stage-2 io needs pitch_frames");
        FS2_CUDA_OK(cudaMemsetAsync(io->pitch_frames, 0, BT * sizeof(float), s));
        FS2_CUDA_OK(cudaMemsetAsync(c->raw_pitch_f, 0, BT * sizeof(float), s));
        predictor(c, s, c->pred[1], fsd, rows, cur, t1, c->raw_pitch_f, curb, fp.actb[1]);
        bucket_embed_add_kernel<<<(rows + 7) / 8, 256, 0, s>>>(
            cur, fsd.meta(), fsd.slot, energy_f ? 2 : 0, rows, c->raw_pitch_f, c->p_targets, c->p_control,
            c->raw.at("variance_adaptor.pitch_bins").ptr, N_BINS - 1,
            c->raw.at("variance_adaptor.pitch_embedding.weight").ptr, io->pitch_frames, nullptr, fp.act[3], fp.actb[3]);
        FS2_LAUNCHED();
        cur = fp.act[3];
        curb = fp.actb[3];
      }
      if (energy_f) {
        require(io->energy_frames != nullptr, FS2_ERR_INVALID, "energy is frame_level: stage-2 io needs energy_frames");
        FS2_CUDA_OK(cudaMemsetAsync(io->energy_frames, 0, BT * sizeof(float), s));
        FS2_CUDA_OK(cudaMemsetAsync(c->raw_energy_f, 0, BT * sizeof(float), s));
        predictor(c, s, c->pred[2], fsd, rows, cur, t1, c->raw_energy_f, curb, fp.actb[1]);
        float* dst = cur == x ? fp.act[3] : x;
        bucket_embed_add_kernel<<<(rows + 7) / 8, 256, 0, s>>>(
            cur, fsd.meta(), fsd.slot, 0, rows, c->raw_energy_f, c->e_targets, c->p_control /* sic */,
            c->raw.at("variance_adaptor.energy_bins").ptr, N_BINS - 1,
            c->raw.at("variance_adaptor.energy_embedding.weight").ptr, io->energy_frames, nullptr, dst);
        FS2_LAUNCHED();
        cur = dst;
      }
      tap(c, s, "va_frames", cur, rows, D_MODEL);
      add_pe_kernel<<<(rows + 7) / 8, 256, 0, s>>>(cur, fsd.meta(), fsd.lens, pe, rows, x, fp.actb[0]);
      FS2_LAUNCHED();
    }
    tap(c, s, "f_start", fsd.starts, 1, B + 1);
    tap(c, s, "dec_in", x, rows, D_MODEL);
    for (int i = 0; i < DEC_LAYERS; ++i) {
      fft_block(c, s, c->dec[i], fsd, fp, rows, B, T, x, t1, t2, true);
      if (c->debug) tap(c, s, ("dec_" + std::to_string(i)).c_str(), x, rows, D_MODEL);
    }
    // ---- mel_linear (fastspeech2.py:134); the first min(10, T_max - T_b) reserved rows get the bias
    ConvGemmArgs a = bf ? gemm_args_b(fp.actb[0], D_MODEL, rows, c->mel_w, c->mel_b, 1, D_MODEL, N_MEL, ACT_NONE, fp.mel, N_MEL,
                                      fp.melb, N_MEL)
                        : gemm_args(x, D_MODEL, rows, c->mel_w, c->mel_b, 1, D_MODEL, N_MEL, ACT_NONE, fp.mel, N_MEL);
    a.row_vpos = fsd.vpos; a.row_room = fsd.room; a.extra = PN_VIRTUAL;
    { ProfScope ps(c, s, "mel_linear"); conv_gemm(c, a, s); }
    tap(c, s, "mel_p", fp.mel, rows, N_MEL);
    // ---- PostNet (Layers.py:129-137) + residual (fastspeech2.py:136), BatchNorm folded
    const float* src = fp.mel;
    const __nv_bfloat16* srcb = fp.melb;
    int ld = N_MEL;
    for (int j = 0; j < PN_LAYERS; ++j) {
      const PostConv& pc = c->post[j];
      const bool last = j == PN_LAYERS - 1;
      float* dst = last ? fp.post : fp.pn[j & 1];
      if (bf) {   // intermediate layers exist only as bf16 operands; the last one writes the fp32 result
        __nv_bfloat16* dstb = last ? nullptr : fp.pnb[j & 1];
        a = gemm_args_b(srcb, ld, rows, pc.w, pc.b, PN_TAPS, pc.cin, pc.cout, last ? ACT_NONE : ACT_TANH,
                        last ? fp.post : nullptr, pc.cout, dstb, pc.cout);
        srcb = dstb;
      } else {
        a = gemm_args(src, ld, rows, pc.w, pc.b, PN_TAPS, pc.cin, pc.cout, last ? ACT_NONE : ACT_TANH, dst, pc.cout);
      }
      a.row_vpos = fsd.vpos; a.row_room = fsd.room; a.extra = PN_VIRTUAL - 2 * (j + 1);
      if (last) { a.residual = fp.mel; a.ldr = N_MEL; }
      { ProfScope ps(c, s, j == 0 ? "postnet.conv_80_512" : (last ? "postnet.conv_512_80" : "postnet.conv_512_512"));
        conv_gemm(c, a, s); }
      src = dst;
      ld = pc.cout;
    }
    tap(c, s, "post_p", fp.post, rows, N_MEL);
  }
  if (T > 0 && (phases & STAGE2_UNPACK)) {
    const int64_t out_rows = (int64_t)B * T;
    {
      ProfScope ps(c, s, "unpack");
      launch_row(unpack_mel_kernel, (unsigned)((out_rows * (N_MEL / 4) + 255) / 256), 256, s, fp.mel, fp.post, fsd.starts, fsd.lens, B, T,
                 c->mel_b, io->mel, io->postnet, io->mel_mask);
      FS2_LAUNCHED();
    }
  }
  c->last_launches += g_launches;
  static const bool timing2 = std::getenv("FS2_TIMING") != nullptr;
  if (timing2) {
    fprintf(stderr, "[fs2] stage2: enqueue %.1f us (%d launches)\n",
            std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_begin2).count(), g_launches);
  }
}

static void prepare(fs2_ctx* c, cudaStream_t s) {
  FS2_CUDA_OK(cudaSetDevice(c->device));
  for (void* p : c->owned) cudaFree(p);
  c->owned.clear();
  c->prepared = false;
  const int d = D_MODEL;
  W(c, "encoder.src_word_emb.weight", {c->cfg.n_src_vocab, d});
  W(c, "encoder.position_enc", {1, c->cfg.max_seq_len + 1, d});
  W(c, "decoder.position_enc", {1, c->cfg.max_seq_len + 1, d});
  for (int i = 0; i < ENC_LAYERS; ++i) prepare_fft(c, "encoder.layer_stack." + std::to_string(i), c->enc[i], s);
  for (int i = 0; i < DEC_LAYERS; ++i) prepare_fft(c, "decoder.layer_stack." + std::to_string(i), c->dec[i], s);
  const char* pn[3] = {"duration", "pitch", "energy"};
  for (int i = 0; i < 3; ++i) prepare_predictor(c, std::string("variance_adaptor.") + pn[i] + "_predictor", c->pred[i], s);
  W(c, "variance_adaptor.pitch_bins", {N_BINS - 1});
  W(c, "variance_adaptor.energy_bins", {N_BINS - 1});
  W(c, "variance_adaptor.pitch_embedding.weight", {N_BINS, d});
  W(c, "variance_adaptor.energy_embedding.weight", {N_BINS, d});
  W(c, "speaker_emb.weight", {c->cfg.n_speaker, d});
  W(c, "emotion_emb.weight", {c->cfg.n_emotion, d / 2});
  W(c, "arousal_emb.weight", {c->cfg.n_arousal, d / 4});
  W(c, "valence_emb.weight", {c->cfg.n_valence, d / 4});
  W(c, "emotion_linear.0.weight", {d, d});
  W(c, "emotion_linear.0.bias", {d});
  c->mel_w = repack_conv(c, W(c, "mel_linear.weight", {N_MEL, d}).ptr, N_MEL, d, 1, nullptr, s);
  c->mel_b = W(c, "mel_linear.bias", {N_MEL}).ptr;
  for (int j = 0; j < PN_LAYERS; ++j) {
    const int cin = j == 0 ? N_MEL : PN_DIM, cout = j == PN_LAYERS - 1 ? N_MEL : PN_DIM;
    const std::string cv = "postnet.convolutions." + std::to_string(j) + ".0.conv";
    const std::string bn = "postnet.convolutions." + std::to_string(j) + ".1";
    float* scale = keep(c, cout);
    float* bias = keep(c, cout);
    bn_fold_kernel<<<(cout + 255) / 256, 256, 0, s>>>(W(c, bn + ".weight", {cout}).ptr, W(c, bn + ".bias", {cout}).ptr,
                                                      W(c, bn + ".running_mean", {cout}).ptr,
                                                      W(c, bn + ".running_var", {cout}).ptr, W(c, cv + ".bias", {cout}).ptr,
                                                      cout, scale, bias);
    FS2_LAUNCHED();
    c->post[j] = PostConv{repack_conv(c, W(c, cv + ".weight", {cout, cin, PN_TAPS}).ptr, cout, cin, PN_TAPS, scale, s), bias,
                          cin, cout};
  }
  FS2_CUDA_OK(cudaStreamSynchronize(s));
  c->prepared = true;
}

// ---- hand-over of the length regulator's input between contexts (rebalancing a sharded batch by frames)
static void export_stage1(fs2_ctx* c, cudaStream_t s, float* hidden, int32_t* reps) {
  require(c->stage1_done && c->lr_adds_energy, FS2_ERR_STATE, "fs2_export_stage1 needs a completed fs2_forward_stage1 of this context");
  require(hidden && reps, FS2_ERR_INVALID, "null argument");
  require(!c->cfg.pitch_frame_level && !c->cfg.energy_frame_level, FS2_ERR_UNSUPPORTED,
          "export / import of stage 1 is implemented for phoneme_level features");
  FS2_CUDA_OK(cudaSetDevice(c->device));
  const int B = c->batch, L = c->max_src_len;
  EnergyAdd en{};
  en.raw = c->raw_energy; en.target = c->e_targets; en.control = c->p_control;
  en.bins = c->raw.at("variance_adaptor.energy_bins").ptr; en.n_bins = N_BINS - 1;
  en.table = c->raw.at("variance_adaptor.energy_embedding.weight").ptr;
  const int warps = B * L;
  export_rows_kernel<<<(warps + 7) / 8, 256, 0, s>>>(c->lr_input, c->ps.starts, c->ps.lens, c->cum, B, L, en, hidden, reps);
  FS2_LAUNCHED();
}

static void import_stage1(fs2_ctx* c, cudaStream_t s, const float* hidden, const int32_t* reps, const int64_t* src_lens, int B,
                          int L, int max_mel_len, int64_t* mel_lens, int64_t* total_frames, int32_t* max_mel_len_out) {
  require(c->prepared, FS2_ERR_STATE, "fs2_import_stage1 called before fs2_prepare");
  require(hidden && reps && src_lens && mel_lens && total_frames && max_mel_len_out, FS2_ERR_INVALID, "null argument");
  require(B > 0 && B <= 65535 && L > 0, FS2_ERR_INVALID, "batch must be in [1,65535] and max_src_len > 0");
  require(!c->cfg.pitch_frame_level && !c->cfg.energy_frame_level, FS2_ERR_UNSUPPORTED,
          "export / import of stage 1 is implemented for phoneme_level features");
  FS2_CUDA_OK(cudaSetDevice(c->device));
  g_launches = 0;
  c->stage1_done = false;
  c->stage2_body_done = false;
  const int64_t bound = (int64_t)GAP_PHON + (int64_t)B * (L + GAP_PHON);
  require(bound < (1LL << 30), FS2_ERR_INVALID, "batch * max_src_len too large");
  const int rows = round_up((int)bound, 128);
  ensure_side(c->ps, B, rows, s, attn_tc::work_bound((int64_t)B * L, B, L));
  ensure_side(c->fs, B, 0, s);
  ensure_pool(c->pp, rows, false, c->cfg.math_mode == FS2_MATH_BF16, s);
  const int64_t BL = (int64_t)B * L;
  if (BL > c->scratch_bl) {
    regrow(c->cum, BL, s);
    regrow(c->raw_pitch, BL, s);
    regrow(c->raw_energy, BL, s);
    c->scratch_bl = BL;
  }
  if (B > c->scratch_b) {
    regrow(c->mel_lens32, B, s);
    regrow(c->cond_spk, (size_t)B * D_MODEL, s);
    regrow(c->cond_emo, (size_t)B * D_MODEL, s);
    c->scratch_b = B;
  }
  RowSide& ps = c->ps;
  layout_scan_kernel<int64_t><<<1, 1024, 0, s>>>(src_lens, B, GAP_PHON, L, L, ps.starts, ps.lens, ps.totals, c->status);
  FS2_LAUNCHED();
  row_meta_kernel<<<(rows + 255) / 256, 256, 0, s>>>(ps.starts, ps.lens, B, GAP_PHON, L, nullptr, rows, ps.utt, ps.vpos, ps.room,
                                                     ps.slot);
  FS2_LAUNCHED();
  float* x = c->pp.act[3];
  import_rows_kernel<<<(rows + 7) / 8, 256, 0, s>>>(hidden, ps.meta(), ps.lens, L, rows, x);
  FS2_LAUNCHED();
  reps_scan_kernel<<<(B + 7) / 8, 256, 0, s>>>(reps, src_lens, B, L, c->cum, mel_lens, c->mel_lens32);
  FS2_LAUNCHED();
  layout_scan_kernel<int32_t><<<1, 1024, 0, s>>>(c->mel_lens32, B, GAP_FRAME, 0, max_mel_len, c->fs.starts, c->fs.lens,
                                                 c->fs.totals, c->status);
  FS2_LAUNCHED();
  FS2_CUDA_OK(cudaMemcpyAsync(c->h_totals, c->fs.totals, 3 * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  FS2_CUDA_OK(cudaMemcpyAsync(c->h_totals + 3, c->status, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  FS2_CUDA_OK(cudaMemsetAsync(c->status, 0, sizeof(int32_t), s));
  FS2_CUDA_OK(cudaStreamSynchronize(s));
  check_status(c, *reinterpret_cast<int32_t*>(c->h_totals + 3));
  c->frame_rows = c->h_totals[0];
  c->max_mel_len = (int)c->h_totals[1];
  *total_frames = c->total_frames = c->h_totals[2];
  *max_mel_len_out = c->max_mel_len;
  require(c->frame_rows < (1LL << 30), FS2_ERR_INVALID, "expanded batch too large");
  c->batch = B;
  c->phon_rows = rows;
  c->max_src_len = L;
  c->lr_input = x;
  c->lr_adds_energy = false;
  c->p_targets = c->e_targets = nullptr;
  c->stage1_done = true;
  c->last_launches = g_launches;
}

template <typename F>
static int guarded(fs2_ctx* c, F&& f) {
  try {
    f();
    return FS2_OK;
  } catch (const Error& e) {
    if (c) c->err = e.msg; else g_create_error = e.msg;
    return e.code;
  } catch (const std::exception& e) {
    if (c) c->err = e.what(); else g_create_error = e.what();
    return FS2_ERR_INVALID;
  }
}

}  // namespace fs2

// ============================================================================ C ABI
extern "C" {

int fs2_version(void) { return 100; }

int fs2_create(const fs2_config* cfg, int device, fs2_ctx** out) {
  return guarded(nullptr, [&] {
    require(cfg && out, FS2_ERR_INVALID, "null argument");
    require(cfg->math_mode == FS2_MATH_TF32 || cfg->math_mode == FS2_MATH_BF16 || cfg->math_mode == FS2_MATH_TF32X3,
            FS2_ERR_UNSUPPORTED, "unknown math_mode");
    require(cfg->n_src_vocab > 0 && cfg->n_speaker > 0 && cfg->n_emotion > 0 && cfg->n_arousal > 0 && cfg->n_valence > 0 &&
                cfg->max_seq_len > 0, FS2_ERR_INVALID, "table sizes must be positive");
    int n_dev = 0;
    FS2_CUDA_OK(cudaGetDeviceCount(&n_dev));
    require(device >= 0 && device < n_dev, FS2_ERR_INVALID, "no such CUDA device");
    FS2_CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop{};
    FS2_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    require(prop.major == 10, FS2_ERR_UNSUPPORTED,
            std::string("libfs2b200 is built for sm_100a (B200) only; device is ") + prop.name);
    auto* c = new fs2_ctx();
    c->device = device;
    c->cfg = *cfg;
    c->status = dalloc<int32_t>(1);
    FS2_CUDA_OK(cudaMemset(c->status, 0, sizeof(int32_t)));
    c->splitk_ws = dalloc<float>(tc2::SPLITK_WS_BYTES / sizeof(float));
    FS2_CUDA_OK(cudaMallocHost(reinterpret_cast<void**>(&c->h_totals), 8 * sizeof(int64_t)));
    *out = c;
  });
}

void fs2_destroy(fs2_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto& kv : c->raw) cudaFree(kv.second.ptr);
  for (void* p : c->owned) cudaFree(p);
  for (auto& kv : c->taps) cudaFree(kv.second.first);
  for (auto& r : c->prof) { cudaEventDestroy(r.beg); cudaEventDestroy(r.end); }
  for (auto& e : c->event_pool) cudaEventDestroy(e);
  for (RowSide* sd : {&c->ps, &c->fs}) {
    cudaFree(sd->starts); cudaFree(sd->lens); cudaFree(sd->utt); cudaFree(sd->vpos); cudaFree(sd->room); cudaFree(sd->slot);
    cudaFree(sd->totals); cudaFree(sd->work); cudaFree(sd->work_count);
  }
  for (Pool* p : {&c->pp, &c->fp}) {
    for (float* a : p->act) cudaFree(a);
    cudaFree(p->qkv); cudaFree(p->hid); cudaFree(p->mel); cudaFree(p->post); cudaFree(p->pn[0]); cudaFree(p->pn[1]);
    for (auto* a : p->actb) cudaFree(a);
    cudaFree(p->hidb); cudaFree(p->qkvb); cudaFree(p->melb); cudaFree(p->pnb[0]); cudaFree(p->pnb[1]);
  }
  cudaFree(c->status); cudaFree(c->cum); cudaFree(c->mel_lens32); cudaFree(c->raw_pitch); cudaFree(c->raw_energy);
  cudaFree(c->cond_spk); cudaFree(c->cond_emo); cudaFree(c->pe_long); cudaFree(c->split_buf); cudaFree(c->splitk_ws); cudaFree(c->ffn_flags); cudaFree(c->raw_pitch_f); cudaFree(c->raw_energy_f);
  cudaFreeHost(c->h_totals);
  delete c;
}

const char* fs2_last_error(const fs2_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int fs2_set_weight(fs2_ctx* c, const char* key, const void* dev_ptr, const int64_t* shape, int ndim) {
  if (!c) return FS2_ERR_INVALID;
  return guarded(c, [&] {
    require(key && dev_ptr && (shape || ndim == 0) && ndim >= 0 && ndim <= 4, FS2_ERR_INVALID, "bad fs2_set_weight argument");
    FS2_CUDA_OK(cudaSetDevice(c->device));
    const std::string k(key);
    if (k.size() > 19 && k.compare(k.size() - 19, 19, "num_batches_tracked") == 0) return;  // int64 counter, unused in eval
    DevTensor t;
    t.numel = 1;
    for (int i = 0; i < ndim; ++i) {
      require(shape[i] > 0, FS2_ERR_INVALID, "non-positive dimension in " + k);
      t.shape.push_back(shape[i]);
      t.numel *= shape[i];
    }
    t.ptr = dalloc<float>(t.numel);
    // the source may have been produced on any stream (a caller's non-blocking stream is not ordered against the legacy
    // stream this blocking copy runs on): wait for the device first.  Weight upload is rare; this costs nothing that matters.
    FS2_CUDA_OK(cudaDeviceSynchronize());
    FS2_CUDA_OK(cudaMemcpy(t.ptr, dev_ptr, t.numel * sizeof(float), cudaMemcpyDeviceToDevice));
    auto it = c->raw.find(k);
    if (it != c->raw.end()) cudaFree(it->second.ptr);
    c->raw[k] = t;
    c->prepared = false;
  });
}

int fs2_prepare(fs2_ctx* c, fs2_stream stream) {
  if (!c) return FS2_ERR_INVALID;
  return guarded(c, [&] { prepare(c, static_cast<cudaStream_t>(stream)); });
}

int fs2_forward_stage1(fs2_ctx* c, fs2_stream stream, const fs2_inputs* in, fs2_stage1_out* out) {
  if (!c) return FS2_ERR_INVALID;
  return guarded(c, [&] { stage1(c, static_cast<cudaStream_t>(stream), in, out); });
}

int fs2_forward_stage2(fs2_ctx* c, fs2_stream stream, const fs2_stage2_io* io) {
  if (!c) return FS2_ERR_INVALID;
  return guarded(c, [&] {
    const bool body_done = c->stage2_body_done;
    c->stage2_body_done = false;
    stage2(c, static_cast<cudaStream_t>(stream), io, body_done ? STAGE2_UNPACK : (STAGE2_BODY | STAGE2_UNPACK));
  });
}

int fs2_set_eager_stage2(fs2_ctx* c, int on) {
  if (!c) return FS2_ERR_INVALID;
  c->eager_stage2 = on != 0;
  return FS2_OK;
}

int fs2_set_stage2_wait_event(fs2_ctx* c, void* event) {
  if (!c) return FS2_ERR_INVALID;
  c->stage2_wait = static_cast<cudaEvent_t>(event);
  return FS2_OK;
}

int fs2_export_stage1(fs2_ctx* c, fs2_stream stream, float* hidden, int32_t* reps) {
  if (!c) return FS2_ERR_INVALID;
  return guarded(c, [&] { export_stage1(c, static_cast<cudaStream_t>(stream), hidden, reps); });
}

int fs2_import_stage1(fs2_ctx* c, fs2_stream stream, const float* hidden, const int32_t* reps, const int64_t* src_lens,
                      int batch, int max_src_len, int max_mel_len, int64_t* mel_lens, int64_t* total_frames,
                      int32_t* max_mel_len_out) {
  if (!c) return FS2_ERR_INVALID;
  return guarded(c, [&] {
    import_stage1(c, static_cast<cudaStream_t>(stream), hidden, reps, src_lens, batch, max_src_len, max_mel_len, mel_lens,
                  total_frames, max_mel_len_out);
  });
}

int fs2_last_launch_count(const fs2_ctx* c) { return c ? c->last_launches : 0; }

int fs2_read_packed_postnet(fs2_ctx* c, fs2_stream stream, float* host_rows, int64_t max_rows, int32_t* host_starts,
                            int64_t* rows_out) {
  if (!c) return FS2_ERR_INVALID;
  return guarded(c, [&] {
    require(c->stage1_done, FS2_ERR_STATE, "fs2_read_packed_postnet needs a completed forward");
    require(host_rows && host_starts && rows_out, FS2_ERR_INVALID, "null argument");
    require(max_rows >= c->frame_rows, FS2_ERR_INVALID, "fs2_read_packed_postnet: destination too small");
    FS2_CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    *rows_out = c->frame_rows;
    if (c->frame_rows > 0 && c->max_mel_len > 0) {
      FS2_CUDA_OK(cudaMemcpyAsync(host_rows, c->fp.post, (size_t)c->frame_rows * N_MEL * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    FS2_CUDA_OK(cudaMemcpyAsync(host_starts, c->fs.starts, (size_t)(c->batch + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  });
}

int fs2_debug_enable(fs2_ctx* c, int on) {
  if (!c) return FS2_ERR_INVALID;
  c->debug = on != 0;
  return FS2_OK;
}

int fs2_debug_fetch(fs2_ctx* c, const char* name, void* host_dst, int64_t max_bytes, int64_t* rows, int64_t* cols) {
  if (!c) return FS2_ERR_INVALID;
  return guarded(c, [&] {
    auto it = c->taps.find(name ? name : "");
    require(it != c->taps.end(), FS2_ERR_INVALID, std::string("no such debug tap: ") + (name ? name : "(null)"));
    const auto& dims = it->second.second;
    if (rows) *rows = dims[0];
    if (cols) *cols = dims[1];
    const int64_t bytes = dims[0] * dims[1] * dims[2];
    if (host_dst) {
      require(max_bytes >= bytes, FS2_ERR_INVALID, "debug_fetch: destination too small");
      FS2_CUDA_OK(cudaSetDevice(c->device));
      FS2_CUDA_OK(cudaDeviceSynchronize());
      FS2_CUDA_OK(cudaMemcpy(host_dst, it->second.first, bytes, cudaMemcpyDeviceToHost));
    }
  });
}

static long long* g_trace_buf = nullptr;
static int g_trace_on = 0;

int fs2_debug_set_flag(int which, int value) {
  if (which == 0) fs2::attn_tc::debug_flag() = value;
  if (which == 2) fs2::tc2::cluster_size_flag() = value == 1 ? 1 : 2;
  if (which == 3) fs2::tc2::a_resident_flag() = value ? 1 : 0;
  if (which == 4) fs2::ffn::enabled_flag() = value;   // 0 off, 1 on, 2 automatic
  if (which == 5) fs2::tc2::n_split_flag() = value ? 1 : 0;
  if (which == 6) fs2::tc2::two_sm_flag() = value ? 1 : 0;
  if (which == 7) fs2::tc2::k_split_flag() = value ? 1 : 0;
  if (which == 8) fs2::attn_tc::pair_force_flag() = value;   // -1 automatic, 0 never, 1 always
  if (which == 9) stage1_fusion_flag() = value;
  if (which == 10) fs2::attn_p::enabled_flag() = value;
  if (which == 1) {
    g_trace_on = value;
    if (value && g_trace_buf == nullptr) cudaMalloc(reinterpret_cast<void**>(&g_trace_buf), 64 * sizeof(long long));
    if (value) cudaMemset(g_trace_buf, 0, 64 * sizeof(long long));
  }
  return FS2_OK;
}

int fs2_debug_read_trace(int64_t* host_dst, int n) {
#ifdef FS2_TRACE_BUILD
  if (n == 256 * 6) {   // per-CTA stamps of the last fused-FFN launch (tools/trace_ffn.py)
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(host_dst, fs2::ffn::g_ffn_cta_trace, (size_t)n * sizeof(long long)) == cudaSuccess ? FS2_OK : FS2_ERR_CUDA;
  }
  if (n == 16 * 8 + 2) {
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(host_dst, fs2::attn_p::g_attn_p_item_trace, 16 * 8 * sizeof(long long)) == cudaSuccess ? FS2_OK : FS2_ERR_CUDA;
  }
  if (n == 64 * 8 + 1) {   // per-tile stamps of one CTA of the last persistent attention launch (tools/trace_attention_persistent.py)
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(host_dst, fs2::attn_p::g_attn_p_tile_trace, 64 * 8 * sizeof(long long)) == cudaSuccess ? FS2_OK : FS2_ERR_CUDA;
  }
  if (n > 64) {   // per-CTA stamps of the last attention launch (tools/trace_attention_ctas.py)
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(host_dst, fs2::attn_tc::g_attn_cta_trace, std::min<size_t>(n, 2048 * 6) * sizeof(long long)) ==
                   cudaSuccess ? FS2_OK : FS2_ERR_CUDA;
  }
#endif
  if (g_trace_buf == nullptr || n > 64) return FS2_ERR_INVALID;
  g_trace_on = g_trace_on ? 1 : 0;
  cudaDeviceSynchronize();
  return cudaMemcpy(host_dst, g_trace_buf, n * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? FS2_OK : FS2_ERR_CUDA;
}

int fs2_profile_enable(fs2_ctx* c, int on) {
  if (!c) return FS2_ERR_INVALID;
  c->profiling = on != 0;
  return FS2_OK;
}

int fs2_profile_read(fs2_ctx* c, char* buf, int64_t buf_bytes) {
  if (!c) return FS2_ERR_INVALID;
  return guarded(c, [&] {
    FS2_CUDA_OK(cudaSetDevice(c->device));
    FS2_CUDA_OK(cudaDeviceSynchronize());
    std::map<std::string, std::pair<int, double>> agg;
    for (auto& r : c->prof) {
      float ms = 0.f;
      FS2_CUDA_OK(cudaEventElapsedTime(&ms, r.beg, r.end));
      auto& a = agg[r.label];
      a.first += 1;
      a.second += ms;
    }
    c->prof_text.clear();
    for (auto& kv : agg) {
      char line[160];
      snprintf(line, sizeof(line), "%s %d %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
      c->prof_text += line;
    }
    require(buf && buf_bytes > (int64_t)c->prof_text.size(), FS2_ERR_INVALID, "profile_read: buffer too small");
    memcpy(buf, c->prof_text.c_str(), c->prof_text.size() + 1);
  });
}

// ---------------------------------------------------------------------------- single operators
static thread_local std::string g_op_error;

int fs2_op_conv_gemm(fs2_stream stream, int math_mode, const float* A, int lda, int rows, const float* Wt,
                     const float* bias, int taps, int pad, int K, int N, int act, const float* residual, int ldr,
                     const int32_t* row_vpos, const int32_t* row_room, int extra, float* C, int ldc) {
  return guarded(nullptr, [&] {
    require(A && Wt && bias && C && rows >= 0 && taps >= 1 && K > 0 && N > 0, FS2_ERR_INVALID, "bad conv_gemm argument");
    require(math_mode == FS2_MATH_TF32 || math_mode == FS2_MATH_TF32X3, FS2_ERR_INVALID, "conv_gemm: TF32 or TF32X3 (bf16 has its own entry)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ConvGemmArgs a{};
    a.A = A; a.lda = lda; a.rows = rows; a.W = Wt; a.bias = bias; a.taps = taps; a.pad = pad; a.K = K; a.N = N; a.act = act;
    a.residual = residual; a.ldr = ldr; a.row_vpos = row_vpos; a.row_room = row_room; a.extra = extra; a.C = C; a.ldc = ldc;
    if (g_trace_on) { a.trace = g_trace_buf + 8 * ((g_trace_on - 1) % 8); ++g_trace_on; }
    if (math_mode == FS2_MATH_TF32X3 && rows > 0) {
      // Wt is the plain fp32 weight [taps][N][K] here: both operands are split into temporaries for this one call
      require(K % 4 == 0 && lda % 4 == 0, FS2_ERR_INVALID, "conv_gemm: K and lda must be multiples of 4");
      const int64_t nw = (int64_t)taps * N * K, n4 = (int64_t)rows * (K / 4);
      float* ws = dalloc<float>(2 * nw);
      float* as = dalloc<float>((size_t)rows * 2 * K);
      // a [taps][N][K] tensor is the [Cout = taps*N][Cin = K][k = 1] case of the repack
      repack_conv_split_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, s>>>(Wt, taps * N, K, 1, nullptr, ws, nw);
      FS2_LAUNCHED();
      split_tf32_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, s>>>(A, rows, K, lda, as);
      FS2_LAUNCHED();
      a.A = as; a.lda = 2 * K; a.W = ws; a.terms = 3;
      tc2::launch(a, s);
      FS2_CUDA_OK(cudaStreamSynchronize(s));
      cudaFree(ws);
      cudaFree(as);
      return;
    }
    if (rows > 0 && rows <= tc2::BM) {   // single row tile: give the launch the K-split workspace the forward owns
      float* sk = dalloc<float>(tc2::SPLITK_WS_BYTES / sizeof(float));
      a.splitk_ws = sk;
      tc2::launch(a, s);
      FS2_CUDA_OK(cudaStreamSynchronize(s));
      cudaFree(sk);
      return;
    }
    tc2::launch(a, s);
  });
}

int fs2_op_conv_gemm_ln(fs2_stream stream, const float* A, int lda, int rows, const float* Wt, const float* bias,
                        int taps, int pad, int K, int act, const float* residual, int ldr, const float* gamma,
                        const float* beta, const int32_t* row_vpos, const int32_t* row_room, int extra, float* C, int ldc,
                        const float* head_w, const float* head_b, float* head_out) {
  return guarded(nullptr, [&] {
    require(A && Wt && bias && gamma && beta && rows >= 0 && taps >= 1 && K > 0, FS2_ERR_INVALID, "bad conv_gemm_ln argument");
    ConvGemmArgs a{};
    a.A = A; a.lda = lda; a.rows = rows; a.W = Wt; a.bias = bias; a.taps = taps; a.pad = pad; a.K = K; a.N = D_MODEL; a.act = act;
    a.residual = residual; a.ldr = ldr; a.row_vpos = row_vpos; a.row_room = row_room; a.extra = extra; a.C = C; a.ldc = ldc;
    a.ln_gamma = gamma; a.ln_beta = beta; a.head_w = head_w; a.head_b = head_b; a.head_out = head_out;
    if (g_trace_on) { a.trace = g_trace_buf + 8 * ((g_trace_on - 1) % 8); ++g_trace_on; }
    tc2::launch(a, static_cast<cudaStream_t>(stream));
  });
}

int fs2_op_ffn_fused(fs2_stream stream, const float* x, int rows, const float* w1, const float* b1, const float* w2,
                     const float* b2, const float* gamma, const float* beta, const int32_t* row_vpos, const int32_t* row_room,
                     int extra, float* y) {
  return guarded(nullptr, [&] {
    require(x && w1 && b1 && w2 && b2 && gamma && beta && y && rows >= 0, FS2_ERR_INVALID, "bad ffn_fused argument");
    ffn::Args f{};
    f.x = x; f.rows = rows; f.w1 = w1; f.b1 = b1; f.w2 = w2; f.b2 = b2; f.gamma = gamma; f.beta = beta;
    f.row_vpos = row_vpos; f.row_room = row_room; f.extra = extra; f.y = y;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float* part = nullptr;
    int32_t* flags = nullptr;
    if (rows > 0) {
      part = dalloc<float>((size_t)rows * D_MODEL);
      flags = dalloc<int32_t>(ffn::flag_count(rows));
      FS2_CUDA_OK(cudaMemsetAsync(flags, 0, ffn::flag_count(rows) * sizeof(int32_t), s));
    }
    f.partial = part; f.flags = flags; f.epoch = 1;
    if (g_trace_on) f.trace = g_trace_buf;
    ffn::launch(f, s);
    FS2_CUDA_OK(cudaStreamSynchronize(s));
    cudaFree(part); cudaFree(flags);
  });
}

int fs2_op_conv_gemm_ex(fs2_stream stream, const float* A, int lda, int rows, const float* Wt, const float* bias, int taps,
                        int dil, int K, int N, int act, float slope, const float* residual, int ldr, int res_inv_lrelu,
                        int act2, const int32_t* row_vpos, const int32_t* row_room, int extra, int mask_shift, float* C,
                        int ldc) {
  return guarded(nullptr, [&] {
    require(A && Wt && bias && C && rows >= 0 && taps >= 1 && dil >= 1 && K > 0 && N > 0 && mask_shift >= 0, FS2_ERR_INVALID,
            "bad conv_gemm_ex argument");
    ConvGemmArgs a{};
    a.A = A; a.lda = lda; a.rows = rows; a.W = Wt; a.bias = bias; a.taps = taps; a.dil = dil; a.pad = dil * (taps - 1) / 2;
    a.K = K; a.N = N; a.act = act; a.slope = slope; a.residual = residual; a.ldr = ldr; a.res_inv_lrelu = res_inv_lrelu;
    a.act2 = act2; a.row_vpos = row_vpos; a.row_room = row_room; a.extra = extra; a.mask_shift = mask_shift; a.C = C; a.ldc = ldc;
    if (g_trace_on) { a.trace = g_trace_buf + 8 * ((g_trace_on - 1) % 8); ++g_trace_on; }
    tc2::launch(a, static_cast<cudaStream_t>(stream));
  });
}

int fs2_op_conv_gemm_bf16(fs2_stream stream, const void* A, int lda, int rows, const void* Wt, const float* bias, int taps,
                          int pad, int K, int N, int act, const float* residual, int ldr, const float* gamma,
                          const float* beta, const int32_t* row_vpos, const int32_t* row_room, int extra, float* C, int ldc,
                          void* C2, int ldc2) {
  return guarded(nullptr, [&] {
    require(A && Wt && bias && (C || C2) && rows >= 0 && taps >= 1 && K > 0 && N > 0, FS2_ERR_INVALID,
            "bad conv_gemm_bf16 argument");
    ConvGemmArgs a{};
    a.A = static_cast<const float*>(A); a.lda = lda; a.rows = rows; a.W = static_cast<const float*>(Wt); a.bias = bias;
    a.taps = taps; a.pad = pad; a.K = K; a.N = N; a.act = act; a.residual = residual; a.ldr = ldr;
    a.row_vpos = row_vpos; a.row_room = row_room; a.extra = extra; a.C = C; a.ldc = ldc; a.C2 = C2; a.ldc2 = ldc2;
    a.ln_gamma = gamma; a.ln_beta = beta; a.a_bf16 = 1;
    tc2::launch(a, static_cast<cudaStream_t>(stream));
  });
}

int fs2_op_attention(fs2_stream stream, const float* qkv, int rows, const int32_t* starts, const int32_t* lens,
                     int batch, int max_len, float* out) {
  return guarded(nullptr, [&] {
    require(qkv && starts && lens && out && rows > 0 && batch > 0 && max_len > 0, FS2_ERR_INVALID, "bad attention argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // the longest-first work list the forward builds in row_meta_kernel, into a temporary
    // paired (K/V multicast) form when there is enough multi-tile work, exactly as the forward decides
    const int q_rows = (max_len <= attn_tc::BQ && attn_tc::pair_force_flag() < 1) ? attn_tc::BQ
                       : attn_tc::query_rows_per_entry((int64_t)batch * max_len, batch, max_len);
    const int cap = attn_tc::work_bound((int64_t)batch * max_len, batch, max_len, q_rows);
    uint32_t* work = dalloc<uint32_t>(cap);
    int32_t* count = dalloc<int32_t>(1);
    attention_work_kernel<<<1, 256, 0, s>>>(lens, batch, work, cap, count, q_rows);
    FS2_LAUNCHED();
    if (attn_tc::use_two_sm(q_rows)) attn2::launch(qkv, rows, starts, lens, work, count, cap, out, s);
    else if (q_rows == attn_tc::BQ && attn_tc::debug_flag() == 0 && attn_p::use_persistent(cap, tc2::sm_count())) {
      if (attn_p::enabled_flag() == 3) attn_q::launch(qkv, rows, starts, lens, work, count, cap, out, s, tc2::sm_count());
      else attn_p::launch(qkv, rows, starts, lens, work, count, cap, out, s, tc2::sm_count());
    }
    else attn_tc::launch(qkv, rows, starts, lens, work, count, cap, q_rows, out, s);
    FS2_CUDA_OK(cudaStreamSynchronize(s));
    cudaFree(work);
    cudaFree(count);
  });
}

int fs2_op_attention_bf16(fs2_stream stream, const void* qkv, int rows, const int32_t* starts, const int32_t* lens, int batch,
                          int max_len, float* out) {
  return guarded(nullptr, [&] {
    require(qkv && starts && lens && out && rows > 0 && batch > 0 && max_len > 0, FS2_ERR_INVALID, "bad attention argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int cap = attn_tc::work_bound((int64_t)batch * max_len, batch, max_len);
    uint32_t* work = dalloc<uint32_t>(cap);
    int32_t* count = dalloc<int32_t>(1);
    attention_work_kernel<<<1, 256, 0, s>>>(lens, batch, work, cap, count, attn_bf::BQ);
    FS2_LAUNCHED();
    attn_bf::launch(static_cast<const __nv_bfloat16*>(qkv), rows, starts, lens, work, count, cap, out, nullptr, s);
    FS2_CUDA_OK(cudaStreamSynchronize(s));
    cudaFree(work);
    cudaFree(count);
  });
}

int fs2_op_durations(fs2_stream stream, const float* d_in, int is_target, float d_control, const int64_t* src_lens,
                     int batch, int max_src_len, float* d_rounded, int32_t* cum, int64_t* mel_lens) {
  return guarded(nullptr, [&] {
    require(d_in && src_lens && cum && batch > 0 && max_src_len > 0, FS2_ERR_INVALID, "bad durations argument");
    durations_kernel<<<(batch + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_in, is_target, d_control, src_lens, batch, max_src_len, d_rounded, cum, mel_lens, nullptr);
    FS2_LAUNCHED();
  });
}

int fs2_op_bucketize(fs2_stream stream, const float* values, int64_t n, const float* bins, int n_bins, int32_t* idx) {
  return guarded(nullptr, [&] {
    require(values && bins && idx && n >= 0 && n_bins >= 0, FS2_ERR_INVALID, "bad bucketize argument");
    if (n == 0) return;
    bucketize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(values, n, bins, n_bins, idx);
    FS2_LAUNCHED();
  });
}

int fs2_op_frame_map(fs2_stream stream, const int32_t* cum, int batch, int max_src_len, int max_mel_len, int32_t* map) {
  return guarded(nullptr, [&] {
    require(cum && map && batch > 0 && max_src_len > 0 && max_mel_len > 0, FS2_ERR_INVALID, "bad frame_map argument");
    const int64_t n = (int64_t)batch * max_mel_len;
    frame_map_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(cum, batch, max_src_len,
                                                                                                max_mel_len, map);
    FS2_LAUNCHED();
  });
}


}  // extern "C"

// ============================================================================ HiFi-GAN generator
struct fs2_voc {
  int device = 0;
  int math_mode = FS2_MATH_TF32;   // FS2_MATH_BF16: bf16 weights AND activations (fp32 accumulation, bias, output wav)
  std::string err;
  bool prepared = false;
  std::map<std::string, fs2::DevTensor> raw;
  std::vector<void*> owned;
  fs2::voc::ConvW pre, res1[fs2::voc::N_UPS][fs2::voc::N_RES][3], res2[fs2::voc::N_UPS][fs2::voc::N_RES][3];
  fs2::voc::UpW ups[fs2::voc::N_UPS];
  const float *post_w = nullptr, *post_b = nullptr;
  fs2::RowSide side;
  int64_t* lens64 = nullptr;
  float* melp = nullptr;        // packed mel [rows, 80]
  float* a0 = nullptr;          // conv_pre output [rows, 512]
  float* buf[7] = {};           // U, T1, P0, P1, R0, R1, R2: rows * 8192 floats each
  int rows_alloc = 0;
  int32_t* status = nullptr;
  int64_t* h_totals = nullptr;   // pinned
  int last_launches = 0;
};

namespace fs2 {
namespace voc {

static const DevTensor& VW(fs2_voc* c, const std::string& key, std::initializer_list<int64_t> shape) {
  auto it = c->raw.find(key);
  require(it != c->raw.end(), FS2_ERR_INVALID, "missing vocoder weight: " + key);
  require(it->second.shape == std::vector<int64_t>(shape), FS2_ERR_INVALID, "unexpected shape for vocoder weight: " + key);
  return it->second;
}
static float* vkeep(fs2_voc* c, size_t n) {
  float* p = dalloc<float>(n);
  c->owned.push_back(p);
  return p;
}
static ConvW conv_w(fs2_voc* c, const std::string& p, int cout, int cin, int k, cudaStream_t s) {
  ConvW w;
  w.cin = cin; w.cout = cout; w.k = k;
  const int64_t n = (int64_t)cout * cin * k;
  w.w = vkeep(c, n);
  if (c->math_mode == FS2_MATH_BF16)
    repack_conv_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(VW(c, p + ".weight", {cout, cin, k}).ptr, cout, cin, k, nullptr,
                                                                        reinterpret_cast<__nv_bfloat16*>(w.w));
  else
    repack_conv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(VW(c, p + ".weight", {cout, cin, k}).ptr, cout, cin, k, nullptr, 1, w.w);
  FS2_LAUNCHED();
  w.b = VW(c, p + ".bias", {cout}).ptr;
  return w;
}

static void prepare(fs2_voc* c, cudaStream_t s) {
  FS2_CUDA_OK(cudaSetDevice(c->device));
  for (void* p : c->owned) cudaFree(p);
  c->owned.clear();
  c->prepared = false;
  c->pre = conv_w(c, "conv_pre", UP_INITIAL, N_MEL, 7, s);          // models.py:117-119
  int ch = UP_INITIAL;
  for (int i = 0; i < N_UPS; ++i) {                                   // models.py:122-135
    UpW& u = c->ups[i];
    u.cin = ch; u.cout = ch / 2; u.s = UP_RATE[i];
    const std::string p = "ups." + std::to_string(i);
    const int64_t n = 3LL * u.s * u.cout * u.cin;
    u.w = vkeep(c, n);
    repack_convT_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(
        VW(c, p + ".weight", {u.cin, u.cout, UP_KERNEL[i]}).ptr, u.cin, u.cout, UP_KERNEL[i], u.s, u.w,
        c->math_mode == FS2_MATH_BF16 ? reinterpret_cast<__nv_bfloat16*>(u.w) : nullptr);
    FS2_LAUNCHED();
    u.b = vkeep(c, (size_t)u.s * u.cout);
    tile_bias_kernel<<<(u.s * u.cout + 255) / 256, 256, 0, s>>>(VW(c, p + ".bias", {u.cout}).ptr, u.cout, u.s, u.b);
    FS2_LAUNCHED();
    ch /= 2;
    for (int j = 0; j < N_RES; ++j) {                                 // models.py:137-143
      const std::string rp = "resblocks." + std::to_string(i * N_RES + j);
      for (int m = 0; m < 3; ++m) {
        c->res1[i][j][m] = conv_w(c, rp + ".convs1." + std::to_string(m), ch, ch, RES_KERNEL[j], s);
        c->res2[i][j][m] = conv_w(c, rp + ".convs2." + std::to_string(m), ch, ch, RES_KERNEL[j], s);
      }
    }
  }
  c->post_w = VW(c, "conv_post.weight", {1, ch, POST_K}).ptr;          // models.py:145
  c->post_b = VW(c, "conv_post.bias", {1}).ptr;
  FS2_CUDA_OK(cudaStreamSynchronize(s));
  c->prepared = true;
}

static void forward(fs2_voc* c, cudaStream_t s, const float* mel, int64_t sb, int64_t sc, int64_t st, int B, int T,
                    const int64_t* mel_lens, float* wav) {
  require(c->prepared, FS2_ERR_STATE, "fs2_voc_forward called before fs2_voc_prepare");
  require(mel && wav && B > 0 && B <= 65535 && T > 0, FS2_ERR_INVALID, "bad vocoder argument");
  FS2_CUDA_OK(cudaSetDevice(c->device));
  g_launches = 0;
  const int64_t bound = (int64_t)VOC_GAP + (int64_t)B * (T + VOC_GAP);
  RowSide& sd = c->side;
  ensure_side(sd, B, 0, s);
  if (c->lens64 == nullptr) regrow(c->lens64, 65536, s);
  if (c->h_totals == nullptr) FS2_CUDA_OK(cudaMallocHost(reinterpret_cast<void**>(&c->h_totals), 4 * sizeof(int64_t)));
  FS2_CUDA_OK(cudaMemsetAsync(c->status, 0, sizeof(int32_t), s));
  FS2_CUDA_OK(cudaMemsetAsync(wav, 0, (size_t)B * T * HOP * sizeof(float), s));
  int64_t used = bound;
  if (mel_lens == nullptr) {   // every frame of the padded batch is data, exactly as vocoder(mels) treats it
    fill_i64_kernel<<<(B + 255) / 256, 256, 0, s>>>(c->lens64, B, T);
    FS2_LAUNCHED();
    mel_lens = c->lens64;
    layout_scan_kernel<int64_t><<<1, 1024, 0, s>>>(mel_lens, B, VOC_GAP, T, T, sd.starts, sd.lens, sd.totals, c->status);
    FS2_LAUNCHED();
  } else {
    // ragged batch: the workspace (32 KB per mel row at the audio-rate stages) is sized from the rows actually in use,
    // which costs one small read-back, instead of from batch * n_frames
    layout_scan_kernel<int64_t><<<1, 1024, 0, s>>>(mel_lens, B, VOC_GAP, T, T, sd.starts, sd.lens, sd.totals, c->status);
    FS2_LAUNCHED();
    FS2_CUDA_OK(cudaMemcpyAsync(c->h_totals, sd.totals, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    FS2_CUDA_OK(cudaMemcpyAsync(c->h_totals + 1, c->status, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    FS2_CUDA_OK(cudaStreamSynchronize(s));
    require(*reinterpret_cast<int32_t*>(c->h_totals + 1) == 0, FS2_ERR_INVALID, "mel_lens outside [0, n_frames]");
    used = c->h_totals[0];
  }
  require(used * HOP < (1LL << 31), FS2_ERR_INVALID, "vocoder batch too large (2^31 audio rows)");
  const int rows = round_up((int)used, 128);
  ensure_side(sd, B, rows, s);
  if (rows > c->rows_alloc) {
    regrow(c->melp, (size_t)rows * N_MEL, s);
    regrow(c->a0, (size_t)rows * UP_INITIAL, s);
    for (auto& b : c->buf) regrow(b, (size_t)rows * 8192, s);
    c->rows_alloc = rows;
  }
  row_meta_kernel<<<(rows + 255) / 256, 256, 0, s>>>(sd.starts, sd.lens, B, VOC_GAP, T, nullptr, rows, sd.utt, sd.vpos, sd.room,
                                                     sd.slot);
  FS2_LAUNCHED();
  const bool bf = c->math_mode == FS2_MATH_BF16;   // every activation buffer then holds bf16 behind its float*
  pack_mel_kernel<<<(rows + 7) / 8, 256, 0, s>>>(mel, sb, sc, st, sd.meta(), sd.lens, rows, c->melp,
                                                 bf ? reinterpret_cast<__nv_bfloat16*>(c->melp) : nullptr);
  FS2_LAUNCHED();
  const int32_t* live = reinterpret_cast<const int32_t*>(sd.totals);

  auto conv = [&](const float* A, int rows_now, int shift, const float* W, const float* bias, int taps, int dil, int K, int N,
                  int act, const float* residual, int act2, float* C, int ldc) {
    ConvGemmArgs a{};
    a.A = A; a.lda = K; a.rows = rows_now; a.W = W; a.bias = bias; a.taps = taps; a.dil = dil; a.pad = dil * (taps - 1) / 2;
    a.K = K; a.N = N; a.act = act; a.slope = SLOPE; a.residual = residual; a.ldr = N; a.res_inv_lrelu = residual != nullptr;
    a.act2 = act2; a.row_vpos = sd.vpos; a.row_room = sd.room; a.extra = 0; a.mask_shift = shift;
    a.live_rows = live;
    if (bf) { a.a_bf16 = 1; a.res_bf16 = 1; a.C2 = C; a.ldc2 = ldc; } else { a.C = C; a.ldc = ldc; }
    tc2::launch(a, s);
  };

  // conv_pre (models.py:149) with the leaky ReLU of the first upsampling stage (:151) applied on the way out
  conv(c->melp, rows, 0, c->pre.w, c->pre.b, 7, 1, N_MEL, UP_INITIAL, ACT_LRELU, nullptr, ACT_NONE, c->a0, UP_INITIAL);
  float *U = c->buf[0], *T1 = c->buf[1], *P[2] = {c->buf[2], c->buf[3]}, *R[3] = {c->buf[4], c->buf[5], c->buf[6]};
  const float* a_in = c->a0;
  int rows_now = rows, shift = 0;
  for (int i = 0; i < N_UPS; ++i) {
    const UpW& u = c->ups[i];
    // ConvTranspose1d (models.py:152) as a 3-tap GEMM over the input-rate rows; output = lrelu(x) for the ResBlocks
    const int n_total = u.s * u.cout;
    for (int n0 = 0; n0 < n_total; n0 += tc2::MAX_N) {
      // columns [n0, n0 + n) of every tap: the weight of tap t starts at row t * n_total, so a column slice needs its
      // own launch only when the whole N does not fit the kernel's bias/parameter stage (never with MAX_N = 2048)
      require(n_total <= tc2::MAX_N, FS2_ERR_UNSUPPORTED, "upsampling GEMM wider than the compiled parameter stage");
      conv(a_in, rows_now, shift, u.w, u.b, 3, 1, u.cin, n_total, ACT_LRELU, nullptr, ACT_NONE, U, n_total);
    }
    rows_now *= u.s;
    shift += u.s == 8 ? 3 : 1;
    const int ch = u.cout;
    for (int j = 0; j < N_RES; ++j) {            // three ResBlocks from the same input (models.py:156-161)
      const float* cur = U;
      for (int m = 0; m < 3; ++m) {              // models.py:97-103
        const ConvW &w1 = c->res1[i][j][m], &w2 = c->res2[i][j][m];
        conv(cur, rows_now, shift, w1.w, w1.b, w1.k, RES_DIL[m], ch, ch, ACT_LRELU, nullptr, ACT_NONE, T1, ch);
        float* out = m == 2 ? R[j] : P[m & 1];
        conv(T1, rows_now, shift, w2.w, w2.b, w2.k, 1, ch, ch, ACT_NONE, cur, m == 2 ? ACT_NONE : ACT_LRELU, out, ch);
        cur = out;
      }
    }
    if (i + 1 < N_UPS) {
      const int64_t n4 = (int64_t)rows_now * ch / 4;
      if (bf)
        sum3_lrelu_bf16_kernel<<<148 * 8, 256, 0, s>>>(reinterpret_cast<const uint4*>(R[0]), reinterpret_cast<const uint4*>(R[1]),
                                                       reinterpret_cast<const uint4*>(R[2]), n4 / 2, reinterpret_cast<uint4*>(T1));
      else
        sum3_lrelu_kernel<<<148 * 8, 256, 0, s>>>(reinterpret_cast<const float4*>(R[0]), reinterpret_cast<const float4*>(R[1]),
                                                  reinterpret_cast<const float4*>(R[2]), n4, reinterpret_cast<float4*>(T1));
      FS2_LAUNCHED();
      a_in = T1;
      std::swap(T1, P[0]);   // T1 now feeds the next upsampling: the next stage uses another scratch buffer
    }
  }
  if (bf) post_kernel<true><<<rows, 256, 0, s>>>(R[0], R[1], R[2], (int64_t)rows_now, c->post_w, c->post_b, sd.meta(), sd.starts, T, wav);
  else post_kernel<false><<<rows, 256, 0, s>>>(R[0], R[1], R[2], (int64_t)rows_now, c->post_w, c->post_b, sd.meta(), sd.starts, T, wav);
  FS2_LAUNCHED();
  c->last_launches = g_launches;
}

template <typename F>
static int vguarded(fs2_voc* c, F&& f) {
  try {
    f();
    return FS2_OK;
  } catch (const Error& e) {
    if (c) c->err = e.msg; else g_create_error = e.msg;
    return e.code;
  } catch (const std::exception& e) {
    if (c) c->err = e.what(); else g_create_error = e.what();
    return FS2_ERR_INVALID;
  }
}

}  // namespace voc
}  // namespace fs2

extern "C" {

int fs2_voc_create(int device, int math_mode, fs2_voc** out) {
  return fs2::voc::vguarded(nullptr, [&] {
    require(out != nullptr, FS2_ERR_INVALID, "null argument");
    require(math_mode == FS2_MATH_TF32 || math_mode == FS2_MATH_BF16, FS2_ERR_UNSUPPORTED, "unknown math_mode");
    int n_dev = 0;
    FS2_CUDA_OK(cudaGetDeviceCount(&n_dev));
    require(device >= 0 && device < n_dev, FS2_ERR_INVALID, "no such CUDA device");
    FS2_CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop{};
    FS2_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    require(prop.major == 10, FS2_ERR_UNSUPPORTED, std::string("libfs2b200 is built for sm_100a (B200) only; device is ") + prop.name);
    auto* c = new fs2_voc();
    c->device = device;
    c->math_mode = math_mode;
    c->status = dalloc<int32_t>(1);
    *out = c;
  });
}

void fs2_voc_destroy(fs2_voc* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto& kv : c->raw) cudaFree(kv.second.ptr);
  for (void* p : c->owned) cudaFree(p);
  RowSide* sd = &c->side;
  cudaFree(sd->starts); cudaFree(sd->lens); cudaFree(sd->utt); cudaFree(sd->vpos); cudaFree(sd->room); cudaFree(sd->slot);
  cudaFree(sd->totals); cudaFree(sd->work); cudaFree(sd->work_count);
  cudaFree(c->lens64); cudaFree(c->melp); cudaFree(c->a0); cudaFree(c->status);
  if (c->h_totals) cudaFreeHost(c->h_totals);
  for (float* b : c->buf) cudaFree(b);
  delete c;
}

const char* fs2_voc_last_error(const fs2_voc* c) { return c ? c->err.c_str() : fs2_last_error(nullptr); }

int fs2_voc_set_weight(fs2_voc* c, const char* key, const void* dev_ptr, const int64_t* shape, int ndim) {
  if (!c) return FS2_ERR_INVALID;
  return fs2::voc::vguarded(c, [&] {
    require(key && dev_ptr && shape && ndim >= 1 && ndim <= 3, FS2_ERR_INVALID, "bad fs2_voc_set_weight argument");
    FS2_CUDA_OK(cudaSetDevice(c->device));
    DevTensor t;
    t.numel = 1;
    for (int i = 0; i < ndim; ++i) {
      require(shape[i] > 0, FS2_ERR_INVALID, std::string("non-positive dimension in ") + key);
      t.shape.push_back(shape[i]);
      t.numel *= shape[i];
    }
    t.ptr = dalloc<float>(t.numel);
    FS2_CUDA_OK(cudaDeviceSynchronize());   // see fs2_set_weight
    FS2_CUDA_OK(cudaMemcpy(t.ptr, dev_ptr, t.numel * sizeof(float), cudaMemcpyDeviceToDevice));
    auto it = c->raw.find(key);
    if (it != c->raw.end()) cudaFree(it->second.ptr);
    c->raw[key] = t;
    c->prepared = false;
  });
}

int fs2_voc_prepare(fs2_voc* c, fs2_stream stream) {
  if (!c) return FS2_ERR_INVALID;
  return fs2::voc::vguarded(c, [&] { fs2::voc::prepare(c, static_cast<cudaStream_t>(stream)); });
}

int fs2_voc_forward(fs2_voc* c, fs2_stream stream, const float* mel, int64_t stride_b, int64_t stride_c, int64_t stride_t,
                    int batch, int n_frames, const int64_t* mel_lens, float* wav) {
  if (!c) return FS2_ERR_INVALID;
  return fs2::voc::vguarded(c, [&] {
    fs2::voc::forward(c, static_cast<cudaStream_t>(stream), mel, stride_b, stride_c, stride_t, batch, n_frames, mel_lens, wav);
  });
}

int fs2_voc_last_launch_count(const fs2_voc* c) { return c ? c->last_launches : 0; }

}  // extern "C"
