// Operator arguments of the implicit-GEMM Conv1d / Linear kernel (gemm_tc2.cuh) and its epilogue options.
#pragma once

#include "common.cuh"

namespace fs2 {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2, ACT_LRELU = 3 /* leaky ReLU, slope = ConvGemmArgs::slope */ };

struct ConvGemmArgs {
  const float* A;        // [rows, lda] activations, token-major
  int lda;
  int rows;              // rows of A / C that exist (reads outside are zero)
  const float* W;        // [taps][N][K], K contiguous ("K-major B operand"), TF32-pre-rounded
  const float* bias;     // [N]
  int taps, pad;         // out[r] = sum_t A[r + t - pad] * W[t]
  int K, N;
  int act;
  const float* residual; // [rows, ldr] or nullptr, added after the activation
  int ldr;
  const int32_t* row_vpos;  // nullptr = no row mask
  const int32_t* row_room;
  int extra;
  float* C;
  int ldc;
  const int32_t* live_rows;  // optional device scalar: tiles at or beyond *live_rows exit early
  // Fused LayerNorm epilogue (requires N == 256):
  //   y = LayerNorm(act(acc + bias) + residual) * ln_gamma + ln_beta, non-live rows -> 0,
  //   optional head: head_out[slot ? slot[row] : row] = y . head_w + head_b[0] on live rows.
  // C may be nullptr when only the head output is wanted.
  const float* ln_gamma;
  const float* ln_beta;
  const float* head_w;
  const float* head_b;
  float* head_out;
  const int32_t* slot;
  // Fused LayerNorm only: y[row] = (y[row] + post_a[u]) + post_b[u] with u = post_utt[row], applied AFTER the row mask on
  // the rows that are live with post_extra reserved rows -- the speaker / emotion conditioning of
  // model/fastspeech2.py:101-110 on the last encoder layer's output (the reference adds the vectors on padding rows too).
  const float* post_a;
  const float* post_b;
  const int32_t* post_utt;
  int post_extra;
  // BF16 operand mode: A is bf16 [rows, lda] and W is bf16 [taps][N][K] (both pointers reinterpret the float*
  // fields); bias, residual and C stay fp32.  C2, when set, receives a bf16 copy of the output [rows, ldc2]
  // (the A operand of the next contraction); C may then be nullptr.
  int a_bf16;
  void* C2;
  int ldc2;
  // Vocoder extensions (zero-initialised = off):
  int dil;            // dilation: tap t reads row r + t*dil - pad (0 is treated as 1; pad is in rows)
  int mask_shift;     // the row mask is looked up at row >> mask_shift (rows upsampled 2^shift times share a frame);
                      // live_rows, when set, counts rows at that coarser rate too
  float slope;        // ACT_LRELU slope
  int act2;           // activation applied AFTER the residual add (ACT_NONE / ACT_LRELU)
  int res_inv_lrelu;  // the residual buffer holds lrelu(x): recover x = y >= 0 ? y : y / slope before adding
  int res_bf16;       // the residual buffer is bf16 [rows, ldr] (vocoder bf16 mode), not fp32
  // FS2_MATH_TF32X3 (split-operand "3xTF32"): A is [rows, 2K] = [hi | lo] (both TF32-exact, written by split_tf32_kernel),
  // W is [2][taps][N][K] = hi block then lo block; the K loop runs three terms  A_hi W_hi + A_lo W_hi + A_hi W_lo.
  int terms;          // 0 / 1 = plain, 3 = split operands
  float* splitk_ws;   // optional workspace (tc2::SPLITK_WS_BYTES) that allows the K-split form for single-row-tile launches
  long long* trace;   // bring-up only: CTA 0 writes globaltimer stamps of its phases (nullptr in normal operation)
};

}  // namespace fs2
