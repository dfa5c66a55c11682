// HiFi-GAN generator (hifigan/models.py:112-174, config hifigan/config.json: V1) on the same persistent tcgen05
// implicit-GEMM engine as the acoustic model -- the step right after FastSpeech2.forward in
// synthesize_chinese_pinyin.py (utils/model.py:74-92 `vocoder_infer`).
//
// Layout: token-major packed rows as everywhere else.  Utterance b owns mel-rate rows [start_b, start_b + len_b)
// separated by VOC_GAP zero rows.  A ConvTranspose1d(k = 2s, stride s, padding s/2) is the 3-tap GEMM
//     out[q, r*Cout + co] = sum_{d in -1..1} x[q - d] . w[:, co, s*d + r + s/2]          (r = output phase)
// whose [rows, s*Cout] row-major output IS the [rows*s, Cout] upsampled tensor, so every stage keeps the layout.  Dilated Conv1d = tap stride in
// rows.  Every epilogue re-zeroes the reserved rows (row mask looked up at mel rate, row >> shift), which is
// exactly the zero padding each utterance sees when the reference runs it alone.
//
// Pre-activations: the generator applies leaky_relu(x, 0.1) BEFORE each conv (models.py:98-103,153-154), and TMA
// cannot transform an operand, so producers store lrelu(x); where the raw x is needed again (the ResBlock
// residual, models.py:102) the consumer's epilogue inverts the (bijective) leaky ReLU.
#pragma once

#include "common.cuh"
#include "gemm_tc2.cuh"
#include "rowops.cuh"

namespace fs2 {
namespace voc {

constexpr int VOC_GAP = 12;        // >= 3 (k=7 at mel rate); x8 after the first upsampling >= 25 (k=11, dilation 5)
constexpr int N_UPS = 4;
constexpr int UP_RATE[N_UPS] = {8, 8, 2, 2};        // hifigan/config.json:11
constexpr int UP_KERNEL[N_UPS] = {16, 16, 4, 4};    // config.json:12
constexpr int UP_INITIAL = 512;                     // config.json:13
constexpr int N_RES = 3;
constexpr int RES_KERNEL[N_RES] = {3, 7, 11};       // config.json:14
constexpr int RES_DIL[3] = {1, 3, 5};               // config.json:15
constexpr float SLOPE = 0.1f;                       // models.py:7
constexpr int HOP = 256;                            // product of the upsample rates

struct ConvW { float* w = nullptr; const float* b = nullptr; int cin = 0, cout = 0, k = 0; };
struct UpW { float* w = nullptr; float* b = nullptr; int cin = 0, cout = 0, s = 0; };

// ConvTranspose1d weight [Cin][Cout][K] (torch layout) -> GEMM form [3 taps][s*Cout][Cin]; tap t reads x[q + t - 1],
// i.e. d = 1 - t, kernel index k = s*d + r + (K - s)/2 (zero when outside [0, K)); operands rounded to TF32.
// out_b != nullptr: the same repack rounded to bf16 (FS2_MATH_BF16 vocoder)
__global__ void repack_convT_kernel(const float* __restrict__ w, int cin, int cout, int K, int s, float* __restrict__ out,
                                    __nv_bfloat16* __restrict__ out_b = nullptr) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per_tap = (int64_t)s * cout * cin;
  if (i >= 3 * per_tap) return;
  const int ci = (int)(i % cin);
  const int n = (int)((i / cin) % ((int64_t)s * cout));
  const int tap = (int)(i / per_tap);
  const int r = n / cout, co = n % cout;
  const int k = s * (1 - tap) + r + (K - s) / 2;
  const float v = (k >= 0 && k < K) ? w[((size_t)ci * cout + co) * K + k] : 0.f;
  if (out_b != nullptr) out_b[i] = __float2bfloat16_rn(v);
  else out[i] = round_tf32(v);
}
__global__ void tile_bias_kernel(const float* __restrict__ b, int cout, int s, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < s * cout) out[i] = b[i % cout];
}
__global__ void fill_i64_kernel(int64_t* p, int n, int64_t v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// mel [B, 80, T] (any strides, in elements) -> packed token-major [rows, 80]; reserved rows are zero.
__global__ void pack_mel_kernel(const float* __restrict__ mel, int64_t sb, int64_t sc, int64_t st, RowMeta meta,
                                const int32_t* __restrict__ lens, int rows, float* __restrict__ out,
                                __nv_bfloat16* __restrict__ out_b = nullptr) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int u = meta.utt[row], vp = meta.vpos[row];
  const bool real = u >= 0 && vp < 0;
  const int t = real ? vp + lens[u] : 0;
  for (int c = lane; c < N_MEL; c += 32) {
    const float v = real ? mel[(int64_t)u * sb + (int64_t)c * sc + (int64_t)t * st] : 0.f;
    if (out_b != nullptr) out_b[(size_t)row * N_MEL + c] = __float2bfloat16_rn(v);
    else out[(size_t)row * N_MEL + c] = v;
  }
}

// x = (r0 + r1) + r2) / 3 (models.py:156-162) -> lrelu(x, 0.1) (models.py:153): the A operand of the next upsampling.
__global__ void sum3_lrelu_kernel(const float4* __restrict__ r0, const float4* __restrict__ r1, const float4* __restrict__ r2,
                                  int64_t n4, float4* __restrict__ y) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = r0[i], b = r1[i], c = r2[i];
    float4 v = make_float4(((a.x + b.x) + c.x) / 3.f, ((a.y + b.y) + c.y) / 3.f, ((a.z + b.z) + c.z) / 3.f,
                           ((a.w + b.w) + c.w) / 3.f);
    v.x = v.x >= 0.f ? v.x : v.x * SLOPE; v.y = v.y >= 0.f ? v.y : v.y * SLOPE;
    v.z = v.z >= 0.f ? v.z : v.z * SLOPE; v.w = v.w >= 0.f ? v.w : v.w * SLOPE;
    y[i] = v;
  }
}

// bf16 activations: eight values per 16-byte access
__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    f[2 * e] = __uint_as_float(w[e] << 16);
    f[2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u);
  }
}
__global__ void sum3_lrelu_bf16_kernel(const uint4* __restrict__ r0, const uint4* __restrict__ r1, const uint4* __restrict__ r2,
                                       int64_t n8, uint4* __restrict__ y) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float a[8], b[8], c[8];
    unpack8(r0[i], a); unpack8(r1[i], b); unpack8(r2[i], c);
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v0 = ((a[2 * e] + b[2 * e]) + c[2 * e]) / 3.f, v1 = ((a[2 * e + 1] + b[2 * e + 1]) + c[2 * e + 1]) / 3.f;
      v0 = v0 >= 0.f ? v0 : v0 * SLOPE;
      v1 = v1 >= 0.f ? v1 : v1 * SLOPE;
      const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
      w[e] = *reinterpret_cast<const uint32_t*>(&h);
    }
    y[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// Last stage fused: x = ((r0 + r1) + r2) / 3 -> leaky_relu(x) with the DEFAULT slope 0.01 (models.py:163, sic) ->
// conv_post (32 -> 1, k = 7, pad 3; models.py:146,164) -> tanh (:165) -> scattered into wav[b, pos] (padded output).
// One block = 256 consecutive audio rows = one mel frame row; thread = one sample.
constexpr int POST_C = 32, POST_K = 7;
template <bool BF>
__global__ void __launch_bounds__(256)
post_kernel(const float* __restrict__ r0, const float* __restrict__ r1, const float* __restrict__ r2, int64_t rows_audio,
            const float* __restrict__ w /*[1][32][7]*/, const float* __restrict__ bias, RowMeta meta,
            const int32_t* __restrict__ starts, int n_frames, float* __restrict__ wav) {
  __shared__ float xs[(256 + POST_K - 1) * (POST_C + 1)];
  __shared__ float ws[POST_K * POST_C];
  const int frame_row = blockIdx.x;
  const int u = meta.utt[frame_row], vp = meta.vpos[frame_row];
  if (u < 0 || vp >= 0) return;                       // reserved row: its 256 samples do not exist
  for (int i = threadIdx.x; i < POST_K * POST_C; i += 256) {   // ws[t][c] = w[0][c][t]
    const int t = i / POST_C, c = i % POST_C;
    ws[i] = w[c * POST_K + t];
  }
  const int64_t base = (int64_t)frame_row * HOP - (POST_K / 2);
  for (int i = threadIdx.x; i < (256 + POST_K - 1) * (POST_C / 4); i += 256) {
    const int rr = i / (POST_C / 4), c4 = i % (POST_C / 4);
    const int64_t r = base + rr;
    float4 v = make_float4(0, 0, 0, 0);
    if (r >= 0 && r < rows_audio) {
      const size_t o = (size_t)r * POST_C + c4 * 4;
      float4 a, b, c;
      if (BF) {   // the three buffers hold bf16: four values = 8 bytes
        auto ldb = [&](const float* base) {
          const uint2 q = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + o);
          return make_float4(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xFFFF0000u), __uint_as_float(q.y << 16),
                             __uint_as_float(q.y & 0xFFFF0000u));
        };
        a = ldb(r0); b = ldb(r1); c = ldb(r2);
      } else {
        a = ld4(r0 + o); b = ld4(r1 + o); c = ld4(r2 + o);
      }
      v = make_float4(((a.x + b.x) + c.x) / 3.f, ((a.y + b.y) + c.y) / 3.f, ((a.z + b.z) + c.z) / 3.f, ((a.w + b.w) + c.w) / 3.f);
      v.x = v.x >= 0.f ? v.x : v.x * 0.01f; v.y = v.y >= 0.f ? v.y : v.y * 0.01f;
      v.z = v.z >= 0.f ? v.z : v.z * 0.01f; v.w = v.w >= 0.f ? v.w : v.w * 0.01f;
    }
    float* d = xs + rr * (POST_C + 1) + c4 * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  float acc = bias[0];
#pragma unroll
  for (int t = 0; t < POST_K; ++t) {
    const float* xr = xs + (threadIdx.x + t) * (POST_C + 1);
#pragma unroll
    for (int c = 0; c < POST_C; ++c) acc = fmaf(xr[c], ws[t * POST_C + c], acc);
  }
  const int64_t pos = (int64_t)(frame_row - starts[u]) * HOP + threadIdx.x;
  wav[(int64_t)u * n_frames * HOP + pos] = tanhf(acc);
}

}  // namespace voc
}  // namespace fs2
