// Varlen 2-head self-attention over packed rows (transformer/SubLayers.py:42-52,
// transformer/Modules.py:14-25): softmax(Q K^T / sqrt(128)) V per utterance, keys masked by the
// utterance length instead of a [2B,S,S] mask, scores never materialised (online softmax).
// TF32 mma.sync tiles, fp32 softmax state.  Block = 64 queries of one (utterance, head);
// K/V tiles of 32 keys are double-buffered with cp.async.
#pragma once

#include "common.cuh"
#include "gemm_mma.cuh"

namespace fs2 {
namespace attn {

constexpr int BQ = 64, BKV = 32, THREADS = 128;
constexpr int P_STRIDE = BKV + 4;
constexpr int SMEM_FLOATS = BQ * D_HEAD + 2 * 2 * BKV * D_HEAD + 4 * 16 * P_STRIDE;
constexpr int SMEM_BYTES = SMEM_FLOATS * 4;
constexpr int LDQKV = 3 * D_MODEL;

// [rows][128] tiles: 16-byte chunk c of row r is stored at chunk c ^ f(r).
__device__ __forceinline__ int idx_qk(int r, int c) { return r * D_HEAD + ((((c >> 2) ^ (r & 7)) << 2) | (c & 3)); }
__device__ __forceinline__ int idx_v(int r, int c) { return r * D_HEAD + ((((c >> 2) ^ ((r & 3) << 1)) << 2) | (c & 3)); }

__global__ void __launch_bounds__(THREADS) attention_kernel(const float* __restrict__ qkv,
                                                            const int32_t* __restrict__ starts,
                                                            const int32_t* __restrict__ lens,
                                                            float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;                       // [64][128]
  float* Ks = Qs + BQ * D_HEAD;           // [2][32][128]
  float* Vs = Ks + 2 * BKV * D_HEAD;      // [2][32][128]
  float* Ps = Vs + 2 * BKV * D_HEAD;      // [4][16][36]

  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int len = lens[b];
  if (q0 >= len) return;
  const int row0 = starts[b];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const float* qbase = qkv + (size_t)row0 * LDQKV + h * D_HEAD;
  const float* kbase = qbase + D_MODEL;
  const float* vbase = qbase + 2 * D_MODEL;

  // Q tile: 64 rows x 32 chunks
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = tid + i * THREADS;
    const int r = c >> 5, ch = c & 31;
    const bool ok = q0 + r < len;
    mma::cp_async16(mma::smem_u32(Qs + r * D_HEAD + ((ch ^ (r & 7)) << 2)),
                    ok ? qbase + (size_t)(q0 + r) * LDQKV + ch * 4 : qkv, ok);
  }
  auto load_kv = [&](int kt, int stage) {
    float* ks = Ks + stage * BKV * D_HEAD;
    float* vs = Vs + stage * BKV * D_HEAD;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = tid + i * THREADS;
      const int r = c >> 5, ch = c & 31;
      const int key = kt * BKV + r;
      const bool ok = key < len;
      mma::cp_async16(mma::smem_u32(ks + r * D_HEAD + ((ch ^ (r & 7)) << 2)),
                      ok ? kbase + (size_t)key * LDQKV + ch * 4 : qkv, ok);
      mma::cp_async16(mma::smem_u32(vs + r * D_HEAD + ((ch ^ ((r & 3) << 1)) << 2)),
                      ok ? vbase + (size_t)key * LDQKV + ch * 4 : qkv, ok);
    }
  };
  const int n_tiles = (len + BKV - 1) / BKV;
  load_kv(0, 0);
  mma::cp_async_commit();

  float o[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;  // rows g and g+8 of this warp's 16
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)D_HEAD);
  float* pw = Ps + warp * 16 * P_STRIDE;
  const int qr = warp * 16 + g;

  for (int kt = 0; kt < n_tiles; ++kt) {
    if (kt + 1 < n_tiles) load_kv(kt + 1, (kt + 1) & 1);
    mma::cp_async_commit();
    mma::cp_async_wait<1>();
    __syncthreads();
    const float* ks = Ks + (kt & 1) * BKV * D_HEAD;
    const float* vs = Vs + (kt & 1) * BKV * D_HEAD;

    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < D_HEAD / 8; ++kk) {
      uint32_t a[4];
      a[0] = mma::to_tf32(Qs[idx_qk(qr, kk * 8 + t)]);
      a[1] = mma::to_tf32(Qs[idx_qk(qr + 8, kk * 8 + t)]);
      a[2] = mma::to_tf32(Qs[idx_qk(qr, kk * 8 + t + 4)]);
      a[3] = mma::to_tf32(Qs[idx_qk(qr + 8, kk * 8 + t + 4)]);
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        uint32_t bb[2];
        bb[0] = mma::to_tf32(ks[idx_qk(ni * 8 + g, kk * 8 + t)]);
        bb[1] = mma::to_tf32(ks[idx_qk(ni * 8 + g, kk * 8 + t + 4)]);
        mma::mma_tf32(s[ni], a, bb);
      }
    }
    // scale into the log2 domain, mask keys beyond the utterance, online softmax
    float mx_lo = m_lo, mx_hi = m_hi;
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const int key = kt * BKV + ni * 8 + 2 * t;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool valid = key + (e & 1) < len;
        s[ni][e] = valid ? s[ni][e] * scale_log2 : -INFINITY;
      }
      mx_lo = fmaxf(mx_lo, fmaxf(s[ni][0], s[ni][1]));
      mx_hi = fmaxf(mx_hi, fmaxf(s[ni][2], s[ni][3]));
    }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float a_lo = exp2f(m_lo - mx_lo), a_hi = exp2f(m_hi - mx_hi);  // 0 on the first tile
    m_lo = mx_lo;
    m_hi = mx_hi;
    l_lo *= a_lo;
    l_hi *= a_hi;
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const float p0 = exp2f(s[ni][0] - m_lo), p1 = exp2f(s[ni][1] - m_lo);
      const float p2 = exp2f(s[ni][2] - m_hi), p3 = exp2f(s[ni][3] - m_hi);
      l_lo += p0 + p1;
      l_hi += p2 + p3;
      *reinterpret_cast<float2*>(pw + g * P_STRIDE + ni * 8 + 2 * t) = make_float2(p0, p1);
      *reinterpret_cast<float2*>(pw + (g + 8) * P_STRIDE + ni * 8 + 2 * t) = make_float2(p2, p3);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      o[i][0] *= a_lo;
      o[i][1] *= a_lo;
      o[i][2] *= a_hi;
      o[i][3] *= a_hi;
    }
    __syncwarp();
#pragma unroll
    for (int kk = 0; kk < BKV / 8; ++kk) {
      uint32_t a[4];
      a[0] = mma::to_tf32(pw[g * P_STRIDE + kk * 8 + t]);
      a[1] = mma::to_tf32(pw[(g + 8) * P_STRIDE + kk * 8 + t]);
      a[2] = mma::to_tf32(pw[g * P_STRIDE + kk * 8 + t + 4]);
      a[3] = mma::to_tf32(pw[(g + 8) * P_STRIDE + kk * 8 + t + 4]);
#pragma unroll
      for (int ni = 0; ni < 16; ++ni) {
        uint32_t bb[2];
        bb[0] = mma::to_tf32(vs[idx_v(kk * 8 + t, ni * 8 + g)]);
        bb[1] = mma::to_tf32(vs[idx_v(kk * 8 + t + 4, ni * 8 + g)]);
        mma::mma_tf32(o[ni], a, bb);
      }
    }
    __syncthreads();  // all warps are done with this K/V stage before it is refilled
  }

  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float inv_lo = 1.f / l_lo, inv_hi = 1.f / l_hi;
  float* obase = out + (size_t)row0 * D_MODEL + h * D_HEAD;
  const int r_lo = q0 + qr, r_hi = r_lo + 8;
#pragma unroll
  for (int ni = 0; ni < 16; ++ni) {
    const int col = ni * 8 + 2 * t;
    if (r_lo < len)
      *reinterpret_cast<float2*>(obase + (size_t)r_lo * D_MODEL + col) = make_float2(o[ni][0] * inv_lo, o[ni][1] * inv_lo);
    if (r_hi < len)
      *reinterpret_cast<float2*>(obase + (size_t)r_hi * D_MODEL + col) = make_float2(o[ni][2] * inv_hi, o[ni][3] * inv_hi);
  }
}

inline void launch(const float* qkv, const int32_t* starts, const int32_t* lens, int batch, int max_len,
                   float* out, cudaStream_t stream) {
  if (batch <= 0 || max_len <= 0) return;
  static bool configured[64] = {};
  int dev = 0;
  FS2_CUDA_OK(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    FS2_CUDA_OK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured[dev & 63] = true;
  }
  dim3 grid((max_len + BQ - 1) / BQ, N_HEAD, batch);
  attention_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(qkv, starts, lens, out);
  FS2_LAUNCHED();
}

}  // namespace attn
}  // namespace fs2
