// Implicit-GEMM Conv1d / Linear over packed token-major rows on the LEGACY tensor-core path
// (mma.sync.m16n8k8 TF32, cp.async staging).  This engine is the bring-up path and the
// cross-check for the tcgen05 engine (gemm_tcgen05.cuh); both implement ConvGemmArgs.
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
  // Fused LayerNorm epilogue (persistent tcgen05 engine only; requires N == 256):
  //   y = LayerNorm(act(acc + bias) + residual) * ln_gamma + ln_beta, non-live rows -> 0,
  //   optional head: head_out[slot ? slot[row] : row] = y . head_w + head_b[0] on live rows.
  // C may be nullptr when only the head output is wanted.
  const float* ln_gamma;
  const float* ln_beta;
  const float* head_w;
  const float* head_b;
  float* head_out;
  const int32_t* slot;
  // BF16 operand mode (persistent tcgen05 engine only): A is bf16 [rows, lda] and W is bf16 [taps][N][K]
  // (both pointers reinterpret the float* fields); bias, residual and C stay fp32.  C2, when set, receives
  // a bf16 copy of the output [rows, ldc2] (the A operand of the next contraction); C may then be nullptr.
  int a_bf16;
  void* C2;
  int ldc2;
  // Vocoder extensions (persistent tcgen05 engine only; zero-initialised = off):
  int dil;            // dilation: tap t reads row r + t*dil - pad (0 is treated as 1; pad is in rows)
  int mask_shift;     // the row mask is looked up at row >> mask_shift (rows upsampled 2^shift times share a frame);
                      // live_rows, when set, counts rows at that coarser rate too
  float slope;        // ACT_LRELU slope
  int act2;           // activation applied AFTER the residual add (ACT_NONE / ACT_LRELU)
  int res_inv_lrelu;  // the residual buffer holds lrelu(x): recover x = y >= 0 ? y : y / slope before adding
  int res_bf16;       // the residual buffer is bf16 [rows, ldr] (vocoder bf16 mode), not fp32
  long long* trace;   // bring-up only: CTA 0 writes globaltimer stamps of its phases (nullptr in normal operation)
};

namespace mma {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3, THREADS = 256;
constexpr int SMEM_BYTES = STAGES * (BM + BN) * BK * 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  int bytes = pred ? 16 : 0;  // src-size 0 => the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// smem tiles are [rows][32 floats] with the 16-byte chunk index XOR-swizzled by (row & 7):
// conflict-free for both the cp.async stores and the mma fragment loads.
__device__ __forceinline__ int swz(int row, int col) { return row * BK + ((((col >> 2) ^ (row & 7)) << 2) | (col & 3)); }

__global__ void __launch_bounds__(THREADS, 2) conv_gemm_kernel(ConvGemmArgs p) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                         // [STAGES][BM*BK]
  float* Bs = smem + STAGES * BM * BK;      // [STAGES][BN*BK]

  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  if (p.live_rows != nullptr && m0 >= *p.live_rows) return;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;
  const int g = lane >> 2, t = lane & 3;

  const int kchunks = (p.K + BK - 1) / BK;
  const int iters = p.taps * kchunks;

  auto load_stage = [&](int it, int stage) {
    const int tap = it / kchunks, kc = it - tap * kchunks;
    const int k0 = kc * BK;
    float* as = As + stage * BM * BK;
    float* bs = Bs + stage * BN * BK;
    // A: 128 rows x 8 chunks of 16 B = 1024 chunks, 4 per thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = tid + i * THREADS;
      const int r = c >> 3, ch = c & 7;
      const int grow = m0 + r + tap - p.pad;
      const int k = k0 + ch * 4;
      const bool ok = grow >= 0 && grow < p.rows && k < p.K;
      const float* src = ok ? p.A + (size_t)grow * p.lda + k : p.A;
      cp_async16(smem_u32(as + r * BK + ((ch ^ (r & 7)) << 2)), src, ok);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = tid + i * THREADS;
      const int r = c >> 3, ch = c & 7;
      const int n = n0 + r;
      const int k = k0 + ch * 4;
      const bool ok = n < p.N && k < p.K;
      const float* src = ok ? p.W + ((size_t)tap * p.N + n) * p.K + k : p.W;
      cp_async16(smem_u32(bs + r * BK + ((ch ^ (r & 7)) << 2)), src, ok);
    }
  };

  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < iters) load_stage(s, s);
    cp_async_commit();
  }

  for (int it = 0; it < iters; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = it + STAGES - 1;
      if (nxt < iters) load_stage(nxt, nxt % STAGES);
      cp_async_commit();
    }
    const float* as = As + (it % STAGES) * BM * BK;
    const float* bs = Bs + (it % STAGES) * BN * BK;
#pragma unroll
    for (int kk = 0; kk < BK / 8; ++kk) {
      uint32_t af[4][4], bf[4][2];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        const int r = wm + mi * 16 + g;
        af[mi][0] = to_tf32(as[swz(r, kk * 8 + t)]);
        af[mi][1] = to_tf32(as[swz(r + 8, kk * 8 + t)]);
        af[mi][2] = to_tf32(as[swz(r, kk * 8 + t + 4)]);
        af[mi][3] = to_tf32(as[swz(r + 8, kk * 8 + t + 4)]);
      }
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int r = wn + ni * 8 + g;
        bf[ni][0] = __float_as_uint(bs[swz(r, kk * 8 + t)]);      // weights are TF32-exact already
        bf[ni][1] = __float_as_uint(bs[swz(r, kk * 8 + t + 4)]);
      }
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) mma_tf32(acc[mi][ni], af[mi], bf[ni]);
    }
  }
  cp_async_wait<0>();

  // epilogue: bias, activation, residual, row mask; float2 stores (c0,c1 are adjacent columns)
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int row = m0 + wm + mi * 16 + g + half * 8;
      if (row >= p.rows) continue;
      bool live = true;
      if (p.row_vpos != nullptr) live = row_live(p.row_vpos[row], p.row_room[row], p.extra);
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int col = n0 + wn + ni * 8 + 2 * t;
        if (col >= p.N) continue;   // N is even, so col+1 < N too
        float v0 = acc[mi][ni][half * 2 + 0] + p.bias[col];
        float v1 = acc[mi][ni][half * 2 + 1] + p.bias[col + 1];
        if (p.act == ACT_RELU) {
          v0 = fmaxf(v0, 0.f);
          v1 = fmaxf(v1, 0.f);
        } else if (p.act == ACT_TANH) {
          v0 = tanhf(v0);
          v1 = tanhf(v1);
        }
        if (p.residual != nullptr) {
          const float2 r2 = *reinterpret_cast<const float2*>(p.residual + (size_t)row * p.ldr + col);
          v0 += r2.x;
          v1 += r2.y;
        }
        if (!live) v0 = v1 = 0.f;
        *reinterpret_cast<float2*>(p.C + (size_t)row * p.ldc + col) = make_float2(v0, v1);
      }
    }
  }
}

inline void launch(const ConvGemmArgs& a, cudaStream_t stream) {
  require(a.K % 4 == 0 && a.lda % 4 == 0, FS2_ERR_INVALID, "conv_gemm: K and lda must be multiples of 4");
  require(a.N % 2 == 0 && a.ldc % 2 == 0, FS2_ERR_INVALID, "conv_gemm: N and ldc must be even");
  static bool configured[64] = {};
  int dev = 0;
  FS2_CUDA_OK(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    FS2_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured[dev & 63] = true;
  }
  dim3 grid((a.rows + BM - 1) / BM, (a.N + BN - 1) / BN);
  if (grid.x == 0) return;
  conv_gemm_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(a);
  FS2_LAUNCHED();
}

}  // namespace mma
}  // namespace fs2
