// Implicit-GEMM Conv1d / Linear on the 5th-generation tensor cores (sm_100a):
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared memory ring -> tcgen05.mma kind::tf32 with
//   the fp32 accumulator tile in TMEM -> tcgen05.ld epilogue (bias, ReLU/tanh, residual, row mask).
// A conv tap is a row offset of the SAME activation matrix, so tap t of K-chunk c is simply the
// TMA box at coordinates (32c, m0 + t - pad); rows outside the tensor (and the K tail of the
// 80-channel PostNet input) are zero-filled by the TMA unit.  Weights are [taps][N][K], i.e. a
// 2-D [taps*N, K] tensor whose box for (tap, n-tile) starts at row tap*N + n0.
//
// Warp roles (192 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator +
// MMA issuer (one elected lane), warps 2..5 = epilogue (TMEM lane quarter = warp % 4; a thread
// owns one output row and moves it with 16-byte stores).
#pragma once

#include "common.cuh"
#include "gemm_mma.cuh"
#include "tc_ptx.cuh"

namespace fs2 {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;            // fp32/tf32 elements per 128-byte swizzle row
constexpr int STAGES = 4;
constexpr int THREADS = 192;
constexpr int A_STAGE_BYTES = BM * 128;

template <int BN>
struct Smem {
  static constexpr int B_STAGE_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int BIAS_OFF = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFF = BIAS_OFF + 1024;                   // BN <= 256 floats
  static constexpr int TOTAL = BAR_OFF + 128 + 1024;                // + alignment slack
  static constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static_assert(B_STAGE_BYTES % 1024 == 0, "every stage must keep 1024-byte alignment (SWIZZLE_128B)");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128");
};

template <int BN>
__global__ void __launch_bounds__(THREADS, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, ConvGemmArgs p) {
  using S = Smem<BN>;
  extern __shared__ uint8_t smem_raw[];
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  if (p.live_rows != nullptr && m0 >= *p.live_rows) return;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* bias_s = reinterpret_cast<float*>(smem + S::BIAS_OFF);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kchunks = (p.K + BK - 1) / BK;
  const int iters = p.taps * kchunks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)S::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], S::STAGE_BYTES);
        const int tap = it / kchunks, kc = it - tap * kchunks;
        uint8_t* a_s = smem + s * S::STAGE_BYTES;
        tma_load_2d(a_s, &tmA, kc * BK, m0 + tap - p.pad, &full[s]);
        tma_load_2d(a_s + A_STAGE_BYTES, &tmW, kc * BK, tap * p.N + n0, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(BM, BN);
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint8_t* a_s = smem + s * S::STAGE_BYTES;
        const uint64_t da = umma_desc(a_s), db = umma_desc(a_s + A_STAGE_BYTES);
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk) {
          // advance 8 tf32 = 32 bytes along K inside the swizzle row: +2 in the (addr >> 4) field
          umma_tf32(tmem_base, da + 2 * kk, db + 2 * kk, idesc, (it | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&empty[s]);  // frees the stage when these MMAs have read it
      }
      umma_commit(tmem_full);    // accumulator complete
    }
  } else {
    // ---- epilogue: 4 warps, thread = one accumulator row (TMEM lane)
    const int et = threadIdx.x - 64;
    for (int i = et; i < BN; i += 128) bias_s[i] = (n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
    asm volatile("bar.sync 1, 128;\n" ::: "memory");
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    const bool in_range = row < p.rows;
    bool live = in_range;
    if (in_range && p.row_vpos != nullptr) live = row_live(p.row_vpos[row], p.row_room[row], p.extra);
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* crow = p.C + (size_t)row * p.ldc + n0;
    const float* rrow = p.residual != nullptr ? p.residual + (size_t)row * p.ldr + n0 : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      const int width = (BN - c0) >= 32 ? 32 : 16;
      if (width == 32) tmem_ld32(taddr + c0, v); else tmem_ld16(taddr + c0, v);
      if (!in_range) continue;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (j >= width) break;
        float4 o;
        float* op = reinterpret_cast<float*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float x = v[j + e] + bias_s[c0 + j + e];
          if (p.act == ACT_RELU) x = fmaxf(x, 0.f);
          else if (p.act == ACT_TANH) x = tanhf(x);
          op[e] = x;
        }
        if (rrow != nullptr) {
          const float4 r4 = *reinterpret_cast<const float4*>(rrow + c0 + j);
          o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
        }
        if (!live) o = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(crow + c0 + j) = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"((uint32_t)S::TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------- host side
template <int BN>
inline void launch_bn(const ConvGemmArgs& a, cudaStream_t stream) {
  using S = Smem<BN>;
  static bool configured[64] = {};
  int dev = 0;
  FS2_CUDA_OK(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    FS2_CUDA_OK(cudaFuncSetAttribute(conv_gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured[dev & 63] = true;
  }
  const CUtensorMap tmA = make_map(a.A, a.rows, a.K, a.lda, BM, /*round_tf32=*/true, false);
  const CUtensorMap tmW = make_map(a.W, (int64_t)a.taps * a.N, a.K, a.K, BN, false, true);
  dim3 grid(a.N / BN, (a.rows + BM - 1) / BM);
  conv_gemm_tc_kernel<BN><<<grid, THREADS, S::TOTAL, stream>>>(tmA, tmW, a);
  FS2_LAUNCHED();
}

inline void launch(const ConvGemmArgs& a, int math_mode, cudaStream_t stream) {
  require(math_mode == FS2_MATH_TF32, FS2_ERR_UNSUPPORTED, "tcgen05 engine: only FS2_MATH_TF32 is built");
  require(a.K % 4 == 0 && a.lda % 4 == 0 && a.ldc % 4 == 0 && (a.residual == nullptr || a.ldr % 4 == 0), FS2_ERR_INVALID,
          "tcgen05 conv_gemm: K and leading dimensions must be multiples of 4 (16-byte rows)");
  require((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.W) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(a.C) & 15) == 0, FS2_ERR_INVALID, "tcgen05 conv_gemm: pointers must be 16-byte aligned");
  if (a.rows <= 0) return;
  if (a.N % 256 == 0) launch_bn<256>(a, stream);
  else if (a.N == 80) launch_bn<80>(a, stream);
  else if (a.N % 128 == 0) launch_bn<128>(a, stream);
  else if (a.N % 64 == 0) launch_bn<64>(a, stream);
  else if (a.N % 16 == 0 && a.N <= 256) {
    throw Error(FS2_ERR_UNSUPPORTED, "tcgen05 conv_gemm: N = " + std::to_string(a.N) + " has no compiled tile");
  } else {
    throw Error(FS2_ERR_UNSUPPORTED, "tcgen05 conv_gemm: N must be a multiple of 16");
  }
}

}  // namespace tc
}  // namespace fs2
