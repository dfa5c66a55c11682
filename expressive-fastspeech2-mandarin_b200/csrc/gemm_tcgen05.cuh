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

#include <cuda.h>  // CUtensorMap and enums only; cuTensorMapEncodeTiled is resolved at run time

#include "common.cuh"
#include "gemm_mma.cuh"

namespace fs2 {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;            // fp32/tf32 elements per 128-byte swizzle row
constexpr int STAGES = 4;
constexpr int THREADS = 192;
constexpr int A_STAGE_BYTES = BM * 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (kills the context) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start address >> 4 in bits [0,14), LBO (unused for swizzled K-major) = 1 in [16,30),
// SBO = 1024 B (one 8-row swizzle atom) >> 4 in [32,46), version 1 in [46,48), layout 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc(const void* smem_tile) {
  uint64_t d = (uint64_t)((smem_u32(smem_tile) >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// cute::UMMA::InstrDescriptor for kind::tf32, fp32 accumulate, both operands K-major.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[32]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

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
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    FS2_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    require(p != nullptr && q == cudaDriverEntryPointSuccess, FS2_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor [rows, cols] with row pitch ld (elements); box = [box_rows, 32 columns]
inline CUtensorMap make_map(const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, bool round_tf32,
                            bool reused, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode_fn()(&m, round_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                                 const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 swizzle,
                                 reused ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  require(r == CUDA_SUCCESS, FS2_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return m;
}

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
