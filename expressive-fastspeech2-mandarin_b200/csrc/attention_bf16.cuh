// Varlen 2-head self-attention on tcgen05 with bf16 operands (FS2_MATH_BF16; transformer/SubLayers.py:42-52,
// Modules.py:14-25).  Same structure as attention_tc.cuh (one CTA = 128 queries of one (utterance, head), taken from the
// longest-first work list; S and O_j double-buffered in TMEM; online softmax by 128 row-owner threads with fp32 running
// maximum / sum / output), but:
//   * Q, K, V are read as bf16 straight out of the packed [rows, 768] bf16 QKV buffer the QKV GEMM writes;
//   * keys / values stream in tiles of 128 (a K + V stage is 64 KB, three stages fit beside the 32 KB Q tile), so one
//     tile is 8 + 8 MMAs of N = 128 instead of 2 x (16 + 8) narrower TF32 ones -- the kernel is bound by the fixed cost of
//     each MMA, which this cuts 2.4x per key;
//   * S = Q K^T runs in the SS form (Q from shared memory), P = exp2(...) is written back to TMEM as PACKED bf16 pairs
//     (column j of the P region holds keys 2j, 2j+1 of the row) and is the A operand of O_j = P V_j from tensor memory;
//   * V is the MN-major B operand: the TMA image of a [128 keys x 64 dims] bf16 box with SWIZZLE_128B is exactly the
//     canonical MN-major SWIZZLE_128B layout (128-byte rows along N, 8 K-rows per 1024-byte swizzle atom).
#pragma once

#include <cuda_bf16.h>

#include <algorithm>

#include "attention_tc.cuh"
#include "common.cuh"
#include "tc_ptx.cuh"

namespace fs2 {
namespace attn_bf {

using namespace tc;

constexpr int BQ = 128, BKV = 128, THREADS = 192;
constexpr int SUB = BQ * 128;                     // one [128 rows x 64 bf16] SWIZZLE_128B sub-tile: 16 KB
constexpr int Q_BYTES = 2 * SUB;                  // 32 KB: dims 0-63 | dims 64-127
constexpr int KV_BYTES = 2 * SUB;                 // K or V tile of one stage
constexpr int KV_STAGES = 3;
constexpr int BAR_OFF = Q_BYTES + KV_STAGES * 2 * KV_BYTES;
constexpr int SMEM_TOTAL = BAR_OFF + 256 + 1024;
constexpr int TMEM_COLS = 512;                    // S0, S1: 2 x 128 (P overwrites the first 64 columns) | O0, O1: 2 x 128
constexpr int LDQKV = 3 * D_MODEL;
static_assert(SMEM_TOTAL <= 232448, "shared memory budget");

__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

// MN-major SWIZZLE_128B descriptor (cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units):
// LBO = distance between the 64-element (128-byte) column blocks, SBO = distance between groups of 8 K-rows.
__device__ __forceinline__ uint64_t umma_desc_mn128(const void* smem_tile, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((smem_u32(smem_tile) >> 4) & 0x3FFF);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;   // UMMA::LayoutType::SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
attention_bf16_kernel(const __grid_constant__ CUtensorMap tmQKV, const int32_t* __restrict__ starts,
                      const int32_t* __restrict__ lens, const uint32_t* __restrict__ work,
                      const int32_t* __restrict__ work_count, float* __restrict__ out, __nv_bfloat16* __restrict__ out_b) {
  extern __shared__ uint8_t smem_raw[];
  const int item = blockIdx.x >> 1, h = blockIdx.x & 1;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = smem;
  auto k_stage = [&](int s) -> uint8_t* { return smem + Q_BYTES + s * 2 * KV_BYTES; };
  auto v_stage = [&](int s) -> uint8_t* { return smem + Q_BYTES + s * 2 * KV_BYTES + KV_BYTES; };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* q_full = bars;            // [1]
  uint64_t* kv_full = bars + 1;       // [3]
  uint64_t* kv_empty = bars + 4;      // [3]  free after P_j V_j
  uint64_t* s_full = bars + 7;        // [2]
  uint64_t* p_full = bars + 9;        // [2]  also: O_{j-2} has been accumulated (program order of the softmax threads)
  uint64_t* o_full = bars + 11;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = warp_index(), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmQKV)) : "memory");
    mbar_init(q_full, 1);
    for (int u = 0; u < KV_STAGES; ++u) {
      mbar_init(&kv_full[u], 1);
      mbar_init(&kv_empty[u], 1);
    }
    for (int u = 0; u < 2; ++u) {
      mbar_init(&s_full[u], 1);
      mbar_init(&p_full[u], 128);
      mbar_init(&o_full[u], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();
  const bool active = item < *work_count;
  const uint32_t wi = active ? work[item] : 0u;
  const int b = (int)(wi >> 16), q0 = (int)(wi & 0xFFFFu) * BQ;
  const int len = active ? lens[b] : 0;
  const int row0 = active ? starts[b] : 0;
  const int n_tiles = (len + BKV - 1) / BKV;
  const uint32_t tmem_s = tmem_base;          // + u*128
  const uint32_t tmem_o = tmem_base + 256;    // + u*128

  if (!active) {
    // nothing to do
  } else if (warp == 0) {
    // ---- TMA producer: Q once, then K_j / V_j into the 3-stage ring (two 64-column boxes each)
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(q_full, Q_BYTES);
      tma_load_2d(q_s, &tmQKV, h * D_HEAD, row0 + q0, q_full);
      tma_load_2d(q_s + SUB, &tmQKV, h * D_HEAD + 64, row0 + q0, q_full);
    }
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      const int sk = j % KV_STAGES;
      mbar_wait(&kv_empty[sk], ((j / KV_STAGES) & 1) ^ 1);
      if (leader) {
        mbar_expect_tx(&kv_full[sk], 2 * KV_BYTES);
        tma_load_2d(k_stage(sk), &tmQKV, D_MODEL + h * D_HEAD, row0 + j * BKV, &kv_full[sk]);
        tma_load_2d(k_stage(sk) + SUB, &tmQKV, D_MODEL + h * D_HEAD + 64, row0 + j * BKV, &kv_full[sk]);
        tma_load_2d(v_stage(sk), &tmQKV, 2 * D_MODEL + h * D_HEAD, row0 + j * BKV, &kv_full[sk]);
        tma_load_2d(v_stage(sk) + SUB, &tmQKV, 2 * D_MODEL + h * D_HEAD + 64, row0 + j * BKV, &kv_full[sk]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---- MMA issuer
    const bool leader = elect_one();
    constexpr uint32_t idesc_qk = idesc_bf16(BQ, BKV, 0);
    constexpr uint32_t idesc_pv = idesc_bf16(BQ, D_HEAD, 1);
    auto issue_qk = [&](int j) {
      const int u = j & 1, sk = j % KV_STAGES;
      mbar_wait(&kv_full[sk], (j / KV_STAGES) & 1);
      tc_fence_after();
      const uint8_t* k_s = k_stage(sk);
      if (leader) {
#pragma unroll
        for (int dc = 0; dc < 2; ++dc) {       // 64 dims per 128-byte swizzle row, 16 dims (32 bytes) per MMA
          const uint64_t da = umma_desc(q_s + dc * SUB), db = umma_desc(k_s + dc * SUB);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem_s + u * BKV, da + 2 * kk, db + 2 * kk, idesc_qk, (dc | kk) != 0);
        }
        umma_commit(&s_full[u]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    tc_fence_after();
    issue_qk(0);
    for (int j = 0; j < n_tiles; ++j) {
      const int u = j & 1;
      if (j + 1 < n_tiles) issue_qk(j + 1);
      const int sk = j % KV_STAGES;
      mbar_wait(&p_full[u], (j >> 1) & 1);      // P_j written; O buffer u drained; V_j landed with K_j
      tc_fence_after();
      const uint64_t dv = umma_desc_mn128(v_stage(sk), SUB, 1024);
      if (leader) {
#pragma unroll
        for (int k16 = 0; k16 < BKV / 16; ++k16)   // 16 keys per MMA: 8 packed columns of P, 16 rows (2048 B) of V
          umma_bf16_ts(tmem_o + u * D_HEAD, tmem_s + u * BKV + k16 * 8, dv + (uint64_t)(k16 * (2048 >> 4)), idesc_pv, k16 != 0);
        umma_commit(&o_full[u]);
        umma_commit(&kv_empty[sk]);
      }
      __syncwarp();
    }
  } else {
    // ---- softmax + accumulation: thread = query row
    const int q = warp & 3;
    const int qrow = q0 + q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const float c = 1.4426950408889634f / sqrtf((float)D_HEAD);
    float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
    float o[D_HEAD];
#pragma unroll
    for (int i = 0; i < D_HEAD; ++i) o[i] = 0.f;

    auto accumulate = [&](int j, float alpha) {
      const int u = j & 1;
      mbar_wait(&o_full[u], (j >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < D_HEAD; c0 += 64) {
        float v0[32], v1[32];
        tmem_ld32_issue(tmem_o + lane_sel + u * D_HEAD + c0, v0);
        tmem_ld32_issue(tmem_o + lane_sel + u * D_HEAD + c0 + 32, v1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          o[c0 + i] = fmaf(o[c0 + i], alpha, v0[i]);
          o[c0 + 32 + i] = fmaf(o[c0 + 32 + i], alpha, v1[i]);
        }
      }
      tc_fence_before();
    };

    for (int j = 0; j < n_tiles; ++j) {
      const int u = j & 1;
      mbar_wait(&s_full[u], (j >> 1) & 1);
      tc_fence_after();
      const int key0 = j * BKV;
      const bool tail = key0 + BKV > len;     // only the last tile has keys beyond the utterance
      const uint32_t s_addr = tmem_s + lane_sel + u * BKV;
      // pass 1: row maximum over the 128 scores of this tile
      float mx[4] = {m, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
      for (int c0 = 0; c0 < BKV; c0 += 64) {
        float s0[32], s1[32];
        tmem_ld32_issue(s_addr + c0, s0);
        tmem_ld32_issue(s_addr + c0 + 32, s1);
        tmem_ld_wait();
        if (tail) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (key0 + c0 + i >= len) s0[i] = -INFINITY;
            if (key0 + c0 + 32 + i >= len) s1[i] = -INFINITY;
          }
        }
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          mx[0] = fmaxf(mx[0], s0[i]);
          mx[1] = fmaxf(mx[1], s0[i + 1]);
          mx[2] = fmaxf(mx[2], s1[i]);
          mx[3] = fmaxf(mx[3], s1[i + 1]);
        }
      }
      const float m_new = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));   // finite: key0 < len
      const float alpha = ex2_approx((m - m_new) * c);                        // 0 on the first tile (m = -inf)
      m = m_new;
      const float mc = m_new * c;
      // pass 2: p = exp2(s*c - m*c) rounded to bf16, packed in pairs, written over the first 64 columns of S
      // (the packed columns [16k, 16k+16) are written after the score columns [32k, 32k+32) they come from were read)
      float sum[2] = {0.f, 0.f};
#pragma unroll 1
      for (int c0 = 0; c0 < BKV; c0 += 32) {
        float s0[32];
        tmem_ld32(s_addr + c0, s0);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a = ex2_approx(fmaf(s0[2 * i], c, -mc)), b2 = ex2_approx(fmaf(s0[2 * i + 1], c, -mc));
          if (tail) {
            if (key0 + c0 + 2 * i >= len) a = 0.f;
            if (key0 + c0 + 2 * i + 1 >= len) b2 = 0.f;
          }
          const __nv_bfloat162 hh = __floats2bfloat162_rn(a, b2);
          pk[i] = *reinterpret_cast<const uint32_t*>(&hh);
          // the running sum uses the ROUNDED probabilities, i.e. exactly what the tensor core multiplies with V
          sum[0] += __low2float(hh);
          sum[1] += __high2float(hh);
        }
        tmem_st16(s_addr + (c0 >> 1), pk);
      }
      l = fmaf(l, alpha, sum[0] + sum[1]);
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      tc_fence_before();
      mbar_arrive(&p_full[u]);
      if (j >= 1) accumulate(j - 1, alpha_prev);
      alpha_prev = alpha;
    }
    accumulate(n_tiles - 1, alpha_prev);

    if (qrow < len) {
      const float inv = 1.f / l;
      if (out_b != nullptr) {
        __nv_bfloat16* dst = out_b + (size_t)(row0 + qrow) * D_MODEL + h * D_HEAD;
#pragma unroll
        for (int i = 0; i < D_HEAD; i += 8) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 hh = __floats2bfloat162_rn(o[i + 2 * e] * inv, o[i + 2 * e + 1] * inv);
            w[e] = *reinterpret_cast<const uint32_t*>(&hh);
          }
          *reinterpret_cast<uint4*>(dst + i) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      if (out != nullptr) {
        float* dst = out + (size_t)(row0 + qrow) * D_MODEL + h * D_HEAD;
#pragma unroll
        for (int i = 0; i < D_HEAD; i += 4)
          *reinterpret_cast<float4*>(dst + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// qkv: bf16 [rows, 768].  out (fp32) and out_bf16 may each be null.
inline void launch(const __nv_bfloat16* qkv, int rows, const int32_t* starts, const int32_t* lens, const uint32_t* work,
                   const int32_t* work_count, int work_cap, float* out, void* out_bf16, cudaStream_t stream) {
  if (work_cap <= 0 || rows <= 0) return;
  static bool configured[64] = {};
  int dev = 0;
  FS2_CUDA_OK(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    FS2_CUDA_OK(cudaFuncSetAttribute(attention_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    configured[dev & 63] = true;
  }
  const CUtensorMap tm = make_map_any(qkv, rows, LDQKV, LDQKV, BQ, 64, MAP_BF16, CU_TENSOR_MAP_SWIZZLE_128B);
  launch_pdl(attention_bf16_kernel, dim3(N_HEAD * work_cap), dim3(THREADS), SMEM_TOTAL, stream, 1, tm, starts, lens, work,
             work_count, out, static_cast<__nv_bfloat16*>(out_bf16));
  FS2_LAUNCHED();
}

}  // namespace attn_bf
}  // namespace fs2
