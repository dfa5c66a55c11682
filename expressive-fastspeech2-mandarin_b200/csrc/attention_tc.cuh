// Varlen 2-head self-attention on tcgen05 (transformer/SubLayers.py:42-52, Modules.py:14-25).
// One CTA = 128 queries of one (utterance, head); keys/values stream in 64-key tiles.
//   S_j = Q K_j^T        tcgen05.mma kind::tf32, A = Q (smem, K-major), B = K_j (smem, K-major)  -> TMEM
//   P_j = exp2(S_j*c - m) computed by 128 softmax threads (thread = query row = TMEM lane), written
//                         back over S_j in TMEM (tcgen05.st)
//   O_j = P_j V_j        tcgen05.mma with A = P_j FROM TMEM and B = V_j (smem, MN-major)          -> TMEM
//   O   = O*alpha_j + O_j accumulated in registers by the softmax threads (online softmax, fp32)
// Q/K/V are read straight out of the packed [rows,768] QKV buffer by TMA (TFLOAT32 tensor map:
// rounded to TF32 on load); keys beyond the utterance are masked by length, never by a mask tensor.
// S/P and O_j are double-buffered in TMEM so that Q K_{j+1}^T overlaps the softmax of tile j.
// Q itself is moved into TMEM once (A operand of Q K^T from tensor memory), which frees its
// 64 KB of shared memory for a third K and V stage: loads run three tiles ahead of the MMAs.
#pragma once

#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace fs2 {
namespace attn_tc {

using namespace tc;

constexpr int BQ = 128, BKV = 64, THREADS = 192;
constexpr int Q_BYTES = BQ * D_HEAD * 4;          // 64 KB: 4 sub-tiles [128 rows x 128 B]
constexpr int K_BYTES = BKV * D_HEAD * 4;         // 32 KB: 4 sub-tiles [64 rows x 128 B]
constexpr int KV_STAGES = 3;
constexpr int BAR_OFF = Q_BYTES + 4 * K_BYTES;    // [Q | K2 V2] [K0 K1] [V0 V1]
constexpr int SMEM_TOTAL = BAR_OFF + 256 + 1024;
constexpr int TMEM_COLS = 512;                    // S0,S1: 2 x 64 | O0,O1: 2 x 128 | Q: 128
constexpr int LDQKV = 3 * D_MODEL;

__host__ __device__ constexpr uint32_t idesc_tf32(int m, int n, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

// MN-major descriptor for a 32-bit operand.  TF32 MN-major operands exist only in the
// SWIZZLE_128B_BASE32B layout (cute::UMMA::Layout_MN_SW128_32B_Atom: 32-byte chunks swizzled inside
// 128-byte rows, 4 rows per atom; TMA writes it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).  128-byte
// rows run along N (the head dimension); LBO = distance between 32-column sub-tiles, SBO = distance
// between 4-key groups.
__device__ __forceinline__ uint64_t umma_desc_mn(const void* smem_tile, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((smem_u32(smem_tile) >> 4) & 0x3FFF);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;   // UMMA::LayoutType::SWIZZLE_128B_BASE32B
  return d;
}

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    const __grid_constant__ CUtensorMap tmV, const int32_t* __restrict__ starts, const int32_t* __restrict__ lens,
                    const uint32_t* __restrict__ work, const int32_t* __restrict__ work_count, float* __restrict__ out,
                    __nv_bfloat16* __restrict__ out_b, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  // work item = 128 queries of one (utterance, head), taken from the longest-first work list (rowops.cuh): CTAs are
  // dispatched in blockIdx order, so the expensive items start first and the grid's tail is made of short utterances
  const int item = blockIdx.x >> 1, h = blockIdx.x & 1;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = smem;
  // stage 2 of both rings reuses the Q region once Q lives in TMEM
  auto k_stage = [&](int s) -> uint8_t* { return s < 2 ? smem + Q_BYTES + s * K_BYTES : smem; };
  auto v_stage = [&](int s) -> uint8_t* { return s < 2 ? smem + Q_BYTES + 2 * K_BYTES + s * K_BYTES : smem + K_BYTES; };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* q_full = bars;            // [1]
  uint64_t* q_moved = bars + 1;       // [1]  Q copied to TMEM: its smem may be overwritten
  uint64_t* kv_full = bars + 2;       // [3]  K_j and V_j of one tile share a stage and a barrier pair
  uint64_t* kv_empty = bars + 5;      // [3]  free after P_j V_j (three stages: loads run two tiles ahead)
  uint64_t* s_full = bars + 8;        // [2]
  uint64_t* p_full = bars + 10;       // [2]  also implies that O_{j-2} has been accumulated (program order
                                      //      of the softmax threads), so P_j V_j may overwrite that buffer
  uint64_t* o_full = bars + 12;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = warp_index(), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmKV)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
    mbar_init(q_full, 1);
    mbar_init(q_moved, 128);
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
  // ---- the prologue above overlaps the previous kernel's tail (programmatic dependent launch);
  // lens / starts / qkv are produced by earlier kernels of this forward
  pdl_trigger();
  pdl_wait();
  const bool active = item < *work_count;     // the grid is sized from a host-side bound: surplus CTAs only tear down
  const uint32_t wi = active ? work[item] : 0u;
  const int b = (int)(wi >> 16), q0 = (int)(wi & 0xFFFFu) * BQ;
  const int len = active ? lens[b] : 0;
  const int row0 = active ? starts[b] : 0;
  const int n_tiles = (len + BKV - 1) / BKV;
  const uint32_t tmem_s = tmem_base;          // + u*64
  const uint32_t tmem_o = tmem_base + 128;    // + u*128
  const uint32_t tmem_q = tmem_base + 384;    // 128 columns
#ifdef FS2_TRACE_BUILD   // phase timestamps for tools/trace_attention.py (block 0 only, dbg == 4)
  const bool trace = dbg == 4 && blockIdx.x == 0;
  const long long t_start = clock64();
  float* tr = out + (size_t)row0 * D_MODEL;
#define FS2_TRACE(tile, k) do { if (trace && (threadIdx.x & 31) == 0) tr[(tile) * 16 + (k)] = (float)(clock64() - t_start); } while (0)
#else
#define FS2_TRACE(tile, k) do { } while (0)
#endif

  if (!active) {
    // nothing to do
  } else if (warp == 0) {
    // ---- TMA producer (whole warp, one elected lane issues): Q once, then K_j / V_j into 3-stage rings
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(q_full, Q_BYTES);
#pragma unroll
      for (int dc = 0; dc < 4; ++dc) tma_load_2d(q_s + dc * (BQ * 128), &tmQ, h * D_HEAD + dc * 32, row0 + q0, q_full);
    }
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      const int sk = j % KV_STAGES;
      const uint32_t par = ((j / KV_STAGES) & 1) ^ 1;
      if (j == 2) mbar_wait(q_moved, 0);      // stage 2 lives where Q was staged
      uint8_t* k_s = k_stage(sk);
      uint8_t* v_s = v_stage(sk);
      mbar_wait(&kv_empty[sk], par);
      if (leader) {
        mbar_expect_tx(&kv_full[sk], 2 * K_BYTES);
#pragma unroll
        for (int dc = 0; dc < 4; ++dc)
          tma_load_2d(k_s + dc * (BKV * 128), &tmKV, D_MODEL + h * D_HEAD + dc * 32, row0 + j * BKV, &kv_full[sk]);
#pragma unroll
        for (int dc = 0; dc < 4; ++dc)
          tma_load_2d(v_s + dc * (BKV * 128), &tmV, 2 * D_MODEL + h * D_HEAD + dc * 32, row0 + j * BKV, &kv_full[sk]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---- MMA issuer (whole warp runs the loop, one elected lane issues)
    const bool leader = elect_one();
    constexpr uint32_t idesc_qk = idesc_tf32(BQ, BKV, 0);
    constexpr uint32_t idesc_pv = idesc_tf32(BQ, D_HEAD, 1);
    auto issue_qk = [&](int j) {
      const int u = j & 1, sk = j % KV_STAGES;
      FS2_TRACE(j, 8);
      mbar_wait(&kv_full[sk], (j / KV_STAGES) & 1);
      FS2_TRACE(j, 9);
      tc_fence_after();
      const uint8_t* k_s = k_stage(sk);
      if (leader) {
#pragma unroll
        for (int dc = 0; dc < 4; ++dc) {
          const uint64_t db = umma_desc(k_s + dc * (BKV * 128));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_tf32_ts(tmem_s + u * BKV, tmem_q + dc * 32 + kk * 8, db + 2 * kk, idesc_qk, (dc | kk) != 0);
        }
        FS2_TRACE(j, 10);
        umma_commit(&s_full[u]);
        FS2_TRACE(j, 11);
      }
      __syncwarp();
    };
    mbar_wait(q_moved, 0);
    tc_fence_after();
    issue_qk(0);
    for (int j = 0; j < n_tiles; ++j) {
      const int u = j & 1;
      const uint32_t par = (j >> 1) & 1;
      if (j + 1 < n_tiles) issue_qk(j + 1);
      const int sk = j % KV_STAGES;
      FS2_TRACE(j, 4);
      mbar_wait(&p_full[u], par);      // P_j written; O buffer u drained (see p_full above); V_j landed with K_j
      FS2_TRACE(j, 5);
      FS2_TRACE(j, 6);
      tc_fence_after();
      const uint8_t* v_s = v_stage(sk);
      const uint64_t dv = umma_desc_mn(v_s, BKV * 128, 512);
      if (leader) {
#pragma unroll
        for (int k8 = 0; k8 < BKV / 8; ++k8)
          umma_tf32_ts(tmem_o + u * D_HEAD, tmem_s + u * BKV + k8 * 8, dv + (uint64_t)(k8 * (1024 >> 4)), idesc_pv, k8 != 0);
        FS2_TRACE(j, 12);
        umma_commit(&o_full[u]);
        umma_commit(&kv_empty[sk]);
        FS2_TRACE(j, 13);
      }
      __syncwarp();
    }
  } else {
    // ---- softmax + accumulation: thread = query row
    const int q = warp & 3;
    const int qrow = q0 + q * 32 + lane;                       // row inside the utterance
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const float c = 1.4426950408889634f / sqrtf((float)D_HEAD);  // log2(e) / temperature
    {
      // Q (TF32-rounded by TMA) from swizzled smem to TMEM: thread = query row = TMEM lane
      mbar_wait(q_full, 0);
      const int r = q * 32 + lane;
      const uint32_t qa = smem_u32(q_s) + r * 128;
      const uint32_t sx = (uint32_t)(r & 7) << 4;
#pragma unroll 1
      for (int dc = 0; dc < 4; ++dc) {
        float v[32];
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          float4 t4;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n"
                       : "=f"(t4.x), "=f"(t4.y), "=f"(t4.z), "=f"(t4.w)
                       : "r"(qa + dc * (BQ * 128) + ((cc << 4) ^ sx)));
          v[cc * 4] = t4.x; v[cc * 4 + 1] = t4.y; v[cc * 4 + 2] = t4.z; v[cc * 4 + 3] = t4.w;
        }
        tmem_st32(tmem_q + lane_sel + dc * 32, v);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic reads before the TMA overwrites
      tc_fence_before();
      mbar_arrive(q_moved);
    }
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
      tc_fence_before();   // ordered before this thread's next p_full arrive, which releases the O buffer
    };

    // m is the running row maximum of the RAW scores; exp2 arguments are s*c - m*c (one FFMA each)
    for (int j = 0; j < n_tiles; ++j) {
      const int u = j & 1;
      FS2_TRACE(j, 0);
      mbar_wait(&s_full[u], (j >> 1) & 1);
      FS2_TRACE(j, 1);
      tc_fence_after();
      float s0[32], s1[32];
      tmem_ld32_issue(tmem_s + lane_sel + u * BKV, s0);
      tmem_ld32_issue(tmem_s + lane_sel + u * BKV + 32, s1);
      tmem_ld_wait();
      const int key0 = j * BKV;
      if (key0 + BKV > len) {   // only the last tile has keys beyond the utterance
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (key0 + i >= len) s0[i] = -INFINITY;
          if (key0 + 32 + i >= len) s1[i] = -INFINITY;
        }
      }
      float mx[4] = {m, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        mx[0] = fmaxf(mx[0], s0[i]);
        mx[1] = fmaxf(mx[1], s0[i + 1]);
        mx[2] = fmaxf(mx[2], s1[i]);
        mx[3] = fmaxf(mx[3], s1[i + 1]);
      }
      const float m_new = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));   // finite: key0 < len
      const float alpha = ex2_approx((m - m_new) * c);                        // 0 on the first tile (m = -inf)
      m = m_new;
      const float mc = m_new * c;
      float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      // P is rounded to TF32 (nearest, ties away) with integer arithmetic: (bits + 0x1000) & ~0x1fff runs on
      // the ALU pipe, whereas cvt.rna.tf32 shares the XU pipe with ex2 and would double its load.  p is in [0, 1].
      auto p_of = [&](float s) {
        const uint32_t bits = (__float_as_uint(ex2_approx(fmaf(s, c, -mc))) + 0x1000u) & 0xFFFFE000u;
        return __uint_as_float(bits);
      };
      for (int i = 0; i < 32; i += 2) {
        s0[i] = p_of(s0[i]);
        s0[i + 1] = p_of(s0[i + 1]);
        s1[i] = p_of(s1[i]);
        s1[i + 1] = p_of(s1[i + 1]);
        sum[0] += s0[i];
        sum[1] += s0[i + 1];
        sum[2] += s1[i];
        sum[3] += s1[i + 1];
      }
      l = fmaf(l, alpha, (sum[0] + sum[1]) + (sum[2] + sum[3]));
      tmem_st32(tmem_s + lane_sel + u * BKV, s0);
      tmem_st32(tmem_s + lane_sel + u * BKV + 32, s1);
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      tc_fence_before();
      mbar_arrive(&p_full[u]);
      FS2_TRACE(j, 2);
      if (j >= 1) accumulate(j - 1, alpha_prev);
      FS2_TRACE(j, 3);
      alpha_prev = alpha;
    }
    accumulate(n_tiles - 1, alpha_prev);

    if (dbg == 4) {
      // timestamps only
    } else
    if (dbg != 0 && qrow < len) {   // raw dumps for bring-up
      float* dst = out + (size_t)(row0 + qrow) * D_MODEL + h * D_HEAD;
      if (dbg == 1) {               // P of tile 0 read back from TMEM (64 values), then l
        float v0[32], v1[32];
        tmem_ld32(tmem_s + lane_sel, v0);
        tmem_ld32(tmem_s + lane_sel + 32, v1);
        for (int i = 0; i < 32; ++i) { dst[i] = v0[i]; dst[32 + i] = v1[i]; }
        dst[64] = l;
      } else {
        for (int i = 0; i < D_HEAD; ++i) dst[i] = o[i];
      }
    } else
    if (qrow < len) {
      const float inv = 1.f / l;
      if (out_b != nullptr) {   // BF16 mode: the context is only ever the A operand of the fc contraction
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
      } else {
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

inline int& debug_flag() {
  static int f = 0;
  return f;
}

// Upper bound of the work-list length known on the host: sum_b ceil(len_b / 128) <= total_rows / 128 + batch.
inline int work_bound(int64_t total_len, int batch, int max_len) {
  const int64_t a = total_len / BQ + batch, b = (int64_t)batch * ((max_len + BQ - 1) / BQ);
  return (int)std::min(a, b);
}

inline void launch(const float* qkv, int rows, const int32_t* starts, const int32_t* lens, const uint32_t* work,
                   const int32_t* work_count, int work_cap, float* out, cudaStream_t stream, void* out_bf16 = nullptr) {
  if (work_cap <= 0 || rows <= 0) return;
  static bool configured[64] = {};
  int dev = 0;
  FS2_CUDA_OK(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    FS2_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    configured[dev & 63] = true;
  }
  const CUtensorMap tmQ = make_map(qkv, rows, LDQKV, LDQKV, BQ, true, false);
  const CUtensorMap tmKV = make_map(qkv, rows, LDQKV, LDQKV, BKV, true, true);
  const CUtensorMap tmV = make_map(qkv, rows, LDQKV, LDQKV, BKV, true, true, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  launch_pdl(attention_tc_kernel, dim3(N_HEAD * work_cap), dim3(THREADS), SMEM_TOTAL, stream, 1, tmQ, tmKV, tmV, starts, lens,
             work, work_count, out, static_cast<__nv_bfloat16*>(out_bf16), debug_flag());
  FS2_LAUNCHED();
}

}  // namespace attn_tc
}  // namespace fs2
