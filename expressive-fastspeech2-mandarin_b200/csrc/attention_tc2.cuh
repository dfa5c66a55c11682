// Varlen 2-head self-attention (transformer/SubLayers.py:42-52, Modules.py:14-25) as a 2-SM tcgen05 kernel: a CTA PAIR owns
// 256 queries of one (utterance, head) -- 128 rows per CTA -- and both contractions are tcgen05.mma.cta_group::2 with M = 256:
//   S_j = Q K_j^T   A = Q (each CTA's own 128 rows, shared memory, K-major), B = K_j: 128 keys, EACH CTA HOLDS 64 OF THEM
//   P_j = exp2(S_j c - m c), TF32-rounded, written over S_j in tensor memory by the 128 row-owner threads of each CTA
//   O_j = P_j V_j   A = P_j from tensor memory, B = V_j (MN-major): 128 keys x 128 head columns, EACH CTA HOLDS 64 COLUMNS
//   O   = O alpha_j + O_j in registers (online softmax, fp32)
// Against attention_tc.cuh (one CTA per 128 queries, 64-key tiles, every CTA loading whole K/V tiles): the K/V bytes a CTA
// pulls from L2 per key are halved with no software coupling between the CTAs, the same shared memory holds two stages of 128
// keys, and every MMA has N = 128 instead of 64.
// MEASURED SLOWER than attention_tc.cuh at config 2 (dec.attention 0.48 vs 0.345 ms per step; traces in
// profiles/r02_attention_experiments.txt): 32 MMAs of ~100 cycles per 128-key tile are 3,200 cycles for 2,048 of tensor work, the
// 128 scores of a row do not fit in registers beside its 128 accumulators (S is read from tensor memory twice), and the coarser
// 256-query / 128-key granularity adds 14 % of tile work.  Not the default: FS2_ATTN_PAIR=2 / debug flag 8 = 2 selects it, and
// tests/test_gpu_ops.py keeps it correct.
// Tensor memory (512 columns per CTA): S0 S1 (2 x 128) | O0 O1 (2 x 128); Q stays in shared memory.
#pragma once

#include "attention_tc.cuh"
#include "gemm_tc2.cuh"

namespace fs2 {
namespace attn2 {

using namespace tc;
using attn_tc::idesc_tf32;
using attn_tc::umma_desc_mn;
using tc2::tma_load_2d_2sm;
using tc2::umma_2sm;
using tc2::umma_commit_2sm;

constexpr int BQ = 128, BKV = 128, THREADS = 224;   // producer, Q K^T issuer, 4 softmax warps, P V issuer
constexpr int LDQKV = 3 * D_MODEL;
constexpr int Q_BYTES = BQ * D_HEAD * 4;              // 64 KB: 4 sub-tiles [128 rows x 128 B]
constexpr int KH_BYTES = (BKV / 2) * D_HEAD * 4;      // 32 KB: this CTA's 64 keys, 4 sub-tiles [64 rows x 128 B]
constexpr int VH_BYTES = BKV * (D_HEAD / 2) * 4;      // 32 KB: this CTA's 64 head columns, 2 sub-tiles [128 keys x 128 B]
constexpr int STAGES = 2;
constexpr int OFF_K = Q_BYTES, OFF_V = OFF_K + STAGES * KH_BYTES, BAR_OFF = OFF_V + STAGES * VH_BYTES;
constexpr int SMEM_TOTAL = BAR_OFF + 256 + 1024;
static_assert(SMEM_TOTAL <= 232448, "shared memory budget");

// the two issuing warps of the even CTA wait most of the time; a bare try_wait loop there takes issue slots from the softmax
// warps that share their schedulers
__device__ __forceinline__ void wait_cl(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (true) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    __nanosleep(64);
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

__device__ __forceinline__ void umma_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
attention_2sm_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const int32_t* __restrict__ starts, const int32_t* __restrict__ lens,
                     const uint32_t* __restrict__ work, const int32_t* __restrict__ work_count, float* __restrict__ out,
                     __nv_bfloat16* __restrict__ out_b, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  FS2_CTA_STAMP(0);
  // cluster = (rank 0, rank 1) of one work item and head; the item list is longest-first (rowops.cuh, built for 256 rows per item)
  const int rank = (int)(blockIdx.x & 1), h = (int)(blockIdx.x >> 1) & 1, item = (int)(blockIdx.x >> 2);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* q_full = bars;            // [1]  rank 0's counts the bytes of both CTAs (as do k_full / v_full)
  uint64_t* k_full = bars + 1;        // [2]
  uint64_t* k_empty = bars + 3;       // [2]  multicast commit after Q K_j^T
  uint64_t* v_full = bars + 5;        // [2]
  uint64_t* v_empty = bars + 7;       // [2]  multicast commit after P_j V_j
  uint64_t* s_full = bars + 9;        // [2]  multicast commit: S_j is in both CTAs' tensor memory
  uint64_t* p_full = bars + 11;       // [2]  rank 0's: 256 arrivals (P_j written AND O_{j-2} accumulated, in both CTAs)
  uint64_t* o_full = bars + 13;       // [2]  multicast commit
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = warp_index(), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmK)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
    mbar_init(q_full, 1);
    for (int u = 0; u < 2; ++u) {
      mbar_init(&k_full[u], 1);
      mbar_init(&k_empty[u], 1);
      mbar_init(&v_full[u], 1);
      mbar_init(&v_empty[u], 1);
      mbar_init(&s_full[u], 1);
      mbar_init(&p_full[u], 256);
      mbar_init(&o_full[u], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {   // the same warp of both CTAs, the same destination offset
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster_sync_all();   // the peer's barriers are initialised before anything is counted on / multicast to them
  pdl_trigger();
  pdl_wait();
  FS2_CTA_STAMP(1);
  const bool listed = item < *work_count;     // the grid is sized from a host-side bound: surplus clusters only tear down
  const uint32_t wi = listed ? work[item] : 0u;
  const int b = (int)(wi >> 16), q0 = ((int)(wi & 0xFFFFu) * 2 + rank) * BQ;
  const int len = listed ? lens[b] : 0;
  const int row0 = listed ? starts[b] : 0;
  // a CTA whose 128 rows lie beyond the utterance (odd number of query tiles) goes through the whole protocol on rows that
  // belong to nobody -- the pair's MMAs are M = 256 either way -- and stores nothing
  const int n_tiles = (len + BKV - 1) / BKV;
  const uint32_t tmem_s = tmem_base;          // + u * 128
  const uint32_t tmem_o = tmem_base + 256;    // + u * 128

#ifdef FS2_TRACE_BUILD   // phase timestamps for tools/trace_attention.py (block 0 only, dbg == 4), written into the output rows
  const bool trace = dbg == 4 && blockIdx.x == 0;
  const long long t_start = clock64();
  float* tr = out + (size_t)row0 * D_MODEL;
#define FS2_TRACE2(tile, k) do { if (trace && (threadIdx.x & 31) == 0 && (tile) < 11) tr[(tile) * 16 + (k)] = (float)(clock64() - t_start); } while (0)
#else
#define FS2_TRACE2(tile, k) do { } while (0)
#endif
  if (!listed || n_tiles == 0) {
    // nothing to do
  } else if (warp == 0) {
    // ---- TMA producer of THIS CTA's operand parts: its Q rows, its 64 keys of every K tile, its 64 head columns of every V tile
    const bool leader = elect_one();
    if (leader) {
      if (rank == 0) mbar_expect_tx(q_full, 2 * Q_BYTES);
#pragma unroll
      for (int dc = 0; dc < 4; ++dc) tma_load_2d_2sm(q_s + dc * (BQ * 128), &tmQ, h * D_HEAD + dc * 32, row0 + q0, q_full);
    }
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j & 1;
      const uint32_t par = ((j >> 1) & 1) ^ 1;
      uint8_t* k_s = smem + OFF_K + s * KH_BYTES;
      uint8_t* v_s = smem + OFF_V + s * VH_BYTES;
      wait_cl(&k_empty[s], par);
      if (leader) {
        if (rank == 0) mbar_expect_tx(&k_full[s], 2 * KH_BYTES);
#pragma unroll
        for (int dc = 0; dc < 4; ++dc)
          tma_load_2d_2sm(k_s + dc * ((BKV / 2) * 128), &tmK, D_MODEL + h * D_HEAD + dc * 32, row0 + j * BKV + rank * (BKV / 2), &k_full[s]);
      }
      __syncwarp();
      wait_cl(&v_empty[s], par);
      if (leader) {
        if (rank == 0) mbar_expect_tx(&v_full[s], 2 * VH_BYTES);
#pragma unroll
        for (int dc = 0; dc < 2; ++dc)
          tma_load_2d_2sm(v_s + dc * (BKV * 128), &tmV, 2 * D_MODEL + h * D_HEAD + (rank * 2 + dc) * 32, row0 + j * BKV, &v_full[s]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---- MMA issuer 1 (even CTA only): S_j = Q K_j^T for the pair's 256 rows
    if (rank == 0) {
      const bool leader = elect_one();
      constexpr uint32_t idesc_qk = idesc_tf32(2 * BQ, BKV, 0);
      wait_cl(q_full, 0);
      tc_fence_after();
      for (int j = 0; j < n_tiles; ++j) {
        const int u = j & 1;
        FS2_TRACE2(j, 8);
        wait_cl(&k_full[u], (j >> 1) & 1);
        FS2_TRACE2(j, 9);
        if (j >= 2) wait_cl(&o_full[u], ((j - 2) >> 1) & 1);   // P V_{j-2} has consumed the P that lives in S buffer u
        FS2_TRACE2(j, 10);
        tc_fence_after();
        const uint8_t* k_s = smem + OFF_K + u * KH_BYTES;
        if (leader) {
#pragma unroll
          for (int dc = 0; dc < 4; ++dc) {
            const uint64_t da = umma_desc(q_s + dc * (BQ * 128)), db = umma_desc(k_s + dc * ((BKV / 2) * 128));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_2sm<false>(tmem_s + u * BKV, da + 2 * kk, db + 2 * kk, idesc_qk, (dc | kk) != 0 ? 1u : 0u);
          }
          umma_commit_2sm(&s_full[u]);
          umma_commit_2sm(&k_empty[u]);
        }
        __syncwarp();
        FS2_TRACE2(j, 11);
      }
    }
  } else if (warp == 6) {
    // ---- MMA issuer 2 (even CTA only): O_j = P_j V_j, A = P from tensor memory
    if (rank == 0) {
      const bool leader = elect_one();
      constexpr uint32_t idesc_pv = idesc_tf32(2 * BQ, D_HEAD, 1);
      for (int j = 0; j < n_tiles; ++j) {
        const int u = j & 1;
        FS2_TRACE2(j, 4);
        wait_cl(&v_full[u], (j >> 1) & 1);
        FS2_TRACE2(j, 5);
        wait_cl(&p_full[u], (j >> 1) & 1);
        FS2_TRACE2(j, 12);
        tc_fence_after();
        const uint64_t dv = umma_desc_mn(smem + OFF_V + u * VH_BYTES, BKV * 128, 512);
        if (leader) {
#pragma unroll
          for (int k8 = 0; k8 < BKV / 8; ++k8)
            umma_ts_2sm(tmem_o + u * D_HEAD, tmem_s + u * BKV + k8 * 8, dv + (uint64_t)(k8 * (1024 >> 4)), idesc_pv, k8 != 0 ? 1u : 0u);
          umma_commit_2sm(&o_full[u]);
          umma_commit_2sm(&v_empty[u]);
        }
        __syncwarp();
        FS2_TRACE2(j, 13);
      }
    }
  } else {
    // ---- softmax + accumulation: thread = query row of this CTA
    const int q = warp & 3;
    const int qrow = q0 + q * 32 + lane;                       // row inside the utterance
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const float c = 1.4426950408889634f / sqrtf((float)D_HEAD);  // log2(e) / temperature
    float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
    float o[D_HEAD];
#pragma unroll
    for (int i = 0; i < D_HEAD; ++i) o[i] = 0.f;
    auto arrive_p = [&](uint64_t* bar) {
      if (rank != 0) mbar_arrive_remote(dsmem_addr(bar, 0));
      else mbar_arrive(bar);
    };
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
    for (int j = 0; j < n_tiles; ++j) {
      const int u = j & 1;
      const uint32_t sa = tmem_s + lane_sel + u * BKV;
      if (q == 0) FS2_TRACE2(j, 0);
      mbar_wait(&s_full[u], (j >> 1) & 1);
      if (q == 0) FS2_TRACE2(j, 1);
      if (j == 0) FS2_CTA_STAMP(2);
      tc_fence_after();
      const int key0 = j * BKV;
      const bool ragged = key0 + BKV > len;   // only the last tile has keys beyond the utterance
      // (masked in a uniform branch of its own: per-element predicates in the main loops serialise the 128 scores through
      // the seven predicate registers -- measured 39 cycles per score instead of 13)
      auto mask_chunk = [&](float (&v)[32], int c4) {
        if (!ragged) return;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (key0 + c4 * 32 + i >= len) v[i] = -INFINITY;
      };
      // pass 1: the row maximum over the tile's 128 scores (the scores do not fit in registers beside the 128 accumulators:
      // they are read from tensor memory twice)
      float mx[4] = {m, -INFINITY, -INFINITY, -INFINITY};
      {
        float va[32], vb[32];
        tmem_ld32_issue(sa, va);
#pragma unroll
        for (int c4 = 0; c4 < 4; c4 += 2) {
          tmem_ld_wait();
          tmem_ld32_issue(sa + (c4 + 1) * 32, vb);
          mask_chunk(va, c4);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            mx[0] = fmaxf(mx[0], va[i]);
            mx[1] = fmaxf(mx[1], va[i + 1]);
          }
          tmem_ld_wait();
          if (c4 + 2 < 4) tmem_ld32_issue(sa + (c4 + 2) * 32, va);
          mask_chunk(vb, c4 + 1);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            mx[2] = fmaxf(mx[2], vb[i]);
            mx[3] = fmaxf(mx[3], vb[i + 1]);
          }
        }
      }
      if (q == 0) FS2_TRACE2(j, 2);
      const float m_new = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));   // finite: key0 < len
      const float alpha = ex2_approx((m - m_new) * c);                        // 0 on the first tile (m = -inf)
      m = m_new;
      const float mc = m_new * c;
      // pass 2: P = exp2(s c - m c), rounded to TF32 (nearest, ties away) with integer arithmetic, written over S
      float sum[4] = {0.f, 0.f, 0.f, 0.f};
      auto p_of = [&](float s) {
        const uint32_t bits = (__float_as_uint(ex2_approx(fmaf(s, c, -mc))) + 0x1000u) & 0xFFFFE000u;
        return __uint_as_float(bits);
      };
      {
        float va[32], vb[32];
        tmem_ld32_issue(sa, va);
#pragma unroll
        for (int c4 = 0; c4 < 4; c4 += 2) {
          tmem_ld_wait();
          tmem_ld32_issue(sa + (c4 + 1) * 32, vb);
          mask_chunk(va, c4);               // exp2(-inf) = 0
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            va[i] = p_of(va[i]);
            va[i + 1] = p_of(va[i + 1]);
            sum[0] += va[i];
            sum[1] += va[i + 1];
          }
          tmem_st32(sa + c4 * 32, va);
          tmem_ld_wait();
          if (c4 + 2 < 4) tmem_ld32_issue(sa + (c4 + 2) * 32, va);
          mask_chunk(vb, c4 + 1);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            vb[i] = p_of(vb[i]);
            vb[i + 1] = p_of(vb[i + 1]);
            sum[2] += vb[i];
            sum[3] += vb[i + 1];
          }
          tmem_st32(sa + (c4 + 1) * 32, vb);
        }
      }
      l = fmaf(l, alpha, (sum[0] + sum[1]) + (sum[2] + sum[3]));
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      tc_fence_before();
      arrive_p(&p_full[u]);
      if (q == 0) FS2_TRACE2(j, 3);
      if (j >= 1) accumulate(j - 1, alpha_prev);
      if (q == 0) FS2_TRACE2(j, 6);
      alpha_prev = alpha;
    }
    accumulate(n_tiles - 1, alpha_prev);
    if (dbg != 4 && qrow < len) {
      const float inv = 1.f / l;
      if (out_b != nullptr) {   // BF16 mirror of the context (A operand of the fc contraction in the bf16 mode)
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
#ifdef FS2_TRACE_BUILD
  if (threadIdx.x == 64 && blockIdx.x < 2048) {
    attn_tc::g_attn_cta_trace[blockIdx.x * 6 + 3] = attn_tc::gtimer();
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    attn_tc::g_attn_cta_trace[blockIdx.x * 6 + 4] = smid;
    attn_tc::g_attn_cta_trace[blockIdx.x * 6 + 5] = listed ? n_tiles : 0;
  }
#endif
  cluster_sync_all();   // the peer may still arrive on this CTA's barriers / read its shared memory through the pair's MMAs
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u) : "memory");
}

// `work` must have been built for 256 query rows per entry (rowops.cuh build_attention_work)
inline void launch(const float* qkv, int rows, const int32_t* starts, const int32_t* lens, const uint32_t* work,
                   const int32_t* work_count, int work_cap, float* out, cudaStream_t stream, void* out_bf16 = nullptr) {
  if (work_cap <= 0 || rows <= 0) return;
  static bool configured[64] = {};
  int dev = 0;
  FS2_CUDA_OK(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    FS2_CUDA_OK(cudaFuncSetAttribute(attention_2sm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    configured[dev & 63] = true;
  }
  const CUtensorMap tmQ = tc2::make_map(qkv, rows, LDQKV, LDQKV, BQ, true, false);
  const CUtensorMap tmK = tc2::make_map(qkv, rows, LDQKV, LDQKV, BKV / 2, true, true);
  const CUtensorMap tmV = tc2::make_map(qkv, rows, LDQKV, LDQKV, BKV, true, true, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  launch_pdl(attention_2sm_kernel, dim3(N_HEAD * work_cap * 2), dim3(THREADS), SMEM_TOTAL, stream, 2, tmQ, tmK, tmV, starts, lens,
             work, work_count, out, static_cast<__nv_bfloat16*>(out_bf16), attn_tc::debug_flag());
  FS2_LAUNCHED();
}

}  // namespace attn2
}  // namespace fs2
