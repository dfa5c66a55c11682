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
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace fs2 {
namespace attn_tc {

using namespace tc;

constexpr int BQ = 128, BKV = 64, THREADS = 192;
constexpr int THREADS2 = 224;   // attention_tc_kernel: + a second MMA-issuing warp
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

#ifdef FS2_TRACE_BUILD   // per-CTA time stamps for tools/trace_attention_ctas.py: [cta][entry, after pdl_wait, first S, exit, smid, tiles]
__device__ long long g_attn_cta_trace[2048 * 6];
__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define FS2_CTA_STAMP(k) do { if (threadIdx.x == 64 && blockIdx.x < 2048) ::fs2::attn_tc::g_attn_cta_trace[blockIdx.x * 6 + (k)] = ::fs2::attn_tc::gtimer(); } while (0)
#else
#define FS2_CTA_STAMP(k) do { } while (0)
#endif

// PAIR: the kernel runs as clusters of two CTAs that own two ADJACENT query tiles of the same (utterance, head).  Each CTA
// loads HALF of every K and V tile (32 of the 64 key rows) and TMA-multicasts it into both CTAs' shared memory, so the L2
// traffic of the kernel -- every query tile of an utterance re-reads all of its keys and values: 420 MB per launch at batch
// 64, which the per-tile trace showed as waits for K/V tiles at ~75 GB/s per SM, the chip's aggregate L2 limit -- is
// halved.  A pair whose second tile lies beyond the utterance keeps its second CTA for the loads only.
template <bool PAIR>
__global__ void __launch_bounds__(THREADS2, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    const __grid_constant__ CUtensorMap tmV, const int32_t* __restrict__ starts, const int32_t* __restrict__ lens,
                    const uint32_t* __restrict__ work, const int32_t* __restrict__ work_count, float* __restrict__ out,
                    __nv_bfloat16* __restrict__ out_b, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  FS2_CTA_STAMP(0);
#ifdef FS2_TRACE_BUILD
  const long long c_entry = clock64();
#endif
  // work item = 128 queries of one (utterance, head), taken from the longest-first work list (rowops.cuh): CTAs are
  // dispatched in blockIdx order, so the expensive items start first and the grid's tail is made of short utterances
  const int rank = PAIR ? (int)(blockIdx.x & 1) : 0;                    // == %cluster_ctarank (cluster of 2 along x)
  const int item = PAIR ? blockIdx.x >> 2 : blockIdx.x >> 1, h = PAIR ? (blockIdx.x >> 1) & 1 : blockIdx.x & 1;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_s = smem;
  // stage 2 of both rings reuses the Q region once Q lives in TMEM
  auto k_stage = [&](int s) -> uint8_t* { return s < 2 ? smem + Q_BYTES + s * K_BYTES : smem; };
  auto v_stage = [&](int s) -> uint8_t* { return s < 2 ? smem + Q_BYTES + 2 * K_BYTES + s * K_BYTES : smem + K_BYTES; };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* q_full = bars;            // [1]
  uint64_t* q_moved = bars + 1;       // [1]  Q copied to TMEM: its smem may be overwritten
  // K and V have their OWN rings: a K stage is free as soon as Q K_j^T has completed, one tile before P_j V_j releases the V
  // stage.  With one shared barrier pair (the first version) K_{j+3} could only be requested when P_j V_j was done, i.e. one
  // tile before Q K_{j+3}^T wants it (the Q K^T products run two tiles ahead of the P V products): the per-tile trace showed
  // ~1,100 cycles of waiting for K on every other tile, at a TMA latency of ~2,100 cycles against a 1,650-cycle tile.
  uint64_t* k_full = bars + 2;        // [3]
  uint64_t* k_empty = bars + 5;       // [3]  free after Q K_j^T
  uint64_t* v_full = bars + 16;       // [3]
  uint64_t* v_empty = bars + 19;      // [3]  free after P_j V_j
  uint64_t* s_full = bars + 8;        // [2]
  uint64_t* p_full = bars + 10;       // [2]  also implies that O_{j-2} has been accumulated (program order
                                      //      of the softmax threads), so P_j V_j may overwrite that buffer
  uint64_t* o_full = bars + 12;       // [2]
  uint64_t* q_free = bars + 14;       // [1]  PAIR: the Q staging area of BOTH CTAs is free (stage 2 may be multicast into it)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = warp_index(), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmKV)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
    mbar_init(q_full, 1);
    mbar_init(q_moved, 128);
    mbar_init(q_free, 256);
    for (int u = 0; u < KV_STAGES; ++u) {
      mbar_init(&k_full[u], 1);
      mbar_init(&v_full[u], 1);
      mbar_init(&k_empty[u], PAIR ? 2 : 1);    // PAIR: a stage is refilled by both CTAs: both consumers release it
      mbar_init(&v_empty[u], PAIR ? 2 : 1);
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
  if (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anything is multicast to it
  // ---- the prologue above overlaps the previous kernel's tail (programmatic dependent launch);
  // lens / starts / qkv are produced by earlier kernels of this forward
  pdl_trigger();
  pdl_wait();
  FS2_CTA_STAMP(1);
  const bool listed = item < *work_count;     // the grid is sized from a host-side bound: surplus CTAs only tear down
  const uint32_t wi = listed ? work[item] : 0u;
  // a work-list entry covers BQ queries, or 2 * BQ in the PAIR form (one tile per CTA of the cluster)
  const int b = (int)(wi >> 16), q0 = ((int)(wi & 0xFFFFu) * (PAIR ? 2 : 1) + rank) * BQ;
  const int len = listed ? lens[b] : 0;
  const int row0 = listed ? starts[b] : 0;
  const bool active = listed && q0 < len;     // PAIR: the second CTA of a pair past the utterance's end only loads
  const bool loads_only = PAIR && listed && !active;
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

  if (!active && !loads_only) {
    // nothing to do
  } else if (warp == 0) {
    // ---- TMA producer (whole warp, one elected lane issues): Q once, then K_j / V_j into 3-stage rings
    const bool leader = elect_one();
    if (leader && active) {
      mbar_expect_tx(q_full, Q_BYTES);
#pragma unroll
      for (int dc = 0; dc < 4; ++dc) tma_load_2d(q_s + dc * (BQ * 128), &tmQ, h * D_HEAD + dc * 32, row0 + q0, q_full);
    }
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      const int sk = j % KV_STAGES;
      const uint32_t par = ((j / KV_STAGES) & 1) ^ 1;
      // stage 2 lives where Q was staged: its first use waits until Q has been moved to tensor memory -- in BOTH CTAs of a
      // pair, because this CTA's multicast lands in the peer's staging area too (q_free: 128 arrivals from each CTA)
      if (j == 2) {
        if (PAIR) mbar_wait_cluster(q_free, 0);
        else mbar_wait(q_moved, 0);
      }
      uint8_t* k_s = k_stage(sk);
      uint8_t* v_s = v_stage(sk);
      const int half = rank * (BKV / 2);   // PAIR: this CTA's 32 key rows of every sub-tile, delivered to both CTAs
      mbar_wait(&k_empty[sk], par);
      if (leader) {
        mbar_expect_tx(&k_full[sk], K_BYTES);
#pragma unroll
        for (int dc = 0; dc < 4; ++dc) {
          if (PAIR) tma_load_2d_mc(k_s + dc * (BKV * 128) + half * 128, &tmKV, D_MODEL + h * D_HEAD + dc * 32, row0 + j * BKV + half,
                                   &k_full[sk], (uint16_t)0x3);
          else tma_load_2d(k_s + dc * (BKV * 128), &tmKV, D_MODEL + h * D_HEAD + dc * 32, row0 + j * BKV, &k_full[sk]);
        }
      }
      __syncwarp();
      mbar_wait(&v_empty[sk], par);
      if (leader) {
        mbar_expect_tx(&v_full[sk], K_BYTES);
#pragma unroll
        for (int dc = 0; dc < 4; ++dc) {
          if (PAIR) tma_load_2d_mc(v_s + dc * (BKV * 128) + half * 128, &tmV, 2 * D_MODEL + h * D_HEAD + dc * 32, row0 + j * BKV + half,
                                   &v_full[sk], (uint16_t)0x3);
          else tma_load_2d(v_s + dc * (BKV * 128), &tmV, 2 * D_MODEL + h * D_HEAD + dc * 32, row0 + j * BKV, &v_full[sk]);
        }
      }
      __syncwarp();
    }
  } else if (loads_only) {
    // ---- PAIR, second CTA beyond the utterance's end: it only loads its halves of the K/V tiles for the peer.  Its Q
    // staging area is free from the start (q_free), and every stage is released, in both CTAs, as soon as its bytes have
    // landed here.
    if (warp >= 2 && warp <= 5) {
      mbar_arrive(q_free);
      mbar_arrive_remote(dsmem_addr(q_free, rank ^ 1));
    } else if (warp == 6) {
      for (int j = 0; j < n_tiles; ++j) {
        const int sk = j % KV_STAGES;
        mbar_wait(&k_full[sk], (j / KV_STAGES) & 1);
        if (lane == 0) {
          mbar_arrive(&k_empty[sk]);
          mbar_arrive_remote(dsmem_addr(&k_empty[sk], rank ^ 1));
        }
        mbar_wait(&v_full[sk], (j / KV_STAGES) & 1);
        if (lane == 0) {
          mbar_arrive(&v_empty[sk]);
          mbar_arrive_remote(dsmem_addr(&v_empty[sk], rank ^ 1));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer 1: S_j = Q K_j^T.  The per-tile trace (tools/trace_attention.py) showed ONE issuing thread as the
    // critical path of the kernel: 16 + 8 MMAs and three commits cost it ~1650 cycles per 64-key tile for 1024 cycles of
    // tensor work (an N = 64 MMA is 32 cycles of work and ~26 cycles to issue).  Q K^T and P V therefore have a warp
    // each; their order on the tensor pipe is no longer program order, so Q K_j^T waits for P V_{j-2} -- which read P
    // from the same TMEM columns -- through its completion barrier.
    const bool leader = elect_one();
    constexpr uint32_t idesc_qk = idesc_tf32(BQ, BKV, 0);
    mbar_wait(q_moved, 0);
    tc_fence_after();
    for (int j = 0; j < n_tiles; ++j) {
      const int u = j & 1, sk = j % KV_STAGES;
      FS2_TRACE(j, 8);
      mbar_wait(&k_full[sk], (j / KV_STAGES) & 1);
      if (j >= 2) mbar_wait(&o_full[u], ((j - 2) >> 1) & 1);
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
        if (PAIR) umma_commit_mc(&k_empty[sk], (uint16_t)0x3); else umma_commit(&k_empty[sk]);
        FS2_TRACE(j, 11);
      }
      __syncwarp();
    }
  } else if (warp == 6) {
    // ---- MMA issuer 2: O_j = P_j V_j (A = P from tensor memory)
    const bool leader = elect_one();
    constexpr uint32_t idesc_pv = idesc_tf32(BQ, D_HEAD, 1);
    for (int j = 0; j < n_tiles; ++j) {
      const int u = j & 1;
      const uint32_t par = (j >> 1) & 1;
      const int sk = j % KV_STAGES;
      FS2_TRACE(j, 4);
      mbar_wait(&v_full[sk], (j / KV_STAGES) & 1);
      mbar_wait(&p_full[u], par);      // P_j written; O buffer u drained (see p_full above)
      FS2_TRACE(j, 5);
      tc_fence_after();
      const uint8_t* v_s = v_stage(sk);
      const uint64_t dv = umma_desc_mn(v_s, BKV * 128, 512);
      if (leader) {
#pragma unroll
        for (int k8 = 0; k8 < BKV / 8; ++k8)
          umma_tf32_ts(tmem_o + u * D_HEAD, tmem_s + u * BKV + k8 * 8, dv + (uint64_t)(k8 * (1024 >> 4)), idesc_pv, k8 != 0);
        FS2_TRACE(j, 12);
        umma_commit(&o_full[u]);
        if (PAIR) umma_commit_mc(&v_empty[sk], (uint16_t)0x3); else umma_commit(&v_empty[sk]);
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
      if (PAIR) {
        mbar_arrive(q_free);
        mbar_arrive_remote(dsmem_addr(q_free, rank ^ 1));
      }
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
      if (j == 0) FS2_CTA_STAMP(2);
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
#ifdef FS2_TRACE_BUILD
  if (threadIdx.x == 64 && blockIdx.x < 2048) {
    g_attn_cta_trace[blockIdx.x * 6 + 3] = gtimer();
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_attn_cta_trace[blockIdx.x * 6 + 4] = (long long)smid | ((clock64() - c_entry) << 16);   // SM id | cycles of the CTA's life
    g_attn_cta_trace[blockIdx.x * 6 + 5] = n_tiles;
  }
#endif
  if (PAIR) cluster_sync_all();   // the peer may still multicast into this CTA's shared memory / arrive on its barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

inline int& debug_flag() {
  static int f = 0;
  return f;
}

// Upper bound of the work-list length known on the host: sum_b ceil(len_b / q_rows) <= total_rows / q_rows + batch.
inline int work_bound(int64_t total_len, int batch, int max_len, int q_rows = BQ) {
  const int64_t a = total_len / q_rows + batch, b = (int64_t)batch * ((max_len + q_rows - 1) / q_rows);
  return (int)std::min(a, b);
}

// Query rows per work-list entry: 256 = the PAIR form (clusters of two CTAs sharing every K/V tile through multicast), chosen
// when the batch has enough work to fill the machine with pairs; 128 = one CTA per entry (single utterances: a pair
// would only add a cluster hand-shake).  The work list must be built with the same value (rowops.cuh).
inline int& pair_force_flag() {   // -1 automatic (default; FS2_ATTN_PAIR overrides), 0 never, 1 multicast pair, 2 the 2-SM kernel
  static int f = [] { const char* e = std::getenv("FS2_ATTN_PAIR"); return e != nullptr ? std::atoi(e) : -1; }();
  return f;
}
inline int query_rows_per_entry(int64_t total_len, int batch, int max_len) {
  const int force = pair_force_flag();
  if (force == 0) return BQ;
  if (force >= 1) return 2 * BQ;
  (void)total_len; (void)batch; (void)max_len;
  // Automatic: one CTA per 128-query tile.  Both paired forms measured SLOWER at config 2 (same box, per step, six decoder
  // launches): this file's multicast pair 0.40 ms, the 2-SM kernel of attention_tc2.cuh 0.48 ms, against 0.345 ms --
  // profiles/r02_attention_experiments.txt has the per-tile and per-CTA traces.  Both stay selectable (FS2_ATTN_PAIR=1 / 2,
  // debug flag 8) and are covered by tests/test_gpu_ops.py.
  return BQ;
}
// which kernel serves 256-row entries
inline bool use_two_sm(int q_rows) { return q_rows == 2 * BQ && pair_force_flag() != 1; }

template <bool PAIR>
inline void launch_t(const float* qkv, int rows, const int32_t* starts, const int32_t* lens, const uint32_t* work,
                     const int32_t* work_count, int work_cap, float* out, cudaStream_t stream, void* out_bf16) {
  static bool configured[64] = {};
  int dev = 0;
  FS2_CUDA_OK(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    FS2_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    configured[dev & 63] = true;
  }
  constexpr int KROWS = PAIR ? BKV / 2 : BKV;   // PAIR: each CTA loads half of the key rows of a tile (and multicasts it)
  const CUtensorMap tmQ = make_map(qkv, rows, LDQKV, LDQKV, BQ, true, false);
  const CUtensorMap tmKV = make_map(qkv, rows, LDQKV, LDQKV, KROWS, true, true);
  const CUtensorMap tmV = make_map(qkv, rows, LDQKV, LDQKV, KROWS, true, true, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  launch_pdl(attention_tc_kernel<PAIR>, dim3(N_HEAD * work_cap * (PAIR ? 2 : 1)), dim3(THREADS2), SMEM_TOTAL, stream, PAIR ? 2 : 1,
             tmQ, tmKV, tmV, starts, lens, work, work_count, out, static_cast<__nv_bfloat16*>(out_bf16), debug_flag());
  FS2_LAUNCHED();
}

// q_rows: query rows per work-list entry (128, or 256 = PAIR form); work / work_count / work_cap describe that list.
inline void launch(const float* qkv, int rows, const int32_t* starts, const int32_t* lens, const uint32_t* work,
                   const int32_t* work_count, int work_cap, int q_rows, float* out, cudaStream_t stream, void* out_bf16 = nullptr) {
  if (work_cap <= 0 || rows <= 0) return;
  if (q_rows == 2 * BQ) launch_t<true>(qkv, rows, starts, lens, work, work_count, work_cap, out, stream, out_bf16);
  else launch_t<false>(qkv, rows, starts, lens, work, work_count, work_cap, out, stream, out_bf16);
}

}  // namespace attn_tc
}  // namespace fs2
