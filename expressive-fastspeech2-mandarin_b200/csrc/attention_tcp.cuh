// Persistent form of the tcgen05 varlen attention of attention_tc.cuh (transformer/SubLayers.py:42-52, Modules.py:14-25):
// the same arithmetic per (utterance, head, 128-query tile) -- S = Q K^T and O = P V as kind::tf32 MMAs with Q and P in
// tensor memory, online softmax by 128 row-owner threads -- but ONE CTA per SM walks the longest-first work list and its
// pipelines run ACROSS items:
//   * Q travels through the K ring as two 64-row tiles (a Q tile is two K stages), so the producer prefetches the next
//     item's Q and first K/V tiles while the current item is still in its last key tiles (3 K stages + 3 V stages);
//   * the softmax threads move the next item's Q into tensor memory right after the P tile of the current item's LAST key
//     tile is handed over (every Q K^T of the item has completed by then), i.e. before they drain the last P V product,
//     normalise and store -- the first S tile of the next item is computed under that tail.
// What this removes (per-CTA traces of the one-CTA-per-item kernel at batch 64, profiles/r02_attention_experiments.txt): a
// CTA lived 3.3 us + 1.2 us per key tile with 0.8 us between two CTAs on an SM, for items of 6 key tiles on average.
#pragma once

#include "attention_tc.cuh"

namespace fs2 {
namespace attn_p {

using namespace tc;
using attn_tc::BKV;
using attn_tc::BQ;
using attn_tc::idesc_tf32;
using attn_tc::LDQKV;
using attn_tc::umma_desc_mn;
using attn_tc::umma_tf32_ts;

constexpr int THREADS = 224;                      // producer | Q K^T issuer | 4 softmax warps | P V issuer
constexpr int TILE_BYTES = BKV * D_HEAD * 4;      // 32 KB: 4 sub-tiles [64 rows x 128 B]
constexpr int K_STAGES = 3, V_STAGES = 3;
// Output staging: each softmax warp owns two 4 KB buffers [32 rows x 32 columns], 128-byte swizzled, and sends every
// 32-column piece of its rows with ONE 2-D TMA store (a warp whose 32 rows are not all inside the utterance -- the last
// query tile of an utterance -- stores from registers instead).  Storing from registers costs the softmax threads
// ~5,000 cycles per item (each 16-byte store instruction of a warp touches 32 different 128-byte lines), and so do
// per-row bulk copies (128 descriptors per item through one TMA unit: measured 2,000-2,500 cycles per 64 columns).
constexpr int STG_CHUNK = 32 * 128;                        // one [32 x 32] fp32 piece
constexpr int STG_OFF = (K_STAGES + V_STAGES) * TILE_BYTES;
constexpr int BAR_OFF = STG_OFF + 4 * 2 * STG_CHUNK;       // 4 warps x 2 buffers
constexpr int SMEM_TOTAL = BAR_OFF + 256;         // (dynamic shared memory starts 1024-byte aligned: checked at entry)
constexpr int TMEM_COLS = 512;                    // S0,S1: 2 x 64 | O0,O1: 2 x 128 | Q: 128

struct Item {
  int row0, q0, len, h, n_tiles;
  bool valid;
};

#ifdef FS2_TRACE_BUILD
// per-tile clock64 stamps of ONE CTA (tools/trace_attention_persistent.py): [tile g][0 softmax waits for S, 1 S there, 2 P handed
// over, 3 iteration done, 4 Q K^T issuer starts waiting, 5 its operands are there, 6 issued + committed, 7 P V issuer's operands there]
__device__ long long g_attn_p_tile_trace[64 * 8];
#define FS2_P_TILE(g, k) do { if (blockIdx.x == 5 && (threadIdx.x & 31) == 0 && (g) < 64) ::fs2::attn_p::g_attn_p_tile_trace[(g) * 8 + (k)] = clock64(); } while (0)
__device__ long long g_attn_p_item_trace[16 * 8];   // per item of that CTA: stamps around the item's tail (see the tool)
#define FS2_P_ITEM(k, i) do { if (blockIdx.x == 5 && threadIdx.x == 64 && (k) < 16) ::fs2::attn_p::g_attn_p_item_trace[(k) * 8 + (i)] = clock64(); } while (0)
#define FS2_P_STAMP(k) do { if (threadIdx.x == 64 && blockIdx.x < 2048) ::fs2::attn_tc::g_attn_cta_trace[blockIdx.x * 6 + (k)] = ::fs2::attn_tc::gtimer(); } while (0)
#else
#define FS2_P_STAMP(k) do { } while (0)
#define FS2_P_TILE(g, k) do { } while (0)
#define FS2_P_ITEM(k, i) do { } while (0)
#endif

__global__ void __launch_bounds__(THREADS, 1)
attention_tcp_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmV,
                     const __grid_constant__ CUtensorMap tmO, const int32_t* __restrict__ starts, const int32_t* __restrict__ lens, const uint32_t* __restrict__ work,
                     const int32_t* __restrict__ work_count, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  FS2_P_STAMP(0);
#ifdef FS2_TRACE_BUILD
  const long long c_entry = clock64();
  int traced_tiles = 0;
#endif
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // the swizzled tiles need 1024-byte alignment and there is no room for slack
  auto k_stage = [&](int s) -> uint8_t* { return smem + s * TILE_BYTES; };
  auto v_stage = [&](int s) -> uint8_t* { return smem + (K_STAGES + s) * TILE_BYTES; };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* k_full = bars;             // [3]  K tile (or half a Q tile) has landed
  uint64_t* k_empty = bars + 4;        // [3]  free after Q K_j^T (a Q half: after the copy to tensor memory)
  uint64_t* v_full = bars + 8;         // [3]
  uint64_t* v_empty = bars + 11;       // [3]  free after P_j V_j
  uint64_t* s_full = bars + 14;        // [2]
  uint64_t* p_full = bars + 16;        // [2]  P written; also: the O buffer of this parity has been drained (program order)
  uint64_t* o_full = bars + 18;        // [2]
  uint64_t* q_moved = bars + 20;       // [1]  this item's Q is in tensor memory (one phase per item)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = warp_index(), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmQK)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmO)) : "memory");
    for (int u = 0; u < K_STAGES; ++u) {
      mbar_init(&k_full[u], 1);
      mbar_init(&k_empty[u], 1);
    }
    for (int u = 0; u < V_STAGES; ++u) {
      mbar_init(&v_full[u], 1);
      mbar_init(&v_empty[u], 1);
    }
    for (int u = 0; u < 2; ++u) {
      mbar_init(&s_full[u], 1);
      mbar_init(&p_full[u], 128);
      mbar_init(&o_full[u], 1);
    }
    mbar_init(q_moved, 128);
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
  pdl_wait();   // lens / starts / work list / qkv are produced by earlier kernels of this forward
  FS2_P_STAMP(1);

  // Items of this CTA: the list is longest-first, (tile, head) pairs are dealt out in snake order (round k runs
  // 0..G-1, round k+1 runs G-1..0) so that every CTA gets a similar number of key tiles.
  const int n_items = N_HEAD * *work_count;
  const int G = (int)gridDim.x, cta = (int)blockIdx.x;
  auto load_item = [&](int k) {
    Item it;
    const int idx = k * G + ((k & 1) ? G - 1 - cta : cta);
    it.valid = idx < n_items;
    const uint32_t wi = it.valid ? work[idx >> 1] : 0u;
    const int b = (int)(wi >> 16);
    it.h = idx & 1;
    it.q0 = (int)(wi & 0xFFFFu) * BQ;
    it.len = it.valid ? lens[b] : 0;
    it.row0 = it.valid ? starts[b] : 0;
    it.n_tiles = (it.len + BKV - 1) / BKV;
    return it;
  };
  const uint32_t tmem_s = tmem_base;          // + u*64
  const uint32_t tmem_o = tmem_base + 128;    // + u*128
  const uint32_t tmem_q = tmem_base + 384;    // 128 columns

  if (warp == 0) {
    // ---- TMA producer: per item  Q rows 0..63 | Q rows 64..127 (K ring) | K_0 V_0 | K_1 V_1 | ...
    const bool leader = elect_one();
    int kc = 0, vc = 0;   // K / V ring positions (stage = position % stages, phase = position / stages)
    Item it = load_item(0);
    for (int k = 0; it.valid; ++k) {
      const Item nx = load_item(k + 1);
      for (int j = -2; j < it.n_tiles; ++j) {
        const int sk = kc % K_STAGES;
        mbar_wait(&k_empty[sk], ((kc / K_STAGES) & 1) ^ 1);
        if (leader) {
          mbar_expect_tx(&k_full[sk], TILE_BYTES);
          const int col = (j < 0 ? 0 : D_MODEL) + it.h * D_HEAD;
          const int row = it.row0 + (j < 0 ? it.q0 + (j + 2) * BKV : j * BKV);
#pragma unroll
          for (int dc = 0; dc < 4; ++dc) tma_load_2d(k_stage(sk) + dc * (BKV * 128), &tmQK, col + dc * 32, row, &k_full[sk]);
        }
        __syncwarp();
        ++kc;
        if (j < 0) continue;
        const int sv = vc % V_STAGES;
        mbar_wait(&v_empty[sv], ((vc / V_STAGES) & 1) ^ 1);
        if (leader) {
          mbar_expect_tx(&v_full[sv], TILE_BYTES);
#pragma unroll
          for (int dc = 0; dc < 4; ++dc)
            tma_load_2d(v_stage(sv) + dc * (BKV * 128), &tmV, 2 * D_MODEL + it.h * D_HEAD + dc * 32, it.row0 + j * BKV, &v_full[sv]);
        }
        __syncwarp();
        ++vc;
      }
      it = nx;
    }
  } else if (warp == 1) {
    // ---- MMA issuer 1: S_g = Q K_g^T (A = Q from tensor memory); g counts key tiles over all items of this CTA
    const bool leader = elect_one();
    constexpr uint32_t idesc_qk = idesc_tf32(BQ, BKV, 0);
    int kc = 0, g = 0;
    Item it = load_item(0);
    for (int k = 0; it.valid; ++k) {
      const Item nx = load_item(k + 1);
      mbar_wait(q_moved, k & 1);
      tc_fence_after();
      // the two ring stages that carried Q are free: every softmax thread has read its row (generic-proxy reads fenced
      // against the async proxy before its arrival on q_moved)
      if (leader) {
        mbar_arrive(&k_empty[kc % K_STAGES]);
        mbar_arrive(&k_empty[(kc + 1) % K_STAGES]);
      }
      __syncwarp();
      kc += 2;
      for (int j = 0; j < it.n_tiles; ++j, ++g, ++kc) {
        const int u = g & 1, sk = kc % K_STAGES;
        FS2_P_TILE(g, 4);
        mbar_wait(&k_full[sk], (kc / K_STAGES) & 1);
        if (g >= 2) mbar_wait(&o_full[u], ((g - 2) >> 1) & 1);   // P V_{g-2} has read P from these columns
        FS2_P_TILE(g, 5);
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
          umma_commit(&s_full[u]);
          umma_commit(&k_empty[sk]);
        }
        __syncwarp();
        FS2_P_TILE(g, 6);
      }
      it = nx;
    }
  } else if (warp == 6) {
    // ---- MMA issuer 2: O_g = P_g V_g (A = P from tensor memory)
    const bool leader = elect_one();
    constexpr uint32_t idesc_pv = idesc_tf32(BQ, D_HEAD, 1);
    int vc = 0, g = 0;
    Item it = load_item(0);
    for (int k = 0; it.valid; ++k) {
      const Item nx = load_item(k + 1);
      for (int j = 0; j < it.n_tiles; ++j, ++g, ++vc) {
        const int u = g & 1, sv = vc % V_STAGES;
        mbar_wait(&v_full[sv], (vc / V_STAGES) & 1);
        mbar_wait(&p_full[u], (g >> 1) & 1);      // P_g written; O buffer u drained
        FS2_P_TILE(g, 7);
        tc_fence_after();
        const uint64_t dv = umma_desc_mn(v_stage(sv), BKV * 128, 512);
        if (leader) {
#pragma unroll
          for (int k8 = 0; k8 < BKV / 8; ++k8)
            umma_tf32_ts(tmem_o + u * D_HEAD, tmem_s + u * BKV + k8 * 8, dv + (uint64_t)(k8 * (1024 >> 4)), idesc_pv, k8 != 0);
          umma_commit(&o_full[u]);
          umma_commit(&v_empty[sv]);
        }
        __syncwarp();
      }
      it = nx;
    }
  } else {
    // ---- softmax + accumulation: thread = query row = TMEM lane
    const int q = warp & 3;
    const int r = q * 32 + lane;                                 // row of the query tile
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const float c = 1.4426950408889634f / sqrtf((float)D_HEAD);  // log2(e) / temperature
    int kc = 0, g = 0;

    // Q (TF32-rounded by TMA) of the item whose K-ring positions start at kq: swizzled smem -> TMEM
    auto move_q = [&](int kq) {
      const int pos = kq + (r >> 6), sk = pos % K_STAGES;
      mbar_wait(&k_full[kq % K_STAGES], (kq / K_STAGES) & 1);          // both halves: the warp's rows live in one of them, but
      mbar_wait(&k_full[(kq + 1) % K_STAGES], ((kq + 1) / K_STAGES) & 1);   // q_moved must mean that both have been consumed
      const int rr = r & 63;
      const uint32_t qa = smem_u32(k_stage(sk)) + rr * 128;
      const uint32_t sx = (uint32_t)(rr & 7) << 4;
#pragma unroll 1
      for (int dc = 0; dc < 4; dc += 2) {   // two 32-column pieces per round: 16 shared-memory loads in flight
        float va[32], vb[32];
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          float4 t4, u4;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n"
                       : "=f"(t4.x), "=f"(t4.y), "=f"(t4.z), "=f"(t4.w)
                       : "r"(qa + dc * (BKV * 128) + ((cc << 4) ^ sx)));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n"
                       : "=f"(u4.x), "=f"(u4.y), "=f"(u4.z), "=f"(u4.w)
                       : "r"(qa + (dc + 1) * (BKV * 128) + ((cc << 4) ^ sx)));
          va[cc * 4] = t4.x; va[cc * 4 + 1] = t4.y; va[cc * 4 + 2] = t4.z; va[cc * 4 + 3] = t4.w;
          vb[cc * 4] = u4.x; vb[cc * 4 + 1] = u4.y; vb[cc * 4 + 2] = u4.z; vb[cc * 4 + 3] = u4.w;
        }
        tmem_st32(tmem_q + lane_sel + dc * 32, va);
        tmem_st32(tmem_q + lane_sel + dc * 32 + 32, vb);
      }
      // (no proxy fence: the shared-memory loads have completed -- their values have been stored to tensor memory -- before
      // this arrival, and the stage is only rewritten by a TMA load issued after the barrier chain q_moved -> k_empty)
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      tc_fence_before();
      mbar_arrive(q_moved);
    };

    uint8_t* stg = smem + STG_OFF;
    const uint32_t sbuf = smem_u32(stg) + q * (2 * STG_CHUNK);   // this warp's two [32 x 32] staging buffers
    const uint32_t swz = (uint32_t)(lane & 7) << 4;
    // two [32 rows x 32 columns] pieces of this warp's rows (columns col .. col + 63): registers -> swizzled staging ->
    // two 2-D TMA stores behind ONE proxy fence (fence.proxy.async is a MEMBAR.ALL.CTA: a few hundred cycles)
    auto tma_send2 = [&](const float (&va)[32], const float (&vb)[32], int col, int row) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");   // the previous pair has read both buffers
      __syncwarp();
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(sbuf + lane * 128 + ((cc << 4) ^ swz)), "f"(va[cc * 4]),
                     "f"(va[cc * 4 + 1]), "f"(va[cc * 4 + 2]), "f"(va[cc * 4 + 3])
                     : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(sbuf + STG_CHUNK + lane * 128 + ((cc << 4) ^ swz)),
                     "f"(vb[cc * 4]), "f"(vb[cc * 4 + 1]), "f"(vb[cc * 4 + 2]), "f"(vb[cc * 4 + 3])
                     : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic writes -> visible to the TMA stores
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                         reinterpret_cast<uint64_t>(&tmO)),
                     "r"(sbuf), "r"(col), "r"(row)
                     : "memory");
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                         reinterpret_cast<uint64_t>(&tmO)),
                     "r"(sbuf + STG_CHUNK), "r"(col + 32), "r"(row)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      }
    };
    // The second 64 columns of an item's output wait in tensor memory (in the O buffer the item's last product was drained
    // from, which the next item's SECOND key tile is the first to write again) and are sent after the next item's first P
    // tile has been handed over: sent at once they would wait ~1,000 cycles for the TMA unit to get through the queued
    // K/V loads and read the two staging buffers.
    bool parked = false;
    uint32_t park_t = 0;
    int park_col = 0, park_row = 0;
    auto flush_parked = [&]() {
      float v0[32], v1[32];
      tmem_ld32_issue(park_t, v0);
      tmem_ld32_issue(park_t + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      tma_send2(v0, v1, park_col, park_row);
      parked = false;
    };
    Item it = load_item(0);
    if (it.valid) move_q(0);
    for (int k = 0; it.valid; ++k) {
      const Item nx = load_item(k + 1);
      kc += 2;
      const int qrow = it.q0 + r;                                // row inside the utterance
      const int len = it.len, n_tiles = it.n_tiles;
      float m = -INFINITY, l = 0.f, alpha_prev = 0.f;
      float o[D_HEAD];
#pragma unroll
      for (int i = 0; i < D_HEAD; ++i) o[i] = 0.f;

      auto accumulate = [&](int gg, float alpha) {
        const int u = gg & 1;
        mbar_wait(&o_full[u], (gg >> 1) & 1);
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
      for (int j = 0; j < n_tiles; ++j, ++g) {
        const int u = g & 1;
        if (warp == 2) FS2_P_TILE(g, 0);
        mbar_wait(&s_full[u], (g >> 1) & 1);
        if (warp == 2) FS2_P_TILE(g, 1);
#ifdef FS2_TRACE_BUILD
        if (g == 0) FS2_P_STAMP(2);
        ++traced_tiles;
#endif
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
        // P is rounded to TF32 (nearest, ties away) with integer arithmetic: (bits + 0x1000) & ~0x1fff runs on
        // the ALU pipe, whereas cvt.rna.tf32 shares the XU pipe with ex2 and would double its load.  p is in [0, 1].
        auto p_of = [&](float s) {
          const uint32_t bits = (__float_as_uint(ex2_approx(fmaf(s, c, -mc))) + 0x1000u) & 0xFFFFE000u;
          return __uint_as_float(bits);
        };
#pragma unroll
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
        if (warp == 2) FS2_P_TILE(g, 2);
        if (parked) flush_parked();   // (warp-uniform) the previous item's second half
        // The S tile of this item's last key tile has been read, so every Q K^T of the item has completed and the Q
        // columns are free: hand the next item's Q over now, and drain the last two P V products under its first Q K^T.
        if (j == n_tiles - 1 && nx.valid) { FS2_P_ITEM(k, 6); move_q(kc + n_tiles); FS2_P_ITEM(k, 7); }
        if (j >= 1) accumulate(g - 1, alpha_prev);
        if (warp == 2) FS2_P_TILE(g, 3);
        alpha_prev = alpha;
      }
      kc += n_tiles;
      {
        // Drain the last P V product, normalise and send the rows: the second half of the product is loaded from tensor
        // memory while the first half is staged and handed to the TMA unit.
        const int u = (g - 1) & 1;
        FS2_P_ITEM(k, 0);
        mbar_wait(&o_full[u], ((g - 1) >> 1) & 1);
        FS2_P_ITEM(k, 1);
        tc_fence_after();
        float v0[32], v1[32];
        tmem_ld32_issue(tmem_o + lane_sel + u * D_HEAD, v0);
        tmem_ld32_issue(tmem_o + lane_sel + u * D_HEAD + 32, v1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          o[i] = fmaf(o[i], alpha_prev, v0[i]);
          o[32 + i] = fmaf(o[32 + i], alpha_prev, v1[i]);
        }
        tmem_ld32_issue(tmem_o + lane_sel + u * D_HEAD + 64, v0);
        tmem_ld32_issue(tmem_o + lane_sel + u * D_HEAD + 96, v1);
        const float inv = 1.f / l;
        // rows of this warp inside the utterance: 32 -> TMA stores, 1..31 -> stores from registers, 0 -> nothing
        const int n_valid = min(max(len - (it.q0 + q * 32), 0), 32);
        const int o_col = it.h * D_HEAD, o_row = it.row0 + it.q0 + q * 32;
        auto send64 = [&](int c0) {   // columns [c0, c0 + 64) of this thread's row
          if (n_valid == 32) {
            float va[32], vb[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              va[i] = o[c0 + i] * inv;
              vb[i] = o[c0 + 32 + i] * inv;
            }
            tma_send2(va, vb, o_col + c0, o_row);
          } else if (qrow < len) {
            float* dst = out + (size_t)(it.row0 + qrow) * D_MODEL + o_col + c0;
#pragma unroll
            for (int i = 0; i < 64; i += 4)
              *reinterpret_cast<float4*>(dst + i) = make_float4(o[c0 + i] * inv, o[c0 + i + 1] * inv, o[c0 + i + 2] * inv, o[c0 + i + 3] * inv);
          }
        };
        FS2_P_ITEM(k, 2);
        send64(0);
        FS2_P_ITEM(k, 3);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          o[64 + i] = fmaf(o[64 + i], alpha_prev, v0[i]);
          o[96 + i] = fmaf(o[96 + i], alpha_prev, v1[i]);
        }
        FS2_P_ITEM(k, 4);
        if (n_valid == 32 && nx.valid) {   // park the second half in the drained O buffer (see flush_parked)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v0[i] = o[64 + i] * inv;
            v1[i] = o[96 + i] * inv;
          }
          park_t = tmem_o + lane_sel + u * D_HEAD;
          tmem_st32(park_t, v0);
          tmem_st32(park_t + 32, v1);
          asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
          park_col = o_col + 64;
          park_row = o_row;
          parked = true;
          tc_fence_before();
        } else {
          tc_fence_before();   // ordered before this thread's next p_full arrive, which releases the O buffer
          send64(64);
        }
        FS2_P_ITEM(k, 5);
      }
      it = nx;
    }
    asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");   // the stores are COMPLETE before the CTA gives up its shared memory
  }
  tc_fence_before();
  __syncthreads();
#ifdef FS2_TRACE_BUILD
  if (threadIdx.x == 64 && blockIdx.x < 2048) {
    attn_tc::g_attn_cta_trace[blockIdx.x * 6 + 3] = attn_tc::gtimer();
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    attn_tc::g_attn_cta_trace[blockIdx.x * 6 + 4] = (long long)smid | ((clock64() - c_entry) << 16);
    attn_tc::g_attn_cta_trace[blockIdx.x * 6 + 5] = traced_tiles;
  }
#endif
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}

// FS2_ATTN_PERSISTENT / debug flag 10: 3 (default) = the persistent kernel with separate softmax and accumulate warpgroups
// (attention_tcq.cuh); 2 = this kernel (one group of row threads) at every size; 1 = this kernel only when the work list can
// exceed one item per SM; 0 = one CTA per item (attention_tc.cuh)
inline int& enabled_flag() {
  static int f = [] { const char* e = std::getenv("FS2_ATTN_PERSISTENT"); return e != nullptr ? std::atoi(e) : 3; }();
  return f;
}
inline bool use_persistent(int work_cap, int sms) { return enabled_flag() >= 2 || (enabled_flag() == 1 && N_HEAD * work_cap > sms); }

inline void launch(const float* qkv, int rows, const int32_t* starts, const int32_t* lens, const uint32_t* work,
                   const int32_t* work_count, int work_cap, float* out, cudaStream_t stream, int sms) {
  if (work_cap <= 0 || rows <= 0) return;
  static bool configured[64] = {};
  int dev = 0;
  FS2_CUDA_OK(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    FS2_CUDA_OK(cudaFuncSetAttribute(attention_tcp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    configured[dev & 63] = true;
  }
  const CUtensorMap tmQK = make_map(qkv, rows, LDQKV, LDQKV, BKV, true, true);
  const CUtensorMap tmV = make_map(qkv, rows, LDQKV, LDQKV, BKV, true, true, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  const CUtensorMap tmO = make_map(out, rows, D_MODEL, D_MODEL, 32, false, false);
  static const int grid_cap = [] { const char* e = std::getenv("FS2_ATTN_GRID"); return e != nullptr ? std::atoi(e) : 0; }();   // experiment
  if (grid_cap > 0) sms = std::min(sms, grid_cap);
  launch_pdl(attention_tcp_kernel, dim3(std::min(sms, N_HEAD * work_cap)), dim3(THREADS), SMEM_TOTAL, stream, 1, tmQK, tmV, tmO, starts,
             lens, work, work_count, out);
  FS2_LAUNCHED();
}

}  // namespace attn_p
}  // namespace fs2
