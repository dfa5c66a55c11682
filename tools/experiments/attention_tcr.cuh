// NOT BUILT -- kept as the record of a measured experiment.  To try it again: copy next to
// attention_tcq.cuh, include it from fs2_api.cu and call attn_r::launch where attn_q::launch is called.  Result on B200,
// config 2, same box, two runs each: dec.attention 0.2606 / 0.2599 ms per step against 0.2476 / 0.2478 for attention_tcq.cuh
// (bit-identical outputs, all attention tests green): the single O buffer makes P V_g -> drain -> P V_{g+1} the chain.
// attention_tcq.cuh with THREE score buffers and ONE product buffer in tensor memory (transformer/SubLayers.py:42-52,
// Modules.py:14-25; same products, same order of floating-point operations: bit-identical outputs).  With two score buffers
// the kernel was bound by the chain S_g -> softmax -> P V_g -> Q K_{g+2}^T (which overwrites P_g) -> S_{g+2}: ~2,500 cycles per
// two key tiles.  Here Q K_{g+3}^T is the first to overwrite P_g, the accumulate group drains the single O buffer right behind
// each product (P V_{g+1} waits for that drain), and the per-tile (alpha, row sum) pair travels through four two-column slots
// of the 64 tensor-memory columns this layout leaves free instead of through shared memory.
//   TMEM columns: S0 S1 S2 (3 x 64) | O (128) | Q (128) | (alpha, sum) slots (4 x 2)
// Roles as in attention_tcq.cuh: warp 0 producer, warp 1 Q K^T issuer, warps 2-5 softmax group, warp 6 P V issuer,
// warps 7-10 accumulate group (o, l, next item's Q, output).
#pragma once

#include "attention_tcq.cuh"

namespace fs2 {
namespace attn_r {

using namespace tc;
using attn_p::Item;
using attn_tc::BKV;
using attn_tc::BQ;
using attn_tc::idesc_tf32;
using attn_tc::LDQKV;
using attn_tc::umma_desc_mn;
using attn_tc::umma_tf32_ts;

constexpr int THREADS = 352;                      // 11 warps (see above)
constexpr int TILE_BYTES = BKV * D_HEAD * 4;      // 32 KB: 4 sub-tiles [64 rows x 128 B]
constexpr int K_STAGES = 3, V_STAGES = 3;
constexpr int STG_CHUNK = 32 * 128;               // one [32 x 32] fp32 piece of the output staging (two per accumulate warp)
constexpr int STG_OFF = (K_STAGES + V_STAGES) * TILE_BYTES;
constexpr int BAR_OFF = STG_OFF + 4 * 2 * STG_CHUNK;
constexpr int SMEM_TOTAL = BAR_OFF + 256;         // (dynamic shared memory starts 1024-byte aligned: checked at entry)
constexpr int TMEM_COLS = 512;                    // S0,S1,S2: 3 x 64 | O: 128 | Q: 128 | 4 x 2 (alpha, sum)
constexpr int NS_BUF = 3, NAB = 4;
static_assert(SMEM_TOTAL <= 232448, "shared memory budget");

__device__ __forceinline__ void tmem_st2(uint32_t taddr, float a, float b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};\n" ::"r"(taddr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b))
               : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float& a, float& b) {
  uint32_t x, y;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n" : "=r"(x), "=r"(y) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  a = __uint_as_float(x);
  b = __uint_as_float(y);
}

#ifdef FS2_TRACE_BUILD
#define FS2_R_STAMP(k) do { if (threadIdx.x == 64 && blockIdx.x < 2048) ::fs2::attn_tc::g_attn_cta_trace[blockIdx.x * 6 + (k)] = ::fs2::attn_tc::gtimer(); } while (0)
// per-tile clock64 stamps of ONE CTA, same array and columns as attention_tcp.cuh: 0 softmax waits for S, 1 S there, 2 P handed
// over, 3 accumulate group done with the tile, 4 Q K^T issuer starts waiting, 5 its operands are there, 6 issued, 7 P V operands there
#define FS2_R_TILE(g, k) do { if (blockIdx.x == 5 && (threadIdx.x & 31) == 0 && (g) < 64) ::fs2::attn_p::g_attn_p_tile_trace[(g) * 8 + (k)] = clock64(); } while (0)
#else
#define FS2_R_STAMP(k) do { } while (0)
#define FS2_R_TILE(g, k) do { } while (0)
#endif

__global__ void __launch_bounds__(THREADS, 1)
attention_tcr_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmV,
                     const __grid_constant__ CUtensorMap tmO, const int32_t* __restrict__ starts,
                     const int32_t* __restrict__ lens, const uint32_t* __restrict__ work,
                     const int32_t* __restrict__ work_count, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  FS2_R_STAMP(0);
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
  uint64_t* s_full = bars + 14;        // [3]  S_g is in tensor memory (buffer g % 3)
  uint64_t* p_full = bars + 17;        // [3]  P_g written over S_g, (alpha_g, sum_g) written to their slot
  uint64_t* pv_done = bars + 20;       // [3]  P V_g has completed: the score buffer g % 3 may be overwritten
  uint64_t* o_full = bars + 23;        // [1]  O_g = P_g V_g is in tensor memory (one phase per key tile)
  uint64_t* o_free = bars + 24;        // [1]  the accumulate group has read O_g and (alpha_g, sum_g) (one phase per key tile)
  uint64_t* q_moved = bars + 25;       // [1]  this item's Q is in tensor memory (one phase per item)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

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
    for (int u = 0; u < NS_BUF; ++u) {
      mbar_init(&s_full[u], 1);
      mbar_init(&p_full[u], 128);
      mbar_init(&pv_done[u], 1);
    }
    mbar_init(o_full, 1);
    mbar_init(o_free, 128);
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
  FS2_R_STAMP(1);

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
  const uint32_t tmem_s = tmem_base;          // + (g % 3) * 64
  const uint32_t tmem_o = tmem_base + 192;    // 128 columns
  const uint32_t tmem_q = tmem_base + 320;    // 128 columns
  const uint32_t tmem_ab = tmem_base + 448;   // + (g % 4) * 2

  if (warp == 0) {
    // ---- TMA producer
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
    // ---- MMA issuer 1: S_g = Q K_g^T; g counts key tiles over all items of this CTA
    const bool leader = elect_one();
    constexpr uint32_t idesc_qk = idesc_tf32(BQ, BKV, 0);
    int kc = 0, g = 0;
    Item it = load_item(0);
    for (int k = 0; it.valid; ++k) {
      const Item nx = load_item(k + 1);
      mbar_wait(q_moved, k & 1);
      tc_fence_after();
      // the two ring stages that carried Q are free: every row has been read before its thread arrived on q_moved
      if (leader) {
        mbar_arrive(&k_empty[kc % K_STAGES]);
        mbar_arrive(&k_empty[(kc + 1) % K_STAGES]);
      }
      __syncwarp();
      kc += 2;
      for (int j = 0; j < it.n_tiles; ++j, ++g, ++kc) {
        const int u = g % NS_BUF, sk = kc % K_STAGES;
        FS2_R_TILE(g, 4);
        mbar_wait(&k_full[sk], (kc / K_STAGES) & 1);
        if (g >= NS_BUF) mbar_wait(&pv_done[u], ((g - NS_BUF) / NS_BUF) & 1);   // P V_{g-3} has read P from these columns
        FS2_R_TILE(g, 5);
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
        FS2_R_TILE(g, 6);
      }
      it = nx;
    }
  } else if (warp == 6) {
    // ---- MMA issuer 2: O_g = P_g V_g (A = P from tensor memory), into the O buffer the accumulate group has drained
    const bool leader = elect_one();
    constexpr uint32_t idesc_pv = idesc_tf32(BQ, D_HEAD, 1);
    int vc = 0, g = 0;
    Item it = load_item(0);
    for (int k = 0; it.valid; ++k) {
      const Item nx = load_item(k + 1);
      for (int j = 0; j < it.n_tiles; ++j, ++g, ++vc) {
        const int u = g % NS_BUF, sv = vc % V_STAGES;
        mbar_wait(&v_full[sv], (vc / V_STAGES) & 1);
        mbar_wait(&p_full[u], (g / NS_BUF) & 1);
        if (g >= 1) mbar_wait(o_free, (g - 1) & 1);   // the accumulate group has drained O_{g-1}
        FS2_R_TILE(g, 7);
        tc_fence_after();
        const uint64_t dv = umma_desc_mn(v_stage(sv), BKV * 128, 512);
        if (leader) {
#pragma unroll
          for (int k8 = 0; k8 < BKV / 8; ++k8)
            umma_tf32_ts(tmem_o, tmem_s + u * BKV + k8 * 8, dv + (uint64_t)(k8 * (1024 >> 4)), idesc_pv, k8 != 0);
          umma_commit(o_full);
          umma_commit(&pv_done[u]);
          umma_commit(&v_empty[sv]);
        }
        __syncwarp();
      }
      it = nx;
    }
  } else if (warp >= 2 && warp <= 5) {
    // ---- softmax group: thread = query row = TMEM lane
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const float c = 1.4426950408889634f / sqrtf((float)D_HEAD);  // log2(e) / temperature
    int g = 0;
    Item it = load_item(0);
    for (int k = 0; it.valid; ++k) {
      const Item nx = load_item(k + 1);
      const int len = it.len, n_tiles = it.n_tiles;
      float m = -INFINITY;   // running row maximum of the RAW scores; exp2 arguments are s*c - m*c (one FFMA each)
      for (int j = 0; j < n_tiles; ++j, ++g) {
        const int u = g % NS_BUF;
        if (warp == 2) FS2_R_TILE(g, 0);
        mbar_wait(&s_full[u], (g / NS_BUF) & 1);
        if (warp == 2) FS2_R_TILE(g, 1);
#ifdef FS2_TRACE_BUILD
        if (g == 0) FS2_R_STAMP(2);
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
        tmem_st32(tmem_s + lane_sel + u * BKV, s0);
        tmem_st32(tmem_s + lane_sel + u * BKV + 32, s1);
        // (alpha_g, sum_g) for the accumulate group: slot g % 4 was last read for tile g - 4, and that read precedes
        // P V_{g-3}, which precedes Q K_g^T
        tmem_st2(tmem_ab + lane_sel + (g % NAB) * 2, alpha, (sum[0] + sum[1]) + (sum[2] + sum[3]));
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        tc_fence_before();
        mbar_arrive(&p_full[u]);
        if (warp == 2) FS2_R_TILE(g, 2);
      }
      it = nx;
    }
  } else {
    // ---- accumulate group (warps 7-10): thread = query row = TMEM lane (a warp reaches the lane quarter warp % 4)
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    int kc = 0, g = 0;

    // Q (TF32-rounded by TMA) of the item whose K-ring positions start at kq: swizzled smem -> TMEM
    auto move_q = [&](int kq) {
      const int pos = kq + (r >> 6), sk = pos % K_STAGES;
      mbar_wait(&k_full[kq % K_STAGES], (kq / K_STAGES) & 1);              // both halves: q_moved must mean that
      mbar_wait(&k_full[(kq + 1) % K_STAGES], ((kq + 1) / K_STAGES) & 1);   // both have been consumed
      const int rr = r & 63;
      const uint32_t qa = smem_u32(k_stage(sk)) + rr * 128;
      const uint32_t sx = (uint32_t)(rr & 7) << 4;
#pragma unroll 1
      for (int dc = 0; dc < 4; ++dc) {
        float v[32];
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          float4 t4;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n"
                       : "=f"(t4.x), "=f"(t4.y), "=f"(t4.z), "=f"(t4.w)
                       : "r"(qa + dc * (BKV * 128) + ((cc << 4) ^ sx)));
          v[cc * 4] = t4.x; v[cc * 4 + 1] = t4.y; v[cc * 4 + 2] = t4.z; v[cc * 4 + 3] = t4.w;
        }
        tmem_st32(tmem_q + lane_sel + dc * 32, v);
      }
      // (no proxy fence: the shared-memory loads have completed -- their values have been stored to tensor memory -- before
      // this arrival, and the stage is only rewritten by a TMA load issued after the barrier chain q_moved -> k_empty)
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      tc_fence_before();
      mbar_arrive(q_moved);
    };

    const uint32_t sbuf = smem_u32(smem + STG_OFF) + q * (2 * STG_CHUNK);   // this warp's two [32 x 32] staging buffers
    const uint32_t swz = (uint32_t)(lane & 7) << 4;
    int n_sent = 0;   // pieces this warp has sent (buffer = n_sent & 1)
    // one [32 rows x 32 columns] piece of this warp's rows: registers -> swizzled staging -> 2-D TMA store
    auto tma_send = [&](const float (&v)[32], int col, int row) {
      const uint32_t sb = sbuf + (n_sent & 1) * STG_CHUNK;
      ++n_sent;
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");   // the store before last has read this buffer
      __syncwarp();
#pragma unroll
      for (int cc = 0; cc < 8; ++cc)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(sb + lane * 128 + ((cc << 4) ^ swz)), "f"(v[cc * 4]),
                     "f"(v[cc * 4 + 1]), "f"(v[cc * 4 + 2]), "f"(v[cc * 4 + 3])
                     : "memory");
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic writes -> visible to the TMA store
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                         reinterpret_cast<uint64_t>(&tmO)),
                     "r"(sb), "r"(col), "r"(row)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      }
    };

    Item it = load_item(0);
    if (it.valid) move_q(0);
    for (int k = 0; it.valid; ++k) {
      const Item nx = load_item(k + 1);
      kc += 2;
      const int n_tiles = it.n_tiles;
      float l = 0.f;
      float o[D_HEAD];
#pragma unroll
      for (int i = 0; i < D_HEAD; ++i) o[i] = 0.f;
      for (int j = 0; j < n_tiles; ++j, ++g) {
        if (nx.valid && j == max(n_tiles - 2, 0)) {
          // Once the S tile of the item's LAST key tile exists, every Q K^T of the item has completed and the Q columns are
          // free: the next item's Q goes in then, and its first score tiles are computed under this item's tail.  That S tile
          // is due about when this group has drained the product of key tile n - 3 (Q K_{n-1}^T is issued when P V_{n-3}
          // completes), so the move sits BEFORE the drains of the last two products: behind them the softmax group waited
          // ~1,600 cycles for the next item's first score tile.
          const int gl = g + (n_tiles - 1 - j);   // the item's last key tile
          mbar_wait(&s_full[gl % NS_BUF], (gl / NS_BUF) & 1);
          tc_fence_after();
          move_q(kc + n_tiles);
        }
        mbar_wait(o_full, g & 1);
        tc_fence_after();
        float2 ab;   // (alpha_g, row sum of P_g)
        tmem_ld2(tmem_ab + lane_sel + (g % NAB) * 2, ab.x, ab.y);
#pragma unroll
        for (int c0 = 0; c0 < D_HEAD; c0 += 32) {
          float v[32];
          tmem_ld32_issue(tmem_o + lane_sel + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c0 + i] = fmaf(o[c0 + i], ab.x, v[i]);
        }
        l = fmaf(l, ab.x, ab.y);
        tc_fence_before();
        mbar_arrive(o_free);
        if (warp == 8) FS2_R_TILE(g, 3);
      }
      kc += n_tiles;
      {
        // normalise and send this warp's rows: 32 inside the utterance -> TMA stores, 1..31 -> from registers, 0 -> nothing
        const float inv = 1.f / l;
        const int qrow = it.q0 + r;
        const int n_valid = min(max(it.len - (it.q0 + q * 32), 0), 32);
        const int o_col = it.h * D_HEAD, o_row = it.row0 + it.q0 + q * 32;
        if (n_valid == 32) {
#pragma unroll
          for (int c0 = 0; c0 < D_HEAD; c0 += 32) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = o[c0 + i] * inv;
            tma_send(v, o_col + c0, o_row);
          }
        } else if (qrow < it.len) {
          float* dst = out + (size_t)(it.row0 + qrow) * D_MODEL + o_col;
#pragma unroll
          for (int i = 0; i < D_HEAD; i += 4)
            *reinterpret_cast<float4*>(dst + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
        }
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

inline void launch(const float* qkv, int rows, const int32_t* starts, const int32_t* lens, const uint32_t* work,
                   const int32_t* work_count, int work_cap, float* out, cudaStream_t stream, int sms) {
  if (work_cap <= 0 || rows <= 0) return;
  static bool configured[64] = {};
  int dev = 0;
  FS2_CUDA_OK(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    FS2_CUDA_OK(cudaFuncSetAttribute(attention_tcr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    configured[dev & 63] = true;
  }
  const CUtensorMap tmQK = make_map(qkv, rows, LDQKV, LDQKV, BKV, true, true);
  const CUtensorMap tmV = make_map(qkv, rows, LDQKV, LDQKV, BKV, true, true, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  const CUtensorMap tmO = make_map(out, rows, D_MODEL, D_MODEL, 32, false, false);
  launch_pdl(attention_tcr_kernel, dim3(std::min(sms, N_HEAD * work_cap)), dim3(THREADS), SMEM_TOTAL, stream, 1, tmQK, tmV, tmO, starts,
             lens, work, work_count, out);
  FS2_LAUNCHED();
}

}  // namespace attn_r
}  // namespace fs2
