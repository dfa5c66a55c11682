// Fused position-wise FFN of an FFT block (transformer/SubLayers.py:85-93) in ONE persistent tcgen05 kernel:
//     y = LayerNorm( w_2( ReLU( w_1 * x ) ) + x ),  w_1 = Conv1d(256 -> 1024, k = 9, pad 4),  w_2 = Conv1d(1024 -> 256, k = 1)
// The 1024-wide hidden row never leaves the SM.  Per 128-row tile and per 256-column chunk c of the hidden layer:
//     acc1[128 x 256]  = sum_{tap, k} x[r + tap - 4, k] . W1[tap][c*256 + n][k]        72 ring steps, operands by TMA
//     acc1            <- tf32( ReLU(acc1 + b1) )          in place in TENSOR MEMORY, by the epilogue warps
//     out [128 x 256] += acc1 . W2[:, c*256 : (c+1)*256]^T                               8 ring steps, A operand FROM TMEM
// and after the fourth chunk the post-LN epilogue of gemm_tc2.cuh runs on `out` (bias, residual by TMA, two-pass
// statistics, affine, row mask, TMA store).  TMEM: acc1 = columns [0,256), out = [256,512).
// Versus the two-kernel form this removes the 110 MB hidden tensor (written once, read once) and the second kernel's
// operand-delivery-bound main loop: the second contraction's A operand is already on chip.
// The tensor pipe executes MMAs in issue order, so conv(c+1) overwriting acc1 needs no barrier against GEMM2(c)'s reads;
// the two hand-offs with the epilogue warps (hidden ready / out drained) are mbarriers.
//
// Scheduling ("stream-K" over hidden chunks).  The unit of work is (row-tile group, hidden chunk): 4 units per group.  The
// launch has one cluster per SM pair and cluster k walks the contiguous unit range [k U / n, (k+1) U / n), so a machine
// that holds 74 clusters takes 105 groups (batch 64: 26.8 k decoder rows) as 5.7 units each instead of one or two whole
// groups each (2 rounds for 1.42 rounds of work: the reason the fused kernel lost to the two-launch form at batch 64).
// A group whose chunks straddle two clusters is finished by the cluster that owns its FIRST chunks: the other cluster
// (which meets the group's tail at the very start of its range) writes its partial `out` tile to a global workspace and
// raises a flag; the owner -- which reaches the group at the end of its range, one to five units later -- loads that
// partial into its `out` accumulator while the conv of its first chunk runs and accumulates on top of it.
// Deadlock freedom does not need the whole grid resident: cluster k only ever waits for the FIRST action of cluster k + 1 (its
// tail dump), which itself waits for nothing.  If fewer clusters fit than the grid has (SMs held by another stream's kernel),
// clusters are dispatched in blockIdx order, every resident cluster except the last one has its successor resident and
// finishes, and each SM it frees admits the next cluster, whose first action releases the one that was spinning.
#pragma once

#include <cstdlib>

#include "common.cuh"
#include "gemm_tc2.cuh"

namespace fs2 {
namespace ffn {

using namespace tc2;   // (which itself brings in the PTX helpers of namespace tc)

constexpr int BM = tc2::BM;              // local names win over the same-named constants of the older engines
constexpr int THREADS = 192;             // TMA producer, MMA issuer, four epilogue warps
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;\n" ::: "memory"); }
constexpr int HC = 256;                       // hidden columns per chunk = N of the first MMA = K of the second
constexpr int N_CHUNKS = D_INNER / HC;        // 4
constexpr int KC1 = D_MODEL / 32;             // 8 K chunks of the conv
constexpr int STEPS1 = FFN_TAPS * KC1;        // 72
constexpr int STEPS2 = HC / 32;               // 8
// CL = 2 (every launch over more than one row tile): the CTA pair issues ONE tcgen05.mma.cta_group::2 (M = 256) per K slice
// and each CTA keeps only ITS half of the weight tile in shared memory (gemm_tc2.cuh's TWO form): a third less shared-memory
// operand traffic per MMA than the 1-SM multicast form (config 2, same box: dec.ffn_fused 1.10 -> 1.08 ms).
// CL = 1 (a single row tile): 1-SM MMAs.
// Tried and dropped: 128-column hidden chunks with TWO hidden buffers in tensor memory and conv(j+1) issued before
// GEMM2(j), which hides the ReLU hand-over (2.9 of 29.4 us per unit) under the next conv.  The N = 128 conv MMAs run
// 3 % slower and the forward got slower, not faster (same box, A/B: dec.ffn_fused 1.13-1.17 ms against 1.08).
// Resident activation tiles.  The k = 9 conv reads the SAME x rows nine times (tap t = the tile shifted by t - 4 rows); streaming
// them with the weights made every ring step 32 KB per CTA: 2.14 GB of L2 reads per launch at config 2 (ncu: L2 throughput 45 %
// of peak).  Now each 32-channel K chunk of the tile is loaded ONCE with its halo (rows [m0 - 4, m0 + 132)) and tap t is an
// MMA whose A descriptor starts t rows further down the same shared-memory tile (the tensor core applies the 128-byte
// swizzle to absolute address bits, gemm_tc2.cuh umma_desc_rowshift); only weight tiles go through the ring: 1.26 GB per
// launch (L2 throughput 26 %), half as many TMA operations, and a six-deep weight ring.  Same box A/B at config 2:
// dec.ffn_fused 1.099 -> 1.061 ms -- the main loop was only partly delivery bound.
constexpr int B_BYTES = 256 * 128;
constexpr int A_ROWS = BM + FFN_TAPS - 1;            // 136 rows per resident K chunk
constexpr int A_TILE_BYTES = A_ROWS * 128;           // 17,408 bytes arrive per chunk
constexpr int A_BUF_BYTES = 18 * 1024;               // buffer stride (1024-byte aligned: the swizzle pattern starts there)
constexpr int MAX_ABUF = 3, MAX_STAGES = 6;
constexpr int RING_BYTES = 150 * 1024;               // pair: 3 x 18 KB activations + 6 x 16 KB weights; single: 2 x 18 + 3 x 32
constexpr int OFF_CST = RING_BYTES;
constexpr int OFF_RES = OFF_CST + 8 * WCHUNK;
constexpr int OFF_PAR = OFF_RES + 8 * WCHUNK;          // b1[1024] | b2[256] | gamma[256] | beta[256]
constexpr int OFF_BAR = OFF_PAR + 8192;
constexpr int SMEM_TOTAL = OFF_BAR + 512 + 1024;
static_assert(SMEM_TOTAL <= 232448, "shared memory budget");

struct Args {
  const float* x;        // [rows, 256]: conv input AND residual
  int rows;
  const float* w1;       // [9][1024][256] tf32
  const float* b1;       // [1024]
  const float* w2;       // [256][1024] tf32
  const float* b2;       // [256]
  const float* gamma;    // [256]
  const float* beta;     // [256]
  const int32_t* row_vpos;
  const int32_t* row_room;
  int extra;
  const int32_t* live_rows;
  // optional: y[row] = (y[row] + post_a[u]) + post_b[u], u = post_utt[row], after the row mask, on the rows live with
  // post_extra reserved rows (the speaker / emotion conditioning on the last encoder layer: ConvGemmArgs::post_a)
  const float* post_a;
  const float* post_b;
  const int32_t* post_utt;
  int post_extra;
  float* y;              // [rows, 256]
  // hand-over of a row-tile group split between two clusters (stream-K schedule): partial [rows, 256] fp32 and
  // flags [flag_count(rows)] int32, each holding the epoch of the launch that last completed the slot (never reset;
  // zero-initialised, epochs start at 1 and are unique per launch on the stream)
  float* partial;
  int32_t* flags;
  int32_t epoch;
  long long* trace;      // trace builds only (tools/trace_ffn.py): phase stamps of cluster 0's first units
};

__device__ __forceinline__ void umma_tf32_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

#ifdef FS2_TRACE_BUILD   // per-CTA stamps for tools/trace_ffn.py: [cta][entry, after pdl_wait, exit, first unit, units, smid]
__device__ long long g_ffn_cta_trace[256 * 6];
#endif

template <int CL>
__global__ void __launch_bounds__(THREADS, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmY,
                 const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmP, Args p) {
  extern __shared__ uint8_t smem_raw[];
  auto stamp = [&](int unit, int k) {
#ifdef FS2_TRACE_BUILD   // phase timestamps for tools/trace_ffn.py: [unit][issuer: start, conv issued, hidden ready, gemm2 issued |
                         // epilogue warp 0: conv complete, ReLU done, out complete, segment epilogue done]
    if (p.trace != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && unit < 8) {
      long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      p.trace[unit * 8 + k] = t;
    }
#endif
  };
  auto cta_stamp = [&](int k, long long v = -1) {
#ifdef FS2_TRACE_BUILD
    if (threadIdx.x == 0 && blockIdx.x < 256) {
      if (v < 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(v));
      g_ffn_cta_trace[blockIdx.x * 6 + k] = v;
    }
#endif
  };
  cta_stamp(0);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* cst = smem + OFF_CST;
  uint8_t* res = smem + OFF_RES;
  float* b1_s = reinterpret_cast<float*>(smem + OFF_PAR);
  float* b2_s = b1_s + D_INNER;
  float* gamma_s = b2_s + 256;
  float* beta_s = gamma_s + 256;
  constexpr bool TWO = CL == 2;
  constexpr int BB = TWO ? B_BYTES / 2 : B_BYTES;          // this CTA's part of a weight tile = one ring stage
  constexpr int STAGE_BYTES = BB, STAGES = TWO ? 6 : 3, NABUF = TWO ? 3 : 2;
  constexpr int OFF_RING = NABUF * A_BUF_BYTES;
  static_assert(OFF_RING + STAGES * STAGE_BYTES <= RING_BYTES && STAGES <= MAX_STAGES && NABUF <= MAX_ABUF, "ring");
  uint8_t* ring = smem + OFF_RING;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* hid_full = empty + MAX_STAGES;    // [1] conv MMAs of a chunk have completed
  uint64_t* hid_ready = hid_full + 1;     // [1] 128 epilogue threads wrote ReLU(hidden) back to TMEM
  uint64_t* out_full = hid_ready + 1;     // [1] the fourth GEMM2 of a tile has completed
  uint64_t* out_empty = out_full + 1;     // [1] 128 epilogue threads have read the out accumulator
  uint64_t* res_full = out_empty + 1;     // [4 warps][2]
  uint64_t* par_full = res_full + 8;      // [4 warps][2]  sub-tiles of another cluster's partial `out` tile
  uint64_t* a_full = par_full + 8;        // [3] a resident activation chunk has landed
  uint64_t* a_empty = a_full + MAX_ABUF;  // [3] every MMA that reads it has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + MAX_ABUF);

  const int warp = warp_index(), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmW1)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmW2)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmR)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmP)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(hid_full, 1);
    mbar_init(hid_ready, 128 * CL);    // 2-SM: the issuer (even CTA) waits for the epilogue threads of both CTAs
    mbar_init(out_full, 1);
    mbar_init(out_empty, 128 * CL);
    for (int u = 0; u < 8; ++u) mbar_init(&res_full[u], 1);
    for (int u = 0; u < 8; ++u) mbar_init(&par_full[u], 1);
    for (int u = 0; u < MAX_ABUF; ++u) {
      mbar_init(&a_full[u], 1);
      mbar_init(&a_empty[u], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    if (TWO) {   // the same warp of both CTAs, the same destination offset
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_hid = tmem_base, tmem_out = tmem_base + 256;
  const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
  if (CL > 1) cluster_sync_all();
  pdl_trigger();
  pdl_wait();
  int rows_live = p.rows;
  if (p.live_rows != nullptr) rows_live = min(rows_live, *p.live_rows);
  const int m_groups = ((rows_live + BM - 1) / BM + CL - 1) / CL;
  auto item_m0 = [&](int w) { return (w * CL + rank) * BM; };
  // this cluster's contiguous range of (group, hidden chunk) units; at least 4 per cluster, so that a group is shared by
  // at most two clusters
  const int n_units = m_groups * N_CHUNKS;
  const int n_cl = min((int)gridDim.x / CL, m_groups), kcl = (int)blockIdx.x / CL;
  const int u0 = kcl < n_cl ? (int)(((long long)kcl * n_units) / n_cl) : 0;
  const int u1 = kcl < n_cl ? (int)(((long long)(kcl + 1) * n_units) / n_cl) : 0;
  cta_stamp(1);
  cta_stamp(3, u0);
  cta_stamp(4, u1 - u0);
#ifdef FS2_TRACE_BUILD
  { uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); cta_stamp(5, smid); }
#endif

  if (warp == 0) {
    // ---------------- TMA producer
    const bool leader = elect_one();
    int it = 0;
    // resident activation chunks form one stream over (unit, K chunk); chunk n + 1 is requested before the weight steps of
    // chunk n are issued, i.e. one chunk (9 ring steps) ahead of the MMAs that read it
    const int n_a_total = (u1 - u0) * KC1;
    auto request_a = [&](int n) {
      if (n >= n_a_total) return;
      const int ab = n % NABUF;
      mbar_wait(&a_empty[ab], ((n / NABUF) & 1) ^ 1);
      if (leader) {
        const int u = u0 + n / KC1, kc = n % KC1;
        const int m0 = item_m0(u / N_CHUNKS);
        uint8_t* a_s = smem + ab * A_BUF_BYTES;
        if (!TWO) {
          mbar_expect_tx(&a_full[ab], A_TILE_BYTES);
          tma_load_2d(a_s, &tmX, kc * 32, m0 - FFN_TAPS / 2, &a_full[ab]);
        } else {   // own rows; both CTAs' bytes are counted on the issuer's barrier
          if (rank == 0) mbar_expect_tx(&a_full[ab], 2 * A_TILE_BYTES);
          tma_load_2d_2sm(a_s, &tmX, kc * 32, m0 - FFN_TAPS / 2, &a_full[ab]);
        }
      }
      __syncwarp();
    };
    request_a(0);
    for (int u = u0; u < u1; ++u) {
      const int c = u % N_CHUNKS;
      for (int kc = 0; kc < KC1; ++kc) {
        request_a((u - u0) * KC1 + kc + 1);
        for (int tap = 0; tap < FFN_TAPS; ++tap, ++it) {   // conv steps: the chunk's W1 tile of (tap, kc)
          const int s = it % STAGES;
          mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
          uint8_t* b_s = ring + s * STAGE_BYTES;
          if (leader) {
            if (!TWO) {
              mbar_expect_tx(&full[s], BB);
              tma_load_2d(b_s, &tmW1, kc * 32, tap * D_INNER + c * HC, &full[s]);
            } else {   // own half of the weight tile
              if (rank == 0) mbar_expect_tx(&full[s], 2 * BB);
              tma_load_2d_2sm(b_s, &tmW1, kc * 32, tap * D_INNER + c * HC + rank * 128, &full[s]);
            }
          }
          __syncwarp();
        }
      }
      for (int kc = 0; kc < STEPS2; ++kc, ++it) {   // second contraction: the W2 tile [256 outputs x 32 hidden columns]
        const int s = it % STAGES;
        mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
        uint8_t* b_s = ring + s * STAGE_BYTES;
        if (leader) {
          if (!TWO) {
            mbar_expect_tx(&full[s], BB);
            tma_load_2d(b_s, &tmW2, c * HC + kc * 32, 0, &full[s]);
          } else {
            if (rank == 0) mbar_expect_tx(&full[s], 2 * BB);
            tma_load_2d_2sm(b_s, &tmW2, c * HC + kc * 32, rank * 128, &full[s]);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (2-SM: the even CTA issues for the pair; each CTA's 128 rows of A come from its own shared
    // or tensor memory, its half of the weight tile from its own shared memory, its 128 accumulator rows are in its own TMEM)
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_tf32(TWO ? 2 * BM : BM, 256);
    int it = 0, n_h = 0, n_seg = 0, n_a = 0;
    for (int u = u0; u < (TWO && rank != 0 ? u0 : u1); ++u, ++n_h) {
      const int c = u % N_CHUNKS;
      // a segment = this cluster's consecutive chunks of one group; `out` is handed over per segment
      const bool seg_first = u == u0 || c == 0, seg_last = u == u1 - 1 || c == N_CHUNKS - 1;
      // head segment of a group whose last chunks belong to the next cluster: `out` starts from that cluster's partial
      const bool preloaded = c == 0 && u1 - u < N_CHUNKS;
      stamp(u - u0, 0);
      {
        // conv chunk c -> acc1 (the previous chunk's GEMM2 reads of acc1 precede these writes in the tensor pipe)
        for (int kc = 0; kc < KC1; ++kc, ++n_a) {
          const int ab = n_a % NABUF;
          if (TWO) mbar_wait_cluster(&a_full[ab], (n_a / NABUF) & 1); else mbar_wait(&a_full[ab], (n_a / NABUF) & 1);
          tc_fence_after();
          const uint64_t da_tile = umma_desc_rowshift(smem_u32(smem + ab * A_BUF_BYTES));
          for (int tap = 0; tap < FFN_TAPS; ++tap, ++it) {   // tap t = the resident chunk, t rows further down (+8 per 128-byte row)
            const int s = it % STAGES;
            if (TWO) mbar_wait_cluster(&full[s], (it / STAGES) & 1); else mbar_wait(&full[s], (it / STAGES) & 1);
            tc_fence_after();
            const uint64_t da = da_tile + (uint64_t)(tap * 8), db = umma_desc(ring + s * STAGE_BYTES);
            if (leader) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                if (TWO) umma_2sm<false>(tmem_hid, da + 2 * kk, db + 2 * kk, idesc, (kc | tap | kk) != 0 ? 1u : 0u);
                else umma_tf32(tmem_hid, da + 2 * kk, db + 2 * kk, idesc, (kc | tap | kk) != 0 ? 1u : 0u);
              }
              if (TWO) umma_commit_2sm(&empty[s]); else umma_commit(&empty[s]);
            }
            __syncwarp();
          }
          if (leader) { if (TWO) umma_commit_2sm(&a_empty[ab]); else umma_commit(&a_empty[ab]); }
          __syncwarp();
        }
        if (leader) { if (TWO) umma_commit_2sm(hid_full); else umma_commit(hid_full); }
        __syncwarp();
        stamp(u - u0, 1);
        // GEMM2 chunk c: out += ReLU(hidden chunk) (TMEM) x W2 tile (smem)
        if (TWO) mbar_wait_cluster(hid_ready, n_h & 1); else mbar_wait(hid_ready, n_h & 1);
        // the previous segment's epilogue has drained `out` (and, for a head segment, loaded the partial into it)
        if (seg_first) { if (TWO) mbar_wait_cluster(out_empty, (n_seg & 1) ^ 1); else mbar_wait(out_empty, (n_seg & 1) ^ 1); }
        tc_fence_after();
        stamp(u - u0, 2);
        for (int i = 0; i < STEPS2; ++i, ++it) {
          const int s = it % STAGES;
          if (TWO) mbar_wait_cluster(&full[s], (it / STAGES) & 1); else mbar_wait(&full[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint64_t db = umma_desc(ring + s * STAGE_BYTES);
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t accum = (!seg_first || preloaded || (i | kk) != 0) ? 1u : 0u;
              if (TWO) umma_tf32_ts_2sm(tmem_out, tmem_hid + i * 32 + kk * 8, db + 2 * kk, idesc, accum);
              else umma_tf32_ts(tmem_out, tmem_hid + i * 32 + kk * 8, db + 2 * kk, idesc, accum);
            }
            if (TWO) umma_commit_2sm(&empty[s]); else umma_commit(&empty[s]);
          }
          __syncwarp();
        }
      }
      stamp(u - u0, 3);
      if (seg_last) {
        if (leader) { if (TWO) umma_commit_2sm(out_full); else umma_commit(out_full); }
        __syncwarp();
        ++n_seg;
      }
    }
  } else {
    // ---------------- epilogue warps: thread = accumulator row
    const int et = threadIdx.x - 64;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    uint8_t* my_cst = cst + q * 2 * WCHUNK;
    uint8_t* my_res = res + q * 2 * WCHUNK;
    uint64_t* my_res_full = res_full + q * 2;
    const uint32_t b1_sa = smem_u32(b1_s), b2_sa = smem_u32(b2_s), gamma_sa = smem_u32(gamma_s), beta_sa = smem_u32(beta_s),
                   cst_sa = smem_u32(my_cst), res_sa = smem_u32(my_res);
    const uint32_t swz_x = (uint32_t)(lane & 7) << 4;
    for (int i = et; i < D_INNER; i += 128) b1_s[i] = p.b1[i];
    for (int i = et; i < 256; i += 128) {
      b2_s[i] = p.b2[i];
      gamma_s[i] = p.gamma[i];
      beta_s[i] = p.beta[i];
    }
    epi_barrier();
    int g_res = 0, g_st = 0, g_par = 0, n_h = 0, n_seg = 0;
    auto arrive_issuer = [&](uint64_t* bar) {   // the issuer lives in the even CTA
      if (TWO && rank != 0) mbar_arrive_remote(dsmem_addr(bar, 0));
      else mbar_arrive(bar);
    };
    uint64_t* my_par_full = par_full + q * 2;
    // groups this cluster FINISHES (LayerNorm + store): those whose first chunk lies in its range
    const int g_fin0 = (u0 + N_CHUNKS - 1) / N_CHUNKS;
    auto finishes = [&](int g) { return g * N_CHUNKS >= u0 && g * N_CHUNKS < u1; };
    if (lane == 0 && finishes(g_fin0)) {   // first residual sub-tile of the first group to finish
      mbar_expect_tx(&my_res_full[0], WCHUNK);
      tma_load_2d(my_res, &tmR, 0, item_m0(g_fin0) + q * 32, &my_res_full[0]);
    }
    int32_t* my_flag_base = p.flags + rank;      // slot of group g: flags[g * CL + rank] (each CTA hands over its own rows)
    // The next group's last chunks belong to the next cluster, which computed them first thing: wait for its flag, load
    // the partial tile into `out` (TMA -> staging -> tcgen05.st), then hand `out` to the issuer, whose second contraction
    // accumulates on top of it.  All of this runs under the conv of the group's first chunk.
    auto preload_partial = [&](int gn) {
      if (et == 0) {
        const long long t0 = clock64();
        int32_t seen;
        do {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(seen) : "l"(my_flag_base + (size_t)gn * CL) : "memory");
          if (clock64() - t0 > 4000000000LL) __trap();
        } while (seen != p.epoch);
      }
      epi_barrier();
      asm volatile("fence.proxy.async;\n" ::: "memory");   // the partial was written through the async proxy of another SM
      if (lane == 0) bulk_wait_read<0>();                   // both staging buffers are free
      __syncwarp();
      const int mn = item_m0(gn);
      auto load_par = [&](int cch) {
        if (lane != 0) return;
        const int buf = (g_par + (cch & 1)) & 1;
        mbar_expect_tx(&my_par_full[buf], WCHUNK);
        tma_load_2d(my_cst + buf * WCHUNK, &tmP, cch * 32, mn + q * 32, &my_par_full[buf]);
      };
      load_par(0);
      float v[32];
#pragma unroll 1
      for (int cch = 0; cch < 8; ++cch) {
        if (cch + 1 < 8) load_par(cch + 1);
        const int buf = (g_par + (cch & 1)) & 1;
        mbar_wait(&my_par_full[buf], ((g_par + cch) >> 1) & 1);
        const uint32_t pb = cst_sa + buf * WCHUNK + lane * 128;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          const float4 t4 = lds4(pb + ((cc << 4) ^ swz_x));
          v[cc * 4] = t4.x; v[cc * 4 + 1] = t4.y; v[cc * 4 + 2] = t4.z; v[cc * 4 + 3] = t4.w;
        }
        tmem_st32(tmem_out + lane_sel + cch * 32, v);
        fence_async_smem();         // every lane has read the buffer before the async proxy writes it again
        __syncwarp();
      }
      g_par += 8;
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      tc_fence_before();
      arrive_issuer(out_empty);
    };
    for (int u = u0; u < u1; ++u, ++n_h) {
      const int w = u / N_CHUNKS, c = u % N_CHUNKS;
      const bool seg_last = u == u1 - 1 || c == N_CHUNKS - 1;
      const int m0 = item_m0(w);
      // ---- hidden chunk: acc1 <- tf32(ReLU(acc1 + b1)) in place
      {
        mbar_wait(hid_full, n_h & 1);
        tc_fence_after();
        if (q == 0) stamp(u - u0, 4);
        float va[32], vb[32];
        tmem_ld32_issue(tmem_hid + lane_sel, va);
#pragma unroll 1
        for (int j = 0; j < 8; j += 2) {
          tmem_ld_wait();
          tmem_ld32_issue(tmem_hid + lane_sel + (j + 1) * 32, vb);
          auto relu_round = [&](float (&v)[32], int jj) {
            const uint32_t ba = b1_sa + (uint32_t)(c * HC + jj * 32) * 4;
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
              const float4 b4 = lds4(ba + cc * 16);
              const float t[4] = {v[cc * 4] + b4.x, v[cc * 4 + 1] + b4.y, v[cc * 4 + 2] + b4.z, v[cc * 4 + 3] + b4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e)   // ReLU, then round to TF32 (nearest, ties away) with integer arithmetic
                v[cc * 4 + e] = __uint_as_float((__float_as_uint(fmaxf(t[e], 0.f)) + 0x1000u) & 0xFFFFE000u);
            }
            tmem_st32(tmem_hid + lane_sel + jj * 32, v);
          };
          relu_round(va, j);
          tmem_ld_wait();
          if (j + 2 < 8) tmem_ld32_issue(tmem_hid + lane_sel + (j + 2) * 32, va);
          relu_round(vb, j + 1);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        tc_fence_before();
        arrive_issuer(hid_ready);
        if (q == 0) stamp(u - u0, 5);
      }
      if (!seg_last) continue;
      // ---- end of this cluster's segment of group w
      mbar_wait(out_full, n_seg & 1);
      ++n_seg;
      tc_fence_after();
      if (q == 0) stamp(u - u0, 6);
      const uint32_t acc = tmem_out + lane_sel;
      const int seg_c0 = (u - c >= u0) ? 0 : (u0 % N_CHUNKS);     // first chunk of the segment
      // the next segment (if any) starts a new group; it is a HEAD segment when the range ends inside that group: its
      // `out` accumulator must then hold the next cluster's partial before it is handed back to the issuer
      const bool next_is_head = u + 1 < u1 && u1 - (u + 1) < N_CHUNKS;
      if (seg_c0 != 0) {
        // ---- TAIL segment: this cluster met the group at the start of its range; the group is finished by the previous
        // cluster.  Raw `out` tile -> workspace (TMA stores), then the flag of this CTA's rows.
        float v[32];
#pragma unroll 1
        for (int cc8 = 0; cc8 < 8; ++cc8) {
          tmem_ld32(acc + cc8 * 32, v);
          if (cc8 == 7 && !next_is_head) {
            tc_fence_before();
            arrive_issuer(out_empty);
          }
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
          uint8_t* sb = my_cst + (g_st & 1) * WCHUNK;
          const uint32_t sa = cst_sa + (g_st & 1) * WCHUNK + lane * 128;
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) sts4(sa + ((cc << 4) ^ swz_x), make_float4(v[cc * 4], v[cc * 4 + 1], v[cc * 4 + 2], v[cc * 4 + 3]));
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmP, sb, cc8 * 32, m0 + q * 32);
            bulk_commit();
          }
          ++g_st;
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");   // the stores are COMPLETE, not just read
        __syncwarp();
        epi_barrier();
        if (et == 0) {
          __threadfence();
          asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(my_flag_base + (size_t)w * CL), "r"(p.epoch) : "memory");
        }
        if (next_is_head) preload_partial(w + 1);
        if (q == 0) stamp(u - u0, 7);
        continue;
      }
      // ---- HEAD (or whole) segment: LayerNorm(out + b2 + x) -> y   (same two-pass scheme as gemm_tc2.cuh's LN epilogue)
      const int row = m0 + r;
      bool live = row < p.rows;
      const float *post_a = nullptr, *post_b = nullptr;
      if (live && p.post_a != nullptr) {
        const int pu = p.post_utt[row];
        if (pu >= 0 && row_live(p.row_vpos[row], p.row_room[row], p.post_extra)) {
          post_a = p.post_a + (size_t)pu * 256;
          post_b = p.post_b + (size_t)pu * 256;
        }
      }
      if (live && p.row_vpos != nullptr) live = row_live(p.row_vpos[row], p.row_room[row], p.extra);
      const int w_next = w + 1;
      auto prefetch_res = [&](int cch) {
        if (lane != 0) return;
        int ww = w, cc = cch + 1;
        if (cc >= 8) { ww = w_next; cc = 0; }
        if (!finishes(ww)) return;
        const int buf = (g_res + 1) & 1;
        mbar_expect_tx(&my_res_full[buf], WCHUNK);
        tma_load_2d(my_res + buf * WCHUNK, &tmR, cc * 32, item_m0(ww) + q * 32, &my_res_full[buf]);
      };
      float mean = 0.f, m2 = 0.f;
      float va[32], vb[32];
      auto pass1 = [&](float (&v)[32], int cch) {
        __syncwarp();
        prefetch_res(cch);
        const uint32_t ba = b2_sa + (uint32_t)(cch * 32) * 4;
        mbar_wait(&my_res_full[g_res & 1], (g_res >> 1) & 1);
        const uint32_t rb = res_sa + (g_res & 1) * WCHUNK + lane * 128;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          const float4 b4 = lds4(ba + cc * 16);
          const float4 r4 = lds4(rb + ((cc << 4) ^ swz_x));
          v[cc * 4 + 0] += b4.x + r4.x; v[cc * 4 + 1] += b4.y + r4.y; v[cc * 4 + 2] += b4.z + r4.z; v[cc * 4 + 3] += b4.w + r4.w;
        }
        ++g_res;
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 32; j += 4) { s4[0] += v[j]; s4[1] += v[j + 1]; s4[2] += v[j + 2]; s4[3] += v[j + 3]; }
        const float cm = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.f / 32.f);
        float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float d0 = v[j] - cm, d1 = v[j + 1] - cm, d2 = v[j + 2] - cm, d3 = v[j + 3] - cm;
          q4[0] = fmaf(d0, d0, q4[0]); q4[1] = fmaf(d1, d1, q4[1]); q4[2] = fmaf(d2, d2, q4[2]); q4[3] = fmaf(d3, d3, q4[3]);
        }
        const float cm2 = (q4[0] + q4[1]) + (q4[2] + q4[3]);
        const float delta = cm - mean;
        const float inv_n = __frcp_rn((float)(cch + 1));
        mean = fmaf(delta, inv_n, mean);
        m2 += cm2 + delta * delta * (32.f * (float)cch * inv_n);
        tmem_st32(acc + cch * 32, v);
      };
      tmem_ld32_issue(acc, va);
#pragma unroll 1
      for (int cch = 0; cch < 8; cch += 2) {
        tmem_ld_wait();
        tmem_ld32_issue(acc + (cch + 1) * 32, vb);
        pass1(va, cch);
        tmem_ld_wait();
        if (cch + 2 < 8) tmem_ld32_issue(acc + (cch + 2) * 32, va);
        pass1(vb, cch + 1);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      const float rstd = 1.f / sqrtf(m2 * (1.f / 256.f) + 1e-5f);
      auto pass2 = [&](float (&v)[32], int cch) {
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          const float4 g4 = lds4(gamma_sa + cch * 128 + cc * 16), b4 = lds4(beta_sa + cch * 128 + cc * 16);
          const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) v[cc * 4 + e] = live ? fmaf((v[cc * 4 + e] - mean) * rstd, gg[e], bb[e]) : 0.f;
        }
        if (post_a != nullptr) {   // (x + speaker) + emotion, in the reference's order
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) {
            const float4 a4 = __ldg(reinterpret_cast<const float4*>(post_a + cch * 32 + cc * 4));
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(post_b + cch * 32 + cc * 4));
            v[cc * 4] = (v[cc * 4] + a4.x) + b4.x; v[cc * 4 + 1] = (v[cc * 4 + 1] + a4.y) + b4.y;
            v[cc * 4 + 2] = (v[cc * 4 + 2] + a4.z) + b4.z; v[cc * 4 + 3] = (v[cc * 4 + 3] + a4.w) + b4.w;
          }
        }
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        uint8_t* sb = my_cst + (g_st & 1) * WCHUNK;
        const uint32_t sa = cst_sa + (g_st & 1) * WCHUNK + lane * 128;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) sts4(sa + ((cc << 4) ^ swz_x), make_float4(v[cc * 4], v[cc * 4 + 1], v[cc * 4 + 2], v[cc * 4 + 3]));
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmY, sb, cch * 32, m0 + q * 32);
          bulk_commit();
        }
        ++g_st;
      };
      tmem_ld32_issue(acc, va);
#pragma unroll 1
      for (int cch = 0; cch < 8; cch += 2) {
        tmem_ld_wait();
        tmem_ld32_issue(acc + (cch + 1) * 32, vb);
        pass2(va, cch);
        tmem_ld_wait();
        if (cch + 2 < 8) {
          tmem_ld32_issue(acc + (cch + 2) * 32, va);
        } else if (!next_is_head) {        // the last read of `out` has landed: hand it back to the issuer
          tc_fence_before();
          arrive_issuer(out_empty);
        }
        pass2(vb, cch + 1);
      }
      if (next_is_head) preload_partial(w + 1);
      if (q == 0) stamp(u - u0, 7);
    }
    if (lane == 0) bulk_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  cta_stamp(2);
  if (CL > 1) cluster_sync_all();
  if (warp == 1) {
    if (TWO) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// 0 = conv9 + w2/LN as two launches, 1 = always the fused kernel, 2 = automatic (default; FS2_FFN_FUSED overrides).
// A (group, chunk) unit costs ~1.16 conv-tile times on one SM against 1 + ~0.28 for the two-launch form, but a cluster
// cannot take fewer than the 4 units of one group: the fused kernel wins once there are at least as many row-tile groups
// as clusters (the stream-K schedule then keeps every SM busy to within one unit), and loses SMs outright below that.
inline int& enabled_flag() {
  static int f = [] {
    const char* e = std::getenv("FS2_FFN_FUSED");
    return e != nullptr ? std::atoi(e) : 2;
  }();
  return f;
}
inline bool use_fused(int rows) {
  const int mode = enabled_flag();
  if (mode != 2) return mode == 1;
  return (rows + BM - 1) / BM >= sm_count();
}
// int32 hand-over flags a launch over `rows` rows may touch
inline size_t flag_count(int rows) { return (size_t)((rows + BM - 1) / BM + 2); }

template <int CL>
inline void launch_cl(const Args& a, cudaStream_t stream) {
  static bool configured[64] = {};
  int dev = 0;
  FS2_CUDA_OK(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    FS2_CUDA_OK(cudaFuncSetAttribute(ffn_fused_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    configured[dev & 63] = true;
  }
  const CUtensorMap tmX = make_map(a.x, a.rows, D_MODEL, D_MODEL, A_ROWS, /*round_tf32=*/true, false);
  const CUtensorMap tmW1 = make_map(a.w1, (int64_t)FFN_TAPS * D_INNER, D_MODEL, D_MODEL, 256 / CL, false, true);
  const CUtensorMap tmW2 = make_map(a.w2, D_MODEL, D_INNER, D_INNER, 256 / CL, false, true);
  const CUtensorMap tmY = make_map(a.y, a.rows, D_MODEL, D_MODEL, 32, false, false);
  const CUtensorMap tmR = make_map(a.x, a.rows, D_MODEL, D_MODEL, 32, false, false);
  const CUtensorMap tmP = make_map(a.partial, a.rows, D_MODEL, D_MODEL, 32, false, false);
  const int items = ((a.rows + BM - 1) / BM + CL - 1) / CL;
  const int grid = std::min(items, sm_count() / CL) * CL;
  launch_pdl(ffn_fused_kernel<CL>, dim3(grid), dim3(THREADS), SMEM_TOTAL, stream, CL, tmX, tmW1, tmW2, tmY, tmR, tmP, a);
  FS2_LAUNCHED();
}

inline void launch(const Args& a, cudaStream_t stream) {
  if (a.rows <= 0) return;
  require(a.x != a.y, FS2_ERR_INVALID, "fused FFN: output must not alias the input (conv halo rows are re-read)");
  require(a.partial != nullptr && a.flags != nullptr && a.epoch != 0, FS2_ERR_INVALID, "fused FFN: hand-over workspace missing");
  if (cluster_size_flag() == 2 && a.rows > BM) launch_cl<2>(a, stream);
  else launch_cl<1>(a, stream);
}

}  // namespace ffn
}  // namespace fs2
