// Persistent implicit-GEMM Conv1d / Linear on tcgen05 with fused epilogues (the product engine).
//
// One CTA per SM loops over output tiles (128 rows x BN columns).  Three pipelines run
// concurrently:
//   warp 0      TMA producer: A (activations, tap = row offset) and B (weights) into a smem ring;
//   warp 1      MMA issuer:   tcgen05.mma kind::tf32 into one of TWO TMEM accumulators, so the
//                             main loop of tile i+1 overlaps the epilogue of tile i;
//   warps 2..5  epilogue:     tcgen05.ld (thread = output row) -> bias / ReLU / tanh / residual /
//                             row mask, or the full post-LN of the FFT block and the predictors
//                             (LayerNorm over the 256-wide row held in TMEM, optional 256->1 head)
//                             -> swizzled smem staging -> TMA store (coalesced, asynchronous).
// With CL = 2 the kernel runs as thread-block clusters of two CTAs that work on vertically adjacent
// row tiles of the SAME column tile: each CTA loads half of the weight tile and TMA-multicasts it into
// both CTAs' shared memory, halving the L2 traffic of the B operand (two thirds of all operand bytes).
// Residual tiles are fetched by TMA into smem (never by per-thread strided loads).  Every epilogue
// warp owns its 32 rows end to end (own staging buffers, own TMA loads/stores, own mbarriers), so
// the steady state has no CTA-wide barrier.
#pragma once

#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"
#include "conv_args.cuh"
#include "tc_ptx.cuh"

namespace fs2 {
namespace tc2 {

using namespace tc;

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int WCHUNK = 32 * 128;        // one warp's [32 rows x 32 fp32] swizzled sub-tile (4 KB)
constexpr int MAX_N = 2048;

// A-resident mode (AR, small K with many taps -- the vocoder's 32/64-channel dilated convs): the activation tile is
// loaded ONCE per output tile with its halo rows ([m0 - pad, m0 - pad + AR_ROWS) for every K chunk) and each tap is an
// MMA whose A descriptor starts tap*dil rows further down the same shared-memory tile; only the weights stream through
// the ring.  Without it every tap re-fetches the whole 128-row tile from L2 (11x the traffic for k = 11).
// The weight ring then moves 16 KB stages that hold several (tap, K chunk) units each: with 4 KB weight tiles a
// one-unit stage is pure mbarrier latency (measured: 530 cycles per tap for 64 cycles of MMA).
constexpr int AR_ROWS = 192;      // 128 + (taps - 1) * dil <= 192  (k = 11, dilation 5: 178)
constexpr int AR_STAGE_BYTES = 16384;

// AR = 0: activations stream with the weights; AR = 1 / 2: resident activation tile of 1 / 2 K chunks (K <= 32 / 64)
// TWO: the CTA pair issues ONE tcgen05.mma.cta_group::2 (M = 256) per K slice; each CTA then keeps only ITS half of the
// weight tile (BN/2 rows) in shared memory, so a stage is 32 KB instead of 48 and four of them fit.
template <int BN, int AR = 0, int NS = 1, bool TWO = false>
struct Cfg {
  // narrow streaming tiles (single utterances) are TMA-latency bound: a fifth 24 KB stage is 25 % more bytes in flight
  static constexpr int STAGES = TWO ? 4 : (AR == 2 ? 3 : (BN > 128 ? 3 : ((BN <= 64 && AR == 0) ? 5 : 4)));
  static constexpr int A_BYTES = AR ? 0 : BM * 128;
  static constexpr int B_BYTES = (TWO ? BN / 2 : BN) * 128;
  static constexpr int STAGE_BYTES = AR ? AR_STAGE_BYTES : A_BYTES + B_BYTES;
  static constexpr int UNITS = AR ? AR_STAGE_BYTES / B_BYTES : 1;          // (tap, K chunk) weight tiles per ring stage
  static constexpr int AR_CHUNK_BYTES = AR_ROWS * 128;
  static constexpr int AR_BUF_BYTES = AR * AR_CHUNK_BYTES;                  // one tile's activations
  // Activation tiles in flight.  Once the issue loop is lean, the k = 3 convs alternate 0.77 / 1.28 us per tile: with two
  // buffers the tile two ahead is only requested when the current one retires, i.e. one HBM latency (~2 us) per two tiles.
  // Three 24 KB buffers (requested two tiles ahead) fit; three 48 KB buffers do not.
  static constexpr int AR_BUFS = AR == 1 ? 3 : 2;
  static constexpr int OFF_RING = AR_BUFS * AR_BUF_BYTES;
  static constexpr int OFF_CST = OFF_RING + STAGES * STAGE_BYTES;      // 4 warps x 2 (or 8 warps x 1) output staging sub-tiles
  static constexpr int OFF_RES = OFF_CST + 8 * WCHUNK;      // 4 warps x 2 (or 8 warps x 1) residual sub-tiles
  static constexpr int OFF_PAR = OFF_RES + 8 * WCHUNK;      // bias[2048] | gamma[256] | beta[256] | head_w[256]
  static constexpr int OFF_BAR = OFF_PAR + 12288;
  static constexpr int OFF_STAT = OFF_BAR + 256;            // N-split: float2 [2 parities][4 source ranks][128 rows]
  static constexpr int OFF_PAIR = OFF_STAT + (NS > 1 ? 8192 : 0);   // two LayerNorm warps per lane quarter: float2 [2][2][128]
  static constexpr int TOTAL = OFF_PAIR + (NS > 1 ? 0 : 4096) + 1024;
  static constexpr int ACC_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
  static constexpr int NCHUNK = (BN + 31) / 32;
  static_assert(B_BYTES % 1024 == 0, "stages must stay 1024-byte aligned (SWIZZLE_128B)");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128");
  static_assert(TOTAL <= 232448, "shared memory budget");
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// explicit shared-space 16-byte accesses (the generic pointers derived from the aligned dynamic
// smem base would otherwise compile to generic LD/ST)
__device__ __forceinline__ float4 lds4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts4(uint32_t saddr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// byte offset of (row r, 16-byte chunk cc) inside a [rows x 128 B] SWIZZLE_128B sub-tile
__device__ __forceinline__ int swz_off(int r, int cc) { return r * 128 + ((cc ^ (r & 7)) << 4); }

// 16-byte shared store of eight values rounded to bf16 (round to nearest even)
__device__ __forceinline__ void sts8_bf16(uint32_t saddr, const float* v) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(saddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
}

// BF = bf16 operands (kind::f16, 64 elements per 128-byte K chunk) instead of TF32 (32 elements).
// K-major SWIZZLE_128B descriptor that starts `rows` rows inside a larger tile whose swizzle pattern begins on a
// 1024-byte boundary (where TMA put it).  Measured on B200 (tools/ar_probe.py): the tensor core applies the XOR to the
// absolute shared-memory address bits [7,10), so a row-shifted start address needs nothing else; the base-offset field
// (bits [49,52)) describes a PATTERN that starts off-boundary and must stay 0 here (setting it to the row phase
// permutes the 16-byte chunks of every row whose phase carries into bit 2).
__device__ __forceinline__ uint64_t umma_desc_rowshift(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// ---- cta_group::2 forms (a CTA pair acting as one 256-row MMA unit; syntax as in CUTLASS cute/arch/*_sm100*.hpp)
// TMA load executed by BOTH CTAs of the pair into their own shared memory; the transaction bytes are counted on the
// mbarrier of the pair's even-ranked CTA (the MMA issuer): clearing bit 24 of the shared::cluster address selects it.
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar) & 0xFEFFFFFFu)
      : "memory");
}
template <bool BF16>
__device__ __forceinline__ void umma_2sm(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (BF16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {   // arrives on this barrier in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)0x3)
               : "memory");
}


// NS > 1 (fused-LayerNorm only): N-split.  A cluster of NS CTAs shares ONE row tile, CTA r owns the 256/NS columns
// [r*BN, (r+1)*BN): each loads 1/NS of the activation tile and multicasts it to the others, streams only its own
// weight columns, and the LayerNorm statistics of the row are combined across the cluster through distributed shared
// memory (two floats per row and CTA).  For a handful of row tiles (single utterances) this puts NS SMs on a K loop
// that one CTA would walk alone: the fused-LN GEMMs were 40 % of the single-utterance latency.
// KS > 1: K-split.  A cluster of KS CTAs shares ONE output tile and each walks 1/KS of the (tap, K chunk) loop; the partial
// accumulators of ranks 1..KS-1 travel through a small global (L2-resident) workspace to rank 0, which adds them to its own
// and runs the epilogue.  For a SINGLE row tile (one utterance) with a long K loop the kernel is otherwise one CTA per
// column tile pulling megabytes of operands through one SM at ~90 GB/s (measured: 17 us for the k = 9 FFN conv of a
// 56-frame utterance, 72 ring steps); KS = 8 puts 8 SMs on that loop.
template <int BN, bool LN, int CL, bool BF, int AR, int NS, bool TWO, int EW, int KS>
__device__ __forceinline__ void conv_gemm_tc2_body(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmC,
                                                   const CUtensorMap& tmR, const CUtensorMap& tmC2, const ConvGemmArgs& p) {
  using C = Cfg<BN, AR, NS, TWO>;
  static_assert(!TWO || (CL == 2 && AR == 0 && NS == 1 && BN == 256), "2-SM MMA: CTA pairs, 256 columns");
  static_assert(!(AR && LN), "A-resident mode: plain epilogue");
  static_assert(!(AR == 2 && BF), "bf16 A-resident tiles hold 64 channels in one K chunk: AR = 1");
  static_assert(NS == 1 || (LN && CL == 1 && !AR && BN * NS == 256), "N-split: fused LayerNorm over 256 columns, cluster along N");
  static_assert(EW == 4 || (EW == 8 && NS == 1), "epilogue warps: one or two per TMEM lane quarter");
  static_assert(!LN || EW == 4 || BN == 256, "two LayerNorm warps per quarter split 8 sub-tiles");
  static_assert(KS == 1 || (CL == 1 && NS == 1 && AR == 0 && !TWO && !LN), "K-split: plain epilogue, cluster along K only");
  constexpr int CSIZE = CL > 1 ? CL : (NS > 1 ? NS : KS);          // CTAs per cluster
  constexpr uint16_t CMASK = (uint16_t)((1u << CSIZE) - 1);
  constexpr int BKE = BF ? 64 : 32;   // operand elements per 128-byte swizzle row
  extern __shared__ uint8_t smem_raw[];
  auto stamp = [&](int k) {
#ifdef FS2_TRACE_BUILD   // phase timestamps for tools/trace_gemm.py
    if (p.trace != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
      long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      p.trace[k] = t;
    }
#endif
  };
  if (threadIdx.x == 0) stamp(0);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* cst = smem + C::OFF_CST;
  uint8_t* res = smem + C::OFF_RES;
  float* bias_s = reinterpret_cast<float*>(smem + C::OFF_PAR);
  float* gamma_s = bias_s + MAX_N;
  float* beta_s = gamma_s + 256;
  float* headw_s = beta_s + 256;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* empty = full + C::STAGES;
  uint64_t* acc_full = empty + C::STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint64_t* res_full = acc_empty + 2;       // [4 warps][2]
  uint64_t* a_full = res_full + 8;          // [3]  AR: the resident activation tile of buffer lt % AR_BUFS has landed
  uint64_t* a_empty = a_full + 3;           // [3]  AR: every MMA that reads that buffer has completed
  uint64_t* stat_full = a_empty + 3;        // [2]  NS: the peers' LayerNorm statistics of this tile parity have arrived
  uint64_t* sk_full = stat_full + 2;        // [1]  KS: the partial accumulators of every other rank are in the workspace
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sk_full + 1);
  float2* stat_s = reinterpret_cast<float2*>(smem + C::OFF_STAT);        // NS: [2 parities][4 source ranks][128 rows]
  float2* pair_stat = reinterpret_cast<float2*>(smem + C::OFF_PAIR);     // EW = 8: [2 parities][2 halves][128 rows]
  uint8_t* ring = smem + C::OFF_RING;

  const int warp = warp_index(), lane = threadIdx.x & 31;
  const int kchunks = (p.K + BKE - 1) / BKE;
  const bool half_chunk = BF && p.K <= 32;   // A-resident bf16 convs with 32 channels: only two of the four K slices hold data
  // FS2_MATH_TF32X3: the K loop runs three terms over split operands (A = [hi | lo] columns, W = hi block then lo block)
  const int terms = p.terms > 1 ? p.terms : 1;
  const int iters1 = p.taps * kchunks;
  const int iters = terms * iters1;
  const int n_tiles_n = (p.N + BN - 1) / BN;
  const bool has_res = p.residual != nullptr;
  const bool has_out = p.C != nullptr;
  const bool has_out2 = p.C2 != nullptr;
  const int dil = p.dil > 0 ? p.dil : 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
    if (has_out) asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
    if (has_out2) asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmC2)) : "memory");
    if (has_res) asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmR)) : "memory");
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], (TWO || KS > 1) ? 1 : CSIZE);   // a stage is refilled by every CTA of the cluster: all consumers must release
                                               // it (2-SM: the one issuer's commit covers both CTAs' operands)
    }
    for (int u = 0; u < 2; ++u) {
      mbar_init(&acc_full[u], 1);
      mbar_init(&acc_empty[u], (TWO ? 2 : 1) * EW * 32);   // 2-SM: the issuer waits for the epilogue threads of both CTAs
    }
    for (int u = 0; u < 8; ++u) mbar_init(&res_full[u], 1);
    for (int u = 0; u < 3; ++u) {
      mbar_init(&a_full[u], 1);
      mbar_init(&a_empty[u], 1);
    }
    for (int u = 0; u < 2; ++u) mbar_init(&stat_full[u], 128 * (NS > 1 ? NS - 1 : 1));
    mbar_init(sk_full, (KS > 1 ? KS - 1 : 1) * EW);   // one arrival per epilogue WARP of every other rank
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    if (TWO) {   // the same warp of both CTAs, the same destination offset
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)C::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)C::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) stamp(1);
  const int rank = CSIZE > 1 ? (int)cluster_ctarank() : 0;
  if (CSIZE > 1) cluster_sync_all();   // the peers' barriers are initialised before anything is multicast to them
  // ---- everything above overlaps the previous kernel's tail; from here on its results are read
  pdl_trigger();
  pdl_wait();
  int rows_live = p.rows;
  if (p.live_rows != nullptr) rows_live = min(rows_live, *p.live_rows << p.mask_shift);
  // work item = CL vertically adjacent row tiles x one column tile; the CTAs of a cluster walk the items in
  // lock step (a trailing odd row tile is processed as an all-zero tile whose stores TMA clips away)
  const int m_groups = ((rows_live + BM - 1) / BM + CL - 1) / CL;
  const int total_items = NS > 1 ? m_groups : m_groups * n_tiles_n;       // N-split: one item = one row tile
  const int w_first = blockIdx.x / CSIZE, w_step = gridDim.x / CSIZE;
  auto item_m0 = [&](int w) { return NS > 1 ? w * BM : ((w / n_tiles_n) * CL + (KS > 1 ? 0 : rank)) * BM; };
  // K-split: this rank's share of the iteration space [it_lo, it_hi)
  const int it_lo = KS > 1 ? (int)(((long long)rank * iters) / KS) : 0;
  const int it_hi = KS > 1 ? (int)(((long long)(rank + 1) * iters) / KS) : iters;
  auto item_n0 = [&](int w) { return NS > 1 ? rank * BN : (w % n_tiles_n) * BN; };

  if (warp == 0) {
    // ---------------- TMA producer (whole warp runs the loop, one elected lane issues)
    const bool leader = elect_one();
    int it = 0, lt = 0;
    // AR with a weight tensor that fits the ring region (all 11 taps of a 32-channel conv are 44 KB): loaded ONCE per CTA
    const bool wres = AR && iters * C::B_BYTES <= C::STAGES * C::STAGE_BYTES;
    if (wres && w_first < total_items) {
      int wrow = item_n0(w_first) + (CL > 1 ? rank * (BN / 2) : 0), kcol = 0;
      uint8_t* b_s = ring + (CL > 1 ? rank * (C::B_BYTES / 2) : 0);
      if (leader) mbar_expect_tx(&full[0], (uint32_t)iters * C::B_BYTES);
      for (int i = 0; i < iters; ++i) {
        if (leader) {
          if (CL == 1) tma_load_2d(b_s, &tmW, kcol, wrow, &full[0]);
          else tma_load_2d_mc(b_s, &tmW, kcol, wrow, &full[0], (uint16_t)0x3);
        }
        b_s += C::B_BYTES;
        kcol += BKE;
        if (kcol >= kchunks * BKE) { kcol = 0; wrow += p.N; }
      }
      __syncwarp();
    }
    // AR: the la-th activation tile of this CTA (with its halo), requested AR_BUFS - 1 tiles ahead of the one being issued
    auto request_a = [&](int la) {
      const int wa = w_first + la * w_step;
      if (wa >= total_items) return;
      const int ua = la % C::AR_BUFS;
      mbar_wait(&a_empty[ua], ((la / C::AR_BUFS) & 1) ^ 1);
      if (leader) {
        mbar_expect_tx(&a_full[ua], (uint32_t)kchunks * C::AR_CHUNK_BYTES);
        for (int kc = 0; kc < kchunks; ++kc)
          tma_load_2d(smem + ua * C::AR_BUF_BYTES + kc * C::AR_CHUNK_BYTES, &tmA, kc * BKE, item_m0(wa) - p.pad, &a_full[ua]);
      }
      __syncwarp();
    };
    if (AR) {
      for (int la = 0; la < C::AR_BUFS - 1; ++la) request_a(la);
    }
    for (int w = w_first; w < total_items; w += w_step, ++lt) {
      const int m0 = item_m0(w), n0 = item_n0(w);
      if (AR) request_a(lt + C::AR_BUFS - 1);
      if (AR && wres) continue;
      if (AR) {   // weights: several (tap, K chunk) units per 16 KB stage (coordinates advance without divisions)
        int wrow = n0 + (CL > 1 ? rank * (BN / 2) : 0), kcol = 0, sidx = lt == 0 ? 0 : it % C::STAGES;
        for (int i0 = 0; i0 < iters; i0 += C::UNITS, ++it) {
          const int s = sidx;
          sidx = sidx + 1 == C::STAGES ? 0 : sidx + 1;
          mbar_wait(&empty[s], ((it / C::STAGES) & 1) ^ 1);
          const int n_units = min(C::UNITS, iters - i0);
          if (leader) mbar_expect_tx(&full[s], (uint32_t)n_units * C::B_BYTES);
          uint8_t* b_s = ring + s * C::STAGE_BYTES + (CL > 1 ? rank * (C::B_BYTES / 2) : 0);
          for (int j = 0; j < n_units; ++j) {
            if (leader) {
              if (CL == 1) tma_load_2d(b_s, &tmW, kcol, wrow, &full[s]);
              else tma_load_2d_mc(b_s, &tmW, kcol, wrow, &full[s], (uint16_t)0x3);
            }
            b_s += C::B_BYTES;
            kcol += BKE;
            if (kcol >= kchunks * BKE) { kcol = 0; wrow += p.N; }
          }
          __syncwarp();
        }
        continue;
      }
      for (int i = it_lo; i < it_hi; ++i, ++it) {
        const int s = it % C::STAGES;
        mbar_wait(&empty[s], ((it / C::STAGES) & 1) ^ 1);
        int term = 0, i1 = i;
        if (terms > 1) { term = i / iters1; i1 = i - term * iters1; }
        const int tap = i1 / kchunks, kc = i1 - tap * kchunks;
        const int a_col = kc * BKE + (term == 1 ? p.K : 0);                 // term 1 reads the lo half of the activations
        const int w_row = tap * p.N + n0 + (term == 2 ? p.taps * p.N : 0);  // term 2 reads the lo block of the weights
        uint8_t* a_s = ring + s * C::STAGE_BYTES;
        if (TWO) {   // own activation rows + own half of the weight tile; both CTAs' bytes are counted on the issuer's barrier
          if (leader) {
            if (rank == 0) mbar_expect_tx(&full[s], 2 * C::STAGE_BYTES);
            tma_load_2d_2sm(a_s, &tmA, a_col, m0 + tap * dil - p.pad, &full[s]);
            tma_load_2d_2sm(a_s + C::A_BYTES, &tmW, kc * BKE, w_row + rank * (BN / 2), &full[s]);
          }
          __syncwarp();
          continue;
        }
        if (leader) {
          mbar_expect_tx(&full[s], C::STAGE_BYTES);
          if (NS > 1) {   // this CTA's 128/NS rows of the shared activation tile, delivered to every CTA of the cluster
            tma_load_2d_mc(a_s + rank * (C::A_BYTES / NS), &tmA, a_col, m0 + tap * dil - p.pad + rank * (BM / NS), &full[s], CMASK);
          } else {
            tma_load_2d(a_s, &tmA, a_col, m0 + tap * dil - p.pad, &full[s]);
          }
          if (CL == 1) {
            tma_load_2d(a_s + C::A_BYTES, &tmW, kc * BKE, w_row, &full[s]);
          } else {   // this CTA's half of the weight tile, delivered to both CTAs of the cluster
            tma_load_2d_mc(a_s + C::A_BYTES + rank * (C::B_BYTES / 2), &tmW, kc * BKE, w_row + rank * (BN / 2), &full[s],
                           (uint16_t)0x3);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (whole warp runs the loop, one elected lane issues)
    const bool leader = elect_one();
    constexpr uint32_t idesc = BF ? umma_idesc_bf16(TWO ? 2 * BM : BM, BN) : umma_idesc_tf32(TWO ? 2 * BM : BM, BN);
    int it = 0, lt = 0;
    for (int w = w_first; TWO && rank == 0 && w < total_items; w += w_step, ++lt) {
      // 2-SM: the even CTA issues for the pair.  One MMA covers the pair's 256 rows (each CTA's 128 rows of A from its own
      // shared memory, its half of the weight tile from its own shared memory, its 128 accumulator rows in its own TMEM).
      const int u = lt & 1;
      mbar_wait_cluster(&acc_empty[u], ((lt >> 1) & 1) ^ 1);   // the epilogue threads of BOTH CTAs drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + u * C::ACC_COLS;
      for (int i = 0; i < iters; ++i, ++it) {
        const int s = it % C::STAGES;
        mbar_wait_cluster(&full[s], (it / C::STAGES) & 1);     // both CTAs' operands of this stage have landed
        tc_fence_after();
        const uint8_t* a_s = ring + s * C::STAGE_BYTES;
        const uint64_t da = umma_desc(a_s), db = umma_desc(a_s + C::A_BYTES);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_2sm<BF>(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (i | kk) != 0 ? 1u : 0u);
          umma_commit_2sm(&empty[s]);
        }
        __syncwarp();
      }
      if (leader) umma_commit_2sm(&acc_full[u]);
      __syncwarp();
    }
    for (int w = w_first; !TWO && w < total_items; w += w_step, ++lt) {
      const int u = lt & 1;
      if (lt < 6) stamp(8 + lt * 4 + 0);
      mbar_wait(&acc_empty[u], ((lt >> 1) & 1) ^ 1);   // epilogue drained this accumulator
      if (lt < 6) stamp(8 + lt * 4 + 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + u * C::ACC_COLS;
      if (AR && iters * C::B_BYTES <= C::STAGES * C::STAGE_BYTES) {   // resident weights (see the producer): no stages at all
        if (lt == 0) mbar_wait(&full[0], 0);
        const int ua = lt % C::AR_BUFS;
        mbar_wait(&a_full[ua], (lt / C::AR_BUFS) & 1);
        tc_fence_after();
        const uint64_t da_tile = umma_desc_rowshift(smem_u32(smem + ua * C::AR_BUF_BYTES));
        const uint64_t da_tap = (uint64_t)(dil * 8);
        uint64_t db = umma_desc(ring);
        int tap = 0, kc = 0;
        uint32_t started = 0;
        for (int i = 0; i < iters; ++i) {
          const uint64_t da = da_tile + (uint64_t)tap * da_tap + (uint64_t)(kc * (C::AR_CHUNK_BYTES >> 4));
          if (leader) {
            if (BF) {
              umma_bf16(d_tmem, da, db, idesc, started);
              umma_bf16(d_tmem, da + 2, db + 2, idesc, 1u);
              if (!half_chunk) {   // 32 bf16 channels fill half of the 64-element chunk: the rest is TMA zero fill
                umma_bf16(d_tmem, da + 4, db + 4, idesc, 1u);
                umma_bf16(d_tmem, da + 6, db + 6, idesc, 1u);
              }
            } else {
              umma_tf32(d_tmem, da, db, idesc, started);
              umma_tf32(d_tmem, da + 2, db + 2, idesc, 1u);
              umma_tf32(d_tmem, da + 4, db + 4, idesc, 1u);
              umma_tf32(d_tmem, da + 6, db + 6, idesc, 1u);
            }
          }
          started = 1u;
          db += (uint64_t)(C::B_BYTES >> 4);
          if (++kc == kchunks) { kc = 0; ++tap; }
        }
        if (leader) {
          umma_commit(&acc_full[u]);
          umma_commit(&a_empty[ua]);
        }
        __syncwarp();
        continue;
      }
      if (AR) {
        const int ua = lt % C::AR_BUFS;
        mbar_wait(&a_full[ua], (lt / C::AR_BUFS) & 1);
        // The MMAs of these shapes are 16-32 cycles of tensor work each, so the ISSUE loop is the critical path (measured:
        // +0.14 us per tap with one integer division and two descriptor builds per unit).  Descriptors advance by additions
        // only: +8 per activation row (128 B >> 4), a constant per K chunk and per weight unit.
        const uint64_t da_tile = umma_desc_rowshift(smem_u32(smem + ua * C::AR_BUF_BYTES));
        const uint64_t da_tap = (uint64_t)(dil * 8);
        int tap = 0, kc = 0;
        uint32_t started = 0;
        for (int i0 = 0; i0 < iters; i0 += C::UNITS, ++it) {
          const int s = it % C::STAGES;
          mbar_wait(&full[s], (it / C::STAGES) & 1);
          tc_fence_after();
          const int n_units = min(C::UNITS, iters - i0);
          uint64_t db = umma_desc(ring + s * C::STAGE_BYTES);
          for (int j = 0; j < n_units; ++j) {   // tap t of K chunk kc = the resident tile, t*dil rows further down
            const uint64_t da = da_tile + (uint64_t)tap * da_tap + (uint64_t)(kc * (C::AR_CHUNK_BYTES >> 4));
            if (leader) {
              if (BF) {
                umma_bf16(d_tmem, da, db, idesc, started);
                umma_bf16(d_tmem, da + 2, db + 2, idesc, 1u);
                if (!half_chunk) {
                  umma_bf16(d_tmem, da + 4, db + 4, idesc, 1u);
                  umma_bf16(d_tmem, da + 6, db + 6, idesc, 1u);
                }
              } else {
                umma_tf32(d_tmem, da, db, idesc, started);
                umma_tf32(d_tmem, da + 2, db + 2, idesc, 1u);
                umma_tf32(d_tmem, da + 4, db + 4, idesc, 1u);
                umma_tf32(d_tmem, da + 6, db + 6, idesc, 1u);
              }
            }
            started = 1u;
            db += (uint64_t)(C::B_BYTES >> 4);
            if (++kc == kchunks) { kc = 0; ++tap; }
          }
          if (leader) {
            if (CL == 1) umma_commit(&empty[s]); else umma_commit_mc(&empty[s], (uint16_t)0x3);
          }
          __syncwarp();
        }
        if (leader) {
          umma_commit(&acc_full[u]);
          umma_commit(&a_empty[ua]);   // the resident tile may be overwritten once these MMAs have read it
        }
        __syncwarp();
        continue;
      }
      for (int i = it_lo; i < it_hi; ++i, ++it) {
        const int s = it % C::STAGES;
        mbar_wait(&full[s], (it / C::STAGES) & 1);
        if (it == 0) stamp(2);
        tc_fence_after();
        const uint8_t* a_s = ring + s * C::STAGE_BYTES;
        const uint64_t da = umma_desc(a_s);
        const uint64_t db = umma_desc(a_s + C::A_BYTES);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {   // four 32-byte K slices per stage: K = 8 (tf32) or 16 (bf16) each
            if (BF) umma_bf16(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (i > it_lo || kk != 0) ? 1u : 0u);
            else umma_tf32(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (i > it_lo || kk != 0) ? 1u : 0u);
          }
          if (CSIZE == 1 || KS > 1) umma_commit(&empty[s]); else umma_commit_mc(&empty[s], CMASK);
        }
        __syncwarp();
      }
      if (leader) umma_commit(&acc_full[u]);
      __syncwarp();
      if (lt == 0) stamp(3);
    }
  } else {
    // ---------------- epilogue: EW warps, thread = one accumulator row, every warp independent of the others.
    // EW = 8: two warps share each TMEM lane quarter (a warp may only touch lanes 32*(warp%4)..+31) and split the tile's
    // 32-column sub-tiles between them, so every scheduler has TWO epilogue warps whose TMEM loads, shared-memory round
    // trips, fences and TMA issues overlap.  Measured with one warp per scheduler (tools/trace_ln.py): ~0.45 us per
    // sub-tile at ~0.2 instructions per cycle -- pure latency -- i.e. 7.7 us per 128 x 256 LayerNorm tile against a
    // 4.5 us main loop.  Each warp then keeps ONE staging and ONE residual buffer (same shared memory as 4 x 2).
    constexpr int EH = EW / 4;                       // warps per lane quarter
    constexpr int NB = EH == 2 ? 1 : 2;              // staging / residual buffers per warp
    constexpr int ET = EW * 32;                      // epilogue threads
    const int et = threadIdx.x - 64;
    const int ew = warp - 2;                         // 0 .. EW-1
    const int q = warp & 3;                          // TMEM lane quarter of this warp
    const int h = EH == 2 ? (ew >> 2) : 0;           // which share of the sub-tiles
    const int r = q * 32 + lane;                     // row inside the tile
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    constexpr int C_SPLIT = EH == 2 ? (C::NCHUNK + 1) / 2 : C::NCHUNK;
    const int c_lo = h * C_SPLIT, c_hi = h == 0 ? C_SPLIT : C::NCHUNK;   // this warp's sub-tiles [c_lo, c_hi)
    uint8_t* my_cst = cst + ew * NB * WCHUNK;
    uint8_t* my_res = res + ew * NB * WCHUNK;
    uint64_t* my_res_full = res_full + ew * NB;
    const uint32_t bias_sa = smem_u32(bias_s), gamma_sa = smem_u32(gamma_s), beta_sa = smem_u32(beta_s),
                   headw_sa = smem_u32(headw_s), cst_sa = smem_u32(my_cst), res_sa = smem_u32(my_res);
    const uint32_t swz_x = (uint32_t)(lane & 7) << 4;
    const int act = p.act;
    for (int i = et; i < p.N; i += ET) bias_s[i] = p.bias[i];
    if (LN) {
      for (int i = et; i < 256; i += ET) {
        gamma_s[i] = p.ln_gamma[i];
        beta_s[i] = p.ln_beta[i];
        headw_s[i] = p.head_w != nullptr ? p.head_w[i] : 0.f;
      }
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"(ET) : "memory");   // the only CTA-level epilogue barrier: parameters are in smem
#ifdef FS2_TRACE_BUILD
    long long res_wait_cycles = 0;   // cycles this thread spent waiting for residual sub-tiles (tools/trace_ln.py)
#endif
    int g_res = 0;   // residual sub-tiles consumed so far (buffer = g % NB, parity = (g / NB) & 1)
    int g_st = 0;    // staging sub-tiles produced so far
    const uint32_t res_bytes = p.res_bf16 ? WCHUNK / 2 : WCHUNK;   // [32 rows x 32 columns] fp32 or bf16
    if (has_res && lane == 0 && w_first < total_items && c_lo < c_hi && (KS == 1 || rank == 0)) {  // first residual sub-tile of the first tile
      mbar_expect_tx(&my_res_full[0], res_bytes);
      tma_load_2d(my_res, &tmR, item_n0(w_first) + c_lo * 32, item_m0(w_first) + q * 32, &my_res_full[0]);
    }
    int lt = 0;
    for (int w = w_first; w < total_items; w += w_step, ++lt) {
      const int m0 = item_m0(w), n0 = item_n0(w);
      const int u = lt & 1;
      const int row = m0 + r;
      const bool in_range = row < p.rows;
      bool live = in_range;
      if (in_range && p.row_vpos != nullptr) {
        const int mr = row >> p.mask_shift;
        live = row_live(p.row_vpos[mr], p.row_room[mr], p.extra);
      }
      const float *post_a = nullptr, *post_b = nullptr;   // this row's conditioning vectors (fused LayerNorm only)
      if (LN && p.post_a != nullptr && in_range) {
        const int pu = p.post_utt[row];
        if (pu >= 0 && row_live(p.row_vpos[row], p.row_room[row], p.post_extra)) {
          post_a = p.post_a + (size_t)pu * 256;
          post_b = p.post_b + (size_t)pu * 256;
        }
      }
      mbar_wait(&acc_full[u], (lt >> 1) & 1);
      if (lt == 0 && warp == 2) stamp(4);
      if (lt < 6 && warp == 2) stamp(8 + lt * 4 + 2);
      tc_fence_after();
      const uint32_t acc = tmem_base + lane_sel + u * C::ACC_COLS;
      const int w_next = w + w_step;
      auto release_acc = [&]() {   // this thread has read its last accumulator value of the tile
        tc_fence_before();
        if (TWO && rank != 0) mbar_arrive_remote(dsmem_addr(&acc_empty[u], 0));   // the issuer lives in the even CTA
        else mbar_arrive(&acc_empty[u]);
      };

      // Request the residual sub-tile that follows (tile w, sub-tile c) in this warp's sequence: its next sub-tile of this
      // tile or its first one of the next tile.  Two buffers: called (by the whole warp, after a __syncwarp that follows
      // the reads of the other buffer) when sub-tile c is taken up; one buffer: called right after sub-tile c's residual
      // has been read.
      auto prefetch_res = [&](int c) {
        if (!has_res || lane != 0) return;
        int ww = w, cc = c + 1;
        if (cc >= c_hi) { ww = w_next; cc = c_lo; }
        if (ww >= total_items) return;
        const int buf = (g_res + 1) % NB;
        mbar_expect_tx(&my_res_full[buf], res_bytes);
        tma_load_2d(my_res + buf * WCHUNK, &tmR, item_n0(ww) + cc * 32, item_m0(ww) + q * 32, &my_res_full[buf]);
      };
      // v = act(acc + bias) (+ residual from this warp's smem sub-tile)
      auto finish = [&](float (&v)[32], int c, int width) {
        const int c0 = c * 32;
        const uint32_t ba = bias_sa + (uint32_t)(n0 + c0) * 4;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          if (cc * 4 < width) {
            const float4 b4 = lds4(ba + cc * 16);
            v[cc * 4 + 0] += b4.x; v[cc * 4 + 1] += b4.y; v[cc * 4 + 2] += b4.z; v[cc * 4 + 3] += b4.w;
          }
        }
        // (the fused-LayerNorm kernels serve the FFT blocks and the predictors only: no tanh, no vocoder options)
        if (act == ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (!LN && act == ACT_TANH) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
        } else if (!LN && act == ACT_LRELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = v[j] >= 0.f ? v[j] : v[j] * p.slope;
        }
        if (has_res) {
#ifdef FS2_TRACE_BUILD
          const long long tw0 = clock64();
#endif
          mbar_wait(&my_res_full[g_res % NB], (g_res / NB) & 1);
#ifdef FS2_TRACE_BUILD
          res_wait_cycles += clock64() - tw0;
#endif
          const uint32_t rbase = res_sa + (g_res % NB) * WCHUNK;
          if (LN || !(p.res_inv_lrelu || p.res_bf16)) {   // plain fp32 residual
            const uint32_t rb = rbase + lane * 128;
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
              if (cc * 4 < width) {
                const float4 r4 = lds4(rb + ((cc << 4) ^ swz_x));
                v[cc * 4 + 0] += r4.x; v[cc * 4 + 1] += r4.y; v[cc * 4 + 2] += r4.z; v[cc * 4 + 3] += r4.w;
              }
            }
          } else {   // vocoder forms: bf16 residual and / or a residual stored as lrelu(x), undone on the fly
            const float inv = p.res_inv_lrelu ? 1.f / p.slope : 1.f;
            if (p.res_bf16) {   // [32 rows x 64 B], SWIZZLE_64B: 16-byte chunk ^= (row >> 1) & 3; bf16 -> fp32 is a 16-bit shift
              const uint32_t rb = rbase + lane * 64;
              const uint32_t sx64 = (uint32_t)((lane >> 1) & 3) << 4;
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) {
                if (cc * 8 < width) {
                  const float4 r4 = lds4(rb + ((cc << 4) ^ sx64));
                  const uint32_t wv[4] = {__float_as_uint(r4.x), __float_as_uint(r4.y), __float_as_uint(r4.z), __float_as_uint(r4.w)};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float lo = __uint_as_float(wv[e] << 16), hi = __uint_as_float(wv[e] & 0xFFFF0000u);
                    v[cc * 8 + 2 * e] += lo >= 0.f ? lo : lo * inv;
                    v[cc * 8 + 2 * e + 1] += hi >= 0.f ? hi : hi * inv;
                  }
                }
              }
            } else {
              const uint32_t rb = rbase + lane * 128;
#pragma unroll
              for (int cc = 0; cc < 8; ++cc) {
                if (cc * 4 < width) {
                  const float4 r4 = lds4(rb + ((cc << 4) ^ swz_x));
                  v[cc * 4 + 0] += r4.x >= 0.f ? r4.x : r4.x * inv; v[cc * 4 + 1] += r4.y >= 0.f ? r4.y : r4.y * inv;
                  v[cc * 4 + 2] += r4.z >= 0.f ? r4.z : r4.z * inv; v[cc * 4 + 3] += r4.w >= 0.f ? r4.w : r4.w * inv;
                }
              }
            }
          }
          if (NB == 1) {   // the single buffer is free again once every lane has read it
            __syncwarp();
            prefetch_res(c);
          }
          ++g_res;
        }
        if (!LN && p.act2 == ACT_LRELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = v[j] >= 0.f ? v[j] : v[j] * p.slope;
        }
      };
      // this warp's [32 x 32] sub-tile -> swizzled staging -> TMA store
      auto stage_out = [&](const float (&v)[32], int c0, int width) {
        const bool tr = lt == 0 && warp == 2 && c0 == 64;   // trace build: phases of the third sub-tile's store
        if (tr) stamp(49);
        if (lane == 0) bulk_wait_read<NB - 1>();   // the store that last used this staging buffer has read it
        __syncwarp();
        if (tr) stamp(50);
        uint8_t* sb = my_cst + (g_st % NB) * WCHUNK;
        const uint32_t sa = cst_sa + (g_st % NB) * WCHUNK + lane * 128;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          if (cc * 4 < width) sts4(sa + ((cc << 4) ^ swz_x), make_float4(v[cc * 4], v[cc * 4 + 1], v[cc * 4 + 2], v[cc * 4 + 3]));
        }
        if (tr) stamp(51);
        fence_async_smem();
        __syncwarp();
        if (tr) stamp(52);
        if (lane == 0) {
          tma_store_2d(&tmC, sb, n0 + c0, m0 + q * 32);
          bulk_commit();
        }
        if (tr) stamp(53);
        ++g_st;
      };
      // the same sub-tile rounded to bf16: [32 rows x 64 B], SWIZZLE_64B (16-byte chunk ^= (row >> 1) & 3)
      auto stage_out_b = [&](const float (&v)[32], int c0, int width) {
        if (lane == 0) bulk_wait_read<NB - 1>();
        __syncwarp();
        uint8_t* sb = my_cst + (g_st % NB) * WCHUNK;
        const uint32_t sa = cst_sa + (g_st % NB) * WCHUNK + lane * 64;
        const uint32_t sx = (uint32_t)((lane >> 1) & 3) << 4;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          if (cc * 8 < width) sts8_bf16(sa + ((cc << 4) ^ sx), &v[cc * 8]);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmC2, sb, n0 + c0, m0 + q * 32);
          bulk_commit();
        }
        ++g_st;
      };

      if (!LN) {
        if (c_lo >= c_hi) release_acc();   // (a narrow tile leaves the second warp of a quarter without sub-tiles)
        // K-split: ranks 1.. write their partial tile to the workspace and signal rank 0; rank 0 waits for all of them
        // Workspace layout [partial][sub-tile][16-byte piece][row]: thread = row on both sides, so a warp's 16-byte accesses
        // to one piece are 512 contiguous bytes (with a row-major tile every lane touched its own 128-byte line: 32
        // wavefronts per instruction, and the reduction of 3 partial tiles took 6 us)
        constexpr size_t SK_TILE = (size_t)C::NCHUNK * 8 * BM * 4;   // floats per partial tile
        float* sk_ws = KS > 1 ? p.splitk_ws + (size_t)w * (KS - 1) * SK_TILE : nullptr;
        if (KS > 1 && rank != 0) {
          float* dst = sk_ws + (size_t)(rank - 1) * SK_TILE + (size_t)r * 4;
#pragma unroll 1
          for (int c = c_lo; c < c_hi; ++c) {
            const int c0 = c * 32;
            const int width = (BN - c0) >= 32 ? 32 : 16;
            float v[32];
            if (width == 32) tmem_ld32(acc + c0, v); else tmem_ld16(acc + c0, v);
#pragma unroll
            for (int cc = 0; cc < 8; ++cc)
              if (cc * 4 < width)
                *reinterpret_cast<float4*>(dst + (size_t)(c * 8 + cc) * BM * 4) = make_float4(v[cc * 4], v[cc * 4 + 1], v[cc * 4 + 2], v[cc * 4 + 3]);
          }
          if (c_lo < c_hi) release_acc();
          // one remote arrival per warp (896 per-thread arrivals on one barrier took 15 us): __syncwarp orders the lanes'
          // stores before lane 0's release at cluster scope, which makes them visible to rank 0's acquiring wait
          __threadfence();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(dsmem_addr(sk_full, 0));
          continue;
        }
        if (KS > 1) {
          if (warp == 2) stamp(54);
          mbar_wait_cluster(sk_full, lt & 1);
          if (warp == 2) stamp(55);
        }
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
          const int c0 = c * 32;
          const int width = (BN - c0) >= 32 ? 32 : 16;
          float v[32];
          if (width == 32) tmem_ld32(acc + c0, v); else tmem_ld16(acc + c0, v);
          if constexpr (KS > 1) {
            // + the other ranks' partial sums.  Every load of the sub-tile is issued before the first add (3 x 8 16-byte
            // loads per thread in flight): with the loads trickling out eight at a time the reduction was a chain of L2
            // round trips (measured: 15 us for 7 partial tiles, against 3.3 us for the K loop itself)
            float4 t4[KS - 1][8];
#pragma unroll
            for (int pr = 0; pr < KS - 1; ++pr) {
              const float* src = sk_ws + (size_t)pr * SK_TILE + (size_t)r * 4;
#pragma unroll
              for (int cc = 0; cc < 8; ++cc)   // (first touch of these lines by this SM: nothing stale in L1)
                if (cc * 4 < width) t4[pr][cc] = *reinterpret_cast<const float4*>(src + (size_t)(c * 8 + cc) * BM * 4);
            }
#pragma unroll
            for (int pr = 0; pr < KS - 1; ++pr) {
#pragma unroll
              for (int cc = 0; cc < 8; ++cc) {
                if (cc * 4 < width) {
                  v[cc * 4] += t4[pr][cc].x; v[cc * 4 + 1] += t4[pr][cc].y; v[cc * 4 + 2] += t4[pr][cc].z; v[cc * 4 + 3] += t4[pr][cc].w;
                }
              }
            }
            if (warp == 2) stamp(56);
          }
          if (NB == 2) prefetch_res(c);   // all lanes passed the __syncwarp of the previous stage_out
          finish(v, c, width);
          if (!live) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          if (c == c_hi - 1) release_acc();   // all TMEM reads of this tile are done: release the accumulator
          if (has_out) stage_out(v, c0, width);
          if (has_out2) stage_out_b(v, c0, width);
        }
      } else {
        // ---- LayerNorm over the 256-wide row held in TMEM (eps 1e-5, biased variance).  Statistics are
        // exact two-pass inside each 32-column chunk (in registers) and merged across chunks with
        // Chan's update, so the accumulator is read only twice.
        float mean = 0.f, var = 0.f;   // var holds M2 = sum (v - mean)^2 until the merge is complete
        float va[32], vb[32];          // double-buffered chunk registers: the TMEM load of chunk c+1 is in
                                       // flight while chunk c is processed
        auto pass1 = [&](float (&v)[32], int c) {
          if (has_res && NB == 2) {
            __syncwarp();             // every lane finished reading the other residual buffer
            prefetch_res(c);
          }
          finish(v, c, 32);
          float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            s4[0] += v[j]; s4[1] += v[j + 1]; s4[2] += v[j + 2]; s4[3] += v[j + 3];
          }
          const float cm = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.f / 32.f);
          float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float d0 = v[j] - cm, d1 = v[j + 1] - cm, d2 = v[j + 2] - cm, d3 = v[j + 3] - cm;
            q4[0] = fmaf(d0, d0, q4[0]); q4[1] = fmaf(d1, d1, q4[1]); q4[2] = fmaf(d2, d2, q4[2]); q4[3] = fmaf(d3, d3, q4[3]);
          }
          const float cm2 = (q4[0] + q4[1]) + (q4[2] + q4[3]);
          const float delta = cm - mean;
          const int k = c - c_lo;                                  // chunks this warp has merged before this one
          const float inv_n = __frcp_rn((float)(k + 1));           // (no division in the loop)
          mean = fmaf(delta, inv_n, mean);
          var += cm2 + delta * delta * (32.f * (float)k * inv_n);
          tmem_st32(acc + c * 32, v);
          if (lt == 0 && warp == 2) stamp(32 + c);
        };
        if (lt == 0 && warp == 2) stamp(31);
        tmem_ld32_issue(acc + c_lo * 32, va);
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 2) {
          tmem_ld_wait();
          tmem_ld32_issue(acc + (c + 1) * 32, vb);
          pass1(va, c);
          tmem_ld_wait();
          if (c + 2 < c_hi) tmem_ld32_issue(acc + (c + 2) * 32, va);
          pass1(vb, c + 1);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        if (EH == 2) {
          // the two warps of a lane quarter hold (mean, M2) over 128 columns each: exchange through shared memory and merge
          // in the order (first half, second half), so that both normalise with bit-identical statistics
          float2* mine = pair_stat + ((lt & 1) * 2 + h) * BM + r;
          const float2* other = pair_stat + ((lt & 1) * 2 + (h ^ 1)) * BM + r;
          *mine = make_float2(mean, var);
          asm volatile("bar.sync %0, 64;\n" ::"r"(2 + q) : "memory");
          const float2 o = *other;
          const float m_a = h == 0 ? mean : o.x, q_a = h == 0 ? var : o.y, m_b = h == 0 ? o.x : mean, q_b = h == 0 ? o.y : var;
          const float d = m_b - m_a;
          mean = fmaf(d, 0.5f, m_a);
          var = (q_a + q_b) + d * d * 64.f;   // n_a n_b / (n_a + n_b) = 128 * 128 / 256
        }
        if (NS > 1) {
          // (mean, M2) over this CTA's BN columns -> every peer; then merge the NS groups in rank order (Chan), so all
          // CTAs of the cluster normalise with bit-identical statistics
          const int par = lt & 1;
#pragma unroll
          for (int pr = 0; pr < NS; ++pr) {
            if (pr == rank) continue;
            dsmem_st2(dsmem_addr(&stat_s[(par * 4 + rank) * BM + r], pr), mean, var);
            mbar_arrive_remote(dsmem_addr(&stat_full[par], pr));
          }
          mbar_wait_cluster(&stat_full[par], (lt >> 1) & 1);
          float gm = 0.f, gm2 = 0.f;
#pragma unroll
          for (int src = 0; src < NS; ++src) {
            float2 st = make_float2(mean, var);
            if (src != rank) st = stat_s[(par * 4 + src) * BM + r];
            const float delta = st.x - gm;
            const float n_old = (float)(BN * src), n_new = (float)(BN * (src + 1));
            gm = fmaf(delta, (float)BN / n_new, gm);
            gm2 += st.y + delta * delta * (n_old * (float)BN / n_new);
          }
          mean = gm;
          var = gm2;
        }
        const float rstd = 1.f / sqrtf(var * (1.f / 256.f) + 1e-5f);
        float dot = 0.f;
        auto pass2 = [&](float (&v)[32], int c) {
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) {
            const float4 g4 = lds4(gamma_sa + (n0 + c * 32) * 4 + cc * 16), b4 = lds4(beta_sa + (n0 + c * 32) * 4 + cc * 16);
            const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) v[cc * 4 + e] = fmaf((v[cc * 4 + e] - mean) * rstd, gg[e], bb[e]);
          }
          if (p.head_out != nullptr) {
            float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
              const float4 w4 = lds4(headw_sa + (n0 + c * 32) * 4 + cc * 16);
              d4[0] = fmaf(v[cc * 4], w4.x, d4[0]); d4[1] = fmaf(v[cc * 4 + 1], w4.y, d4[1]);
              d4[2] = fmaf(v[cc * 4 + 2], w4.z, d4[2]); d4[3] = fmaf(v[cc * 4 + 3], w4.w, d4[3]);
            }
            dot += (d4[0] + d4[1]) + (d4[2] + d4[3]);
          }
          if (!live) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          if (post_a != nullptr) {   // (x + speaker) + emotion, in the reference's order; 32 consecutive rows mostly share u
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
              const float4 a4 = __ldg(reinterpret_cast<const float4*>(post_a + n0 + c * 32 + cc * 4));
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(post_b + n0 + c * 32 + cc * 4));
              v[cc * 4] = (v[cc * 4] + a4.x) + b4.x; v[cc * 4 + 1] = (v[cc * 4 + 1] + a4.y) + b4.y;
              v[cc * 4 + 2] = (v[cc * 4 + 2] + a4.z) + b4.z; v[cc * 4 + 3] = (v[cc * 4 + 3] + a4.w) + b4.w;
            }
          }
          if (has_out) stage_out(v, c * 32, 32);
          if (has_out2) stage_out_b(v, c * 32, 32);
          if (lt == 0 && warp == 2) stamp(40 + c);
        };
        tmem_ld32_issue(acc + c_lo * 32, va);
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 2) {
          tmem_ld_wait();
          tmem_ld32_issue(acc + (c + 1) * 32, vb);
          pass2(va, c);
          tmem_ld_wait();
          if (c + 2 < c_hi) tmem_ld32_issue(acc + (c + 2) * 32, va);
          else release_acc();           // the last TMEM read of this tile has landed: release the accumulator
          pass2(vb, c + 1);
        }
        if (p.head_out != nullptr && live) {
          const int dst = p.slot != nullptr ? p.slot[row] : row;
          if (dst >= 0) {
            // two warps per row: each adds its half of the dot product to the zero-initialised slot (two terms: the
            // result does not depend on the order)
            if (EH == 2) atomicAdd(&p.head_out[dst], h == 0 ? dot + p.head_b[0] : dot);
            else p.head_out[dst] = dot + p.head_b[0];
          }
        }
      }
      if (lt < 6 && warp == 2) stamp(8 + lt * 4 + 3);
    }
    if (warp == 2) stamp(5);
#ifdef FS2_TRACE_BUILD
    if (p.trace != nullptr && blockIdx.x == 0 && warp == 2 && lane == 0) p.trace[48] = res_wait_cycles;
#endif
    if (lane == 0) bulk_wait_read<0>();   // smem must outlive the last TMA stores
    if (warp == 2) stamp(6);
  }
  tc_fence_before();
  __syncthreads();
  if (CSIZE > 1) cluster_sync_all();   // the peers may still multicast commits into this CTA's barriers until they are done too
  if (warp == 1) {
    if (TWO) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    stamp(7);
  }
}

template <int BN, bool LN, int CL, bool BF, int AR = 0, int NS = 1, bool TWO = false, int EW = 4, int KS = 1>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
conv_gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                     const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                     const __grid_constant__ CUtensorMap tmC2, const __grid_constant__ ConvGemmArgs p) {
  conv_gemm_tc2_body<BN, LN, CL, BF, AR, NS, TWO, EW, KS>(tmA, tmW, tmC, tmR, tmC2, p);
}

// Two independent contractions of the SAME shape in one launch (grid.y = 2; blockIdx.y picks the operand set): the duration
// and pitch predictors read the same rows (model/modules.py:115-121), so their conv + LayerNorm layers run side by side
// instead of back to back -- on the phoneme side a single predictor layer fills a quarter of the machine.
struct ConvGemmMaps {
  CUtensorMap A, W, C, R, C2;
};
template <int BN, bool LN, int CL, bool BF, int AR = 0, int NS = 1, bool TWO = false, int EW = 4, int KS = 1>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
conv_gemm_tc2_dual_kernel(const __grid_constant__ ConvGemmMaps m0, const __grid_constant__ ConvGemmArgs p0,
                          const __grid_constant__ ConvGemmMaps m1, const __grid_constant__ ConvGemmArgs p1) {
  if (blockIdx.y == 0) conv_gemm_tc2_body<BN, LN, CL, BF, AR, NS, TWO, EW, KS>(m0.A, m0.W, m0.C, m0.R, m0.C2, p0);
  else conv_gemm_tc2_body<BN, LN, CL, BF, AR, NS, TWO, EW, KS>(m1.A, m1.W, m1.C, m1.R, m1.C2, p1);
}

inline int sm_count() {
  static int n[64] = {};
  int dev = 0;
  FS2_CUDA_OK(cudaGetDevice(&dev));
  if (n[dev & 63] == 0) FS2_CUDA_OK(cudaDeviceGetAttribute(&n[dev & 63], cudaDevAttrMultiProcessorCount, dev));
  return n[dev & 63];
}

inline int& cluster_size_flag() {   // 2 = weight tiles multicast across CTA pairs (default), 1 = independent CTAs
  static int f = 2;
  return f;
}

inline int& epi_warps_flag() {   // 8 (default) = two epilogue warps per TMEM lane quarter where the variant has them, 4 = one
  static int f = [] {
    const char* e = std::getenv("FS2_EPI_WARPS");
    return e != nullptr ? std::atoi(e) : 8;
  }();
  return f;
}

template <int BN, bool LN, int CL, bool BF, int AR = 0, int NS = 1, bool TWO = false, int EW = 0, int KS = 1>
inline void launch_bn_cl(const ConvGemmArgs& a, cudaStream_t stream, const ConvGemmArgs* second = nullptr) {
  if constexpr (EW == 0) {   // pick the epilogue width: 8 warps for the full-width streaming variants
    constexpr bool CAN8 = NS == 1 && AR == 0 && (!LN || BN == 256) && BN >= 64;
    if constexpr (CAN8) {
      if (epi_warps_flag() == 8) { launch_bn_cl<BN, LN, CL, BF, AR, NS, TWO, 8, KS>(a, stream, second); return; }
    }
    launch_bn_cl<BN, LN, CL, BF, AR, NS, TWO, 4, KS>(a, stream, second);
    return;
  } else {
  using C = Cfg<BN, AR, NS, TWO>;
  static bool configured[64] = {};
  int dev = 0;
  FS2_CUDA_OK(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    FS2_CUDA_OK(cudaFuncSetAttribute(conv_gemm_tc2_kernel<BN, LN, CL, BF, AR, NS, TWO, EW, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
    configured[dev & 63] = true;
  }
  constexpr CUtensorMapSwizzle SW128 = CU_TENSOR_MAP_SWIZZLE_128B;
  auto maps_of = [&](const ConvGemmArgs& a) {
    ConvGemmMaps m;
    const int split = a.terms > 1 ? 2 : 1;   // 3xTF32: activations [rows, hi | lo], weights [hi block ; lo block]
    m.A = BF ? make_map_any(a.A, a.rows, a.K, a.lda, AR ? AR_ROWS : BM / NS, 64, MAP_BF16, SW128)
             : make_map(a.A, a.rows, (int64_t)a.K * split, a.lda, AR ? AR_ROWS : BM / NS, /*round_tf32=*/true, false);
    m.W = BF ? make_map_any(a.W, (int64_t)a.taps * a.N, a.K, a.K, BN / CL, 64, MAP_BF16, SW128)
             : make_map(a.W, (int64_t)a.taps * a.N * split, a.K, a.K, BN / CL, false, true);
    m.C = a.C != nullptr ? make_map(a.C, a.rows, a.N, a.ldc, 32, false, false) : m.A;
    m.R = a.residual == nullptr ? m.A
          : a.res_bf16 ? make_map_any(a.residual, a.rows, a.N, a.ldr, 32, 32, MAP_BF16, CU_TENSOR_MAP_SWIZZLE_64B)
                       : make_map(a.residual, a.rows, a.N, a.ldr, 32, false, false);
    m.C2 = a.C2 != nullptr ? make_map_any(a.C2, a.rows, a.N, a.ldc2, 32, 32, MAP_BF16, CU_TENSOR_MAP_SWIZZLE_64B) : m.A;
    return m;
  };
  const ConvGemmMaps m = maps_of(a);
  constexpr int CSIZE = CL > 1 ? CL : (NS > 1 ? NS : KS);
  const int items = NS > 1 ? (a.rows + BM - 1) / BM : (((a.rows + BM - 1) / BM + CL - 1) / CL) * ((a.N + BN - 1) / BN);
  const int grid = std::min(items, sm_count() / CSIZE) * CSIZE;
  require(KS == 1 || (items * KS <= sm_count() && a.splitk_ws != nullptr), FS2_ERR_INVALID, "K-split: one resident cluster per tile and a workspace");
  if (second != nullptr) {
    if constexpr (LN && KS == 1 && AR == 0) {   // the dual entry point exists for the fused-LayerNorm forms (the predictors)
      static bool configured2[64] = {};
      if (!configured2[dev & 63]) {
        FS2_CUDA_OK(cudaFuncSetAttribute(conv_gemm_tc2_dual_kernel<BN, LN, CL, BF, AR, NS, TWO, EW, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
        configured2[dev & 63] = true;
      }
      const ConvGemmMaps m1 = maps_of(*second);
      launch_pdl(conv_gemm_tc2_dual_kernel<BN, LN, CL, BF, AR, NS, TWO, EW, KS>, dim3(grid, 2), dim3(64 + 32 * EW), C::TOTAL, stream, CSIZE,
                 m, a, m1, *second);
      FS2_LAUNCHED();
      return;
    } else {
      throw Error(FS2_ERR_UNSUPPORTED, "paired launch: fused-LayerNorm contractions only");
    }
  }
  launch_pdl(conv_gemm_tc2_kernel<BN, LN, CL, BF, AR, NS, TWO, EW, KS>, dim3(grid), dim3(64 + 32 * EW), C::TOTAL, stream, CSIZE, m.A, m.W, m.C, m.R, m.C2, a);
  FS2_LAUNCHED();
  }
}

inline int& n_split_flag() {   // 1 = N-split fused-LN GEMMs for small row counts (default), 0 = always one CTA per row tile
  static int f = 1;
  return f;
}

constexpr size_t SPLITK_WS_BYTES = 8u << 20;   // K-split workspace a caller provides (ConvGemmArgs::splitk_ws)
inline int& k_split_flag() {   // 1 (default) = K-split clusters for single-row-tile GEMMs with long K loops, 0 = off (FS2_K_SPLIT)
  static int f = [] {
    const char* e = std::getenv("FS2_K_SPLIT");
    return e != nullptr ? std::atoi(e) : 1;
  }();
  return f;
}

inline int& a_resident_flag() {   // 1 = use the A-resident variant where it applies (default), 0 = always stream A
  static int f = 1;
  return f;
}

template <int BN>
inline void launch_ar(const ConvGemmArgs& a, cudaStream_t stream) {
  const bool pair = cluster_size_flag() == 2 && a.rows > BM;
  if (a.a_bf16) {   // 64 bf16 channels fit one 128-byte K chunk
    if (pair) launch_bn_cl<BN, false, 2, true, 1>(a, stream); else launch_bn_cl<BN, false, 1, true, 1>(a, stream);
  } else if (a.K <= 32) {
    if (pair) launch_bn_cl<BN, false, 2, false, 1>(a, stream); else launch_bn_cl<BN, false, 1, false, 1>(a, stream);
  } else {
    if (pair) launch_bn_cl<BN, false, 2, false, 2>(a, stream); else launch_bn_cl<BN, false, 1, false, 2>(a, stream);
  }
}

// 1 (default) = cta_group::2 MMAs for the 256-column plain GEMMs with a long K loop (conv9, PostNet 512 -> 512, the
// vocoder's 256-channel stage), 0 = off (FS2_TWO_SM / debug flag 6).  Measured at config 2: conv9 754 -> 834 TFLOP/s TF32,
// PostNet 0.327 -> 0.288 ms, forward 8.82 -> 9.26 M frames/s; the K = 256 single-tap QKV GEMM is 4 % slower with it and
// stays on the multicast form.
inline int& two_sm_flag() {
  static int f = [] {
    const char* e = std::getenv("FS2_TWO_SM");
    return e != nullptr ? std::atoi(e) : 1;
  }();
  return f;
}

template <int BN, bool LN>
inline void launch_bn(const ConvGemmArgs& a, cudaStream_t stream, const ConvGemmArgs* second = nullptr) {
  // a single row tile has no partner to share weights with
  const bool pair = cluster_size_flag() == 2 && a.rows > BM;
  if constexpr (BN == 256 && !LN) {
    if (pair && two_sm_flag() && a.N % 256 == 0 && a.taps * a.K >= 512) {
      if (a.a_bf16) launch_bn_cl<256, false, 2, true, 0, 1, true>(a, stream); else launch_bn_cl<256, false, 2, false, 0, 1, true>(a, stream);
      return;
    }
  }
  if (a.a_bf16) {
    if (pair) launch_bn_cl<BN, LN, 2, true>(a, stream, second); else launch_bn_cl<BN, LN, 1, true>(a, stream, second);
  } else {
    if (pair) launch_bn_cl<BN, LN, 2, false>(a, stream, second); else launch_bn_cl<BN, LN, 1, false>(a, stream, second);
  }
}

// second: an independent contraction of the same shape and options (fused-LayerNorm forms only) that shares the launch
inline void launch(const ConvGemmArgs& a, cudaStream_t stream, const ConvGemmArgs* second = nullptr) {   // the operand type travels with the arguments (a_bf16, terms)
  if (second != nullptr)
    require(a.ln_gamma != nullptr && second->ln_gamma != nullptr && a.rows == second->rows && a.K == second->K && a.N == second->N &&
                a.taps == second->taps && a.a_bf16 == second->a_bf16 && a.terms <= 1 && second->terms <= 1 &&
                (a.head_out != nullptr) == (second->head_out != nullptr) && (a.residual != nullptr) == (second->residual != nullptr) &&
                (a.C != nullptr) == (second->C != nullptr) && (a.C2 != nullptr) == (second->C2 != nullptr),
            FS2_ERR_INVALID, "paired launch: two fused-LayerNorm contractions of one shape");
  require(a.terms <= 1 || (a.terms == 3 && !a.a_bf16), FS2_ERR_INVALID, "split-operand contraction: three TF32 terms");
  const int am = a.a_bf16 ? 8 : 4;   // elements per 16 bytes of the A / W rows
  require(a.K % am == 0 && a.lda % am == 0 && (a.C == nullptr || a.ldc % 4 == 0) && (a.residual == nullptr || a.ldr % (a.res_bf16 ? 8 : 4) == 0) &&
              (a.C2 == nullptr || a.ldc2 % 8 == 0), FS2_ERR_INVALID,
          "tcgen05 conv_gemm: K and leading dimensions must describe 16-byte-aligned rows");
  require((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.W) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(a.C) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.residual) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(a.C2) & 15) == 0,
          FS2_ERR_INVALID, "tcgen05 conv_gemm: pointers must be 16-byte aligned");
  if (a.rows <= 0) return;
  require(a.N <= MAX_N, FS2_ERR_UNSUPPORTED, "tcgen05 conv_gemm: N > 2048");
  const bool ln = a.ln_gamma != nullptr;
  if (ln) {
    require(a.N == 256 && a.ln_beta != nullptr, FS2_ERR_INVALID, "fused LayerNorm needs N == 256 and both affine vectors");
    require(a.C != nullptr || a.C2 != nullptr || a.head_out != nullptr, FS2_ERR_INVALID, "fused LayerNorm: nothing to write");
    require(a.head_out == nullptr || (a.head_w != nullptr && a.head_b != nullptr), FS2_ERR_INVALID, "head needs weight and bias");
    require(a.post_a == nullptr || (a.post_b != nullptr && a.post_utt != nullptr && a.row_vpos != nullptr && a.mask_shift == 0),
            FS2_ERR_INVALID, "post-LayerNorm add needs both vectors and the row metadata");
    // few row tiles (single utterances, the encoder of a small batch): split the 256 columns over a cluster of 4 or 2 CTAs
    const int m_tiles_ln = (a.rows + BM - 1) / BM;
    static const int ns_force = [] { const char* e = std::getenv("FS2_LN_NS"); return e != nullptr ? std::atoi(e) : 0; }();
    int ns = (n_split_flag() == 0 || a.head_out != nullptr) ? 1 : (m_tiles_ln * 4 <= sm_count() ? 4 : (m_tiles_ln * 2 <= sm_count() ? 2 : 1));
    if (ns_force > 0 && a.head_out == nullptr && m_tiles_ln * 2 > sm_count()) ns = ns_force;   // experiment: N-split beyond one wave
    if (ns == 4) {
      if (a.a_bf16) launch_bn_cl<64, true, 1, true, 0, 4>(a, stream, second); else launch_bn_cl<64, true, 1, false, 0, 4>(a, stream, second);
    } else if (ns == 2) {
      if (a.a_bf16) launch_bn_cl<128, true, 1, true, 0, 2>(a, stream, second); else launch_bn_cl<128, true, 1, false, 0, 2>(a, stream, second);
    } else if (two_sm_flag() && cluster_size_flag() == 2 && a.rows > BM && a.taps * a.K >= 512) {
      // long K loop (w2 + LayerNorm, K = 1024): operand-delivery bound; the 2-SM form halves the weight bytes each CTA takes in
      if (a.a_bf16) launch_bn_cl<256, true, 2, true, 0, 1, true>(a, stream, second); else launch_bn_cl<256, true, 2, false, 0, 1, true>(a, stream, second);
    } else {
      launch_bn<256, true>(a, stream, second);
    }
    return;
  }
  require(a.C != nullptr || a.C2 != nullptr, FS2_ERR_INVALID, "conv_gemm: null output");
  {   // small K, several taps: keep the activation tile resident and shift the descriptor per tap
    const int d = a.dil > 0 ? a.dil : 1;
    if (a_resident_flag() && a.terms <= 1 && a.taps > 1 && a.K <= 64 && BM + (a.taps - 1) * d <= AR_ROWS) {
      if (a.N == 32) { launch_ar<32>(a, stream); return; }
      if (a.N == 64) { launch_ar<64>(a, stream); return; }
    }
  }
  // Small problems (single utterances): a 128 x 256 tile would leave most SMs idle while one CTA
  // walks the whole K loop at 512 cycles per stage, so narrower tiles spread the columns over more
  // CTAs whose stages are proportionally shorter.
  const int m_tiles = (a.rows + BM - 1) / BM;
  const int sms = sm_count();
  if (m_tiles == 1 && a.splitk_ws != nullptr && k_split_flag()) {
    // one utterance: the K loop is the critical path of the launch -- spread it over a cluster (KS CTAs per column tile)
    const int ke = a.a_bf16 ? 64 : 32;
    const int iters = (a.terms > 1 ? a.terms : 1) * a.taps * ((a.K + ke - 1) / ke);
    if (a.N % 64 == 0 || a.N == 80) {
      const int tiles = a.N == 80 ? 1 : a.N / 64;
      constexpr int KSPLIT = 4;   // 3 partial tiles: their loads fit the registers in one round (see the epilogue)
      const bool fits = (size_t)tiles * (KSPLIT - 1) * BM * (a.N == 80 ? 96 : 64) * sizeof(float) <= SPLITK_WS_BYTES;
      if (iters >= 32 && tiles * KSPLIT <= sms && fits) {
        if (a.N == 80) { if (a.a_bf16) launch_bn_cl<80, false, 1, true, 0, 1, false, 0, KSPLIT>(a, stream); else launch_bn_cl<80, false, 1, false, 0, 1, false, 0, KSPLIT>(a, stream); }
        else { if (a.a_bf16) launch_bn_cl<64, false, 1, true, 0, 1, false, 0, KSPLIT>(a, stream); else launch_bn_cl<64, false, 1, false, 0, 1, false, 0, KSPLIT>(a, stream); }
        return;
      }
    }
  }
  // experiment (FS2_QKV_BN=128): 128-column tiles for the single-tap K = 256 GEMMs whose 256-column tiles leave a partly
  // empty last wave (QKV at config 2: 630 tiles = 4.26 waves)
  static const int qkv_bn = [] { const char* e = std::getenv("FS2_QKV_BN"); return e != nullptr ? std::atoi(e) : 0; }();
  if (qkv_bn == 128 && a.taps == 1 && a.K == 256 && a.N == 768) { launch_bn<128, false>(a, stream); return; }
  if (a.N % 256 == 0 && m_tiles * (a.N / 256) * 4 <= sms) { launch_bn<64, false>(a, stream); return; }
  if (a.N % 256 == 0 && m_tiles * (a.N / 256) * 2 <= sms) { launch_bn<128, false>(a, stream); return; }
  if (a.N % 256 == 0) launch_bn<256, false>(a, stream);
  else if (a.N == 80) launch_bn<80, false>(a, stream);
  else if (a.N % 128 == 0) launch_bn<128, false>(a, stream);
  else if (a.N % 64 == 0) launch_bn<64, false>(a, stream);
  else if (a.N == 32) launch_bn<32, false>(a, stream);
  else throw Error(FS2_ERR_UNSUPPORTED, "tcgen05 conv_gemm: N = " + std::to_string(a.N) + " has no compiled tile");
}

}  // namespace tc2
}  // namespace fs2
