// HBM-bound kernels of the path: row layout, embedding + positional encoding, conditioning,
// LayerNorm, bucketize + embedding add, durations + scans, length regulator, output unpack,
// weight repack.  All activations are fp32, token-major, 256 columns unless noted; a warp owns
// one row and moves it with 16-byte accesses.
#pragma once

#include <cuda_bf16.h>

#include "common.cuh"

namespace fs2 {

// device-side status word: nonzero => the host raises after the stage's read-back
enum { ERR_BAD_LEN = 1, ERR_BAD_ID = 2, ERR_MAXLEN_SMALL = 4, ERR_BAD_INDEX = 8 };

// First statement of the row kernels of the forward: when the kernel is launched with the programmatic-serialization
// attribute (fs2_api.cu: launch_row) its blocks may become resident while the previous kernel is still running -- nothing is
// read or written before the wait returns (= the previous kernel has completed and its writes are visible) -- and the next
// kernel may start its own prologue.  Launched the ordinary way both instructions are no-ops.
__device__ __forceinline__ void row_pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// optional bf16 copy of a row (BF16 mode: the copy is the A operand of the next contraction); no-op when p == nullptr
__device__ __forceinline__ void st4b(__nv_bfloat16* p, float4 v) {
  if (p == nullptr) return;
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------
// Row layout: exclusive scan of (len + gap) over the batch.  One block.
// totals[0] = rows in use (end of the last gap), totals[1] = max_len (max over lens, or the
// caller's value when forced_max > 0), totals[2] = sum of lens.  (utils/tools.py:152-160 builds
// masks from the same lengths; here they define the packed layout instead.)
template <typename LenT>
__global__ void layout_scan_kernel(const LenT* __restrict__ lens, int batch, int gap, int len_limit,
                                   int forced_max, int32_t* __restrict__ starts, int32_t* __restrict__ lens32,
                                   int64_t* __restrict__ totals, int32_t* __restrict__ status,
                                   int64_t* __restrict__ host_out = nullptr) {
  // host_out (pinned host memory, device-addressable): the three totals and the accumulated status word are also written
  // there and the status word is cleared for the next forward -- the caller then needs a stream synchronisation only, not
  // two device->host copies and a memset behind this kernel.
  row_pdl_sync();
  // 1024 threads = 32 warps: inclusive scan of the per-thread sums by warp shuffles, one exchange of the 32 warp totals through
  // shared memory, and the maxima / real-length sums reduced the same way (two block barriers; the Hillis-Steele scan over
  // 1024 shared-memory partials took ~30 of them and 5 us on the critical path of both stages)
  __shared__ long long w_sum[32], w_real[32];
  __shared__ int w_max[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (batch + blockDim.x - 1) / blockDim.x;
  const int lo = min(tid * per, batch), hi = min(lo + per, batch);
  long long s = 0;
  int mx = 0;
  long long real = 0;
  for (int b = lo; b < hi; ++b) {
    long long l = (long long)lens[b];
    if (l < 0 || (len_limit > 0 && l > len_limit)) {
      atomicOr(status, ERR_BAD_LEN);
      l = l < 0 ? 0 : len_limit;
    }
    s += l + gap;
    real += l;
    mx = max(mx, (int)l);
  }
  long long incl = s;   // inclusive scan inside the warp
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const long long v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    real += __shfl_xor_sync(0xffffffffu, real, off);
  }
  if (lane == 31) w_sum[warp] = incl;
  if (lane == 0) {
    w_max[warp] = mx;
    w_real[warp] = real;
  }
  __syncthreads();
  long long before = 0, total_s = 0, total_real = 0;   // sum of the warps before this one; block totals
  int max_len = 0;
  for (int w = 0; w < 32; ++w) {
    const long long v = w_sum[w];
    if (w < warp) before += v;
    total_s += v;
    total_real += w_real[w];
    max_len = max(max_len, w_max[w]);
  }
  long long run = gap + before + (incl - s);
  for (int b = lo; b < hi; ++b) {
    long long l = (long long)lens[b];
    l = l < 0 ? 0 : ((len_limit > 0 && l > len_limit) ? len_limit : l);
    starts[b] = (int32_t)run;
    lens32[b] = (int32_t)l;
    run += l + gap;
  }
  const long long total_rows = gap + total_s;
  if (tid == 0) {
    if (forced_max > 0) {
      if (forced_max < max_len) atomicOr(status, ERR_MAXLEN_SMALL);
      max_len = forced_max;
    }
    starts[batch] = (int32_t)total_rows;
    totals[0] = total_rows;
    totals[1] = max_len;
    totals[2] = total_real;
    if (host_out != nullptr) {   // (every earlier kernel of the stage has completed; this block's own atomics precede the barriers above)
      host_out[0] = total_rows;
      host_out[1] = max_len;
      host_out[2] = total_real;
      host_out[3] = (int64_t)atomicExch(status, 0);
      __threadfence_system();
    }
  }
}

// ---------------------------------------------------------------------------------------
// Work list of the attention kernel (transformer/SubLayers.py:42-52 runs every utterance padded to the batch maximum;
// here one entry = 128 queries of one utterance).  Entries are (utterance << 16 | query tile), utterances in DESCENDING
// order of their key-tile count -- the cost of each of their entries -- so that the hardware's in-order CTA dispatch is a
// longest-processing-time-first schedule: long items go out first and the tail of the grid is made of short ones.
// Counting sort over 1024 cost classes with shared-memory atomics; the order inside a class is arbitrary (the entries are
// independent, the results do not depend on it).  Called by every thread of ONE block; bins = 1024 ints of shared memory.
// q_rows = query rows one entry covers: 128, or 256 for the paired attention kernel (attention_tc.cuh).
__device__ inline void build_attention_work(const int32_t* __restrict__ lens32, int batch, uint32_t* __restrict__ work,
                                            int cap, int32_t* __restrict__ count, int* bins, int q_rows) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < 1024; i += nt) bins[i] = 0;
  __syncthreads();
  for (int b = tid; b < batch; b += nt) {
    const int l = lens32[b];
    if (l > 0) atomicAdd(&bins[1023 - min((l + 63) >> 6, 1023)], (l + q_rows - 1) / q_rows);
  }
  __syncthreads();
  if (tid < 32) {   // exclusive scan of the 1024 classes by one warp: 32 consecutive classes per lane
    int v[32], sum = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) { v[i] = bins[tid * 32 + i]; sum += v[i]; }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (tid >= o) inc += t;
    }
    int run = inc - sum;
#pragma unroll
    for (int i = 0; i < 32; ++i) { bins[tid * 32 + i] = run; run += v[i]; }
    if (tid == 31) *count = min(inc, cap);
  }
  __syncthreads();
  for (int b = tid; b < batch; b += nt) {
    const int l = lens32[b];
    if (l <= 0) continue;
    const int q = (l + q_rows - 1) / q_rows;
    const int pos = atomicAdd(&bins[1023 - min((l + 63) >> 6, 1023)], q);
    for (int t = 0; t < q; ++t)
      if (pos + t < cap) work[pos + t] = ((uint32_t)b << 16) | (uint32_t)t;
  }
}

__global__ void attention_work_kernel(const int32_t* __restrict__ lens32, int batch, uint32_t* __restrict__ work, int cap,
                                      int32_t* __restrict__ count, int q_rows) {
  __shared__ int bins[1024];
  build_attention_work(lens32, batch, work, cap, count, bins, q_rows);
}

// [B, L] arrays the forward fills only at real positions (predictions scattered through `slot`) start from zero, and the
// source padding mask (utils/tools.py:152-160) is written, by the same grid that builds the row metadata: one launch
// instead of a memset per array.
struct SlotInit {
  float* zero[6];            // each [n] or nullptr
  int64_t n;                 // B * L
  uint8_t* src_mask;         // [n] or nullptr: 1 = padding
  const int64_t* src_lens;   // [B]
  int max_src_len;
};

// Per-row metadata from the starts.  max_len comes from a device scalar (totals[1]) when
// max_len_dev != nullptr (frame side: T_max is only known on the device when this is enqueued).
__global__ void row_meta_kernel(const int32_t* __restrict__ starts, const int32_t* __restrict__ lens, int batch,
                                int gap, int max_len_host, const int64_t* __restrict__ max_len_dev, int rows_alloc,
                                int32_t* __restrict__ utt, int32_t* __restrict__ vpos, int32_t* __restrict__ room,
                                int32_t* __restrict__ slot, uint32_t* __restrict__ work = nullptr, int work_cap = 0,
                                int32_t* __restrict__ work_count = nullptr, SlotInit init = SlotInit{}, int work_q_rows = 128) {
  row_pdl_sync();
  if (work != nullptr && blockIdx.x == gridDim.x - 1) {   // the last block also builds the attention work list of this side
    __shared__ int bins[1024];
    build_attention_work(lens, batch, work, work_cap, work_count, bins, work_q_rows);
  }
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = r; i < init.n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (init.zero[k] != nullptr) init.zero[k][i] = 0.f;
    if (init.src_mask != nullptr) init.src_mask[i] = (i % init.max_src_len) >= init.src_lens[i / init.max_src_len] ? 1 : 0;
  }
  if (r >= rows_alloc) return;
  const int max_len = max_len_dev ? (int)*max_len_dev : max_len_host;
  int u = -1, vp = VPOS_DEAD, rm = 0, sl = -1;
  if (batch > 0 && r >= starts[0]) {
    int lo = 0, hi = batch - 1;  // largest b with starts[b] <= r
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (starts[mid] <= r) lo = mid; else hi = mid - 1;
    }
    const int pos = r - starts[lo], len = lens[lo];
    if (pos < len + gap) {
      u = lo;
      vp = pos - len;
      rm = max_len - len;
      if (pos < max_len) sl = lo * max_len + pos;
    }
  }
  utt[r] = u;
  vpos[r] = vp;
  room[r] = rm;
  slot[r] = sl;
}

// ---------------------------------------------------------------------------------------
// Encoder input: src_word_emb[texts] + position_enc[:L]  (transformer/Models.py:82-91).
__global__ void embed_pe_kernel(const int64_t* __restrict__ texts, int max_src_len, const float* __restrict__ emb,
                                int n_vocab, const float* __restrict__ pe, RowMeta meta, const int32_t* __restrict__ lens,
                                int rows, float* __restrict__ x, int32_t* __restrict__ status,
                                __nv_bfloat16* __restrict__ xb = nullptr) {
  row_pdl_sync();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int u = meta.utt[row], vp = meta.vpos[row];
  float* dst = x + (size_t)row * D_MODEL;
  __nv_bfloat16* dstb = xb != nullptr ? xb + (size_t)row * D_MODEL : nullptr;
  if (u < 0 || vp >= 0) {
    st4(dst + lane * 4, make_float4(0, 0, 0, 0));
    st4(dst + 128 + lane * 4, make_float4(0, 0, 0, 0));
    st4b(dstb ? dstb + lane * 4 : nullptr, make_float4(0, 0, 0, 0));
    st4b(dstb ? dstb + 128 + lane * 4 : nullptr, make_float4(0, 0, 0, 0));
    return;
  }
  const int pos = vp + lens[u];
  long long id = texts[(size_t)u * max_src_len + pos];
  if (id < 0 || id >= n_vocab) {
    if (lane == 0) atomicOr(status, ERR_BAD_ID);
    id = 0;
  }
  const float* e = emb + (size_t)id * D_MODEL;
  const float* p = pe + (size_t)pos * D_MODEL;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = h * 128 + lane * 4;
    const float4 v = add4(ld4(e + c), ld4(p + c));
    st4(dst + c, v);
    st4b(dstb ? dstb + c : nullptr, v);
  }
}

// ---------------------------------------------------------------------------------------
// Conditioning vectors (model/fastspeech2.py:101-110): spk[b] = speaker_emb[s];
// emo[b] = ReLU(W . cat(emotion_emb[e], arousal_emb[a], valence_emb[v]) + bias).  COND_PARTS blocks per utterance, each
// owning 256 / COND_PARTS output rows (a block per utterance walked eight dependent passes over W: ~5 us on the forward's
// critical path, and on a single utterance's).
constexpr int COND_PARTS = 4;
__global__ void cond_kernel(const int64_t* __restrict__ speakers, const int64_t* __restrict__ emotions,
                            const int64_t* __restrict__ arousals, const int64_t* __restrict__ valences,
                            const float* __restrict__ spk_emb, int n_spk, const float* __restrict__ emo_emb, int n_emo,
                            const float* __restrict__ aro_emb, int n_aro, const float* __restrict__ val_emb, int n_val,
                            const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ spk_out,
                            float* __restrict__ emo_out, int32_t* __restrict__ status) {
  row_pdl_sync();
  __shared__ float e[D_MODEL];
  const int b = blockIdx.x / COND_PARTS, part = blockIdx.x % COND_PARTS, tid = threadIdx.x;
  long long s = speakers[b], em = emotions[b], ar = arousals[b], va = valences[b];
  if (s < 0 || s >= n_spk || em < 0 || em >= n_emo || ar < 0 || ar >= n_aro || va < 0 || va >= n_val) {
    if (tid == 0 && part == 0) atomicOr(status, ERR_BAD_INDEX);
    s = min(max(s, 0LL), (long long)n_spk - 1);
    em = min(max(em, 0LL), (long long)n_emo - 1);
    ar = min(max(ar, 0LL), (long long)n_aro - 1);
    va = min(max(va, 0LL), (long long)n_val - 1);
  }
  if (tid < 128) e[tid] = emo_emb[em * 128 + tid];
  else if (tid < 192) e[tid] = aro_emb[ar * 64 + (tid - 128)];
  else e[tid] = val_emb[va * 64 + (tid - 192)];
  if (part == 0) spk_out[(size_t)b * D_MODEL + tid] = spk_emb[s * D_MODEL + tid];
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  // each warp owns 32 / COND_PARTS output rows; four rows per pass keep 32 independent loads in flight per lane (the
  // row-at-a-time loop was a chain of 32 dependent L2 round trips: 24 us for 64 tiny blocks)
  constexpr int ROWS_W = 32 / COND_PARTS;
  const int row_lo = part * (D_MODEL / COND_PARTS) + warp * ROWS_W;
  for (int n0 = row_lo; n0 < row_lo + ROWS_W; n0 += 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = lane; k < D_MODEL; k += 32) {
      const float ek = e[k];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(W[(size_t)(n0 + j) * D_MODEL + k], ek, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float v = warp_sum(acc[j]);
      if (lane == 0) emo_out[(size_t)b * D_MODEL + n0 + j] = fmaxf(v + bias[n0 + j], 0.f);
    }
  }
}

// x_cond = (enc + spk[b]) + emo[b] on real rows AND on the first min(2, L_max - L_b) reserved
// rows (the reference adds the vectors on padding rows too and the predictors read them --
// SURVEY.md B.4); every other row is zero.
__global__ void add_cond_kernel(const float* __restrict__ x, RowMeta meta, const float* __restrict__ spk,
                                const float* __restrict__ emo, int extra, int rows, float* __restrict__ y,
                                __nv_bfloat16* __restrict__ yb = nullptr) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int u = meta.utt[row];
  float* dst = y + (size_t)row * D_MODEL;
  __nv_bfloat16* dstb = yb != nullptr ? yb + (size_t)row * D_MODEL : nullptr;
  if (u < 0 || !row_live(meta.vpos[row], meta.room[row], extra)) {
    st4(dst + lane * 4, make_float4(0, 0, 0, 0));
    st4(dst + 128 + lane * 4, make_float4(0, 0, 0, 0));
    st4b(dstb ? dstb + lane * 4 : nullptr, make_float4(0, 0, 0, 0));
    st4b(dstb ? dstb + 128 + lane * 4 : nullptr, make_float4(0, 0, 0, 0));
    return;
  }
  const float* src = x + (size_t)row * D_MODEL;
  const float* s = spk + (size_t)u * D_MODEL;
  const float* e = emo + (size_t)u * D_MODEL;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = h * 128 + lane * 4;
    const float4 v = add4(add4(ld4(src + c), ld4(s + c)), ld4(e + c));
    st4(dst + c, v);
    st4b(dstb ? dstb + c : nullptr, v);
  }
}

// ---------------------------------------------------------------------------------------
// torch.bucketize(v, bins, right=False): number of boundaries strictly below v; NaN -> n_bins.
__device__ __forceinline__ int bucket_of(float v, const float* __restrict__ bins, int n_bins) {
  int lo = 0, hi = n_bins;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (bins[mid] >= v) hi = mid; else lo = mid + 1;
  }
  return lo;
}

__global__ void bucketize_kernel(const float* __restrict__ v, int64_t n, const float* __restrict__ bins, int n_bins,
                                 int32_t* __restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) idx[i] = bucket_of(v[i], bins, n_bins);
}

// get_pitch_embedding / get_energy_embedding (model/modules.py:80-100) + the add (:121,:126):
// value = target ? target : raw * control; returned prediction = target ? raw : raw * control;
// y[row] = x[row] + table[bucketize(value)].  Runs on real rows and on the `extra` virtual rows
// (whose raw prediction is the masked 0, modules.py:248-249); other rows are zeroed.
__global__ void bucket_embed_add_kernel(const float* __restrict__ x, RowMeta meta, const int32_t* __restrict__ slot,
                                        int extra, int rows, const float* __restrict__ raw,
                                        const float* __restrict__ target, float control,
                                        const float* __restrict__ bins, int n_bins, const float* __restrict__ table,
                                        float* __restrict__ pred_out, int32_t* __restrict__ idx_out,
                                        float* __restrict__ y, __nv_bfloat16* __restrict__ yb = nullptr) {
  row_pdl_sync();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* dst = y + (size_t)row * D_MODEL;
  __nv_bfloat16* dstb = yb != nullptr ? yb + (size_t)row * D_MODEL : nullptr;
  const int sl = slot[row];
  if (meta.utt[row] < 0 || sl < 0 || !row_live(meta.vpos[row], meta.room[row], extra)) {
    st4(dst + lane * 4, make_float4(0, 0, 0, 0));
    st4(dst + 128 + lane * 4, make_float4(0, 0, 0, 0));
    st4b(dstb ? dstb + lane * 4 : nullptr, make_float4(0, 0, 0, 0));
    st4b(dstb ? dstb + 128 + lane * 4 : nullptr, make_float4(0, 0, 0, 0));
    return;
  }
  const bool real = meta.vpos[row] < 0;
  const float r = real ? raw[sl] : 0.f;
  const float scaled = r * control;
  const float value = target != nullptr ? target[sl] : scaled;
  const int idx = bucket_of(value, bins, n_bins);
  if (lane == 0 && real) {
    pred_out[sl] = target != nullptr ? r : scaled;
    if (idx_out != nullptr) idx_out[sl] = idx;
  }
  const float* src = x + (size_t)row * D_MODEL;
  const float* e = table + (size_t)idx * D_MODEL;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = h * 128 + lane * 4;
    const float4 v = add4(ld4(src + c), ld4(e + c));
    st4(dst + c, v);
    st4b(dstb ? dstb + c : nullptr, v);
  }
}

// ---------------------------------------------------------------------------------------
// Durations (model/modules.py:132-135) -> repeat counts max(int(d),0) (modules.py:186-187) ->
// inclusive scan per utterance.  Warp per utterance.  Padding positions expand to nothing.
// When energy_out != nullptr the same pass also writes the energy prediction the forward returns (modules.py:93-100:
// raw * control, or the raw prediction when a target is embedded instead; 0 at padding) -- its bucketize + embedding add
// happens inside the length regulator.
__global__ void durations_kernel(const float* __restrict__ d_in, int is_target, float d_control,
                                 const int64_t* __restrict__ src_lens, int batch, int max_src_len,
                                 float* __restrict__ d_rounded, int32_t* __restrict__ cum,
                                 int64_t* __restrict__ mel_lens, int32_t* __restrict__ mel_lens32,
                                 const float* __restrict__ energy_raw = nullptr, float energy_scale = 1.f,
                                 float* __restrict__ energy_out = nullptr) {
  row_pdl_sync();
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= batch) return;
  long long len = src_lens[b];
  len = len < 0 ? 0 : (len > max_src_len ? max_src_len : len);
  int run = 0;
  for (int j0 = 0; j0 < max_src_len; j0 += 32) {
    const int j = j0 + lane;
    int reps = 0;
    if (j < max_src_len) {
      const size_t i = (size_t)b * max_src_len + j;
      float d;
      if (is_target) {
        d = d_in[i];
      } else {
        const float logd = j < len ? d_in[i] : 0.f;           // masked_fill(mask, 0) (modules.py:248-249)
        d = fmaxf(rintf(expf(logd) - 1.f) * d_control, 0.f);   // round half-to-even BEFORE scaling
        if (d_rounded != nullptr) d_rounded[i] = d;
      }
      if (j < len) {
        const float t = truncf(d);
        reps = t > 0.f ? (t < 1048576.f ? (int)t : 1048576) : 0;
      }
      if (energy_out != nullptr) energy_out[i] = j < len ? energy_raw[i] * energy_scale : 0.f;
    }
    int inc = reps;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (j < max_src_len) cum[(size_t)b * max_src_len + j] = run + inc;
    run += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) {
    if (mel_lens != nullptr) mel_lens[b] = run;
    if (mel_lens32 != nullptr) mel_lens32[b] = run;
  }
}

__device__ __forceinline__ int phoneme_of_frame(const int32_t* __restrict__ cum, int n, int t) {
  int lo = 0, hi = n;  // first j with cum[j] > t
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cum[mid] > t) hi = mid; else lo = mid + 1;
  }
  return lo;
}

__global__ void frame_map_kernel(const int32_t* __restrict__ cum, int batch, int max_src_len, int max_mel_len,
                                 int32_t* __restrict__ map) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)batch * max_mel_len) return;
  const int b = (int)(i / max_mel_len), t = (int)(i % max_mel_len);
  const int32_t* c = cum + (size_t)b * max_src_len;
  map[i] = t < c[max_src_len - 1] ? phoneme_of_frame(c, max_src_len, t) : -1;
}

// LengthRegulator (model/modules.py:167-194, utils/tools.py:360-378) fused with the decoder's
// positional-encoding add (transformer/Models.py:145-162): frame row -> phoneme row by binary
// search in the scan, one coalesced 1 KB row copy per warp; reserved rows are zeroed.
__global__ void length_regulate_kernel(const float* __restrict__ x, const int32_t* __restrict__ p_starts,
                                       const int32_t* __restrict__ cum, int max_src_len, RowMeta fmeta,
                                       const int32_t* __restrict__ f_lens, const float* __restrict__ pe, int rows,
                                       float* __restrict__ y, __nv_bfloat16* __restrict__ yb = nullptr) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int u = fmeta.utt[row], vp = fmeta.vpos[row];
  float* dst = y + (size_t)row * D_MODEL;
  __nv_bfloat16* dstb = yb != nullptr ? yb + (size_t)row * D_MODEL : nullptr;
  if (u < 0 || vp >= 0) {
    st4(dst + lane * 4, make_float4(0, 0, 0, 0));
    st4(dst + 128 + lane * 4, make_float4(0, 0, 0, 0));
    st4b(dstb ? dstb + lane * 4 : nullptr, make_float4(0, 0, 0, 0));
    st4b(dstb ? dstb + 128 + lane * 4 : nullptr, make_float4(0, 0, 0, 0));
    return;
  }
  const int t = vp + f_lens[u];
  const int j = phoneme_of_frame(cum + (size_t)u * max_src_len, max_src_len, t);
  const float* src = x + (size_t)(p_starts[u] + j) * D_MODEL;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = h * 128 + lane * 4;
    float4 v = ld4(src + c);
    if (pe != nullptr) v = add4(v, ld4(pe + (size_t)t * D_MODEL + c));   // nullptr: a frame_level predictor runs first
    st4(dst + c, v);
    st4b(dstb ? dstb + c : nullptr, v);
  }
}

// Phoneme-level energy (model/modules.py:122-126): the bucketize + embedding add of the energy feature is applied to the
// phoneme row while it sits in registers, instead of in a pass of its own.  table == nullptr: off.
struct EnergyAdd {
  const float* raw;      // [B, L] raw predictions (0 at padding)
  const float* target;   // [B, L] or nullptr
  float control;         // p_control (sic: modules.py:123-125)
  const float* bins;
  int n_bins;
  const float* table;    // [256, 256]
  float* va_out;         // optional [rows_p, 256] copy of x + embedding (debug tap "va_x")
};

// The same expansion driven from the SOURCE side (the product path): a warp owns one phoneme row, keeps its 1 KB in
// registers and streams it to the frames [cum[j-1], cum[j]) it expands to, adding the positional row of each frame.  No
// per-frame binary search and one dependent load chain per phoneme instead of per frame, so the kernel is bound by the
// coalesced 1 KB stores.  Warps beyond the phoneme rows zero the reserved frame rows (leading gap, the GAP rows after
// every utterance, the tail up to `rows_f`).
__global__ void length_regulate_scatter_kernel(const float* __restrict__ x, RowMeta pmeta, const int32_t* __restrict__ p_lens,
                                               int rows_p, const int32_t* __restrict__ cum, int max_src_len,
                                               const int32_t* __restrict__ f_starts, const int32_t* __restrict__ f_lens,
                                               int batch, int gap, const int64_t* __restrict__ f_totals,
                                               const float* __restrict__ pe, int rows_f, float* __restrict__ y,
                                               __nv_bfloat16* __restrict__ yb, EnergyAdd en = EnergyAdd{}) {
  row_pdl_sync();
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const float4 zero = make_float4(0, 0, 0, 0);
  auto put = [&](int row, int c, float4 v) {
    st4(y + (size_t)row * D_MODEL + c, v);
    if (yb != nullptr) st4b(yb + (size_t)row * D_MODEL + c, v);
  };
  if (w < rows_p) {
    const int u = pmeta.utt[w], vp = pmeta.vpos[w];
    if (u < 0 || vp >= 0) return;                      // reserved phoneme row: expands to nothing
    const int j = vp + p_lens[u];
    const int32_t* c = cum + (size_t)u * max_src_len;
    const int t0 = j > 0 ? c[j - 1] : 0, t1 = c[j];
    if (t1 <= t0 && en.va_out == nullptr) return;      // (the debug tap wants the row even when it expands to nothing)
    const float* src = x + (size_t)w * D_MODEL;
    float4 a = ld4(src + lane * 4), b = ld4(src + 128 + lane * 4);
    if (en.table != nullptr) {   // x + energy_embedding[bucketize(energy)] (model/modules.py:93-100,126), fused into the expansion
      const size_t sl = (size_t)u * max_src_len + j;
      const float value = en.target != nullptr ? en.target[sl] : en.raw[sl] * en.control;
      const float* e = en.table + (size_t)bucket_of(value, en.bins, en.n_bins) * D_MODEL;
      a = add4(a, ld4(e + lane * 4));
      b = add4(b, ld4(e + 128 + lane * 4));
      if (en.va_out != nullptr) {   // debug tap: the variance adaptor's output row
        st4(en.va_out + (size_t)w * D_MODEL + lane * 4, a);
        st4(en.va_out + (size_t)w * D_MODEL + 128 + lane * 4, b);
      }
    }
    if (t1 <= t0) return;
    const int base = f_starts[u];
    if (pe == nullptr) {                               // a frame_level predictor runs first: plain copies
      for (int t = t0; t < t1; ++t) {
        put(base + t, lane * 4, a);
        put(base + t, 128 + lane * 4, b);
      }
      return;
    }
    int t = t0;
    for (; t + 4 <= t1; t += 4) {                      // four positional rows in flight per lane
      float4 pa[4], pb[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        pa[k] = ld4(pe + (size_t)(t + k) * D_MODEL + lane * 4);
        pb[k] = ld4(pe + (size_t)(t + k) * D_MODEL + 128 + lane * 4);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        put(base + t + k, lane * 4, add4(a, pa[k]));
        put(base + t + k, 128 + lane * 4, add4(b, pb[k]));
      }
    }
    for (; t < t1; ++t) {
      put(base + t, lane * 4, add4(a, ld4(pe + (size_t)t * D_MODEL + lane * 4)));
      put(base + t, 128 + lane * 4, add4(b, ld4(pe + (size_t)t * D_MODEL + 128 + lane * 4)));
    }
    return;
  }
  // reserved frame rows: g = 0 .. (batch + 1) * gap - 1  ->  leading gap, then the gap after each utterance; then the tail
  const int g = w - rows_p;
  int row;
  if (g < gap) {
    row = g;
  } else if (g < (batch + 1) * gap) {
    const int b = g / gap - 1;
    row = f_starts[b] + f_lens[b] + g % gap;
  } else {
    row = (int)f_totals[0] + (g - (batch + 1) * gap);
  }
  if (row < rows_f) {
    put(row, lane * 4, zero);
    put(row, 128 + lane * 4, zero);
  }
}

// ---------------------------------------------------------------------------------------
// Hand-over of the length regulator's input between contexts (multi-GPU rebalancing by FRAMES: the durations are only
// known after stage 1).  export: padded [B, L, 256] rows = x + energy_embedding[bucketize(energy)]
// (what model/modules.py:126 feeds the LengthRegulator) and the integer repeat counts; import: the inverse.
__global__ void export_rows_kernel(const float* __restrict__ x, const int32_t* __restrict__ p_starts,
                                   const int32_t* __restrict__ p_lens, const int32_t* __restrict__ cum, int batch,
                                   int max_src_len, EnergyAdd en, float* __restrict__ hidden, int32_t* __restrict__ reps) {
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= batch * max_src_len) return;
  const int b = w / max_src_len, j = w - b * max_src_len;
  float* dst = hidden + (size_t)w * D_MODEL;
  if (j >= p_lens[b]) {
    st4(dst + lane * 4, make_float4(0, 0, 0, 0));
    st4(dst + 128 + lane * 4, make_float4(0, 0, 0, 0));
    if (lane == 0) reps[w] = 0;
    return;
  }
  const float* src = x + (size_t)(p_starts[b] + j) * D_MODEL;
  float4 a = ld4(src + lane * 4), c = ld4(src + 128 + lane * 4);
  if (en.table != nullptr) {
    const float value = en.target != nullptr ? en.target[w] : en.raw[w] * en.control;
    const float* e = en.table + (size_t)bucket_of(value, en.bins, en.n_bins) * D_MODEL;
    a = add4(a, ld4(e + lane * 4));
    c = add4(c, ld4(e + 128 + lane * 4));
  }
  st4(dst + lane * 4, a);
  st4(dst + 128 + lane * 4, c);
  if (lane == 0) reps[w] = cum[w] - (j > 0 ? cum[w - 1] : 0);
}

__global__ void import_rows_kernel(const float* __restrict__ hidden, RowMeta meta, const int32_t* __restrict__ p_lens,
                                   int max_src_len, int rows, float* __restrict__ x) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int u = meta.utt[row], vp = meta.vpos[row];
  float* dst = x + (size_t)row * D_MODEL;
  if (u < 0 || vp >= 0) {
    st4(dst + lane * 4, make_float4(0, 0, 0, 0));
    st4(dst + 128 + lane * 4, make_float4(0, 0, 0, 0));
    return;
  }
  const float* src = hidden + ((size_t)u * max_src_len + (vp + p_lens[u])) * D_MODEL;
  st4(dst + lane * 4, ld4(src + lane * 4));
  st4(dst + 128 + lane * 4, ld4(src + 128 + lane * 4));
}

// Inclusive scan of given repeat counts per utterance (the second half of durations_kernel).  Warp per utterance.
__global__ void reps_scan_kernel(const int32_t* __restrict__ reps, const int64_t* __restrict__ src_lens, int batch,
                                 int max_src_len, int32_t* __restrict__ cum, int64_t* __restrict__ mel_lens,
                                 int32_t* __restrict__ mel_lens32) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= batch) return;
  long long len = src_lens[b];
  len = len < 0 ? 0 : (len > max_src_len ? max_src_len : len);
  int run = 0;
  for (int j0 = 0; j0 < max_src_len; j0 += 32) {
    const int j = j0 + lane;
    int inc = (j < len) ? max(reps[(size_t)b * max_src_len + j], 0) : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (j < max_src_len) cum[(size_t)b * max_src_len + j] = run + inc;
    run += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) {
    if (mel_lens != nullptr) mel_lens[b] = run;
    mel_lens32[b] = run;
  }
}

// Decoder input when a frame_level predictor sits between the LengthRegulator and the decoder
// (model/modules.py:139-148, transformer/Models.py:154-162): real rows get x + position_enc[t], every
// reserved row returns to zero (the FFT stacks need zero gaps).  In place is fine (row-wise).
__global__ void add_pe_kernel(const float* __restrict__ x, RowMeta fmeta, const int32_t* __restrict__ f_lens,
                              const float* __restrict__ pe, int rows, float* __restrict__ y,
                              __nv_bfloat16* __restrict__ yb = nullptr) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int u = fmeta.utt[row], vp = fmeta.vpos[row];
  float* dst = y + (size_t)row * D_MODEL;
  __nv_bfloat16* dstb = yb != nullptr ? yb + (size_t)row * D_MODEL : nullptr;
  const bool real = u >= 0 && vp < 0;
  const float* src = x + (size_t)row * D_MODEL;
  const float* p = pe + (size_t)(real ? vp + f_lens[u] : 0) * D_MODEL;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = h * 128 + lane * 4;
    const float4 v = real ? add4(ld4(src + c), ld4(p + c)) : make_float4(0, 0, 0, 0);
    st4(dst + c, v);
    st4b(dstb ? dstb + c : nullptr, v);
  }
}

// ---------------------------------------------------------------------------------------
// Packed -> padded outputs (model/fastspeech2.py:138-149): mel / postnet [B,T_max,80] and the
// mel mask.  Padding rows of both tensors receive mel_linear.bias (what the reference's mel has
// there; its postnet padding rows are outside the contract).  Warp per output row.
// Thread per 16 bytes of the padded output (20 per row): consecutive threads write consecutive memory, every lane works.
__global__ void unpack_mel_kernel(const float* __restrict__ mel_p, const float* __restrict__ post_p,
                                  const int32_t* __restrict__ f_starts, const int32_t* __restrict__ f_lens, int batch,
                                  int max_mel_len, const float* __restrict__ bias, float* __restrict__ mel,
                                  float* __restrict__ post, uint8_t* __restrict__ mask) {
  row_pdl_sync();
  constexpr int Q = N_MEL / 4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)batch * max_mel_len * Q) return;
  const int64_t o = i / Q;
  const int q = (int)(i - o * Q);
  const int b = (int)(o / max_mel_len), t = (int)(o - (int64_t)b * max_mel_len);
  const bool real = t < f_lens[b];
  if (q == 0 && mask != nullptr) mask[o] = real ? 0 : 1;
  float4 m, p;
  if (real) {
    const size_t src = (size_t)(f_starts[b] + t) * N_MEL + q * 4;
    m = ld4(mel_p + src);
    p = ld4(post_p + src);
  } else {
    m = p = ld4(bias + q * 4);
  }
  st4(mel + i * 4, m);
  st4(post + i * 4, p);
}

// ---------------------------------------------------------------------------------------
// Weight repack (once, in fs2_prepare).
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// Conv1d weight [Cout][Cin][k] (torch) -> [k][Cout][Cin], scaled per output channel (BatchNorm
// folding: transformer/Layers.py:129-137 with eval-mode BatchNorm1d), operands rounded to TF32.
__global__ void repack_conv_kernel(const float* __restrict__ w, int cout, int cin, int k,
                                   const float* __restrict__ scale, int round_operand, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)cout * cin * k) return;
  const int ci = (int)(i % cin);
  const int co = (int)((i / cin) % cout);
  const int t = (int)(i / ((int64_t)cin * cout));
  float v = w[((size_t)co * cin + ci) * k + t];
  if (scale != nullptr) v *= scale[co];
  out[i] = round_operand ? round_tf32(v) : v;
}

// FS2_MATH_TF32X3: w = hi + lo with both parts exactly representable in TF32; out = [hi block ; lo block], each [k][Cout][Cin].
__global__ void repack_conv_split_kernel(const float* __restrict__ w, int cout, int cin, int k,
                                         const float* __restrict__ scale, float* __restrict__ out, int64_t lo_off) {
  const int64_t n = (int64_t)cout * cin * k;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int ci = (int)(i % cin);
  const int co = (int)((i / cin) % cout);
  const int t = (int)(i / ((int64_t)cin * cout));
  float v = w[((size_t)co * cin + ci) * k + t];
  if (scale != nullptr) v *= scale[co];
  const float hi = round_tf32(v);
  out[i] = hi;
  out[lo_off + i] = round_tf32(v - hi);   // lo_off = distance from a hi element to its lo element (>= n)
}

// Activations of a split-operand contraction: x [rows, K] (pitch ldx) -> out [rows, 2K] = [hi | lo].
__global__ void split_tf32_kernel(const float* __restrict__ x, int rows, int K, int ldx, float* __restrict__ out) {
  const int k4 = K >> 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * k4) return;
  const int64_t r = i / k4;
  const int c = (int)(i - r * k4) * 4;
  const float4 v = ld4(x + r * ldx + c);
  const float4 hi = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
  const float4 lo = make_float4(round_tf32(v.x - hi.x), round_tf32(v.y - hi.y), round_tf32(v.z - hi.z), round_tf32(v.w - hi.w));
  st4(out + r * 2 * K + c, hi);
  st4(out + r * 2 * K + K + c, lo);
}

// The same repack with the operands rounded to bf16 (round to nearest even) -- FS2_MATH_BF16.
__global__ void repack_conv_bf16_kernel(const float* __restrict__ w, int cout, int cin, int k,
                                        const float* __restrict__ scale, __nv_bfloat16* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)cout * cin * k) return;
  const int ci = (int)(i % cin);
  const int co = (int)((i / cin) % cout);
  const int t = (int)(i / ((int64_t)cin * cout));
  float v = w[((size_t)co * cin + ci) * k + t];
  if (scale != nullptr) v *= scale[co];
  out[i] = __float2bfloat16_rn(v);
}

// s = gamma / sqrt(var + eps);  b' = (b - mean) * s + beta
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var,
                               const float* __restrict__ conv_bias, int n, float* __restrict__ scale,
                               float* __restrict__ bias_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double s = (double)gamma[i] / sqrt((double)var[i] + 1e-5);
  scale[i] = (float)s;
  bias_out[i] = (float)(((double)conv_bias[i] - (double)mean[i]) * s + (double)beta[i]);
}

}  // namespace fs2
