// Shared declarations for libfs2b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/fs2_b200.h"

namespace fs2 {

// ---- fixed architecture (config/ESD-Chinese-Singing-MFA/model.yaml) -----------------------
constexpr int D_MODEL = 256;
constexpr int N_HEAD = 2;
constexpr int D_HEAD = 128;
constexpr int D_INNER = 1024;
constexpr int FFN_TAPS = 9;
constexpr int ENC_LAYERS = 4;
constexpr int DEC_LAYERS = 6;
constexpr int VP_TAPS = 3;
constexpr int N_BINS = 256;
constexpr int N_MEL = 80;
constexpr int PN_DIM = 512;
constexpr int PN_TAPS = 5;
constexpr int PN_LAYERS = 5;

// ---- packed, gapped, token-major row layout --------------------------------------------
// Utterance b owns rows [start_b, start_b + len_b); it is followed by GAP reserved rows and the
// whole buffer starts with GAP reserved rows:   start_0 = GAP, start_{b+1} = start_b + len_b + GAP.
// Reserved rows carry what the reference's padding rows carry at that point of the network
// (zeros inside the FFT stacks; conditioning rows for the predictors; mel_linear.bias rows
// shrinking by 2 per PostNet layer -- SURVEY.md B.4) and the last rows of every gap are always
// zero, so one flat convolution over the buffer equals the reference's per-utterance padded one.
constexpr int GAP_PHON = 4;    // >= 4 zero halo rows for the k=9 conv; 2 virtual + 2 zero rows for the predictors
constexpr int GAP_FRAME = 12;  // 10 PostNet virtual rows + 2 zero rows; >= 4 for the k=9 conv
constexpr int PN_VIRTUAL = 10; // PostNet receptive field: 5 layers x 2 rows

// Per-row metadata: vpos = pos - len (negative on real rows, 0.. on the trailing gap of the same
// utterance, INT_MAX/2 on the leading gap and beyond the end); room = max_len - len.
// A row is "live" for a kernel with extra = e  iff  vpos < min(e, room).
constexpr int VPOS_DEAD = 1 << 29;

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct RowMeta {
  const int32_t* utt;   // utterance id, -1 for dead rows
  const int32_t* vpos;
  const int32_t* room;
};

__device__ __forceinline__ bool row_live(int vpos, int room, int extra) {
  return vpos < (extra < room ? extra : room);
}

#define FS2_CUDA_OK(call)                                                             \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      (void)cudaGetLastError(); /* clear the sticky state: the error travels as an exception */ \
      char buf__[512];                                                                 \
      snprintf(buf__, sizeof(buf__), "%s:%d: %s -> %s", __FILE__, __LINE__, #call,     \
               cudaGetErrorString(e__));                                               \
      throw fs2::Error(FS2_ERR_CUDA, buf__);                                           \
    }                                                                                  \
  } while (0)

struct Error {
  int code;
  std::string msg;
  Error(int c, std::string m) : code(c), msg(std::move(m)) {}
};

inline void require(bool ok, int code, const std::string& msg) {
  if (!ok) throw Error(code, msg);
}

// Launch counter (the bench's gpu_launches claim is counted here, not estimated).
extern thread_local int g_launches;
#define FS2_LAUNCHED()                         \
  do {                                         \
    ++fs2::g_launches;                         \
    FS2_CUDA_OK(cudaPeekAtLastError());        \
  } while (0)

}  // namespace fs2
