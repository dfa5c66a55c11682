// tcgen05 + TMA implicit-GEMM engine (placeholder until the engine lands).
#pragma once
#include "common.cuh"
#include "gemm_mma.cuh"
namespace fs2 { namespace tc {
inline void launch(const ConvGemmArgs&, int, cudaStream_t) {
  throw Error(FS2_ERR_UNSUPPORTED, "tcgen05 engine not built yet");
}
}}
