// Host API of the hand-written sm_100a GEMM family (no torch dependency).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hb {

enum GemmAct : int { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2, ACT_GELU_TANH = 3, ACT_SILU = 4 };
enum GemmAuxMode : int {
  AUX_NONE = 0,
  AUX_ADD = 1,        // out = acc + aux_in                       (residual add)
  AUX_DGELU = 2,      // out = acc * gelu'(aux_in)                (fc2 dgrad -> d(pre-activation))
  AUX_DRELU = 3,      // out = acc * (aux_in > 0)
  AUX_DGELU_TANH = 4,
  AUX_DSILU = 5,
};

enum class GemmOut : int { BF16 = 0, FP32 = 1 };

// C[M,N] = epi(alpha * A[M,K] * B[K,N]).
//   a_mn_major == false: A stored [M, lda] (K contiguous);  true: A stored [K, lda] (M contiguous)
//   b_mn_major == false: B stored [N, ldb] (K contiguous);  true: B stored [K, ldb] (N contiguous)
struct GemmCall {
  const void* A = nullptr;
  const void* B = nullptr;
  void* C = nullptr;
  int M = 0, N = 0, K = 0;
  int64_t lda = 0, ldb = 0, ldc = 0;
  bool a_mn_major = false, b_mn_major = false;
  GemmOut out = GemmOut::BF16;
  const void* bias = nullptr;     // bf16 [N]
  const void* aux_in = nullptr;   // bf16 [M, ld_aux]
  void* aux_out = nullptr;        // bf16 [M, ld_aux]
  int64_t ld_aux = 0;
  int act = 0;        // GemmAct
  int aux_mode = 0;   // GemmAuxMode
  bool accumulate = false;
  float alpha = 1.0f;
  // fused GEMM -> reduce-scatter over symmetric memory (see GemmParams::peer_c)
  void* const* peer_c = nullptr;   // host array of `world` mapped staging buffers, each [world, rows_per_rank, N]
  int world = 1, my_rank = 0, rows_per_rank = 0;
  int cta_group = 0;  // 0 = auto (2), 1 or 2 to force
  int block_n = 0;    // accumulator columns per tile: 0 = auto (wave-quantisation heuristic), 256 or 128 to force (128: bf16, cta_group 2)
  // fused all-gather -> GEMM (see GemmParams::ag_src): A is the local gather target, ag_src[r] rank r's symmetric shard
  const void* const* ag_src = nullptr;
  uint32_t* ag_flags = nullptr;
  int ag_world = 1, ag_rank = 0, ag_rows_per_rank = 0;
  // fp8 path: A [M, lda] and B [N, ldb] hold e4m3 bytes (K-major), the result is scaled by row_scale[m] * col_scale[n]
  bool fp8 = false;
  const float* row_scale = nullptr;
  const float* col_scale = nullptr;
};

// Returns cudaSuccess or an error; never silently falls back to a library.
cudaError_t gemm_bf16(const GemmCall& call, cudaStream_t stream);

// Number of kernel launches issued by this module since process start.
int64_t gemm_launch_count();

}  // namespace hb
