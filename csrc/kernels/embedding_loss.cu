// Embedding gather/scatter and fused softmax-cross-entropy (single-GPU and
// vocab-parallel pieces) for bf16 logits.
//
// Capability parity: hetu/impl/kernel/EmbeddingLookup.cu:91,140;
// SoftmaxCrossEntropySparse.cu; VocabParallelCrossEntropyLoss.cu:57,114 and the
// host op hetu/graph/ops/VocabParallelCrossEntropyLoss.cc:58-110.
#include "common.cuh"
#include "kernels.h"

namespace hb {
namespace {

__global__ void embedding_fwd_kernel(const int64_t* __restrict__ ids, const int32_t* __restrict__ pos,
                                     const void* __restrict__ wte, const void* __restrict__ wpe, void* __restrict__ y,
                                     int64_t tokens, int hidden, int64_t vocab) {
  const int hv = hidden >> 3;
  const int64_t total = tokens * hv;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t t = i / hv;
    const int c = int(i - t * hv);
    int64_t id = ids[t];
    float f[8];
    if (id >= 0 && id < vocab) unpack8(ld8(wte, id * hv + c), f);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = 0.f;
    }
    if (wpe != nullptr) {
      float g[8];
      unpack8(ld8(wpe, int64_t(pos ? pos[t] : 0) * hv + c), g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += g[j];
    }
    st8(y, i, pack8(f));
  }
}

__global__ void embedding_bwd_kernel(const int64_t* __restrict__ ids, const int32_t* __restrict__ pos,
                                     const void* __restrict__ dy, float* __restrict__ dwte, float* __restrict__ dwpe,
                                     int64_t tokens, int hidden, int64_t vocab) {
  const int hv = hidden >> 3;
  const int64_t total = tokens * hv;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t t = i / hv;
    const int c = int(i - t * hv);
    float f[8];
    unpack8(ld8_stream(dy, i), f);
    const int64_t id = ids[t];
    if (dwte != nullptr && id >= 0 && id < vocab) {
      float* d = dwte + id * hidden + c * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(d + j, f[j]);
    }
    if (dwpe != nullptr) {
      float* d = dwpe + int64_t(pos ? pos[t] : 0) * hidden + c * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(d + j, f[j]);
    }
  }
}

constexpr int kCeThreads = 512;

// online (max, sum-exp) merge
struct MS { float m, s; };
__device__ __forceinline__ MS ms_merge(MS a, MS b) {
  const float m = fmaxf(a.m, b.m);
  MS r;
  r.m = m;
  r.s = (a.m == -INFINITY ? 0.f : a.s * __expf(a.m - m)) + (b.m == -INFINITY ? 0.f : b.s * __expf(b.m - m));
  return r;
}
__device__ __forceinline__ MS block_ms(MS v, MS* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MS t;
    t.m = __shfl_xor_sync(0xffffffffu, v.m, o);
    t.s = __shfl_xor_sync(0xffffffffu, v.s, o);
    v = ms_merge(v, t);
  }
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  MS r;
  if (lane < nw) r = red[lane]; else { r.m = -INFINITY; r.s = 0.f; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MS t;
    t.m = __shfl_xor_sync(0xffffffffu, r.m, o);
    t.s = __shfl_xor_sync(0xffffffffu, r.s, o);
    r = ms_merge(r, t);
  }
  return r;
}

__device__ __forceinline__ MS row_max_sum(const __nv_bfloat16* row, int cols, MS* red) {
  MS acc; acc.m = -INFINITY; acc.s = 0.f;
  const int nvec = cols >> 3;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float f[8];
    unpack8(ld8(row, v), f);
    float m = f[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) m = fmaxf(m, f[j]);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += __expf(f[j] - m);
    MS t; t.m = m; t.s = s;
    acc = ms_merge(acc, t);
  }
  for (int c = (nvec << 3) + threadIdx.x; c < cols; c += blockDim.x) {
    MS t; t.m = __bfloat162float(row[c]); t.s = 1.f;
    acc = ms_merge(acc, t);
  }
  return block_ms(acc, red);
}

__global__ void __launch_bounds__(kCeThreads) ce_fwd_bwd_kernel(__nv_bfloat16* __restrict__ logits,
                                                                const int64_t* __restrict__ labels,
                                                                float* __restrict__ loss, float* __restrict__ lse_out,
                                                                int cols, int64_t ld, int64_t ignore_index,
                                                                float grad_scale, const float* __restrict__ grad_scale_ptr,
                                                                int write_grad) {
  __shared__ MS red[32];
  if (grad_scale_ptr != nullptr) grad_scale *= *grad_scale_ptr;
  const int64_t r = blockIdx.x;
  __nv_bfloat16* row = logits + r * ld;
  const int64_t label = labels[r];
  const bool valid = (label != ignore_index) && label >= 0 && label < cols;
  const MS ms = row_max_sum(row, cols, red);
  const float lse = ms.m + __logf(ms.s);
  if (threadIdx.x == 0) {
    loss[r] = valid ? lse - __bfloat162float(row[label]) : 0.f;
    if (lse_out) lse_out[r] = lse;
  }
  if (!write_grad) return;
  __syncthreads();  // row[label] read above must precede the in-place overwrite
  const float gs = valid ? grad_scale : 0.f;
  const int nvec = cols >> 3;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float f[8];
    unpack8(ld8(row, v), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float p = __expf(f[j] - lse);
      if (v * 8 + j == label) p -= 1.0f;
      f[j] = p * gs;
    }
    st8(row, v, pack8(f));
  }
  for (int c = (nvec << 3) + threadIdx.x; c < cols; c += blockDim.x) {
    float p = __expf(__bfloat162float(row[c]) - lse);
    if (c == label) p -= 1.0f;
    row[c] = __float2bfloat16(p * gs);
  }
}

__global__ void __launch_bounds__(kCeThreads) vp_max_kernel(const __nv_bfloat16* __restrict__ logits,
                                                            float* __restrict__ row_max, int cols, int64_t ld) {
  __shared__ float red[32];
  const int64_t r = blockIdx.x;
  const __nv_bfloat16* row = logits + r * ld;
  float m = -INFINITY;
  const int nvec = cols >> 3;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float f[8];
    unpack8(ld8(row, v), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) m = fmaxf(m, f[j]);
  }
  for (int c = (nvec << 3) + threadIdx.x; c < cols; c += blockDim.x) m = fmaxf(m, __bfloat162float(row[c]));
  m = block_max(m, red);
  if (threadIdx.x == 0) row_max[r] = m;
}

__global__ void __launch_bounds__(kCeThreads) vp_sum_kernel(const __nv_bfloat16* __restrict__ logits,
                                                            const int64_t* __restrict__ labels,
                                                            const float* __restrict__ row_max, float* __restrict__ sum_exp,
                                                            float* __restrict__ target_logit, int cols, int64_t ld,
                                                            int64_t vocab_start) {
  __shared__ float red[32];
  const int64_t r = blockIdx.x;
  const __nv_bfloat16* row = logits + r * ld;
  const float m = row_max[r];
  float s = 0.f;
  const int nvec = cols >> 3;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float f[8];
    unpack8(ld8(row, v), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += __expf(f[j] - m);
  }
  for (int c = (nvec << 3) + threadIdx.x; c < cols; c += blockDim.x) s += __expf(__bfloat162float(row[c]) - m);
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    sum_exp[r] = s;
    const int64_t l = labels[r] - vocab_start;
    target_logit[r] = (l >= 0 && l < cols) ? __bfloat162float(row[l]) : 0.f;
  }
}

__global__ void __launch_bounds__(kCeThreads) vp_finish_kernel(__nv_bfloat16* __restrict__ logits,
                                                               const int64_t* __restrict__ labels,
                                                               const float* __restrict__ row_max,
                                                               const float* __restrict__ sum_exp,
                                                               const float* __restrict__ target_logit,
                                                               float* __restrict__ loss, int cols, int64_t ld,
                                                               int64_t vocab_start, int64_t ignore_index,
                                                               float grad_scale, int write_grad) {
  const int64_t r = blockIdx.x;
  __nv_bfloat16* row = logits + r * ld;
  const int64_t label = labels[r];
  const bool valid = label != ignore_index;
  const float lse = row_max[r] + __logf(sum_exp[r]);
  if (threadIdx.x == 0) loss[r] = valid ? lse - target_logit[r] : 0.f;
  if (!write_grad) return;
  const int64_t l = label - vocab_start;
  const float gs = valid ? grad_scale : 0.f;
  const int nvec = cols >> 3;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float f[8];
    unpack8(ld8(row, v), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float p = __expf(f[j] - lse);
      if (v * 8 + j == l) p -= 1.0f;
      f[j] = p * gs;
    }
    st8(row, v, pack8(f));
  }
  for (int c = (nvec << 3) + threadIdx.x; c < cols; c += blockDim.x) {
    float p = __expf(__bfloat162float(row[c]) - lse);
    if (c == l) p -= 1.0f;
    row[c] = __float2bfloat16(p * gs);
  }
}

inline int grid_for(int64_t n) {
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = int64_t(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

cudaError_t embedding_fwd(const int64_t* ids, const int32_t* pos, const void* wte, const void* wpe, void* y,
                          int64_t tokens, int hidden, int64_t vocab, cudaStream_t s) {
  if (tokens == 0) return cudaSuccess;
  if (hidden & 7) return cudaErrorInvalidValue;
  embedding_fwd_kernel<<<grid_for(tokens * (hidden >> 3)), 256, 0, s>>>(ids, pos, wte, wpe, y, tokens, hidden, vocab);
  count_launch();
  return cudaGetLastError();
}
cudaError_t embedding_bwd(const int64_t* ids, const int32_t* pos, const void* dy, float* dwte, float* dwpe,
                          int64_t tokens, int hidden, int64_t vocab, cudaStream_t s) {
  if (tokens == 0) return cudaSuccess;
  if (hidden & 7) return cudaErrorInvalidValue;
  embedding_bwd_kernel<<<grid_for(tokens * (hidden >> 3)), 256, 0, s>>>(ids, pos, dy, dwte, dwpe, tokens, hidden, vocab);
  count_launch();
  return cudaGetLastError();
}
cudaError_t softmax_ce_fwd_bwd(void* logits, const int64_t* labels, float* loss, float* lse, int64_t rows, int cols,
                               int64_t ld, int64_t ignore_index, float grad_scale, bool write_grad, cudaStream_t s,
                               const float* grad_scale_ptr) {
  if (rows == 0) return cudaSuccess;
  if ((ld & 7) || (reinterpret_cast<uintptr_t>(logits) & 15)) return cudaErrorMisalignedAddress;
  ce_fwd_bwd_kernel<<<(unsigned)rows, kCeThreads, 0, s>>>((__nv_bfloat16*)logits, labels, loss, lse, cols, ld,
                                                          ignore_index, grad_scale, grad_scale_ptr, write_grad ? 1 : 0);
  count_launch();
  return cudaGetLastError();
}
cudaError_t vp_ce_local_max(const void* logits, float* row_max, int64_t rows, int cols, int64_t ld, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  if ((ld & 7) || (reinterpret_cast<uintptr_t>(logits) & 15)) return cudaErrorMisalignedAddress;
  vp_max_kernel<<<(unsigned)rows, kCeThreads, 0, s>>>((const __nv_bfloat16*)logits, row_max, cols, ld);
  count_launch();
  return cudaGetLastError();
}
cudaError_t vp_ce_local_sum(const void* logits, const int64_t* labels, const float* row_max, float* sum_exp,
                            float* target_logit, int64_t rows, int cols, int64_t ld, int64_t vocab_start,
                            cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  vp_sum_kernel<<<(unsigned)rows, kCeThreads, 0, s>>>((const __nv_bfloat16*)logits, labels, row_max, sum_exp,
                                                      target_logit, cols, ld, vocab_start);
  count_launch();
  return cudaGetLastError();
}
cudaError_t vp_ce_finish(void* logits, const int64_t* labels, const float* row_max, const float* sum_exp,
                         const float* target_logit, float* loss, int64_t rows, int cols, int64_t ld,
                         int64_t vocab_start, int64_t ignore_index, float grad_scale, bool write_grad, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  vp_finish_kernel<<<(unsigned)rows, kCeThreads, 0, s>>>((__nv_bfloat16*)logits, labels, row_max, sum_exp, target_logit,
                                                         loss, cols, ld, vocab_start, ignore_index, grad_scale,
                                                         write_grad ? 1 : 0);
  count_launch();
  return cudaGetLastError();
}

}  // namespace hb
