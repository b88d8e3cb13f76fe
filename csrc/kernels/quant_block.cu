// Blockwise absmax quantisation for frozen / low-precision weights (QLoRA-style): int8, NF4 and FP4 codes with one fp32
// absmax per block of `bs` consecutive elements; 4-bit codes are packed two per byte (even element in the high nibble).
// One warp owns one block: the absmax is a warp reduction, every lane then encodes its own pairs -- a single pass
// over the input, no intermediate tensors.  Dequantisation is a table lookup times the block scale.
// (capability parity: hetu/impl/kernel/quantization.cu:13 QuantizationCuda, :412 DeQuantizationCuda -- wrappers around
//  bitsandbytes' kQuantizeBlockwise / kDequantizeBlockwise)
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"

namespace hb {
namespace {

__constant__ float kNF4[16] = {-1.0f, -0.6961928f, -0.5250730f, -0.3949175f, -0.2844414f, -0.1848449f, -0.0910500f, 0.0f,
                               0.0795803f, 0.1609302f, 0.2461123f, 0.3379152f, 0.4407098f, 0.5626170f, 0.7229568f, 1.0f};
__constant__ float kFP4[16] = {0.0f, 0.0052083f, 0.6666667f, 1.0f, 0.3333333f, 0.5f, 0.1666667f, 0.25f,
                               -0.0f, -0.0052083f, -0.6666667f, -1.0f, -0.3333333f, -0.5f, -0.1666667f, -0.25f};

template <typename T>
__device__ __forceinline__ float load_as_float(const T* p, int64_t i);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p, int64_t i) { return p[i]; }
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) { return __bfloat162float(p[i]); }

__device__ __forceinline__ int nearest_code(float v, const float* table) {
  int best = 0;
  float bd = fabsf(v - table[0]);
#pragma unroll
  for (int c = 1; c < 16; ++c) {
    const float d = fabsf(v - table[c]);
    if (d < bd) { bd = d; best = c; }     // first minimum wins (matches argmin)
  }
  return best;
}

// kind: 0 int8, 1 nf4, 2 fp4
template <typename T>
__global__ void quantize_blockwise_kernel(const T* __restrict__ x, void* __restrict__ q, float* __restrict__ absmax, int64_t n,
                                          int bs, int kind) {
  const int lane = threadIdx.x & 31;
  const int64_t blk = blockIdx.x * int64_t(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nb = (n + bs - 1) / bs;
  if (blk >= nb) return;
  const int64_t base = blk * bs;
  float m = 0.f;
  for (int i = lane; i < bs; i += 32) {
    const int64_t k = base + i;
    if (k < n) m = fmaxf(m, fabsf(load_as_float(x, k)));
  }
  m = fmaxf(warp_max(m), 1e-12f);
  if (lane == 0) absmax[blk] = m;
  if (kind == 0) {
    int8_t* out = reinterpret_cast<int8_t*>(q);
    for (int i = lane; i < bs; i += 32) {
      const int64_t k = base + i;
      if (k < n) out[k] = (int8_t)fminf(fmaxf(rintf(load_as_float(x, k) / m * 127.0f), -127.f), 127.f);
    }
    return;
  }
  const float* table = kind == 2 ? kFP4 : kNF4;
  uint8_t* out = reinterpret_cast<uint8_t*>(q);
  for (int i = lane; i < bs / 2; i += 32) {                 // one byte = elements (2i, 2i+1) of the block
    const int64_t k0 = base + 2 * i, k1 = k0 + 1;
    const float v0 = k0 < n ? load_as_float(x, k0) / m : 0.f;     // true division: same normalised value as the host formulation
    const float v1 = k1 < n ? load_as_float(x, k1) / m : 0.f;
    out[(base >> 1) + i] = (uint8_t)((nearest_code(v0, table) << 4) | nearest_code(v1, table));
  }
}

template <typename T>
__device__ __forceinline__ void store_from_float(T* p, int64_t i, float v);
template <>
__device__ __forceinline__ void store_from_float<float>(float* p, int64_t i, float v) { p[i] = v; }
template <>
__device__ __forceinline__ void store_from_float<__nv_bfloat16>(__nv_bfloat16* p, int64_t i, float v) { p[i] = __float2bfloat16(v); }

template <typename T>
__global__ void dequantize_blockwise_kernel(const void* __restrict__ q, const float* __restrict__ absmax, T* __restrict__ y,
                                            int64_t n, int bs, int kind) {
  const float* table = kind == 2 ? kFP4 : kNF4;
  for (int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; k < n; k += int64_t(gridDim.x) * blockDim.x) {
    const float s = absmax[k / bs];
    float v;
    if (kind == 0) v = float(reinterpret_cast<const int8_t*>(q)[k]) * (1.0f / 127.0f);
    else {
      const uint8_t b = reinterpret_cast<const uint8_t*>(q)[k >> 1];
      v = table[(k & 1) ? (b & 15) : (b >> 4)];
    }
    store_from_float(y, k, v * s);
  }
}

}  // namespace

cudaError_t quantize_blockwise(const void* x, bool x_is_bf16, void* q, float* absmax, int64_t n, int blocksize, int kind,
                               cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  if (blocksize < 2 || (blocksize & 1) || kind < 0 || kind > 2) return cudaErrorInvalidValue;
  const int64_t nb = (n + blocksize - 1) / blocksize;
  const unsigned grid = (unsigned)((nb + 7) / 8);
  if (x_is_bf16) quantize_blockwise_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, q, absmax, n, blocksize, kind);
  else quantize_blockwise_kernel<float><<<grid, 256, 0, s>>>((const float*)x, q, absmax, n, blocksize, kind);
  count_launch();
  return cudaGetLastError();
}

cudaError_t dequantize_blockwise(const void* q, const float* absmax, void* y, bool y_is_bf16, int64_t n, int blocksize, int kind,
                                 cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  if (blocksize < 2 || kind < 0 || kind > 2) return cudaErrorInvalidValue;
  int64_t blocks = (n + 255) / 256;
  if (blocks > int64_t(sm_count()) * 16) blocks = int64_t(sm_count()) * 16;
  if (y_is_bf16) dequantize_blockwise_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, s>>>(q, absmax, (__nv_bfloat16*)y, n, blocksize, kind);
  else dequantize_blockwise_kernel<float><<<(unsigned)blocks, 256, 0, s>>>(q, absmax, (float*)y, n, blocksize, kind);
  count_launch();
  return cudaGetLastError();
}

}  // namespace hb
