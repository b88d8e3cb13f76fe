// fp8 (e4m3) quantisation for the block-scaled training GEMMs: one fp32 scale per 1 x K block (= per row of the
// K-major operand), so the scales factor out of the contraction and are applied in the GEMM epilogue
// (C = diag(sa) * (A8 B8^T) * diag(sb)).  Two kernels:
//   quantize_rowwise_e4m3       x[R, C] bf16 -> q[R, C] e4m3, scale[R]          (activations, weights [N, K])
//   quantize_transpose_e4m3     w[R, C] bf16 -> q[C, R] e4m3, scale[C]          (W^T for the input-gradient GEMM)
// (capability parity: the reference's 4-bit/8-bit paths go through bitsandbytes -- hetu/impl/kernel/Quantization.cu)
#include <cuda_fp8.h>

#include "common.cuh"
#include "kernels.h"

namespace hb {
namespace {

constexpr float kE4M3Max = 448.0f;

__device__ __forceinline__ uint32_t pack4_e4m3(float a, float b, float c, float d) {
  const __nv_fp8x2_storage_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
  const __nv_fp8x2_storage_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
  return uint32_t(lo) | (uint32_t(hi) << 16);
}

__global__ void __launch_bounds__(256) quant_rowwise_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ q,
                                                            float* __restrict__ scale, int cols, int64_t ldx, int64_t ldq) {
  __shared__ float red[32];
  const int64_t r = blockIdx.x;
  const __nv_bfloat16* row = x + r * ldx;
  const int nvec = cols >> 3;
  float amax = 0.f;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float f[8];
    unpack8(ld8(row, v), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) amax = fmaxf(amax, fabsf(f[j]));
  }
  for (int c = (nvec << 3) + threadIdx.x; c < cols; c += blockDim.x) amax = fmaxf(amax, fabsf(__bfloat162float(row[c])));
  amax = block_max(amax, red);
  // amax * (1 / 448) (not amax / 448): the same rounding as frameworks that divide by a scalar through its reciprocal
  const float sc = amax > 0.f ? amax * (1.0f / kE4M3Max) : 1.0f;
  if (threadIdx.x == 0) scale[r] = sc;
  uint8_t* qrow = q + r * ldq;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {   // second read of the row hits L1/L2
    float f[8];
    unpack8(ld8(row, v), f);
    uint2 o;
    o.x = pack4_e4m3(f[0] / sc, f[1] / sc, f[2] / sc, f[3] / sc);     // true division: bit-identical to x / scale references
    o.y = pack4_e4m3(f[4] / sc, f[5] / sc, f[6] / sc, f[7] / sc);
    reinterpret_cast<uint2*>(qrow)[v] = o;
  }
  for (int c = (nvec << 3) + threadIdx.x; c < cols; c += blockDim.x)
    qrow[c] = (uint8_t)__nv_cvt_float_to_fp8(__bfloat162float(row[c]) / sc, __NV_SATFINITE, __NV_E4M3);
}

// column-wise absolute maximum -> scale[c] (atomicMax on the float bit pattern: values are non-negative)
__global__ void __launch_bounds__(256) col_amax_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ amax, int64_t rows,
                                                       int cols, int rows_per_block) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  const int64_t r0 = int64_t(blockIdx.y) * rows_per_block;
  const int64_t r1 = min(r0 + rows_per_block, rows);
  float m = 0.f;
  for (int64_t r = r0; r < r1; ++r) m = fmaxf(m, fabsf(__bfloat162float(x[r * cols + c])));
  atomicMax(reinterpret_cast<int*>(amax + c), __float_as_int(m));
}
__global__ void amax_to_scale_kernel(float* __restrict__ s, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) s[i] = s[i] > 0.f ? s[i] * (1.0f / kE4M3Max) : 1.0f;
}
// q[c, r] = e4m3(x[r, c] / scale[c]) through a 32 x 32 shared tile
__global__ void __launch_bounds__(256) quant_transpose_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale,
                                                              uint8_t* __restrict__ q, int64_t rows, int cols, int64_t ldq) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t r0 = int64_t(blockIdx.y) * 32;
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i;
    const int c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? __bfloat162float(x[r * cols + c]) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i;
    const int64_t r = r0 + tx;
    if (c < cols && r < rows)
      q[int64_t(c) * ldq + r] = (uint8_t)__nv_cvt_float_to_fp8(tile[tx][i] / scale[c], __NV_SATFINITE, __NV_E4M3);
  }
}

}  // namespace

cudaError_t quantize_rowwise_e4m3(const void* x, void* q, float* scale, int64_t rows, int cols, int64_t ldx, int64_t ldq,
                                  cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(q) & 7) || (ldx & 7) || (ldq & 7)) return cudaErrorMisalignedAddress;
  quant_rowwise_kernel<<<(unsigned)rows, 256, 0, s>>>((const __nv_bfloat16*)x, (uint8_t*)q, scale, cols, ldx, ldq);
  count_launch();
  return cudaGetLastError();
}
cudaError_t quantize_transpose_e4m3(const void* x, void* q, float* scale, int64_t rows, int cols, int64_t ldq, cudaStream_t s) {
  if (rows == 0 || cols == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(scale, 0, sizeof(float) * cols, s);
  if (e != cudaSuccess) return e;
  const int rpb = 256;
  col_amax_kernel<<<dim3((cols + 255) / 256, (unsigned)((rows + rpb - 1) / rpb)), 256, 0, s>>>((const __nv_bfloat16*)x, scale, rows, cols, rpb);
  amax_to_scale_kernel<<<(cols + 255) / 256, 256, 0, s>>>(scale, cols);
  quant_transpose_kernel<<<dim3((cols + 31) / 32, (unsigned)((rows + 31) / 32)), 256, 0, s>>>((const __nv_bfloat16*)x, scale, (uint8_t*)q, rows,
                                                                                            cols, ldq);
  count_launch(3);
  return cudaGetLastError();
}

}  // namespace hb
