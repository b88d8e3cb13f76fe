// Scaled (masked / causal) row softmax forward + backward for bf16 attention scores, for graphs that materialise the
// score matrix instead of using the fused flash-attention kernels.
// (capability parity: tools/Galvatron/galvatron/site_package/megatron/fused_kernels/scaled_masked_softmax.h,
//  scaled_upper_triang_masked_softmax.h; hetu/impl/kernel/Softmax.cu)
#include "common.cuh"
#include "kernels.h"

namespace hb {
namespace {

// mode 0: plain, 1: boolean mask (1 = masked out) broadcast over heads: mask[b, 1, sq, sk], 2: causal (col > row masked)
__global__ void __launch_bounds__(128) scaled_softmax_fwd_kernel(const __nv_bfloat16* __restrict__ x, const uint8_t* __restrict__ mask,
                                                                 __nv_bfloat16* __restrict__ y, int cols, int rows_per_batch, int sq,
                                                                 float scale, int mode) {
  __shared__ float red[32];
  const int64_t r = blockIdx.x;
  const __nv_bfloat16* xr = x + r * cols;
  __nv_bfloat16* yr = y + r * cols;
  const int q = int(r % sq);
  const uint8_t* mr = mode == 1 ? mask + (r / rows_per_batch) * int64_t(sq) * cols + int64_t(q) * cols : nullptr;
  const int limit = mode == 2 ? q + 1 + (cols - sq) : cols;      // bottom-right aligned causal window
  float m = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const bool dead = c >= limit || (mr && mr[c]);
    if (!dead) m = fmaxf(m, __bfloat162float(xr[c]) * scale);
  }
  m = block_max(m, red);
  float s = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const bool dead = c >= limit || (mr && mr[c]);
    if (!dead) s += __expf(__bfloat162float(xr[c]) * scale - m);
  }
  s = block_sum(s, red);
  const float inv = s > 0.f ? 1.0f / s : 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const bool dead = c >= limit || (mr && mr[c]);
    yr[c] = __float2bfloat16(dead ? 0.f : __expf(__bfloat162float(xr[c]) * scale - m) * inv);
  }
}
// dx = scale * y * (dy - sum(dy * y))
__global__ void __launch_bounds__(128) scaled_softmax_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                                                 __nv_bfloat16* __restrict__ dx, int cols, float scale) {
  __shared__ float red[32];
  const int64_t r = blockIdx.x;
  float dot = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) dot += __bfloat162float(dy[r * cols + c]) * __bfloat162float(y[r * cols + c]);
  dot = block_sum(dot, red);
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float yy = __bfloat162float(y[r * cols + c]);
    dx[r * cols + c] = __float2bfloat16(scale * yy * (__bfloat162float(dy[r * cols + c]) - dot));
  }
}
}  // namespace

cudaError_t scaled_softmax_fwd(const void* x, const uint8_t* mask, void* y, int64_t rows, int cols, int rows_per_batch, int sq, float scale,
                               int mode, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  scaled_softmax_fwd_kernel<<<(unsigned)rows, 128, 0, s>>>((const __nv_bfloat16*)x, mask, (__nv_bfloat16*)y, cols, rows_per_batch, sq, scale,
                                                           mode);
  count_launch();
  return cudaGetLastError();
}
cudaError_t scaled_softmax_bwd(const void* dy, const void* y, void* dx, int64_t rows, int cols, float scale, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  scaled_softmax_bwd_kernel<<<(unsigned)rows, 128, 0, s>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, (__nv_bfloat16*)dx, cols, scale);
  count_launch();
  return cudaGetLastError();
}

}  // namespace hb
