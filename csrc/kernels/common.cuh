// Shared device helpers for the memory-bound kernels: 16-byte vector access,
// bf16 pack/unpack, warp/block reductions, launch accounting.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace hb {

extern std::atomic<int64_t> g_kernel_launches;
inline void count_launch(int n = 1) { g_kernel_launches.fetch_add(n); }

struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};

__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return p;
}
__device__ __forceinline__ bf16x8 ld8(const void* base, int64_t vec_idx) {
  return reinterpret_cast<const bf16x8*>(base)[vec_idx];
}
__device__ __forceinline__ void st8(void* base, int64_t vec_idx, const bf16x8& v) {
  reinterpret_cast<bf16x8*>(base)[vec_idx] = v;
}
// streaming variants (do not pollute L1)
__device__ __forceinline__ bf16x8 ld8_stream(const void* base, int64_t vec_idx) {
  const uint4 u = __ldcs(reinterpret_cast<const uint4*>(base) + vec_idx);
  bf16x8 r;
  *reinterpret_cast<uint4*>(&r) = u;
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Block-wide sum for blocks of up to 1024 threads; `red` is >= 32 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (lane < nw) ? red[lane] : 0.0f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (lane < nw) ? red[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// counter-based RNG: 4 uniform floats per (seed, counter)
__device__ __forceinline__ uint32_t mulhilo32(uint32_t a, uint32_t b, uint32_t* hi) {
  *hi = __umulhi(a, b);
  return a * b;
}
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, hi1;
    const uint32_t lo0 = mulhilo32(M0, ctr.x, &hi0);
    const uint32_t lo1 = mulhilo32(M1, ctr.z, &hi1);
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// keep / drop decision of dropout for the 8 elements of vector `v` (16 bits of Philox output per element): shared by the
// standalone dropout kernel and the fused dropout + residual + norm kernel so their masks agree
__device__ __forceinline__ void dropout_keep8(uint64_t counter, uint64_t seed, uint32_t thresh, bool keep[8]) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) keep[j] = ((rr[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu) >= thresh;
}

}  // namespace hb
