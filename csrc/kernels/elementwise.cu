// Memory-bound elementwise / reduction kernels (bf16 activations, 16-byte vector
// access, grid sized to a multiple of the SM count with grid-stride loops).
//
// Capability parity: hetu/impl/kernel/{Activation,Gelu,Relu,Sigmoid,Tanh,SwiGLU,
// BinaryElewise,DataTransfer,Dropout,ReduceSum}.cu.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace hb {
namespace {

constexpr int kThreads = 256;

inline int grid_for(int64_t nvec) {
  int64_t blocks = (nvec + kThreads - 1) / kThreads;
  const int64_t cap = int64_t(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

__device__ __forceinline__ float u_fwd(int op, float x) {
  switch (op) {
    case U_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
    case U_RELU: return fmaxf(x, 0.f);
    case U_SILU: return x / (1.0f + __expf(-x));
    case U_SIGMOID: return 1.0f / (1.0f + __expf(-x));
    case U_TANH: return tanhf(x);
    case U_GELU_TANH: {
      const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
      return 0.5f * x * (1.0f + tanhf(u));
    }
    case U_EXP: return __expf(x);
    case U_NEG: return -x;
    case U_SQRT: return sqrtf(x);
    case U_RSQRT: return rsqrtf(x);
    case U_ABS: return fabsf(x);
    default: return x;
  }
}
__device__ __forceinline__ float u_bwd(int op, float x) {
  switch (op) {
    case U_GELU: {
      const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
      return cdf + x * 0.39894228040143267f * __expf(-0.5f * x * x);
    }
    case U_RELU: return x > 0.f ? 1.f : 0.f;
    case U_SILU: {
      const float s = 1.0f / (1.0f + __expf(-x));
      return s * (1.0f + x * (1.0f - s));
    }
    case U_SIGMOID: {
      const float s = 1.0f / (1.0f + __expf(-x));
      return s * (1.0f - s);
    }
    case U_TANH: {
      const float t = tanhf(x);
      return 1.0f - t * t;
    }
    case U_GELU_TANH: {
      const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
      const float t = tanhf(u);
      return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * 0.7978845608028654f * (1.0f + 0.134145f * x * x);
    }
    case U_EXP: return __expf(x);
    case U_NEG: return -1.f;
    case U_SQRT: return 0.5f * rsqrtf(x);
    case U_RSQRT: return -0.5f * rsqrtf(x) / x;
    case U_ABS: return x >= 0.f ? 1.f : -1.f;
    default: return 1.f;
  }
}

__global__ void unary_fwd_kernel(int op, const void* __restrict__ x, void* __restrict__ y, int64_t n) {
  const int64_t nvec = n >> 3;
  for (int64_t v = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; v < nvec; v += int64_t(gridDim.x) * blockDim.x) {
    float f[8];
    unpack8(ld8_stream(x, v), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = u_fwd(op, f[j]);
    st8(y, v, pack8(f));
  }
  if (blockIdx.x == 0) {
    for (int64_t i = (nvec << 3) + threadIdx.x; i < n; i += blockDim.x)
      ((__nv_bfloat16*)y)[i] = __float2bfloat16(u_fwd(op, __bfloat162float(((const __nv_bfloat16*)x)[i])));
  }
}
__global__ void unary_bwd_kernel(int op, const void* __restrict__ dy, const void* __restrict__ x, void* __restrict__ dx,
                                 int64_t n) {
  const int64_t nvec = n >> 3;
  for (int64_t v = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; v < nvec; v += int64_t(gridDim.x) * blockDim.x) {
    float f[8], g[8];
    unpack8(ld8_stream(x, v), f);
    unpack8(ld8_stream(dy, v), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] *= u_bwd(op, f[j]);
    st8(dx, v, pack8(g));
  }
  if (blockIdx.x == 0) {
    for (int64_t i = (nvec << 3) + threadIdx.x; i < n; i += blockDim.x)
      ((__nv_bfloat16*)dx)[i] = __float2bfloat16(__bfloat162float(((const __nv_bfloat16*)dy)[i]) *
                                                 u_bwd(op, __bfloat162float(((const __nv_bfloat16*)x)[i])));
  }
}

__global__ void swiglu_fwd_kernel(const void* __restrict__ x, void* __restrict__ y, int64_t rows, int d) {
  const int dv = d >> 3;
  const int64_t total = rows * dv;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / dv;
    const int c = int(i - r * dv);
    float a[8], b[8];
    unpack8(ld8_stream(x, r * 2 * dv + c), a);
    unpack8(ld8_stream(x, r * 2 * dv + dv + c), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = a[j] / (1.0f + __expf(-a[j])) * b[j];
    st8(y, i, pack8(a));
  }
}
__global__ void swiglu_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ x, void* __restrict__ dx,
                                  int64_t rows, int d) {
  const int dv = d >> 3;
  const int64_t total = rows * dv;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / dv;
    const int c = int(i - r * dv);
    float a[8], b[8], g[8], da[8], dbv[8];
    unpack8(ld8_stream(x, r * 2 * dv + c), a);
    unpack8(ld8_stream(x, r * 2 * dv + dv + c), b);
    unpack8(ld8_stream(dy, i), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = 1.0f / (1.0f + __expf(-a[j]));
      da[j] = g[j] * b[j] * s * (1.0f + a[j] * (1.0f - s));
      dbv[j] = g[j] * a[j] * s;
    }
    st8(dx, r * 2 * dv + c, pack8(da));
    st8(dx, r * 2 * dv + dv + c, pack8(dbv));
  }
}

// interleaved layout: x[r, 2i] = gate_i, x[r, 2i+1] = up_i (tensor-parallel splits keep the pairs together)
__global__ void swiglu_il_fwd_kernel(const void* __restrict__ x, void* __restrict__ y, int64_t nvec_out) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < nvec_out; i += int64_t(gridDim.x) * blockDim.x) {
    float a[8], b[8], o[8];
    unpack8(ld8_stream(x, 2 * i), a);
    unpack8(ld8_stream(x, 2 * i + 1), b);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[j] = a[2 * j] / (1.0f + __expf(-a[2 * j])) * a[2 * j + 1];
      o[4 + j] = b[2 * j] / (1.0f + __expf(-b[2 * j])) * b[2 * j + 1];
    }
    st8(y, i, pack8(o));
  }
}
__global__ void swiglu_il_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ x, void* __restrict__ dx,
                                     int64_t nvec_out) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < nvec_out; i += int64_t(gridDim.x) * blockDim.x) {
    float a[16], g[8], d[16];
    unpack8(ld8_stream(x, 2 * i), a);
    unpack8(ld8_stream(x, 2 * i + 1), a + 8);
    unpack8(ld8_stream(dy, i), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gate = a[2 * j], up = a[2 * j + 1];
      const float s = 1.0f / (1.0f + __expf(-gate));
      d[2 * j] = g[j] * up * s * (1.0f + gate * (1.0f - s));
      d[2 * j + 1] = g[j] * gate * s;
    }
    st8(dx, 2 * i, pack8(d));
    st8(dx, 2 * i + 1, pack8(d + 8));
  }
}

__global__ void add_kernel(const void* __restrict__ a, const void* __restrict__ b, void* __restrict__ out, int64_t n) {
  const int64_t nvec = n >> 3;
  for (int64_t v = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; v < nvec; v += int64_t(gridDim.x) * blockDim.x) {
    float f[8], g[8];
    unpack8(ld8(a, v), f);
    unpack8(ld8_stream(b, v), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += g[j];
    st8(out, v, pack8(f));
  }
  if (blockIdx.x == 0)
    for (int64_t i = (nvec << 3) + threadIdx.x; i < n; i += blockDim.x)
      ((__nv_bfloat16*)out)[i] = __float2bfloat16(__bfloat162float(((const __nv_bfloat16*)a)[i]) +
                                                  __bfloat162float(((const __nv_bfloat16*)b)[i]));
}

__global__ void accum_kernel(const void* __restrict__ src, float* __restrict__ dst, int64_t n, int accumulate) {
  const int64_t nvec = n >> 3;
  for (int64_t v = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; v < nvec; v += int64_t(gridDim.x) * blockDim.x) {
    float f[8];
    unpack8(ld8_stream(src, v), f);
    float4* d = reinterpret_cast<float4*>(dst + v * 8);
    float4 d0 = make_float4(f[0], f[1], f[2], f[3]), d1 = make_float4(f[4], f[5], f[6], f[7]);
    if (accumulate) {
      const float4 o0 = d[0], o1 = d[1];
      d0.x += o0.x; d0.y += o0.y; d0.z += o0.z; d0.w += o0.w;
      d1.x += o1.x; d1.y += o1.y; d1.z += o1.z; d1.w += o1.w;
    }
    d[0] = d0; d[1] = d1;
  }
  if (blockIdx.x == 0)
    for (int64_t i = (nvec << 3) + threadIdx.x; i < n; i += blockDim.x) {
      const float f = __bfloat162float(((const __nv_bfloat16*)src)[i]);
      dst[i] = accumulate ? dst[i] + f : f;
    }
}

__global__ void cast_f2b_kernel(const float* __restrict__ src, void* __restrict__ dst, int64_t n) {
  const int64_t nvec = n >> 3;
  for (int64_t v = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; v < nvec; v += int64_t(gridDim.x) * blockDim.x) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(src + v * 8));
    const float4 b = __ldcs(reinterpret_cast<const float4*>(src + v * 8) + 1);
    float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    st8(dst, v, pack8(f));
  }
  if (blockIdx.x == 0)
    for (int64_t i = (nvec << 3) + threadIdx.x; i < n; i += blockDim.x) ((__nv_bfloat16*)dst)[i] = __float2bfloat16(src[i]);
}

__global__ void dropout_kernel(const void* __restrict__ x, void* __restrict__ y, int64_t n, float p, uint64_t seed,
                               uint64_t offset) {
  const int64_t nvec = n >> 3;
  const float scale = 1.0f / (1.0f - p);
  const uint32_t thresh = (uint32_t)(p * 65536.0f);
  for (int64_t v = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; v < nvec; v += int64_t(gridDim.x) * blockDim.x) {
    bool keep[8];
    dropout_keep8(uint64_t(v) + offset, seed, thresh, keep);
    float f[8];
    unpack8(ld8_stream(x, v), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = keep[j] ? f[j] * scale : 0.f;
    st8(y, v, pack8(f));
  }
}

// column sums of a bf16 [rows, cols] matrix; each block owns 64 columns x a row slice
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int64_t rows, int cols,
                              int rows_per_block) {
  __shared__ float sm[8][64];
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);
  const int ry = threadIdx.x >> 6;  // 0..3
  const int64_t r0 = int64_t(blockIdx.y) * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float s = 0.f;
  if (c < cols)
    for (int64_t r = r0 + ry; r < r1; r += 4) s += __bfloat162float(x[r * cols + c]);
  sm[ry][threadIdx.x & 63] = s;
  __syncthreads();
  if (ry == 0 && c < cols) {
    s = sm[0][threadIdx.x] + sm[1][threadIdx.x] + sm[2][threadIdx.x] + sm[3][threadIdx.x];
    atomicAdd(out + c, s);
  }
}
// Vectorised column sum: a warp covers 256 columns with 16-byte loads, the 8 warps of a CTA take interleaved rows
// (4 independent loads in flight each), partial sums meet in shared memory and one atomicAdd per column leaves the CTA.
__global__ void __launch_bounds__(256) colsum_vec_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out,
                                                         int64_t rows, int cols, int rows_per_block) {
  __shared__ float sm[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + lane * 8;
  const int64_t r0 = int64_t(blockIdx.y) * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c0 < cols) {
    const char* base = reinterpret_cast<const char*>(x + c0);
    const int64_t pitch = int64_t(cols) * 2;
    int64_t r = r0 + warp;
    for (; r + 56 < r1; r += 64) {
      bf16x8 q[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) q[u] = ld8_stream(base + (r + u * 8) * pitch, 0);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float f[8];
        unpack8(q[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    }
    for (; r < r1; r += 8) {
      float f[8];
      unpack8(ld8_stream(base + r * pitch, 0), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < cols) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sm[w][threadIdx.x];
    atomicAdd(out + c, s);
  }
}
__global__ void scale_rows_kernel(const void* __restrict__ x, void* __restrict__ y, const float* __restrict__ scale,
                                  int per_row, int64_t nvec, int64_t vec_per_row) {
  for (int64_t v = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; v < nvec; v += int64_t(gridDim.x) * blockDim.x) {
    const float sc = per_row ? scale[v / vec_per_row] : scale[0];
    float f[8];
    unpack8(ld8_stream(x, v), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= sc;
    st8(y, v, pack8(f));
  }
}
__global__ void zero_f32_kernel(float* p, int64_t n) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) p[i] = 0.f;
}

__global__ void rotary_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                              const int32_t* __restrict__ pos, int64_t tokens, int heads, int head_dim, int rot_dim,
                              float log2_base, int inverse, int64_t token_stride, int group_size, int group_stride) {
  // one thread per (token, head, pair i < rot_dim/2): (x[i], x[i + rot/2]) rotated by pos * base^(-2i/rot)
  const int half = rot_dim >> 1;
  const int64_t total = tokens * heads * half;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int i = int(idx % half);
    const int64_t th = idx / half;
    const int h = int(th % heads);
    const int64_t t = th / heads;
    const float inv_freq = exp2f(-log2_base * (2.0f * i) / rot_dim);
    const float ang = float(pos ? pos[t] : (int32_t)t) * inv_freq;
    float sn, cs;
    sincosf(ang, &sn, &cs);
    if (inverse) sn = -sn;
    // grouped layouts: rotated heads come in runs of `group_size`, runs are `group_stride` elements apart
    const int64_t base = t * token_stride + (group_size > 0 ? int64_t(h / group_size) * group_stride + int64_t(h % group_size) * head_dim
                                                            : int64_t(h) * head_dim);
    const float a = __bfloat162float(x[base + i]);
    const float b = __bfloat162float(x[base + i + half]);
    y[base + i] = __float2bfloat16(a * cs - b * sn);
    y[base + i + half] = __float2bfloat16(b * cs + a * sn);
  }
}

}  // namespace

#define CHECK_ALIGN16(p) \
  if (reinterpret_cast<uintptr_t>(p) & 15) return cudaErrorMisalignedAddress

cudaError_t unary_fwd(int op, const void* x, void* y, int64_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  CHECK_ALIGN16(x); CHECK_ALIGN16(y);
  unary_fwd_kernel<<<grid_for(n >> 3), kThreads, 0, s>>>(op, x, y, n);
  count_launch();
  return cudaGetLastError();
}
cudaError_t unary_bwd(int op, const void* dy, const void* x, void* dx, int64_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  CHECK_ALIGN16(x); CHECK_ALIGN16(dy); CHECK_ALIGN16(dx);
  unary_bwd_kernel<<<grid_for(n >> 3), kThreads, 0, s>>>(op, dy, x, dx, n);
  count_launch();
  return cudaGetLastError();
}
cudaError_t swiglu_interleaved_fwd(const void* x, void* y, int64_t rows, int d, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  CHECK_ALIGN16(x); CHECK_ALIGN16(y);
  const int64_t nv = rows * (d >> 3);
  swiglu_il_fwd_kernel<<<grid_for(nv), kThreads, 0, s>>>(x, y, nv);
  count_launch();
  return cudaGetLastError();
}
cudaError_t swiglu_interleaved_bwd(const void* dy, const void* x, void* dx, int64_t rows, int d, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  CHECK_ALIGN16(x); CHECK_ALIGN16(dy); CHECK_ALIGN16(dx);
  const int64_t nv = rows * (d >> 3);
  swiglu_il_bwd_kernel<<<grid_for(nv), kThreads, 0, s>>>(dy, x, dx, nv);
  count_launch();
  return cudaGetLastError();
}
cudaError_t swiglu_fwd(const void* x, void* y, int64_t rows, int d, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  if (d & 7) return cudaErrorInvalidValue;
  swiglu_fwd_kernel<<<grid_for(rows * (d >> 3)), kThreads, 0, s>>>(x, y, rows, d);
  count_launch();
  return cudaGetLastError();
}
cudaError_t swiglu_bwd(const void* dy, const void* x, void* dx, int64_t rows, int d, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  if (d & 7) return cudaErrorInvalidValue;
  swiglu_bwd_kernel<<<grid_for(rows * (d >> 3)), kThreads, 0, s>>>(dy, x, dx, rows, d);
  count_launch();
  return cudaGetLastError();
}
cudaError_t add_bf16(const void* a, const void* b, void* out, int64_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  CHECK_ALIGN16(a); CHECK_ALIGN16(b); CHECK_ALIGN16(out);
  add_kernel<<<grid_for(n >> 3), kThreads, 0, s>>>(a, b, out, n);
  count_launch();
  return cudaGetLastError();
}
cudaError_t accum_bf16_into_fp32(const void* src, float* dst, int64_t n, bool accumulate, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  CHECK_ALIGN16(src); CHECK_ALIGN16(dst);
  accum_kernel<<<grid_for(n >> 3), kThreads, 0, s>>>(src, dst, n, accumulate ? 1 : 0);
  count_launch();
  return cudaGetLastError();
}
cudaError_t cast_fp32_to_bf16(const float* src, void* dst, int64_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  CHECK_ALIGN16(src); CHECK_ALIGN16(dst);
  cast_f2b_kernel<<<grid_for(n >> 3), kThreads, 0, s>>>(src, dst, n);
  count_launch();
  return cudaGetLastError();
}
cudaError_t cast_bf16_to_fp32(const void* src, float* dst, int64_t n, cudaStream_t s) {
  return accum_bf16_into_fp32(src, dst, n, false, s);
}
cudaError_t dropout_fwd(const void* x, void* y, int64_t n, float p, uint64_t seed, uint64_t offset, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  if (n & 7) return cudaErrorInvalidValue;
  dropout_kernel<<<grid_for(n >> 3), kThreads, 0, s>>>(x, y, n, p, seed, offset);
  count_launch();
  return cudaGetLastError();
}
cudaError_t colsum_bf16(const void* x, float* out, int64_t rows, int cols, bool accumulate, cudaStream_t s) {
  if (!accumulate) {
    zero_f32_kernel<<<(cols + 255) / 256, 256, 0, s>>>(out, cols);
    count_launch();
  }
  if (rows == 0) return cudaGetLastError();
  if ((cols & 7) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int cb = (cols + 255) / 256;
    int rb = (sm_count() * 4 + cb - 1) / cb;
    if (rb > (rows + 63) / 64) rb = (int)((rows + 63) / 64);
    if (rb < 1) rb = 1;
    const int rpb = (int)((rows + rb - 1) / rb);
    rb = (int)((rows + rpb - 1) / rpb);
    colsum_vec_kernel<<<dim3(cb, rb), 256, 0, s>>>((const __nv_bfloat16*)x, out, rows, cols, rpb);
    count_launch();
    return cudaGetLastError();
  }
  const int col_blocks = (cols + 63) / 64;
  int row_blocks = (sm_count() * 4 + col_blocks - 1) / col_blocks;
  if (row_blocks > rows) row_blocks = (int)rows;
  const int rpb = (int)((rows + row_blocks - 1) / row_blocks);
  row_blocks = (int)((rows + rpb - 1) / rpb);
  colsum_kernel<<<dim3(col_blocks, row_blocks), 256, 0, s>>>((const __nv_bfloat16*)x, out, rows, cols, rpb);
  count_launch();
  return cudaGetLastError();
}
cudaError_t scale_rows_bf16(const void* x, void* y, const float* scale, bool per_row, int64_t rows, int64_t cols,
                            cudaStream_t s) {
  if (rows * cols == 0) return cudaSuccess;
  if ((cols & 7) || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) return cudaErrorMisalignedAddress;
  const int64_t nvec = rows * cols / 8;
  scale_rows_kernel<<<grid_for(nvec), kThreads, 0, s>>>(x, y, scale, per_row ? 1 : 0, nvec, cols / 8);
  count_launch();
  return cudaGetLastError();
}
cudaError_t rotary_apply(const void* x, void* y, const int32_t* pos, int64_t tokens, int heads, int head_dim,
                         int rot_dim, float base, bool inverse, int64_t token_stride, cudaStream_t s, int group_size,
                         int group_stride) {
  if (tokens == 0) return cudaSuccess;
  const int64_t total = tokens * heads * (rot_dim / 2);
  rotary_kernel<<<grid_for(total), kThreads, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, pos, tokens, heads,
                                                     head_dim, rot_dim, log2f(base), inverse ? 1 : 0, token_stride, group_size,
                                                     group_stride);
  count_launch();
  return cudaGetLastError();
}

}  // namespace hb
