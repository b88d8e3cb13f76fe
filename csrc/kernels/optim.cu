// Fused optimizers over flat fp32 shards (ZeRO-friendly): one launch updates the
// fp32 master weights, both Adam moments and the bf16 compute copy, optionally
// zeroing the gradient accumulator in the same pass.  The step counter, grad
// scale and learning rate are read from device memory so a whole training step
// can be captured in a CUDA graph.
//
// Capability parity: hetu/impl/kernel/Optimizers.cu:13-188 (SGDUpdate,
// SGDUpdateWithGradScaler, AdamCuda), CheckFinite.cu, hetu/graph/optim/optimizer.cc.
#include "common.cuh"
#include "kernels.h"

namespace hb {
namespace {

struct AdamDev {
  float* master; float* m; float* v;
  const void* grad; int grad_is_bf16;
  void* param_bf16;
  int64_t n;
  float lr, beta1, beta2, eps, wd;
  const int64_t* step_ptr; int step_add;
  const float* grad_scale_ptr;
  const float* lr_ptr;
  int zero_grad;
};

__global__ void __launch_bounds__(256) adam_kernel(AdamDev a) {
  const float step = a.step_ptr ? float(*a.step_ptr + a.step_add) : 1.0f;
  const float gscale = a.grad_scale_ptr ? *a.grad_scale_ptr : 1.0f;
  const float lr = a.lr_ptr ? *a.lr_ptr : a.lr;
  const float bc1 = 1.0f - __powf(a.beta1, step);
  const float bc2 = 1.0f - __powf(a.beta2, step);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const int64_t nvec = a.n >> 2;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < nvec; i += int64_t(gridDim.x) * blockDim.x) {
    float g[4];
    if (a.grad_is_bf16) {
      const uint2 raw = reinterpret_cast<const uint2*>(a.grad)[i];
      const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&raw);
      const float2 f0 = __bfloat1622float2(b[0]), f1 = __bfloat1622float2(b[1]);
      g[0] = f0.x; g[1] = f0.y; g[2] = f1.x; g[3] = f1.y;
      if (a.zero_grad) reinterpret_cast<uint2*>(const_cast<void*>(a.grad))[i] = make_uint2(0u, 0u);
    } else {
      const float4 f = reinterpret_cast<const float4*>(a.grad)[i];
      g[0] = f.x; g[1] = f.y; g[2] = f.z; g[3] = f.w;
      if (a.zero_grad) reinterpret_cast<float4*>(const_cast<void*>(a.grad))[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 p4 = reinterpret_cast<float4*>(a.master)[i];
    float4 m4 = reinterpret_cast<float4*>(a.m)[i];
    float4 v4 = reinterpret_cast<float4*>(a.v)[i];
    float p[4] = {p4.x, p4.y, p4.z, p4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = g[j] * gscale;
      m[j] = a.beta1 * m[j] + (1.0f - a.beta1) * gj;
      v[j] = a.beta2 * v[j] + (1.0f - a.beta2) * gj * gj;
      const float denom = sqrtf(v[j]) * inv_sqrt_bc2 + a.eps;
      p[j] = p[j] - step_size * (m[j] / denom) - lr * a.wd * p[j];
    }
    reinterpret_cast<float4*>(a.master)[i] = make_float4(p[0], p[1], p[2], p[3]);
    reinterpret_cast<float4*>(a.m)[i] = make_float4(m[0], m[1], m[2], m[3]);
    reinterpret_cast<float4*>(a.v)[i] = make_float4(v[0], v[1], v[2], v[3]);
    if (a.param_bf16) {
      uint2 o;
      __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
      ob[0] = __floats2bfloat162_rn(p[0], p[1]);
      ob[1] = __floats2bfloat162_rn(p[2], p[3]);
      reinterpret_cast<uint2*>(a.param_bf16)[i] = o;
    }
  }
  // scalar tail
  if (blockIdx.x == 0) {
    for (int64_t i = (nvec << 2) + threadIdx.x; i < a.n; i += blockDim.x) {
      float gj = a.grad_is_bf16 ? __bfloat162float(((const __nv_bfloat16*)a.grad)[i]) : ((const float*)a.grad)[i];
      if (a.zero_grad) {
        if (a.grad_is_bf16) ((__nv_bfloat16*)const_cast<void*>(a.grad))[i] = __float2bfloat16(0.f);
        else ((float*)const_cast<void*>(a.grad))[i] = 0.f;
      }
      gj *= gscale;
      const float mm = a.beta1 * a.m[i] + (1.0f - a.beta1) * gj;
      const float vv = a.beta2 * a.v[i] + (1.0f - a.beta2) * gj * gj;
      float pp = a.master[i];
      pp = pp - step_size * (mm / (sqrtf(vv) * inv_sqrt_bc2 + a.eps)) - lr * a.wd * pp;
      a.m[i] = mm; a.v[i] = vv; a.master[i] = pp;
      if (a.param_bf16) ((__nv_bfloat16*)a.param_bf16)[i] = __float2bfloat16(pp);
    }
  }
}

__global__ void sgd_kernel(float* master, float* mom, const void* grad, int grad_is_bf16, void* param_bf16, int64_t n,
                           float lr, float momentum, int nesterov, float wd) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    float g = grad_is_bf16 ? __bfloat162float(((const __nv_bfloat16*)grad)[i]) : ((const float*)grad)[i];
    float p = master[i];
    g += wd * p;
    if (mom != nullptr && momentum != 0.f) {
      const float b = momentum * mom[i] + g;
      mom[i] = b;
      g = nesterov ? g + momentum * b : b;
    }
    p -= lr * g;
    master[i] = p;
    if (param_bf16) ((__nv_bfloat16*)param_bf16)[i] = __float2bfloat16(p);
  }
}

__global__ void inc_step_kernel(int64_t* s) { *s += 1; }

__global__ void check_finite_kernel(const float* x, int64_t n, float* found) {
  bool bad = false;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    bad |= !isfinite(x[i]);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *found = 1.0f;
}

__global__ void sumsq_kernel(const float* x, int64_t n, float* out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    s += x[i] * x[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
}
__global__ void zero1_kernel(float* p) { *p = 0.f; }

inline int grid_for(int64_t n) {
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = int64_t(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

cudaError_t adam_update(const AdamArgs& a, cudaStream_t s) {
  if (a.n == 0) return cudaSuccess;
  if ((reinterpret_cast<uintptr_t>(a.master) | reinterpret_cast<uintptr_t>(a.m) | reinterpret_cast<uintptr_t>(a.v) |
       reinterpret_cast<uintptr_t>(a.grad)) & 15)
    return cudaErrorMisalignedAddress;
  if (a.param_bf16 && (reinterpret_cast<uintptr_t>(a.param_bf16) & 7)) return cudaErrorMisalignedAddress;
  AdamDev d;
  d.master = a.master; d.m = a.m; d.v = a.v; d.grad = a.grad; d.grad_is_bf16 = a.grad_is_bf16 ? 1 : 0;
  d.param_bf16 = a.param_bf16; d.n = a.n; d.lr = a.lr; d.beta1 = a.beta1; d.beta2 = a.beta2; d.eps = a.eps;
  d.wd = a.weight_decay; d.step_ptr = a.step_ptr; d.step_add = a.step_add; d.grad_scale_ptr = a.grad_scale_ptr; d.lr_ptr = a.lr_ptr;
  d.zero_grad = a.zero_grad ? 1 : 0;
  adam_kernel<<<grid_for(a.n >> 2), 256, 0, s>>>(d);
  count_launch();
  return cudaGetLastError();
}
namespace {
struct AdamZeroDev {
  float* master; float* m; float* v;
  const __nv_bfloat16* slots; int nslots; int64_t slot_stride;
  __nv_bfloat16* peer[8]; int world;
  int64_t n;
  float lr, beta1, beta2, eps, wd, gscale;
  const int64_t* step_ptr; int step_add;
};
__global__ void __launch_bounds__(256) adam_zero_fused_kernel(AdamZeroDev a) {
  const float step = float((a.step_ptr ? *a.step_ptr : 0) + a.step_add);
  const float bc1 = 1.0f - __powf(a.beta1, step);
  const float bc2 = 1.0f - __powf(a.beta2, step);
  const float step_size = a.lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const int64_t nvec = a.n >> 3;   // 8 elements (16 bytes of bf16) per thread-iteration
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < nvec; i += int64_t(gridDim.x) * blockDim.x) {
    float g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = 0.f;
    for (int s = 0; s < a.nslots; ++s) {
      const uint4 raw = __ldcs(reinterpret_cast<const uint4*>(a.slots + int64_t(s) * a.slot_stride) + i);
      const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(b[j]);
        g[2 * j] += f.x; g[2 * j + 1] += f.y;
      }
    }
    float p[8], m[8], v[8];
    *reinterpret_cast<float4*>(p) = reinterpret_cast<const float4*>(a.master)[2 * i];
    *reinterpret_cast<float4*>(p + 4) = reinterpret_cast<const float4*>(a.master)[2 * i + 1];
    *reinterpret_cast<float4*>(m) = reinterpret_cast<const float4*>(a.m)[2 * i];
    *reinterpret_cast<float4*>(m + 4) = reinterpret_cast<const float4*>(a.m)[2 * i + 1];
    *reinterpret_cast<float4*>(v) = reinterpret_cast<const float4*>(a.v)[2 * i];
    *reinterpret_cast<float4*>(v + 4) = reinterpret_cast<const float4*>(a.v)[2 * i + 1];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gj = g[j] * a.gscale;
      m[j] = a.beta1 * m[j] + (1.0f - a.beta1) * gj;
      v[j] = a.beta2 * v[j] + (1.0f - a.beta2) * gj * gj;
      const float denom = sqrtf(v[j]) * inv_sqrt_bc2 + a.eps;
      p[j] = p[j] - step_size * (m[j] / denom) - a.lr * a.wd * p[j];
    }
    reinterpret_cast<float4*>(a.master)[2 * i] = *reinterpret_cast<const float4*>(p);
    reinterpret_cast<float4*>(a.master)[2 * i + 1] = *reinterpret_cast<const float4*>(p + 4);
    reinterpret_cast<float4*>(a.m)[2 * i] = *reinterpret_cast<const float4*>(m);
    reinterpret_cast<float4*>(a.m)[2 * i + 1] = *reinterpret_cast<const float4*>(m + 4);
    reinterpret_cast<float4*>(a.v)[2 * i] = *reinterpret_cast<const float4*>(v);
    reinterpret_cast<float4*>(a.v)[2 * i + 1] = *reinterpret_cast<const float4*>(v + 4);
    uint4 o;
    __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) ob[j] = __floats2bfloat162_rn(p[2 * j], p[2 * j + 1]);
    for (int r = 0; r < a.world; ++r) reinterpret_cast<uint4*>(a.peer[r])[i] = o;   // all-gather by peer stores
  }
  if (blockIdx.x == 0) {   // scalar tail (n not a multiple of 8)
    for (int64_t i = (nvec << 3) + threadIdx.x; i < a.n; i += blockDim.x) {
      float g = 0.f;
      for (int s = 0; s < a.nslots; ++s) g += __bfloat162float(a.slots[int64_t(s) * a.slot_stride + i]);
      g *= a.gscale;
      const float m = a.beta1 * a.m[i] + (1.0f - a.beta1) * g;
      const float v = a.beta2 * a.v[i] + (1.0f - a.beta2) * g * g;
      float p = a.master[i];
      p = p - step_size * (m / (sqrtf(v) * inv_sqrt_bc2 + a.eps)) - a.lr * a.wd * p;
      a.master[i] = p; a.m[i] = m; a.v[i] = v;
      for (int r = 0; r < a.world; ++r) a.peer[r][i] = __float2bfloat16(p);
    }
  }
}
// ---- ZeRO step over NVLS: the NVSwitch sums the ranks' gradient copies while this rank loads its shard
// (multimem.ld_reduce, fp32 accumulation in the switch), AdamW runs on the fp32 master shard, and one multicast store
// (multimem.st) writes the new bf16 shard into EVERY rank's parameter tensor: reduce-scatter + optimizer + all-gather in
// one pass with no staging slots and no per-peer store loop.
struct AdamNvlsDev {
  float* master; float* m; float* v;
  const char* grad_mc;   // multicast address of this rank's shard inside the (symmetric) gradient region
  char* param_mc;        // multicast address of this rank's shard inside the (symmetric) parameter tensor
  int64_t n;
  float lr, beta1, beta2, eps, wd, gscale;
  const int64_t* step_ptr; int step_add;
};
__global__ void __launch_bounds__(256) adam_nvls_kernel(AdamNvlsDev a) {
  const float step = float((a.step_ptr ? *a.step_ptr : 0) + a.step_add);
  const float bc1 = 1.0f - __powf(a.beta1, step);
  const float bc2 = 1.0f - __powf(a.beta2, step);
  const float step_size = a.lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const int64_t nvec = a.n >> 3;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < nvec; i += int64_t(gridDim.x) * blockDim.x) {
    uint4 raw;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "l"(a.grad_mc + i * 16) : "memory");
    const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&raw);
    float g[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(b[j]);
      g[2 * j] = f.x; g[2 * j + 1] = f.y;
    }
    float p[8], m[8], v[8];
    *reinterpret_cast<float4*>(p) = reinterpret_cast<const float4*>(a.master)[2 * i];
    *reinterpret_cast<float4*>(p + 4) = reinterpret_cast<const float4*>(a.master)[2 * i + 1];
    *reinterpret_cast<float4*>(m) = reinterpret_cast<const float4*>(a.m)[2 * i];
    *reinterpret_cast<float4*>(m + 4) = reinterpret_cast<const float4*>(a.m)[2 * i + 1];
    *reinterpret_cast<float4*>(v) = reinterpret_cast<const float4*>(a.v)[2 * i];
    *reinterpret_cast<float4*>(v + 4) = reinterpret_cast<const float4*>(a.v)[2 * i + 1];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gj = g[j] * a.gscale;
      m[j] = a.beta1 * m[j] + (1.0f - a.beta1) * gj;
      v[j] = a.beta2 * v[j] + (1.0f - a.beta2) * gj * gj;
      const float denom = sqrtf(v[j]) * inv_sqrt_bc2 + a.eps;
      p[j] = p[j] - step_size * (m[j] / denom) - a.lr * a.wd * p[j];
    }
    reinterpret_cast<float4*>(a.master)[2 * i] = *reinterpret_cast<const float4*>(p);
    reinterpret_cast<float4*>(a.master)[2 * i + 1] = *reinterpret_cast<const float4*>(p + 4);
    reinterpret_cast<float4*>(a.m)[2 * i] = *reinterpret_cast<const float4*>(m);
    reinterpret_cast<float4*>(a.m)[2 * i + 1] = *reinterpret_cast<const float4*>(m + 4);
    reinterpret_cast<float4*>(a.v)[2 * i] = *reinterpret_cast<const float4*>(v);
    reinterpret_cast<float4*>(a.v)[2 * i + 1] = *reinterpret_cast<const float4*>(v + 4);
    uint4 o;
    __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) ob[j] = __floats2bfloat162_rn(p[2 * j], p[2 * j + 1]);
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(a.param_mc + i * 16), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
  }
}

__global__ void increment_many_kernel(int64_t* const* table, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) *table[i] += 1;
}
}  // namespace

cudaError_t adam_zero_fused(const AdamZeroArgs& a, cudaStream_t s) {
  if (a.n == 0) return cudaSuccess;
  if (a.world > 8 || a.nslots < 1) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(a.master) | reinterpret_cast<uintptr_t>(a.m) | reinterpret_cast<uintptr_t>(a.v) |
       reinterpret_cast<uintptr_t>(a.slots)) & 15)
    return cudaErrorMisalignedAddress;
  if ((a.slot_stride & 7) != 0 && a.nslots > 1) return cudaErrorMisalignedAddress;
  AdamZeroDev d;
  d.master = a.master; d.m = a.m; d.v = a.v;
  d.slots = reinterpret_cast<const __nv_bfloat16*>(a.slots); d.nslots = a.nslots; d.slot_stride = a.slot_stride;
  for (int r = 0; r < 8; ++r) {
    d.peer[r] = reinterpret_cast<__nv_bfloat16*>(r < a.world ? a.peer_param[r] : nullptr);
    if (r < a.world && (reinterpret_cast<uintptr_t>(a.peer_param[r]) & 15)) return cudaErrorMisalignedAddress;
  }
  d.world = a.world; d.n = a.n;
  d.lr = a.lr; d.beta1 = a.beta1; d.beta2 = a.beta2; d.eps = a.eps; d.wd = a.weight_decay; d.gscale = a.grad_scale;
  d.step_ptr = a.step_ptr; d.step_add = a.step_add;
  adam_zero_fused_kernel<<<grid_for(a.n >> 3), 256, 0, s>>>(d);
  count_launch();
  return cudaGetLastError();
}
cudaError_t adam_zero_nvls(const AdamNvlsArgs& a, cudaStream_t s) {
  if (a.n == 0) return cudaSuccess;
  if (a.n & 7) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(a.master) | reinterpret_cast<uintptr_t>(a.m) | reinterpret_cast<uintptr_t>(a.v) |
       reinterpret_cast<uintptr_t>(a.grad_mc) | reinterpret_cast<uintptr_t>(a.param_mc)) & 15)
    return cudaErrorMisalignedAddress;
  AdamNvlsDev d;
  d.master = a.master; d.m = a.m; d.v = a.v;
  d.grad_mc = reinterpret_cast<const char*>(a.grad_mc); d.param_mc = reinterpret_cast<char*>(a.param_mc);
  d.n = a.n;
  d.lr = a.lr; d.beta1 = a.beta1; d.beta2 = a.beta2; d.eps = a.eps; d.wd = a.weight_decay; d.gscale = a.grad_scale;
  d.step_ptr = a.step_ptr; d.step_add = a.step_add;
  adam_nvls_kernel<<<grid_for(a.n >> 3), 256, 0, s>>>(d);
  count_launch();
  return cudaGetLastError();
}
cudaError_t increment_many_i64(int64_t* const* table, int count, cudaStream_t s) {
  if (count <= 0) return cudaSuccess;
  increment_many_kernel<<<(count + 255) / 256, 256, 0, s>>>(table, count);
  count_launch();
  return cudaGetLastError();
}
cudaError_t sgd_update(float* master, float* momentum_buf, const void* grad, bool grad_is_bf16, void* param_bf16,
                       int64_t n, float lr, float momentum, bool nesterov, float weight_decay, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  sgd_kernel<<<grid_for(n), 256, 0, s>>>(master, momentum_buf, grad, grad_is_bf16 ? 1 : 0, param_bf16, n, lr, momentum,
                                         nesterov ? 1 : 0, weight_decay);
  count_launch();
  return cudaGetLastError();
}
cudaError_t increment_step(int64_t* step_ptr, cudaStream_t s) {
  inc_step_kernel<<<1, 1, 0, s>>>(step_ptr);
  count_launch();
  return cudaGetLastError();
}
cudaError_t check_finite(const float* x, int64_t n, float* found_inf, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  check_finite_kernel<<<grid_for(n), 256, 0, s>>>(x, n, found_inf);
  count_launch();
  return cudaGetLastError();
}
cudaError_t sumsq_fp32(const float* x, int64_t n, float* out, bool accumulate, cudaStream_t s) {
  if (!accumulate) { zero1_kernel<<<1, 1, 0, s>>>(out); count_launch(); }
  if (n == 0) return cudaGetLastError();
  sumsq_kernel<<<grid_for(n), 256, 0, s>>>(x, n, out);
  count_launch();
  return cudaGetLastError();
}

}  // namespace hb
