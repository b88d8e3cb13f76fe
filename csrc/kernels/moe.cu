// HetuMoE device kernels: softmax top-k gating, capacity-based slot assignment
// (deterministic, token-order like the reference's cumsum formulation), token
// dispatch into expert-major buffers and gated combine.
//
// Capability parity: hetu/v1/src/ops/{TopKIdx,TopKVal,CumSum,Scatter1D,
// LayoutTransform}.cu and hetu/v1/python/hetu/layers/TopGate.py:14-57.
#include "common.cuh"
#include "kernels.h"

namespace hb {
namespace {

constexpr int kMaxExpertsPerLane = 8;  // up to 256 experts
constexpr int kMaxK = 8;

__global__ void gate_topk_kernel(const __nv_bfloat16* __restrict__ logits, float* __restrict__ probs,
                                 int32_t* __restrict__ topk_idx, float* __restrict__ topk_val, int64_t tokens,
                                 int experts, int k) {
  const int lane = threadIdx.x & 31;
  const int64_t t = blockIdx.x * int64_t(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= tokens) return;
  float v[kMaxExpertsPerLane];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < kMaxExpertsPerLane; ++i) {
    const int e = lane + i * 32;
    v[i] = e < experts ? __bfloat162float(logits[t * experts + e]) : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxExpertsPerLane; ++i) {
    v[i] = (lane + i * 32 < experts) ? __expf(v[i] - mx) : 0.f;
    sum += v[i];
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < kMaxExpertsPerLane; ++i) {
    v[i] *= inv;
    const int e = lane + i * 32;
    if (e < experts && probs) probs[t * experts + e] = v[i];
  }
  for (int kk = 0; kk < k; ++kk) {
    float best = -1.f; int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < kMaxExpertsPerLane; ++i) {
      const int e = lane + i * 32;
      if (e < experts && v[i] > best) { best = v[i]; bi = e; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { topk_idx[t * k + kk] = bi; topk_val[t * k + kk] = best; }
#pragma unroll
    for (int i = 0; i < kMaxExpertsPerLane; ++i)
      if (lane + i * 32 == bi) v[i] = -2.f;
  }
}

// One warp per expert; walks choices in (k-major, token-minor) order so the first choices of all
// tokens are placed before any second choice (GShard ordering).
__global__ void assign_slots_kernel(const int32_t* __restrict__ topk_idx, int32_t* __restrict__ location,
                                    int32_t* __restrict__ expert_count, int64_t tokens, int experts, int k,
                                    int capacity) {
  const int lane = threadIdx.x & 31;
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (e >= experts) return;
  int count = 0;
  for (int kk = 0; kk < k; ++kk) {
    for (int64_t t0 = 0; t0 < tokens; t0 += 32) {
      const int64_t t = t0 + lane;
      const bool mine = t < tokens && topk_idx[t * k + kk] == e;
      const unsigned ball = __ballot_sync(0xffffffffu, mine);
      if (mine) {
        const int loc = count + __popc(ball & ((1u << lane) - 1));
        location[t * k + kk] = loc < capacity ? loc : -1;
      }
      count += __popc(ball);
    }
  }
  if (lane == 0 && expert_count) expert_count[e] = count < capacity ? count : capacity;
}

__global__ void dispatch_kernel(const void* __restrict__ x, const int32_t* __restrict__ topk_idx,
                                const int32_t* __restrict__ location, const float* __restrict__ scale,
                                void* __restrict__ dispatched, int64_t tokens, int hidden, int k, int capacity) {
  const int hv = hidden >> 3;
  const int64_t total = tokens * k * hv;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % hv);
    const int64_t tk = i / hv;
    const int loc = location[tk];
    if (loc < 0) continue;
    const int64_t t = tk / k;
    const int e = topk_idx[tk];
    bf16x8 val = ld8(x, t * hv + c);
    if (scale) {
      float f[8];
      unpack8(val, f);
      const float sc = scale[tk];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= sc;
      val = pack8(f);
    }
    st8(dispatched, (int64_t(e) * capacity + loc) * hv + c, val);
  }
}

__global__ void combine_kernel(const void* __restrict__ expert_out, const int32_t* __restrict__ topk_idx,
                               const int32_t* __restrict__ location, const float* __restrict__ gate,
                               void* __restrict__ y, int64_t tokens, int hidden, int k, int capacity) {
  const int hv = hidden >> 3;
  const int64_t total = tokens * hv;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % hv);
    const int64_t t = i / hv;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int kk = 0; kk < k; ++kk) {
      const int loc = location[t * k + kk];
      if (loc < 0) continue;
      const int e = topk_idx[t * k + kk];
      const float gt = gate ? gate[t * k + kk] : 1.0f;
      float f[8];
      unpack8(ld8(expert_out, (int64_t(e) * capacity + loc) * hv + c), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += gt * f[j];
    }
    st8(y, i, pack8(acc));
  }
}

__global__ void combine_bwd_gate_kernel(const void* __restrict__ dy, const void* __restrict__ expert_out,
                                        const int32_t* __restrict__ topk_idx, const int32_t* __restrict__ location,
                                        float* __restrict__ dgate, int64_t tokens, int hidden, int k, int capacity) {
  const int lane = threadIdx.x & 31;
  const int64_t tk = blockIdx.x * int64_t(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tk >= tokens * k) return;
  const int loc = location[tk];
  float s = 0.f;
  if (loc >= 0) {
    const int64_t t = tk / k;
    const int e = topk_idx[tk];
    const int hv = hidden >> 3;
    for (int c = lane; c < hv; c += 32) {
      float a[8], b[8];
      unpack8(ld8(dy, t * hv + c), a);
      unpack8(ld8(expert_out, (int64_t(e) * capacity + loc) * hv + c), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += a[j] * b[j];
    }
  }
  s = warp_sum(s);
  if (lane == 0) dgate[tk] = s;
}

__device__ __forceinline__ char* peer_row(const MoePeers& pr, int e, int loc, int capacity, int hidden) {
  const int r = e / pr.experts_per_rank, el = e - r * pr.experts_per_rank;
  return static_cast<char*>(pr.base[r]) + ((int64_t(el) * pr.ep + pr.src_rank) * capacity + loc) * int64_t(hidden) * 2;
}
__global__ void dispatch_peers_kernel(const void* __restrict__ x, const int32_t* __restrict__ topk_idx,
                                      const int32_t* __restrict__ location, const float* __restrict__ scale, MoePeers pr,
                                      int64_t tokens, int hidden, int k, int capacity) {
  const int hv = hidden >> 3;
  const int64_t total = tokens * k * hv;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % hv);
    const int64_t tk = i / hv;
    const int loc = location[tk];
    if (loc < 0) continue;
    const int64_t t = tk / k;
    bf16x8 val = ld8(x, t * hv + c);
    if (scale) {
      float f[8];
      unpack8(val, f);
      const float sc = scale[tk];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= sc;
      val = pack8(f);
    }
    st8(peer_row(pr, topk_idx[tk], loc, capacity, hidden), c, val);     // store over NVLink into the expert's rank
  }
}
__global__ void combine_peers_kernel(MoePeers pr, const int32_t* __restrict__ topk_idx, const int32_t* __restrict__ location,
                                     const float* __restrict__ gate, void* __restrict__ y, int64_t tokens, int hidden, int k,
                                     int capacity) {
  const int hv = hidden >> 3;
  const int64_t total = tokens * hv;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % hv);
    const int64_t t = i / hv;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int kk = 0; kk < k; ++kk) {
      const int loc = location[t * k + kk];
      if (loc < 0) continue;
      const float gt = gate ? gate[t * k + kk] : 1.0f;
      float f[8];
      unpack8(ld8_stream(peer_row(pr, topk_idx[t * k + kk], loc, capacity, hidden), c), f);   // load over NVLink
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += gt * f[j];
    }
    st8(y, i, pack8(acc));
  }
}
__global__ void combine_bwd_gate_peers_kernel(const void* __restrict__ dy, MoePeers pr, const int32_t* __restrict__ topk_idx,
                                              const int32_t* __restrict__ location, float* __restrict__ dgate, int64_t tokens,
                                              int hidden, int k, int capacity) {
  const int lane = threadIdx.x & 31;
  const int64_t tk = blockIdx.x * int64_t(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tk >= tokens * k) return;
  const int loc = location[tk];
  float s = 0.f;
  if (loc >= 0) {
    const int64_t t = tk / k;
    const int hv = hidden >> 3;
    const char* row = peer_row(pr, topk_idx[tk], loc, capacity, hidden);
    for (int c = lane; c < hv; c += 32) {
      float a[8], b[8];
      unpack8(ld8(dy, t * hv + c), a);
      unpack8(ld8_stream(row, c), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += a[j] * b[j];
    }
  }
  s = warp_sum(s);
  if (lane == 0) dgate[tk] = s;
}

inline int grid_for(int64_t n) {
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = int64_t(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

cudaError_t moe_gate_topk(const void* logits_bf16, float* probs, int32_t* topk_idx, float* topk_val, int64_t tokens,
                          int experts, int k, cudaStream_t s) {
  if (tokens == 0) return cudaSuccess;
  if (experts > 32 * kMaxExpertsPerLane || k > kMaxK || k > experts) return cudaErrorInvalidValue;
  gate_topk_kernel<<<(unsigned)((tokens + 7) / 8), 256, 0, s>>>((const __nv_bfloat16*)logits_bf16, probs, topk_idx,
                                                                topk_val, tokens, experts, k);
  count_launch();
  return cudaGetLastError();
}
cudaError_t moe_assign_slots(const int32_t* topk_idx, int32_t* location, int32_t* expert_count, int64_t tokens,
                             int experts, int k, int capacity, cudaStream_t s) {
  assign_slots_kernel<<<(experts + 3) / 4, 128, 0, s>>>(topk_idx, location, expert_count, tokens, experts, k, capacity);
  count_launch();
  return cudaGetLastError();
}
cudaError_t moe_dispatch(const void* x, const int32_t* topk_idx, const int32_t* location, const float* scale,
                         void* dispatched, int64_t tokens, int hidden, int experts, int k, int capacity,
                         cudaStream_t s) {
  if (hidden & 7) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(dispatched, 0, size_t(experts) * capacity * hidden * 2, s);
  if (e != cudaSuccess) return e;
  if (tokens == 0) return cudaSuccess;
  dispatch_kernel<<<grid_for(tokens * k * (hidden >> 3)), 256, 0, s>>>(x, topk_idx, location, scale, dispatched, tokens,
                                                                       hidden, k, capacity);
  count_launch();
  return cudaGetLastError();
}
cudaError_t moe_combine(const void* expert_out, const int32_t* topk_idx, const int32_t* location, const float* gate,
                        void* y, int64_t tokens, int hidden, int experts, int k, int capacity, cudaStream_t s) {
  if (hidden & 7) return cudaErrorInvalidValue;
  if (tokens == 0) return cudaSuccess;
  combine_kernel<<<grid_for(tokens * (hidden >> 3)), 256, 0, s>>>(expert_out, topk_idx, location, gate, y, tokens,
                                                                  hidden, k, capacity);
  count_launch();
  return cudaGetLastError();
}
cudaError_t moe_dispatch_peers(const void* x, const int32_t* topk_idx, const int32_t* location, const float* scale,
                               const MoePeers& peers, int64_t tokens, int hidden, int k, int capacity, cudaStream_t s) {
  if (hidden & 7) return cudaErrorInvalidValue;
  if (tokens == 0) return cudaSuccess;
  dispatch_peers_kernel<<<grid_for(tokens * k * (hidden >> 3)), 256, 0, s>>>(x, topk_idx, location, scale, peers, tokens, hidden, k,
                                                                             capacity);
  count_launch();
  return cudaGetLastError();
}
cudaError_t moe_combine_peers(const MoePeers& peers, const int32_t* topk_idx, const int32_t* location, const float* gate,
                              void* y, int64_t tokens, int hidden, int k, int capacity, cudaStream_t s) {
  if (hidden & 7) return cudaErrorInvalidValue;
  if (tokens == 0) return cudaSuccess;
  combine_peers_kernel<<<grid_for(tokens * (hidden >> 3)), 256, 0, s>>>(peers, topk_idx, location, gate, y, tokens, hidden, k, capacity);
  count_launch();
  return cudaGetLastError();
}
cudaError_t moe_combine_bwd_gate_peers(const void* dy, const MoePeers& peers, const int32_t* topk_idx, const int32_t* location,
                                       float* dgate, int64_t tokens, int hidden, int k, int capacity, cudaStream_t s) {
  if (tokens == 0) return cudaSuccess;
  combine_bwd_gate_peers_kernel<<<(unsigned)((tokens * k + 7) / 8), 256, 0, s>>>(dy, peers, topk_idx, location, dgate, tokens, hidden, k,
                                                                                 capacity);
  count_launch();
  return cudaGetLastError();
}
cudaError_t moe_combine_bwd_gate(const void* dy, const void* expert_out, const int32_t* topk_idx,
                                 const int32_t* location, float* dgate, int64_t tokens, int hidden, int experts, int k,
                                 int capacity, cudaStream_t s) {
  if (tokens == 0) return cudaSuccess;
  combine_bwd_gate_kernel<<<(unsigned)((tokens * k + 7) / 8), 256, 0, s>>>(dy, expert_out, topk_idx, location, dgate,
                                                                           tokens, hidden, k, capacity);
  count_launch();
  return cudaGetLastError();
}

}  // namespace hb
