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


// ---------------------------------------------------------------- BASE-layer balanced assignment
// Every expert receives exactly `capacity` (= ceil(T / E)) tokens.  Rounds of (choose, accept): each unassigned token
// proposes to its best expert that still has room; an expert with more proposers than room keeps the highest scoring
// ones (radix select of the r-th largest score, ties in token order).  Each round either places every remaining token or
// fills at least one expert, so E rounds always suffice -- no host round trips (the reference's auction,
// hetu/v1/python/hetu/gpu_ops/BalanceAssignment.py, copies bids to the host every iteration).
__global__ void balance_choose_kernel(const float* __restrict__ scores, const int32_t* __restrict__ idx,
                                      const int32_t* __restrict__ filled, int32_t* __restrict__ choice, int64_t tokens,
                                      int experts, int capacity) {
  const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (t >= tokens) return;
  int c = -1;
  if (idx[t] < 0) {
    float best = -INFINITY;
    for (int e = 0; e < experts; ++e) {
      if (filled[e] >= capacity) continue;
      const float v = scores[t * experts + e];
      if (c < 0 || v > best) { best = v; c = e; }
    }
  }
  choice[t] = c;
}

__device__ __forceinline__ uint32_t order_key(float v) {      // larger float <=> larger key (NaN-free inputs)
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ int block_sum_int(int v, int* smem) {
  v = __reduce_add_sync(0xffffffffu, v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
  __syncthreads();
  int tot = 0;
  for (int w = 0; w < (blockDim.x >> 5); ++w) tot += smem[w];
  return tot;
}

__global__ void balance_accept_kernel(const float* __restrict__ scores, const int32_t* __restrict__ choice,
                                      int32_t* __restrict__ idx, int32_t* __restrict__ loc, int32_t* __restrict__ filled,
                                      int64_t tokens, int experts, int capacity) {
  __shared__ int red[32];
  __shared__ int scan_base[2];
  const int e = blockIdx.x;
  const int have = filled[e];
  const int room = capacity - have;
  if (room <= 0) return;
  int n = 0;
  for (int64_t t = threadIdx.x; t < tokens; t += blockDim.x) n += (choice[t] == e);
  n = block_sum_int(n, red);
  if (n == 0) return;
  uint32_t thr = 0;          // accept keys > thr, plus the first `need_eq` proposers with key == thr
  int need_eq = 0;
  if (n > room) {
    int need = room;
    uint32_t prefix = 0;
    for (int b = 31; b >= 0; --b) {
      const uint32_t test = prefix | (1u << b);
      const uint32_t mask = ~((1u << b) - 1u);
      int c1 = 0;
      for (int64_t t = threadIdx.x; t < tokens; t += blockDim.x)
        if (choice[t] == e) c1 += ((order_key(scores[t * experts + e]) & mask) == test);
      c1 = block_sum_int(c1, red);
      if (c1 >= need) prefix = test;
      else need -= c1;
    }
    thr = prefix;
    need_eq = need;
  } else {
    need_eq = n;             // accept everything: treat all as "equal" candidates in token order
  }
  const bool all = n <= room;
  // placement in token order: running counts of accepted (-> slot) and of accepted ties
  if (threadIdx.x == 0) { scan_base[0] = 0; scan_base[1] = 0; }
  __syncthreads();
  for (int64_t base = 0; base < tokens; base += blockDim.x) {
    const int64_t t = base + threadIdx.x;
    bool mine = t < tokens && choice[t] == e;
    uint32_t key = mine ? order_key(scores[t * experts + e]) : 0u;
    const bool gt = mine && !all && key > thr;
    const bool eq = mine && (all || key == thr);
    // block-wide exclusive scans of eq and of (gt | accepted eq); two passes through warp ballots
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned eq_mask = __ballot_sync(0xffffffffu, eq);
    __syncthreads();
    if (lane == 0) red[warp] = __popc(eq_mask);
    __syncthreads();
    int eq_before = scan_base[1];
    for (unsigned w = 0; w < warp; ++w) eq_before += red[w];
    int eq_total = 0;
    for (unsigned w = 0; w < (blockDim.x >> 5); ++w) eq_total += red[w];
    eq_before += __popc(eq_mask & ((1u << lane) - 1u));
    const bool take = gt || (eq && eq_before < need_eq);
    const unsigned tk_mask = __ballot_sync(0xffffffffu, take);
    __syncthreads();
    if (lane == 0) red[warp] = __popc(tk_mask);
    __syncthreads();
    int tk_before = scan_base[0];
    for (unsigned w = 0; w < warp; ++w) tk_before += red[w];
    int tk_total = 0;
    for (unsigned w = 0; w < (blockDim.x >> 5); ++w) tk_total += red[w];
    tk_before += __popc(tk_mask & ((1u << lane) - 1u));
    if (take) { idx[t] = e; loc[t] = have + tk_before; }
    __syncthreads();
    if (threadIdx.x == 0) { scan_base[0] += tk_total; scan_base[1] += eq_total; }
    __syncthreads();
  }
  if (threadIdx.x == 0) filled[e] = have + scan_base[0];
}

// ---------------------------------------------------------------- hierarchical all-to-all layout transform
// x viewed as [a][b][chunk] -> y[b][a][chunk]: regroups the per-destination chunks between the intra-node and the
// inter-node exchange of the two-level all-to-all (ref: hetu/v1/src/ops/H_A2A_LayoutTransform.cu, node-major <-> gpu-major)
__global__ void chunk_transpose_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int a, int b, int64_t vecs) {
  const int64_t total = int64_t(a) * b * vecs;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t v = i % vecs;
    const int64_t ab = i / vecs;          // destination chunk index = bi * a + ai
    const int ai = int(ab % a), bi = int(ab / a);
    y[i] = x[(int64_t(ai) * b + bi) * vecs + v];
  }
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

namespace hb {
cudaError_t moe_balance_assign(const float* scores, int32_t* idx, int32_t* loc, int32_t* filled, int32_t* choice,
                               int64_t tokens, int experts, int capacity, cudaStream_t s) {
  cudaError_t err = cudaMemsetAsync(idx, 0xff, sizeof(int32_t) * tokens, s);       // -1 = unassigned
  if (err != cudaSuccess) return err;
  err = cudaMemsetAsync(filled, 0, sizeof(int32_t) * experts, s);
  if (err != cudaSuccess) return err;
  const int threads = 256;
  const unsigned blocks = (unsigned)((tokens + threads - 1) / threads);
  for (int round = 0; round < experts; ++round) {
    balance_choose_kernel<<<blocks, threads, 0, s>>>(scores, idx, filled, choice, tokens, experts, capacity);
    balance_accept_kernel<<<experts, 1024, 0, s>>>(scores, choice, idx, loc, filled, tokens, experts, capacity);
    count_launch();
    count_launch();
  }
  return cudaGetLastError();
}
cudaError_t chunk_transpose(const void* x, void* y, int a, int b, int64_t chunk_bytes, cudaStream_t s) {
  if ((chunk_bytes & 15) || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) return cudaErrorMisalignedAddress;
  const int64_t vecs = chunk_bytes / 16, total = int64_t(a) * b * vecs;
  if (total == 0) return cudaSuccess;
  int64_t blocks = (total + 255) / 256;
  if (blocks > int64_t(sm_count()) * 8) blocks = int64_t(sm_count()) * 8;
  chunk_transpose_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), a, b, vecs);
  count_launch();
  return cudaGetLastError();
}
}  // namespace hb
