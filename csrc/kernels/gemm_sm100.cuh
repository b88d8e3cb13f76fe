// Persistent, warp-specialised bf16 GEMM for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] * B[K,N] )        fp32 accumulation in TMEM
//
// Design (B200-first, not a cuBLAS wrapper):
//   * one CTA per SM (or a CTA *pair* on one TPC when kCtaGroup == 2, issuing
//     tcgen05.mma.cta_group::2 with a 256 x 256 x 16 instruction shape),
//   * warp 0 = TMA producer, warp 1 = single-thread MMA issuer, warp 2 = TMEM
//     allocator, warps 4..7 = epilogue (TMEM -> registers -> fused epilogue ->
//     global),
//   * kStages-deep smem ring (128B-swizzled tiles written by TMA, read by
//     tcgen05.mma through shared-memory descriptors),
//   * two 256-column TMEM accumulator stages so the epilogue of tile i overlaps
//     the main loop of tile i+1,
//   * both operands may be K-major ("row-major [rows, K]") or MN-major
//     ("[K, rows]"), which covers forward (X * W^T), dgrad (dY * W) and wgrad
//     (dY^T * X) of a linear layer without any transposing copies.
//
// Capability parity: replaces the cuBLAS call sites of the reference
// (hetu/impl/kernel/MatMul.cu:29-77, Linear.cu:25-105, CUDABlas.cc:119-213).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_sm100.h"
#include "ptx.cuh"

namespace hb {

struct GemmParams {
  int M, N, K;
  void* C;               // [M, ldc] bf16 or fp32
  int64_t ldc;
  const __nv_bfloat16* bias;    // [N] or nullptr
  const __nv_bfloat16* aux_in;  // [M, ld_aux] or nullptr
  __nv_bfloat16* aux_out;       // [M, ld_aux] (value before activation) or nullptr
  int64_t ld_aux;
  int act;
  int aux_mode;
  int accumulate;        // C += result (read-modify-write)
  float alpha;           // scale applied to the accumulator
  // fused GEMM -> reduce-scatter: when rows_per_rank > 0 the tile owning rows [o*rows_per_rank, (o+1)*rows_per_rank) is
  // stored straight into rank o's staging buffer (mapped peer memory over NVLink), slot `my_rank`; ldc = N
  void* peer_c[8];
  int rows_per_rank;
  int my_rank;
  // fp8 (e4m3) operands: C = (A8 * B8) * row_scale[m] * col_scale[n]  -- scales of a 1 x K block-scaled quantisation
  const float* row_scale;   // [M] or nullptr
  const float* col_scale;   // [N] or nullptr
  // fused all-gather -> GEMM: A is a local [M, K] buffer whose row block r (ag_rows_per_rank rows) is rank r's shard.
  // Warp 3 of every CTA copies this CTA's slice of each 128-row chunk from the owner's symmetric buffer (NVLink peer
  // loads) into A and bumps ag_flags[chunk]; the TMA producer waits until all CTAs have delivered a chunk before it
  // loads tiles from it.  Tiles are visited starting at this rank's own rows, in the order the chunks arrive.
  const void* ag_src[8];    // per-rank shard [ag_rows_per_rank, K] bf16 (mapped peer memory); null = plain GEMM
  void* ag_dst;             // == A
  uint32_t* ag_flags;       // [M / 128] zero-initialised
  int ag_world, ag_rank, ag_rows_per_rank;
};

namespace gemm_detail {

constexpr int BLOCK_M = 128;   // rows of A per CTA
constexpr int BLOCK_N = 256;   // default UMMA N (columns of the accumulator); the kernel's kBN parameter may narrow it to 128
constexpr int BLOCK_K = 64;    // 128 bytes of bf16 = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 8;     // 2 warps per TMEM lane quarter, each owning half the columns
constexpr int kNumThreads = 128 + kEpiWarps * 32;
constexpr int kAccumStages = 2;
constexpr int kTmemCols = 512;

__host__ __device__ constexpr int a_stage_bytes() { return BLOCK_M * BLOCK_K * 2; }
__host__ __device__ constexpr int b_stage_bytes(int cta_group, int bn = BLOCK_N) { return (bn / cta_group) * BLOCK_K * 2; }
__host__ __device__ constexpr int smem_bytes(int cta_group, int stages, int bn = BLOCK_N) {
  return stages * (a_stage_bytes() + b_stage_bytes(cta_group, bn)) + 1024 /*align slack*/ + 256 /*barriers*/ +
         kEpiWarps * 4096 /*epilogue store staging*/;
}

// erf via Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7): two MUFU ops + a short FMA chain,
// ~6x cheaper than erff() in the epilogue where 128..256 threads cover a whole tile.
__device__ __forceinline__ float fast_erf(float x) {
  const float ax = fabsf(x);
  const float t = __frcp_rn(fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float r = 1.0f - p * __expf(-ax * ax);
  return copysignf(r, x);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + fast_erf(x * 0.70710678118654752f)); }
// d/dx gelu(x) = Phi(x) + x * phi(x); the erf polynomial and the density share one exponential (exp(-x^2/2))
__device__ __forceinline__ float dgelu_erf(float x) {
  const float ax = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = __expf(-ax * ax);
  const float erf_abs = 1.0f - p * e;
  const float cdf = fmaf(0.5f, copysignf(erf_abs, x), 0.5f);
  return fmaf(x * 0.39894228040143267f, e, cdf);
}
__device__ __forceinline__ float gelu_tanh(float x) {
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.0f + tanhf(u));
}
__device__ __forceinline__ float dgelu_tanh(float x) {
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  const float t = tanhf(u);
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * 0.7978845608028654f * (1.0f + 3.0f * 0.044715f * x * x);
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float dsilu(float x) {
  const float s = 1.0f / (1.0f + __expf(-x));
  return s * (1.0f + x * (1.0f - s));
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_GELU: return gelu_erf(v);
    case ACT_RELU: return fmaxf(v, 0.0f);
    case ACT_GELU_TANH: return gelu_tanh(v);
    case ACT_SILU: return silu(v);
    default: return v;
  }
}
__device__ __forceinline__ float apply_aux(float v, float a, int mode) {
  switch (mode) {
    case AUX_ADD: return v + a;
    case AUX_DGELU: return v * dgelu_erf(a);
    case AUX_DRELU: return a > 0.0f ? v : 0.0f;
    case AUX_DGELU_TANH: return v * dgelu_tanh(a);
    case AUX_DSILU: return v * dsilu(a);
    default: return v;
  }
}

// Group-swizzled tile order: consecutive tile ids walk down a band of kGroupM
// M-blocks before moving to the next N-block, so the ~74 clusters in flight
// share a small set of A and B panels in L2.
__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int& tm, int& tn, int m_rotate = 0) {
  constexpr int kGroupM = 8;
  const int per_group = kGroupM * tiles_n;
  const int g = tile / per_group;
  const int first_m = g * kGroupM;
  const int gm = min(tiles_m - first_m, kGroupM);
  const int in_g = tile - g * per_group;
  tm = first_m + in_g % gm;
  tn = in_g / gm;
  if (m_rotate) { tm += m_rotate; if (tm >= tiles_m) tm -= tiles_m; }
}

}  // namespace gemm_detail

// kFp8: both operands are e4m3 bytes (K-major only); a 128-byte swizzle row then holds 128 K elements and one
// tcgen05.mma.kind::f8f6f4 consumes 32 of them -- the smem stage size, descriptor strides and pipeline are unchanged.
// kBN: accumulator columns per tile.  256 is the default; 128 halves the tile for outputs whose 256-wide tile count does not
// fill the last wave of the persistent grid (e.g. 8192 x 2048: 256 tiles on 74 CTA pairs = 3.46 waves; 512 half tiles = 6.92).
template <int kCtaGroup, bool kAMN, bool kBMN, int kStages, typename OutT, bool kFp8 = false, int kBN = 256>
__global__ void __launch_bounds__(gemm_detail::kNumThreads, 1)
gemm_bf16_sm100_kernel(const __grid_constant__ CUtensorMap tmap_a,
                       const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  using namespace gemm_detail;
  constexpr int A_STAGE = a_stage_bytes();
  constexpr int B_STAGE = b_stage_bytes(kCtaGroup, kBN);
  constexpr int LOAD_N = kBN / kCtaGroup;
  constexpr int TILE_M = BLOCK_M * kCtaGroup;
  constexpr uint32_t kTxBytes = (A_STAGE + B_STAGE) * kCtaGroup;
  constexpr int kAtomBytes = BLOCK_K * 64 * 2;  // one 64(MN) x BLOCK_K MN-major box

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (A_STAGE + B_STAGE));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tfull_bar = bars + 2 * kStages;
  uint64_t* tempty_bar = bars + 2 * kStages + kAccumStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2 * kAccumStages);
  uint8_t* epi_staging = smem + kStages * (A_STAGE + B_STAGE) + 256;   // kEpiWarps x 4 KB

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (kCtaGroup == 2) ? ptx::cluster_ctarank() : 0;
  const bool leader = rank == 0;

  const int tiles_m = (p.M + TILE_M - 1) / TILE_M;
  const int tiles_n = (p.N + kBN - 1) / kBN;
  const int num_tiles = tiles_m * tiles_n;
  constexpr int BLOCK_K_E = kFp8 ? BLOCK_K * 2 : BLOCK_K;   // K elements per stage (128 bytes per row either way)
  static_assert(!kFp8 || (!kAMN && !kBMN), "fp8 operands must be K-major");
  const int num_kb = (p.K + BLOCK_K_E - 1) / BLOCK_K_E;
  const int cluster_id = blockIdx.x / kCtaGroup;
  const int num_clusters = gridDim.x / kCtaGroup;
  const bool ag_fused = p.ag_world > 1;
  const int m_rotate = ag_fused ? (p.ag_rank * p.ag_rows_per_rank) / TILE_M : 0;   // start at this rank's own rows

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_bar[i], kCtaGroup);  // one producer arrive per CTA (+ tx bytes)
      ptx::mbar_init(&empty_bar[i], 1);         // one tcgen05.commit
    }
    for (int i = 0; i < kAccumStages; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);                 // one tcgen05.commit
      ptx::mbar_init(&tempty_bar[i], kEpiWarps * kCtaGroup);  // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<kCtaGroup>(tmem_ptr_smem, kTmemCols);
    ptx::tmem_relinquish<kCtaGroup>();
  }
  ptx::tc_fence_before();
  if constexpr (kCtaGroup == 2) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ========================= TMA producer =========================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int tm, tn; tile_coords(tile, tiles_m, tiles_n, tm, tn, m_rotate);
        const int m0 = tm * TILE_M + int(rank) * BLOCK_M;
        const int n0 = tn * kBN + int(rank) * LOAD_N;
        if (ag_fused && m0 < p.M) {
          // the 128 rows of this CTA's A tile are one all-gather chunk: wait until every CTA delivered its slice
          const uint32_t* f = p.ag_flags + (m0 >> 7);
          uint32_t spins = 0;
          while (ptx::ld_acquire_gpu(f) < 2 * gridDim.x) {        // two puller warps per CTA
            __nanosleep(64);
            if (++spins > (1u << 26)) __trap();   // a lost peer must surface as an error, not as a hung GPU
          }
          ptx::fence_proxy_async();               // generic-proxy copies -> async-proxy (TMA) reads
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[s], ph ^ 1);
          if (kCtaGroup == 1 || leader) ptx::mbar_arrive_expect_tx(&full_bar[s], kTxBytes);
          uint8_t* sa = smem_a + s * A_STAGE;
          uint8_t* sb = smem_b + s * B_STAGE;
          const int k0 = kb * BLOCK_K_E;
          if constexpr (kCtaGroup == 1) {
            if constexpr (!kAMN) ptx::tma_load_2d(sa, &tmap_a, &full_bar[s], k0, m0);
            else
              for (int i = 0; i < BLOCK_M / 64; ++i) ptx::tma_load_2d(sa + i * kAtomBytes, &tmap_a, &full_bar[s], m0 + i * 64, k0);
            if constexpr (!kBMN) ptx::tma_load_2d(sb, &tmap_b, &full_bar[s], k0, n0);
            else
              for (int i = 0; i < LOAD_N / 64; ++i) ptx::tma_load_2d(sb + i * kAtomBytes, &tmap_b, &full_bar[s], n0 + i * 64, k0);
          } else {
            if constexpr (!kAMN) ptx::tma_load_2d_2sm(sa, &tmap_a, &full_bar[s], k0, m0);
            else
              for (int i = 0; i < BLOCK_M / 64; ++i) ptx::tma_load_2d_2sm(sa + i * kAtomBytes, &tmap_a, &full_bar[s], m0 + i * 64, k0);
            if constexpr (!kBMN) ptx::tma_load_2d_2sm(sb, &tmap_b, &full_bar[s], k0, n0);
            else
              for (int i = 0; i < LOAD_N / 64; ++i) ptx::tma_load_2d_2sm(sb + i * kAtomBytes, &tmap_b, &full_bar[s], n0 + i * 64, k0);
            if (!leader) ptx::mbar_arrive_cluster(&full_bar[s], 0);
          }
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ========================= MMA issuer (leader CTA, one thread) =========================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = kFp8 ? ptx::make_idesc(TILE_M, kBN, 0, 0, false, false)      // e4m3 x e4m3 -> f32
                                      : ptx::make_idesc(TILE_M, kBN, 1, 1, kAMN, kBMN);
      const uint64_t a_desc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_a), kAMN ? kAtomBytes : 0, 1024);
      const uint64_t b_desc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_b), kBMN ? kAtomBytes : 0, 1024);
      // descriptor-address increments (units of 16 B) per UMMA_K step
      constexpr uint32_t a_kstep = kAMN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      constexpr uint32_t b_kstep = kBMN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
      int s = 0; uint32_t ph = 0;
      int as = 0; uint32_t aph = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        ptx::mbar_wait(&tempty_bar[as], aph ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kBN;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[s], ph);
          ptx::tc_fence_after();
          const uint64_t a_desc = a_desc0 + uint64_t(s * (A_STAGE >> 4));
          const uint64_t b_desc = b_desc0 + uint64_t(s * (B_STAGE >> 4));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            if constexpr (kFp8)
              ptx::mma_f8_ss<kCtaGroup>(tmem_d, a_desc + uint64_t(k * a_kstep), b_desc + uint64_t(k * b_kstep), idesc,
                                        (kb > 0 || k > 0) ? 1u : 0u);
            else
              ptx::mma_f16_ss<kCtaGroup>(tmem_d, a_desc + uint64_t(k * a_kstep), b_desc + uint64_t(k * b_kstep), idesc,
                                         (kb > 0 || k > 0) ? 1u : 0u);
          }
          ptx::mma_commit<kCtaGroup>(&empty_bar[s]);
          if (kb == num_kb - 1) ptx::mma_commit<kCtaGroup>(&tfull_bar[as]);
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
        if (++as == kAccumStages) { as = 0; aph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 2 || warp == 3) {
    // ========================= all-gather pullers (fused AG -> GEMM only) =========================
    // warps 2 (idle after the TMEM allocation) and 3 of every CTA stream this CTA's slice of each 128-row chunk from the
    // owner's symmetric shard into A: 8 independent 16-byte peer loads in flight per lane
    if (ag_fused) {
      const int chunks_per_rank = p.ag_rows_per_rank >> 7;
      const int64_t chunk_vecs = int64_t(128) * p.K * 2 / 16;        // uint4 per 128-row chunk
      const int64_t lane_id = (int64_t(blockIdx.x) * 2 + (warp - 2)) * 32 + lane;
      const int64_t stride = int64_t(gridDim.x) * 64;
      for (int i = 0; i < p.ag_world; ++i) {
        const int r = (p.ag_rank + i) % p.ag_world;                  // own shard first, then the ring order the tiles follow
        const uint4* src = reinterpret_cast<const uint4*>(p.ag_src[r]);
        uint4* dst = reinterpret_cast<uint4*>(p.ag_dst) + int64_t(r) * chunks_per_rank * chunk_vecs;
        // 8 chunks (1024 rows) are pulled as one batch: a lane owns only a few 16-byte pieces of a single chunk, batching
        // gives it ~8x more independent peer loads in flight; the 8 chunk flags are raised together afterwards
        for (int c0 = 0; c0 < chunks_per_rank; c0 += 8) {
          const int nc = min(8, chunks_per_rank - c0);
          const int64_t batch_vecs = int64_t(nc) * chunk_vecs;
          const uint4* s4 = src + int64_t(c0) * chunk_vecs;
          uint4* d4 = dst + int64_t(c0) * chunk_vecs;
          int64_t v = lane_id;
          for (; v + 7 * stride < batch_vecs; v += 8 * stride) {
            uint4 t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = ptx::ld_global_relaxed_sys(s4 + v + u * stride);
#pragma unroll
            for (int u = 0; u < 8; ++u) d4[v + u * stride] = t[u];
          }
          for (; v < batch_vecs; v += stride) d4[v] = ptx::ld_global_relaxed_sys(s4 + v);
          __syncwarp();
          __threadfence();
          if (lane < nc) atomicAdd(p.ag_flags + r * chunks_per_rank + c0 + lane, 1u);
        }
      }
    }
  } else if (warp >= 4) {
    // ========================= epilogue =========================
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int chalf = (warp - 4) >> 2;  // which half of the accumulator columns this warp drains
    constexpr int kChunksPerWarp = kBN / 32 / (kEpiWarps / 4);
    int as = 0; uint32_t aph = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      int tm, tn; tile_coords(tile, tiles_m, tiles_n, tm, tn, m_rotate);
      const int row = tm * TILE_M + int(rank) * BLOCK_M + q * 32 + lane;
      const int n_base = tn * kBN;
      ptx::mbar_wait(&tfull_bar[as], aph);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + as * kBN;
      const bool row_ok = row < p.M;
      OutT* crow = reinterpret_cast<OutT*>(p.C) + int64_t(row) * p.ldc;
      if (p.rows_per_rank > 0 && row_ok) {
        const int owner = row / p.rows_per_rank;
        crow = reinterpret_cast<OutT*>(p.peer_c[owner]) +
               (int64_t(p.my_rank) * p.rows_per_rank + (row - owner * p.rows_per_rank)) * p.ldc;
      }
      const int row_base = tm * TILE_M + int(rank) * BLOCK_M + q * 32;   // global row of lane 0
      uint8_t* stg = epi_staging + (warp - 4) * 4096;
      auto c_row_ptr = [&](int grow) -> OutT* {
        if (p.rows_per_rank > 0) {   // fused GEMM -> reduce-scatter: the row lives in its owner's staging slot
          const int owner = grow / p.rows_per_rank;
          return reinterpret_cast<OutT*>(p.peer_c[owner]) +
                 (int64_t(p.my_rank) * p.rows_per_rank + (grow - owner * p.rows_per_rank)) * p.ldc;
        }
        return reinterpret_cast<OutT*>(p.C) + int64_t(grow) * p.ldc;
      };
      // aux operand (residual / pre-activation) of the NEXT chunk is fetched while the current one is processed: the
      // ~1 us global latency would otherwise sit between every TMEM drain and its store
      uint4 aux_next[4];
      auto fetch_aux = [&](int cc_) {
        const int col0_ = n_base + (chalf * kChunksPerWarp + cc_) * 32;
        if (p.aux_mode != AUX_NONE && row_ok && col0_ + 32 <= p.N) {
          const uint4* ai4 = reinterpret_cast<const uint4*>(p.aux_in + int64_t(row) * p.ld_aux + col0_);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) aux_next[j4] = __ldcs(ai4 + j4);
        }
      };
      fetch_aux(0);
#pragma unroll 1
      for (int cc = 0; cc < kChunksPerWarp; ++cc) {
        const int c = chalf * kChunksPerWarp + cc;
        uint4 aux_cur[4];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) aux_cur[j4] = aux_next[j4];
        if (cc + 1 < kChunksPerWarp) fetch_aux(cc + 1);
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(taddr + c * 32, r);
        ptx::tmem_ld_wait();
        const int col0 = n_base + c * 32;
        if (col0 >= p.N) continue;  // warp-uniform
        const bool full_chunk = (col0 + 32 <= p.N);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (p.alpha != 1.0f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
        }
        if constexpr (kFp8) {
          const float rs = (p.row_scale != nullptr && row_ok) ? p.row_scale[row] : 1.0f;
          if (p.col_scale != nullptr && full_chunk) {
            const float4* cs4 = reinterpret_cast<const float4*>(p.col_scale + col0);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 c4 = __ldg(cs4 + j4);
              v[j4 * 4] *= rs * c4.x; v[j4 * 4 + 1] *= rs * c4.y; v[j4 * 4 + 2] *= rs * c4.z; v[j4 * 4 + 3] *= rs * c4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= rs * ((p.col_scale != nullptr && col0 + j < p.N) ? p.col_scale[col0 + j] : 1.0f);
          }
        }
        if (p.bias != nullptr) {
          if (full_chunk) {
            const uint4* bp = reinterpret_cast<const uint4*>(p.bias + col0);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const uint4 b = __ldg(bp + j4);
              const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 f = __bfloat1622float2(b2[t]);
                v[j4 * 8 + t * 2] += f.x; v[j4 * 8 + t * 2 + 1] += f.y;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] += __bfloat162float(p.bias[col0 + j]);
          }
        }
        // ---- stores go through a per-warp 32-row x 128-byte staging tile (16-byte units XOR-swizzled by row): every
        // lane parks its own row, then the warp writes 4 rows x 128 contiguous bytes per instruction (full lines for
        // HBM and for NVLink peer stores) instead of 32 scattered 16-byte pieces.
        {
          if (p.aux_out != nullptr) {
            // pre-activation (bf16): staged in this chunk's half of the tile, flushed at once (64-byte row pieces)
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              uint4 o; __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int t = 0; t < 4; ++t) o2[t] = __floats2bfloat162_rn(v[j4 * 8 + t * 2], v[j4 * 8 + t * 2 + 1]);
              *reinterpret_cast<uint4*>(stg + lane * 128 + ((((cc & 1) * 4 + j4) ^ (lane & 7)) << 4)) = o;
            }
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int rl = it * 8 + (lane >> 2), u = lane & 3;
              const int grow = row_base + rl, gcol = col0 + u * 8;
              if (grow < p.M && gcol < p.N) {
                const uint4 val = *reinterpret_cast<const uint4*>(stg + rl * 128 + ((((cc & 1) * 4 + u) ^ (rl & 7)) << 4));
                __nv_bfloat16* dst = p.aux_out + int64_t(grow) * p.ld_aux + gcol;
                if (gcol + 8 <= p.N) *reinterpret_cast<uint4*>(dst) = val;
                else {
                  const uint32_t w4[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
                  for (int t = 0; t < 8; ++t)
                    if (gcol + t < p.N) dst[t] = __ushort_as_bfloat16((unsigned short)(w4[t >> 1] >> ((t & 1) * 16)));
                }
              }
            }
            __syncwarp();
          }
          if (p.act == ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
          } else if (p.act != ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act);
          }
          if (p.aux_mode != AUX_NONE) {
            const __nv_bfloat16* ai = p.aux_in + int64_t(row) * p.ld_aux + col0;
            if (full_chunk) {
              float a[32];
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&aux_cur[j4]);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const float2 f = __bfloat1622float2(a2[t]);
                  a[j4 * 8 + t * 2] = f.x; a[j4 * 8 + t * 2 + 1] = f.y;
                }
              }
              if (p.aux_mode == AUX_ADD) {          // the mode is uniform: keep the switch out of the unrolled loops
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += a[j];
              } else if (p.aux_mode == AUX_DGELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] *= dgelu_erf(a[j]);
              } else if (p.aux_mode == AUX_DSILU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] *= dsilu(a[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = apply_aux(v[j], a[j], p.aux_mode);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (row_ok && col0 + j < p.N) v[j] = apply_aux(v[j], __bfloat162float(ai[j]), p.aux_mode);
            }
          }
          if constexpr (sizeof(OutT) == 2) {
            if (p.accumulate && row_ok) {   // rare (bf16 accumulation): old values fetched by the owning lane
              const __nv_bfloat16* cp = reinterpret_cast<const __nv_bfloat16*>(crow) + col0;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) v[j] += __bfloat162float(cp[j]);
            }
            // stage this chunk's 32 bf16 columns into its half of the 64-column tile
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              uint4 o; __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int t = 0; t < 4; ++t) o2[t] = __floats2bfloat162_rn(v[j4 * 8 + t * 2], v[j4 * 8 + t * 2 + 1]);
              *reinterpret_cast<uint4*>(stg + lane * 128 + ((((cc & 1) * 4 + j4) ^ (lane & 7)) << 4)) = o;
            }
            const bool flush = (cc & 1) || (cc + 1 == kChunksPerWarp) || (col0 + 32 >= p.N);
            if (flush) {
              __syncwarp();
              const int span0 = col0 - (cc & 1) * 32;
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const int rl = it * 4 + (lane >> 3), u = lane & 7;
                const int grow = row_base + rl, gcol = span0 + u * 8;
                if (grow < p.M && gcol < p.N && (u < 4 || (cc & 1))) {
                  const uint4 val = *reinterpret_cast<const uint4*>(stg + rl * 128 + ((u ^ (rl & 7)) << 4));
                  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(c_row_ptr(grow)) + gcol;
                  if (gcol + 8 <= p.N) *reinterpret_cast<uint4*>(dst) = val;
                  else {
                    const uint32_t w4[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                      if (gcol + t < p.N) dst[t] = __ushort_as_bfloat16((unsigned short)(w4[t >> 1] >> ((t & 1) * 16)));
                  }
                }
              }
              __syncwarp();
            }
          } else {
            // fp32 output: the 32 columns of this chunk are one 128-byte row of the staging tile
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
              *reinterpret_cast<float4*>(stg + lane * 128 + ((j4 ^ (lane & 7)) << 4)) =
                  make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int rl = it * 4 + (lane >> 3), u = lane & 7;
              const int grow = row_base + rl, gcol = col0 + u * 4;
              if (grow < p.M && gcol < p.N) {
                float4 val = *reinterpret_cast<const float4*>(stg + rl * 128 + ((u ^ (rl & 7)) << 4));
                float* dst = reinterpret_cast<float*>(c_row_ptr(grow)) + gcol;
                if (gcol + 4 <= p.N) {
                  if (p.accumulate) {
                    const float4 old = *reinterpret_cast<const float4*>(dst);
                    val.x += old.x; val.y += old.y; val.z += old.z; val.w += old.w;
                  }
                  *reinterpret_cast<float4*>(dst) = val;
                } else {
                  const float e4[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
                  for (int t = 0; t < 4; ++t)
                    if (gcol + t < p.N) dst[t] = p.accumulate ? dst[t] + e4[t] : e4[t];
                }
              }
            }
            __syncwarp();
          }
        }
      }
      // release this accumulator stage back to the MMA issuer (leader CTA's barrier)
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kCtaGroup == 1) ptx::mbar_arrive(&tempty_bar[as]);
        else ptx::mbar_arrive_cluster(&tempty_bar[as], 0);
      }
      if (++as == kAccumStages) { as = 0; aph ^= 1; }
    }
  }

  // ========================= teardown =========================
  ptx::tc_fence_before();
  if constexpr (kCtaGroup == 2) ptx::cluster_sync(); else __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kCtaGroup>(tmem_base, kTmemCols);
  }
}

}  // namespace hb
