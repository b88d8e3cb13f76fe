// Device-side collectives over symmetric (IPC-mapped) peer memory on NVLink 5 / NVSwitch:
// system-scope flag barrier, pull all-gather, pull reduce-scatter (fp32 accumulation), two-shot all-reduce,
// all-to-all, and the slot reduction that completes the fused GEMM -> reduce-scatter epilogue.
// Every data access is a plain 16-byte ld.global / st.global on a mapped peer pointer, issued from all SMs.
#include <cuda_bf16.h>

#include <atomic>

#include "../runtime/symm_mem.h"
#include "common.cuh"

namespace hb {

static std::atomic<int64_t> g_symm_launches{0};
int64_t symm_launch_count() { return g_symm_launches.load(); }

namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// one CTA: thread r publishes this rank's arrival to peer r, then waits for peer r's arrival
__global__ void barrier_kernel(uint32_t* const* flags, int rank, int world, uint32_t epoch, int slot) {
  const int r = threadIdx.x;
  if (r < world) {
    __threadfence_system();
    st_release_sys(flags[r] + slot * kMaxPeers + rank, epoch);
    const uint32_t* mine = flags[rank] + slot * kMaxPeers + r;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) { __nanosleep(20); }
  }
}

__global__ void all_gather_kernel(void* const* peers, size_t src_off, uint4* __restrict__ out, size_t vecs_per_rank, int world) {
  const size_t total = vecs_per_rank * world;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int r = int(i / vecs_per_rank);
    const size_t j = i - size_t(r) * vecs_per_rank;
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(peers[r]) + src_off);
    out[i] = src[j];
  }
}

template <bool kBf16>
__global__ void reduce_scatter_kernel(void* const* peers, size_t src_off, void* __restrict__ out, size_t elems_per_rank, int rank,
                                      int world) {
  constexpr int kPer = kBf16 ? 8 : 4;
  const size_t vecs = elems_per_rank / kPer;
  for (size_t v = blockIdx.x * size_t(blockDim.x) + threadIdx.x; v < vecs; v += size_t(gridDim.x) * blockDim.x) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const size_t idx = size_t(rank) * vecs + v;   // my chunk inside every rank's array
#pragma unroll 1
    for (int rr = 0; rr < world; ++rr) {
      const int r = (rank + rr) % world;          // stagger peers so the ranks do not all hit the same source at once
      const uint4 raw = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(peers[r]) + src_off)[idx];
      if (kBf16) {
        const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = __bfloat1622float2(b2[t]);
          acc[2 * t] += f.x; acc[2 * t + 1] += f.y;
        }
      } else {
        const float* f = reinterpret_cast<const float*>(&raw);
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[t] += f[t];
      }
    }
    if (kBf16) {
      uint4 o;
      __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int t = 0; t < 4; ++t) o2[t] = __floats2bfloat162_rn(acc[2 * t], acc[2 * t + 1]);
      reinterpret_cast<uint4*>(out)[v] = o;
    } else {
      reinterpret_cast<float4*>(out)[v] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
  }
}

__global__ void all_to_all_kernel(void* const* peers, size_t src_off, uint4* __restrict__ out, size_t vecs_per_chunk, int rank,
                                  int world) {
  const size_t total = vecs_per_chunk * world;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int r = int(i / vecs_per_chunk);
    const size_t j = i - size_t(r) * vecs_per_chunk;
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(peers[r]) + src_off);
    out[i] = src[size_t(rank) * vecs_per_chunk + j];
  }
}

__global__ void reduce_slots_kernel(const __nv_bfloat16* __restrict__ slots, int world, __nv_bfloat16* __restrict__ out,
                                    const __nv_bfloat16* __restrict__ bias, const __nv_bfloat16* __restrict__ residual,
                                    int64_t rows, int cols) {
  const int cv = cols >> 3;
  const int64_t total = rows * cv;
  const int64_t slot_vecs = total;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int r = 0; r < world; ++r) {
      float f[8];
      unpack8(ld8_stream(slots, r * slot_vecs + i), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    const int c = int(i % cv);
    if (bias) {
      float f[8];
      unpack8(ld8(bias, c), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    if (residual) {
      float f[8];
      unpack8(ld8_stream(residual, i), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    st8(out, i, pack8(acc));
  }
}


// ---- NVLS: the NVSwitch reduces (multimem.ld_reduce) and replicates (multimem.st) -------------------------------------
__device__ __forceinline__ uint4 mc_ld_reduce_bf16(const void* mc_addr) {       // 8 bf16, fp32 accumulation in the switch
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc_addr) : "memory");
  return v;
}
__device__ __forceinline__ uint4 mc_ld_reduce_f32(const void* mc_addr) {        // 4 fp32
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc_addr) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(void* mc_addr, const uint4& v) {           // 16 bytes into every rank's buffer
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(mc_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <bool kBf16, bool kBroadcast>
__global__ void mc_reduce_kernel(char* mc_base, size_t src_off, uint4* __restrict__ out, size_t vecs_per_rank, int rank) {
  char* mine = mc_base + src_off + size_t(rank) * vecs_per_rank * 16;
  for (size_t v = blockIdx.x * size_t(blockDim.x) + threadIdx.x; v < vecs_per_rank; v += size_t(gridDim.x) * blockDim.x) {
    const uint4 r = kBf16 ? mc_ld_reduce_bf16(mine + v * 16) : mc_ld_reduce_f32(mine + v * 16);
    if (kBroadcast) mc_st(mine + v * 16, r);
    else out[v] = r;
  }
}

__global__ void mc_all_gather_kernel(char* mc_base, size_t dst_off, const uint4* __restrict__ src, size_t vecs_per_rank, int rank) {
  char* dst = mc_base + dst_off + size_t(rank) * vecs_per_rank * 16;
  for (size_t v = blockIdx.x * size_t(blockDim.x) + threadIdx.x; v < vecs_per_rank; v += size_t(gridDim.x) * blockDim.x)
    mc_st(dst + v * 16, src[v]);
}

inline int grid_for(size_t n) {
  size_t blocks = (n + 255) / 256;
  const size_t cap = size_t(sm_count()) * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace

cudaError_t symm_barrier(SymmBuffer& b, cudaStream_t s) {
  b.epoch += 1;
  barrier_kernel<<<1, 32, 0, s>>>(b.d_flags, b.rank, b.world, b.epoch, 0);
  g_symm_launches.fetch_add(1);
  return cudaGetLastError();
}

cudaError_t symm_barrier_slot(SymmBuffer& b, int slot, uint32_t epoch, cudaStream_t s) {
  if (slot < 1 || slot >= kFlagWords / kMaxPeers) return cudaErrorInvalidValue;
  barrier_kernel<<<1, 32, 0, s>>>(b.d_flags, b.rank, b.world, epoch, slot);
  g_symm_launches.fetch_add(1);
  return cudaGetLastError();
}

cudaError_t symm_all_gather(SymmBuffer& b, size_t src_off, void* out, size_t bytes_per_rank, cudaStream_t s) {
  if ((bytes_per_rank & 15) || (src_off & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return cudaErrorMisalignedAddress;
  cudaError_t e = symm_barrier(b, s);                   // every rank's source is complete
  if (e != cudaSuccess) return e;
  all_gather_kernel<<<grid_for(bytes_per_rank / 16 * b.world), 256, 0, s>>>(b.d_peer, src_off, reinterpret_cast<uint4*>(out),
                                                                            bytes_per_rank / 16, b.world);
  g_symm_launches.fetch_add(1);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return symm_barrier(b, s);                            // sources may be overwritten again
}

cudaError_t symm_reduce_scatter(SymmBuffer& b, size_t src_off, void* out, size_t elems_per_rank, bool bf16, cudaStream_t s) {
  const size_t per = bf16 ? 8 : 4;
  if ((elems_per_rank % per) || (src_off & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return cudaErrorMisalignedAddress;
  cudaError_t e = symm_barrier(b, s);
  if (e != cudaSuccess) return e;
  if (bf16) reduce_scatter_kernel<true><<<grid_for(elems_per_rank / per), 256, 0, s>>>(b.d_peer, src_off, out, elems_per_rank, b.rank, b.world);
  else reduce_scatter_kernel<false><<<grid_for(elems_per_rank / per), 256, 0, s>>>(b.d_peer, src_off, out, elems_per_rank, b.rank, b.world);
  g_symm_launches.fetch_add(1);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return symm_barrier(b, s);
}

cudaError_t symm_all_reduce(SymmBuffer& b, size_t src_off, size_t elems, bool bf16, cudaStream_t s) {
  const size_t es = bf16 ? 2 : 4;
  if (elems % (size_t(b.world) * (bf16 ? 8 : 4))) return cudaErrorInvalidValue;
  const size_t per_rank = elems / b.world;
  // shot 1: my chunk <- sum over peers (written in place into my own copy: only I write chunk `rank` of my buffer,
  //         and peers read chunk `rank` of my buffer only in shot 2, after the barrier)
  cudaError_t e = symm_barrier(b, s);
  if (e != cudaSuccess) return e;
  char* mine = reinterpret_cast<char*>(b.local) + src_off;
  // reduce into a scratch area past the data (peers are still reading my chunk `rank` during shot 1)
  void* scratch = mine + elems * es;
  if (src_off + 2 * elems * es > b.bytes + 0 && src_off + elems * es + per_rank * es > b.bytes) return cudaErrorInvalidValue;
  if (bf16) reduce_scatter_kernel<true><<<grid_for(per_rank / 8), 256, 0, s>>>(b.d_peer, src_off, scratch, per_rank, b.rank, b.world);
  else reduce_scatter_kernel<false><<<grid_for(per_rank / 4), 256, 0, s>>>(b.d_peer, src_off, scratch, per_rank, b.rank, b.world);
  g_symm_launches.fetch_add(1);
  e = symm_barrier(b, s);                               // everyone finished reading the un-reduced data
  if (e != cudaSuccess) return e;
  // shot 2: gather the reduced chunks (each rank's chunk sits in its scratch area)
  all_gather_kernel<<<grid_for(per_rank * es / 16 * b.world), 256, 0, s>>>(b.d_peer, src_off + elems * es,
                                                                          reinterpret_cast<uint4*>(mine), per_rank * es / 16, b.world);
  g_symm_launches.fetch_add(1);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return symm_barrier(b, s);
}

cudaError_t symm_reduce_slots(const void* slots, int world, void* out, const void* bias, const void* residual, int64_t rows,
                              int cols, cudaStream_t s) {
  if (cols & 7) return cudaErrorInvalidValue;
  reduce_slots_kernel<<<grid_for(size_t(rows) * (cols >> 3)), 256, 0, s>>>(
      (const __nv_bfloat16*)slots, world, (__nv_bfloat16*)out, (const __nv_bfloat16*)bias, (const __nv_bfloat16*)residual, rows, cols);
  g_symm_launches.fetch_add(1);
  return cudaGetLastError();
}

cudaError_t symm_all_to_all(SymmBuffer& b, size_t src_off, void* out, size_t bytes_per_chunk, cudaStream_t s) {
  if ((bytes_per_chunk & 15) || (src_off & 15)) return cudaErrorMisalignedAddress;
  cudaError_t e = symm_barrier(b, s);
  if (e != cudaSuccess) return e;
  all_to_all_kernel<<<grid_for(bytes_per_chunk / 16 * b.world), 256, 0, s>>>(b.d_peer, src_off, reinterpret_cast<uint4*>(out),
                                                                             bytes_per_chunk / 16, b.rank, b.world);
  g_symm_launches.fetch_add(1);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return symm_barrier(b, s);
}

cudaError_t symm_mc_all_reduce(SymmBuffer& b, size_t src_off, size_t elems, bool bf16, cudaStream_t s) {
  if (b.mc == nullptr) return cudaErrorNotSupported;
  const size_t per = bf16 ? 8 : 4;
  if (elems % (size_t(b.world) * per) || (src_off & 15)) return cudaErrorInvalidValue;
  const size_t vecs = elems / b.world / per;
  cudaError_t e = symm_barrier(b, s);                     // every rank's contribution is complete
  if (e != cudaSuccess) return e;
  char* mc = reinterpret_cast<char*>(b.mc);
  if (bf16) mc_reduce_kernel<true, true><<<grid_for(vecs), 256, 0, s>>>(mc, src_off, nullptr, vecs, b.rank);
  else mc_reduce_kernel<false, true><<<grid_for(vecs), 256, 0, s>>>(mc, src_off, nullptr, vecs, b.rank);
  g_symm_launches.fetch_add(1);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return symm_barrier(b, s);                              // all slices have been replicated everywhere
}

cudaError_t symm_mc_reduce_scatter(SymmBuffer& b, size_t src_off, void* out, size_t elems_per_rank, bool bf16, cudaStream_t s) {
  if (b.mc == nullptr) return cudaErrorNotSupported;
  const size_t per = bf16 ? 8 : 4;
  if ((elems_per_rank % per) || (src_off & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return cudaErrorMisalignedAddress;
  const size_t vecs = elems_per_rank / per;
  cudaError_t e = symm_barrier(b, s);
  if (e != cudaSuccess) return e;
  char* mc = reinterpret_cast<char*>(b.mc);
  if (bf16) mc_reduce_kernel<true, false><<<grid_for(vecs), 256, 0, s>>>(mc, src_off, reinterpret_cast<uint4*>(out), vecs, b.rank);
  else mc_reduce_kernel<false, false><<<grid_for(vecs), 256, 0, s>>>(mc, src_off, reinterpret_cast<uint4*>(out), vecs, b.rank);
  g_symm_launches.fetch_add(1);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return symm_barrier(b, s);                              // sources may be overwritten again
}

cudaError_t symm_mc_all_gather(SymmBuffer& b, const void* src, size_t dst_off, size_t bytes_per_rank, cudaStream_t s) {
  if (b.mc == nullptr) return cudaErrorNotSupported;
  if ((bytes_per_rank & 15) || (dst_off & 15) || (reinterpret_cast<uintptr_t>(src) & 15)) return cudaErrorMisalignedAddress;
  cudaError_t e = symm_barrier(b, s);                     // previous readers of the destination are done
  if (e != cudaSuccess) return e;
  mc_all_gather_kernel<<<grid_for(bytes_per_rank / 16), 256, 0, s>>>(reinterpret_cast<char*>(b.mc), dst_off,
                                                                     reinterpret_cast<const uint4*>(src), bytes_per_rank / 16, b.rank);
  g_symm_launches.fetch_add(1);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return symm_barrier(b, s);                              // every rank's shard has landed in every buffer
}

}  // namespace hb
