// Symmetric memory over NVLink/NVSwitch: every rank allocates the same buffer, exchanges CUDA IPC handles and maps all
// peers' buffers, so kernels can load / store / reduce peer memory directly (no NCCL in the data path).
// A small flag pad per buffer provides system-scope release/acquire signalling for device-side barriers and per-tile
// readiness flags of the fused compute+collective kernels.
// (the reference has no equivalent: all GPU traffic goes through NCCL -- hetu/impl/communication/nccl_comm_group.cu)
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "../core/base.h"

namespace hb {

constexpr int kMaxPeers = 8;
constexpr int kFlagWords = 1024;   // uint32 flags per rank per buffer

struct SymmBuffer {
  std::string name;
  size_t bytes = 0;
  int rank = 0, world = 1;
  void* local = nullptr;                 // this rank's allocation (data followed by the flag pad)
  void* peer[kMaxPeers] = {nullptr};     // mapped pointer of every rank's allocation (peer[rank] == local)
  void** d_peer = nullptr;               // device copy of peer[] (data region)
  uint32_t** d_flags = nullptr;          // device copy of every rank's flag pad pointer
  uint32_t* flags_local = nullptr;
  uint32_t epoch = 0;                    // barrier generation (host-tracked, identical on all ranks)
  // ---- VMM / NVLS multicast variant (symm_vmm.cc) ----
  bool vmm = false;                      // allocated with cuMemCreate (shareable fd) instead of cudaMalloc + IPC
  int device = 0;
  size_t map_bytes = 0, granularity = 0;
  unsigned long long mem_handle = 0, mc_handle = 0, peer_handles[kMaxPeers] = {0};
  int mem_fd = -1, mc_fd = -1;
  bool mc_joined = false;                // this device was added to the multicast object
  void* mc = nullptr;                    // multicast mapping of the whole allocation: multimem.st / multimem.ld_reduce address
};

class SymmMem {
 public:
  static SymmMem& get();
  // step 1: allocate locally, returns the 64-byte IPC handle to publish
  std::string alloc(const std::string& name, size_t bytes, int rank, int world);
  // step 2: map the peers (handles[r] = handle published by rank r)
  void open(const std::string& name, const std::vector<std::string>& handles);
  SymmBuffer& buffer(const std::string& name);
  bool has(const std::string& name) const { return bufs_.count(name) > 0; }
  void free_all();
  // ---- VMM allocation with an NVLS multicast mapping (symm_vmm.cc); three collective steps, see the file header
  static bool multicast_supported();
  std::string alloc_vmm(const std::string& name, size_t bytes, int rank, int world);
  void open_vmm(const std::string& name, const std::vector<std::string>& descs);
  void bind_multicast(const std::string& name);

 private:
  void free_vmm(SymmBuffer& b);
  std::map<std::string, SymmBuffer> bufs_;
};

// ---- device-side collectives over a symmetric buffer (symm_comm.cu) ----------------------------------------
// all ranks must call these in the same order; `stream` is the launching stream on every rank
cudaError_t symm_barrier(SymmBuffer& b, cudaStream_t s);
// the same barrier on an independent flag slot (1..127) with a caller-tracked generation: lets a second stream
// synchronise the ranks without disturbing the barriers of the main stream
cudaError_t symm_barrier_slot(SymmBuffer& b, int slot, uint32_t epoch, cudaStream_t s);
// out[r * n : (r+1) * n] = rank r's `n` elements at byte offset `src_off` of the symmetric buffer
cudaError_t symm_all_gather(SymmBuffer& b, size_t src_off, void* out, size_t bytes_per_rank, cudaStream_t s);
// out[0 : n] = sum_r (rank r's chunk `rank` of the `world * n`-element bf16/fp32 array at src_off); fp32 accumulation
cudaError_t symm_reduce_scatter(SymmBuffer& b, size_t src_off, void* out, size_t elems_per_rank, bool bf16, cudaStream_t s);
// in-place all-reduce of `elems` elements at src_off (two-shot: reduce-scatter into place, then all-gather)
cudaError_t symm_all_reduce(SymmBuffer& b, size_t src_off, size_t elems, bool bf16, cudaStream_t s);
// GEMM->reduce-scatter second stage: out[rows, cols] = sum over `world` staging slots (+ bias + residual), bf16
cudaError_t symm_reduce_slots(const void* slots, int world, void* out, const void* bias, const void* residual, int64_t rows,
                              int cols, cudaStream_t s);
// all-to-all of equal chunks: out chunk r = rank r's chunk `rank`
cudaError_t symm_all_to_all(SymmBuffer& b, size_t src_off, void* out, size_t bytes_per_chunk, cudaStream_t s);
// ---- NVLS (multicast) collectives: need b.mc != nullptr; the switch does the reduction / replication ----------
// in-place all-reduce of `elems` bf16 / fp32 elements at src_off: every rank load-reduces its 1/world slice through the
// multicast address (fp32 accumulation in the switch) and multicast-stores the result into all ranks' buffers
cudaError_t symm_mc_all_reduce(SymmBuffer& b, size_t src_off, size_t elems, bool bf16, cudaStream_t s);
// out[0 : n] = sum over ranks of chunk `rank` of the world * n element array at src_off
cudaError_t symm_mc_reduce_scatter(SymmBuffer& b, size_t src_off, void* out, size_t elems_per_rank, bool bf16, cudaStream_t s);
// every rank's buffer [dst_off + r * bytes_per_rank, ...) = rank r's `src` (local memory), one multicast store per rank
cudaError_t symm_mc_all_gather(SymmBuffer& b, const void* src, size_t dst_off, size_t bytes_per_rank, cudaStream_t s);
int64_t symm_launch_count();

}  // namespace hb
