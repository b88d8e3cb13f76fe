#include "symm_mem.h"

#include <cstring>

namespace hb {

SymmMem& SymmMem::get() {
  static SymmMem m;
  return m;
}

#define CU_OK(x)                                                                      \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    HB_CHECK(e_ == cudaSuccess) << #x << " failed: " << cudaGetErrorString(e_);       \
  } while (0)

std::string SymmMem::alloc(const std::string& name, size_t bytes, int rank, int world) {
  HB_CHECK(world <= kMaxPeers) << "symmetric memory supports up to " << kMaxPeers << " ranks per node";
  HB_CHECK(!bufs_.count(name)) << "symmetric buffer " << name << " already exists";
  SymmBuffer b;
  b.name = name;
  b.bytes = (bytes + 255) / 256 * 256;
  b.rank = rank;
  b.world = world;
  const size_t total = b.bytes + size_t(kFlagWords) * sizeof(uint32_t) * world;
  CU_OK(cudaMalloc(&b.local, total));
  CU_OK(cudaMemset(b.local, 0, total));
  CU_OK(cudaDeviceSynchronize());
  b.flags_local = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(b.local) + b.bytes);
  cudaIpcMemHandle_t h;
  CU_OK(cudaIpcGetMemHandle(&h, b.local));
  bufs_[name] = b;
  return std::string(reinterpret_cast<const char*>(&h), sizeof(h));
}

void SymmMem::open(const std::string& name, const std::vector<std::string>& handles) {
  SymmBuffer& b = buffer(name);
  HB_CHECK((int)handles.size() == b.world) << "need one IPC handle per rank";
  void* flag_ptrs[kMaxPeers] = {nullptr};
  for (int r = 0; r < b.world; ++r) {
    if (r == b.rank) b.peer[r] = b.local;
    else {
      cudaIpcMemHandle_t h;
      HB_CHECK(handles[r].size() == sizeof(h)) << "malformed IPC handle from rank " << r;
      std::memcpy(&h, handles[r].data(), sizeof(h));
      CU_OK(cudaIpcOpenMemHandle(&b.peer[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    flag_ptrs[r] = reinterpret_cast<char*>(b.peer[r]) + b.bytes;
  }
  CU_OK(cudaMalloc(&b.d_peer, sizeof(void*) * kMaxPeers));
  CU_OK(cudaMalloc(&b.d_flags, sizeof(void*) * kMaxPeers));
  CU_OK(cudaMemcpy(b.d_peer, b.peer, sizeof(void*) * kMaxPeers, cudaMemcpyHostToDevice));
  CU_OK(cudaMemcpy(b.d_flags, flag_ptrs, sizeof(void*) * kMaxPeers, cudaMemcpyHostToDevice));
}

SymmBuffer& SymmMem::buffer(const std::string& name) {
  auto it = bufs_.find(name);
  HB_CHECK(it != bufs_.end()) << "unknown symmetric buffer " << name;
  return it->second;
}

void SymmMem::free_all() {
  for (auto& kv : bufs_) {
    SymmBuffer& b = kv.second;
    if (b.vmm) { free_vmm(b); continue; }
    for (int r = 0; r < b.world; ++r)
      if (r != b.rank && b.peer[r]) cudaIpcCloseMemHandle(b.peer[r]);
    if (b.d_peer) cudaFree(b.d_peer);
    if (b.d_flags) cudaFree(b.d_flags);
    if (b.local) cudaFree(b.local);
  }
  bufs_.clear();
}

}  // namespace hb
