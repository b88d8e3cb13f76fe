// Symmetric memory on the CUDA virtual-memory-management API with an NVLink-SHARP (NVLS) multicast mapping.
//
// Every rank creates one physical allocation (cuMemCreate, exportable as a POSIX file descriptor), the ranks import each
// other's allocations (the descriptor is duplicated out of the owning process with pidfd_getfd, or -- where the
// container forbids that -- received as an SCM_RIGHTS message over an abstract unix socket) and map them into their
// address space, exactly like the IPC path of symm_mem.cc.  In addition rank 0
// creates a *multicast object* spanning all GPUs of the group; every rank adds its device, binds its physical
// allocation and maps the object: a store to the multicast address is replicated by the NVSwitch into every rank's
// buffer (multimem.st), a load-reduce returns the sum over all ranks' buffers computed inside the switch
// (multimem.ld_reduce) -- the kernels of symm_comm.cu build all-gather / reduce-scatter / all-reduce and the ZeRO
// reduce + AdamW + broadcast step on these two instructions.
//
// Three collective steps (the caller exchanges the descriptors, e.g. with all_gather_object, and puts a barrier
// between open_vmm and bind_multicast because every device must have joined the object before memory is bound):
//     desc = alloc_vmm(name, bytes, rank, world)   ->   open_vmm(name, descs)   ->   [barrier]   ->   bind_multicast(name)
//
// The driver entry points are resolved through the runtime (cudaGetDriverEntryPoint): nothing links libcuda, the
// module still imports on a machine without a driver.
// (the reference has no equivalent: all GPU traffic goes through NCCL -- hetu/impl/communication/nccl_comm_group.cu)
#include <cuda.h>
#include <sys/prctl.h>
#include <sys/socket.h>
#include <sys/syscall.h>
#include <sys/un.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <cstring>
#include <thread>

#include <map>
#include <memory>

#include "symm_mem.h"

namespace hb {

namespace {

template <typename F>
F drv(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
  HB_CHECK(e == cudaSuccess && p != nullptr) << "cannot resolve driver entry point " << name << ": " << cudaGetErrorString(e);
  return reinterpret_cast<F>(p);
}

#define DRV_OK(call)                                                         \
  do {                                                                       \
    CUresult r_ = (call);                                                    \
    HB_CHECK(r_ == CUDA_SUCCESS) << #call << " failed with CUresult " << (int)r_; \
  } while (0)

using PFN_cuMemCreate = CUresult (*)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
using PFN_cuMemRelease = CUresult (*)(CUmemGenericAllocationHandle);
using PFN_cuMemExport = CUresult (*)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
using PFN_cuMemImport = CUresult (*)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
using PFN_cuMemAddressReserve = CUresult (*)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
using PFN_cuMemAddressFree = CUresult (*)(CUdeviceptr, size_t);
using PFN_cuMemMap = CUresult (*)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
using PFN_cuMemUnmap = CUresult (*)(CUdeviceptr, size_t);
using PFN_cuMemSetAccess = CUresult (*)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
using PFN_cuMemGetGranularity = CUresult (*)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
using PFN_cuMulticastCreate = CUresult (*)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
using PFN_cuMulticastAddDevice = CUresult (*)(CUmemGenericAllocationHandle, CUdevice);
using PFN_cuMulticastBindMem = CUresult (*)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t,
                                            unsigned long long);
using PFN_cuMulticastGetGranularity = CUresult (*)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
using PFN_cuDeviceGet = CUresult (*)(CUdevice*, int);
using PFN_cuDeviceGetAttribute = CUresult (*)(int*, CUdevice_attribute, CUdevice);

struct Driver {
  PFN_cuMemCreate memCreate = drv<PFN_cuMemCreate>("cuMemCreate");
  PFN_cuMemRelease memRelease = drv<PFN_cuMemRelease>("cuMemRelease");
  PFN_cuMemExport memExport = drv<PFN_cuMemExport>("cuMemExportToShareableHandle");
  PFN_cuMemImport memImport = drv<PFN_cuMemImport>("cuMemImportFromShareableHandle");
  PFN_cuMemAddressReserve addressReserve = drv<PFN_cuMemAddressReserve>("cuMemAddressReserve");
  PFN_cuMemAddressFree addressFree = drv<PFN_cuMemAddressFree>("cuMemAddressFree");
  PFN_cuMemMap memMap = drv<PFN_cuMemMap>("cuMemMap");
  PFN_cuMemUnmap memUnmap = drv<PFN_cuMemUnmap>("cuMemUnmap");
  PFN_cuMemSetAccess setAccess = drv<PFN_cuMemSetAccess>("cuMemSetAccess");
  PFN_cuMemGetGranularity getGranularity = drv<PFN_cuMemGetGranularity>("cuMemGetAllocationGranularity");
  PFN_cuMulticastCreate mcCreate = drv<PFN_cuMulticastCreate>("cuMulticastCreate");
  PFN_cuMulticastAddDevice mcAddDevice = drv<PFN_cuMulticastAddDevice>("cuMulticastAddDevice");
  PFN_cuMulticastBindMem mcBindMem = drv<PFN_cuMulticastBindMem>("cuMulticastBindMem");
  PFN_cuMulticastGetGranularity mcGranularity = drv<PFN_cuMulticastGetGranularity>("cuMulticastGetGranularity");
  PFN_cuDeviceGet deviceGet = drv<PFN_cuDeviceGet>("cuDeviceGet");
  PFN_cuDeviceGetAttribute deviceGetAttribute = drv<PFN_cuDeviceGetAttribute>("cuDeviceGetAttribute");
};
Driver& D() {
  static Driver d;
  return d;
}

struct VmmDesc {          // what a rank publishes (plain bytes, exchanged by the caller)
  int32_t pid;
  int32_t mem_fd;
  int32_t mc_fd;          // rank 0 only, -1 elsewhere (or when multicast is unsupported)
  int32_t mc_ok;          // this rank's device supports multicast
  uint64_t size;          // physical size (rounded to the granularity)
  int32_t sock_seq;       // this rank serves its descriptors on the abstract unix socket "hetu_symm_<pid>_<sock_seq>"
  int32_t pad;
};

// ---- handing a file descriptor to a peer process.  Fast path: pidfd_getfd (needs ptrace permission over the owner --
// granted to everybody by the owner through PR_SET_PTRACER_ANY when Yama's ptrace_scope is 1).  Containers that forbid it
// fall back to the classic SCM_RIGHTS message over an abstract unix socket served by a thread of the owner.
void fill_addr(sockaddr_un* a, socklen_t* len, int pid, int seq) {
  std::memset(a, 0, sizeof(*a));
  a->sun_family = AF_UNIX;
  const std::string name = "hetu_symm_" + std::to_string(pid) + "_" + std::to_string(seq);
  std::memcpy(a->sun_path + 1, name.data(), name.size());          // leading NUL: abstract namespace, nothing on disk
  *len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + name.size());
}

struct FdServer {
  int listen_fd = -1;
  int fds[2] = {-1, -1};            // [0] memory handle, [1] multicast handle
  std::thread th;
  std::atomic<bool> stop{false};
  void start(int seq) {
    listen_fd = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    HB_CHECK(listen_fd >= 0) << "socket(AF_UNIX) failed: " << std::strerror(errno);
    sockaddr_un a;
    socklen_t len;
    fill_addr(&a, &len, (int)getpid(), seq);
    HB_CHECK(bind(listen_fd, (sockaddr*)&a, len) == 0) << "bind of the descriptor socket failed: " << std::strerror(errno);
    HB_CHECK(listen(listen_fd, 32) == 0) << "listen failed: " << std::strerror(errno);
    th = std::thread([this] {
      while (!stop.load()) {
        const int c = accept(listen_fd, nullptr, nullptr);
        if (c < 0) break;
        int32_t which = 0;
        if (recv(c, &which, sizeof(which), MSG_WAITALL) == (ssize_t)sizeof(which) && which >= 0 && which < 2 && fds[which] >= 0) {
          char payload = 'f';
          iovec io{&payload, 1};
          alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
          std::memset(ctrl, 0, sizeof(ctrl));
          msghdr msg{};
          msg.msg_iov = &io; msg.msg_iovlen = 1;
          msg.msg_control = ctrl; msg.msg_controllen = sizeof(ctrl);
          cmsghdr* cm = CMSG_FIRSTHDR(&msg);
          cm->cmsg_level = SOL_SOCKET; cm->cmsg_type = SCM_RIGHTS; cm->cmsg_len = CMSG_LEN(sizeof(int));
          std::memcpy(CMSG_DATA(cm), &fds[which], sizeof(int));
          sendmsg(c, &msg, 0);
        }
        close(c);
      }
    });
  }
  ~FdServer() { shutdown_server(); }
  void shutdown_server() {
    stop.store(true);
    if (listen_fd >= 0) { ::shutdown(listen_fd, SHUT_RDWR); close(listen_fd); listen_fd = -1; }
    if (th.joinable()) th.join();
  }
};
std::map<std::string, std::unique_ptr<FdServer>>& servers() {
  static std::map<std::string, std::unique_ptr<FdServer>> m;
  return m;
}

int recv_fd_over_socket(int pid, int seq, int which) {
  const int c = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
  HB_CHECK(c >= 0) << "socket(AF_UNIX) failed: " << std::strerror(errno);
  sockaddr_un a;
  socklen_t len;
  fill_addr(&a, &len, pid, seq);
  int rc = -1;
  for (int attempt = 0; attempt < 200 && rc != 0; ++attempt) {        // the peer's server thread may not be listening yet
    rc = connect(c, (sockaddr*)&a, len);
    if (rc != 0) usleep(10000);
  }
  HB_CHECK(rc == 0) << "cannot reach the descriptor socket of process " << pid << ": " << std::strerror(errno);
  const int32_t w = which;
  HB_CHECK(send(c, &w, sizeof(w), 0) == (ssize_t)sizeof(w)) << "descriptor request failed";
  char payload = 0;
  iovec io{&payload, 1};
  alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
  msghdr msg{};
  msg.msg_iov = &io; msg.msg_iovlen = 1;
  msg.msg_control = ctrl; msg.msg_controllen = sizeof(ctrl);
  const ssize_t n = recvmsg(c, &msg, 0);
  close(c);
  HB_CHECK(n == 1) << "no descriptor received from process " << pid;
  cmsghdr* cm = CMSG_FIRSTHDR(&msg);
  HB_CHECK(cm != nullptr && cm->cmsg_level == SOL_SOCKET && cm->cmsg_type == SCM_RIGHTS) << "malformed descriptor message";
  int fd = -1;
  std::memcpy(&fd, CMSG_DATA(cm), sizeof(int));
  return fd;
}

// a duplicate of descriptor `fd` (which: 0 memory handle, 1 multicast handle) of the peer described by (pid, seq)
int dup_from_process(int pid, int fd, int seq, int which) {
  if (pid == (int)getpid()) return dup(fd);
  const int pidfd = (int)syscall(SYS_pidfd_open, pid, 0);
  if (pidfd >= 0) {
    const int got = (int)syscall(SYS_pidfd_getfd, pidfd, fd, 0);
    close(pidfd);
    if (got >= 0) return got;
  }
  return recv_fd_over_socket(pid, seq, which);
}

CUmemAccessDesc rw_access(int device) {
  CUmemAccessDesc a;
  std::memset(&a, 0, sizeof(a));
  a.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  a.location.id = device;
  a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  return a;
}

void* map_handle(CUmemGenericAllocationHandle h, size_t size, size_t align, int device) {
  CUdeviceptr va = 0;
  DRV_OK(D().addressReserve(&va, size, align, 0, 0));
  DRV_OK(D().memMap(va, size, 0, h, 0));
  CUmemAccessDesc a = rw_access(device);
  DRV_OK(D().setAccess(va, size, &a, 1));
  return reinterpret_cast<void*>(va);
}

}  // namespace

bool SymmMem::multicast_supported() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  try {
    CUdevice cd;
    if (D().deviceGet(&cd, dev) != CUDA_SUCCESS) return false;
    int ok = 0;
    if (D().deviceGetAttribute(&ok, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cd) != CUDA_SUCCESS) return false;
    return ok != 0;
  } catch (const std::exception&) {
    return false;
  }
}

std::string SymmMem::alloc_vmm(const std::string& name, size_t bytes, int rank, int world) {
  HB_CHECK(world <= kMaxPeers) << "symmetric memory supports up to " << kMaxPeers << " ranks per node";
  HB_CHECK(!bufs_.count(name)) << "symmetric buffer " << name << " already exists";
  SymmBuffer b;
  b.name = name;
  b.bytes = (bytes + 255) / 256 * 256;
  b.rank = rank;
  b.world = world;
  b.vmm = true;
  int dev = 0;
  HB_CHECK(cudaGetDevice(&dev) == cudaSuccess) << "no CUDA device";
  HB_CHECK(cudaFree(0) == cudaSuccess) << "CUDA context";
  b.device = dev;
  const bool mc_ok = multicast_supported();

  CUmemAllocationProp prop;
  std::memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = dev;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0;
  DRV_OK(D().getGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  const size_t total = b.bytes + size_t(kFlagWords) * sizeof(uint32_t) * world;
  CUmulticastObjectProp mp;
  std::memset(&mp, 0, sizeof(mp));
  mp.numDevices = (unsigned)world;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  if (mc_ok) {
    mp.size = total;
    size_t mgran = 0;
    DRV_OK(D().mcGranularity(&mgran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    if (mgran > gran) gran = mgran;
  }
  b.map_bytes = (total + gran - 1) / gran * gran;
  b.granularity = gran;

  CUmemGenericAllocationHandle h;
  DRV_OK(D().memCreate(&h, b.map_bytes, &prop, 0));
  b.mem_handle = (unsigned long long)h;
  int fd = -1;
  DRV_OK(D().memExport(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  b.mem_fd = fd;
  b.local = map_handle(h, b.map_bytes, gran, dev);
  HB_CHECK(cudaMemset(b.local, 0, b.map_bytes) == cudaSuccess) << "cudaMemset of the symmetric buffer";
  HB_CHECK(cudaDeviceSynchronize() == cudaSuccess) << "sync";
  b.flags_local = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(b.local) + b.bytes);

  static std::atomic<int> sock_seq{0};
  VmmDesc d;
  std::memset(&d, 0, sizeof(d));
  d.pid = (int32_t)getpid();
  d.mem_fd = fd;
  d.mc_fd = -1;
  d.mc_ok = mc_ok ? 1 : 0;
  d.size = b.map_bytes;
  d.sock_seq = sock_seq.fetch_add(1);
  prctl(PR_SET_PTRACER, PR_SET_PTRACER_ANY, 0, 0, 0);      // let the peers use pidfd_getfd under Yama ptrace_scope = 1
  if (mc_ok && rank == 0 && world > 1) {
    mp.size = b.map_bytes;
    CUmemGenericAllocationHandle mc;
    DRV_OK(D().mcCreate(&mc, &mp));
    b.mc_handle = (unsigned long long)mc;
    int mfd = -1;
    DRV_OK(D().memExport(&mfd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    b.mc_fd = mfd;
    d.mc_fd = mfd;
  }
  {
    auto srv = std::make_unique<FdServer>();
    srv->fds[0] = b.mem_fd;
    srv->fds[1] = b.mc_fd;
    srv->start(d.sock_seq);
    servers()[name] = std::move(srv);
  }
  bufs_[name] = b;
  return std::string(reinterpret_cast<const char*>(&d), sizeof(d));
}

void SymmMem::open_vmm(const std::string& name, const std::vector<std::string>& descs) {
  SymmBuffer& b = buffer(name);
  HB_CHECK(b.vmm) << name << " is not a VMM symmetric buffer";
  HB_CHECK((int)descs.size() == b.world) << "need one descriptor per rank";
  std::vector<VmmDesc> ds(b.world);
  bool mc_all = b.world > 1;
  for (int r = 0; r < b.world; ++r) {
    HB_CHECK(descs[r].size() == sizeof(VmmDesc)) << "malformed descriptor from rank " << r;
    std::memcpy(&ds[r], descs[r].data(), sizeof(VmmDesc));
    HB_CHECK(ds[r].size == b.map_bytes) << "rank " << r << " allocated " << ds[r].size << " bytes, this rank " << b.map_bytes;
    mc_all = mc_all && ds[r].mc_ok != 0;
  }
  mc_all = mc_all && ds[0].mc_fd >= 0;
  void* flag_ptrs[kMaxPeers] = {nullptr};
  for (int r = 0; r < b.world; ++r) {
    if (r == b.rank) b.peer[r] = b.local;
    else {
      const int fd = dup_from_process(ds[r].pid, ds[r].mem_fd, ds[r].sock_seq, 0);
      CUmemGenericAllocationHandle h;
      DRV_OK(D().memImport(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
      close(fd);
      b.peer[r] = map_handle(h, b.map_bytes, b.granularity, b.device);
      b.peer_handles[r] = (unsigned long long)h;
    }
    flag_ptrs[r] = reinterpret_cast<char*>(b.peer[r]) + b.bytes;
  }
  HB_CHECK(cudaMalloc(&b.d_peer, sizeof(void*) * kMaxPeers) == cudaSuccess) << "cudaMalloc";
  HB_CHECK(cudaMalloc(&b.d_flags, sizeof(void*) * kMaxPeers) == cudaSuccess) << "cudaMalloc";
  HB_CHECK(cudaMemcpy(b.d_peer, b.peer, sizeof(void*) * kMaxPeers, cudaMemcpyHostToDevice) == cudaSuccess) << "cudaMemcpy";
  HB_CHECK(cudaMemcpy(b.d_flags, flag_ptrs, sizeof(void*) * kMaxPeers, cudaMemcpyHostToDevice) == cudaSuccess) << "cudaMemcpy";
  if (!mc_all) return;        // no NVLS on this system: the peer-pointer path still works
  CUmemGenericAllocationHandle mc;
  if (b.rank == 0) mc = (CUmemGenericAllocationHandle)b.mc_handle;
  else {
    const int fd = dup_from_process(ds[0].pid, ds[0].mc_fd, ds[0].sock_seq, 1);
    DRV_OK(D().memImport(&mc, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    close(fd);
    b.mc_handle = (unsigned long long)mc;
  }
  CUdevice cd;
  DRV_OK(D().deviceGet(&cd, b.device));
  DRV_OK(D().mcAddDevice(mc, cd));
  b.mc_joined = true;
}

void SymmMem::bind_multicast(const std::string& name) {
  SymmBuffer& b = buffer(name);
  if (!b.mc_joined) return;
  CUmemGenericAllocationHandle mc = (CUmemGenericAllocationHandle)b.mc_handle;
  DRV_OK(D().mcBindMem(mc, 0, (CUmemGenericAllocationHandle)b.mem_handle, 0, b.map_bytes, 0));
  b.mc = map_handle(mc, b.map_bytes, b.granularity, b.device);
}

void SymmMem::free_vmm(SymmBuffer& b) {
  auto sit = servers().find(b.name);
  if (sit != servers().end()) { sit->second->shutdown_server(); servers().erase(sit); }
  auto unmap = [&](void* p) {
    if (!p) return;
    D().memUnmap((CUdeviceptr)p, b.map_bytes);
    D().addressFree((CUdeviceptr)p, b.map_bytes);
  };
  unmap(b.mc);
  for (int r = 0; r < b.world; ++r) {
    if (r == b.rank) continue;
    unmap(b.peer[r]);
    if (b.peer_handles[r]) D().memRelease((CUmemGenericAllocationHandle)b.peer_handles[r]);
  }
  unmap(b.local);
  if (b.mc_handle) D().memRelease((CUmemGenericAllocationHandle)b.mc_handle);
  if (b.mem_handle) D().memRelease((CUmemGenericAllocationHandle)b.mem_handle);
  if (b.mem_fd >= 0) close(b.mem_fd);
  if (b.mc_fd >= 0) close(b.mc_fd);
  if (b.d_peer) cudaFree(b.d_peer);
  if (b.d_flags) cudaFree(b.d_flags);
}

}  // namespace hb
