#include "memory_pool.h"

#include <map>
#include <unistd.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <sstream>

#include "../core/base.h"

namespace hb {

// ------------------------------------------------------------------ backends
namespace {
class HostBackend : public MemoryBackend {
 public:
  explicit HostBackend(bool pinned) : pinned_(pinned) {}
  void* raw_alloc(size_t bytes) override {
    void* p = nullptr;
    if (pinned_) {
      if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
      return p;
    }
    if (posix_memalign(&p, 4096, bytes) != 0) return nullptr;
    return p;
  }
  void raw_free(void* p) override {
    if (pinned_) cudaFreeHost(p);
    else std::free(p);
  }
  uint64_t record_event(int64_t) override { return 0; }
  bool event_done(uint64_t) override { return true; }
  void event_sync(uint64_t) override {}
  const char* name() const override { return pinned_ ? "host-pinned" : "host"; }

 private:
  bool pinned_;
};

class CudaBackend : public MemoryBackend {
 public:
  explicit CudaBackend(int device) : device_(device) {}
  void* raw_alloc(size_t bytes) override {
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device_);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    cudaSetDevice(prev);
    if (e != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
  }
  void raw_free(void* p) override { cudaFree(p); }
  uint64_t record_event(int64_t stream) override {
    cudaEvent_t ev;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return 0;
    cudaEventRecord(ev, reinterpret_cast<cudaStream_t>(stream));
    const uint64_t id = ++next_;
    events_[id] = ev;
    return id;
  }
  bool event_done(uint64_t id) override {
    auto it = events_.find(id);
    if (it == events_.end()) return true;
    if (cudaEventQuery(it->second) == cudaErrorNotReady) { cudaGetLastError(); return false; }
    cudaEventDestroy(it->second);
    events_.erase(it);
    return true;
  }
  void event_sync(uint64_t id) override {
    auto it = events_.find(id);
    if (it == events_.end()) return;
    cudaEventSynchronize(it->second);
    cudaEventDestroy(it->second);
    events_.erase(it);
  }
  const char* name() const override { return "cuda"; }

 private:
  int device_;
  uint64_t next_ = 0;
  std::unordered_map<uint64_t, cudaEvent_t> events_;
};
// POSIX shared-memory segments (shm_open + mmap): host buffers another process can map by name -- the asynchronous
// checkpoint writer reads the parameter snapshot from here while training continues
// (ref: CPUMemoryPool::AllocShareMemory, hetu/impl/memory/CPUMemoryPool.cc)
class ShmBackend : public MemoryBackend {
 public:
  explicit ShmBackend(std::string prefix) : prefix_(std::move(prefix)) {}
  ~ShmBackend() override {
    for (auto& kv : segs_) { munmap(kv.first, kv.second.second); shm_unlink(kv.second.first.c_str()); }
  }
  void* raw_alloc(size_t bytes) override {
    const std::string name = "/" + prefix_ + "_" + std::to_string(getpid()) + "_" + std::to_string(++next_);
    int fd = shm_open(name.c_str(), O_CREAT | O_RDWR | O_EXCL, 0600);
    if (fd < 0) return nullptr;
    if (ftruncate(fd, (off_t)bytes) != 0) { close(fd); shm_unlink(name.c_str()); return nullptr; }
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { shm_unlink(name.c_str()); return nullptr; }
    segs_[p] = {name, bytes};
    return p;
  }
  void raw_free(void* p) override {
    auto it = segs_.find(p);
    if (it == segs_.end()) return;
    munmap(p, it->second.second);
    shm_unlink(it->second.first.c_str());
    segs_.erase(it);
  }
  uint64_t record_event(int64_t) override { return 0; }
  bool event_done(uint64_t) override { return true; }
  void event_sync(uint64_t) override {}
  const char* name() const override { return "shm"; }
  // segment that contains p -> (shm name, offset) so that a peer process can map the same bytes
  bool locate(const void* p, std::string* shm_name, size_t* offset) const {
    for (auto& kv : segs_) {
      const char* base = static_cast<const char*>(kv.first);
      if (p >= base && p < base + kv.second.second) { *shm_name = kv.second.first; *offset = (size_t)(static_cast<const char*>(p) - base); return true; }
    }
    return false;
  }

 private:
  std::string prefix_;
  uint64_t next_ = 0;
  std::map<void*, std::pair<std::string, size_t>> segs_;
};
}  // namespace

std::unique_ptr<MemoryBackend> make_shm_backend(const std::string& prefix) { return std::make_unique<ShmBackend>(prefix); }
bool shm_locate(MemoryBackend* backend, const void* p, std::string* name, size_t* offset) {
  auto* b = dynamic_cast<ShmBackend*>(backend);
  return b != nullptr && b->locate(p, name, offset);
}

// ------------------------------------------------------------------ per-device registry (MemoryManager role)
std::shared_ptr<CachingMemoryPool> MemoryPoolRegistry::get(const std::string& device) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = pools_.find(device);
  if (it != pools_.end()) return it->second;
  std::unique_ptr<MemoryBackend> be;
  if (device.rfind("cuda", 0) == 0) {
    const size_t c = device.find(':');
    be = make_cuda_backend(c == std::string::npos ? 0 : std::stoi(device.substr(c + 1)));
  } else if (device == "pinned") be = make_host_backend(true);
  else if (device == "shm") be = make_shm_backend("hetu_b200");
  else be = make_host_backend(false);
  auto pool = std::make_shared<CachingMemoryPool>(std::move(be), CachingMemoryPool::options_from_env());
  pools_[device] = pool;
  return pool;
}
std::vector<std::string> MemoryPoolRegistry::devices() const {
  std::lock_guard<std::mutex> lk(mu_);
  std::vector<std::string> out;
  for (auto& kv : pools_) out.push_back(kv.first);
  return out;
}
size_t MemoryPoolRegistry::empty_all_caches() {
  std::lock_guard<std::mutex> lk(mu_);
  size_t n = 0;
  for (auto& kv : pools_) n += kv.second->empty_cache();
  return n;
}
MemoryPoolRegistry& MemoryPoolRegistry::instance() {
  static MemoryPoolRegistry* r = new MemoryPoolRegistry();
  return *r;
}

std::unique_ptr<MemoryBackend> make_host_backend(bool pinned) { return std::make_unique<HostBackend>(pinned); }
std::unique_ptr<MemoryBackend> make_cuda_backend(int device) { return std::make_unique<CudaBackend>(device); }

// ------------------------------------------------------------------ pool
CachingMemoryPool::Options CachingMemoryPool::options_from_env() {
  Options o;
  const int64_t split = env_int("HETU_MAX_SPLIT_SIZE_MB", 0);
  if (split > 0) o.max_split_size = (size_t)split << 20;
  const int64_t frag = env_int("HETU_MAX_INTERNAL_FRAGMENT_SIZE_MB", 0);
  if (frag > 0) o.max_internal_fragment = (size_t)frag << 20;
  const int64_t pre = env_int("HETU_PRE_ALLOCATE_SIZE_MB", 0);
  if (pre > 0) o.pre_allocate = (size_t)pre << 20;
  return o;
}

CachingMemoryPool::CachingMemoryPool(std::unique_ptr<MemoryBackend> backend, Options opt) : backend_(std::move(backend)), opt_(opt) {
  if (opt_.pre_allocate > 0) {
    std::lock_guard<std::mutex> lk(mu_);
    Block* b = new_segment(opt_.pre_allocate, 0, false);
    if (b != nullptr) list_for(0, false).insert(b);
  }
}

CachingMemoryPool::~CachingMemoryPool() {
  std::set<char*> segs;
  for (auto& kv : free_) for (Block* b : kv.second) { segs.insert(b->segment); }
  for (auto& kv : live_) if (!kv.second->borrowed) segs.insert(kv.second->segment);
  for (Block* b : pending_) segs.insert(b->segment);
  for (char* s : segs) backend_->raw_free(s);
  for (auto& kv : free_) for (Block* b : kv.second) delete b;
  for (auto& kv : live_) delete kv.second;
  for (Block* b : pending_) delete b;
}

CachingMemoryPool::Block* CachingMemoryPool::new_segment(size_t size, int64_t stream, bool small) {
  size_t seg = small ? opt_.small_segment : std::max(round(size), size < opt_.large_segment_min ? opt_.large_segment_min : round(size));
  if (st_.reserved + seg > opt_.limit) {
    seg = round(size);
    if (st_.reserved + seg > opt_.limit) return nullptr;
  }
  void* p = backend_->raw_alloc(seg);
  if (p == nullptr && seg > round(size)) { seg = round(size); p = backend_->raw_alloc(seg); }
  if (p == nullptr) return nullptr;
  st_.reserved += seg;
  st_.peak_reserved = std::max(st_.peak_reserved, st_.reserved);
  ++st_.num_segment_alloc;
  Block* b = new Block{static_cast<char*>(p), seg, stream};
  b->segment = b->ptr;
  b->segment_size = seg;
  return b;
}

CachingMemoryPool::Block* CachingMemoryPool::find_free(size_t size, int64_t stream, bool small) {
  FreeList& fl = list_for(stream, small);
  Block key{nullptr, size, stream};
  auto it = fl.lower_bound(&key);
  if (it == fl.end()) return nullptr;
  Block* b = *it;
  if (b->size - size > opt_.max_internal_fragment && b->size > opt_.max_split_size) return nullptr;   // would waste too much, cannot split
  fl.erase(it);
  ++st_.cache_hits;
  return b;
}

void* CachingMemoryPool::alloc(size_t bytes, int64_t stream) {
  if (bytes == 0) return nullptr;
  std::lock_guard<std::mutex> lk(mu_);
  process_pending();
  const size_t size = round(bytes);
  const bool small = size <= opt_.small_block;
  Block* b = find_free(size, stream, small);
  if (b == nullptr) b = new_segment(size, stream, small);
  if (b == nullptr) {
    // out of memory: give every cached segment back and retry once (the reference's EmptyCache-on-OOM)
    mu_.unlock();
    empty_cache();
    mu_.lock();
    b = new_segment(size, stream, small);
    HB_CHECK(b != nullptr) << "memory pool (" << backend_->name() << "): out of memory allocating " << bytes << " bytes; " << summary();
  }
  // split when the remainder is worth keeping
  const size_t remain = b->size - size;
  if (remain >= opt_.round_to && b->size <= opt_.max_split_size && (small || remain > opt_.small_block)) {
    Block* r = new Block{b->ptr + size, remain, b->stream};
    r->segment = b->segment; r->segment_size = b->segment_size;
    r->prev = b; r->next = b->next;
    if (b->next) b->next->prev = r;
    b->next = r;
    b->size = size;
    list_for(r->stream, small).insert(r);
    ++st_.num_split;
  }
  b->in_use = true;
  live_[b->ptr] = b;
  st_.allocated += b->size;
  st_.peak_allocated = std::max(st_.peak_allocated, st_.allocated);
  ++st_.num_alloc;
  return b->ptr;
}

void* CachingMemoryPool::borrow(void* p, size_t bytes) {
  std::lock_guard<std::mutex> lk(mu_);
  Block* b = new Block{static_cast<char*>(p), bytes, 0};
  b->in_use = true; b->borrowed = true; b->segment = b->ptr; b->segment_size = bytes;
  live_[p] = b;
  return p;
}

void CachingMemoryPool::release_block(Block* b) {
  const bool small = b->segment_size == opt_.small_segment;
  b->in_use = false;
  // merge with free neighbours of the same stream
  auto try_merge = [&](Block* dst, Block* src) {
    if (src == nullptr || src->in_use || src->stream != dst->stream || !src->events.empty()) return false;
    FreeList& fl = list_for(src->stream, small);
    auto it = fl.find(src);
    if (it == fl.end()) return false;
    fl.erase(it);
    if (src == dst->prev) {
      dst->ptr = src->ptr;
      dst->prev = src->prev;
      if (src->prev) src->prev->next = dst;
    } else {
      dst->next = src->next;
      if (src->next) src->next->prev = dst;
    }
    dst->size += src->size;
    delete src;
    ++st_.num_merge;
    return true;
  };
  try_merge(b, b->prev);
  try_merge(b, b->next);
  list_for(b->stream, small).insert(b);
}

void CachingMemoryPool::free(void* p) {
  if (p == nullptr) return;
  std::lock_guard<std::mutex> lk(mu_);
  auto it = live_.find(p);
  HB_CHECK(it != live_.end()) << "memory pool: free of an unknown pointer";
  Block* b = it->second;
  live_.erase(it);
  ++st_.num_free;
  if (b->borrowed) { delete b; return; }
  st_.allocated -= b->size;
  if (!b->events.empty()) { pending_.push_back(b); return; }   // another stream may still be reading it
  release_block(b);
}

void CachingMemoryPool::mark_used_by_stream(void* p, int64_t stream) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = live_.find(p);
  if (it == live_.end() || it->second->stream == stream) return;
  const uint64_t ev = backend_->record_event(stream);
  if (ev != 0) it->second->events.push_back(ev);
}

void CachingMemoryPool::wait(void* p) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = live_.find(p);
  if (it == live_.end()) return;
  for (uint64_t ev : it->second->events) backend_->event_sync(ev);
  it->second->events.clear();
}

void CachingMemoryPool::process_pending() {
  for (size_t i = 0; i < pending_.size();) {
    Block* b = pending_[i];
    b->events.erase(std::remove_if(b->events.begin(), b->events.end(), [&](uint64_t ev) { return backend_->event_done(ev); }),
                    b->events.end());
    if (b->events.empty()) {
      pending_.erase(pending_.begin() + (long)i);
      release_block(b);
    } else ++i;
  }
}

size_t CachingMemoryPool::empty_cache() {
  std::lock_guard<std::mutex> lk(mu_);
  process_pending();
  size_t released = 0;
  for (auto& kv : free_) {
    for (auto it = kv.second.begin(); it != kv.second.end();) {
      Block* b = *it;
      if (b->prev == nullptr && b->next == nullptr && b->size == b->segment_size) {   // a whole free segment
        backend_->raw_free(b->segment);
        released += b->size;
        st_.reserved -= b->size;
        it = kv.second.erase(it);
        delete b;
      } else ++it;
    }
  }
  return released;
}

PoolStats CachingMemoryPool::stats() const {
  std::lock_guard<std::mutex> lk(mu_);
  return st_;
}

std::string CachingMemoryPool::summary() const {
  std::ostringstream os;
  os << backend_->name() << " pool: reserved " << (st_.reserved >> 20) << " MiB (peak " << (st_.peak_reserved >> 20) << "), allocated "
     << (st_.allocated >> 20) << " MiB (peak " << (st_.peak_allocated >> 20) << "), " << st_.num_alloc << " allocs, " << st_.cache_hits
     << " cache hits, " << st_.num_split << " splits, " << st_.num_merge << " merges, " << st_.num_segment_alloc << " segments";
  return os.str();
}

}  // namespace hb
