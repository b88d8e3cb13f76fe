// Caching device-memory pool: per-stream free lists of size-binned blocks carved out of large segments, split on
// allocation and merged with free neighbours on release, event-based reuse across streams, statistics and limits.
// Backends: CUDA (cudaMalloc / cudaMallocAsync-style stream-ordered reuse) and host (pinned or pageable) -- the host
// backend also makes the allocator logic testable without a GPU.
// (capability parity: hetu/core/memory_pool.h (AllocDataSpace / BorrowDataSpace / FreeDataSpace /
//  MarkDataSpaceUsedByStream / WaitDataSpace / EmptyCache), hetu/impl/memory/CUDACachingMemoryPool.cu, CUDABFCMemoryPool,
//  CUDAStreamOrderedMemoryPool, CPUMemoryPool; env knobs HETU_MAX_SPLIT_SIZE_MB, HETU_MAX_INTERNAL_FRAGMENT_SIZE_MB,
//  HETU_PRE_ALLOCATE_SIZE_MB)
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

namespace hb {

struct PoolStats {
  size_t reserved = 0, allocated = 0, peak_reserved = 0, peak_allocated = 0;
  size_t num_alloc = 0, num_free = 0, num_segment_alloc = 0, num_split = 0, num_merge = 0, cache_hits = 0;
};

class MemoryBackend {
 public:
  virtual ~MemoryBackend() = default;
  virtual void* raw_alloc(size_t bytes) = 0;          // nullptr on failure
  virtual void raw_free(void* p) = 0;
  // cross-stream safety: record "block is in use on `stream`" and query / wait for it
  virtual uint64_t record_event(int64_t stream) = 0;
  virtual bool event_done(uint64_t ev) = 0;
  virtual void event_sync(uint64_t ev) = 0;
  virtual const char* name() const = 0;
};
std::unique_ptr<MemoryBackend> make_host_backend(bool pinned);
std::unique_ptr<MemoryBackend> make_cuda_backend(int device);
std::unique_ptr<MemoryBackend> make_shm_backend(const std::string& prefix);     // POSIX shared memory (AllocShareMemory)
bool shm_locate(MemoryBackend* backend, const void* p, std::string* shm_name, size_t* offset);

class CachingMemoryPool {
 public:
  struct Options {
    size_t small_block = 1 << 20;          // requests <= 1 MiB are served from 2 MiB segments
    size_t small_segment = 2 << 20;
    size_t large_segment_min = 20 << 20;   // larger requests get at least this much reserved
    size_t round_to = 512;
    size_t max_split_size = SIZE_MAX;      // blocks larger than this are never split (HETU_MAX_SPLIT_SIZE_MB)
    size_t max_internal_fragment = 8 << 20;   // do not hand out a cached block wasting more than this
    size_t pre_allocate = 0;               // reserve this much up front (HETU_PRE_ALLOCATE_SIZE_MB)
    size_t limit = SIZE_MAX;               // hard cap on reserved bytes
  };
  CachingMemoryPool(std::unique_ptr<MemoryBackend> backend, Options opt);
  explicit CachingMemoryPool(std::unique_ptr<MemoryBackend> backend) : CachingMemoryPool(std::move(backend), Options()) {}
  ~CachingMemoryPool();

  void* alloc(size_t bytes, int64_t stream = 0);                 // AllocDataSpace
  void free(void* p);                                            // FreeDataSpace
  void* borrow(void* p, size_t bytes);                           // BorrowDataSpace: track external memory (never freed by us)
  void mark_used_by_stream(void* p, int64_t stream);             // MarkDataSpaceUsedByStream
  void wait(void* p);                                            // WaitDataSpace: block until all marked streams are done
  size_t empty_cache();                                          // release every fully free segment; returns bytes released
  PoolStats stats() const;
  std::string summary() const;
  MemoryBackend* backend() { return backend_.get(); }
  static Options options_from_env();

 private:
  struct Block {
    char* ptr; size_t size; int64_t stream; bool in_use = false; bool borrowed = false;
    Block* prev = nullptr; Block* next = nullptr;     // neighbours inside the segment
    char* segment; size_t segment_size;
    std::vector<uint64_t> events;                      // other streams that touched the block
  };
  struct Cmp { bool operator()(const Block* a, const Block* b) const { return a->size != b->size ? a->size < b->size : a->ptr < b->ptr; } };
  using FreeList = std::set<Block*, Cmp>;
  FreeList& list_for(int64_t stream, bool small) { return free_[{stream, small}]; }
  Block* find_free(size_t size, int64_t stream, bool small);
  Block* new_segment(size_t size, int64_t stream, bool small);
  void release_block(Block* b);
  void process_pending();
  size_t round(size_t n) const { return (n + opt_.round_to - 1) / opt_.round_to * opt_.round_to; }

  std::unique_ptr<MemoryBackend> backend_;
  Options opt_;
  mutable std::mutex mu_;
  std::map<std::pair<int64_t, bool>, FreeList> free_;
  std::unordered_map<void*, Block*> live_;
  std::vector<Block*> pending_;       // freed but still referenced by another stream's events
  PoolStats st_;
};

// per-device registry of pools: "cuda:<i>", "cpu", "pinned", "shm" (ref: GetMemoryPool / RegisterMemoryPool, MemoryManager)
class MemoryPoolRegistry {
 public:
  static MemoryPoolRegistry& instance();
  std::shared_ptr<CachingMemoryPool> get(const std::string& device);
  std::vector<std::string> devices() const;
  size_t empty_all_caches();

 private:
  mutable std::mutex mu_;
  std::map<std::string, std::shared_ptr<CachingMemoryPool>> pools_;
};

}  // namespace hb
