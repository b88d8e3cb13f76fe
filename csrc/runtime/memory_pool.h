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
  size_t cross_stream_syncs = 0;
};
std::unique_ptr<MemoryBackend> make_host_backend(bool pinned);
std::unique_ptr<MemoryBackend> make_cuda_backend(int device);
std::unique_ptr<MemoryBackend> make_shm_backend(const std::string& prefix);     // POSIX shared memory (AllocShareMemory)
bool shm_locate(MemoryBackend* backend, const void* p, std::string* shm_name, size_t* offset);

// What the tensor path needs from a pool; three implementations: CachingMemoryPool (per-stream size-binned free lists over
// many segments), BFCMemoryPool (best-fit with coalescing over a few large, doubling regions) and StreamOrderedMemoryPool
// (the driver's stream-ordered allocator, cudaMallocAsync on a cudaMemPool with a release threshold).
// `HETU_MEMORY_POOL = caching | bfc | stream_ordered` picks the one that backs CUDA tensors (HETU_NATIVE_ALLOCATOR=1).
class DeviceAllocator {
 public:
  virtual ~DeviceAllocator() = default;
  virtual void* alloc(size_t bytes, int64_t stream = 0) = 0;
  virtual void free(void* p) = 0;
  virtual void mark_used_by_stream(void* p, int64_t stream) = 0;
  virtual void wait(void* p) = 0;
  virtual size_t empty_cache() = 0;
  virtual PoolStats stats() const = 0;
  virtual std::string summary() const = 0;
  virtual const char* kind() const = 0;
};

class CachingMemoryPool : public DeviceAllocator {
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

  void* alloc(size_t bytes, int64_t stream = 0) override;        // AllocDataSpace
  void free(void* p) override;                                   // FreeDataSpace
  void* borrow(void* p, size_t bytes);                           // BorrowDataSpace: track external memory (never freed by us)
  void mark_used_by_stream(void* p, int64_t stream) override;    // MarkDataSpaceUsedByStream
  void wait(void* p) override;                                   // WaitDataSpace: block until all marked streams are done
  size_t empty_cache() override;                                 // release every fully free segment; returns bytes released
  PoolStats stats() const override;
  std::string summary() const override;
  const char* kind() const override { return "caching"; }
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

// Best-fit-with-coalescing pool (ref: hetu/impl/memory/CUDABFCMemoryPool -- capability parity, own design): memory is reserved
// in a few large regions (each new one twice the previous), free chunks sit in power-of-two size bins ordered by (size,
// address), an allocation takes the smallest fitting chunk of the lowest non-empty bin and splits it, a release coalesces
// with both address neighbours.  Chunks are not partitioned by stream: a free chunk remembers which streams last used its
// bytes; a request prefers chunks of its own stream and, when it takes one that another stream used, first waits for that
// stream's queued work (rare on the step path, where almost every allocation is made on the compute stream).
class BFCMemoryPool : public DeviceAllocator {
 public:
  struct Options {
    size_t min_chunk = 256;                  // allocation granularity
    size_t initial_region = 64 << 20;        // first region (HETU_PRE_ALLOCATE_SIZE_MB overrides)
    size_t max_internal_fragment = 128 << 20;   // a fitting chunk is split when it wastes more than this (or is >= 2x the request)
    size_t limit = SIZE_MAX;
  };
  BFCMemoryPool(std::unique_ptr<MemoryBackend> backend, Options opt);
  explicit BFCMemoryPool(std::unique_ptr<MemoryBackend> backend) : BFCMemoryPool(std::move(backend), Options()) {}
  ~BFCMemoryPool() override;
  void* alloc(size_t bytes, int64_t stream = 0) override;
  void free(void* p) override;
  void mark_used_by_stream(void* p, int64_t stream) override;
  void wait(void* p) override;
  size_t empty_cache() override;
  PoolStats stats() const override;
  std::string summary() const override;
  const char* kind() const override { return "bfc"; }
  // diagnostics: number of regions, free chunks per bin, largest free chunk, external fragmentation = 1 - largest_free / total_free
  size_t num_regions() const;
  std::vector<size_t> bin_occupancy() const;
  size_t largest_free_chunk() const;
  double fragmentation() const;
  static constexpr int kNumBins = 24;

 private:
  struct Chunk {
    char* ptr; size_t size; size_t requested = 0; bool in_use = false;
    Chunk* prev = nullptr; Chunk* next = nullptr;        // address neighbours inside the region
    int region; int64_t stream = 0;          // owner while in use
    std::vector<int64_t> used_by;             // free chunk: streams whose queued work may still touch these bytes
    std::vector<uint64_t> events;             // in use: foreign readers (mark_used_by_stream)
  };
  struct Cmp { bool operator()(const Chunk* a, const Chunk* b) const { return a->size != b->size ? a->size < b->size : a->ptr < b->ptr; } };
  struct Region { char* base; size_t size; };
  int bin_of(size_t size) const;
  static bool only_stream(const Chunk* c, int64_t stream);
  Chunk* take(size_t rounded, int64_t stream);
  bool extend(size_t rounded);
  void insert_free(Chunk* c);
  void remove_free(Chunk* c);
  void coalesce_and_insert(Chunk* c);
  void process_pending();

  std::unique_ptr<MemoryBackend> backend_;
  Options opt_;
  mutable std::mutex mu_;
  std::vector<Region> regions_;
  std::vector<std::set<Chunk*, Cmp>> bins_;
  std::unordered_map<void*, Chunk*> live_;
  std::vector<Chunk*> pending_;
  size_t next_region_ = 0;
  PoolStats st_;
};

// The driver's stream-ordered allocator behind the same interface (ref: hetu/impl/memory/CUDAStreamOrderedMemoryPool):
// cudaMallocAsync / cudaFreeAsync on the device's default cudaMemPool with a release threshold, so freed memory stays cached
// in the driver pool and reuse is ordered by the stream.  Cross-stream use is made safe by making the freeing stream wait
// for an event of every stream that was marked.  Without a CUDA device the class degrades to plain host allocations with the
// same bookkeeping (keeps the accounting testable).
class StreamOrderedMemoryPool : public DeviceAllocator {
 public:
  explicit StreamOrderedMemoryPool(int device, size_t release_threshold = SIZE_MAX);
  ~StreamOrderedMemoryPool() override;
  void* alloc(size_t bytes, int64_t stream = 0) override;
  void free(void* p) override;
  void mark_used_by_stream(void* p, int64_t stream) override;
  void wait(void* p) override;
  size_t empty_cache() override;            // cudaMemPoolTrimTo(0)
  PoolStats stats() const override;
  std::string summary() const override;
  const char* kind() const override { return "stream_ordered"; }
  bool on_device() const { return cuda_; }

 private:
  struct Rec { size_t size; int64_t stream; std::vector<int64_t> other_streams; };
  int device_;
  bool cuda_ = false;
  void* mempool_ = nullptr;                 // cudaMemPool_t
  mutable std::mutex mu_;
  std::unordered_map<void*, Rec> live_;
  PoolStats st_;
};

// per-device registry of pools: "cuda:<i>", "cpu", "pinned", "shm" (ref: GetMemoryPool / RegisterMemoryPool, MemoryManager)
class MemoryPoolRegistry {
 public:
  static MemoryPoolRegistry& instance();
  std::shared_ptr<CachingMemoryPool> get(const std::string& device);
  // the allocator that backs CUDA tensors of `device` ("cuda:<i>"): kind from HETU_MEMORY_POOL (default caching)
  std::shared_ptr<DeviceAllocator> tensor_allocator(const std::string& device);
  std::vector<std::string> devices() const;
  size_t empty_all_caches();

 private:
  mutable std::mutex mu_;
  std::map<std::string, std::shared_ptr<CachingMemoryPool>> pools_;
  std::map<std::string, std::shared_ptr<DeviceAllocator>> tensor_allocs_;
};

}  // namespace hb
