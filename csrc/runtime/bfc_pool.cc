// BFCMemoryPool and StreamOrderedMemoryPool (see memory_pool.h).
#include <cuda_runtime.h>

#include <algorithm>
#include <sstream>

#include "../core/base.h"
#include "memory_pool.h"

namespace hb {

// ------------------------------------------------------------------ best fit with coalescing
BFCMemoryPool::BFCMemoryPool(std::unique_ptr<MemoryBackend> backend, Options opt)
    : backend_(std::move(backend)), opt_(opt), bins_(kNumBins) {
  next_region_ = std::max(opt_.initial_region, opt_.min_chunk);
}

BFCMemoryPool::~BFCMemoryPool() {
  for (auto& b : bins_)
    for (Chunk* c : b) delete c;
  for (auto& kv : live_) delete kv.second;
  for (Chunk* c : pending_) delete c;
  for (auto& r : regions_) backend_->raw_free(r.base);
}

int BFCMemoryPool::bin_of(size_t size) const {
  size_t v = std::max<size_t>(size / opt_.min_chunk, 1);
  int b = 0;
  while (v > 1 && b < kNumBins - 1) { v >>= 1; ++b; }
  return b;
}

void BFCMemoryPool::insert_free(Chunk* c) { bins_[bin_of(c->size)].insert(c); }
void BFCMemoryPool::remove_free(Chunk* c) { bins_[bin_of(c->size)].erase(c); }

bool BFCMemoryPool::only_stream(const Chunk* c, int64_t stream) {
  for (int64_t s : c->used_by)
    if (s != stream) return false;
  return true;
}

BFCMemoryPool::Chunk* BFCMemoryPool::take(size_t rounded, int64_t stream) {
  // pass 0: best fit among chunks this stream (or nobody) used last; pass 1: any chunk, after waiting for its last users
  for (int pass = 0; pass < 2; ++pass) {
    for (int b = bin_of(rounded); b < kNumBins; ++b) {
      auto& bin = bins_[b];
      Chunk probe{nullptr, rounded};
      for (auto it = bin.lower_bound(&probe); it != bin.end(); ++it) {
        Chunk* c = *it;
        if (pass == 0 && !only_stream(c, stream)) continue;
        bin.erase(it);
        for (int64_t s : c->used_by) {
          if (s == stream) continue;
          const uint64_t ev = backend_->record_event(s);     // everything that stream queued so far covers its use of the bytes
          if (ev != 0) { backend_->event_sync(ev); ++backend_->cross_stream_syncs; }
        }
        // what is left is this stream's own earlier use (ordered by the stream itself); a split remainder inherits it
        c->used_by.erase(std::remove_if(c->used_by.begin(), c->used_by.end(), [&](int64_t s) { return s != stream; }), c->used_by.end());
        return c;
      }
    }
  }
  return nullptr;
}

bool BFCMemoryPool::extend(size_t rounded) {
  size_t want = next_region_;
  while (want < rounded) want *= 2;
  const size_t room = opt_.limit > st_.reserved ? opt_.limit - st_.reserved : 0;
  want = std::min(want, room);
  char* base = nullptr;
  while (want >= rounded) {
    base = static_cast<char*>(backend_->raw_alloc(want));
    if (base != nullptr) break;
    if (want == rounded) break;
    want = std::max(rounded, want / 2);            // the device is tighter than the growth schedule: back off
  }
  if (base == nullptr) return false;
  regions_.push_back({base, want});
  next_region_ = want * 2;
  st_.reserved += want;
  st_.peak_reserved = std::max(st_.peak_reserved, st_.reserved);
  ++st_.num_segment_alloc;
  Chunk* c = new Chunk{base, want};
  c->region = (int)regions_.size() - 1;
  insert_free(c);
  return true;
}

void* BFCMemoryPool::alloc(size_t bytes, int64_t stream) {
  if (bytes == 0) return nullptr;
  std::lock_guard<std::mutex> lk(mu_);
  process_pending();
  const size_t rounded = (bytes + opt_.min_chunk - 1) / opt_.min_chunk * opt_.min_chunk;
  Chunk* c = take(rounded, stream);
  if (c != nullptr) ++st_.cache_hits;
  if (c == nullptr) {
    if (!extend(rounded)) return nullptr;
    c = take(rounded, stream);
    if (c == nullptr) return nullptr;
  }
  const size_t rest = c->size - rounded;
  if (rest >= opt_.min_chunk && (c->size >= 2 * rounded || rest >= opt_.max_internal_fragment)) {
    Chunk* r = new Chunk{c->ptr + rounded, rest};
    r->region = c->region;
    r->used_by = c->used_by;
    r->prev = c; r->next = c->next;
    if (c->next) c->next->prev = r;
    c->next = r;
    c->size = rounded;
    insert_free(r);
    ++st_.num_split;
  }
  c->in_use = true; c->requested = bytes; c->stream = stream; c->events.clear();
  live_[c->ptr] = c;
  st_.allocated += c->size;
  st_.peak_allocated = std::max(st_.peak_allocated, st_.allocated);
  ++st_.num_alloc;
  return c->ptr;
}

void BFCMemoryPool::coalesce_and_insert(Chunk* c) {
  c->in_use = false;
  if (std::find(c->used_by.begin(), c->used_by.end(), c->stream) == c->used_by.end()) c->used_by.push_back(c->stream);
  // free address neighbours are absorbed; the merged chunk remembers every stream that used any part of it
  auto absorb = [&](Chunk* n) {
    if (n == nullptr || n->in_use || !n->events.empty()) return;      // in use, or parked until a foreign reader finishes
    remove_free(n);
    if (n == c->prev) {
      c->ptr = n->ptr;
      c->prev = n->prev;
      if (n->prev) n->prev->next = c;
    } else {
      c->next = n->next;
      if (n->next) n->next->prev = c;
    }
    c->size += n->size;
    for (int64_t s : n->used_by)
      if (std::find(c->used_by.begin(), c->used_by.end(), s) == c->used_by.end()) c->used_by.push_back(s);
    delete n;
    ++st_.num_merge;
  };
  absorb(c->prev);
  absorb(c->next);
  insert_free(c);
}

void BFCMemoryPool::free(void* p) {
  if (p == nullptr) return;
  std::lock_guard<std::mutex> lk(mu_);
  auto it = live_.find(p);
  HB_CHECK(it != live_.end()) << "bfc pool: free of an unknown pointer";
  Chunk* c = it->second;
  live_.erase(it);
  ++st_.num_free;
  st_.allocated -= c->size;
  if (!c->events.empty()) { c->in_use = false; pending_.push_back(c); return; }   // another stream may still be reading it
  coalesce_and_insert(c);
}

void BFCMemoryPool::mark_used_by_stream(void* p, int64_t stream) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = live_.find(p);
  if (it == live_.end() || it->second->stream == stream) return;
  const uint64_t ev = backend_->record_event(stream);
  if (ev != 0) it->second->events.push_back(ev);
}

void BFCMemoryPool::wait(void* p) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = live_.find(p);
  if (it == live_.end()) return;
  for (uint64_t ev : it->second->events) backend_->event_sync(ev);
  it->second->events.clear();
}

void BFCMemoryPool::process_pending() {
  for (size_t i = 0; i < pending_.size();) {
    Chunk* c = pending_[i];
    c->events.erase(std::remove_if(c->events.begin(), c->events.end(), [&](uint64_t ev) { return backend_->event_done(ev); }), c->events.end());
    if (c->events.empty()) {
      pending_.erase(pending_.begin() + (long)i);
      coalesce_and_insert(c);
    } else ++i;
  }
}

size_t BFCMemoryPool::empty_cache() {
  std::lock_guard<std::mutex> lk(mu_);
  process_pending();
  size_t released = 0;
  for (size_t r = 0; r < regions_.size(); ++r) {
    if (regions_[r].base == nullptr) continue;
    // a region can go back to the device when one free chunk covers it
    Chunk* whole = nullptr;
    for (auto& b : bins_)
      for (Chunk* c : b)
        if (c->ptr == regions_[r].base && c->size == regions_[r].size && c->prev == nullptr && c->next == nullptr) whole = c;
    if (whole == nullptr) continue;
    for (int64_t st : whole->used_by) {             // the device may still be working on these bytes
      const uint64_t ev = backend_->record_event(st);
      if (ev != 0) backend_->event_sync(ev);
    }
    remove_free(whole);
    backend_->raw_free(regions_[r].base);
    released += regions_[r].size;
    st_.reserved -= regions_[r].size;
    regions_[r].base = nullptr;
    regions_[r].size = 0;
    delete whole;
  }
  if (st_.reserved == 0) next_region_ = std::max(opt_.initial_region, opt_.min_chunk);
  return released;
}

PoolStats BFCMemoryPool::stats() const {
  std::lock_guard<std::mutex> lk(mu_);
  return st_;
}
size_t BFCMemoryPool::num_regions() const {
  std::lock_guard<std::mutex> lk(mu_);
  size_t n = 0;
  for (auto& r : regions_) n += r.base != nullptr;
  return n;
}
std::vector<size_t> BFCMemoryPool::bin_occupancy() const {
  std::lock_guard<std::mutex> lk(mu_);
  std::vector<size_t> v;
  for (auto& b : bins_) v.push_back(b.size());
  return v;
}
size_t BFCMemoryPool::largest_free_chunk() const {
  std::lock_guard<std::mutex> lk(mu_);
  size_t m = 0;
  for (auto& b : bins_)
    if (!b.empty()) m = std::max(m, (*b.rbegin())->size);
  return m;
}
double BFCMemoryPool::fragmentation() const {
  std::lock_guard<std::mutex> lk(mu_);
  size_t total = 0, largest = 0;
  for (auto& b : bins_)
    for (Chunk* c : b) { total += c->size; largest = std::max(largest, c->size); }
  return total == 0 ? 0.0 : 1.0 - double(largest) / double(total);
}
std::string BFCMemoryPool::summary() const {
  std::ostringstream os;
  os << backend_->name() << " bfc pool: " << num_regions() << " regions, reserved " << (st_.reserved >> 20) << " MiB (peak "
     << (st_.peak_reserved >> 20) << "), allocated " << (st_.allocated >> 20) << " MiB (peak " << (st_.peak_allocated >> 20) << "), "
     << st_.num_alloc << " allocs, " << st_.cache_hits << " served from free chunks, " << st_.num_split << " splits, " << st_.num_merge
     << " merges, largest free chunk " << (largest_free_chunk() >> 20) << " MiB, fragmentation " << fragmentation();
  return os.str();
}

// ------------------------------------------------------------------ stream ordered
StreamOrderedMemoryPool::StreamOrderedMemoryPool(int device, size_t release_threshold) : device_(device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) == cudaSuccess && device < n) {
    cudaMemPool_t mp = nullptr;
    int supported = 0;
    cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, device);
    if (supported && cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess) {
      uint64_t thr = release_threshold == SIZE_MAX ? UINT64_MAX : (uint64_t)release_threshold;
      cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &thr);
      mempool_ = mp;
      cuda_ = true;
    }
  }
  cudaGetLastError();
}
StreamOrderedMemoryPool::~StreamOrderedMemoryPool() {
  for (auto& kv : live_) {
    if (cuda_) cudaFreeAsync(kv.first, reinterpret_cast<cudaStream_t>(kv.second.stream));
    else std::free(kv.first);
  }
}
void* StreamOrderedMemoryPool::alloc(size_t bytes, int64_t stream) {
  if (bytes == 0) return nullptr;
  void* p = nullptr;
  if (cuda_) {
    int prev = 0;
    cudaGetDevice(&prev);
    if (prev != device_) cudaSetDevice(device_);
    cudaError_t e = cudaMallocFromPoolAsync(&p, bytes, static_cast<cudaMemPool_t>(mempool_), reinterpret_cast<cudaStream_t>(stream));
    if (prev != device_) cudaSetDevice(prev);
    if (e != cudaSuccess) { cudaGetLastError(); return nullptr; }
  } else if (posix_memalign(&p, 256, bytes) != 0) {
    return nullptr;
  }
  std::lock_guard<std::mutex> lk(mu_);
  live_[p] = Rec{bytes, stream, {}};
  st_.allocated += bytes;
  st_.peak_allocated = std::max(st_.peak_allocated, st_.allocated);
  ++st_.num_alloc;
  return p;
}
void StreamOrderedMemoryPool::free(void* p) {
  if (p == nullptr) return;
  Rec r;
  {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = live_.find(p);
    HB_CHECK(it != live_.end()) << "stream-ordered pool: free of an unknown pointer";
    r = it->second;
    live_.erase(it);
    st_.allocated -= r.size;
    ++st_.num_free;
  }
  if (!cuda_) { std::free(p); return; }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(r.stream);
  for (int64_t o : r.other_streams) {            // the free is ordered after every stream that touched the block
    cudaEvent_t ev;
    if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) continue;
    cudaEventRecord(ev, reinterpret_cast<cudaStream_t>(o));
    cudaStreamWaitEvent(s, ev, 0);
    cudaEventDestroy(ev);
  }
  cudaFreeAsync(p, s);
}
void StreamOrderedMemoryPool::mark_used_by_stream(void* p, int64_t stream) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = live_.find(p);
  if (it == live_.end() || it->second.stream == stream) return;
  auto& v = it->second.other_streams;
  if (std::find(v.begin(), v.end(), stream) == v.end()) v.push_back(stream);
}
void StreamOrderedMemoryPool::wait(void* p) {
  std::vector<int64_t> streams;
  {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = live_.find(p);
    if (it == live_.end()) return;
    streams = it->second.other_streams;
    streams.push_back(it->second.stream);
    it->second.other_streams.clear();
  }
  if (cuda_)
    for (int64_t s : streams) cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(s));
}
size_t StreamOrderedMemoryPool::empty_cache() {
  if (!cuda_) return 0;
  uint64_t before = 0, after = 0;
  cudaMemPool_t mp = static_cast<cudaMemPool_t>(mempool_);
  cudaMemPoolGetAttribute(mp, cudaMemPoolAttrReservedMemCurrent, &before);
  cudaDeviceSynchronize();
  cudaMemPoolTrimTo(mp, 0);
  cudaMemPoolGetAttribute(mp, cudaMemPoolAttrReservedMemCurrent, &after);
  return before > after ? (size_t)(before - after) : 0;
}
PoolStats StreamOrderedMemoryPool::stats() const {
  std::lock_guard<std::mutex> lk(mu_);
  PoolStats s = st_;
  if (cuda_) {
    uint64_t cur = 0, high = 0;
    cudaMemPool_t mp = static_cast<cudaMemPool_t>(mempool_);
    cudaMemPoolGetAttribute(mp, cudaMemPoolAttrReservedMemCurrent, &cur);
    cudaMemPoolGetAttribute(mp, cudaMemPoolAttrReservedMemHigh, &high);
    s.reserved = cur; s.peak_reserved = high;
  } else {
    s.reserved = s.allocated; s.peak_reserved = s.peak_allocated;
  }
  return s;
}
std::string StreamOrderedMemoryPool::summary() const {
  PoolStats s = stats();
  std::ostringstream os;
  os << (cuda_ ? "cuda" : "host") << " stream-ordered pool: reserved " << (s.reserved >> 20) << " MiB (peak " << (s.peak_reserved >> 20)
     << "), allocated " << (s.allocated >> 20) << " MiB (peak " << (s.peak_allocated >> 20) << "), " << s.num_alloc << " allocs, "
     << s.num_free << " frees";
  return os.str();
}

// ------------------------------------------------------------------ which pool backs the tensors of a device
std::shared_ptr<DeviceAllocator> MemoryPoolRegistry::tensor_allocator(const std::string& device) {
  {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = tensor_allocs_.find(device);
    if (it != tensor_allocs_.end()) return it->second;
  }
  const std::string kind = env_str("HETU_MEMORY_POOL", "caching");
  const size_t c = device.find(':');
  const int index = c == std::string::npos ? 0 : std::stoi(device.substr(c + 1));
  std::shared_ptr<DeviceAllocator> a;
  if (kind == "bfc" || kind == "BFC") {
    BFCMemoryPool::Options o;
    const int64_t pre = env_int("HETU_PRE_ALLOCATE_SIZE_MB", 0);
    if (pre > 0) o.initial_region = (size_t)pre << 20;
    const int64_t frag = env_int("HETU_MAX_INTERNAL_FRAGMENT_SIZE_MB", 0);
    if (frag > 0) o.max_internal_fragment = (size_t)frag << 20;
    a = std::make_shared<BFCMemoryPool>(device.rfind("cuda", 0) == 0 ? make_cuda_backend(index) : make_host_backend(false), o);
  } else if (kind == "stream_ordered" || kind == "async") {
    a = std::make_shared<StreamOrderedMemoryPool>(index);
  } else {
    HB_CHECK(kind == "caching") << "HETU_MEMORY_POOL must be caching, bfc or stream_ordered, got '" << kind << "'";
    a = get(device);
  }
  std::lock_guard<std::mutex> lk(mu_);
  tensor_allocs_[device] = a;
  return a;
}

}  // namespace hb
