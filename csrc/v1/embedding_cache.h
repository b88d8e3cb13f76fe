// HET-style client-side embedding cache with bounded staleness (LRU / LFU / LFUOpt policies).
// Rows live in a dense fp32 slab; each row carries a version and a pending-update counter so the
// trainer can push gradients lazily and pull fresh rows only when the staleness bound is exceeded.
// (capability parity: hetu/v1/src/hetu_cache/{cache,lru_cache,lfu_cache,lfuopt_cache}.h,
//  python/hetu/cstable.py)
#pragma once
#include <cstdint>
#include <list>
#include <map>
#include <unordered_map>
#include <vector>

#include "../core/base.h"

namespace hb {

enum class CachePolicy : int { LRU = 0, LFU = 1, LFUOPT = 2 };

struct CacheStats {
  int64_t lookups = 0, hits = 0, evictions = 0, pushes = 0, pulls = 0;
};

class EmbeddingCache {
 public:
  EmbeddingCache(int64_t capacity, int width, CachePolicy policy, int64_t pull_bound, int64_t push_bound);
  int width() const { return width_; }
  int64_t size() const { return (int64_t)index_.size(); }
  int64_t capacity() const { return capacity_; }
  const CacheStats& stats() const { return stats_; }

  // Lookup `keys`; rows present and fresh enough (server_version - version <= pull_bound) are copied to `out`
  // and marked hit. Returns the positions that must be fetched from the server.
  std::vector<int64_t> lookup(const std::vector<int64_t>& keys, const std::vector<int64_t>& server_versions, float* out);
  // Insert rows fetched from the server (evicting by policy). Evicted rows with pending updates are returned
  // so the caller can push them: (key, accumulated gradient row).
  void insert(const std::vector<int64_t>& keys, const float* rows, const std::vector<int64_t>& versions,
              std::vector<int64_t>* evicted_keys, std::vector<float>* evicted_grads);
  // Accumulate gradients locally; keys whose pending updates exceed push_bound are returned for a push.
  void update(const std::vector<int64_t>& keys, const float* grads, float lr, std::vector<int64_t>* push_keys,
              std::vector<float>* push_grads);
  // Flush every pending update.
  void flush(std::vector<int64_t>* push_keys, std::vector<float>* push_grads);
  bool contains(int64_t key) const { return index_.count(key) > 0; }

 private:
  struct Line {
    int64_t key;
    int64_t slot;
    int64_t version;
    int64_t pending;   // number of un-pushed updates
    int64_t freq;
    std::list<int64_t>::iterator lru_it;
  };
  void touch(Line& l);
  int64_t pick_victim();
  int64_t capacity_;
  int width_;
  CachePolicy policy_;
  int64_t pull_bound_, push_bound_;
  std::vector<float> data_, grad_;
  std::unordered_map<int64_t, Line> index_;
  std::list<int64_t> lru_;                       // most recent at front
  std::multimap<int64_t, int64_t> by_freq_;      // freq -> key (LFU)
  std::vector<int64_t> free_slots_;
  CacheStats stats_;
};

}  // namespace hb
