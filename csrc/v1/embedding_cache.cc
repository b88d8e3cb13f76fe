#include "embedding_cache.h"

#include <algorithm>
#include <cstring>

namespace hb {

EmbeddingCache::EmbeddingCache(int64_t capacity, int width, CachePolicy policy, int64_t pull_bound, int64_t push_bound)
    : capacity_(capacity), width_(width), policy_(policy), pull_bound_(pull_bound), push_bound_(push_bound) {
  HB_CHECK(capacity > 0 && width > 0) << "cache needs positive capacity and width";
  data_.assign((size_t)capacity * width, 0.f);
  grad_.assign((size_t)capacity * width, 0.f);
  for (int64_t i = capacity - 1; i >= 0; --i) free_slots_.push_back(i);
}

void EmbeddingCache::touch(Line& l) {
  if (policy_ == CachePolicy::LRU) {
    lru_.erase(l.lru_it);
    lru_.push_front(l.key);
    l.lru_it = lru_.begin();
  } else {
    auto range = by_freq_.equal_range(l.freq);
    for (auto it = range.first; it != range.second; ++it)
      if (it->second == l.key) { by_freq_.erase(it); break; }
    l.freq += 1;
    by_freq_.insert({l.freq, l.key});
  }
}

int64_t EmbeddingCache::pick_victim() {
  if (policy_ == CachePolicy::LRU) return lru_.back();
  if (policy_ == CachePolicy::LFU) return by_freq_.begin()->second;
  // LFUOpt: among the least-frequent lines prefer one without pending updates (cheaper eviction)
  auto lo = by_freq_.begin();
  auto range = by_freq_.equal_range(lo->first);
  for (auto it = range.first; it != range.second; ++it)
    if (index_[it->second].pending == 0) return it->second;
  return lo->second;
}

std::vector<int64_t> EmbeddingCache::lookup(const std::vector<int64_t>& keys, const std::vector<int64_t>& server_versions,
                                            float* out) {
  std::vector<int64_t> miss;
  for (size_t i = 0; i < keys.size(); ++i) {
    stats_.lookups++;
    auto it = index_.find(keys[i]);
    const int64_t sv = i < server_versions.size() ? server_versions[i] : 0;
    if (it != index_.end() && sv - it->second.version <= pull_bound_) {
      stats_.hits++;
      std::memcpy(out + i * width_, &data_[(size_t)it->second.slot * width_], sizeof(float) * width_);
      touch(it->second);
    } else {
      miss.push_back((int64_t)i);
    }
  }
  return miss;
}

void EmbeddingCache::insert(const std::vector<int64_t>& keys, const float* rows, const std::vector<int64_t>& versions,
                            std::vector<int64_t>* evicted_keys, std::vector<float>* evicted_grads) {
  for (size_t i = 0; i < keys.size(); ++i) {
    stats_.pulls++;
    auto it = index_.find(keys[i]);
    if (it != index_.end()) {  // refresh a stale line in place, keep its pending gradient
      std::memcpy(&data_[(size_t)it->second.slot * width_], rows + i * width_, sizeof(float) * width_);
      it->second.version = i < versions.size() ? versions[i] : 0;
      touch(it->second);
      continue;
    }
    if (free_slots_.empty()) {
      const int64_t vk = pick_victim();
      Line& v = index_[vk];
      stats_.evictions++;
      if (v.pending > 0 && evicted_keys) {
        evicted_keys->push_back(vk);
        evicted_grads->insert(evicted_grads->end(), grad_.begin() + v.slot * width_, grad_.begin() + (v.slot + 1) * width_);
      }
      std::fill(grad_.begin() + v.slot * width_, grad_.begin() + (v.slot + 1) * width_, 0.f);
      if (policy_ == CachePolicy::LRU) lru_.erase(v.lru_it);
      else {
        auto range = by_freq_.equal_range(v.freq);
        for (auto fit = range.first; fit != range.second; ++fit)
          if (fit->second == vk) { by_freq_.erase(fit); break; }
      }
      free_slots_.push_back(v.slot);
      index_.erase(vk);
    }
    Line l;
    l.key = keys[i];
    l.slot = free_slots_.back();
    free_slots_.pop_back();
    l.version = i < versions.size() ? versions[i] : 0;
    l.pending = 0;
    l.freq = 1;
    std::memcpy(&data_[(size_t)l.slot * width_], rows + i * width_, sizeof(float) * width_);
    if (policy_ == CachePolicy::LRU) {
      lru_.push_front(l.key);
      l.lru_it = lru_.begin();
    } else by_freq_.insert({l.freq, l.key});
    index_[l.key] = l;
  }
}

void EmbeddingCache::update(const std::vector<int64_t>& keys, const float* grads, float lr, std::vector<int64_t>* push_keys,
                            std::vector<float>* push_grads) {
  for (size_t i = 0; i < keys.size(); ++i) {
    auto it = index_.find(keys[i]);
    if (it == index_.end()) {  // not cached: forward straight to the server
      push_keys->push_back(keys[i]);
      push_grads->insert(push_grads->end(), grads + i * width_, grads + (i + 1) * width_);
      stats_.pushes++;
      continue;
    }
    Line& l = it->second;
    float* g = &grad_[(size_t)l.slot * width_];
    float* d = &data_[(size_t)l.slot * width_];
    for (int c = 0; c < width_; ++c) {
      g[c] += grads[i * width_ + c];
      d[c] -= lr * grads[i * width_ + c];   // local view stays fresh for this worker
    }
    l.pending += 1;
    if (l.pending > push_bound_) {
      push_keys->push_back(l.key);
      push_grads->insert(push_grads->end(), g, g + width_);
      std::fill(g, g + width_, 0.f);
      l.pending = 0;
      l.version += 1;
      stats_.pushes++;
    }
  }
}

void EmbeddingCache::flush(std::vector<int64_t>* push_keys, std::vector<float>* push_grads) {
  for (auto& kv : index_) {
    Line& l = kv.second;
    if (l.pending == 0) continue;
    float* g = &grad_[(size_t)l.slot * width_];
    push_keys->push_back(l.key);
    push_grads->insert(push_grads->end(), g, g + width_);
    std::fill(g, g + width_, 0.f);
    l.pending = 0;
    stats_.pushes++;
  }
}

}  // namespace hb
