#include "runtime.h"

#include <algorithm>
#include <map>
#include <numeric>
#include <random>

#include "../core/base.h"

namespace hb {

// ------------------------------------------------------------------ streams
const char* stream_role_name(int role) {
  static const char* names[] = {"blocking", "computing", "switch_computing", "h2d", "d2h", "p2p", "collective", "switch_collective",
                                "bridge", "offload"};
  return role >= 0 && role < 10 ? names[role] : "user";
}

cudaStream_t logical_stream(int device, int index) {
  HB_CHECK(index >= 0 && index < kNumLogicalStreams) << "logical stream index " << index << " out of range";
  if (index == kBlockingStream) return nullptr;
  static std::mutex mu;
  static std::map<std::pair<int, int>, cudaStream_t> streams;
  std::lock_guard<std::mutex> lk(mu);
  auto key = std::make_pair(device, index);
  auto it = streams.find(key);
  if (it != streams.end()) return it->second;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  cudaStream_t s = nullptr;
  // collectives / p2p get a higher priority so that their small kernels are not queued behind the compute stream
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  const int prio = (index == kCollectiveStream || index == kP2PStream || index == kSwitchCollectiveStream) ? hi : lo;
  cudaError_t e = cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, prio);
  cudaSetDevice(prev);
  HB_CHECK(e == cudaSuccess) << "cudaStreamCreate failed: " << cudaGetErrorString(e);
  streams[key] = s;
  return s;
}
void sync_logical_stream(int device, int index) {
  cudaError_t e = cudaStreamSynchronize(logical_stream(device, index));
  HB_CHECK(e == cudaSuccess) << "stream sync failed: " << cudaGetErrorString(e);
}

TimedEvent::TimedEvent(int device, bool timing) : device_(device) {
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  cudaEventCreateWithFlags(&ev_, timing ? cudaEventDefault : cudaEventDisableTiming);
  cudaSetDevice(prev);
}
TimedEvent::~TimedEvent() { if (ev_) cudaEventDestroy(ev_); }
void TimedEvent::record(cudaStream_t s) { cudaEventRecord(ev_, s); }
void TimedEvent::sync() { cudaEventSynchronize(ev_); }
bool TimedEvent::query() {
  const cudaError_t e = cudaEventQuery(ev_);
  if (e == cudaErrorNotReady) { cudaGetLastError(); return false; }
  return true;
}
void TimedEvent::block(cudaStream_t waiting_stream) { cudaStreamWaitEvent(waiting_stream, ev_, 0); }
float TimedEvent::elapsed_ms_since(const TimedEvent& start) {
  float ms = 0.f;
  cudaEventElapsedTime(&ms, start.ev_, ev_);
  return ms;
}

// ------------------------------------------------------------------ random state
RandomState& RandomState::get() {
  static RandomState rs;
  return rs;
}
void RandomState::set_seed(uint64_t seed) {
  seed_.store(seed);
  offset_.store(0);
}
uint64_t RandomState::next_offset(uint64_t count) { return offset_.fetch_add(count); }

// ------------------------------------------------------------------ data loader
NativeDataloader::NativeDataloader(at::Tensor data, int64_t batch_size, bool shuffle, bool drop_last, int dp_rank, int dp_size,
                                   uint64_t seed, int prefetch, bool pin_memory)
    : data_(data.contiguous()), batch_size_(batch_size), shuffle_(shuffle), drop_last_(drop_last), dp_rank_(dp_rank), dp_size_(dp_size),
      seed_(seed), prefetch_(std::max(1, prefetch)), pin_(pin_memory) {
  HB_CHECK(data_.dim() >= 1 && batch_size > 0 && dp_size > 0 && dp_rank >= 0 && dp_rank < dp_size) << "bad dataloader arguments";
  HB_CHECK(num_batches() > 0) << "dataset of " << data_.size(0) << " rows yields no batch of " << batch_size;
  thread_ = std::thread([this] { worker(); });
}
NativeDataloader::~NativeDataloader() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
  }
  cv_put_.notify_all();
  cv_get_.notify_all();
  if (thread_.joinable()) thread_.join();
}
int64_t NativeDataloader::num_batches() const {
  const int64_t n = data_.size(0);
  return drop_last_ ? n / batch_size_ : (n + batch_size_ - 1) / batch_size_;
}
at::Tensor NativeDataloader::make_batch(int64_t epoch, int64_t index) {
  const int64_t n = data_.size(0);
  std::vector<int64_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  if (shuffle_) {
    std::mt19937_64 rng(seed_ + (uint64_t)epoch);
    std::shuffle(order.begin(), order.end(), rng);
  }
  const int64_t lo = index * batch_size_, hi = std::min(n, lo + batch_size_);
  std::vector<int64_t> rows;
  for (int64_t i = lo + dp_rank_; i < hi; i += dp_size_) rows.push_back(order[i]);
  at::Tensor idx = at::tensor(rows, at::TensorOptions().dtype(at::kLong));
  at::Tensor out = data_.index_select(0, idx);
  if (pin_ && at::hasCUDA()) out = out.pin_memory();
  return out;
}
void NativeDataloader::worker() {
  for (;;) {
    int64_t ep, idx;
    uint64_t gen;
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_put_.wait(lk, [&] { return stop_ || (int)queue_.size() < prefetch_; });
      if (stop_) return;
      ep = produce_epoch_; idx = next_index_; gen = generation_;
      if (++next_index_ >= num_batches()) { next_index_ = 0; ++produce_epoch_; }
    }
    at::Tensor b = make_batch(ep, idx);
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (gen == generation_) queue_.push_back(b);      // dropped when reset() happened meanwhile
    }
    cv_get_.notify_one();
  }
}
at::Tensor NativeDataloader::next() {
  std::unique_lock<std::mutex> lk(mu_);
  cv_get_.wait(lk, [&] { return stop_ || !queue_.empty(); });
  HB_CHECK(!queue_.empty()) << "dataloader stopped";
  at::Tensor b = queue_.front();
  queue_.pop_front();
  lk.unlock();
  cv_put_.notify_one();
  return b;
}
void NativeDataloader::reset(int64_t start_batch) {
  std::lock_guard<std::mutex> lk(mu_);
  queue_.clear();
  ++generation_;
  next_index_ = start_batch % num_batches();
  produce_epoch_ = start_batch / num_batches();
  epoch_.store(produce_epoch_);
  cv_put_.notify_all();
}

}  // namespace hb
