// Small native runtime pieces: logical streams / events, the global random state and a prefetching data loader.
// (capability parity: hetu/core/stream.h + hetu/impl/stream/CUDAStream.{h,cc} -- 16 logical streams per device with
//  fixed roles, events with timing; hetu/impl/random/* -- seed + Philox offset bookkeeping;
//  hetu/graph/data/dataloader.{h,cc} -- prefetch queue over an array with data-parallel slicing)
#pragma once
#include <ATen/ATen.h>
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "../core/device.h"

namespace hb {

// ------------------------------------------------------------------ streams
// roles: core/device.h StreamIndex (blocking, computing, switch-computing, h2d, d2h, p2p, collective, switch-collective,
// bridge, offload, ..., join)
constexpr int kNumLogicalStreams = kNumStreams;
const char* stream_role_name(int role);
// lazily created non-blocking CUDA stream for (device, logical index); index 0 is the legacy default stream
cudaStream_t logical_stream(int device, int index);
void sync_logical_stream(int device, int index);

class TimedEvent {
 public:
  explicit TimedEvent(int device = 0, bool timing = true);
  ~TimedEvent();
  void record(cudaStream_t s);
  void sync();
  bool query();
  void block(cudaStream_t waiting_stream);      // make `waiting_stream` wait for this event
  float elapsed_ms_since(const TimedEvent& start);

 private:
  cudaEvent_t ev_ = nullptr;
  int device_;
};

// ------------------------------------------------------------------ random state
// One process-wide seed; every random kernel launch reserves a contiguous range of Philox counter offsets, so
// re-running the same graph from the same (seed, offset) -- e.g. activation recompute -- reproduces identical numbers.
class RandomState {
 public:
  static RandomState& get();
  void set_seed(uint64_t seed);
  uint64_t seed() const { return seed_.load(); }
  uint64_t next_offset(uint64_t count);      // reserve `count` counters, returns the first
  uint64_t offset() const { return offset_.load(); }
  void set_offset(uint64_t o) { offset_.store(o); }

 private:
  std::atomic<uint64_t> seed_{0x5DEECE66Dull}, offset_{0};
};

// ------------------------------------------------------------------ data loader
class NativeDataloader {
 public:
  // data: [N, ...] tensor on the host; batches of `batch_size` rows, this rank takes rows dp_rank::dp_size of each batch
  NativeDataloader(at::Tensor data, int64_t batch_size, bool shuffle, bool drop_last, int dp_rank, int dp_size, uint64_t seed,
                   int prefetch, bool pin_memory);
  ~NativeDataloader();
  int64_t num_batches() const;
  at::Tensor next();        // blocks for the next batch; wraps around at the end of an epoch
  int64_t epoch() const { return epoch_.load(); }
  void reset(int64_t start_batch);

 private:
  void worker();
  at::Tensor make_batch(int64_t epoch, int64_t index);
  at::Tensor data_;
  int64_t batch_size_;
  bool shuffle_, drop_last_;
  int dp_rank_, dp_size_;
  uint64_t seed_;
  int prefetch_;
  bool pin_;
  std::thread thread_;
  std::mutex mu_;
  std::condition_variable cv_put_, cv_get_;
  std::deque<at::Tensor> queue_;
  std::atomic<int64_t> epoch_{0};
  int64_t next_index_ = 0, produce_epoch_ = 0;
  uint64_t generation_ = 0;
  bool stop_ = false;
};

}  // namespace hb
