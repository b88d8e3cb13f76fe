// Parameter server core (v1 PS mode): sharded key-value store of dense tensors and sparse embedding tables with
// server-side optimizers, stale-synchronous-parallel clocks, row versions for the HET cache, and partial reduce.
// One instance serves the worker threads of a process; `python/hetu_b200/v1/ps.py` exposes it over TCP for
// multi-process jobs.  (capability parity: ps-lite as modified by Hetu -- hetu/v1/ps-lite/**: PSFunc dense/sparse
// push-pull (PSFHandle), SSP (ssp_handler.h), preduce (preduce_handler.h), cache sync (CacheSync PSF);
// python/hetu/v1 communicator bindings)
#pragma once
#include <condition_variable>
#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../core/base.h"

namespace hb {

enum class PsOptimizer : int { NONE = 0, SGD = 1, MOMENTUM = 2, ADAGRAD = 3, ADAM = 4 };

struct PsParamConfig {
  PsOptimizer opt = PsOptimizer::SGD;
  float lr = 0.01f, momentum = 0.9f, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-7f;
};

class ParameterServer {
 public:
  explicit ParameterServer(int num_workers);
  int num_workers() const { return num_workers_; }

  // ---- dense parameters
  void init_dense(int64_t key, const std::vector<float>& value, const PsParamConfig& cfg);
  void push_dense(int64_t key, const std::vector<float>& grad);                  // server-side optimizer step
  std::vector<float> pull_dense(int64_t key);
  std::vector<float> push_pull_dense(int64_t key, const std::vector<float>& grad);

  // ---- sparse (embedding) parameters: rows x width
  void init_sparse(int64_t key, int64_t rows, int width, const std::vector<float>& value, const PsParamConfig& cfg);
  void push_sparse(int64_t key, const std::vector<int64_t>& rows, const std::vector<float>& grads);
  std::vector<float> pull_sparse(int64_t key, const std::vector<int64_t>& rows);
  std::vector<int64_t> row_versions(int64_t key, const std::vector<int64_t>& rows);   // HET cache bounded staleness
  // cache synchronisation: push accumulated updates, get back the rows whose server version ran ahead by > bound
  void sync_cache(int64_t key, const std::vector<int64_t>& rows, const std::vector<int64_t>& client_versions, int64_t bound,
                  std::vector<int64_t>* stale_rows, std::vector<float>* fresh_values, std::vector<int64_t>* fresh_versions);

  // ---- consistency
  void barrier(int worker);                                   // BSP barrier over all workers
  void ssp_init(int staleness);
  void ssp_sync(int worker, int clock);                       // blocks while clock - min(clocks) > staleness
  // partial reduce: group the first `min_workers` arrivals within the wait window, average their vectors
  std::vector<float> preduce(int worker, int64_t key, const std::vector<float>& value, int min_workers, int wait_ms,
                             std::vector<int>* partners);
  std::map<std::string, int64_t> stats() const;

 private:
  struct Dense { std::vector<float> w, s1, s2; PsParamConfig cfg; int64_t step = 0; };
  struct Sparse { std::vector<float> w, s1, s2; std::vector<int64_t> version; int64_t rows = 0; int width = 0; PsParamConfig cfg; int64_t step = 0; };
  static void apply(const PsParamConfig& cfg, int64_t step, float* w, float* s1, float* s2, const float* g, int64_t n);

  int num_workers_;
  mutable std::mutex mu_;
  std::unordered_map<int64_t, Dense> dense_;
  std::unordered_map<int64_t, Sparse> sparse_;
  // barrier
  std::condition_variable bar_cv_;
  int bar_count_ = 0;
  int64_t bar_gen_ = 0;
  // ssp
  std::condition_variable ssp_cv_;
  std::vector<int> clocks_;
  int staleness_ = 0;
  // preduce
  struct PGroup { std::vector<int> members; std::vector<float> sum; bool closed = false; int taken = 0; int64_t id = 0; };
  std::condition_variable pr_cv_;
  std::map<int64_t, PGroup> open_;          // key -> group currently collecting
  std::map<int64_t, PGroup> done_;          // group id -> finished group until all members fetched it
  int64_t next_group_ = 1;
  std::map<std::pair<int, int64_t>, int64_t> my_group_;
  int64_t n_push_ = 0, n_pull_ = 0, n_preduce_ = 0;
};

}  // namespace hb
