// Network transport of the parameter server: `PsNetServer` exposes a ParameterServer on a TCP port (one handler thread per
// worker connection, so blocking calls -- barrier, SSP wait, partial reduce -- of one worker never stall the others),
// `PsNetClient` is the worker-side stub with the same methods.  Binary frames: u32 length | u8 opcode | i32 worker | i64 key |
// op-specific scalars and arrays (u64 count + raw little-endian data).
// (capability parity: ps-lite's van / customer / postoffice layers -- hetu/v1/ps-lite/src/{van,zmq_van,customer}.cc -- and
// the PSF request handlers of hetu/v1/ps-lite/include/ps/psf/)
#pragma once
#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ps_server.h"

namespace hb {

class PsNetServer {
 public:
  PsNetServer(std::shared_ptr<ParameterServer> ps, int port = 0, const std::string& bind_addr = "0.0.0.0");
  ~PsNetServer();
  int port() const { return port_; }
  void stop();
  int64_t requests() const { return requests_.load(); }

 private:
  void accept_loop();
  void serve(int fd);
  std::shared_ptr<ParameterServer> ps_;
  int listen_fd_ = -1, port_ = 0;
  std::atomic<bool> stop_{false};
  std::atomic<int64_t> requests_{0};
  std::thread acceptor_;
  std::mutex mu_;
  std::vector<std::thread> handlers_;
  std::vector<int> fds_;
};

class PsNetClient {
 public:
  PsNetClient(const std::string& host, int port, double connect_timeout_s = 60.0);
  ~PsNetClient();
  void init_dense(int64_t key, const std::vector<float>& value, const PsParamConfig& cfg);
  void push_dense(int64_t key, const std::vector<float>& grad);
  std::vector<float> pull_dense(int64_t key);
  std::vector<float> push_pull_dense(int64_t key, const std::vector<float>& grad);
  void init_sparse(int64_t key, int64_t rows, int width, const std::vector<float>& value, const PsParamConfig& cfg);
  void push_sparse(int64_t key, const std::vector<int64_t>& rows, const std::vector<float>& grads);
  std::vector<float> pull_sparse(int64_t key, const std::vector<int64_t>& rows);
  std::vector<int64_t> row_versions(int64_t key, const std::vector<int64_t>& rows);
  void sync_cache(int64_t key, const std::vector<int64_t>& rows, const std::vector<int64_t>& client_versions, int64_t bound,
                  std::vector<int64_t>* stale_rows, std::vector<float>* fresh_values, std::vector<int64_t>* fresh_versions);
  void barrier(int worker);
  void ssp_init(int staleness);
  void ssp_sync(int worker, int clock);
  std::vector<float> preduce(int worker, int64_t key, const std::vector<float>& value, int min_workers, int wait_ms, std::vector<int>* partners);
  std::map<std::string, int64_t> stats();
  int num_workers();

 private:
  std::string roundtrip(const std::string& req);
  int fd_ = -1;
  std::mutex mu_;
};

}  // namespace hb
