// Native worker-side client of the DeviceController rendezvous service (length-prefixed JSON over TCP, the same wire
// format as hetu_b200/rpc/server.py): Connect / GetRank, typed key-value store, NCCL-id exchange, barrier / consistency
// check, and a background heart-beat thread on its own connection so a blocking call never delays the liveness signal.
// (ref: hetu/impl/communication/rpc_client.{h,cc} DeviceClientImpl, rpc_comm.cc bootstrap)
#pragma once
#include <atomic>
#include <cstdint>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace hb {

// minimal JSON helpers (enough for the controller's replies; values are passed through as raw JSON text)
std::string json_quote(const std::string& s);
std::string json_unquote(const std::string& s);                              // "a\"b" -> a"b ; non-strings are returned as is
bool json_field(const std::string& obj, const std::string& key, std::string* raw_value);   // top-level field of an object
std::string base64_encode(const std::string& bytes);
std::string base64_decode(const std::string& text);

class RpcClient {
 public:
  RpcClient(std::string host, int port, std::string hostname = "", double heartbeat_interval_s = 2.0, double connect_timeout_s = 60.0);
  ~RpcClient();
  RpcClient(const RpcClient&) = delete;

  // raw call: `args_json` is the JSON object of keyword arguments; returns the raw JSON of the reply's value
  std::string call(const std::string& method, const std::string& args_json);

  // bootstrap: registers this worker, blocks until every worker has connected -> rank / local device / world size
  void connect(bool start_heartbeat = true);
  int rank() const { return rank_; }
  int local_device() const { return local_device_; }
  int world_size() const { return world_size_; }
  const std::string& client_id() const { return client_id_; }

  void put_int(const std::string& key, int64_t v);
  int64_t get_int(const std::string& key);
  void put_double(const std::string& key, double v);
  double get_double(const std::string& key);
  void put_string(const std::string& key, const std::string& v);
  std::string get_string(const std::string& key);
  void put_bytes(const std::string& key, const std::string& v);
  std::string get_bytes(const std::string& key);
  void put_json(const std::string& key, const std::string& raw_json);
  std::string get_json(const std::string& key);
  bool remove(const std::string& key, const std::string& kind = "json");

  void commit_hostname();
  std::string get_hostname(int rank);
  void commit_nccl_id(std::vector<int> ranks, int stream, const std::string& id_bytes);
  std::string get_nccl_id(std::vector<int> ranks, int stream);

  void barrier(const std::vector<int>& ranks = {}, const std::string& tag = "");
  bool consistent(const std::string& raw_json_value, const std::vector<int>& ranks = {}, const std::string& tag = "");
  void worker_stop();
  bool already_stop();
  void exit();
  int64_t heartbeats_sent() const { return beats_.load(); }

 private:
  int open_socket(double timeout_s) const;
  static void send_frame(int fd, const std::string& payload);
  static std::string recv_frame(int fd);
  static std::string ranks_json(const std::vector<int>& ranks);
  void heartbeat_loop();

  std::string host_, hostname_, client_id_;
  int port_;
  double hb_interval_, connect_timeout_;
  int fd_ = -1, hb_fd_ = -1;
  int rank_ = -1, local_device_ = -1, world_size_ = -1;
  std::mutex mu_;
  std::thread hb_thread_;
  std::atomic<bool> hb_stop_{false}, stop_requested_{false};
  std::atomic<int64_t> beats_{0};
};

}  // namespace hb
