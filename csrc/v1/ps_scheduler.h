// Scheduler node of the parameter-server deployment: servers and workers register with it, learn their rank and the
// address table, synchronise on group barriers, report liveness by heartbeat and check out at the end; it also hands out
// the key ranges that partition dense parameters over the servers.
// (capability parity: ps-lite's scheduler role -- Postoffice node management, ADD_NODE / BARRIER / HEARTBEAT control
//  messages of hetu/v1/ps-lite/src/{postoffice,van}.cc, GetServerKeyRanges -- own protocol over the frames of ps_wire.h)
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace hb {

struct PsNodeInfo {
  int role = 0;            // 0 server, 1 worker
  int rank = 0;            // rank within the role
  std::string host;
  int port = 0;            // servers: the PsNetServer port; workers: 0
};
enum PsGroup : int { kServerGroup = 1, kWorkerGroup = 2, kAllGroup = 3 };

class PsScheduler {
 public:
  PsScheduler(int num_servers, int num_workers, int port = 0, const std::string& bind_addr = "0.0.0.0");
  ~PsScheduler();
  int port() const { return port_; }
  void stop();
  // nodes that have registered but not sent a heartbeat (or any request) for `timeout_s`
  std::vector<PsNodeInfo> dead_nodes(double timeout_s) const;
  int registered() const;
  int finalized() const;
  // block until every node has checked out (or the timeout passes); true when all did
  bool wait_finalized(double timeout_s);

 private:
  struct Node { PsNodeInfo info; std::chrono::steady_clock::time_point seen; bool finalized = false; };
  void accept_loop();
  void serve(int fd);
  int num_servers_, num_workers_;
  int listen_fd_ = -1, port_ = 0;
  std::atomic<bool> stop_{false};
  std::thread acceptor_;
  mutable std::mutex mu_;
  std::condition_variable cv_;
  std::vector<Node> nodes_;                     // registration order
  std::map<int, std::pair<int, uint64_t>> barrier_;   // group -> (arrived, generation)
  struct PReduceRound {
    bool open = false;
    uint64_t gen = 0;
    std::vector<int> members;
    std::chrono::steady_clock::time_point deadline;
    std::map<uint64_t, std::pair<std::vector<int>, int>> done;   // sealed rounds: members + how many have not collected them yet
  };
  std::map<int, PReduceRound> preduce_;
  std::vector<std::thread> handlers_;
  std::vector<int> fds_;
};

class PsSchedulerClient {
 public:
  // registers and blocks until every expected node has registered; servers pass the port their PsNetServer listens on
  PsSchedulerClient(const std::string& sched_host, int sched_port, int role, const std::string& my_host, int my_port,
                    double connect_timeout_s = 60.0);
  ~PsSchedulerClient();
  int role() const { return role_; }
  int rank() const { return rank_; }
  int num_servers() const { return (int)servers_.size(); }
  int num_workers() const { return num_workers_; }
  const std::vector<PsNodeInfo>& servers() const { return servers_; }
  void barrier(int group);
  // partial reduce: the workers that ask under `key` within `wait_ms` of the first one (at most `max_worker`) form a group
  std::vector<int> preduce_partners(int key, int rank, int max_worker, double wait_ms);
  void heartbeat();
  std::vector<PsNodeInfo> dead_nodes(double timeout_s);
  // contiguous split of [0, total) over the servers: begin offsets (size num_servers + 1)
  std::vector<int64_t> key_ranges(int64_t total) const;
  void finalize();
  // background heartbeat every `interval_s` until the client is destroyed / finalized
  void start_heartbeat(double interval_s);

 private:
  std::string roundtrip(const std::string& req);
  int fd_ = -1, role_ = 0, rank_ = 0, node_id_ = -1, num_workers_ = 0;
  std::vector<PsNodeInfo> servers_;
  std::mutex mu_;
  std::thread hb_;
  std::atomic<bool> hb_stop_{false};
  bool finalized_ = false;
};

}  // namespace hb
