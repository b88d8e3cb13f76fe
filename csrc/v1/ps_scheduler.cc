#include <algorithm>
#include "ps_scheduler.h"

#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/tcp.h>

#include "../core/base.h"
#include "ps_wire.h"

namespace hb {
using namespace ps_wire;
namespace {
enum SOp : uint8_t { S_REGISTER = 1, S_BARRIER, S_HEARTBEAT, S_DEAD, S_FINALIZE, S_PREDUCE };

void put_node(Writer& w, const PsNodeInfo& n) { w.put<int32_t>(n.role); w.put<int32_t>(n.rank); w.str(n.host); w.put<int32_t>(n.port); }
PsNodeInfo get_node(Reader& r) {
  PsNodeInfo n;
  n.role = r.get<int32_t>(); n.rank = r.get<int32_t>(); n.host = r.str(); n.port = r.get<int32_t>();
  return n;
}
}  // namespace

PsScheduler::PsScheduler(int num_servers, int num_workers, int port, const std::string& bind_addr)
    : num_servers_(num_servers), num_workers_(num_workers) {
  HB_CHECK(num_servers > 0 && num_workers > 0) << "scheduler needs at least one server and one worker";
  listen_fd_ = ::socket(AF_INET, SOCK_STREAM, 0);
  HB_CHECK(listen_fd_ >= 0) << "ps scheduler: cannot create a socket";
  int one = 1;
  setsockopt(listen_fd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
  sockaddr_in a{};
  a.sin_family = AF_INET;
  a.sin_port = htons((uint16_t)port);
  a.sin_addr.s_addr = bind_addr == "0.0.0.0" ? INADDR_ANY : inet_addr(bind_addr.c_str());
  HB_CHECK(::bind(listen_fd_, (sockaddr*)&a, sizeof a) == 0) << "ps scheduler: cannot bind port " << port;
  HB_CHECK(::listen(listen_fd_, 256) == 0) << "ps scheduler: listen failed";
  socklen_t len = sizeof a;
  getsockname(listen_fd_, (sockaddr*)&a, &len);
  port_ = ntohs(a.sin_port);
  acceptor_ = std::thread([this] { accept_loop(); });
}
PsScheduler::~PsScheduler() { stop(); }

void PsScheduler::stop() {
  if (stop_.exchange(true)) return;
  ::shutdown(listen_fd_, SHUT_RDWR);
  ::close(listen_fd_);
  if (acceptor_.joinable()) acceptor_.join();
  std::vector<std::thread> hs;
  {
    std::lock_guard<std::mutex> g(mu_);
    for (int fd : fds_) ::shutdown(fd, SHUT_RDWR);
    hs.swap(handlers_);
  }
  cv_.notify_all();
  for (auto& t : hs) if (t.joinable()) t.join();
}

void PsScheduler::accept_loop() {
  while (!stop_) {
    int fd = ::accept(listen_fd_, nullptr, nullptr);
    if (fd < 0) { if (stop_) return; continue; }
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
    std::lock_guard<std::mutex> g(mu_);
    fds_.push_back(fd);
    handlers_.emplace_back([this, fd] { serve(fd); ::close(fd); });
  }
}

void PsScheduler::serve(int fd) {
  std::string req;
  while (!stop_ && recv_frame(fd, &req)) {
    Writer out;
    out.put<uint8_t>(0);
    try {
      Reader r(req);
      const SOp op = (SOp)r.get<uint8_t>();
      const int node = r.get<int32_t>();
      std::unique_lock<std::mutex> lk(mu_);
      if (node >= 0 && node < (int)nodes_.size()) nodes_[node].seen = std::chrono::steady_clock::now();
      switch (op) {
        case S_REGISTER: {
          PsNodeInfo n = get_node(r);
          int rank = 0;
          for (auto& o : nodes_) rank += o.info.role == n.role;
          HB_CHECK(rank < (n.role == 0 ? num_servers_ : num_workers_)) << "more " << (n.role == 0 ? "servers" : "workers") << " registered than expected";
          n.rank = rank;
          const int id = (int)nodes_.size();
          nodes_.push_back({n, std::chrono::steady_clock::now(), false});
          cv_.notify_all();
          cv_.wait(lk, [&] { return stop_ || (int)nodes_.size() == num_servers_ + num_workers_; });
          HB_CHECK(!stop_) << "scheduler stopped during registration";
          out.put<int32_t>(id); out.put<int32_t>(rank); out.put<int32_t>(num_workers_);
          out.put<int32_t>(num_servers_);
          for (int s = 0; s < num_servers_; ++s)
            for (auto& o : nodes_)
              if (o.info.role == 0 && o.info.rank == s) put_node(out, o.info);
          break;
        }
        case S_BARRIER: {
          const int group = r.get<int32_t>();
          const int expect = ((group & kServerGroup) ? num_servers_ : 0) + ((group & kWorkerGroup) ? num_workers_ : 0);
          auto& b = barrier_[group];
          const uint64_t gen = b.second;
          if (++b.first == expect) {
            b.first = 0;
            ++b.second;
            cv_.notify_all();
          } else {
            cv_.wait(lk, [&] { return stop_ || barrier_[group].second != gen; });
            HB_CHECK(!stop_) << "scheduler stopped inside a barrier";
          }
          break;
        }
        case S_HEARTBEAT: break;      // `seen` was refreshed above
        case S_DEAD: {
          const double timeout = r.get<double>();
          lk.unlock();
          auto dead = dead_nodes(timeout);
          out.put<int32_t>((int32_t)dead.size());
          for (auto& d : dead) put_node(out, d);
          break;
        }
        case S_FINALIZE: {
          if (node >= 0 && node < (int)nodes_.size()) nodes_[node].finalized = true;
          cv_.notify_all();
          break;
        }
        case S_PREDUCE: {
          // partial reduce matchmaking: the first worker to ask under `key` opens a round; whoever asks before the round's
          // deadline (or until `max_worker` have gathered) is in it, and all of them receive the same sorted member list
          const int key = r.get<int32_t>(), rank = r.get<int32_t>(), max_worker = r.get<int32_t>();
          const double wait_ms = r.get<double>();
          auto& st = preduce_[key];
          if (!st.open) {
            st.open = true;
            ++st.gen;
            st.members.clear();
            st.deadline = std::chrono::steady_clock::now() + std::chrono::microseconds((int64_t)(wait_ms * 1e3));
          }
          const uint64_t gen = st.gen;
          st.members.push_back(rank);
          auto close = [&] {
            std::sort(st.members.begin(), st.members.end());
            st.done[gen] = {st.members, (int)st.members.size()};
            st.open = false;
            cv_.notify_all();
          };
          if ((int)st.members.size() >= std::max(1, max_worker)) {
            close();
          } else {
            cv_.wait_until(lk, st.deadline, [&] { return stop_ || st.done.count(gen) > 0; });
            HB_CHECK(!stop_) << "scheduler stopped inside a partial-reduce round";
            if (!st.done.count(gen)) close();          // the deadline passed: this waiter seals the round for everybody in it
          }
          auto& res = st.done[gen];
          out.put<int32_t>((int32_t)res.first.size());
          for (int m : res.first) out.put<int32_t>(m);
          if (--res.second == 0) st.done.erase(gen);
          break;
        }
        default: HB_FAIL() << "unknown scheduler request " << (int)op;
      }
    } catch (const std::exception& e) {
      out = Writer();
      out.put<uint8_t>(1);
      out.str(e.what());
    }
    try {
      send_frame(fd, out.b);
    } catch (...) {
      return;
    }
  }
}

std::vector<PsNodeInfo> PsScheduler::dead_nodes(double timeout_s) const {
  std::lock_guard<std::mutex> g(mu_);
  std::vector<PsNodeInfo> out;
  const auto now = std::chrono::steady_clock::now();
  for (auto& n : nodes_)
    if (!n.finalized && std::chrono::duration<double>(now - n.seen).count() > timeout_s) out.push_back(n.info);
  return out;
}
int PsScheduler::registered() const {
  std::lock_guard<std::mutex> g(mu_);
  return (int)nodes_.size();
}
int PsScheduler::finalized() const {
  std::lock_guard<std::mutex> g(mu_);
  int n = 0;
  for (auto& x : nodes_) n += x.finalized;
  return n;
}
bool PsScheduler::wait_finalized(double timeout_s) {
  std::unique_lock<std::mutex> lk(mu_);
  return cv_.wait_for(lk, std::chrono::duration<double>(timeout_s), [&] {
    if ((int)nodes_.size() < num_servers_ + num_workers_) return false;
    for (auto& x : nodes_) if (!x.finalized) return false;
    return true;
  });
}

// ------------------------------------------------------------------------------------------------------------------ client
PsSchedulerClient::PsSchedulerClient(const std::string& host, int port, int role, const std::string& my_host, int my_port, double timeout)
    : role_(role) {
  auto deadline = std::chrono::steady_clock::now() + std::chrono::duration<double>(timeout);
  for (;;) {
    addrinfo hints{}, *res = nullptr;
    hints.ai_family = AF_INET;
    hints.ai_socktype = SOCK_STREAM;
    if (getaddrinfo(host.c_str(), std::to_string(port).c_str(), &hints, &res) == 0) {
      for (addrinfo* a = res; a && fd_ < 0; a = a->ai_next) {
        int fd = ::socket(a->ai_family, a->ai_socktype, a->ai_protocol);
        if (fd < 0) continue;
        if (::connect(fd, a->ai_addr, a->ai_addrlen) == 0) fd_ = fd; else ::close(fd);
      }
      freeaddrinfo(res);
    }
    if (fd_ >= 0) break;
    if (std::chrono::steady_clock::now() > deadline) throw std::runtime_error("ps: cannot reach the scheduler at " + host + ":" + std::to_string(port));
    std::this_thread::sleep_for(std::chrono::milliseconds(100));
  }
  int one = 1;
  setsockopt(fd_, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
  Writer w;
  w.put<uint8_t>(S_REGISTER); w.put<int32_t>(-1);
  PsNodeInfo me;
  me.role = role; me.host = my_host; me.port = my_port;
  put_node(w, me);
  std::string rep = roundtrip(w.b);
  Reader r(rep);
  r.get<uint8_t>();
  node_id_ = r.get<int32_t>(); rank_ = r.get<int32_t>(); num_workers_ = r.get<int32_t>();
  const int ns = r.get<int32_t>();
  for (int i = 0; i < ns; ++i) servers_.push_back(get_node(r));
}
PsSchedulerClient::~PsSchedulerClient() {
  hb_stop_ = true;
  if (hb_.joinable()) hb_.join();
  if (fd_ >= 0) ::close(fd_);
}
std::string PsSchedulerClient::roundtrip(const std::string& req) {
  std::string rep;
  {
    std::lock_guard<std::mutex> g(mu_);
    send_frame(fd_, req);
    if (!recv_frame(fd_, &rep)) throw std::runtime_error("ps: scheduler closed the connection");
  }
  if (rep.empty()) throw std::runtime_error("ps: empty scheduler reply");
  if (rep[0] != 0) {
    Reader r(rep);
    r.get<uint8_t>();
    throw std::runtime_error("ps scheduler: " + r.str());
  }
  return rep;
}
void PsSchedulerClient::barrier(int group) {
  Writer w;
  w.put<uint8_t>(S_BARRIER); w.put<int32_t>(node_id_); w.put<int32_t>(group);
  roundtrip(w.b);
}
std::vector<int> PsSchedulerClient::preduce_partners(int key, int rank, int max_worker, double wait_ms) {
  Writer w;
  w.put<uint8_t>(S_PREDUCE); w.put<int32_t>(node_id_); w.put<int32_t>(key); w.put<int32_t>(rank); w.put<int32_t>(max_worker);
  w.put<double>(wait_ms);
  std::string rep = roundtrip(w.b);
  Reader r(rep);
  r.get<uint8_t>();
  const int n = r.get<int32_t>();
  std::vector<int> out(n);
  for (int i = 0; i < n; ++i) out[i] = r.get<int32_t>();
  return out;
}
void PsSchedulerClient::heartbeat() {
  Writer w;
  w.put<uint8_t>(S_HEARTBEAT); w.put<int32_t>(node_id_);
  roundtrip(w.b);
}
std::vector<PsNodeInfo> PsSchedulerClient::dead_nodes(double timeout_s) {
  Writer w;
  w.put<uint8_t>(S_DEAD); w.put<int32_t>(node_id_); w.put<double>(timeout_s);
  std::string rep = roundtrip(w.b);
  Reader r(rep);
  r.get<uint8_t>();
  const int n = r.get<int32_t>();
  std::vector<PsNodeInfo> out;
  for (int i = 0; i < n; ++i) out.push_back(get_node(r));
  return out;
}
std::vector<int64_t> PsSchedulerClient::key_ranges(int64_t total) const {
  const int64_t s = (int64_t)servers_.size();
  std::vector<int64_t> begin(s + 1);
  for (int64_t i = 0; i <= s; ++i) begin[i] = total / s * i + std::min<int64_t>(i, total % s);
  return begin;
}
void PsSchedulerClient::finalize() {
  if (finalized_) return;
  finalized_ = true;
  hb_stop_ = true;
  if (hb_.joinable()) hb_.join();
  Writer w;
  w.put<uint8_t>(S_FINALIZE); w.put<int32_t>(node_id_);
  roundtrip(w.b);
}
void PsSchedulerClient::start_heartbeat(double interval_s) {
  if (hb_.joinable()) return;
  hb_stop_ = false;
  hb_ = std::thread([this, interval_s] {
    auto next = std::chrono::steady_clock::now();
    while (!hb_stop_) {
      if (std::chrono::steady_clock::now() >= next) {
        try {
          heartbeat();
        } catch (...) {
          return;
        }
        next = std::chrono::steady_clock::now() + std::chrono::duration_cast<std::chrono::steady_clock::duration>(std::chrono::duration<double>(interval_s));
      }
      std::this_thread::sleep_for(std::chrono::milliseconds(10));
    }
  });
}

}  // namespace hb
