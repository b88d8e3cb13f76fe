#include "ps_net.h"

#include "ps_wire.h"

#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <chrono>
#include <cstring>
#include <stdexcept>

namespace hb {
using namespace ps_wire;
namespace {

enum Op : uint8_t { INIT_DENSE = 1, PUSH_DENSE, PULL_DENSE, PUSH_PULL_DENSE, INIT_SPARSE, PUSH_SPARSE, PULL_SPARSE, ROW_VERSIONS, SYNC_CACHE,
                    BARRIER, SSP_INIT, SSP_SYNC, PREDUCE, STATS, NUM_WORKERS };

}  // namespace

// ------------------------------------------------------------------------------------------------------------------ server
PsNetServer::PsNetServer(std::shared_ptr<ParameterServer> ps, int port, const std::string& bind_addr) : ps_(std::move(ps)) {
  listen_fd_ = ::socket(AF_INET, SOCK_STREAM, 0);
  HB_CHECK(listen_fd_ >= 0) << "ps: cannot create a socket";
  int one = 1;
  setsockopt(listen_fd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
  sockaddr_in a{};
  a.sin_family = AF_INET;
  a.sin_port = htons((uint16_t)port);
  a.sin_addr.s_addr = bind_addr == "0.0.0.0" ? INADDR_ANY : inet_addr(bind_addr.c_str());
  HB_CHECK(::bind(listen_fd_, (sockaddr*)&a, sizeof a) == 0) << "ps: cannot bind port " << port;
  HB_CHECK(::listen(listen_fd_, 128) == 0) << "ps: listen failed";
  socklen_t len = sizeof a;
  getsockname(listen_fd_, (sockaddr*)&a, &len);
  port_ = ntohs(a.sin_port);
  acceptor_ = std::thread([this] { accept_loop(); });
}

PsNetServer::~PsNetServer() { stop(); }

void PsNetServer::stop() {
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
  for (auto& t : hs) if (t.joinable()) t.join();
}

void PsNetServer::accept_loop() {
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

void PsNetServer::serve(int fd) {
  std::string req;
  while (!stop_ && recv_frame(fd, &req)) {
    requests_.fetch_add(1);
    Writer out;
    out.put<uint8_t>(0);   // status, patched on error
    try {
      Reader r(req);
      const Op op = (Op)r.get<uint8_t>();
      const int worker = r.get<int32_t>();
      const int64_t key = r.get<int64_t>();
      switch (op) {
        case INIT_DENSE: { auto c = r.cfg(); ps_->init_dense(key, r.arr<float>(), c); break; }
        case PUSH_DENSE: ps_->push_dense(key, r.arr<float>()); break;
        case PULL_DENSE: out.arr(ps_->pull_dense(key)); break;
        case PUSH_PULL_DENSE: out.arr(ps_->push_pull_dense(key, r.arr<float>())); break;
        case INIT_SPARSE: {
          auto c = r.cfg();
          const int64_t rows = r.get<int64_t>();
          const int width = r.get<int32_t>();
          ps_->init_sparse(key, rows, width, r.arr<float>(), c);
          break;
        }
        case PUSH_SPARSE: { auto rows = r.arr<int64_t>(); ps_->push_sparse(key, rows, r.arr<float>()); break; }
        case PULL_SPARSE: out.arr(ps_->pull_sparse(key, r.arr<int64_t>())); break;
        case ROW_VERSIONS: out.arr(ps_->row_versions(key, r.arr<int64_t>())); break;
        case SYNC_CACHE: {
          auto rows = r.arr<int64_t>();
          auto vers = r.arr<int64_t>();
          const int64_t bound = r.get<int64_t>();
          std::vector<int64_t> stale, fv;
          std::vector<float> vals;
          ps_->sync_cache(key, rows, vers, bound, &stale, &vals, &fv);
          out.arr(stale); out.arr(vals); out.arr(fv);
          break;
        }
        case BARRIER: ps_->barrier(worker); break;
        case SSP_INIT: ps_->ssp_init(r.get<int32_t>()); break;
        case SSP_SYNC: ps_->ssp_sync(worker, r.get<int32_t>()); break;
        case PREDUCE: {
          const int min_workers = r.get<int32_t>(), wait_ms = r.get<int32_t>();
          std::vector<int> partners;
          out.arr(ps_->preduce(worker, key, r.arr<float>(), min_workers, wait_ms, &partners));
          out.arr(std::vector<int32_t>(partners.begin(), partners.end()));
          break;
        }
        case STATS: {
          auto st = ps_->stats();
          out.put<uint64_t>(st.size());
          for (auto& kv : st) { out.str(kv.first); out.put<int64_t>(kv.second); }
          break;
        }
        case NUM_WORKERS: out.put<int32_t>(ps_->num_workers()); break;
        default: throw std::runtime_error("ps: unknown opcode");
      }
    } catch (const std::exception& e) {
      out.b.clear();
      out.put<uint8_t>(1);
      out.str(e.what());
    }
    try { send_frame(fd, out.b); } catch (const std::exception&) { return; }
  }
}

// ------------------------------------------------------------------------------------------------------------------ client
PsNetClient::PsNetClient(const std::string& host, int port, double timeout) {
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
    if (std::chrono::steady_clock::now() > deadline) throw std::runtime_error("ps: cannot reach the server at " + host + ":" + std::to_string(port));
    std::this_thread::sleep_for(std::chrono::milliseconds(100));
  }
  int one = 1;
  setsockopt(fd_, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
}
PsNetClient::~PsNetClient() { if (fd_ >= 0) ::close(fd_); }

std::string PsNetClient::roundtrip(const std::string& req) {
  std::string rep;
  {
    std::lock_guard<std::mutex> g(mu_);
    send_frame(fd_, req);
    if (!recv_frame(fd_, &rep)) throw std::runtime_error("ps: server closed the connection");
  }
  if (rep.empty()) throw std::runtime_error("ps: empty reply");
  if (rep[0] != 0) {
    Reader r(rep);
    r.get<uint8_t>();
    throw std::runtime_error("ps server: " + r.str());
  }
  return rep;
}

static Writer header(Op op, int worker, int64_t key) {
  Writer w;
  w.put<uint8_t>(op); w.put<int32_t>(worker); w.put<int64_t>(key);
  return w;
}
static Reader body(const std::string& rep) { Reader r(rep); r.get<uint8_t>(); return r; }

void PsNetClient::init_dense(int64_t key, const std::vector<float>& v, const PsParamConfig& c) { auto w = header(INIT_DENSE, 0, key); w.cfg(c); w.arr(v); roundtrip(w.b); }
void PsNetClient::push_dense(int64_t key, const std::vector<float>& g) { auto w = header(PUSH_DENSE, 0, key); w.arr(g); roundtrip(w.b); }
std::vector<float> PsNetClient::pull_dense(int64_t key) { auto rep = roundtrip(header(PULL_DENSE, 0, key).b); return body(rep).arr<float>(); }
std::vector<float> PsNetClient::push_pull_dense(int64_t key, const std::vector<float>& g) {
  auto w = header(PUSH_PULL_DENSE, 0, key); w.arr(g);
  auto rep = roundtrip(w.b);
  return body(rep).arr<float>();
}
void PsNetClient::init_sparse(int64_t key, int64_t rows, int width, const std::vector<float>& v, const PsParamConfig& c) {
  auto w = header(INIT_SPARSE, 0, key); w.cfg(c); w.put<int64_t>(rows); w.put<int32_t>(width); w.arr(v); roundtrip(w.b);
}
void PsNetClient::push_sparse(int64_t key, const std::vector<int64_t>& rows, const std::vector<float>& g) {
  auto w = header(PUSH_SPARSE, 0, key); w.arr(rows); w.arr(g); roundtrip(w.b);
}
std::vector<float> PsNetClient::pull_sparse(int64_t key, const std::vector<int64_t>& rows) {
  auto w = header(PULL_SPARSE, 0, key); w.arr(rows);
  auto rep = roundtrip(w.b);
  return body(rep).arr<float>();
}
std::vector<int64_t> PsNetClient::row_versions(int64_t key, const std::vector<int64_t>& rows) {
  auto w = header(ROW_VERSIONS, 0, key); w.arr(rows);
  auto rep = roundtrip(w.b);
  return body(rep).arr<int64_t>();
}
void PsNetClient::sync_cache(int64_t key, const std::vector<int64_t>& rows, const std::vector<int64_t>& vers, int64_t bound,
                             std::vector<int64_t>* stale, std::vector<float>* vals, std::vector<int64_t>* fv) {
  auto w = header(SYNC_CACHE, 0, key); w.arr(rows); w.arr(vers); w.put<int64_t>(bound);
  auto rep = roundtrip(w.b);
  Reader r = body(rep);
  *stale = r.arr<int64_t>(); *vals = r.arr<float>(); *fv = r.arr<int64_t>();
}
void PsNetClient::barrier(int worker) { roundtrip(header(BARRIER, worker, 0).b); }
void PsNetClient::ssp_init(int s) { auto w = header(SSP_INIT, 0, 0); w.put<int32_t>(s); roundtrip(w.b); }
void PsNetClient::ssp_sync(int worker, int clock) { auto w = header(SSP_SYNC, worker, 0); w.put<int32_t>(clock); roundtrip(w.b); }
std::vector<float> PsNetClient::preduce(int worker, int64_t key, const std::vector<float>& v, int min_workers, int wait_ms, std::vector<int>* partners) {
  auto w = header(PREDUCE, worker, key); w.put<int32_t>(min_workers); w.put<int32_t>(wait_ms); w.arr(v);
  auto rep = roundtrip(w.b);
  Reader r = body(rep);
  auto out = r.arr<float>();
  auto p = r.arr<int32_t>();
  partners->assign(p.begin(), p.end());
  return out;
}
std::map<std::string, int64_t> PsNetClient::stats() {
  auto rep = roundtrip(header(STATS, 0, 0).b);
  Reader r = body(rep);
  std::map<std::string, int64_t> m;
  const uint64_t n = r.get<uint64_t>();
  for (uint64_t i = 0; i < n; ++i) { std::string k = r.str(); m[k] = r.get<int64_t>(); }
  return m;
}
int PsNetClient::num_workers() { auto rep = roundtrip(header(NUM_WORKERS, 0, 0).b); return body(rep).get<int32_t>(); }

}  // namespace hb
