// Wire format shared by the parameter-server transport and the scheduler: length-prefixed binary frames, little-endian
// scalars, arrays as u64 count + raw data.
#pragma once
#include <sys/socket.h>
#include <unistd.h>

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "ps_server.h"

namespace hb {
namespace ps_wire {

struct Writer {
  std::string b;
  template <typename T> void put(T v) { b.append(reinterpret_cast<const char*>(&v), sizeof v); }
  template <typename T> void arr(const std::vector<T>& v) {
    put<uint64_t>(v.size());
    if (!v.empty()) b.append(reinterpret_cast<const char*>(v.data()), v.size() * sizeof(T));
  }
  void str(const std::string& s) { put<uint64_t>(s.size()); b += s; }
  void cfg(const PsParamConfig& c) { put<int32_t>((int32_t)c.opt); put(c.lr); put(c.momentum); put(c.beta1); put(c.beta2); put(c.eps); }
};
struct Reader {
  const std::string& b;
  size_t i = 0;
  explicit Reader(const std::string& s) : b(s) {}
  template <typename T> T get() {
    if (i + sizeof(T) > b.size()) throw std::runtime_error("ps: truncated frame");
    T v;
    memcpy(&v, b.data() + i, sizeof v);
    i += sizeof v;
    return v;
  }
  template <typename T> std::vector<T> arr() {
    const uint64_t n = get<uint64_t>();
    if (i + n * sizeof(T) > b.size()) throw std::runtime_error("ps: truncated array");
    std::vector<T> v(n);
    if (n) memcpy(v.data(), b.data() + i, n * sizeof(T));
    i += n * sizeof(T);
    return v;
  }
  std::string str() {
    const uint64_t n = get<uint64_t>();
    std::string s = b.substr(i, n);
    i += n;
    return s;
  }
  PsParamConfig cfg() {
    PsParamConfig c;
    c.opt = (PsOptimizer)get<int32_t>(); c.lr = get<float>(); c.momentum = get<float>(); c.beta1 = get<float>(); c.beta2 = get<float>(); c.eps = get<float>();
    return c;
  }
};

inline void write_all(int fd, const char* p, size_t n) {
  while (n) {
    ssize_t w = ::send(fd, p, n, MSG_NOSIGNAL);
    if (w <= 0) throw std::runtime_error("ps: connection lost");
    p += w; n -= (size_t)w;
  }
}
inline bool read_all(int fd, char* p, size_t n) {
  while (n) {
    ssize_t r = ::recv(fd, p, n, 0);
    if (r <= 0) return false;
    p += r; n -= (size_t)r;
  }
  return true;
}
inline void send_frame(int fd, const std::string& s) {
  uint32_t n = (uint32_t)s.size();
  std::string buf(reinterpret_cast<const char*>(&n), 4);
  buf += s;
  write_all(fd, buf.data(), buf.size());
}
inline bool recv_frame(int fd, std::string* s) {
  uint32_t n = 0;
  if (!read_all(fd, reinterpret_cast<char*>(&n), 4)) return false;
  s->assign(n, '\0');
  return n == 0 || read_all(fd, &(*s)[0], n);
}


}  // namespace ps_wire
}  // namespace hb
