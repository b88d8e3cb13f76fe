#include "rpc_client.h"

#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <random>
#include <sstream>
#include <stdexcept>

namespace hb {

// ---------------------------------------------------------------------------------------------------------------- JSON
std::string json_quote(const std::string& s) {
  std::string o = "\"";
  for (unsigned char c : s) {
    switch (c) {
      case '"': o += "\\\""; break;
      case '\\': o += "\\\\"; break;
      case '\n': o += "\\n"; break;
      case '\r': o += "\\r"; break;
      case '\t': o += "\\t"; break;
      default:
        if (c < 0x20) {
          char buf[8];
          snprintf(buf, sizeof buf, "\\u%04x", c);
          o += buf;
        } else {
          o += (char)c;
        }
    }
  }
  return o + "\"";
}

static void append_utf8(std::string& o, unsigned cp) {
  if (cp < 0x80) o += (char)cp;
  else if (cp < 0x800) { o += (char)(0xC0 | (cp >> 6)); o += (char)(0x80 | (cp & 0x3F)); }
  else if (cp < 0x10000) { o += (char)(0xE0 | (cp >> 12)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
  else { o += (char)(0xF0 | (cp >> 18)); o += (char)(0x80 | ((cp >> 12) & 0x3F)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
}

std::string json_unquote(const std::string& s) {
  if (s.size() < 2 || s.front() != '"') return s;
  std::string o;
  for (size_t i = 1; i + 1 < s.size(); ++i) {
    char c = s[i];
    if (c != '\\') { o += c; continue; }
    char e = s[++i];
    switch (e) {
      case 'n': o += '\n'; break;
      case 'r': o += '\r'; break;
      case 't': o += '\t'; break;
      case 'b': o += '\b'; break;
      case 'f': o += '\f'; break;
      case 'u': {
        unsigned cp = (unsigned)std::stoul(s.substr(i + 1, 4), nullptr, 16);
        i += 4;
        if (cp >= 0xD800 && cp < 0xDC00 && i + 6 < s.size() && s[i + 1] == '\\' && s[i + 2] == 'u') {   // surrogate pair
          unsigned lo = (unsigned)std::stoul(s.substr(i + 3, 4), nullptr, 16);
          cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
          i += 6;
        }
        append_utf8(o, cp);
        break;
      }
      default: o += e;   // \" \\ \/
    }
  }
  return o;
}

// index one past the JSON value that starts at s[i]
static size_t skip_value(const std::string& s, size_t i) {
  while (i < s.size() && isspace((unsigned char)s[i])) ++i;
  if (i >= s.size()) throw std::runtime_error("json: truncated");
  if (s[i] == '"') {
    for (++i; i < s.size(); ++i) {
      if (s[i] == '\\') ++i;
      else if (s[i] == '"') return i + 1;
    }
    throw std::runtime_error("json: unterminated string");
  }
  if (s[i] == '{' || s[i] == '[') {
    int depth = 0;
    for (; i < s.size(); ++i) {
      if (s[i] == '"') { i = skip_value(s, i) - 1; continue; }
      if (s[i] == '{' || s[i] == '[') ++depth;
      if (s[i] == '}' || s[i] == ']') { if (--depth == 0) return i + 1; }
    }
    throw std::runtime_error("json: unbalanced");
  }
  while (i < s.size() && s[i] != ',' && s[i] != '}' && s[i] != ']' && !isspace((unsigned char)s[i])) ++i;
  return i;
}

bool json_field(const std::string& obj, const std::string& key, std::string* raw) {
  size_t i = 0;
  while (i < obj.size() && obj[i] != '{') ++i;
  ++i;
  while (i < obj.size()) {
    while (i < obj.size() && (isspace((unsigned char)obj[i]) || obj[i] == ',')) ++i;
    if (i >= obj.size() || obj[i] == '}') return false;
    size_t ke = skip_value(obj, i);
    std::string k = json_unquote(obj.substr(i, ke - i));
    i = ke;
    while (i < obj.size() && (isspace((unsigned char)obj[i]) || obj[i] == ':')) ++i;
    size_t ve = skip_value(obj, i);
    if (k == key) { *raw = obj.substr(i, ve - i); return true; }
    i = ve;
  }
  return false;
}

static const char kB64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
std::string base64_encode(const std::string& in) {
  std::string o;
  size_t i = 0;
  for (; i + 2 < in.size(); i += 3) {
    unsigned v = ((unsigned char)in[i] << 16) | ((unsigned char)in[i + 1] << 8) | (unsigned char)in[i + 2];
    o += kB64[v >> 18]; o += kB64[(v >> 12) & 63]; o += kB64[(v >> 6) & 63]; o += kB64[v & 63];
  }
  if (i + 1 == in.size()) {
    unsigned v = (unsigned char)in[i] << 16;
    o += kB64[v >> 18]; o += kB64[(v >> 12) & 63]; o += "==";
  } else if (i + 2 == in.size()) {
    unsigned v = ((unsigned char)in[i] << 16) | ((unsigned char)in[i + 1] << 8);
    o += kB64[v >> 18]; o += kB64[(v >> 12) & 63]; o += kB64[(v >> 6) & 63]; o += '=';
  }
  return o;
}
std::string base64_decode(const std::string& in) {
  int T[256];
  std::fill(T, T + 256, -1);
  for (int i = 0; i < 64; ++i) T[(unsigned char)kB64[i]] = i;
  std::string o;
  unsigned acc = 0;
  int bits = 0;
  for (unsigned char c : in) {
    if (T[c] < 0) continue;
    acc = (acc << 6) | (unsigned)T[c];
    bits += 6;
    if (bits >= 8) { bits -= 8; o += (char)((acc >> bits) & 0xFF); }
  }
  return o;
}

// -------------------------------------------------------------------------------------------------------------- socket
static void write_all(int fd, const char* p, size_t n) {
  while (n) {
    ssize_t w = ::send(fd, p, n, MSG_NOSIGNAL);
    if (w <= 0) throw std::runtime_error("rpc: connection lost while sending");
    p += w; n -= (size_t)w;
  }
}
static void read_all(int fd, char* p, size_t n) {
  while (n) {
    ssize_t r = ::recv(fd, p, n, 0);
    if (r <= 0) throw std::runtime_error("rpc: peer closed the connection");
    p += r; n -= (size_t)r;
  }
}
void RpcClient::send_frame(int fd, const std::string& payload) {
  uint32_t n = htonl((uint32_t)payload.size());
  std::string buf((const char*)&n, 4);
  buf += payload;
  write_all(fd, buf.data(), buf.size());
}
std::string RpcClient::recv_frame(int fd) {
  uint32_t n = 0;
  read_all(fd, (char*)&n, 4);
  std::string s(ntohl(n), '\0');
  if (!s.empty()) read_all(fd, &s[0], s.size());
  return s;
}

int RpcClient::open_socket(double timeout_s) const {
  auto deadline = std::chrono::steady_clock::now() + std::chrono::duration<double>(timeout_s);
  for (;;) {
    addrinfo hints{}, *res = nullptr;
    hints.ai_family = AF_UNSPEC;
    hints.ai_socktype = SOCK_STREAM;
    if (getaddrinfo(host_.c_str(), std::to_string(port_).c_str(), &hints, &res) == 0) {
      for (addrinfo* a = res; a; a = a->ai_next) {
        int fd = ::socket(a->ai_family, a->ai_socktype, a->ai_protocol);
        if (fd < 0) continue;
        if (::connect(fd, a->ai_addr, a->ai_addrlen) == 0) {
          int one = 1;
          setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
          freeaddrinfo(res);
          return fd;
        }
        ::close(fd);
      }
      freeaddrinfo(res);
    }
    if (std::chrono::steady_clock::now() > deadline)
      throw std::runtime_error("rpc: cannot reach the device controller at " + host_ + ":" + std::to_string(port_));
    std::this_thread::sleep_for(std::chrono::milliseconds(200));
  }
}

// -------------------------------------------------------------------------------------------------------------- client
RpcClient::RpcClient(std::string host, int port, std::string hostname, double hb, double timeout)
    : host_(std::move(host)), hostname_(std::move(hostname)), port_(port), hb_interval_(hb), connect_timeout_(timeout) {
  if (hostname_.empty()) {
    char buf[256] = {0};
    gethostname(buf, sizeof buf - 1);
    hostname_ = buf;
  }
  std::random_device rd;
  std::ostringstream id;
  id << std::hex << rd() << rd() << rd() << rd();
  client_id_ = id.str();
  fd_ = open_socket(connect_timeout_);
}

RpcClient::~RpcClient() {
  hb_stop_ = true;
  if (hb_fd_ >= 0) ::shutdown(hb_fd_, SHUT_RDWR);
  if (hb_thread_.joinable()) hb_thread_.join();
  if (hb_fd_ >= 0) ::close(hb_fd_);
  if (fd_ >= 0) ::close(fd_);
}

std::string RpcClient::call(const std::string& method, const std::string& args_json) {
  std::string reply;
  {
    std::lock_guard<std::mutex> g(mu_);
    send_frame(fd_, "{\"method\": " + json_quote(method) + ", \"args\": " + (args_json.empty() ? "{}" : args_json) + "}");
    reply = recv_frame(fd_);
  }
  std::string ok, value;
  if (!json_field(reply, "ok", &ok) || ok != "true") {
    std::string err;
    json_field(reply, "error", &err);
    throw std::runtime_error("rpc " + method + " failed: " + json_unquote(err));
  }
  json_field(reply, "value", &value);
  return value;
}

void RpcClient::connect(bool start_heartbeat) {
  const std::string id = "{\"client_id\": " + json_quote(client_id_);
  call("Connect", id + ", \"hostname\": " + json_quote(hostname_) + "}");
  std::string info = call("GetRank", id + "}"), v;
  if (json_field(info, "rank", &v)) rank_ = std::stoi(v);
  if (json_field(info, "local_device", &v)) local_device_ = std::stoi(v);
  if (json_field(info, "world_size", &v)) world_size_ = std::stoi(v);
  if (start_heartbeat && !hb_thread_.joinable()) {
    hb_fd_ = open_socket(10.0);
    hb_thread_ = std::thread([this] { heartbeat_loop(); });
  }
}

void RpcClient::heartbeat_loop() {
  const auto step = std::chrono::milliseconds(20);
  while (!hb_stop_) {
    auto until = std::chrono::steady_clock::now() + std::chrono::duration<double>(hb_interval_);
    while (!hb_stop_ && std::chrono::steady_clock::now() < until) std::this_thread::sleep_for(step);
    if (hb_stop_) return;
    try {
      send_frame(hb_fd_, "{\"method\": \"HeartBeat\", \"args\": {\"rank\": " + std::to_string(rank_) + "}}");
      std::string rep = recv_frame(hb_fd_), v;
      beats_.fetch_add(1);
      if (json_field(rep, "value", &v) && v == "true") stop_requested_ = true;
    } catch (const std::exception&) {
      return;   // controller gone: the elastic layer notices through its own monitor
    }
  }
}

static std::string kv_args(const std::string& key, const std::string& kind, const std::string* raw_value) {
  std::string a = "{\"key\": " + json_quote(key) + ", \"kind\": " + json_quote(kind);
  if (raw_value) a += ", \"value\": " + *raw_value;
  return a + "}";
}
void RpcClient::put_int(const std::string& k, int64_t v) { std::string r = std::to_string(v); call("Put", kv_args(k, "int", &r)); }
int64_t RpcClient::get_int(const std::string& k) { return std::stoll(call("Get", kv_args(k, "int", nullptr))); }
void RpcClient::put_double(const std::string& k, double v) {
  char buf[64];
  snprintf(buf, sizeof buf, "%.17g", v);
  std::string r = buf;
  if (r.find_first_of(".eEn") == std::string::npos) r += ".0";
  call("Put", kv_args(k, "double", &r));
}
double RpcClient::get_double(const std::string& k) { return std::stod(call("Get", kv_args(k, "double", nullptr))); }
void RpcClient::put_string(const std::string& k, const std::string& v) { std::string r = json_quote(v); call("Put", kv_args(k, "string", &r)); }
std::string RpcClient::get_string(const std::string& k) { return json_unquote(call("Get", kv_args(k, "string", nullptr))); }
void RpcClient::put_bytes(const std::string& k, const std::string& v) { std::string r = json_quote(base64_encode(v)); call("Put", kv_args(k, "bytes", &r)); }
std::string RpcClient::get_bytes(const std::string& k) { return base64_decode(json_unquote(call("Get", kv_args(k, "bytes", nullptr)))); }
void RpcClient::put_json(const std::string& k, const std::string& raw) { call("Put", kv_args(k, "json", &raw)); }
std::string RpcClient::get_json(const std::string& k) { return call("Get", kv_args(k, "json", nullptr)); }
bool RpcClient::remove(const std::string& k, const std::string& kind) { return call("Remove", kv_args(k, kind, nullptr)) == "true"; }

void RpcClient::commit_hostname() {
  call("CommitHostName", "{\"rank\": " + std::to_string(rank_) + ", \"hostname\": " + json_quote(hostname_) + "}");
}
std::string RpcClient::get_hostname(int rank) { return json_unquote(call("GetHostName", "{\"rank\": " + std::to_string(rank) + "}")); }

static std::string nccl_key(std::vector<int> ranks, int stream) {
  std::sort(ranks.begin(), ranks.end());
  std::string k = "[";                        // same text as Python's f"{sorted(ranks)}:{stream}"
  for (size_t i = 0; i < ranks.size(); ++i) k += (i ? ", " : "") + std::to_string(ranks[i]);
  return k + "]:" + std::to_string(stream);
}
void RpcClient::commit_nccl_id(std::vector<int> ranks, int stream, const std::string& id) {
  call("CommitNcclId", "{\"key\": " + json_quote(nccl_key(std::move(ranks), stream)) + ", \"nccl_id\": " + json_quote(base64_encode(id)) + "}");
}
std::string RpcClient::get_nccl_id(std::vector<int> ranks, int stream) {
  return base64_decode(json_unquote(call("GetNcclId", "{\"key\": " + json_quote(nccl_key(std::move(ranks), stream)) + "}")));
}

std::string RpcClient::ranks_json(const std::vector<int>& ranks) {
  if (ranks.empty()) return "null";
  std::string s = "[";
  for (size_t i = 0; i < ranks.size(); ++i) s += (i ? ", " : "") + std::to_string(ranks[i]);
  return s + "]";
}
void RpcClient::barrier(const std::vector<int>& ranks, const std::string& tag) {
  call("Barrier", "{\"rank\": " + std::to_string(rank_) + ", \"world_ranks\": " + ranks_json(ranks) + ", \"tag\": " + json_quote(tag) + "}");
}
bool RpcClient::consistent(const std::string& raw, const std::vector<int>& ranks, const std::string& tag) {
  return call("Consistent", "{\"rank\": " + std::to_string(rank_) + ", \"value\": " + raw + ", \"world_ranks\": " + ranks_json(ranks) +
                                ", \"tag\": " + json_quote(tag) + "}") == "true";
}
void RpcClient::worker_stop() { call("WorkerStop", "{}"); }
bool RpcClient::already_stop() { return stop_requested_ || call("AlreadyStop", "{}") == "true"; }
void RpcClient::exit() {
  hb_stop_ = true;
  call("Exit", "{\"rank\": " + std::to_string(rank_) + "}");
}

}  // namespace hb
