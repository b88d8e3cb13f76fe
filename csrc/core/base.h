// Logging, error handling and small utilities shared by the whole native core.
// (capability parity: hetu/common/{logging,except,timing}.h)
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace hb {

enum class LogLevel : int { TRACE = 0, DEBUG = 1, INFO = 2, WARN = 3, ERROR = 4, FATAL = 5 };

// Level comes from HETU_INTERNAL_LOG_LEVEL (same knob as the reference), default WARN.
LogLevel log_level();
void set_log_level(LogLevel l);
void set_log_prefix(const std::string& prefix);  // e.g. "[cuda:3]"
const std::string& log_prefix();

class LogMessage {
 public:
  LogMessage(LogLevel lvl, const char* file, int line);
  ~LogMessage();
  std::ostream& stream() { return os_; }

 private:
  LogLevel lvl_;
  std::ostringstream os_;
};

#define HB_LOG(LVL)                                         \
  if (::hb::LogLevel::LVL < ::hb::log_level()) {            \
  } else                                                    \
    ::hb::LogMessage(::hb::LogLevel::LVL, __FILE__, __LINE__).stream()

template <typename T>
std::ostream& operator<<(std::ostream& os, const std::vector<T>& v) {
  os << "[";
  for (size_t i = 0; i < v.size(); ++i) {
    if (i) os << ", ";
    os << v[i];
  }
  return os << "]";
}

// "\n  at <module>+0x<offset> ..." for the calling frames (resolve with addr2line -e <module> <offset>); empty unless
// HETU_BACKTRACE=1.  Makes a failed check on a remote GPU box diagnosable without a debugger.
std::string capture_backtrace();

class Error : public std::runtime_error {
 public:
  explicit Error(const std::string& m) : std::runtime_error(m) {}
};

// Stream-style exception builder:  HB_CHECK(x > 0) << "x must be positive, got " << x;
class ErrorBuilder {
 public:
  ErrorBuilder(const char* file, int line, const char* cond) {
    os_ << file << ":" << line << ": check failed: " << cond << " ";
  }
  template <typename T>
  ErrorBuilder& operator<<(const T& v) {
    os_ << v;
    return *this;
  }
  [[noreturn]] ~ErrorBuilder() noexcept(false) { throw Error(os_.str() + capture_backtrace()); }

 private:
  std::ostringstream os_;
};

#define HB_CHECK(cond) \
  if (cond) {          \
  } else               \
    ::hb::ErrorBuilder(__FILE__, __LINE__, #cond)
#define HB_FAIL() ::hb::ErrorBuilder(__FILE__, __LINE__, "unreachable")

inline double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

inline std::string env_str(const char* name, const std::string& dflt = "") {
  const char* v = std::getenv(name);
  return v ? std::string(v) : dflt;
}
inline int64_t env_int(const char* name, int64_t dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoll(v) : dflt;
}

}  // namespace hb
