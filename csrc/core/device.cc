#include "device.h"
#include <dlfcn.h>
#include <execinfo.h>
#include <cstring>

#include <cstring>
#include <mutex>

namespace hb {

// ------------------------------------------------------------------ logging
static LogLevel g_level = [] {
  std::string s = env_str("HETU_INTERNAL_LOG_LEVEL", "WARN");
  for (auto& c : s) c = toupper(c);
  if (s == "TRACE") return LogLevel::TRACE;
  if (s == "DEBUG") return LogLevel::DEBUG;
  if (s == "INFO") return LogLevel::INFO;
  if (s == "WARN" || s == "WARNING") return LogLevel::WARN;
  if (s == "ERROR") return LogLevel::ERROR;
  if (s == "FATAL") return LogLevel::FATAL;
  return LogLevel::WARN;
}();
static std::string g_prefix;
static std::mutex g_log_mu;

LogLevel log_level() { return g_level; }

std::string capture_backtrace() {
  const char* e = std::getenv("HETU_BACKTRACE");
  if (!e || e[0] == '0') return "";
  void* frames[48];
  const int n = ::backtrace(frames, 48);
  std::ostringstream os;
  os << "\n  backtrace:";
  for (int i = 1; i < n; ++i) {
    Dl_info info;
    if (dladdr(frames[i], &info) && info.dli_fname) {
      const char* base = std::strrchr(info.dli_fname, '/');
      os << "\n    " << (base ? base + 1 : info.dli_fname) << "+0x" << std::hex
         << (uintptr_t)((char*)frames[i] - (char*)info.dli_fbase) << std::dec;
      if (info.dli_sname) os << " (" << info.dli_sname << ")";
    }
  }
  return os.str();
}
void set_log_level(LogLevel l) { g_level = l; }
void set_log_prefix(const std::string& p) { g_prefix = p; }
const std::string& log_prefix() { return g_prefix; }

LogMessage::LogMessage(LogLevel lvl, const char* file, int line) : lvl_(lvl) {
  static const char* names[] = {"TRACE", "DEBUG", "INFO", "WARN", "ERROR", "FATAL"};
  const char* base = std::strrchr(file, '/');
  os_ << "[" << names[int(lvl)] << "]" << g_prefix << " " << (base ? base + 1 : file) << ":" << line << " ";
}
LogMessage::~LogMessage() {
  std::lock_guard<std::mutex> lk(g_log_mu);
  std::cerr << os_.str() << std::endl;
  if (lvl_ == LogLevel::FATAL) std::abort();
}

// ------------------------------------------------------------------ Device
Device::Device(const std::string& spec_in) : type_(DeviceType::UNDETERMINED), index_(0), multiplex_(0) {
  std::string spec = spec_in;
  auto slash = spec.find('/');
  if (slash != std::string::npos) {
    host_ = spec.substr(0, slash);
    spec = spec.substr(slash + 1);
  }
  auto hashp = spec.find('#');
  if (hashp != std::string::npos) {
    multiplex_ = std::atoi(spec.substr(hashp + 1).c_str());
    spec = spec.substr(0, hashp);
  }
  std::string ty = spec;
  auto colon = spec.find(':');
  if (colon != std::string::npos) {
    ty = spec.substr(0, colon);
    index_ = std::atoi(spec.substr(colon + 1).c_str());
  }
  for (auto& c : ty) c = tolower(c);
  if (ty == "cuda" || ty == "gpu") type_ = DeviceType::CUDA;
  else if (ty == "cpu") type_ = DeviceType::CPU;
  else HB_FAIL() << "cannot parse device spec '" << spec_in << "'";
  if (host_ == "localhost") host_.clear();
}

bool Device::local() const {
  if (host_.empty()) return true;
  return host_ == env_str("HETU_LOCAL_HOSTNAME", "");
}

std::string Device::str() const {
  std::ostringstream os;
  if (!host_.empty()) os << host_ << "/";
  os << (type_ == DeviceType::CUDA ? "cuda" : type_ == DeviceType::CPU ? "cpu" : "undetermined");
  if (type_ == DeviceType::CUDA || index_ != 0) os << ":" << index_;
  if (multiplex_) os << "#" << multiplex_;
  return os.str();
}
std::ostream& operator<<(std::ostream& os, const Device& d) { return os << d.str(); }

std::string DeviceGroup::str() const {
  std::ostringstream os;
  os << "DeviceGroup(";
  for (size_t i = 0; i < devs_.size(); ++i) os << (i ? ", " : "") << devs_[i];
  os << ")";
  return os.str();
}
std::ostream& operator<<(std::ostream& os, const DeviceGroup& g) { return os << g.str(); }

DeviceGroupUnion DeviceGroupUnion::merge(const DeviceGroupUnion& a, const DeviceGroupUnion& b) {
  HB_CHECK(a.size() == b.size()) << "cannot merge device group unions of different size";
  std::vector<DeviceGroup> out;
  for (size_t i = 0; i < a.size(); ++i) {
    std::vector<Device> v = a.get(i).devices();
    for (auto& d : b.get(i).devices()) if (std::find(v.begin(), v.end(), d) == v.end()) v.push_back(d);
    out.emplace_back(v);
  }
  return DeviceGroupUnion(out);
}

DeviceGroupUnion DeviceGroupUnion::device_group_to_union(const DeviceGroup& g, int /*hint*/, int num) {
  HB_CHECK(num > 0 && g.num_devices() % num == 0) << "cannot split " << g.num_devices() << " devices into " << num;
  const size_t per = g.num_devices() / num;
  std::vector<DeviceGroup> out;
  for (int i = 0; i < num; ++i) {
    std::vector<Device> v(g.devices().begin() + i * per, g.devices().begin() + (i + 1) * per);
    out.emplace_back(v);
  }
  return DeviceGroupUnion(out);
}

// ------------------------------------------------------------------ dtypes
size_t dtype_size(DataType t) {
  switch (t) {
    case DataType::UINT8: case DataType::INT8: case DataType::BOOL: case DataType::FLOAT4: case DataType::NFLOAT4:
    case DataType::FLOAT8_E4M3: case DataType::FLOAT8_E5M2: return 1;
    case DataType::INT16: case DataType::FLOAT16: case DataType::BFLOAT16: return 2;
    case DataType::INT32: case DataType::FLOAT32: return 4;
    case DataType::INT64: case DataType::FLOAT64: return 8;
    default: return 0;
  }
}
static const char* kDtypeNames[] = {"uint8", "int8", "int16", "int32", "int64", "float16", "float32", "float64",
                                    "bfloat16", "float4", "nfloat4", "bool", "float8_e4m3", "float8_e5m2",
                                    "undetermined"};
const char* dtype_name(DataType t) { return kDtypeNames[int(t)]; }
DataType dtype_from_name(const std::string& s) {
  for (int i = 0; i <= int(DataType::UNDETERMINED); ++i) if (s == kDtypeNames[i]) return DataType(i);
  if (s == "float" ) return DataType::FLOAT32;
  if (s == "double") return DataType::FLOAT64;
  if (s == "half") return DataType::FLOAT16;
  if (s == "long") return DataType::INT64;
  if (s == "int") return DataType::INT32;
  HB_FAIL() << "unknown dtype name " << s;
}

const char* reduction_name(ReductionType r) {
  static const char* n[] = {"sum", "mean", "max", "min", "prod", "none"};
  return n[int(r)];
}
ReductionType reduction_from_name(const std::string& s_in) {
  std::string s = s_in;
  for (auto& c : s) c = tolower(c);
  if (s == "sum") return ReductionType::SUM;
  if (s == "mean" || s == "avg") return ReductionType::MEAN;
  if (s == "max") return ReductionType::MAX;
  if (s == "min") return ReductionType::MIN;
  if (s == "prod") return ReductionType::PROD;
  if (s == "none") return ReductionType::NONE;
  HB_FAIL() << "unknown reduction " << s_in;
}

}  // namespace hb
