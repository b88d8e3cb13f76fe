// Device / DeviceGroup / unions and hierarchies, DataType, Stream roles.
// (capability parity: hetu/core/{device,dtype,stream}.h and the DeviceGroupUnion /
//  DeviceGroupHierarchy classes of hetu/graph/distributed_states.h:360-606)
#pragma once
#include <algorithm>
#include <functional>
#include <string>
#include <vector>

#include "base.h"

namespace hb {

enum class DeviceType : int8_t { CPU = 0, CUDA = 1, UNDETERMINED = 2 };

class Device {
 public:
  Device() : type_(DeviceType::UNDETERMINED), index_(0), multiplex_(0) {}
  Device(DeviceType t, int index = 0, const std::string& host = "", int multiplex = 0)
      : type_(t), index_(t == DeviceType::CPU && index < 0 ? 0 : index), multiplex_(multiplex), host_(host) {}
  // "cuda:3", "cpu", "cpu:1", "host1/cuda:3", "cuda:3#1" (multiplex)
  explicit Device(const std::string& spec);

  DeviceType type() const { return type_; }
  int index() const { return index_; }
  int multiplex() const { return multiplex_; }
  const std::string& hostname() const { return host_; }
  bool is_cpu() const { return type_ == DeviceType::CPU; }
  bool is_cuda() const { return type_ == DeviceType::CUDA; }
  bool is_undetermined() const { return type_ == DeviceType::UNDETERMINED; }
  bool local() const;  // hostname empty or equal to HETU_LOCAL_HOSTNAME
  std::string str() const;

  bool operator==(const Device& o) const {
    return type_ == o.type_ && index_ == o.index_ && multiplex_ == o.multiplex_ && host_ == o.host_;
  }
  bool operator!=(const Device& o) const { return !(*this == o); }
  bool operator<(const Device& o) const {
    if (host_ != o.host_) return host_ < o.host_;
    if (type_ != o.type_) return type_ < o.type_;
    if (index_ != o.index_) return index_ < o.index_;
    return multiplex_ < o.multiplex_;
  }
  size_t hash() const {
    return std::hash<std::string>()(host_) ^ (size_t(type_) << 20) ^ (size_t(index_) << 4) ^ size_t(multiplex_);
  }

 private:
  DeviceType type_;
  int index_;
  int multiplex_;
  std::string host_;
};
std::ostream& operator<<(std::ostream& os, const Device& d);

// Ordered list of devices (the order matters for heterogeneous pipelines).
class DeviceGroup {
 public:
  DeviceGroup() = default;
  explicit DeviceGroup(std::vector<Device> devs) : devs_(std::move(devs)) {}
  explicit DeviceGroup(const std::vector<std::string>& specs) {
    for (auto& s : specs) devs_.emplace_back(s);
  }
  size_t num_devices() const { return devs_.size(); }
  bool empty() const { return devs_.empty(); }
  bool contains(const Device& d) const { return std::find(devs_.begin(), devs_.end(), d) != devs_.end(); }
  const Device& get(size_t i) const {
    HB_CHECK(i < devs_.size()) << "device index " << i << " out of range " << devs_.size();
    return devs_[i];
  }
  int get_index(const Device& d) const {
    auto it = std::find(devs_.begin(), devs_.end(), d);
    return it == devs_.end() ? -1 : int(it - devs_.begin());
  }
  const std::vector<Device>& devices() const { return devs_; }
  bool operator==(const DeviceGroup& o) const { return devs_ == o.devs_; }
  bool operator!=(const DeviceGroup& o) const { return !(*this == o); }
  bool is_subset(const DeviceGroup& o) const {
    for (auto& d : devs_) if (!o.contains(d)) return false;
    return true;
  }
  std::string str() const;

 private:
  std::vector<Device> devs_;
};
std::ostream& operator<<(std::ostream& os, const DeviceGroup& g);

// One DeviceGroup per heterogeneous sub-group (e.g. pipelines with different TP degree).
class DeviceGroupUnion {
 public:
  DeviceGroupUnion() = default;
  explicit DeviceGroupUnion(std::vector<DeviceGroup> u) : union_(std::move(u)) {}
  size_t size() const { return union_.size(); }
  const DeviceGroup& get(size_t i) const {
    HB_CHECK(i < union_.size()) << "union index out of range";
    return union_[i];
  }
  void add(const DeviceGroup& g) { union_.push_back(g); }
  const std::vector<DeviceGroup>& raw() const { return union_; }
  bool has(const Device& d) const {
    for (auto& g : union_) if (g.contains(d)) return true;
    return false;
  }
  // index of the sub-group containing d (-1 when absent)
  int get_index(const Device& d) const {
    for (size_t i = 0; i < union_.size(); ++i) if (union_[i].contains(d)) return int(i);
    return -1;
  }
  DeviceGroup all() const {
    std::vector<Device> v;
    for (auto& g : union_) for (auto& d : g.devices()) v.push_back(d);
    return DeviceGroup(v);
  }
  bool operator==(const DeviceGroupUnion& o) const { return union_ == o.union_; }
  static DeviceGroupUnion merge(const DeviceGroupUnion& a, const DeviceGroupUnion& b);
  // Re-partition the same devices into `num` sub-groups (used when the hetero dim changes)
  static DeviceGroupUnion device_group_to_union(const DeviceGroup& g, int device_num_per_group_hint, int num);

 private:
  std::vector<DeviceGroup> union_;
};

// One union per parallel strategy (hot switching keeps several strategies alive).
class DeviceGroupHierarchy {
 public:
  DeviceGroupHierarchy() = default;
  explicit DeviceGroupHierarchy(std::vector<DeviceGroupUnion> h) : h_(std::move(h)) {}
  size_t size() const { return h_.size(); }
  const DeviceGroupUnion& get(size_t i) const {
    HB_CHECK(i < h_.size()) << "strategy id " << i << " out of range " << h_.size();
    return h_[i];
  }
  void add(const DeviceGroupUnion& u) { h_.push_back(u); }
  const std::vector<DeviceGroupUnion>& raw() const { return h_; }

 private:
  std::vector<DeviceGroupUnion> h_;
};

// ------------------------------------------------------------------ dtypes
enum class DataType : int8_t {
  UINT8 = 0, INT8, INT16, INT32, INT64, FLOAT16, FLOAT32, FLOAT64, BFLOAT16, FLOAT4, NFLOAT4, BOOL,
  FLOAT8_E4M3, FLOAT8_E5M2, UNDETERMINED
};
size_t dtype_size(DataType t);          // bytes per element (4-bit types report 1: two per byte handled by ops)
const char* dtype_name(DataType t);
DataType dtype_from_name(const std::string& s);
inline bool dtype_is_float(DataType t) {
  return t == DataType::FLOAT16 || t == DataType::FLOAT32 || t == DataType::FLOAT64 || t == DataType::BFLOAT16 ||
         t == DataType::FLOAT8_E4M3 || t == DataType::FLOAT8_E5M2;
}

// ------------------------------------------------------------------ stream roles
// Fixed logical streams per device (same role table as hetu/core/stream.h:7-20).
enum StreamIndex : int {
  kBlockingStream = 0, kComputingStream = 1, kSwitchComputingStream = 2, kH2DStream = 3, kD2HStream = 4,
  kP2PStream = 5, kCollectiveStream = 6, kSwitchCollectiveStream = 7, kBridgeStream = 8, kOffloadStream = 9,
  kJoinStream = 15, kNumStreams = 16
};

enum class ReductionType : int8_t { SUM = 0, MEAN, MAX, MIN, PROD, NONE };
const char* reduction_name(ReductionType r);
ReductionType reduction_from_name(const std::string& s);

}  // namespace hb
