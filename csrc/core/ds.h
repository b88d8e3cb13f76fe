// DistributedStates algebra: how a logical tensor maps onto a device group.
//   states: {dim -> number of shards}, dim -1 = duplicate, dim -2 = partial (sum pending)
//   order : which state varies slowest..fastest over the device index
//   zero  : parameter whose optimizer copy is additionally sharded over the duplicate axis
// Union = one DS per heterogeneous sub-group (along `hetero_dim`);
// Hierarchy = one union per parallel strategy (hot switching).
//
// The data structure is a small sorted vector (not hash maps), the comm classifier is
// table driven, and everything is pure host logic that is unit-tested on CPU.
// (capability parity: hetu/graph/distributed_states.{h,cc}, hetu/graph/ops/Communication.cc:109-330)
#pragma once
#include <map>
#include <string>
#include <vector>

#include "device.h"

namespace hb {

constexpr int kDupDim = -1;
constexpr int kPartialDim = -2;
constexpr int kNullHeteroDim = -3;

class DistributedStates {
 public:
  DistributedStates() : device_num_(-1), zero_(false) {}
  DistributedStates(int device_num, const std::map<int, int>& states, const std::vector<int>& order = {},
                    bool zero = false);

  bool is_none() const { return device_num_ == -1; }
  bool is_valid() const { return device_num_ == 1 || (device_num_ > 1 && !order_.empty()); }
  int device_num() const { return device_num_; }
  bool zero() const { return zero_; }
  void set_zero(bool z) { zero_ = z; }
  // always contains -2 and -1 entries (value 1 when absent), plus split dims with n > 1
  const std::map<int, int>& states() const { return states_; }
  const std::vector<int>& order() const { return order_; }
  int get_dim(int dim) const {
    auto it = states_.find(dim);
    return it == states_.end() ? 1 : it->second;
  }
  int states(int dim) const { return get_dim(dim); }

  bool check_equal(const DistributedStates& o) const {
    return device_num_ == o.device_num_ && states_ == o.states_ && order_ == o.order_;
  }
  bool operator==(const DistributedStates& o) const { return check_equal(o); }
  bool check_pure_duplicate() const { return device_num_ == get_dim(kDupDim); }
  bool check_max_dim(int max_dim) const;

  // fold the src dims into dst (used to express "all-reduce turns partial into duplicate", ...)
  static std::map<int, int> combine_states(const std::vector<int>& src, int dst, const std::map<int, int>& states);
  static std::vector<int> combine_order(const std::vector<int>& src, int dst, const std::vector<int>& order);
  std::map<int, int> combine_states(const std::vector<int>& src, int dst) const { return combine_states(src, dst, states_); }
  std::vector<int> combine_order(const std::vector<int>& src, int dst) const { return combine_order(src, dst, order_); }
  bool check_combine(const DistributedStates& dst, const std::vector<int>& src, int dst_dim) const;
  std::map<int, int> reduce_states(int dim) const;
  std::vector<int> reduce_order(int dim) const;
  bool check_reduce_dim(const DistributedStates& dst, int dim) const;

  // relations between a source (this) and destination layout on the SAME device group
  bool check_split(const DistributedStates& dst) const;
  bool check_scatter(const DistributedStates& dst) const;
  bool check_allreduce(const DistributedStates& dst) const;
  bool check_allgather(const DistributedStates& dst) const;
  bool check_reducescatter(const DistributedStates& dst) const;
  bool check_broadcast(const DistributedStates& dst) const;
  bool check_reduce(const DistributedStates& dst) const;

  // the (single) tensor dim that is split more finely here than in `merged`; -3 when none
  int get_split_dim(const DistributedStates& merged) const;
  std::vector<int> get_loop_sizes() const;
  // device index -> {state dim -> coordinate}
  std::map<int, int> map_device_to_state_index(int device_index) const;
  int get_dup_group_index(int device_index) const;
  // devices that differ from `local_device_idx` only along `dim`
  std::vector<int> get_device_indices_by_dim(int dim, int local_device_idx) const;
  DeviceGroup get_devices_by_dim(int dim, int local_device_idx, const DeviceGroup& group) const;
  std::string str() const;

  // local shape of a tensor with this layout and the slice (offset, size) owned by a device
  std::vector<int64_t> local_shape(const std::vector<int64_t>& global_shape) const;
  std::vector<int64_t> global_shape(const std::vector<int64_t>& local_shape) const;
  void local_slice(const std::vector<int64_t>& global_shape, int device_index, std::vector<int64_t>* begin,
                   std::vector<int64_t>* size) const;

 private:
  void set_states(const std::map<int, int>& states);
  void set_order(const std::vector<int>& order);
  int device_num_;
  bool zero_;
  std::map<int, int> states_;
  std::vector<int> order_;
};
std::ostream& operator<<(std::ostream& os, const DistributedStates& ds);

class DistributedStatesUnion {
 public:
  DistributedStatesUnion() : hetero_dim_(kNullHeteroDim), contiguous_(true) {}
  explicit DistributedStatesUnion(std::vector<DistributedStates> u, int hetero_dim = kNullHeteroDim,
                                  bool contiguous = true)
      : union_(std::move(u)), hetero_dim_(hetero_dim), contiguous_(contiguous) {}
  size_t size() const { return union_.size(); }
  bool is_hetero() const { return hetero_dim_ != kNullHeteroDim; }
  int hetero_dim() const { return hetero_dim_; }
  void set_hetero_dim(int d) { hetero_dim_ = d; }
  bool contiguous() const { return contiguous_; }
  void set_contiguous(bool c) { contiguous_ = c; }
  const DistributedStates& get(size_t i) const {
    HB_CHECK(i < union_.size()) << "ds union index " << i << " out of range " << union_.size();
    return union_[i];
  }
  DistributedStates& get_mut(size_t i) { return union_[i]; }
  // DS of sub-group i seen in isolation (the hetero dim is not split inside a sub-group)
  DistributedStates get_local(size_t i) const { return get(i); }
  void add(const DistributedStates& ds) { union_.push_back(ds); }
  const std::vector<DistributedStates>& raw() const { return union_; }
  bool check_equal(const DistributedStatesUnion& o) const;
  // re-express the union along another hetero dim (`dim` must hold >= union size shards in every member)
  DistributedStatesUnion to_hetero(int dim, int num) const;
  std::string str() const;

 private:
  std::vector<DistributedStates> union_;
  int hetero_dim_;
  bool contiguous_;
};

class DistributedStatesHierarchy {
 public:
  DistributedStatesHierarchy() = default;
  explicit DistributedStatesHierarchy(std::vector<DistributedStatesUnion> h) : h_(std::move(h)) {}
  size_t size() const { return h_.size(); }
  const DistributedStatesUnion& get(size_t i) const {
    HB_CHECK(i < h_.size()) << "strategy id " << i << " out of range " << h_.size();
    return h_[i];
  }
  DistributedStatesUnion& get_mut(size_t i) { return h_[i]; }
  void add(const DistributedStatesUnion& u) { h_.push_back(u); }
  const std::vector<DistributedStatesUnion>& raw() const { return h_; }
  // convenience for the homogeneous case
  const DistributedStates& get_default_ds(size_t strategy = 0) const { return get(strategy).get(0); }

 private:
  std::vector<DistributedStatesUnion> h_;
};

// ------------------------------------------------------------------ comm classification
enum class CommType : int {
  UNUSED = 0, P2P, COMM_SPLIT, SCATTER, ALL_REDUCE, ALL_GATHER, REDUCE_SCATTER, BROADCAST, REDUCE,
  SPLIT_ALL_REDUCE, SPLIT_REDUCE_SCATTER, SPLIT_ALL_GATHER, BATCHED_ISEND_IRECV, ALL_TO_ALL
};
const char* comm_type_name(CommType t);

struct CommPlan {
  CommType type = CommType::UNUSED;
  int dim = kNullHeteroDim;       // gather / scatter tensor dim when applicable
  std::vector<int> group;         // indices (into the device group) of the peers of `device_index`
};

// Classify the transformation src layout/group -> dst layout/group for one (homogeneous) member.
CommType classify_comm(const DistributedStates& src, const DeviceGroup& src_group, const DistributedStates& dst,
                       const DeviceGroup& dst_group);
// Union-aware classification (heterogeneous strategies).
CommType classify_comm_union(const DistributedStatesUnion& src, const DeviceGroupUnion& src_groups,
                             const DistributedStatesUnion& dst, const DeviceGroupUnion& dst_groups);
// Full plan for one device of a homogeneous group (which peers, which dim).
CommPlan plan_comm(const DistributedStates& src, const DistributedStates& dst, const DeviceGroup& group,
                   int device_index);

// ------------------------------------------------------------------ re-sharding (hot switch / irregular comm)
// A rectangular slice of a global tensor owned by one device.
struct SliceSpec {
  std::vector<int64_t> begin, size;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : size) n *= s;
    return n;
  }
};
struct TransferItem {
  int src_device, dst_device;  // indices into the respective groups' global rank lists
  SliceSpec global;            // region of the global tensor moved by this item
};
enum class SwitchAlgorithm : int { FCFS = 0, ROUND_ROBIN, MULTI_NODE_ROUND_ROBIN, GREEDY, NEW_GREEDY };
SwitchAlgorithm switch_algorithm_from_env();
// Plan the peer-to-peer transfers that turn layout (src_ds on src_ranks) into (dst_ds on dst_ranks) for a
// tensor of `global_shape`.  Every destination slice is cut by the source tiling; when several sources hold
// a replica of a piece, `algo` picks the sender (load is balanced by bytes already assigned).
std::vector<TransferItem> plan_resharding(const std::vector<int64_t>& global_shape, const DistributedStates& src_ds,
                                          const std::vector<int>& src_ranks, const DistributedStates& dst_ds,
                                          const std::vector<int>& dst_ranks, SwitchAlgorithm algo,
                                          std::vector<int64_t>* send_load = nullptr, int devices_per_node = 8);

}  // namespace hb
