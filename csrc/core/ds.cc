#include "ds.h"

#include <algorithm>
#include <numeric>

namespace hb {

// ------------------------------------------------------------------ construction
DistributedStates::DistributedStates(int device_num, const std::map<int, int>& states, const std::vector<int>& order,
                                     bool zero)
    : device_num_(device_num), zero_(zero) {
  set_states(states);
  set_order(order);
}

void DistributedStates::set_states(const std::map<int, int>& states) {
  HB_CHECK(device_num_ != -1) << "device_num must be set before states";
  int prod = 1;
  states_.clear();
  for (auto& kv : states) {
    HB_CHECK(kv.first >= kPartialDim) << "invalid state dim " << kv.first;
    if (kv.second > 1) {
      states_[kv.first] = kv.second;
      prod *= kv.second;
    }
  }
  if (!states_.count(kPartialDim)) states_[kPartialDim] = 1;
  if (!states_.count(kDupDim)) states_[kDupDim] = 1;
  HB_CHECK(prod == device_num_) << "states use " << prod << " devices but the layout was declared for " << device_num_;
}

void DistributedStates::set_order(const std::vector<int>& order) {
  order_.clear();
  if (order.empty()) {
    for (auto& kv : states_) if (kv.second > 1) order_.push_back(kv.first);  // std::map iterates sorted
    return;
  }
  for (auto& kv : states_)
    if (kv.second > 1)
      HB_CHECK(std::find(order.begin(), order.end(), kv.first) != order.end())
          << "order " << order << " does not mention split dim " << kv.first;
  for (int o : order) {
    auto it = states_.find(o);
    if (it != states_.end() && it->second > 1) order_.push_back(o);
  }
}

bool DistributedStates::check_max_dim(int max_dim) const {
  for (auto& kv : states_) if (kv.first >= max_dim) return false;
  return true;
}

// ------------------------------------------------------------------ combine / reduce
std::map<int, int> DistributedStates::combine_states(const std::vector<int>& src, int dst,
                                                     const std::map<int, int>& ori) {
  std::map<int, int> st = ori;
  int value = 1;
  for (int s : src) {
    HB_CHECK(s != dst) << "cannot combine dim " << s << " into itself";
    if (s < 0) {
      value *= st[s];
      st[s] = 1;
    } else {
      auto it = st.find(s);
      if (it != st.end()) {
        value *= it->second;
        st.erase(it);
      }
      // tensor dims above the removed one shift down by one
      std::map<int, int> shifted;
      for (auto& kv : st) shifted[(kv.first > s) ? kv.first - 1 : kv.first] = kv.second;
      st.swap(shifted);
    }
  }
  if (dst < 0) {
    st[dst] = (st.count(dst) ? st[dst] : 1) * value;
  } else {
    int d = dst;
    for (int s : src) if (s >= 0 && dst > s) d -= 1;
    st[d] = (st.count(d) ? st[d] : 1) * value;
  }
  return st;
}

std::vector<int> DistributedStates::combine_order(const std::vector<int>& src, int dst, const std::vector<int>& ori) {
  std::vector<int> order = ori;
  std::vector<int> inds;
  auto collect = [&](int dim) {
    auto it = std::find(order.begin(), order.end(), dim);
    if (it != order.end()) inds.push_back(int(it - order.begin()));
  };
  for (int s : src) collect(s);
  collect(dst);
  std::sort(inds.begin(), inds.end());
  if (!inds.empty()) {
    for (size_t i = 1; i < inds.size(); ++i)
      HB_CHECK(inds[i] == inds[0] + int(i)) << "cannot combine state dims that are not adjacent in the device order";
    for (size_t i = inds.size(); i-- > 1;) order.erase(order.begin() + inds[i]);
    order[inds[0]] = dst;
    for (auto& o : order)
      if (o > 0)
        for (int s : src) if (s >= 0 && o > s) o -= 1;
  }
  return order;
}

bool DistributedStates::check_combine(const DistributedStates& dst, const std::vector<int>& src, int dst_dim) const {
  // normalise: entries equal to 1 for tensor dims never appear; -1/-2 always present
  auto st = combine_states(src, dst_dim);
  for (auto it = st.begin(); it != st.end();)
    if (it->first >= 0 && it->second <= 1) it = st.erase(it); else ++it;
  return st == dst.states_ && combine_order(src, dst_dim) == dst.order_;
}

std::map<int, int> DistributedStates::reduce_states(int dim) const {
  auto st = states_;
  if (dim < 0) st[dim] = 1; else st.erase(dim);
  return st;
}
std::vector<int> DistributedStates::reduce_order(int dim) const {
  auto o = order_;
  auto it = std::find(o.begin(), o.end(), dim);
  if (it != o.end()) o.erase(it);
  return o;
}
bool DistributedStates::check_reduce_dim(const DistributedStates& dst, int dim) const {
  return reduce_states(dim) == dst.states_ && reduce_order(dim) == dst.order_;
}

// ------------------------------------------------------------------ relations
bool DistributedStates::check_split(const DistributedStates& dst) const {
  int split_num = 1;
  for (int o : dst.order_) {
    if (o >= 0) {
      if (dst.get_dim(o) % get_dim(o) != 0) return false;
      split_num *= dst.get_dim(o) / get_dim(o);
    }
  }
  return dst.get_dim(kPartialDim) == get_dim(kPartialDim) && split_num > 1 && get_dim(kDupDim) == split_num;
}
bool DistributedStates::check_scatter(const DistributedStates& dst) const {
  const int d = dst.get_split_dim(*this);
  if (d < 0) return false;
  return get_dim(kDupDim) > 1 && check_combine(dst, {kDupDim}, d);
}
bool DistributedStates::check_allreduce(const DistributedStates& dst) const {
  return get_dim(kPartialDim) > 1 && check_combine(dst, {kPartialDim}, kDupDim);
}
bool DistributedStates::check_allgather(const DistributedStates& dst) const {
  const int d = get_split_dim(dst);
  if (d < 0) return false;
  return get_dim(d) > 1 && dst.get_dim(kDupDim) > 1 && dst.check_combine(*this, {kDupDim}, d);
}
bool DistributedStates::check_reducescatter(const DistributedStates& dst) const {
  const int d = dst.get_split_dim(*this);
  if (d < 0) return false;
  return get_dim(kPartialDim) > 1 && check_combine(dst, {kPartialDim}, d);
}
bool DistributedStates::check_broadcast(const DistributedStates& dst) const {
  return dst.get_dim(kDupDim) > 1 && dst.check_reduce_dim(*this, kDupDim);
}
bool DistributedStates::check_reduce(const DistributedStates& dst) const {
  return get_dim(kPartialDim) > 1 && check_reduce_dim(dst, kPartialDim);
}

int DistributedStates::get_split_dim(const DistributedStates& merged) const {
  int split_dim = kNullHeteroDim;
  for (auto& kv : states_) {
    if (kv.first < 0 || kv.second <= 1) continue;
    if (merged.get_dim(kv.first) < kv.second) {
      if (split_dim != kNullHeteroDim) return kNullHeteroDim;  // more than one dim differs: not a simple gather
      split_dim = kv.first;
    }
  }
  return split_dim;
}

std::vector<int> DistributedStates::get_loop_sizes() const {
  std::vector<int> sizes(order_.size(), 1);
  int acc = 1;
  for (size_t i = order_.size(); i-- > 0;) {
    sizes[i] = acc;
    acc *= get_dim(order_[i]);
  }
  return sizes;
}

std::map<int, int> DistributedStates::map_device_to_state_index(int device_index) const {
  std::map<int, int> idx;
  for (size_t i = order_.size(); i-- > 0;) {
    const int n = get_dim(order_[i]);
    idx[order_[i]] = device_index % n;
    device_index /= n;
  }
  return idx;
}

int DistributedStates::get_dup_group_index(int device_index) const {
  auto cur = map_device_to_state_index(device_index);
  std::vector<int> order = order_;
  std::sort(order.begin(), order.end());
  int idx = 0, interval = 1;
  for (size_t i = order.size(); i-- > 0;) {
    if (order[i] < 0) break;
    idx += cur[order[i]] * interval;
    interval *= get_dim(order[i]);
  }
  return idx;
}

std::vector<int> DistributedStates::get_device_indices_by_dim(int dim, int local_device_idx) const {
  auto it = std::find(order_.begin(), order_.end(), dim);
  if (it == order_.end()) return {local_device_idx};
  int interval = 1;
  for (auto c = it + 1; c != order_.end(); ++c) interval *= get_dim(*c);
  const int macro = interval * get_dim(dim);
  const int start = local_device_idx - local_device_idx % macro + local_device_idx % interval;
  std::vector<int> out;
  for (int i = start; i < start + macro; i += interval) out.push_back(i);
  return out;
}

DeviceGroup DistributedStates::get_devices_by_dim(int dim, int local_device_idx, const DeviceGroup& group) const {
  std::vector<Device> v;
  for (int i : get_device_indices_by_dim(dim, local_device_idx)) v.push_back(group.get(i));
  return DeviceGroup(v);
}

std::string DistributedStates::str() const {
  std::ostringstream os;
  os << "DS(n=" << device_num_ << ", states={";
  bool first = true;
  for (auto& kv : states_) {
    if (kv.second <= 1) continue;
    os << (first ? "" : ", ") << kv.first << ":" << kv.second;
    first = false;
  }
  os << "}, order=" << order_ << (zero_ ? ", zero" : "") << ")";
  return os.str();
}
std::ostream& operator<<(std::ostream& os, const DistributedStates& ds) { return os << ds.str(); }

std::vector<int64_t> DistributedStates::local_shape(const std::vector<int64_t>& global) const {
  std::vector<int64_t> out = global;
  for (auto& kv : states_) {
    if (kv.first < 0 || kv.second <= 1) continue;
    HB_CHECK(kv.first < (int)global.size()) << "split dim " << kv.first << " out of rank " << global.size();
    HB_CHECK(global[kv.first] % kv.second == 0)
        << "dim " << kv.first << " of size " << global[kv.first] << " is not divisible by " << kv.second;
    out[kv.first] = global[kv.first] / kv.second;
  }
  return out;
}
std::vector<int64_t> DistributedStates::global_shape(const std::vector<int64_t>& local) const {
  std::vector<int64_t> out = local;
  for (auto& kv : states_) {
    if (kv.first < 0 || kv.second <= 1) continue;
    HB_CHECK(kv.first < (int)local.size()) << "split dim " << kv.first << " out of rank " << local.size();
    out[kv.first] = local[kv.first] * kv.second;
  }
  return out;
}
void DistributedStates::local_slice(const std::vector<int64_t>& global, int device_index, std::vector<int64_t>* begin,
                                    std::vector<int64_t>* size) const {
  *size = local_shape(global);
  begin->assign(global.size(), 0);
  auto idx = map_device_to_state_index(device_index);
  for (auto& kv : idx)
    if (kv.first >= 0) (*begin)[kv.first] = int64_t(kv.second) * (*size)[kv.first];
}

// ------------------------------------------------------------------ union
bool DistributedStatesUnion::check_equal(const DistributedStatesUnion& o) const {
  if (union_.size() != o.union_.size() || hetero_dim_ != o.hetero_dim_) return false;
  for (size_t i = 0; i < union_.size(); ++i) if (!union_[i].check_equal(o.union_[i])) return false;
  return true;
}

DistributedStatesUnion DistributedStatesUnion::to_hetero(int dim, int num) const {
  // A homogeneous layout over N devices whose `dim` axis has >= num shards can be seen as `num` members,
  // each owning 1/num of that axis (the member layouts keep the remaining shards of the axis).
  HB_CHECK(union_.size() == 1) << "to_hetero expects a homogeneous union";
  const DistributedStates& ds = union_[0];
  HB_CHECK(ds.get_dim(dim) % num == 0) << "dim " << dim << " has " << ds.get_dim(dim) << " shards, cannot form " << num
                                       << " hetero members";
  auto st = ds.states();
  st[dim] = ds.get_dim(dim) / num;
  std::vector<DistributedStates> members;
  for (int i = 0; i < num; ++i) members.emplace_back(ds.device_num() / num, st, ds.order(), ds.zero());
  return DistributedStatesUnion(members, dim, true);
}

std::string DistributedStatesUnion::str() const {
  std::ostringstream os;
  os << "DSUnion([";
  for (size_t i = 0; i < union_.size(); ++i) os << (i ? ", " : "") << union_[i];
  os << "], hetero_dim=" << hetero_dim_ << ")";
  return os.str();
}

// ------------------------------------------------------------------ classification
const char* comm_type_name(CommType t) {
  static const char* n[] = {"UNUSED", "P2P", "COMM_SPLIT", "SCATTER", "ALL_REDUCE", "ALL_GATHER", "REDUCE_SCATTER",
                            "BROADCAST", "REDUCE", "SPLIT_ALL_REDUCE", "SPLIT_REDUCE_SCATTER", "SPLIT_ALL_GATHER",
                            "BATCHED_ISEND_IRECV", "ALL_TO_ALL"};
  return n[int(t)];
}

CommType classify_comm(const DistributedStates& src, const DeviceGroup& src_group, const DistributedStates& dst,
                       const DeviceGroup& dst_group) {
  if (src.check_equal(dst)) return src_group == dst_group ? CommType::UNUSED : CommType::P2P;
  if (src_group == dst_group) {
    // ordered rule table: first match wins (same precedence as the reference)
    struct Rule { bool (DistributedStates::*pred)(const DistributedStates&) const; CommType type; };
    static const Rule rules[] = {
        {&DistributedStates::check_scatter, CommType::SCATTER},
        {&DistributedStates::check_split, CommType::COMM_SPLIT},
        {&DistributedStates::check_allreduce, CommType::ALL_REDUCE},
        {&DistributedStates::check_allgather, CommType::ALL_GATHER},
        {&DistributedStates::check_reducescatter, CommType::REDUCE_SCATTER},
    };
    for (auto& r : rules) if ((src.*(r.pred))(dst)) return r.type;
    // same group, no partial sums involved: any other re-tiling is a (batched) point-to-point exchange
    if (src.get_dim(kPartialDim) == dst.get_dim(kPartialDim)) return CommType::BATCHED_ISEND_IRECV;
    HB_FAIL() << "no communication pattern turns " << src << " into " << dst;
  }
  HB_CHECK(src.get_dim(kPartialDim) == dst.get_dim(kPartialDim))
      << "cross-group communication with a pending reduction is not supported: " << src << " -> " << dst;
  return CommType::BATCHED_ISEND_IRECV;
}

CommType classify_comm_union(const DistributedStatesUnion& src, const DeviceGroupUnion& sg,
                             const DistributedStatesUnion& dst, const DeviceGroupUnion& dg) {
  bool no_reduction = true;
  for (auto& d : src.raw()) if (d.get_dim(kPartialDim) != 1) no_reduction = false;
  for (auto& d : dst.raw()) if (d.get_dim(kPartialDim) != 1) no_reduction = false;
  if (sg.size() != dg.size() || !src.contiguous()) {
    HB_CHECK(no_reduction) << "unions of different size (or non-contiguous input) cannot carry a reduction";
    if (src.check_equal(dst) && sg == dg) return CommType::UNUSED;
    bool split_ag = sg.size() == dg.size() && sg == dg && src.hetero_dim() == 0 && !src.contiguous() &&
                    dst.hetero_dim() == kDupDim && dst.contiguous();
    if (split_ag)
      for (size_t i = 0; i < sg.size(); ++i) if (!src.get_local(i).check_equal(dst.get_local(i))) split_ag = false;
    return split_ag ? CommType::SPLIT_ALL_GATHER : CommType::BATCHED_ISEND_IRECV;
  }
  if (src.hetero_dim() == dst.hetero_dim()) {
    // every member transforms independently; all members must agree on the pattern
    CommType t = classify_comm(src.get(0), sg.get(0), dst.get(0), dg.get(0));
    for (size_t i = 1; i < sg.size(); ++i) {
      CommType ti = classify_comm(src.get(i), sg.get(i), dst.get(i), dg.get(i));
      if (ti != t) return CommType::BATCHED_ISEND_IRECV;
    }
    return t;
  }
  // the hetero axis itself changes: collectives run across members on slices of the tensor
  if (no_reduction) {
    if (src.hetero_dim() >= 0 && dst.hetero_dim() == kDupDim && sg == dg) return CommType::SPLIT_ALL_GATHER;
    return CommType::BATCHED_ISEND_IRECV;
  }
  if (src.hetero_dim() == kPartialDim && dst.hetero_dim() == kDupDim) return CommType::SPLIT_ALL_REDUCE;
  if (src.hetero_dim() == kPartialDim && dst.hetero_dim() >= 0) return CommType::SPLIT_REDUCE_SCATTER;
  HB_FAIL() << "unsupported heterogeneous communication " << src.str() << " -> " << dst.str();
}

CommPlan plan_comm(const DistributedStates& src, const DistributedStates& dst, const DeviceGroup& group,
                   int device_index) {
  CommPlan p;
  p.type = classify_comm(src, group, dst, group);
  switch (p.type) {
    case CommType::ALL_REDUCE:
      p.group = src.get_device_indices_by_dim(kPartialDim, device_index);
      break;
    case CommType::ALL_GATHER:
      p.dim = src.get_split_dim(dst);
      p.group = dst.get_device_indices_by_dim(kDupDim, device_index);
      break;
    case CommType::REDUCE_SCATTER:
      p.dim = dst.get_split_dim(src);
      p.group = src.get_device_indices_by_dim(kPartialDim, device_index);
      break;
    case CommType::SCATTER:
      p.dim = dst.get_split_dim(src);
      p.group = src.get_device_indices_by_dim(kDupDim, device_index);
      break;
    default:
      break;
  }
  return p;
}

// ------------------------------------------------------------------ re-sharding planner
SwitchAlgorithm switch_algorithm_from_env() {
  std::string s = env_str("HETU_SWITCH_ALGORITHM", "NEW_GREEDY");
  for (auto& c : s) c = toupper(c);
  if (s == "FCFS") return SwitchAlgorithm::FCFS;
  if (s == "ROUND_ROBIN") return SwitchAlgorithm::ROUND_ROBIN;
  if (s == "MULTI_NODE_ROUND_ROBIN") return SwitchAlgorithm::MULTI_NODE_ROUND_ROBIN;
  if (s == "GREEDY") return SwitchAlgorithm::GREEDY;
  return SwitchAlgorithm::NEW_GREEDY;
}

namespace {
bool intersect(const SliceSpec& a, const SliceSpec& b, SliceSpec* out) {
  const size_t r = a.begin.size();
  out->begin.resize(r);
  out->size.resize(r);
  for (size_t i = 0; i < r; ++i) {
    const int64_t lo = std::max(a.begin[i], b.begin[i]);
    const int64_t hi = std::min(a.begin[i] + a.size[i], b.begin[i] + b.size[i]);
    if (hi <= lo) return false;
    out->begin[i] = lo;
    out->size[i] = hi - lo;
  }
  return true;
}
}  // namespace

std::vector<TransferItem> plan_resharding(const std::vector<int64_t>& global_shape, const DistributedStates& src_ds,
                                          const std::vector<int>& src_ranks, const DistributedStates& dst_ds,
                                          const std::vector<int>& dst_ranks, SwitchAlgorithm algo,
                                          std::vector<int64_t>* send_load_out, int devices_per_node) {
  HB_CHECK((int)src_ranks.size() == src_ds.device_num() && (int)dst_ranks.size() == dst_ds.device_num())
      << "rank lists must match the layouts";
  HB_CHECK(src_ds.get_dim(kPartialDim) == 1) << "cannot re-shard a tensor with a pending reduction";
  // distinct source tiles (one per dup-group) and the replicas holding each
  struct Tile { SliceSpec s; std::vector<int> holders; };
  std::map<int, Tile> tiles;
  for (int i = 0; i < src_ds.device_num(); ++i) {
    const int g = src_ds.get_dup_group_index(i);
    auto& t = tiles[g];
    if (t.holders.empty()) src_ds.local_slice(global_shape, i, &t.s.begin, &t.s.size);
    t.holders.push_back(i);
  }
  int max_rank = 0;
  for (int r : src_ranks) max_rank = std::max(max_rank, r);
  for (int r : dst_ranks) max_rank = std::max(max_rank, r);
  std::vector<int64_t> load(max_rank + 1, 0);
  std::vector<TransferItem> plan;
  int rr = 0;
  for (int j = 0; j < dst_ds.device_num(); ++j) {
    SliceSpec want;
    dst_ds.local_slice(global_shape, j, &want.begin, &want.size);
    const int dst_rank = dst_ranks[j];
    for (auto& kv : tiles) {
      SliceSpec piece;
      if (!intersect(kv.second.s, want, &piece)) continue;
      const auto& holders = kv.second.holders;
      int chosen = -1;
      // a replica already living on the destination rank costs nothing
      for (int hidx : holders) if (src_ranks[hidx] == dst_rank) chosen = hidx;
      if (chosen < 0) {
        switch (algo) {
          case SwitchAlgorithm::FCFS:
            chosen = holders[0];
            break;
          case SwitchAlgorithm::ROUND_ROBIN:
            chosen = holders[(rr++) % holders.size()];
            break;
          case SwitchAlgorithm::MULTI_NODE_ROUND_ROBIN: {
            // prefer a replica on the destination's node, round-robin inside that set
            std::vector<int> same;
            for (int hidx : holders) if (src_ranks[hidx] / devices_per_node == dst_rank / devices_per_node) same.push_back(hidx);
            const auto& pool = same.empty() ? holders : same;
            chosen = pool[(rr++) % pool.size()];
            break;
          }
          case SwitchAlgorithm::GREEDY:
          case SwitchAlgorithm::NEW_GREEDY: {
            int64_t best = -1;
            for (int hidx : holders) {
              int64_t cost = load[src_ranks[hidx]];
              // NEW_GREEDY additionally penalises cross-node senders
              if (algo == SwitchAlgorithm::NEW_GREEDY && src_ranks[hidx] / devices_per_node != dst_rank / devices_per_node)
                cost += piece.numel();
              if (best < 0 || cost < best) { best = cost; chosen = hidx; }
            }
            break;
          }
        }
      }
      TransferItem it;
      it.src_device = src_ranks[chosen];
      it.dst_device = dst_rank;
      it.global = piece;
      if (it.src_device != it.dst_device) load[it.src_device] += piece.numel();
      plan.push_back(it);
    }
  }
  if (send_load_out) *send_load_out = load;
  return plan;
}

}  // namespace hb
