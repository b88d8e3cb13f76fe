// Graph IR: Tensor / Operator / Graph, op registry, autodiff and topo sort.
//
// Local-view SPMD semantics (as in the reference): every op computes on the
// rank-local shard; each tensor carries a DistributedStates hierarchy that says how
// the shard relates to the logical tensor under each parallel strategy, and `comm`
// ops declare a target layout that the executor lowers to collectives / P2P.
//
// Differences from the reference that matter on B200: op bodies are thin closures
// over at::Tensor (shape inference = the same closure on meta tensors), the hot
// ops bind to hand-written sm_100a kernels, and the executor runs a compiled
// per-rank plan instead of interpreting the define graph.
// (capability parity: hetu/graph/{graph,tensor,operator}.h, graph.cc:30-117)
#pragma once
#include <ATen/ATen.h>

#include <functional>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <unordered_map>
#include <variant>
#include <vector>

#include "../core/device.h"
#include "../core/ds.h"
#include "../core/symbol.h"

namespace hb {

class Graph;
class OpDef;
class TensorDef;
using Tensor = std::shared_ptr<TensorDef>;
using Operator = std::shared_ptr<OpDef>;
using TensorList = std::vector<Tensor>;
using TensorId = int64_t;
using OpId = int64_t;

at::ScalarType to_aten_dtype(DataType t);
DataType from_aten_dtype(at::ScalarType t);

// ------------------------------------------------------------------ attributes
using AttrValue = std::variant<bool, int64_t, double, std::string, std::vector<int64_t>, std::vector<double>>;
class AttrMap {
 public:
  bool has(const std::string& k) const { return m_.count(k) > 0; }
  void set(const std::string& k, AttrValue v) { m_[k] = std::move(v); }
  template <typename T>
  T get(const std::string& k) const {
    auto it = m_.find(k);
    HB_CHECK(it != m_.end()) << "missing attribute '" << k << "'";
    const T* p = std::get_if<T>(&it->second);
    HB_CHECK(p != nullptr) << "attribute '" << k << "' has another type";
    return *p;
  }
  template <typename T>
  T get_or(const std::string& k, T dflt) const {
    auto it = m_.find(k);
    if (it == m_.end()) return dflt;
    const T* p = std::get_if<T>(&it->second);
    return p ? *p : dflt;
  }
  int64_t i(const std::string& k, int64_t d = 0) const {
    auto it = m_.find(k);
    if (it == m_.end()) return d;
    if (auto* p = std::get_if<int64_t>(&it->second)) return *p;
    if (auto* p = std::get_if<bool>(&it->second)) return *p ? 1 : 0;
    if (auto* p = std::get_if<double>(&it->second)) return (int64_t)*p;
    return d;
  }
  double f(const std::string& k, double d = 0.0) const {
    auto it = m_.find(k);
    if (it == m_.end()) return d;
    if (auto* p = std::get_if<double>(&it->second)) return *p;
    if (auto* p = std::get_if<int64_t>(&it->second)) return (double)*p;
    if (auto* p = std::get_if<bool>(&it->second)) return *p ? 1.0 : 0.0;
    return d;
  }
  bool b(const std::string& k, bool d = false) const { return i(k, d ? 1 : 0) != 0; }
  std::string s(const std::string& k, const std::string& d = "") const { return get_or<std::string>(k, d); }
  std::vector<int64_t> ints(const std::string& k) const { return get_or<std::vector<int64_t>>(k, {}); }
  std::vector<double> floats(const std::string& k) const { return get_or<std::vector<double>>(k, {}); }
  const std::map<std::string, AttrValue>& raw() const { return m_; }

 private:
  std::map<std::string, AttrValue> m_;
};

// ------------------------------------------------------------------ op meta
struct OpMeta {
  std::string name;
  int stream_index = -1;                 // -1: role default (compute / collective / p2p ...)
  DeviceGroupHierarchy dg_hierarchy;     // placement per strategy (empty: inherit from inputs)
  std::vector<Tensor> extra_deps;        // control dependencies
  bool is_cpu = false;
  std::vector<bool> recompute;           // per strategy
  std::vector<bool> cpu_offload;         // per strategy
  std::string subgraph;                  // module path ("" = top level)
  int autocast_id = -1;
};

// ------------------------------------------------------------------ tensor
class TensorDef {
 public:
  TensorId id = -1;
  std::string name;
  OpDef* producer = nullptr;
  int output_index = 0;
  DataType dtype = DataType::FLOAT32;
  std::vector<int64_t> shape;     // rank-local shape under the graph's current strategy
  SyShape symbolic_shape;         // optional (empty when fully static)
  DistributedStatesHierarchy ds_hierarchy;
  bool requires_grad = false;
  bool is_grad = false;
  Graph* graph = nullptr;
  std::vector<OpDef*> consumers;
  at::Tensor eager_data;          // eager / define-by-run graphs only
  Tensor grad;                    // filled by backward() in eager graphs

  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
  int ndim() const { return (int)shape.size(); }
  bool has_ds(size_t strategy = 0) const { return ds_hierarchy.size() > strategy && ds_hierarchy.get(strategy).size() > 0; }
  const DistributedStates& ds(size_t strategy = 0) const { return ds_hierarchy.get(strategy).get(0); }
  std::vector<int64_t> global_shape(size_t strategy = 0) const {
    return has_ds(strategy) ? ds(strategy).global_shape(shape) : shape;
  }
};

// ------------------------------------------------------------------ op kernels
struct RunCtx;   // executor-side context (streams, micro-batch index, comm handles ...) -- see exec.h
class OpDef;

using ComputeFn = std::function<std::vector<at::Tensor>(const OpDef&, const std::vector<at::Tensor>&, RunCtx*)>;
// builds gradient ops: returns one (possibly null) grad tensor per input
using GradientFn = std::function<TensorList(OpDef&, const TensorList& grad_outputs)>;
// deduces output DS for one strategy; default copies input 0 (or stays empty)
using DeduceStatesFn = std::function<void(OpDef&, size_t strategy)>;
// explicit shape inference; default = run compute on meta tensors
using InferMetaFn = std::function<void(OpDef&)>;

enum OpFlag : uint32_t {
  kFlagPlaceholder = 1u << 0, kFlagVariable = 1u << 1, kFlagComm = 1u << 2, kFlagOptimizerUpdate = 1u << 3,
  kFlagInplace = 1u << 4, kFlagLoss = 1u << 5, kFlagGroup = 1u << 6, kFlagDataTransfer = 1u << 7,
  kFlagAttention = 1u << 8, kFlagNoMetaExec = 1u << 9, kFlagGradOf = 1u << 10, kFlagConst = 1u << 11,
  kFlagNondiff = 1u << 12,
};

struct OpKernel {
  std::string type;
  int num_outputs = 1;        // -1: variadic, decided by infer_meta
  uint32_t flags = 0;
  ComputeFn compute;
  GradientFn gradient;        // may be empty -> generic VJP by recompute (autograd)
  DeduceStatesFn deduce_states;
  InferMetaFn infer_meta;
};

class OpRegistry {
 public:
  static OpRegistry& get();
  void add(OpKernel k);
  const OpKernel* find(const std::string& type) const;
  std::vector<std::string> list() const;

 private:
  std::unordered_map<std::string, OpKernel> kernels_;
};
struct OpRegistrar {
  explicit OpRegistrar(OpKernel k) { OpRegistry::get().add(std::move(k)); }
};

// ------------------------------------------------------------------ operator
class OpDef {
 public:
  OpId id = -1;
  std::string type;
  TensorList inputs;
  TensorList outputs;
  AttrMap attrs;
  OpMeta meta;
  const OpKernel* kernel = nullptr;
  Graph* graph = nullptr;
  OpId fw_op_id = -1;       // forward op this gradient op belongs to (-1 for forward ops)
  bool is_bwd = false;
  bool pruned = false;      // define-by-run: removed by Graph::prune()
  // side payloads that do not fit AttrMap
  DistributedStatesHierarchy dst_ds;      // comm / placeholder / variable target layouts
  SyShape sy_shape;                        // symbolic target shape (reshape / slice ...)
  at::Tensor const_data;                   // constants / provided initial values

  const std::string& name() const { return meta.name; }
  bool has_flag(uint32_t f) const { return kernel && (kernel->flags & f); }
  Tensor input(size_t i) const { return inputs.at(i); }
  Tensor output(size_t i) const { return outputs.at(i); }
  // placement of this op for a strategy (union of all hetero members); empty = everywhere
  DeviceGroup placement(size_t strategy = 0) const {
    if (meta.dg_hierarchy.size() > strategy) return meta.dg_hierarchy.get(strategy).all();
    if (meta.dg_hierarchy.size() == 1) return meta.dg_hierarchy.get(0).all();
    return DeviceGroup();
  }
};

// ------------------------------------------------------------------ graph
enum class GraphKind : int { EAGER = 0, DEFINE_BY_RUN = 1, DEFINE_AND_RUN = 2, EXECUTABLE = 3 };
enum class RunLevel : int { UPDATE = 0, GRAD = 1, COMPUTE_ONLY = 2, ALLOC = 3, TOPO = 4 };

struct SubGraphInfo {
  std::string name;
  std::string module_type;  // MODULE / PIPELINE / OPTIMIZE_COMPUTE_BRIDGE / COMPUTE_OPTIMIZE_BRIDGE / TERMINATE
  std::string parent;
  std::vector<OpId> fwd_ops, bwd_ops, update_ops;
};

class Executor;  // exec.h

class Graph {
 public:
  Graph(GraphKind kind, const std::string& name, int num_strategy = 1);
  ~Graph();

  GraphKind kind() const { return kind_; }
  const std::string& name() const { return name_; }
  int num_strategy() const { return num_strategy_; }
  void set_num_strategy(int n) { num_strategy_ = n; }
  int cur_strategy() const { return cur_strategy_; }
  void set_cur_strategy(int s) { cur_strategy_ = s; }

  // op construction (runs meta inference + DS deduction; eager graphs also compute)
  TensorList make_op(const std::string& type, const TensorList& inputs, AttrMap attrs = {}, OpMeta meta = {},
                     std::function<void(OpDef&)> init = nullptr);
  Tensor make_op1(const std::string& type, const TensorList& inputs, AttrMap attrs = {}, OpMeta meta = {},
                  std::function<void(OpDef&)> init = nullptr) {
    return make_op(type, inputs, std::move(attrs), std::move(meta), std::move(init)).at(0);
  }

  const std::vector<Operator>& ops() const { return ops_; }
  Operator op(OpId id) const { return ops_.at(id); }
  size_t num_ops() const { return ops_.size(); }
  // reverse-mode autodiff: d(ys)/d(xs); grad_ys may be empty (ones)
  TensorList gradients(const TensorList& ys, const TensorList& xs, const TensorList& grad_ys = {});
  // ops needed to compute `fetches`, topologically sorted (bfs-depth + id tie-break)
  std::vector<OpDef*> topo_sort(const TensorList& fetches) const;

  // parameters (variable ops) and their state
  std::vector<Tensor> parameters() const;
  std::unordered_map<TensorId, at::Tensor>& param_data() { return param_data_; }
  bool has_param_data(TensorId id) const { return param_data_.count(id) > 0; }

  // module subgraphs
  void push_subgraph(const std::string& name, const std::string& module_type);
  void pop_subgraph();
  std::string cur_subgraph() const { return subgraph_stack_.empty() ? "" : subgraph_stack_.back(); }
  std::map<std::string, SubGraphInfo>& subgraphs() { return subgraphs_; }

  // contexts applied to every new op (python context managers push/pop these)
  struct Ctx {
    DeviceGroupHierarchy dg_hierarchy;
    int stream_index = -1;
    std::vector<Tensor> extra_deps;
    std::vector<bool> recompute, cpu_offload;
    int autocast_dtype = -1;   // DataType or -1
    bool building_backward = false;
  };
  Ctx& ctx() { return ctx_; }
  std::vector<Ctx>& ctx_stack() { return ctx_stack_; }

  Executor* executor();   // lazily created (define-and-run)
  int64_t next_tensor_id() { return next_tensor_id_++; }

  // re-run shape inference for every op under `strategy` (local shapes are strategy dependent)
  void reinfer_shapes(int strategy);

  // eager helpers
  void eager_backward(const Tensor& loss, const Tensor& grad = nullptr);

  // ---- define-by-run graphs (ref: hetu/graph/define_by_run_graph.{h,cc}): ops are recorded, not executed;
  // * an op identical to an existing one (same type, inputs, attributes; deterministic, not in place) is not created
  //   again -- its outputs are reused (FindReusableOp);
  // * a tensor's value is computed on demand by running exactly the not-yet-evaluated part of its ancestry, results
  //   are cached on the tensors (materialise);
  // * ops none of whose outputs is referenced any more (by user handles or by live consumers) are pruned: their cached
  //   values are dropped and they leave the consumer lists (prune)
  at::Tensor materialize(const Tensor& t);
  size_t prune();
  size_t num_live_ops() const;
  int64_t reuse_hits() const { return reuse_hits_; }

  static std::shared_ptr<Graph> make(GraphKind kind, const std::string& name, int num_strategy = 1);
  static std::shared_ptr<Graph> default_eager();

 private:
  void deduce_states(OpDef& op);
  void infer_meta(OpDef& op);
  GraphKind kind_;
  std::string name_;
  int num_strategy_;
  int cur_strategy_ = 0;
  std::vector<Operator> ops_;
  int64_t next_tensor_id_ = 0;
  std::unordered_map<TensorId, at::Tensor> param_data_;
  std::vector<std::string> subgraph_stack_;
  std::map<std::string, SubGraphInfo> subgraphs_;
  Ctx ctx_;
  std::vector<Ctx> ctx_stack_;
  std::unique_ptr<Executor> executor_;
  Operator find_reusable_op(const OpKernel* k, const std::string& type, const TensorList& inputs, const AttrMap& attrs) const;
  int64_t reuse_hits_ = 0;
};

// helpers used by op definitions -----------------------------------------------------------
// run `compute` on meta tensors to obtain output shapes / dtypes
void infer_meta_by_meta_exec(OpDef& op);
// default DS rule: outputs inherit input 0's layout
void deduce_states_like_input(OpDef& op, size_t strategy, size_t input_index = 0);
// generic gradient through ATen autograd (recompute forward under grad mode)
TensorList autograd_gradient(OpDef& op, const TensorList& grad_outputs);

#define HB_REGISTER_OP(NAME, ...) static ::hb::OpRegistrar _hb_op_reg_##NAME(::hb::OpKernel{__VA_ARGS__})

}  // namespace hb
