#include <torch/csrc/autograd/autograd.h>
#include <torch/csrc/autograd/grad_mode.h>

#include <algorithm>
#include <deque>
#include <queue>

#include "exec.h"
#include "ir.h"

namespace hb {

// ------------------------------------------------------------------ dtype bridge
at::ScalarType to_aten_dtype(DataType t) {
  switch (t) {
    case DataType::UINT8: return at::kByte;
    case DataType::INT8: return at::kChar;
    case DataType::INT16: return at::kShort;
    case DataType::INT32: return at::kInt;
    case DataType::INT64: return at::kLong;
    case DataType::FLOAT16: return at::kHalf;
    case DataType::FLOAT32: return at::kFloat;
    case DataType::FLOAT64: return at::kDouble;
    case DataType::BFLOAT16: return at::kBFloat16;
    case DataType::BOOL: return at::kBool;
    case DataType::FLOAT8_E4M3: return at::kFloat8_e4m3fn;
    case DataType::FLOAT8_E5M2: return at::kFloat8_e5m2;
    case DataType::FLOAT4: case DataType::NFLOAT4: return at::kByte;  // packed storage
    default: HB_FAIL() << "dtype has no ATen equivalent";
  }
}
DataType from_aten_dtype(at::ScalarType t) {
  switch (t) {
    case at::kByte: return DataType::UINT8;
    case at::kChar: return DataType::INT8;
    case at::kShort: return DataType::INT16;
    case at::kInt: return DataType::INT32;
    case at::kLong: return DataType::INT64;
    case at::kHalf: return DataType::FLOAT16;
    case at::kFloat: return DataType::FLOAT32;
    case at::kDouble: return DataType::FLOAT64;
    case at::kBFloat16: return DataType::BFLOAT16;
    case at::kBool: return DataType::BOOL;
    case at::kFloat8_e4m3fn: return DataType::FLOAT8_E4M3;
    case at::kFloat8_e5m2: return DataType::FLOAT8_E5M2;
    default: HB_FAIL() << "unsupported ATen dtype " << int(t);
  }
}

// ------------------------------------------------------------------ registry
OpRegistry& OpRegistry::get() {
  static OpRegistry r;
  return r;
}
void OpRegistry::add(OpKernel k) {
  const std::string t = k.type;
  kernels_[t] = std::move(k);
}
const OpKernel* OpRegistry::find(const std::string& type) const {
  auto it = kernels_.find(type);
  return it == kernels_.end() ? nullptr : &it->second;
}
std::vector<std::string> OpRegistry::list() const {
  std::vector<std::string> v;
  for (auto& kv : kernels_) v.push_back(kv.first);
  std::sort(v.begin(), v.end());
  return v;
}

// ------------------------------------------------------------------ graph
Graph::Graph(GraphKind kind, const std::string& name, int num_strategy)
    : kind_(kind), name_(name), num_strategy_(num_strategy < 1 ? 1 : num_strategy) {}
Graph::~Graph() = default;

std::shared_ptr<Graph> Graph::make(GraphKind kind, const std::string& name, int num_strategy) {
  return std::make_shared<Graph>(kind, name, num_strategy);
}
std::shared_ptr<Graph> Graph::default_eager() {
  static std::shared_ptr<Graph> g = std::make_shared<Graph>(GraphKind::EAGER, "default_eager", 1);
  return g;
}

Executor* Graph::executor() {
  if (!executor_) executor_ = std::make_unique<Executor>(this);
  return executor_.get();
}

void Graph::push_subgraph(const std::string& name, const std::string& module_type) {
  const std::string parent = cur_subgraph();
  const std::string full = parent.empty() ? name : parent + "." + name;
  if (!subgraphs_.count(full)) {
    SubGraphInfo info;
    info.name = full;
    info.module_type = module_type;
    info.parent = parent;
    subgraphs_[full] = info;
  }
  subgraph_stack_.push_back(full);
}
void Graph::pop_subgraph() {
  HB_CHECK(!subgraph_stack_.empty()) << "subgraph stack underflow";
  subgraph_stack_.pop_back();
}

void infer_meta_by_meta_exec(OpDef& op) {
  std::vector<at::Tensor> ins;
  for (auto& t : op.inputs)
    ins.push_back(at::empty(t->shape, at::TensorOptions().dtype(to_aten_dtype(t->dtype)).device(at::kMeta)));
  HB_CHECK(op.kernel->compute) << "op " << op.type << " has no compute function";
  std::vector<at::Tensor> outs = op.kernel->compute(op, ins, nullptr);
  HB_CHECK(op.kernel->num_outputs < 0 || (int)outs.size() == op.kernel->num_outputs)
      << "op " << op.type << " produced " << outs.size() << " outputs, expected " << op.kernel->num_outputs;
  op.outputs.resize(outs.size());
  for (size_t i = 0; i < outs.size(); ++i) {
    if (!op.outputs[i]) op.outputs[i] = std::make_shared<TensorDef>();
    op.outputs[i]->shape = outs[i].sizes().vec();
    op.outputs[i]->dtype = from_aten_dtype(outs[i].scalar_type());
  }
}

void deduce_states_like_input(OpDef& op, size_t strategy, size_t input_index) {
  if (op.inputs.size() <= input_index) return;
  const Tensor& in = op.inputs[input_index];
  if (!in->has_ds(strategy)) return;
  for (auto& out : op.outputs) {
    while (out->ds_hierarchy.size() <= strategy) out->ds_hierarchy.add(DistributedStatesUnion());
    out->ds_hierarchy.get_mut(strategy) = in->ds_hierarchy.get(strategy);
  }
}

void Graph::infer_meta(OpDef& op) {
  if (op.kernel->infer_meta) op.kernel->infer_meta(op);
  else infer_meta_by_meta_exec(op);
}
void Graph::deduce_states(OpDef& op) {
  for (int s = 0; s < num_strategy_; ++s) {
    if (op.kernel->deduce_states) op.kernel->deduce_states(op, s);
    else deduce_states_like_input(op, s, 0);
  }
}

TensorList Graph::make_op(const std::string& type, const TensorList& inputs, AttrMap attrs, OpMeta meta,
                          std::function<void(OpDef&)> init) {
  const OpKernel* k = OpRegistry::get().find(type);
  HB_CHECK(k != nullptr) << "unknown op type '" << type << "'";
  auto op = std::make_shared<OpDef>();
  op->id = (OpId)ops_.size();
  op->type = type;
  op->inputs = inputs;
  if (kind_ == GraphKind::DEFINE_BY_RUN && !init && !ctx_.building_backward) {
    if (Operator hit = find_reusable_op(k, type, inputs, attrs)) {
      ++reuse_hits_;
      return hit->outputs;
    }
  }
  op->attrs = std::move(attrs);
  op->meta = std::move(meta);
  op->kernel = k;
  op->graph = this;
  op->is_bwd = ctx_.building_backward;
  for (auto& t : inputs) HB_CHECK(t != nullptr) << "null input to op " << type;
  // inherit context
  if (op->meta.dg_hierarchy.size() == 0) {
    if (ctx_.dg_hierarchy.size() > 0) op->meta.dg_hierarchy = ctx_.dg_hierarchy;
    else {
      for (auto& t : inputs)
        if (t->producer && t->producer->meta.dg_hierarchy.size() > 0) {
          op->meta.dg_hierarchy = t->producer->meta.dg_hierarchy;
          break;
        }
    }
  }
  if (op->meta.stream_index < 0) op->meta.stream_index = ctx_.stream_index;
  for (auto& d : ctx_.extra_deps) op->meta.extra_deps.push_back(d);
  if (op->meta.recompute.empty()) op->meta.recompute = ctx_.recompute;
  if (op->meta.cpu_offload.empty()) op->meta.cpu_offload = ctx_.cpu_offload;
  if (op->meta.subgraph.empty()) op->meta.subgraph = cur_subgraph();
  if (op->meta.name.empty()) op->meta.name = type + "_" + std::to_string(op->id);
  if (init) init(*op);

  infer_meta(*op);
  bool any_grad = false;
  for (auto& t : inputs) any_grad |= t->requires_grad;
  for (size_t i = 0; i < op->outputs.size(); ++i) {
    auto& o = op->outputs[i];
    o->id = next_tensor_id();
    o->producer = op.get();
    o->output_index = (int)i;
    o->graph = this;
    if (o->name.empty()) o->name = op->meta.name + (op->outputs.size() > 1 ? ":" + std::to_string(i) : "");
    if (!(k->flags & (kFlagVariable | kFlagPlaceholder | kFlagConst)))
      o->requires_grad = any_grad && dtype_is_float(o->dtype) && !(k->flags & kFlagNondiff);
    o->is_grad = op->is_bwd;
  }
  deduce_states(*op);
  for (auto& t : inputs) t->consumers.push_back(op.get());
  ops_.push_back(op);
  if (!op->meta.subgraph.empty()) {
    auto& sg = subgraphs_[op->meta.subgraph];
    if (sg.name.empty()) sg.name = op->meta.subgraph;
    if (k->flags & kFlagOptimizerUpdate) sg.update_ops.push_back(op->id);
    else if (op->is_bwd) sg.bwd_ops.push_back(op->id);
    else sg.fwd_ops.push_back(op->id);
  }

  if (kind_ == GraphKind::EAGER) {
    std::vector<at::Tensor> ins;
    for (auto& t : inputs) {
      HB_CHECK(t->eager_data.defined()) << "eager op " << type << " got input " << t->name << " without data";
      ins.push_back(t->eager_data);
    }
    RunCtx rc;
    rc.graph = this;
    at::NoGradGuard ng;
    auto outs = k->compute(*op, ins, &rc);
    HB_CHECK(outs.size() == op->outputs.size()) << "eager op " << type << " output count mismatch";
    for (size_t i = 0; i < outs.size(); ++i) {
      op->outputs[i]->eager_data = outs[i];
      op->outputs[i]->shape = outs[i].sizes().vec();
    }
  }
  return op->outputs;
}

// ---------------------------------------------------------------------------------------- define-by-run
Operator Graph::find_reusable_op(const OpKernel* k, const std::string& type, const TensorList& inputs, const AttrMap& attrs) const {
  // only pure ops: no variables / placeholders / constants (identity matters), nothing in place, nothing that draws random
  // numbers or communicates
  if (inputs.empty() || (k->flags & (kFlagVariable | kFlagPlaceholder | kFlagConst | kFlagInplace | kFlagOptimizerUpdate | kFlagComm)))
    return nullptr;
  if (type == "dropout" || type == "dropout_add_norm" || type.find("rand") != std::string::npos) return nullptr;
  for (OpDef* c : inputs[0]->consumers) {
    if (c->pruned || c->type != type || c->inputs.size() != inputs.size() || c->is_bwd) continue;
    bool same = true;
    for (size_t i = 0; i < inputs.size() && same; ++i) same = c->inputs[i] == inputs[i];
    if (!same || !(c->attrs.raw() == attrs.raw())) continue;
    return ops_.at(c->id);
  }
  return nullptr;
}

at::Tensor Graph::materialize(const Tensor& t) {
  HB_CHECK(t != nullptr) << "materialize(null)";
  if (t->eager_data.defined()) return t->eager_data;
  if (has_param_data(t->id)) return param_data_[t->id];
  HB_CHECK(t->producer != nullptr) << "tensor " << t->name << " has no producer and no data";
  RunCtx rc;
  rc.graph = this;
  at::NoGradGuard ng;
  for (OpDef* op : topo_sort({t})) {
    bool done = !op->outputs.empty();
    for (auto& o : op->outputs) done = done && (o->eager_data.defined() || has_param_data(o->id));
    if (done) continue;
    HB_CHECK(!op->pruned) << "op " << op->name() << " was pruned but is needed again";
    std::vector<at::Tensor> ins;
    for (auto& in : op->inputs) {
      if (in->eager_data.defined()) ins.push_back(in->eager_data);
      else if (has_param_data(in->id)) ins.push_back(param_data_[in->id]);
      else HB_CHECK(false) << "input " << in->name << " of " << op->name() << " has no value (unfed placeholder in a define-by-run graph?)";
    }
    if (op->has_flag(kFlagVariable)) {
      at::Tensor v = executor()->get_param(op->outputs[0]);
      op->outputs[0]->eager_data = v;
      continue;
    }
    auto outs = op->kernel->compute(*op, ins, &rc);
    HB_CHECK(outs.size() == op->outputs.size()) << "op " << op->name() << " output count mismatch";
    for (size_t i = 0; i < outs.size(); ++i) {
      op->outputs[i]->eager_data = outs[i];
      if (outs[i].defined()) op->outputs[i]->shape = outs[i].sizes().vec();
    }
  }
  return t->eager_data;
}

size_t Graph::prune() {
  // reverse creation order: consumers are visited before their producers, so whole dead chains disappear in one call
  size_t removed = 0;
  for (auto it = ops_.rbegin(); it != ops_.rend(); ++it) {
    OpDef* op = it->get();
    if (op->pruned || op->has_flag(kFlagVariable) || op->has_flag(kFlagPlaceholder)) continue;
    bool dead = true;
    for (auto& o : op->outputs) {
      // referenced from outside (python handle, another container)?  the op itself holds exactly one reference
      if (o.use_count() > 1) { dead = false; break; }
      for (OpDef* c : o->consumers) if (!c->pruned) { dead = false; break; }
      if (!dead) break;
    }
    if (!dead) continue;
    for (auto& in : op->inputs) {
      auto& cs = in->consumers;
      cs.erase(std::remove(cs.begin(), cs.end(), op), cs.end());
    }
    for (auto& o : op->outputs) o->eager_data = at::Tensor();
    op->inputs.clear();
    op->pruned = true;
    ++removed;
  }
  return removed;
}

size_t Graph::num_live_ops() const {
  size_t n = 0;
  for (auto& op : ops_) n += op->pruned ? 0 : 1;
  return n;
}

void Graph::reinfer_shapes(int strategy) {
  const int saved = cur_strategy_;
  cur_strategy_ = strategy;
  for (auto& op : ops_) {
    std::vector<Tensor> outs = op->outputs;  // keep tensor identities, refresh shapes in place
    try {
      infer_meta(*op);
    } catch (const std::exception& e) {
      std::ostringstream os;
      os << "shape re-inference of op " << op->name() << " (" << op->type << ") under strategy " << strategy << " with input shapes";
      for (auto& t : op->inputs) os << " " << t->name << t->shape;
      os << ": " << e.what();
      throw Error(os.str());
    }
    HB_CHECK(op->outputs.size() == outs.size()) << "shape re-inference changed the arity of " << op->name();
    for (size_t i = 0; i < outs.size(); ++i) {
      if (op->outputs[i] != outs[i]) {
        outs[i]->shape = op->outputs[i]->shape;
        outs[i]->dtype = op->outputs[i]->dtype;
        op->outputs[i] = outs[i];
      }
    }
  }
  cur_strategy_ = saved;
}

std::vector<Tensor> Graph::parameters() const {
  std::vector<Tensor> v;
  for (auto& op : ops_)
    if (op->has_flag(kFlagVariable) && op->outputs[0]->requires_grad) v.push_back(op->outputs[0]);
  return v;
}

// ------------------------------------------------------------------ topo sort
std::vector<OpDef*> Graph::topo_sort(const TensorList& fetches) const {
  // collect the ancestor set
  std::set<OpId> needed;
  std::vector<OpDef*> stack;
  for (auto& t : fetches) if (t && t->producer) stack.push_back(t->producer);
  while (!stack.empty()) {
    OpDef* op = stack.back();
    stack.pop_back();
    if (!needed.insert(op->id).second) continue;
    for (auto& in : op->inputs) if (in->producer) stack.push_back(in->producer);
    for (auto& d : op->meta.extra_deps) if (d && d->producer) stack.push_back(d->producer);
  }
  // ops are created in a valid topological order (inputs precede consumers), so ascending id is a
  // topological order that also preserves program order for in-place ops; this is deterministic across ranks
  std::vector<OpDef*> order;
  order.reserve(needed.size());
  for (OpId id : needed) order.push_back(ops_[id].get());
  return order;
}

// ------------------------------------------------------------------ autodiff
// tensors the caller of gradients() asked about by name: a frozen variable among them still gets its gradient
static thread_local std::set<TensorId> tl_requested_xs;

TensorList Graph::gradients(const TensorList& ys, const TensorList& xs, const TensorList& grad_ys) {
  HB_CHECK(grad_ys.empty() || grad_ys.size() == ys.size()) << "grad_ys must match ys";
  struct Requested {
    std::set<TensorId> saved;
    explicit Requested(const TensorList& xs) : saved(tl_requested_xs) { for (auto& x : xs) tl_requested_xs.insert(x->id); }
    ~Requested() { tl_requested_xs = saved; }
  } requested(xs);
  auto order = topo_sort(ys);
  // which ops lie on a path from xs to ys
  std::set<TensorId> from_x;
  for (auto& x : xs) from_x.insert(x->id);
  std::set<OpId> on_path;
  for (OpDef* op : order) {
    bool dep = false;
    for (auto& in : op->inputs) if (from_x.count(in->id)) dep = true;
    if (dep) {
      on_path.insert(op->id);
      for (auto& o : op->outputs) from_x.insert(o->id);
    }
  }
  std::unordered_map<TensorId, TensorList> pending;
  const bool prev_bwd = ctx_.building_backward;
  ctx_.building_backward = true;
  for (size_t i = 0; i < ys.size(); ++i) {
    Tensor g = grad_ys.empty() ? nullptr : grad_ys[i];
    if (!g) {
      OpMeta m;
      if (ys[i]->producer) m.dg_hierarchy = ys[i]->producer->meta.dg_hierarchy;
      g = make_op1("ones_like", {ys[i]}, {}, m);
    }
    pending[ys[i]->id].push_back(g);
  }
  auto sum_grads = [&](const Tensor& of, TensorList& gs) -> Tensor {
    if (gs.empty()) return nullptr;
    if (gs.size() == 1) return gs[0];
    OpMeta m;
    if (of->producer) m.dg_hierarchy = of->producer->meta.dg_hierarchy;
    if (gs.size() == 2) {
      // residual stream: d(x) = d(skip) + dgrad(branch).  Fold the addition into the dgrad GEMM epilogue (its residual
      // operand) instead of a separate elementwise pass; the plain dgrad op becomes dead and is never scheduled.
      for (int k = 0; k < 2; ++k) {
        const Tensor& a = gs[k];
        const Tensor& other = gs[1 - k];
        OpDef* p = a->producer;
        if (p != nullptr && p->type == "linear_dgrad" && p->inputs.size() == 2 && a->consumers.empty() && a->shape == other->shape &&
            a->dtype == other->dtype && a->ds_hierarchy.size() == other->ds_hierarchy.size()) {
          bool same = true;
          for (size_t s = 0; s < a->ds_hierarchy.size() && same; ++s)
            if (a->has_ds((int)s) != other->has_ds((int)s) || (a->has_ds((int)s) && !a->ds((int)s).check_equal(other->ds((int)s)))) same = false;
          if (!same) continue;
          AttrMap at = p->attrs;
          at.set("residual_add", true);
          return make_op1("linear_dgrad", {p->inputs[0], p->inputs[1], other}, at, p->meta);
        }
        // the same for a norm: d(x) = norm_bwd(...).dx + d(skip) -- the skip gradient becomes the norm backward's `dx_add`
        // operand.  The replaced op's dgamma / dbeta outputs are re-pointed in the pending lists; the old op is dead.
        if (p != nullptr && p->type == "norm_bwd" && p->inputs.size() == 5 && a == p->outputs[0] && a->consumers.empty() &&
            a->shape == other->shape && a->dtype == other->dtype && a->ds_hierarchy.size() == other->ds_hierarchy.size() &&
            env_int("HETU_FUSE_NORM_BWD_ADD", 1) != 0) {
          bool same = true;
          for (size_t s = 0; s < a->ds_hierarchy.size() && same; ++s)
            if (a->has_ds((int)s) != other->has_ds((int)s) || (a->has_ds((int)s) && !a->ds((int)s).check_equal(other->ds((int)s)))) same = false;
          bool params_unused = true;
          for (size_t o = 1; o < p->outputs.size(); ++o) params_unused = params_unused && p->outputs[o]->consumers.empty();
          if (!same || !params_unused) continue;
          TensorList ins = p->inputs;
          ins.push_back(other);
          TensorList outs = make_op("norm_bwd", ins, p->attrs, p->meta);
          for (auto& kv : pending)
            for (auto& t : kv.second)
              for (size_t o = 1; o < p->outputs.size() && o < outs.size(); ++o)
                if (t == p->outputs[o]) t = outs[o];
          return outs[0];
        }
      }
    }
    return make_op1("sum_n", gs, {}, m);
  };
  for (auto it = order.rbegin(); it != order.rend(); ++it) {
    OpDef* op = *it;
    if (!on_path.count(op->id)) continue;
    TensorList gouts(op->outputs.size());
    bool any = false;
    for (size_t i = 0; i < op->outputs.size(); ++i) {
      auto pit = pending.find(op->outputs[i]->id);
      if (pit != pending.end()) {
        gouts[i] = sum_grads(op->outputs[i], pit->second);
        any |= (gouts[i] != nullptr);
      }
    }
    if (!any) continue;
    // gradient ops inherit placement / recompute context of the forward op
    Ctx saved = ctx_;
    ctx_.dg_hierarchy = op->meta.dg_hierarchy;
    ctx_.recompute.clear();
    ctx_.cpu_offload.clear();
    const size_t first_new = ops_.size();
    TensorList gins;
    if (op->kernel->gradient) gins = op->kernel->gradient(*op, gouts);
    else if (!(op->kernel->flags & (kFlagNondiff | kFlagPlaceholder | kFlagVariable | kFlagConst)))
      gins = autograd_gradient(*op, gouts);
    ctx_ = saved;
    ctx_.building_backward = true;
    for (size_t j = first_new; j < ops_.size(); ++j) {
      ops_[j]->fw_op_id = op->id;
      if (!op->meta.subgraph.empty() && ops_[j]->meta.subgraph.empty()) ops_[j]->meta.subgraph = op->meta.subgraph;
    }
    for (size_t i = 0; i < gins.size() && i < op->inputs.size(); ++i) {
      if (!gins[i]) continue;
      if (!op->inputs[i]->requires_grad && !from_x.count(op->inputs[i]->id)) continue;
      // A gradient must come back in the layout of the tensor it belongs to ("partial" read as "duplicate"):
      // e.g. dX of a column-parallel linear is a partial sum over the TP group and is all-reduced right here.
      // Parameter gradients are exempt -- their (data-parallel) reduction is deferred and bucketed by the optimizer.
      {
        const Tensor& x = op->inputs[i];
        bool leaf = x->producer && (x->producer->has_flag(kFlagVariable) || x->producer->has_flag(kFlagPlaceholder));
        // a parameter shipped to another pipeline stage (tied embedding / lm_head: "share_weight_comm") is still that
        // parameter: its gradient travels back un-reduced and joins the owner's deferred data-parallel reduction once
        if (x->producer && x->producer->type == "comm" && !x->producer->inputs.empty() && x->producer->inputs[0]->producer &&
            x->producer->inputs[0]->producer->has_flag(kFlagVariable))
          leaf = true;
        if (!leaf && x->ds_hierarchy.size() > 0 && gins[i]->ds_hierarchy.size() > 0 && op->type != "comm") {
          DistributedStatesHierarchy target;
          bool differs = false;
          for (size_t s = 0; s < x->ds_hierarchy.size(); ++s) {
            DistributedStatesUnion u;
            const auto& xu = x->ds_hierarchy.get(s);
            for (size_t m = 0; m < xu.size(); ++m) {
              const DistributedStates& ds = xu.get(m);
              if (ds.get_dim(kPartialDim) > 1)
                u.add(DistributedStates(ds.device_num(), ds.combine_states({kPartialDim}, kDupDim), ds.combine_order({kPartialDim}, kDupDim)));
              else u.add(ds);
            }
            u.set_hetero_dim(xu.hetero_dim() == kPartialDim ? kDupDim : xu.hetero_dim());
            target.add(u);
            if (gins[i]->has_ds(s) && u.size() > 0 && !gins[i]->ds(s).check_equal(u.get(0))) differs = true;
          }
          if (differs) {
            OpMeta m;
            if (x->producer) m.dg_hierarchy = x->producer->meta.dg_hierarchy;
            const size_t before = ops_.size();
            gins[i] = make_op1("comm", {gins[i]}, {}, m, [&](OpDef& o) { o.dst_ds = target; });
            for (size_t j = before; j < ops_.size(); ++j) ops_[j]->fw_op_id = op->id;
          }
        }
      }
      pending[op->inputs[i]->id].push_back(gins[i]);
    }
  }
  TensorList out;
  for (auto& x : xs) {
    auto pit = pending.find(x->id);
    out.push_back(pit == pending.end() ? nullptr : sum_grads(x, pit->second));
  }
  ctx_.building_backward = prev_bwd;
  return out;
}

void Graph::eager_backward(const Tensor& loss, const Tensor& grad) {
  HB_CHECK(kind_ == GraphKind::EAGER) << "backward() is only available on eager graphs";
  TensorList xs;
  for (auto& op : ops_)
    if ((op->has_flag(kFlagVariable) || op->has_flag(kFlagConst)) && op->outputs[0]->requires_grad)
      xs.push_back(op->outputs[0]);
  TensorList gy;
  if (grad) gy.push_back(grad);
  TensorList gs = gradients({loss}, xs, gy);
  for (size_t i = 0; i < xs.size(); ++i) {
    if (!gs[i]) continue;
    if (xs[i]->grad && xs[i]->grad->eager_data.defined() && gs[i]->eager_data.defined()) {
      xs[i]->grad->eager_data = xs[i]->grad->eager_data + gs[i]->eager_data;
    } else {
      xs[i]->grad = gs[i];
    }
  }
}

// ------------------------------------------------------------------ generic VJP through ATen autograd
// op "autograd_vjp": inputs = [fw inputs..., grad_outputs (non-null ones)...]; attrs carry the forward op id.
TensorList autograd_gradient(OpDef& fw, const TensorList& gouts) {
  Graph* g = fw.graph;
  TensorList ins = fw.inputs;
  std::vector<int64_t> gout_idx;
  for (size_t i = 0; i < gouts.size(); ++i)
    if (gouts[i]) {
      ins.push_back(gouts[i]);
      gout_idx.push_back((int64_t)i);
    }
  std::vector<int64_t> diff_inputs;
  for (size_t i = 0; i < fw.inputs.size(); ++i) {
    const Tensor& t = fw.inputs[i];
    if (!dtype_is_float(t->dtype)) continue;
    // a variable that was declared non-trainable (running statistics, frozen tables, masks) never needs a gradient -- and some
    // library ops refuse to be traced with respect to it (native_batch_norm and its running mean / variance)
    if (t->producer != nullptr && t->producer->type == "variable" && !t->requires_grad && !tl_requested_xs.count(t->id)) continue;
    diff_inputs.push_back((int64_t)i);
  }
  if (diff_inputs.empty()) return TensorList(fw.inputs.size());
  AttrMap a;
  a.set("fw_op", (int64_t)fw.id);
  a.set("gout_idx", gout_idx);
  a.set("diff_inputs", diff_inputs);
  a.set("num_fw_inputs", (int64_t)fw.inputs.size());
  TensorList outs = g->make_op("autograd_vjp", ins, a);
  TensorList res(fw.inputs.size());
  for (size_t k = 0; k < diff_inputs.size(); ++k) res[diff_inputs[k]] = outs[k];
  return res;
}

static std::vector<at::Tensor> vjp_compute(const OpDef& op, const std::vector<at::Tensor>& in, RunCtx* rc) {
  const OpDef& fw = *op.graph->op(op.attrs.i("fw_op"));
  const auto gout_idx = op.attrs.ints("gout_idx");
  const auto diff_inputs = op.attrs.ints("diff_inputs");
  const int64_t nfw = op.attrs.i("num_fw_inputs");
  if (in.size() > 0 && in[0].is_meta()) {
    std::vector<at::Tensor> outs;
    for (auto i : diff_inputs) outs.push_back(at::empty_like(in[i]));
    return outs;
  }
  std::vector<at::Tensor> fw_in(in.begin(), in.begin() + nfw);
  std::vector<at::Tensor> leaves;
  {
    at::AutoGradMode gm(true);
    for (auto i : diff_inputs) {
      fw_in[i] = fw_in[i].detach().requires_grad_(true);
      leaves.push_back(fw_in[i]);
    }
    std::vector<at::Tensor> fw_out = fw.kernel->compute(fw, fw_in, rc);
    std::vector<at::Tensor> outs_sel, gouts_sel;
    for (size_t k = 0; k < gout_idx.size(); ++k) {
      const at::Tensor& o = fw_out[gout_idx[k]];
      if (!o.requires_grad()) continue;
      outs_sel.push_back(o);
      gouts_sel.push_back(in[nfw + k].to(o.scalar_type()).expand_as(o));
    }
    std::vector<at::Tensor> grads;
    if (!outs_sel.empty())
      grads = torch::autograd::grad(outs_sel, leaves, gouts_sel, /*retain_graph=*/false, /*create_graph=*/false,
                                    /*allow_unused=*/true);
    std::vector<at::Tensor> res;
    for (size_t k = 0; k < leaves.size(); ++k) {
      if (k < grads.size() && grads[k].defined()) res.push_back(grads[k].detach());
      else res.push_back(at::zeros_like(leaves[k]).detach());
    }
    return res;
  }
}

static void vjp_infer(OpDef& op) {
  const auto diff_inputs = op.attrs.ints("diff_inputs");
  op.outputs.resize(diff_inputs.size());
  for (size_t k = 0; k < diff_inputs.size(); ++k) {
    if (!op.outputs[k]) op.outputs[k] = std::make_shared<TensorDef>();
    op.outputs[k]->shape = op.inputs[diff_inputs[k]]->shape;
    op.outputs[k]->dtype = op.inputs[diff_inputs[k]]->dtype;
  }
}
static void vjp_deduce(OpDef& op, size_t s) {
  const auto diff_inputs = op.attrs.ints("diff_inputs");
  for (size_t k = 0; k < diff_inputs.size(); ++k) {
    const Tensor& in = op.inputs[diff_inputs[k]];
    if (!in->has_ds(s)) continue;
    auto& out = op.outputs[k];
    while (out->ds_hierarchy.size() <= s) out->ds_hierarchy.add(DistributedStatesUnion());
    out->ds_hierarchy.get_mut(s) = in->ds_hierarchy.get(s);
  }
}
static OpRegistrar _reg_vjp(OpKernel{"autograd_vjp", -1, kFlagNondiff, vjp_compute, nullptr, vjp_deduce, vjp_infer});

}  // namespace hb
