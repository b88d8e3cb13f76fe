// Leaf ops, elementwise / reduction / shape ops.  Long-tail ops are thin ATen
// closures (shape inference = the closure on meta tensors, gradients through the
// generic autograd VJP); ops on the training hot path get explicit gradients.
// (capability parity: hetu/graph/ops/{Arithmetics,Unary,Reduce,Reshape,Slice,Concat,
//  Split,Transpose,Broadcast,...}.cc and the Python names of codegen/ops.yml)
#include <ATen/ATen.h>

#include "exec.h"
#include "ir.h"
#include "op_utils.h"

namespace hb {

using Ts = std::vector<at::Tensor>;

// ------------------------------------------------------------------ leaves
static void leaf_infer(OpDef& op) {
  op.outputs.resize(1);
  if (!op.outputs[0]) op.outputs[0] = std::make_shared<TensorDef>();
  auto& o = op.outputs[0];
  o->dtype = dtype_from_name(op.attrs.s("dtype", "float32"));
  const auto gshape = op.attrs.ints("global_shape");
  const int s = op.graph ? op.graph->cur_strategy() : 0;
  if (op.dst_ds.size() > (size_t)s && op.dst_ds.get(s).size() > 0) o->shape = op.dst_ds.get(s).get(0).local_shape(gshape);
  else o->shape = gshape;
  o->ds_hierarchy = op.dst_ds;
  o->requires_grad = op.attrs.b("requires_grad", false);
  if (!op.sy_shape.empty()) o->symbolic_shape = op.sy_shape;
}
static void leaf_deduce(OpDef& op, size_t) { op.outputs[0]->ds_hierarchy = op.dst_ds; }
static Ts leaf_compute(const OpDef& op, const Ts&, RunCtx*) {
  HB_FAIL() << "leaf op " << op.name() << " must be fed / materialised by the executor";
}
HB_REGISTER_OP(placeholder, "placeholder", 1, kFlagPlaceholder | kFlagNondiff | kFlagNoMetaExec, leaf_compute, nullptr,
               leaf_deduce, leaf_infer);
HB_REGISTER_OP(variable, "variable", 1, kFlagVariable | kFlagNondiff | kFlagNoMetaExec, leaf_compute, nullptr,
               leaf_deduce, leaf_infer);

static void const_infer(OpDef& op) {
  op.outputs.resize(1);
  if (!op.outputs[0]) op.outputs[0] = std::make_shared<TensorDef>();
  HB_CHECK(op.const_data.defined()) << "const op without data";
  op.outputs[0]->shape = op.const_data.sizes().vec();
  op.outputs[0]->dtype = from_aten_dtype(op.const_data.scalar_type());
  op.outputs[0]->requires_grad = op.attrs.b("requires_grad", false);
  op.outputs[0]->ds_hierarchy = op.dst_ds;
}
static Ts const_compute(const OpDef& op, const Ts&, RunCtx*) { return {op.const_data}; }
HB_REGISTER_OP(const_tensor, "const", 1, kFlagConst | kFlagNondiff | kFlagNoMetaExec, const_compute, nullptr,
               leaf_deduce, const_infer);

// ------------------------------------------------------------------ helper macros
#define ATEN_OP(NAME, NOUT, ...)                                                                 \
  static Ts NAME##_compute(const OpDef& op, const Ts& in, RunCtx* rc) {                           \
    (void)op; (void)rc;                                                                           \
    __VA_ARGS__                                                                                   \
  }                                                                                               \
  HB_REGISTER_OP(NAME, #NAME, NOUT, 0, NAME##_compute, nullptr, nullptr, nullptr)

#define ATEN_OP_NODIFF(NAME, NOUT, ...)                                                          \
  static Ts NAME##_compute(const OpDef& op, const Ts& in, RunCtx* rc) {                           \
    (void)op; (void)rc;                                                                           \
    __VA_ARGS__                                                                                   \
  }                                                                                               \
  HB_REGISTER_OP(NAME, #NAME, NOUT, kFlagNondiff, NAME##_compute, nullptr, nullptr, nullptr)

// native generic kernel first, ATen when it does not apply (CPU, integer dtypes, ...)
#define GEN_OR(EXPR, FALLBACK)                    \
  {                                               \
    at::Tensor _r = (EXPR);                       \
    if (_r.defined()) return {_r};                \
    return {FALLBACK};                            \
  }

static at::Tensor like_const(const at::Tensor& ref, double v) {
  at::Tensor r = g_full(ref.sizes(), ref.options(), v);
  return r.defined() ? r : at::full_like(ref, v);
}

// ------------------------------------------------------------------ creation-like
ATEN_OP_NODIFF(ones_like, 1, return {like_const(in[0], 1.0)};);
ATEN_OP_NODIFF(zeros_like, 1, return {like_const(in[0], 0.0)};);
ATEN_OP_NODIFF(full_like, 1, return {like_const(in[0], op.attrs.f("value"))};);
ATEN_OP_NODIFF(arange, 0 + 1, {
  auto o = at::TensorOptions().dtype(to_aten_dtype(dtype_from_name(op.attrs.s("dtype", "int64"))));
  return {at::arange(op.attrs.f("start"), op.attrs.f("end"), op.attrs.f("step", 1.0), o)};
});

// n-ary sum (gradient accumulation of fan-out tensors)
static Ts sum_n_compute(const OpDef&, const Ts& in, RunCtx*) {
  at::Tensor acc = in[0];
  for (size_t i = 1; i < in.size(); ++i) acc = acc + in[i];
  return {acc};
}
static TensorList sum_n_grad(OpDef& op, const TensorList& g) { return TensorList(op.inputs.size(), g[0]); }
HB_REGISTER_OP(sum_n, "sum_n", 1, 0, sum_n_compute, sum_n_grad, nullptr, nullptr);
HB_REGISTER_OP(sum, "sum", 1, 0, sum_n_compute, sum_n_grad, nullptr, nullptr);

// ------------------------------------------------------------------ binary arithmetic (broadcasting)
// gradient of a broadcasting binary op: reduce the grad back to the operand's shape
static Tensor reduce_to_shape(Graph* g, const Tensor& grad, const Tensor& like) {
  if (grad->shape == like->shape) return grad;
  AttrMap a;
  a.set("shape", like->shape);
  // which target dims were broadcast (size 1 against a larger gradient dim): the other dims follow the gradient's RUNTIME
  // size, so the op stays valid when the token count changes from run to run
  const int lead = (int)grad->shape.size() - (int)like->shape.size();
  std::vector<int64_t> bcast;
  for (size_t j = 0; j < like->shape.size(); ++j)
    bcast.push_back(lead >= 0 && like->shape[j] == 1 && grad->shape[j + lead] != 1 ? 1 : 0);
  if (lead >= 0) a.set("bcast", bcast);
  return g->make_op1("reduce_to_shape", {grad}, a);
}
static Ts reduce_to_shape_compute(const OpDef& op, const Ts& in, RunCtx*) {
  std::vector<int64_t> target = op.attrs.ints("shape");
  const auto bcast = op.attrs.ints("bcast");
  if (bcast.size() == target.size()) {
    const int64_t lead = in[0].dim() - (int64_t)target.size();
    for (size_t j = 0; j < target.size(); ++j) target[j] = bcast[j] ? 1 : in[0].size((int64_t)j + lead);
  }
  return {at::sum_to(in[0], target)};
}
static void reduce_to_shape_deduce(OpDef& op, size_t s) {
  // summed-away dims that were sharded leave partial sums; surviving dims keep their split (shifted to the new rank)
  const Tensor& in = op.inputs[0];
  if (!in->has_ds(s)) return;
  const DistributedStates& ds = in->ds(s);
  const auto target = op.attrs.ints("shape");
  const int nd_in = in->ndim(), nd_out = (int)target.size();
  const int lead = nd_in - nd_out;
  std::map<int, int> st;
  int partial = ds.get_dim(kPartialDim);
  std::map<int, int> dim_map;
  for (auto& kv : ds.states()) {
    if (kv.second <= 1) continue;
    if (kv.first < 0) { if (kv.first == kDupDim) st[kDupDim] = kv.second; continue; }
    const int od = kv.first - lead;
    const bool reduced = od < 0 || (target[od] == 1 && in->shape[kv.first] != 1);
    if (reduced) partial *= kv.second;
    else { st[od] = kv.second; dim_map[kv.first] = od; }
  }
  if (partial > 1) st[kPartialDim] = partial;
  std::vector<int> order;
  for (int o : ds.order()) {
    int m = o < 0 ? o : (dim_map.count(o) ? dim_map[o] : kPartialDim);
    if (st.count(m) && std::find(order.begin(), order.end(), m) == order.end()) order.push_back(m);
  }
  set_out_ds(op, 0, s, DistributedStates(ds.device_num(), st, order));
}
HB_REGISTER_OP(reduce_to_shape, "reduce_to_shape", 1, 0, reduce_to_shape_compute, nullptr, reduce_to_shape_deduce, nullptr);

static Ts add_compute(const OpDef& op, const Ts& in, RunCtx*) {
  if (in.size() == 1) GEN_OR(g_unary(G_ADD_SCALAR, in[0], (float)op.attrs.f("value")), in[0] + op.attrs.f("value"));
  if (in[0].sizes() == in[1].sizes()) return {native_add(in[0], in[1])};
  GEN_OR(g_binary(B_ADD, in[0], in[1]), in[0] + in[1]);
}
static TensorList add_grad(OpDef& op, const TensorList& g) {
  TensorList r(op.inputs.size());
  for (size_t i = 0; i < op.inputs.size(); ++i) r[i] = reduce_to_shape(op.graph, g[0], op.inputs[i]);
  return r;
}
HB_REGISTER_OP(add, "add", 1, 0, add_compute, add_grad, nullptr, nullptr);

static Ts sub_compute(const OpDef& op, const Ts& in, RunCtx*) {
  if (in.size() == 1) {
    if (op.attrs.b("from_const")) GEN_OR(g_unary(G_RSUB_SCALAR, in[0], (float)op.attrs.f("value")), op.attrs.f("value") - in[0]);
    GEN_OR(g_unary(G_ADD_SCALAR, in[0], -(float)op.attrs.f("value")), in[0] - op.attrs.f("value"));
  }
  GEN_OR(g_binary(B_SUB, in[0], in[1]), in[0] - in[1]);
}
static TensorList sub_grad(OpDef& op, const TensorList& g) {
  Graph* gr = op.graph;
  if (op.inputs.size() == 1) return {op.attrs.b("from_const") ? gr->make_op1("neg", {g[0]}) : g[0]};
  return {reduce_to_shape(gr, g[0], op.inputs[0]), reduce_to_shape(gr, gr->make_op1("neg", {g[0]}), op.inputs[1])};
}
HB_REGISTER_OP(sub, "sub", 1, 0, sub_compute, sub_grad, nullptr, nullptr);

static Ts mul_compute(const OpDef& op, const Ts& in, RunCtx*) {
  if (in.size() == 1) GEN_OR(g_unary(G_MUL_SCALAR, in[0], (float)op.attrs.f("value")), in[0] * op.attrs.f("value"));
  GEN_OR(g_binary(B_MUL, in[0], in[1]), in[0] * in[1]);
}
static TensorList mul_grad(OpDef& op, const TensorList& g) {
  Graph* gr = op.graph;
  if (op.inputs.size() == 1) {
    AttrMap a;
    a.set("value", op.attrs.f("value"));
    return {gr->make_op1("mul", {g[0]}, a)};
  }
  return {reduce_to_shape(gr, gr->make_op1("mul", {g[0], op.inputs[1]}), op.inputs[0]),
          reduce_to_shape(gr, gr->make_op1("mul", {g[0], op.inputs[0]}), op.inputs[1])};
}
HB_REGISTER_OP(mul, "mul", 1, 0, mul_compute, mul_grad, nullptr, nullptr);

ATEN_OP(div, 1, {
  if (in.size() == 1) {
    if (op.attrs.b("from_const")) GEN_OR(g_unary(G_RDIV_SCALAR, in[0], (float)op.attrs.f("value")), at::reciprocal(in[0]) * op.attrs.f("value"));
    GEN_OR(g_unary(G_DIV_SCALAR, in[0], (float)op.attrs.f("value")), in[0] / op.attrs.f("value"));
  }
  GEN_OR(g_binary(B_DIV, in[0], in[1]), in[0] / in[1]);
});
ATEN_OP(pow, 1, GEN_OR(g_unary(G_POW, in[0], (float)op.attrs.f("exponent")), at::pow(in[0], op.attrs.f("exponent"))));
ATEN_OP(neg, 1, GEN_OR(g_unary(G_NEG, in[0]), at::neg(in[0])));
ATEN_OP(reciprocal, 1, GEN_OR(g_unary(G_RECIPROCAL, in[0]), at::reciprocal(in[0])));
ATEN_OP(abs, 1, GEN_OR(g_unary(G_ABS, in[0]), at::abs(in[0])));
ATEN_OP_NODIFF(ceil, 1, GEN_OR(g_unary(G_CEIL, in[0]), at::ceil(in[0])));
ATEN_OP_NODIFF(floor, 1, GEN_OR(g_unary(G_FLOOR, in[0]), at::floor(in[0])));
ATEN_OP_NODIFF(round, 1, GEN_OR(g_unary(G_ROUND, in[0]), at::round(in[0])));
ATEN_OP(exp, 1, GEN_OR(g_unary(G_EXP, in[0]), at::exp(in[0])));
ATEN_OP(log, 1, GEN_OR(g_unary(G_LOG, in[0]), at::log(in[0])));
ATEN_OP(sqrt, 1, GEN_OR(g_unary(G_SQRT, in[0]), at::sqrt(in[0])));
ATEN_OP(rsqrt, 1, GEN_OR(g_unary(G_RSQRT, in[0]), at::rsqrt(in[0])));
ATEN_OP(sin, 1, GEN_OR(g_unary(G_SIN, in[0]), at::sin(in[0])));
ATEN_OP(cos, 1, GEN_OR(g_unary(G_COS, in[0]), at::cos(in[0])));
ATEN_OP(clamp, 1, GEN_OR(g_unary(G_CLAMP, in[0], (float)op.attrs.f("min"), (float)op.attrs.f("max")), at::clamp(in[0], op.attrs.f("min"), op.attrs.f("max"))));
ATEN_OP_NODIFF(bool_op, 1, return {in[0] != 0};);
ATEN_OP(where, 1, return {at::where(in[0].to(at::kBool), in[1], in[2])};);
ATEN_OP(masked_fill, 1, return {at::masked_fill(in[0], in[1].to(at::kBool), op.attrs.f("value"))};);
ATEN_OP(triu, 1, return {op.attrs.b("lower") ? at::tril(in[0], op.attrs.i("diagonal")) : at::triu(in[0], op.attrs.i("diagonal"))};);
ATEN_OP_NODIFF(onehot, 1, return {at::one_hot(in[0].to(at::kLong), op.attrs.i("num_classes")).to(at::kFloat)};);
ATEN_OP_NODIFF(checknumeric, 1, return {(at::isnan(in[0]).sum() + at::isinf(in[0]).sum()).to(at::kFloat)};);
ATEN_OP_NODIFF(check_finite, 1, return {at::logical_not(at::isfinite(in[0])).any().to(at::kFloat).reshape({1})};);
ATEN_OP_NODIFF(range_mask, 1, {
  auto x = in[0];
  return {at::logical_or(x < op.attrs.i("min"), x > op.attrs.i("max")).to(x.scalar_type())};
});

// ------------------------------------------------------------------ activations (long tail; hot ones live in ops_nn.cc)
ATEN_OP(sigmoid, 1, GEN_OR(g_unary(G_SIGMOID, in[0]), at::sigmoid(in[0])));
ATEN_OP(tanh, 1, GEN_OR(g_unary(G_TANH, in[0]), at::tanh(in[0])));
ATEN_OP(leakyrelu, 1, GEN_OR(g_unary(G_LEAKYRELU, in[0], (float)op.attrs.f("alpha", 0.01)), at::leaky_relu(in[0], op.attrs.f("alpha", 0.01))));
ATEN_OP(elu, 1, GEN_OR(g_unary(G_ELU, in[0], (float)op.attrs.f("alpha", 1.0), (float)op.attrs.f("scale", 1.0)), at::elu(in[0], op.attrs.f("alpha", 1.0), op.attrs.f("scale", 1.0))));
ATEN_OP(hardshrink, 1, GEN_OR(g_unary(G_HARDSHRINK, in[0], (float)op.attrs.f("lambda", 0.5)), at::hardshrink(in[0], op.attrs.f("lambda", 0.5))));
ATEN_OP(hardsigmoid, 1, GEN_OR(g_unary(G_HARDSIGMOID, in[0]), at::hardsigmoid(in[0])));
ATEN_OP(hardtanh, 1, GEN_OR(g_unary(G_HARDTANH, in[0], (float)op.attrs.f("min_val", -1.0), (float)op.attrs.f("max_val", 1.0)), at::hardtanh(in[0], op.attrs.f("min_val", -1.0), op.attrs.f("max_val", 1.0))));
ATEN_OP(hardswish, 1, GEN_OR(g_unary(G_HARDSWISH, in[0]), at::hardswish(in[0])));
ATEN_OP(logsigmoid, 1, GEN_OR(g_unary(G_LOGSIGMOID, in[0]), at::log_sigmoid(in[0])));
ATEN_OP(mish, 1, GEN_OR(g_unary(G_MISH, in[0]), at::mish(in[0])));
ATEN_OP(softplus, 1, GEN_OR(g_unary(G_SOFTPLUS, in[0], (float)op.attrs.f("beta", 1.0), (float)op.attrs.f("threshold", 20.0)), at::softplus(in[0], op.attrs.f("beta", 1.0), op.attrs.f("threshold", 20.0))));
ATEN_OP(softshrink, 1, GEN_OR(g_unary(G_SOFTSHRINK, in[0], (float)op.attrs.f("lambda", 0.5)), at::softshrink(in[0], op.attrs.f("lambda", 0.5))));
ATEN_OP(softmax, 1, GEN_OR(g_softmax(false, in[0], op.attrs.i("dim", -1)), at::softmax(in[0], op.attrs.i("dim", -1))));
ATEN_OP(log_softmax, 1, GEN_OR(g_softmax(true, in[0], op.attrs.i("dim", -1)), at::log_softmax(in[0], op.attrs.i("dim", -1))));

// ------------------------------------------------------------------ reductions
static Ts reduce_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const std::string mode = op.attrs.s("mode", "sum");
  auto axes = op.attrs.ints("axes");
  const bool keep = op.attrs.b("keepdims");
  const at::Tensor& x = in[0];
  if (axes.empty()) for (int64_t i = 0; i < x.dim(); ++i) axes.push_back(i);
  {
    const int gm = mode == "sum" ? R_SUM : (mode == "mean" || mode == "avg") ? R_MEAN : mode == "max" ? R_MAX : mode == "min" ? R_MIN
                   : mode == "prod" ? R_PROD : -1;
    if (gm >= 0) {
      at::Tensor r = g_reduce(gm, x, axes, keep);
      if (r.defined()) return {r};
    }
  }
  if (mode == "sum") return {at::sum(x, axes, keep)};
  if (mode == "mean" || mode == "avg") return {at::mean(x, axes, keep)};
  if (mode == "max") return {at::amax(x, axes, keep)};
  if (mode == "min") return {at::amin(x, axes, keep)};
  if (mode == "prod") {
    at::Tensor r = x;
    std::vector<int64_t> sorted = axes;
    for (auto& a : sorted) if (a < 0) a += x.dim();
    std::sort(sorted.rbegin(), sorted.rend());
    for (auto a : sorted) r = at::prod(r, a, keep);
    return {r};
  }
  HB_FAIL() << "unknown reduce mode " << mode;
}
static void reduce_deduce(OpDef& op, size_t s) {
  // reducing over a split dim leaves a partial result; other split dims shift down when not keepdims
  const Tensor& in = op.inputs[0];
  if (!in->has_ds(s)) return;
  const DistributedStates& ds = in->ds(s);
  auto axes = op.attrs.ints("axes");
  const int nd = in->ndim();
  if (axes.empty()) for (int i = 0; i < nd; ++i) axes.push_back(i);
  for (auto& a : axes) if (a < 0) a += nd;
  const bool keep = op.attrs.b("keepdims");
  std::map<int, int> st;
  std::vector<int> order;
  int partial = ds.get_dim(kPartialDim);
  auto remap = [&](int d) {
    if (d < 0 || keep) return d;
    int shift = 0;
    for (auto a : axes) if (a < d) ++shift;
    return d - shift;
  };
  for (auto& kv : ds.states()) {
    if (kv.second <= 1) continue;
    if (kv.first >= 0 && std::find(axes.begin(), axes.end(), (int64_t)kv.first) != axes.end()) partial *= kv.second;
    else if (kv.first != kPartialDim) st[remap(kv.first)] = kv.second;
  }
  if (partial > 1) st[kPartialDim] = partial;
  bool partial_placed = false;
  for (int o : ds.order()) {
    const bool reduced = o >= 0 && std::find(axes.begin(), axes.end(), (int64_t)o) != axes.end();
    if (reduced || o == kPartialDim) {
      if (!partial_placed) { order.push_back(kPartialDim); partial_placed = true; }
    } else order.push_back(remap(o));
  }
  auto& out = op.outputs[0];
  while (out->ds_hierarchy.size() <= s) out->ds_hierarchy.add(DistributedStatesUnion());
  out->ds_hierarchy.get_mut(s) = DistributedStatesUnion({DistributedStates(ds.device_num(), st, order)});
}
HB_REGISTER_OP(reduce, "reduce", 1, 0, reduce_compute, nullptr, reduce_deduce, nullptr);
ATEN_OP(mean, 1, GEN_OR(g_reduce(R_MEAN, in[0], {}, false), at::mean(in[0])));
ATEN_OP(norm, 1, return {at::norm(in[0], op.attrs.f("p", 2.0), op.attrs.ints("axes").empty() ? std::vector<int64_t>{} : op.attrs.ints("axes"), op.attrs.b("keepdims"))};);

// ------------------------------------------------------------------ shape / view ops
static std::vector<int64_t> resolve_shape(const OpDef& op) {
  if (!op.sy_shape.empty()) return sy_shape_values(op.sy_shape);
  return op.attrs.ints("shape");
}
static Ts reshape_compute(const OpDef& op, const Ts& in, RunCtx*) { return {in[0].reshape(resolve_shape(op))}; }
static TensorList reshape_grad(OpDef& op, const TensorList& g) {
  AttrMap a;
  a.set("shape", op.inputs[0]->shape);
  Graph* gr = op.graph;
  OpDef* fw = &op;
  return {gr->make_op1("reshape", {g[0]}, a, {}, [fw](OpDef& o) {
    if (!fw->inputs[0]->symbolic_shape.empty()) o.sy_shape = fw->inputs[0]->symbolic_shape;
  })};
}
static void reshape_deduce(OpDef& op, size_t s) {
  // supported: the split dims are leading dims preserved by the reshape, or the reshape merges / splits
  // trailing dims only.  A split of input dim d maps to the output dim with the same leading element
  // offset (prefix product match).
  const Tensor& in = op.inputs[0];
  if (!in->has_ds(s)) return;
  const DistributedStates& ds = in->ds(s);
  const auto& ishape = in->shape;
  const auto& oshape = op.outputs[0]->shape;
  std::map<int, int> st;
  std::map<int, int> dim_map;
  for (auto& kv : ds.states()) {
    if (kv.second <= 1) continue;
    if (kv.first < 0) { st[kv.first] = kv.second; continue; }
    int64_t prefix = 1;
    for (int i = 0; i < kv.first; ++i) prefix *= ishape[i];
    int64_t acc = 1;
    int od = -1;
    for (int j = 0; j < (int)oshape.size(); ++j) {
      if (acc == prefix) { od = j; break; }
      acc *= oshape[j];
    }
    HB_CHECK(od >= 0) << "reshape of " << op.inputs[0]->name << " moves split dim " << kv.first
                      << " to a position that is not a dim boundary";
    st[od] = kv.second;
    dim_map[kv.first] = od;
  }
  std::vector<int> order;
  for (int o : ds.order()) order.push_back(o < 0 ? o : dim_map[o]);
  auto& out = op.outputs[0];
  while (out->ds_hierarchy.size() <= s) out->ds_hierarchy.add(DistributedStatesUnion());
  out->ds_hierarchy.get_mut(s) = DistributedStatesUnion({DistributedStates(ds.device_num(), st, order)});
}
HB_REGISTER_OP(reshape, "reshape", 1, 0, reshape_compute, reshape_grad, reshape_deduce, nullptr);

static Ts transpose_compute(const OpDef& op, const Ts& in, RunCtx*) {
  auto perm = op.attrs.ints("perm");
  if (perm.empty()) for (int64_t i = in[0].dim() - 1; i >= 0; --i) perm.push_back(i);
  return {in[0].permute(perm).contiguous()};
}
static TensorList transpose_grad(OpDef& op, const TensorList& g) {
  auto perm = op.attrs.ints("perm");
  const int nd = op.inputs[0]->ndim();
  if (perm.empty()) for (int i = nd - 1; i >= 0; --i) perm.push_back(i);
  std::vector<int64_t> inv(nd);
  for (int i = 0; i < nd; ++i) inv[perm[i]] = i;
  AttrMap a;
  a.set("perm", inv);
  return {op.graph->make_op1("transpose", {g[0]}, a)};
}
static void transpose_deduce(OpDef& op, size_t s) {
  const Tensor& in = op.inputs[0];
  if (!in->has_ds(s)) return;
  const DistributedStates& ds = in->ds(s);
  auto perm = op.attrs.ints("perm");
  const int nd = in->ndim();
  if (perm.empty()) for (int i = nd - 1; i >= 0; --i) perm.push_back(i);
  std::vector<int> inv(nd);
  for (int i = 0; i < nd; ++i) inv[perm[i]] = i;
  std::map<int, int> st;
  for (auto& kv : ds.states()) if (kv.second > 1) st[kv.first < 0 ? kv.first : inv[kv.first]] = kv.second;
  std::vector<int> order;
  for (int o : ds.order()) order.push_back(o < 0 ? o : inv[o]);
  auto& out = op.outputs[0];
  while (out->ds_hierarchy.size() <= s) out->ds_hierarchy.add(DistributedStatesUnion());
  out->ds_hierarchy.get_mut(s) = DistributedStatesUnion({DistributedStates(ds.device_num(), st, order)});
}
HB_REGISTER_OP(transpose, "transpose", 1, 0, transpose_compute, transpose_grad, transpose_deduce, nullptr);

ATEN_OP(contiguous, 1, return {in[0].contiguous()};);
ATEN_OP(as_strided, 1, return {at::as_strided(in[0], op.attrs.ints("shape"), op.attrs.ints("stride"), op.attrs.i("storage_offset")).contiguous()};);
ATEN_OP(diagonal, 1, return {at::diagonal(in[0], op.attrs.i("offset"), op.attrs.i("dim1", 0), op.attrs.i("dim2", 1)).contiguous()};);
ATEN_OP(broadcast, 1, {
  auto shape = op.attrs.ints("shape");
  auto add_axes = op.attrs.ints("add_axes");
  at::Tensor x = in[0];
  for (auto a : add_axes) x = x.unsqueeze(a);
  return {x.expand(shape).contiguous()};
});
ATEN_OP(repeat, 1, return {in[0].repeat(op.attrs.ints("repeats"))};);
ATEN_OP(roll, 1, return {at::roll(in[0], op.attrs.ints("shifts"), op.attrs.ints("dims"))};);
ATEN_OP(pad, 1, return {at::constant_pad_nd(in[0], op.attrs.ints("paddings"), op.attrs.f("value"))};);
ATEN_OP(gather, 1, return {at::gather(in[0], op.attrs.i("dim"), in[1].to(at::kLong))};);
// index / order statistics of the v1 op set (ref: hetu/v1 Argmax, Argsort, TopKIdx / TopKVal, Cumsum, Sign, Scatter)
ATEN_OP_NODIFF(argmax, 1, return {at::argmax(in[0], op.attrs.i("dim", -1), op.attrs.b("keepdims"))};);
ATEN_OP_NODIFF(argsort, 1, return {at::argsort(in[0], op.attrs.i("dim", -1), op.attrs.b("descending"))};);
ATEN_OP(topk, 2, {
  auto r = at::topk(in[0], op.attrs.i("k", 1), op.attrs.i("dim", -1), op.attrs.b("largest", true), true);
  return {std::get<0>(r), std::get<1>(r)};
});
ATEN_OP(cumsum, 1, return {at::cumsum(in[0], op.attrs.i("dim", -1))};);
ATEN_OP_NODIFF(sign, 1, return {at::sign(in[0])};);
ATEN_OP(scatter, 1, return {at::scatter(in[0], op.attrs.i("dim"), in[1].to(at::kLong), in[2])};);
ATEN_OP_NODIFF(unique_consecutive_count, 2, {
  // sorted unique values of a 1-D tensor and, per input element, the index of its value among them (fixed-size outputs: the unique
  // list is padded with the last value to the input length; `count` tells how many entries are real)
  auto flat = in[0].reshape({-1});
  auto r = at::_unique2(flat, true, true, false);
  at::Tensor u = std::get<0>(r), inv = std::get<1>(r);
  at::Tensor padded = u.numel() > 0 ? at::cat({u, u.slice(0, u.numel() - 1, u.numel()).expand({flat.numel() - u.numel()})}) : flat;
  return {padded, inv};
});
ATEN_OP(index_add, 1, return {at::index_add(in[0], op.attrs.i("dim"), in[1].to(at::kLong), in[2])};);
ATEN_OP(interpolate, 1, {
  auto size = op.attrs.ints("size");
  const std::string mode = op.attrs.s("mode", "bilinear");
  if (mode == "nearest") return {at::upsample_nearest2d(in[0], size)};
  if (mode == "bicubic") return {at::upsample_bicubic2d(in[0], size, op.attrs.b("align_corners"))};
  return {at::upsample_bilinear2d(in[0], size, op.attrs.b("align_corners"))};
});

static Ts slice_compute(const OpDef& op, const Ts& in, RunCtx*) {
  std::vector<int64_t> begin = op.attrs.ints("begin"), size = op.attrs.ints("size");
  if (!op.sy_shape.empty()) size = sy_shape_values(op.sy_shape);
  at::Tensor x = in[0];
  for (size_t d = 0; d < begin.size(); ++d) {
    const int64_t len = size[d] < 0 ? x.size(d) - begin[d] : size[d];
    x = x.narrow((int64_t)d, begin[d], len);
  }
  return {x.contiguous()};
}
HB_REGISTER_OP(slice, "slice", 1, 0, slice_compute, nullptr, nullptr, nullptr);

static Ts split_compute(const OpDef& op, const Ts& in, RunCtx*) {
  // split(x, num_chunks | sections, dim) -> chunks
  const int64_t dim = op.attrs.i("dim");
  auto sections = op.attrs.ints("sections");
  std::vector<at::Tensor> parts;
  if (sections.empty()) parts = at::chunk(in[0], op.attrs.i("num_chunks"), dim);
  else parts = at::split_with_sizes(in[0], sections, dim);
  for (auto& p : parts) p = p.contiguous();
  return parts;
}
static TensorList split_grad(OpDef& op, const TensorList& g) {
  Graph* gr = op.graph;
  TensorList parts;
  for (size_t i = 0; i < g.size(); ++i) parts.push_back(g[i] ? g[i] : gr->make_op1("zeros_like", {op.outputs[i]}));
  AttrMap a;
  a.set("dim", op.attrs.i("dim"));
  return {gr->make_op1("concat", parts, a)};
}
HB_REGISTER_OP(split, "split", -1, 0, split_compute, split_grad, nullptr, nullptr);

static Ts concat_compute(const OpDef& op, const Ts& in, RunCtx*) { GEN_OR(g_concat(in, op.attrs.i("dim")), at::cat(in, op.attrs.i("dim"))); }
static TensorList concat_grad(OpDef& op, const TensorList& g) {
  std::vector<int64_t> sections;
  const int64_t dim = op.attrs.i("dim");
  for (auto& t : op.inputs) sections.push_back(t->shape[dim < 0 ? dim + t->ndim() : dim]);
  AttrMap a;
  a.set("dim", dim);
  a.set("sections", sections);
  return op.graph->make_op("split", {g[0]}, a);
}
HB_REGISTER_OP(concat, "concat", 1, 0, concat_compute, concat_grad, nullptr, nullptr);
HB_REGISTER_OP(dynamic_concat, "dynamic_concat", 1, 0, concat_compute, concat_grad, nullptr, nullptr);

// dtype / device transfer (autocast inserts these)
static Ts data_transfer_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const at::ScalarType to = to_aten_dtype(dtype_from_name(op.attrs.s("dtype")));
  GEN_OR(g_cast(in[0], to), in[0].to(to));
}
static TensorList data_transfer_grad(OpDef& op, const TensorList& g) {
  AttrMap a;
  a.set("dtype", std::string(dtype_name(op.inputs[0]->dtype)));
  return {op.graph->make_op1("data_transfer", {g[0]}, a)};
}
HB_REGISTER_OP(data_transfer, "data_transfer", 1, kFlagDataTransfer, data_transfer_compute, data_transfer_grad, nullptr,
               nullptr);

// group: pure control op tying several tensors (train_op)
static Ts group_compute(const OpDef&, const Ts& in, RunCtx*) {
  (void)in;
  return {at::zeros({1})};
}
static void group_infer(OpDef& op) {
  op.outputs.resize(1);
  if (!op.outputs[0]) op.outputs[0] = std::make_shared<TensorDef>();
  op.outputs[0]->shape = {1};
  op.outputs[0]->dtype = DataType::FLOAT32;
}
HB_REGISTER_OP(group, "group", 1, kFlagGroup | kFlagNondiff | kFlagNoMetaExec, group_compute, nullptr, nullptr, group_infer);

// ------------------------------------------------------------------ classic NN long tail (CNN tests)
ATEN_OP(conv2d, 1, {
  c10::optional<at::Tensor> bias;
  if (in.size() > 2) bias = in[2];
  const int64_t p = op.attrs.i("padding"), st = op.attrs.i("stride", 1);
  return {at::conv2d(in[0], in[1], bias, {st, st}, {p, p})};
});
ATEN_OP(avgpool, 1, {
  const int64_t k1 = op.attrs.i("kernel_H"), k2 = op.attrs.i("kernel_W"), p = op.attrs.i("padding"), st = op.attrs.i("stride", 1);
  return {at::avg_pool2d(in[0], {k1, k2}, {st, st}, {p, p})};
});
ATEN_OP(maxpool, 1, {
  const int64_t k1 = op.attrs.i("kernel_H"), k2 = op.attrs.i("kernel_W"), p = op.attrs.i("padding"), st = op.attrs.i("stride", 1);
  return {at::max_pool2d(in[0], {k1, k2}, {st, st}, {p, p})};
});
ATEN_OP(batch_norm, 1, {
  return {at::batch_norm(in[0], in[1], in[2], in[3], in[4], true, op.attrs.f("momentum", 0.1), op.attrs.f("eps", 1e-5), false)};
});
ATEN_OP(instance_norm, 1, {
  return {at::instance_norm(in[0], {}, {}, {}, {}, true, 0.1, op.attrs.f("eps", 1e-7), false)};
});
ATEN_OP(bmm, 1, return {at::bmm(in[0], in[1])};);
ATEN_OP(dot, 1, return {at::matmul(in[0], in[1])};);
ATEN_OP(einsum, 1, return {at::einsum(op.attrs.s("equation"), in)};);
ATEN_OP(outer, 1, return {at::outer(in[0], in[1])};);

// losses (mean / sum / none reductions)
static at::Tensor apply_reduction(const at::Tensor& l, const std::string& r) {
  if (r == "mean") return l.mean();
  if (r == "sum") return l.sum();
  return l;
}
ATEN_OP(mse_loss, 1, return {apply_reduction(at::pow(in[0] - in[1], 2), op.attrs.s("reduction", "mean"))};);
ATEN_OP(binary_cross_entropy, 1, {
  auto l = -(in[1] * at::log(in[0].clamp_min(1e-12)) + (1 - in[1]) * at::log((1 - in[0]).clamp_min(1e-12)));
  return {apply_reduction(l, op.attrs.s("reduction", "mean"))};
});
ATEN_OP(nll_loss, 1, {
  auto l = -at::gather(in[0], 1, in[1].to(at::kLong).unsqueeze(1)).squeeze(1);
  return {apply_reduction(l, op.attrs.s("reduction", "mean"))};
});
ATEN_OP(kl_div, 1, {
  auto l = in[1] * (at::log(in[1].clamp_min(1e-12)) - in[0]);
  return {apply_reduction(l, op.attrs.s("reduction", "mean"))};
});
ATEN_OP(softmax_cross_entropy, 1, {
  auto l = -(in[1] * at::log_softmax(in[0], -1)).sum(-1);
  return {apply_reduction(l, op.attrs.s("reduction", "mean"))};
});

}  // namespace hb

// ------------------------------------------------------------------ blockwise quantisation (int8 / nf4 / fp4)
// (capability parity: hetu/graph/ops/Quantization.cc + hetu/impl/kernel/Quantization.cu, which wrap bitsandbytes'
//  blockwise absmax quantisers; 4-bit codes are packed two per byte, high nibble first)
namespace hb {
namespace {
at::Tensor code_table(const std::string& kind, const at::TensorOptions& o) {
  static const float nf4[16] = {-1.0f, -0.6961928f, -0.5250730f, -0.3949175f, -0.2844414f, -0.1848449f, -0.0910500f, 0.0f,
                                0.0795803f, 0.1609302f, 0.2461123f, 0.3379152f, 0.4407098f, 0.5626170f, 0.7229568f, 1.0f};
  static const float fp4[16] = {0.0f, 0.0052083f, 0.6666667f, 1.0f, 0.3333333f, 0.5f, 0.1666667f, 0.25f,
                                -0.0f, -0.0052083f, -0.6666667f, -1.0f, -0.3333333f, -0.5f, -0.1666667f, -0.25f};
  const float* t = kind == "fp4" ? fp4 : nf4;
  return at::from_blob(const_cast<float*>(t), {16}, at::TensorOptions().dtype(at::kFloat)).clone().to(o.device());
}
int kind_code(const std::string& kind) { return kind == "int8" ? 0 : (kind == "fp4" ? 2 : 1); }
bool quant_native(const at::Tensor& t) {
  return t.is_cuda() && env_int("HETU_B200_FORCE_CPU", 0) == 0 && (t.scalar_type() == at::kBFloat16 || t.scalar_type() == at::kFloat);
}
std::vector<at::Tensor> quantize_blockwise(const at::Tensor& x, const std::string& kind, int64_t bs) {
  if (quant_native(x) && bs % 2 == 0) {
    // native sm_100a kernel: one warp per block, absmax + encode in a single pass (csrc/kernels/quant_block.cu)
    at::Tensor xc = x.contiguous();
    const int64_t n = xc.numel(), nb = (n + bs - 1) / bs;
    at::Tensor absmax = at::empty({nb}, xc.options().dtype(at::kFloat));
    at::Tensor q = kind == "int8" ? at::empty(xc.sizes(), xc.options().dtype(at::kChar)) : at::empty({nb * bs / 2}, xc.options().dtype(at::kByte));
    cuda_ok(hb::quantize_blockwise(xc.data_ptr(), xc.scalar_type() == at::kBFloat16, q.data_ptr(), absmax.data_ptr<float>(), n, (int)bs,
                                   kind_code(kind), cur_stream()), "quantize_blockwise");
    return {q, absmax};
  }
  at::Tensor flat = x.to(at::kFloat).reshape({-1});
  const int64_t n = flat.numel(), nb = (n + bs - 1) / bs;
  at::Tensor padded = at::zeros({nb * bs}, flat.options());
  padded.narrow(0, 0, n).copy_(flat);
  at::Tensor blocks = padded.reshape({nb, bs});
  at::Tensor absmax = blocks.abs().amax(1).clamp_min(1e-12);
  at::Tensor norm = blocks / absmax.unsqueeze(1);
  if (kind == "int8") {
    at::Tensor q = at::round(norm * 127.0).clamp(-127, 127).to(at::kChar);
    return {q.reshape({-1}).narrow(0, 0, n).reshape(x.sizes()), absmax};
  }
  at::Tensor table = code_table(kind, flat.options());
  at::Tensor idx = (norm.unsqueeze(-1) - table.reshape({1, 1, 16})).abs().argmin(-1).to(at::kByte).reshape({-1});   // nearest code
  at::Tensor hi = idx.slice(0, 0, nb * bs, 2), lo = idx.slice(0, 1, nb * bs, 2);
  return {(hi * 16 + lo).to(at::kByte), absmax};
}
at::Tensor dequantize_blockwise(const at::Tensor& q, const at::Tensor& absmax, const std::string& kind, int64_t bs,
                                const std::vector<int64_t>& shape, at::ScalarType dt) {
  int64_t n = 1;
  for (auto s : shape) n *= s;
  if (q.is_cuda() && env_int("HETU_B200_FORCE_CPU", 0) == 0 && absmax.scalar_type() == at::kFloat && q.is_contiguous() &&
      (dt == at::kBFloat16 || dt == at::kFloat)) {
    at::Tensor out = at::empty(shape, q.options().dtype(dt));
    cuda_ok(hb::dequantize_blockwise(q.data_ptr(), absmax.contiguous().data_ptr<float>(), out.data_ptr(), dt == at::kBFloat16, n, (int)bs,
                                     kind_code(kind), cur_stream()), "dequantize_blockwise");
    return out;
  }
  at::Tensor vals;
  if (kind == "int8") vals = q.to(at::kFloat).reshape({-1}) / 127.0;
  else {
    at::Tensor table = code_table(kind, absmax.options());
    at::Tensor b = q.reshape({-1}).to(at::kLong);
    at::Tensor idx = at::stack({at::floor_divide(b, 16), at::remainder(b, 16)}, 1).reshape({-1});
    vals = table.index_select(0, idx);
  }
  const int64_t nb = absmax.numel();
  at::Tensor padded = at::zeros({nb * bs}, vals.options());
  padded.narrow(0, 0, std::min<int64_t>(vals.numel(), nb * bs)).copy_(vals.narrow(0, 0, std::min<int64_t>(vals.numel(), nb * bs)));
  at::Tensor out = (padded.reshape({nb, bs}) * absmax.to(at::kFloat).unsqueeze(1)).reshape({-1}).narrow(0, 0, n);
  return out.reshape(shape).to(dt);
}
}  // namespace

ATEN_OP_NODIFF(quantize_blockwise, 2, {
  if (in[0].is_meta()) {
    const std::string kind = op.attrs.s("kind", "int8");
    const int64_t bs = op.attrs.i("blocksize", 64), n = in[0].numel(), nb = (n + bs - 1) / bs;
    if (kind == "int8") return {at::empty(in[0].sizes(), in[0].options().dtype(at::kChar)), at::empty({nb}, in[0].options().dtype(at::kFloat))};
    return {at::empty({nb * bs / 2}, in[0].options().dtype(at::kByte)), at::empty({nb}, in[0].options().dtype(at::kFloat))};
  }
  return quantize_blockwise(in[0], op.attrs.s("kind", "int8"), op.attrs.i("blocksize", 64));
});
ATEN_OP_NODIFF(dequantize_blockwise, 1, {
  const std::vector<int64_t> shape = op.attrs.ints("shape");
  const at::ScalarType dt = to_aten_dtype(dtype_from_name(op.attrs.s("dtype", "float32")));
  if (in[0].is_meta()) return {at::empty(shape, in[0].options().dtype(dt))};
  return {dequantize_blockwise(in[0], in[1], op.attrs.s("kind", "int8"), op.attrs.i("blocksize", 64), shape, dt)};
});
// y = x @ dequant(w_q)^T : 4-bit frozen base weights (QLoRA-style); differentiable w.r.t. x only
static Ts matmul4bit_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const std::vector<int64_t> wshape = op.attrs.ints("weight_shape");
  if (in[0].is_meta()) {
    std::vector<int64_t> o = in[0].sizes().vec();
    o.back() = wshape[0];
    return {at::empty(o, in[0].options())};
  }
  at::Tensor w = dequantize_blockwise(in[1], in[2], op.attrs.s("kind", "nf4"), op.attrs.i("blocksize", 64), wshape, in[0].scalar_type());
  // bf16 on the GPU: 4-bit weights are expanded once into a bf16 tile stream and multiplied on the tcgen05 GEMM
  at::Tensor x2 = in[0].reshape({-1, in[0].size(-1)}).contiguous();
  if (is_native(x2) && is_native(w) && x2.size(1) % 8 == 0 && (reinterpret_cast<uintptr_t>(x2.data_ptr()) % 16) == 0) {
    std::vector<int64_t> o = in[0].sizes().vec();
    o.back() = wshape[0];
    at::Tensor y = at::empty({x2.size(0), wshape[0]}, x2.options());
    GemmCall g;
    g.A = x2.data_ptr(); g.B = w.data_ptr(); g.C = y.data_ptr();
    g.M = (int)x2.size(0); g.N = (int)wshape[0]; g.K = (int)x2.size(1);
    g.lda = x2.stride(0); g.ldb = w.stride(0); g.ldc = wshape[0];
    g.out = GemmOut::BF16;
    if (gemm_bf16(g, cur_stream()) == cudaSuccess) return {y.reshape(o)};
  }
  return {at::matmul(in[0], w.t())};
}
static TensorList matmul4bit_grad(OpDef& op, const TensorList& g) {
  AttrMap a = op.attrs;
  a.set("transpose", true);
  return {op.graph->make_op1("matmul4bit_dgrad", {g[0], op.inputs[1], op.inputs[2]}, a), nullptr, nullptr};
}
static Ts matmul4bit_dgrad_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const std::vector<int64_t> wshape = op.attrs.ints("weight_shape");
  if (in[0].is_meta()) {
    std::vector<int64_t> o = in[0].sizes().vec();
    o.back() = wshape[1];
    return {at::empty(o, in[0].options())};
  }
  at::Tensor w = dequantize_blockwise(in[1], in[2], op.attrs.s("kind", "nf4"), op.attrs.i("blocksize", 64), wshape, in[0].scalar_type());
  return {at::matmul(in[0], w)};
}
HB_REGISTER_OP(matmul4bit, "matmul4bit", 1, 0, matmul4bit_compute, matmul4bit_grad, nullptr, nullptr);
HB_REGISTER_OP(matmul4bit_dgrad, "matmul4bit_dgrad", 1, kFlagNondiff, matmul4bit_dgrad_compute, nullptr, nullptr, nullptr);
}  // namespace hb

// ------------------------------------------------------------------ integer id arithmetic, sparse matmul (CTR / GNN / compression)
namespace hb {
ATEN_OP_NODIFF(remainder, 1, return {at::remainder(in[0], (int64_t)op.attrs.i("divisor"))};);
ATEN_OP_NODIFF(floor_divide, 1, return {at::floor_divide(in[0], (int64_t)op.attrs.i("divisor"))};);
// universal hashing of ids: ((a * id + b) mod p) mod m   (p a Mersenne prime, computed in int64)
ATEN_OP_NODIFF(hash_ids, 1, {
  const int64_t a = op.attrs.i("a", 1000003), b = op.attrs.i("b", 12345), p = op.attrs.i("p", 2147483647), m = op.attrs.i("buckets");
  at::Tensor x = in[0].to(at::kLong);
  return {at::remainder(at::remainder(x * a + b, p), m)};
});
// y = A x with A given in COO form (indices [2, nnz], values [nnz]) -- differentiable w.r.t. values and x (generic VJP)
ATEN_OP(spmm, 1, {
  const int64_t rows = op.attrs.i("rows");
  if (in[2].is_meta()) return {at::empty({rows, in[2].size(1)}, in[2].options())};
  at::Tensor A = at::sparse_coo_tensor(in[0].to(at::kLong), in[1].to(in[2].scalar_type()), {rows, in[2].size(0)});
  return {at::_sparse_mm(A, in[2])};
});
}  // namespace hb

namespace hb {
// value passes through, gradient does not (straight-through estimators, frozen branches)
ATEN_OP_NODIFF(stop_gradient, 1, return {in[0]};);
}  // namespace hb

namespace hb {
// elementwise comparison against a scalar -> 0/1 mask in `dtype` (float32 by default)
ATEN_OP_NODIFF(compare_scalar, 1, {
  const std::string m = op.attrs.s("mode", "lt");
  const double v = op.attrs.f("value");
  at::Tensor r = m == "lt" ? in[0] < v : m == "le" ? in[0] <= v : m == "gt" ? in[0] > v : m == "ge" ? in[0] >= v : m == "eq" ? in[0] == v : in[0] != v;
  return {r.to(to_aten_dtype(dtype_from_name(op.attrs.s("dtype", "float32"))))};
});
}  // namespace hb
