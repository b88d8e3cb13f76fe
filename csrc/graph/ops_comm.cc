// Communication ops: the abstract `comm` op (declares a target layout; lowered by
// the executor), explicit collectives with rank lists, vocab-parallel cross entropy
// and context-parallel (ring) attention.
// (capability parity: hetu/graph/ops/Communication.{h,cc}, VocabParallelCrossEntropyLoss.cc,
//  ParallelAttention.{h,cc})
#include <sstream>
#include <fstream>
#include <chrono>
#include <ATen/ATen.h>

#include "exec.h"
#include "ir.h"
#include "op_utils.h"

namespace hb {

using Ts = std::vector<at::Tensor>;

// ------------------------------------------------------------------ abstract comm
static void comm_infer(OpDef& op) {
  const Tensor& x = op.inputs[0];
  make_out(op, 0, x->shape, x->dtype);
  const int s = op.graph ? op.graph->cur_strategy() : 0;
  if (x->has_ds(s) && op.dst_ds.size() > (size_t)s && op.dst_ds.get(s).size() > 0) {
    const auto g = x->ds(s).global_shape(x->shape);
    op.outputs[0]->shape = op.dst_ds.get(s).get(0).local_shape(g);
  }
  if (!x->symbolic_shape.empty()) op.outputs[0]->symbolic_shape.clear();
}
static void comm_deduce(OpDef& op, size_t s) {
  auto& out = op.outputs[0];
  while (out->ds_hierarchy.size() <= s) out->ds_hierarchy.add(DistributedStatesUnion());
  if (op.dst_ds.size() > s) out->ds_hierarchy.get_mut(s) = op.dst_ds.get(s);
  else if (op.dst_ds.size() == 1) out->ds_hierarchy.get_mut(s) = op.dst_ds.get(0);
}
static Ts comm_compute(const OpDef&, const Ts& in, RunCtx*) {
  // single-process / eager graphs: a layout change is the identity
  return {in[0]};
}
static TensorList comm_grad(OpDef& op, const TensorList& g) {
  // the gradient returns to the layout of the forward input, with "partial" read as "duplicate"
  // (every contributor of a partial sum receives the same gradient)
  const Tensor& x = op.inputs[0];
  DistributedStatesHierarchy target;
  // a parameter moved to another device group (tied weights across pipeline stages): its gradient comes home in the layout it
  // has -- still a partial sum over the data-parallel replicas -- and is reduced once, by the owner's deferred gradient sync
  bool param_in = x->producer && x->producer->has_flag(kFlagVariable) && g[0] && g[0]->ds_hierarchy.size() == x->ds_hierarchy.size();
  // (only when both ends use the same number of devices: stages with different tensor-parallel degrees need the re-sharding
  // to the owner's layout below)
  for (size_t s = 0; param_in && s < x->ds_hierarchy.size(); ++s)
    if (g[0]->ds_hierarchy.get(s).get(0).device_num() != x->ds_hierarchy.get(s).get(0).device_num()) param_in = false;
  if (param_in) target = g[0]->ds_hierarchy;
  for (size_t s = 0; !param_in && s < x->ds_hierarchy.size(); ++s) {
    DistributedStatesUnion u;
    const auto& src_u = x->ds_hierarchy.get(s);
    for (size_t i = 0; i < src_u.size(); ++i) {
      const DistributedStates& ds = src_u.get(i);
      if (ds.get_dim(kPartialDim) > 1) {
        auto st = ds.combine_states({kPartialDim}, kDupDim);
        auto order = ds.combine_order({kPartialDim}, kDupDim);
        u.add(DistributedStates(ds.device_num(), st, order));
      } else u.add(ds);
    }
    u.set_hetero_dim(src_u.hetero_dim() == kPartialDim ? kDupDim : src_u.hetero_dim());
    target.add(u);
  }
  OpMeta m;
  if (x->producer) m.dg_hierarchy = x->producer->meta.dg_hierarchy;  // gradient lives where the input lived
  Tensor gx = op.graph->make_op1("comm", {g[0]}, {}, m, [&](OpDef& o) { o.dst_ds = target; });
  return {gx};
}
HB_REGISTER_OP(comm, "comm", 1, kFlagComm | kFlagNoMetaExec, comm_compute, comm_grad, comm_deduce, comm_infer);

// ------------------------------------------------------------------ heterogeneous data parallelism (Malleus / Ampelos unions)
// grouped_all_reduce: the gradient of a parameter that other pipelines shard with a DIFFERENT tensor-parallel degree is
// synchronised slice by slice -- slice i (offset, length along `dim`) is all-reduced over group i (the holders of that finest
// shard in every pipeline); replicated parameters are all-reduced among one leader per pipeline and broadcast inside the
// pipeline's tensor-parallel group.  (ref: SplitAllReduce / hetero CommOp lowering in hetu/graph/ops/Communication.cc)
static Ts grouped_all_reduce_compute(const OpDef& op, const Ts& in, RunCtx*) {
  at::Tensor g = in[0];
  if (g.is_meta()) return {at::empty_like(g)};
  auto& comm = CommRuntime::get();
  if (!comm.initialized()) return {g};
  const int64_t dim = op.attrs.i("dim", 0);
  const std::vector<int64_t> offs = op.attrs.ints("offsets"), lens = op.attrs.ints("lengths"), sizes = op.attrs.ints("group_sizes"),
                             flat = op.attrs.ints("ranks_flat"), bcast = op.attrs.ints("bcast_ranks");
  at::Tensor out = g.contiguous().clone();
  size_t pos = 0;
  for (size_t i = 0; i < sizes.size(); ++i) {
    std::vector<int> ranks(flat.begin() + pos, flat.begin() + pos + sizes[i]);
    pos += sizes[i];
    if (ranks.size() < 2) continue;
    at::Tensor piece = out.narrow(dim, offs[i], lens[i]);
    at::Tensor red = comm.all_reduce(piece.contiguous(), ranks, ReductionType::SUM);
    piece.copy_(red);
  }
  if (bcast.size() > 1) {
    std::vector<int> ranks(bcast.begin(), bcast.end());
    out = comm.broadcast(out, ranks, ranks[0]);
  }
  return {out};
}
HB_REGISTER_OP(grouped_all_reduce, "grouped_all_reduce", 1, kFlagComm | kFlagNondiff, grouped_all_reduce_compute, nullptr,
               [](OpDef& op, size_t s) { copy_out_ds(op, 0, s, op.inputs[0]); }, nullptr);

// grouped_reduce_scatter / grouped_all_gather: the sharded-state (ZeRO) counterparts of grouped_all_reduce.  Slice i of the
// gradient is reduce-scattered over group i along `dim` (every member keeps 1 / |group| of the reduced slice, in group
// order); the kept parts are concatenated.  grouped_all_gather is the inverse: the parts are all-gathered slice by slice and
// written back at the slice offsets.  (ref: SplitReduceScatterOp / SplitAllGatherOp, hetu/graph/ops/Communication.cc:660-852)
static std::vector<std::vector<int>> slice_groups(const OpDef& op) {
  const auto sizes = op.attrs.ints("group_sizes"), flat = op.attrs.ints("ranks_flat");
  std::vector<std::vector<int>> out;
  size_t pos = 0;
  for (int64_t n : sizes) {
    out.emplace_back(flat.begin() + pos, flat.begin() + pos + n);
    pos += (size_t)n;
  }
  return out;
}
static Ts grouped_reduce_scatter_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const at::Tensor& g = in[0];
  const int64_t dim = op.attrs.i("dim", 0);
  const auto offs = op.attrs.ints("offsets"), lens = op.attrs.ints("lengths");
  auto groups = slice_groups(op);
  int64_t kept = 0;
  for (size_t i = 0; i < lens.size(); ++i) kept += lens[i] / std::max<int64_t>((int64_t)groups[i].size(), 1);
  auto shp = g.sizes().vec();
  shp[dim] = kept;
  if (g.is_meta()) return {at::empty(shp, g.options())};
  auto& comm = CommRuntime::get();
  std::vector<at::Tensor> parts;
  for (size_t i = 0; i < lens.size(); ++i) {
    at::Tensor piece = g.narrow(dim, offs[i], lens[i]).contiguous();
    if (groups[i].size() > 1 && comm.initialized()) {
      HB_CHECK(lens[i] % (int64_t)groups[i].size() == 0) << "slice of " << lens[i] << " rows is not divisible over " << groups[i].size() << " holders";
      piece = comm.reduce_scatter(piece, groups[i], (int)dim, ReductionType::SUM, op.attrs.b("fp32_reduce"));
    }
    parts.push_back(piece);
  }
  return {parts.empty() ? at::empty(shp, g.options()) : at::cat(parts, dim)};
}
static Ts grouped_all_gather_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const at::Tensor& x = in[0];
  const int64_t dim = op.attrs.i("dim", 0);
  const auto offs = op.attrs.ints("offsets"), lens = op.attrs.ints("lengths");
  auto groups = slice_groups(op);
  int64_t full = 0;
  for (size_t i = 0; i < lens.size(); ++i) full = std::max(full, offs[i] + lens[i]);
  auto shp = x.sizes().vec();
  shp[dim] = full;
  if (x.is_meta()) return {at::empty(shp, x.options())};
  auto& comm = CommRuntime::get();
  at::Tensor out = at::zeros(shp, x.options());
  int64_t pos = 0;
  for (size_t i = 0; i < lens.size(); ++i) {
    const int64_t n = std::max<int64_t>((int64_t)groups[i].size(), 1), part = lens[i] / n;
    at::Tensor piece = x.narrow(dim, pos, part).contiguous();
    pos += part;
    if (n > 1 && comm.initialized()) piece = comm.all_gather(piece, groups[i], (int)dim);
    out.narrow(dim, offs[i], lens[i]).copy_(piece);
  }
  return {out};
}
HB_REGISTER_OP(grouped_reduce_scatter, "grouped_reduce_scatter", 1, kFlagComm | kFlagNondiff, grouped_reduce_scatter_compute, nullptr, nullptr, nullptr);
HB_REGISTER_OP(grouped_all_gather, "grouped_all_gather", 1, kFlagComm | kFlagNondiff, grouped_all_gather_compute, nullptr, nullptr, nullptr);

// ------------------------------------------------------------------ explicit collectives (rank lists in attrs)
static std::vector<int> ranks_attr(const OpDef& op) {
  std::vector<int> r;
  for (auto v : op.attrs.ints("ranks")) r.push_back((int)v);
  return r;
}
static bool single(const OpDef& op) { return !CommRuntime::get().initialized() || op.attrs.ints("ranks").size() <= 1; }

static Ts all_reduce_compute(const OpDef& op, const Ts& in, RunCtx*) {
  if (in[0].is_meta() || single(op)) return {in[0]};
  return {CommRuntime::get().all_reduce(in[0], ranks_attr(op), reduction_from_name(op.attrs.s("reduction", "sum")),
                                        op.attrs.b("fp32_reduce"))};
}
static TensorList all_reduce_grad(OpDef& op, const TensorList& g) { return {g[0]}; }
HB_REGISTER_OP(all_reduce, "all_reduce", 1, kFlagComm, all_reduce_compute, all_reduce_grad, nullptr, nullptr);

static Ts all_gather_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const int64_t n = (int64_t)op.attrs.ints("ranks").size();
  const int64_t dim = op.attrs.i("dim", 0);
  if (in[0].is_meta()) {
    auto shp = in[0].sizes().vec();
    shp[dim] *= std::max<int64_t>(n, 1);
    return {at::empty(shp, in[0].options())};
  }
  if (single(op)) return {in[0]};
  return {CommRuntime::get().all_gather(in[0], ranks_attr(op), (int)dim)};
}
static Ts reduce_scatter_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const int64_t n = (int64_t)op.attrs.ints("ranks").size();
  const int64_t dim = op.attrs.i("dim", 0);
  if (in[0].is_meta()) {
    auto shp = in[0].sizes().vec();
    shp[dim] /= std::max<int64_t>(n, 1);
    return {at::empty(shp, in[0].options())};
  }
  if (single(op)) return {in[0]};
  return {CommRuntime::get().reduce_scatter(in[0], ranks_attr(op), (int)dim, ReductionType::SUM, op.attrs.b("fp32_reduce"))};
}
static TensorList all_gather_grad(OpDef& op, const TensorList& g) {
  return {op.graph->make_op1("reduce_scatter", {g[0]}, op.attrs)};
}
static TensorList reduce_scatter_grad(OpDef& op, const TensorList& g) {
  return {op.graph->make_op1("all_gather", {g[0]}, op.attrs)};
}
HB_REGISTER_OP(all_gather, "all_gather", 1, kFlagComm, all_gather_compute, all_gather_grad, nullptr, nullptr);
HB_REGISTER_OP(reduce_scatter, "reduce_scatter", 1, kFlagComm, reduce_scatter_compute, reduce_scatter_grad, nullptr, nullptr);

// all_to_all over `ranks`: split dim `split_dim` into n chunks, chunk j goes to rank j, received chunks are
// concatenated along `concat_dim` (MoE dispatch / combine, Ulysses-style head<->sequence exchange)
static Ts all_to_all_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const int64_t n = std::max<int64_t>((int64_t)op.attrs.ints("ranks").size(), 1);
  const int64_t sd = op.attrs.i("split_dim", 0), cd = op.attrs.i("concat_dim", 0);
  if (in[0].is_meta()) {
    auto shp = in[0].sizes().vec();
    shp[sd] /= n;
    shp[cd] *= n;
    return {at::empty(shp, in[0].options())};
  }
  if (single(op)) return {in[0]};
  return {CommRuntime::get().all_to_all(in[0], ranks_attr(op), (int)sd, (int)cd)};
}
static TensorList all_to_all_grad(OpDef& op, const TensorList& g) {
  AttrMap a = op.attrs;
  a.set("split_dim", op.attrs.i("concat_dim", 0));
  a.set("concat_dim", op.attrs.i("split_dim", 0));
  return {op.graph->make_op1("all_to_all", {g[0]}, a)};
}
HB_REGISTER_OP(all_to_all, "all_to_all", 1, kFlagComm, all_to_all_compute, all_to_all_grad, nullptr, nullptr);

// hierarchical (two-level) all-to-all: ranks = nodes x gpus_per_node (node-major).  Stage 1 exchanges inside the node so
// that local peer l holds everything this node sends to the GPUs with local index l; stage 2 exchanges between the nodes
// among GPUs of equal local index -- every inter-node message is gpus_per_node times larger than in the flat algorithm
// (one aggregated message per node pair and rail instead of one per GPU pair).  Between the stages the per-destination
// chunks are regrouped by a layout-transform kernel.  Same result as all_to_all(split_dim=0, concat_dim=0).
// (ref: hetu/v1/src/communication/mpi_nccl_communication.cu:152-243 _ncclHAllToAll / HA2AGather / HA2AScatter,
//  hetu/v1/src/ops/H_A2A_LayoutTransform.cu, python halltoall_op)
static at::Tensor chunk_transpose_any(const at::Tensor& x, int64_t a, int64_t b) {
  const int64_t chunk_bytes = (int64_t)x.nbytes() / (a * b);
  if (x.is_cuda() && x.is_contiguous() && chunk_bytes % 16 == 0) {
    at::Tensor y = at::empty_like(x);
    cuda_ok(chunk_transpose(x.data_ptr(), y.data_ptr(), (int)a, (int)b, chunk_bytes, cur_stream()), "chunk_transpose");
    return y;
  }
  auto shp = x.sizes().vec();
  at::Tensor v = x.reshape({a, b, -1}).transpose(0, 1).contiguous();
  return v.reshape(shp);
}
static Ts hall_to_all_compute(const OpDef& op, const Ts& in, RunCtx*) {
  if (in[0].is_meta() || single(op)) return {in[0]};
  const std::vector<int> ranks = ranks_attr(op);
  const int64_t n = (int64_t)ranks.size();
  int64_t local = op.attrs.i("gpus_per_node", n);
  if (local <= 0 || n % local != 0) local = n;
  const int64_t nodes = n / local;
  auto& comm = CommRuntime::get();
  if (nodes == 1 || local == 1) return {comm.all_to_all(in[0], ranks, 0, 0)};
  HB_CHECK(in[0].size(0) % n == 0) << "hall_to_all: dim 0 (" << in[0].size(0) << ") must be divisible by " << n << " ranks";
  int64_t me = -1;
  for (int64_t i = 0; i < n; ++i) if (ranks[i] == comm.rank()) me = i;
  HB_CHECK(me >= 0) << "hall_to_all: rank " << comm.rank() << " is not a member of the group";
  const int64_t my_node = me / local, my_l = me % local;
  // sub-groups are created collectively: every rank requests all node groups and all rail groups in the same order
  std::vector<int> node_group, rail_group;
  for (int64_t q = 0; q < nodes; ++q) {
    std::vector<int> grp;
    for (int64_t l = 0; l < local; ++l) grp.push_back(ranks[q * local + l]);
    comm.group(grp);
    if (q == my_node) node_group = grp;
  }
  for (int64_t l = 0; l < local; ++l) {
    std::vector<int> grp;
    for (int64_t q = 0; q < nodes; ++q) grp.push_back(ranks[q * local + l]);
    comm.group(grp);
    if (l == my_l) rail_group = grp;
  }
  at::Tensor x = in[0].contiguous();
  // [node q][local l] -> [l][q]; intra-node exchange by l
  at::Tensor s1 = comm.all_to_all(chunk_transpose_any(x, nodes, local), node_group, 0, 0);      // now [from p][to node q]
  // [p][q] -> [q][p]; inter-node exchange by q
  at::Tensor s2 = comm.all_to_all(chunk_transpose_any(s1, local, nodes), rail_group, 0, 0);     // now [from node q'][from local p]
  return {s2};
}
static TensorList hall_to_all_grad(OpDef& op, const TensorList& g) {
  return {op.graph->make_op1("hall_to_all", {g[0]}, op.attrs)};      // the exchange is its own transpose
}
HB_REGISTER_OP(hall_to_all, "hall_to_all", 1, kFlagComm, hall_to_all_compute, hall_to_all_grad, nullptr, nullptr);

static Ts broadcast_comm_compute(const OpDef& op, const Ts& in, RunCtx*) {
  if (in[0].is_meta() || single(op)) return {in[0]};
  return {CommRuntime::get().broadcast(in[0], ranks_attr(op), (int)op.attrs.i("root"))};
}
HB_REGISTER_OP(broadcast_comm, "broadcast_comm", 1, kFlagComm | kFlagNondiff, broadcast_comm_compute, nullptr, nullptr, nullptr);

// ------------------------------------------------------------------ explicit pipeline hand-off (v1 API: pipeline_send_op / pipeline_receive_op)
// pipeline_send(x; dst, channel) ships x to rank dst and yields a 1-element token (so the node can be fetched / depended on);
// pipeline_recv(; src, shape, dtype, channel) yields the tensor sent by rank src.  The DS-lowered pipelines never emit these --
// their stage boundaries are `comm` ops the executor turns into batched P2P -- they serve hand-placed v1 graphs.
static void pipeline_send_infer(OpDef& op) { make_out(op, 0, {1}, op.inputs[0]->dtype); }
static Ts pipeline_send_compute(const OpDef& op, const Ts& in, RunCtx*) {
  if (in[0].is_meta()) return {at::empty({1}, in[0].options())};
  auto& comm = CommRuntime::get();
  if (comm.initialized() && (int)op.attrs.i("dst") != comm.rank()) {
    comm.send(in[0].contiguous(), (int)op.attrs.i("dst"), (int)op.attrs.i("channel", 0));
    comm.flush_sends();
  }
  return {at::zeros({1}, in[0].options())};
}
HB_REGISTER_OP(pipeline_send, "pipeline_send", 1, kFlagComm | kFlagNondiff | kFlagNoMetaExec, pipeline_send_compute, nullptr, nullptr,
               pipeline_send_infer);

static void pipeline_recv_infer(OpDef& op) { make_out(op, 0, op.attrs.ints("shape"), dtype_from_name(op.attrs.s("dtype", "float32"))); }
static Ts pipeline_recv_compute(const OpDef& op, const Ts&, RunCtx* rc) {
  auto& comm = CommRuntime::get();
  const auto dt = to_aten_dtype(dtype_from_name(op.attrs.s("dtype", "float32")));
  (void)rc;
  const at::Device dev = aten_device();
  HB_CHECK(comm.initialized()) << "pipeline_recv needs an initialised communication runtime";
  return {comm.recv(op.attrs.ints("shape"), dt, dev, (int)op.attrs.i("src"), (int)op.attrs.i("channel", 0))};
}
HB_REGISTER_OP(pipeline_recv, "pipeline_recv", 1, kFlagComm | kFlagNondiff | kFlagNoMetaExec, pipeline_recv_compute, nullptr, nullptr,
               pipeline_recv_infer);

// ------------------------------------------------------------------ vocab-parallel cross entropy
// logits [T, V/t] sharded on the vocab dim, labels hold *global* vocabulary ids.
// Three small all-reduces (max, sum-exp, target logit) between local kernels.
static std::vector<int> tp_ranks(const OpDef& op, RunCtx* rc) {
  auto explicit_ranks = op.attrs.ints("ranks");
  std::vector<int> r;
  if (!explicit_ranks.empty()) {
    for (auto v : explicit_ranks) r.push_back((int)v);
    return r;
  }
  // derive from the logits layout: peers along the vocab (last) dim
  const Tensor& lg = op.inputs[0];
  const int s = rc ? rc->strategy : 0;
  if (!lg->has_ds(s) || !rc || !rc->exec) return {};
  const DistributedStates& ds = lg->ds(s);
  const DeviceGroup grp = op.placement(s);
  if (grp.empty()) return {};
  const int me = rc->exec->local_device_index(grp);
  if (me < 0) return {};
  for (int i : ds.get_device_indices_by_dim(lg->ndim() - 1, me)) r.push_back(grp.get(i).index());
  return r;
}
static Ts vp_ce_compute(const OpDef& op, const Ts& in, RunCtx* rc) {
  const at::Tensor& logits = in[0];
  const at::Tensor& labels = in[1];
  const int64_t ignore = op.attrs.i("ignore_index", -1);
  const std::string red = op.attrs.s("reduction", "mean");
  const int64_t Vl = logits.size(-1);
  std::vector<int64_t> lshape(logits.sizes().begin(), logits.sizes().end() - 1);
  auto fopt = logits.options().dtype(at::kFloat);
  if (logits.is_meta()) return {red == "none" ? at::empty(lshape, fopt) : at::empty({}, fopt), at::empty_like(logits)};
  std::vector<int> ranks = tp_ranks(op, rc);
  auto& comm = CommRuntime::get();
  int my_pos = 0;
  for (size_t i = 0; i < ranks.size(); ++i) if (ranks[i] == comm.rank()) my_pos = (int)i;
  const bool dist = comm.initialized() && ranks.size() > 1;
  const int64_t vstart = op.attrs.has("vocab_start") ? op.attrs.i("vocab_start") : (int64_t)my_pos * Vl;
  at::Tensor lab = labels.to(at::kLong).reshape({-1}).contiguous();
  const int64_t rows = lab.numel();
  at::Tensor per_tok = at::empty({rows}, fopt), unit;
  if (is_native(logits) && logits.is_contiguous() && Vl % 8 == 0) {
    unit = op.attrs.b("donate_logits") ? logits : logits.clone();
    at::Tensor rmax = at::empty({rows}, fopt), sexp = at::empty({rows}, fopt), tgt = at::empty({rows}, fopt);
    cuda_ok(vp_ce_local_max(unit.data_ptr(), rmax.data_ptr<float>(), rows, (int)Vl, Vl, cur_stream()), "vp_ce_max");
    if (dist) rmax = comm.all_reduce(rmax, ranks, ReductionType::MAX);
    cuda_ok(vp_ce_local_sum(unit.data_ptr(), lab.data_ptr<int64_t>(), rmax.data_ptr<float>(), sexp.data_ptr<float>(),
                            tgt.data_ptr<float>(), rows, (int)Vl, Vl, vstart, cur_stream()), "vp_ce_sum");
    if (dist) {
      at::Tensor both = at::stack({sexp, tgt});
      both = comm.all_reduce(both, ranks, ReductionType::SUM);
      sexp = both[0].contiguous();
      tgt = both[1].contiguous();
    }
    cuda_ok(vp_ce_finish(unit.data_ptr(), lab.data_ptr<int64_t>(), rmax.data_ptr<float>(), sexp.data_ptr<float>(),
                         tgt.data_ptr<float>(), per_tok.data_ptr<float>(), rows, (int)Vl, Vl, vstart, ignore, 1.0f, true,
                         cur_stream()), "vp_ce_finish");
  } else {
    at::Tensor lf = logits.to(at::kFloat).reshape({rows, Vl});
    at::Tensor rmax = std::get<0>(lf.max(-1));
    if (dist) rmax = comm.all_reduce(rmax, ranks, ReductionType::MAX);
    at::Tensor ex = at::exp(lf - rmax.unsqueeze(1));
    at::Tensor sexp = ex.sum(-1);
    at::Tensor local = lab - vstart;
    at::Tensor mine = (local >= 0).logical_and(local < Vl);
    at::Tensor safe = at::where(mine, local, at::zeros_like(local));
    at::Tensor tgt = lf.gather(1, safe.unsqueeze(1)).squeeze(1) * mine.to(at::kFloat);
    if (dist) {
      at::Tensor both = comm.all_reduce(at::stack({sexp, tgt}), ranks, ReductionType::SUM);
      sexp = both[0];
      tgt = both[1];
    }
    at::Tensor valid = (lab != ignore).to(at::kFloat);
    at::Tensor lse = rmax + at::log(sexp);
    per_tok = (lse - tgt) * valid;
    at::Tensor u = ex / sexp.unsqueeze(1);
    u.scatter_add_(1, safe.unsqueeze(1), -mine.to(at::kFloat).unsqueeze(1));
    unit = (u * valid.unsqueeze(1)).to(logits.scalar_type()).reshape(logits.sizes());
  }
  at::Tensor loss;
  if (red == "none") loss = per_tok.reshape(lshape);
  else if (red == "sum") loss = per_tok.sum();
  else loss = per_tok.sum() / (lab != ignore).sum().to(at::kFloat).clamp_min(1.0);
  return {loss, unit.reshape(logits.sizes())};
}
static TensorList vp_ce_grad(OpDef& op, const TensorList& g) {
  AttrMap a;
  a.set("reduction", op.attrs.s("reduction", "mean"));
  a.set("ignore_index", op.attrs.i("ignore_index", -1));
  return {op.graph->make_op1("softmax_ce_sparse_bwd", {g[0], op.outputs[1], op.inputs[1]}, a), nullptr};
}
static void vp_ce_deduce(OpDef& op, size_t s) {
  const Tensor& lg = op.inputs[0];
  const Tensor& lab = op.inputs[1];
  if (!lg->has_ds(s)) return;
  copy_out_ds(op, 1, s, lg);
  const std::string red = op.attrs.s("reduction", "mean");
  if (red == "none" && lab->has_ds(s)) { copy_out_ds(op, 0, s, lab); return; }
  const int n = lg->ds(s).device_num();
  set_out_ds(op, 0, s, DistributedStates(n, {{kDupDim, n}}, {kDupDim}));
}
HB_REGISTER_OP(vocab_parallel_cross_entropy, "vocab_parallel_cross_entropy", 2, 0, vp_ce_compute, vp_ce_grad, vp_ce_deduce,
               nullptr);

// ------------------------------------------------------------------ context-parallel attention
// q, k, v hold this rank's sequence chunk(s); KV blocks travel around the CP ring with batched
// send/recv while the local flash-attention kernel runs, and partial results are merged with their
// log-sum-exp.  Causal load balance: SYM (zig-zag) split -- rank i owns chunks i and 2c-1-i.
// Backward re-walks the ring, circulating dK/dV accumulators with the KV blocks.
static std::vector<int> cp_ranks(const OpDef& op) {
  std::vector<int> r;
  for (auto v : op.attrs.ints("ranks")) r.push_back((int)v);
  return r;
}
struct AttnPiece { at::Tensor o, lse; };
static AttnPiece local_attn(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, double scale, bool causal) {
  static const OpKernel* kern = OpRegistry::get().find("attn");
  OpDef tmp;
  tmp.attrs.set("causal", causal);
  tmp.attrs.set("softmax_scale", scale);
  tmp.kernel = kern;
  auto r = kern->compute(tmp, {q, k, v}, nullptr);
  return {r[0], r[1]};
}
// merge two normalised partial attention results
static void merge_piece(at::Tensor& o, at::Tensor& lse, const AttnPiece& p) {
  if (!o.defined()) { o = p.o.to(at::kFloat); lse = p.lse.clone(); return; }
  at::Tensor nl = at::logaddexp(lse, p.lse);
  at::Tensor w0 = at::exp(lse - nl).permute({0, 2, 1}).unsqueeze(-1);   // [B,S,H,1]
  at::Tensor w1 = at::exp(p.lse - nl).permute({0, 2, 1}).unsqueeze(-1);
  w0 = at::where(at::isfinite(w0), w0, at::zeros_like(w0));
  w1 = at::where(at::isfinite(w1), w1, at::zeros_like(w1));
  o = o * w0 + p.o.to(at::kFloat) * w1;
  lse = nl;
}
// ---- variable-length (packed) rows under context parallelism: `cu` holds the document boundaries in the coordinates of
// the whole row (length cs * number of chunks).  A (query chunk, key chunk) block decomposes into one sub-block per
// document that intersects both chunks: the same document inside the diagonal block is causal, a document spanning an
// earlier key chunk is fully visible, everything else is masked (ref: AttnInfo / valid_cu_seqlens, ParallelAttention.cc:125-330)
struct DocPiece { int64_t q_off, q_len, k_off, k_len; bool causal; };
static std::vector<int64_t> host_boundaries(const at::Tensor& cu, int64_t row_len) {
  at::Tensor h = cu.to(at::kCPU, at::kLong).contiguous();
  std::vector<int64_t> v(h.data_ptr<int64_t>(), h.data_ptr<int64_t>() + h.numel());
  std::vector<int64_t> out;
  for (int64_t b : v) {
    b = std::min(b, row_len);
    if (out.empty() || b > out.back()) out.push_back(b);
  }
  if (out.empty() || out.front() != 0) out.insert(out.begin(), 0);
  if (out.back() != row_len) out.push_back(row_len);
  return out;
}
static std::vector<DocPiece> doc_pieces(const std::vector<int64_t>& cu, int qpos, int kpos, int64_t cs, bool causal) {
  std::vector<DocPiece> out;
  const int64_t q_lo = qpos * cs, q_hi = q_lo + cs, k_lo = kpos * cs, k_hi = k_lo + cs;
  for (size_t d = 0; d + 1 < cu.size(); ++d) {
    const int64_t qa = std::max(q_lo, cu[d]), qb = std::min(q_hi, cu[d + 1]);
    const int64_t ka = std::max(k_lo, cu[d]), kb = std::min(k_hi, cu[d + 1]);
    if (qa >= qb || ka >= kb) continue;
    out.push_back({qa - q_lo, qb - qa, ka - k_lo, kb - ka, causal && qpos == kpos});
  }
  return out;
}
// merge a partial result into rows [off, off + len) of the accumulators (o [B,cs,H,D] fp32, lse [B,H,cs], -inf = empty)
static void merge_rows(at::Tensor& o, at::Tensor& lse, int64_t off, int64_t len, const AttnPiece& p) {
  at::Tensor osub = o.narrow(1, off, len), lsub = lse.narrow(2, off, len);
  at::Tensor nl = at::logaddexp(lsub, p.lse);
  at::Tensor w0 = at::exp(lsub - nl).permute({0, 2, 1}).unsqueeze(-1);
  at::Tensor w1 = at::exp(p.lse - nl).permute({0, 2, 1}).unsqueeze(-1);
  w0 = at::where(at::isfinite(w0), w0, at::zeros_like(w0));
  w1 = at::where(at::isfinite(w1), w1, at::zeros_like(w1));
  osub.copy_(osub * w0 + p.o.to(at::kFloat) * w1);
  lsub.copy_(nl);
}
// position of this rank's chunks: SYM -> two half-chunks (i, 2c-1-i); NORMAL -> one chunk i
static std::vector<int> owned_chunks(int idx, int c, bool sym) {
  if (sym) return {idx, 2 * c - 1 - idx};
  return {idx};
}
static Ts parallel_attn_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const at::Tensor& q = in[0];
  const at::Tensor& k = in[1];
  const at::Tensor& v = in[2];
  auto fopt = q.options().dtype(at::kFloat);
  if (q.is_meta()) return {at::empty(q.sizes(), q.options()), at::empty({q.size(0), q.size(2), q.size(1)}, fopt)};
  const bool causal = op.attrs.b("causal", true);
  const double scale = op.attrs.f("softmax_scale", 0.0) > 0 ? op.attrs.f("softmax_scale") : 1.0 / std::sqrt((double)q.size(3));
  std::vector<int> ranks = cp_ranks(op);
  auto& comm = CommRuntime::get();
  const int c = (int)ranks.size();
  const bool varlen = in.size() > 3 && in[3].defined();
  if (c <= 1 || !comm.initialized()) {
    if (!varlen) {
      AttnPiece p = local_attn(q, k, v, scale, causal);
      return {p.o, p.lse};
    }
    const std::vector<int64_t> cu1 = host_boundaries(in[3], q.size(1));
    at::Tensor o1 = at::zeros(q.sizes(), fopt), l1 = at::full({q.size(0), q.size(2), q.size(1)}, -INFINITY, fopt);
    for (auto& pc : doc_pieces(cu1, 0, 0, q.size(1), causal))
      merge_rows(o1, l1, pc.q_off, pc.q_len, local_attn(q.narrow(1, pc.q_off, pc.q_len), k.narrow(1, pc.k_off, pc.k_len),
                                                        v.narrow(1, pc.k_off, pc.k_len), scale, pc.causal));
    return {o1.to(q.scalar_type()), l1};
  }
  int idx = 0;
  for (int i = 0; i < c; ++i) if (ranks[i] == comm.rank()) idx = i;
  const bool sym = causal && op.attrs.s("split_pattern", env_str("HETU_PARALLEL_ATTN_SPLIT_PATTERN", "SYM")) == "SYM";
  const int64_t S = q.size(1);
  const int nchunk = sym ? 2 : 1;
  HB_CHECK(S % nchunk == 0) << "sequence chunk not divisible for the SYM split";
  const int64_t cs = S / nchunk;
  auto my_chunks = owned_chunks(idx, c, sym);
  std::vector<int64_t> cu;
  if (varlen) cu = host_boundaries(in[3], cs * nchunk * c);
  std::vector<at::Tensor> o_acc(nchunk), lse_acc(nchunk);
  if (varlen)
    for (int qi = 0; qi < nchunk; ++qi) {
      o_acc[qi] = at::zeros({q.size(0), cs, q.size(2), q.size(3)}, fopt);
      lse_acc[qi] = at::full({q.size(0), q.size(2), cs}, -INFINITY, fopt);
    }
  at::Tensor kv_cur = at::stack({k, v}).contiguous();
  const int next = ranks[(idx + 1) % c], prev = ranks[(idx - 1 + c) % c];
  // HETU_PARALLEL_ATTN=ANALYSIS: per ring round, attention time, blocks computed / skipped by the causal mask and the KV
  // bytes rotated (synchronising per round, so only for analysis); appended to HETU_PARALLEL_ATTN_LOG_FILE.rank<r>
  // (ref: AttnCommRing::Profile, ops/ParallelAttention.cc:1140)
  const bool analysis = env_str("HETU_PARALLEL_ATTN", "") == "ANALYSIS";
  std::ostringstream report;
  auto sync_now = [&]() {
    if (q.is_cuda()) at::cuda::getCurrentCUDAStream().synchronize();
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  };
  for (int round = 0; round < c; ++round) {
    const double t_round = analysis ? sync_now() : 0.0;
    int blocks = 0, skipped = 0;
    at::Tensor kv_next;
    const int src_idx = (idx - round + c) % c;  // owner of the KV block we hold this round
    std::vector<std::pair<at::Tensor, int>> sends, recvs;
    if (round + 1 < c) {
      kv_next = at::empty_like(kv_cur);
      sends.push_back({kv_cur, next});
      recvs.push_back({kv_next, prev});
      comm.batched_send_recv(sends, recvs);   // overlaps with the attention kernels below (async on NCCL)
    }
    auto src_chunks = owned_chunks(src_idx, c, sym);
    for (int qi = 0; qi < nchunk; ++qi) {
      at::Tensor qc = q.narrow(1, qi * cs, cs);
      for (int ki = 0; ki < nchunk; ++ki) {
        const int qpos = my_chunks[qi], kpos = src_chunks[ki];
        if (causal && kpos > qpos) { ++skipped; continue; }  // fully masked block: skipped
        at::Tensor kc = kv_cur[0].narrow(1, ki * cs, cs), vc = kv_cur[1].narrow(1, ki * cs, cs);
        if (varlen) {
          auto pieces = doc_pieces(cu, qpos, kpos, cs, causal);
          for (auto& pc : pieces)
            merge_rows(o_acc[qi], lse_acc[qi], pc.q_off, pc.q_len,
                       local_attn(qc.narrow(1, pc.q_off, pc.q_len), kc.narrow(1, pc.k_off, pc.k_len), vc.narrow(1, pc.k_off, pc.k_len), scale, pc.causal));
          if (pieces.empty()) ++skipped; else ++blocks;
          continue;
        }
        AttnPiece p = local_attn(qc, kc, vc, scale, causal && kpos == qpos);
        merge_piece(o_acc[qi], lse_acc[qi], p);
        ++blocks;
      }
    }
    if (analysis) {
      const double t_attn = sync_now() - t_round;
      if (round + 1 < c) comm.flush_sends();
      report << (round ? ", " : "") << "{\"round\": " << round << ", \"kv_from\": " << ranks[src_idx] << ", \"attn_ms\": " << t_attn
             << ", \"blocks\": " << blocks << ", \"skipped\": " << skipped << ", \"kv_bytes_sent\": "
             << (round + 1 < c ? (double)kv_cur.nbytes() : 0.0) << "}";
    }
    if (round + 1 < c) kv_cur = kv_next;
  }
  if (analysis) {
    const std::string line = "{\"op\": \"" + op.name() + "\", \"rank\": " + std::to_string(comm.rank()) + ", \"cp\": " + std::to_string(c) +
                             ", \"pattern\": \"" + (sym ? "SYM" : "NORMAL") + "\", \"rounds\": [" + report.str() + "]}";
    const std::string path = env_str("HETU_PARALLEL_ATTN_LOG_FILE", "");
    if (path.empty()) HB_LOG(INFO) << "[parallel_attn] " << line;
    else {
      std::ofstream f(path + ".rank" + std::to_string(comm.rank()), std::ios::app);
      f << line << "\n";
    }
  }
  at::Tensor o = at::cat(o_acc, 1).to(q.scalar_type());
  at::Tensor lse = at::cat(lse_acc, 2);
  return {o, lse};
}
// inputs: do, q, k, v, o, lse -> dq, dk, dv
static Ts parallel_attn_bwd_compute(const OpDef& op, const Ts& in, RunCtx*) {
  const at::Tensor& d_o = in[0];
  const at::Tensor& q = in[1];
  const at::Tensor& k = in[2];
  const at::Tensor& v = in[3];
  const at::Tensor& o = in[4];
  const at::Tensor& lse = in[5];
  if (q.is_meta()) return {at::empty(q.sizes(), q.options()), at::empty(k.sizes(), k.options()), at::empty(v.sizes(), v.options())};
  const bool causal = op.attrs.b("causal", true);
  const double scale = op.attrs.f("softmax_scale", 0.0) > 0 ? op.attrs.f("softmax_scale") : 1.0 / std::sqrt((double)q.size(3));
  static const OpKernel* bwd = OpRegistry::get().find("attn_bwd");
  auto run_bwd = [&](const at::Tensor& dO, const at::Tensor& Q, const at::Tensor& K, const at::Tensor& V, const at::Tensor& O,
                     const at::Tensor& L, bool cz) {
    OpDef tmp;
    tmp.attrs.set("causal", cz);
    tmp.attrs.set("softmax_scale", scale);
    tmp.kernel = bwd;
    return bwd->compute(tmp, {dO.contiguous(), Q.contiguous(), K.contiguous(), V.contiguous(), O.contiguous(), L.contiguous()}, nullptr);
  };
  std::vector<int> ranks = cp_ranks(op);
  auto& comm = CommRuntime::get();
  const int c = (int)ranks.size();
  const bool varlen = in.size() > 6 && in[6].defined();
  if (c <= 1 || !comm.initialized()) {
    if (!varlen) return run_bwd(d_o, q, k, v, o, lse, causal);
    const std::vector<int64_t> cu1 = host_boundaries(in[6], q.size(1));
    at::Tensor dq1 = at::zeros(q.sizes(), q.options().dtype(at::kFloat)), dk1 = at::zeros_like(k, at::kFloat), dv1 = at::zeros_like(v, at::kFloat);
    for (auto& pc : doc_pieces(cu1, 0, 0, q.size(1), causal)) {
      auto r = run_bwd(d_o.narrow(1, pc.q_off, pc.q_len), q.narrow(1, pc.q_off, pc.q_len), k.narrow(1, pc.k_off, pc.k_len),
                       v.narrow(1, pc.k_off, pc.k_len), o.narrow(1, pc.q_off, pc.q_len), lse.narrow(2, pc.q_off, pc.q_len), pc.causal);
      dq1.narrow(1, pc.q_off, pc.q_len).add_(r[0].to(at::kFloat));
      dk1.narrow(1, pc.k_off, pc.k_len).add_(r[1].to(at::kFloat));
      dv1.narrow(1, pc.k_off, pc.k_len).add_(r[2].to(at::kFloat));
    }
    return {dq1.to(q.scalar_type()), dk1.to(k.scalar_type()), dv1.to(v.scalar_type())};
  }
  int idx = 0;
  for (int i = 0; i < c; ++i) if (ranks[i] == comm.rank()) idx = i;
  const bool sym = causal && op.attrs.s("split_pattern", env_str("HETU_PARALLEL_ATTN_SPLIT_PATTERN", "SYM")) == "SYM";
  const int64_t S = q.size(1);
  const int nchunk = sym ? 2 : 1;
  const int64_t cs = S / nchunk;
  auto my_chunks = owned_chunks(idx, c, sym);
  std::vector<int64_t> cu;
  if (varlen) cu = host_boundaries(in[6], cs * nchunk * c);
  at::Tensor dq = at::zeros(q.sizes(), q.options().dtype(at::kFloat));
  // the travelling buffer carries [k, v, dk, dv] so gradients return to the owner after a full loop
  at::Tensor buf = at::stack({k.to(at::kFloat), v.to(at::kFloat), at::zeros_like(k, at::kFloat), at::zeros_like(v, at::kFloat)}).contiguous();
  const int next = ranks[(idx + 1) % c], prev = ranks[(idx - 1 + c) % c];
  for (int round = 0; round < c; ++round) {
    const int src_idx = (idx - round + c) % c;
    auto src_chunks = owned_chunks(src_idx, c, sym);
    at::Tensor kb = buf[0].to(q.scalar_type()), vb = buf[1].to(q.scalar_type());
    for (int qi = 0; qi < nchunk; ++qi) {
      for (int ki = 0; ki < nchunk; ++ki) {
        const int qpos = my_chunks[qi], kpos = src_chunks[ki];
        if (causal && kpos > qpos) continue;
        if (varlen) {
          for (auto& pc : doc_pieces(cu, qpos, kpos, cs, causal)) {
            const int64_t qo = qi * cs + pc.q_off, ko = ki * cs + pc.k_off;
            auto r = run_bwd(d_o.narrow(1, qo, pc.q_len), q.narrow(1, qo, pc.q_len), kb.narrow(1, ko, pc.k_len), vb.narrow(1, ko, pc.k_len),
                             o.narrow(1, qo, pc.q_len), lse.narrow(2, qo, pc.q_len), pc.causal);
            dq.narrow(1, qo, pc.q_len).add_(r[0].to(at::kFloat));
            buf[2].narrow(1, ko, pc.k_len).add_(r[1].to(at::kFloat));
            buf[3].narrow(1, ko, pc.k_len).add_(r[2].to(at::kFloat));
          }
          continue;
        }
        // the local kernel recomputes P from the *global* lse of the query rows, so per-block calls compose
        auto r = run_bwd(d_o.narrow(1, qi * cs, cs), q.narrow(1, qi * cs, cs), kb.narrow(1, ki * cs, cs), vb.narrow(1, ki * cs, cs),
                         o.narrow(1, qi * cs, cs), lse.narrow(2, qi * cs, cs), causal && kpos == qpos);
        dq.narrow(1, qi * cs, cs).add_(r[0].to(at::kFloat));
        buf[2].narrow(1, ki * cs, cs).add_(r[1].to(at::kFloat));
        buf[3].narrow(1, ki * cs, cs).add_(r[2].to(at::kFloat));
      }
    }
    // rotate (after the last round the buffer is sent once more so every block returns home)
    at::Tensor nb = at::empty_like(buf);
    std::vector<std::pair<at::Tensor, int>> sends = {{buf, next}}, recvs = {{nb, prev}};
    comm.batched_send_recv(sends, recvs);
    buf = nb;
  }
  return {dq.to(q.scalar_type()), buf[2].to(k.scalar_type()), buf[3].to(v.scalar_type())};
}
static TensorList parallel_attn_grad(OpDef& op, const TensorList& g) {
  TensorList ins = {g[0], op.inputs[0], op.inputs[1], op.inputs[2], op.outputs[0], op.outputs[1]};
  if (op.inputs.size() > 3) ins.push_back(op.inputs[3]);          // cu_seqlens of a packed row
  TensorList r = op.graph->make_op("parallel_attn_bwd", ins, op.attrs);
  if (op.inputs.size() > 3) return {r[0], r[1], r[2], nullptr};
  return {r[0], r[1], r[2]};
}
static void pattn_deduce(OpDef& op, size_t s) { copy_out_ds(op, 0, s, op.inputs[0]); }
static void pattn_bwd_deduce(OpDef& op, size_t s) {
  copy_out_ds(op, 0, s, op.inputs[1]);
  copy_out_ds(op, 1, s, op.inputs[2]);
  copy_out_ds(op, 2, s, op.inputs[3]);
}
HB_REGISTER_OP(parallel_attn, "parallel_attn", 2, kFlagAttention | kFlagComm, parallel_attn_compute, parallel_attn_grad, pattn_deduce, nullptr);
HB_REGISTER_OP(parallel_attn_bwd, "parallel_attn_bwd", 3, kFlagAttention | kFlagComm, parallel_attn_bwd_compute, nullptr, pattn_bwd_deduce, nullptr);

}  // namespace hb
