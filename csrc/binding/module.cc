// Python binding of the native core (module hetu_b200._C).
// (capability parity: python/hetu/_binding/** -- device, DeviceGroup, dtype, IntSymbol, Tensor,
//  DistributedStates(+Union), Graph, graph.run, comm-group init, plus the generic op entry point
//  that the generated Python op wrappers call)
#include <pybind11/functional.h>
#include <pybind11/stl.h>
#include <torch/extension.h>

#include "../core/device.h"
#include "../core/ds.h"
#include "../core/symbol.h"
#include "../core/utils.h"
#include "../graph/exec.h"
#include "../graph/ir.h"
#include "../graph/op_utils.h"
#include "../kernels/attention_sm100.h"
#include "../kernels/gemm_sm100.h"
#include "../kernels/kernels.h"
#include "../planner/dp_core.h"
#include "../v1/embedding_cache.h"
#include <torch/csrc/cuda/CUDAPluggableAllocator.h>

#include "../runtime/symm_mem.h"
#include "../runtime/memory_pool.h"
#include "../v1/ps_server.h"
#include "../v1/ps_net.h"
#include "../v1/ps_scheduler.h"
#include "../runtime/runtime.h"
#include "../runtime/rpc_client.h"

namespace py = pybind11;
using namespace hb;

namespace {

AttrMap attrs_from_dict(const py::dict& d) {
  AttrMap a;
  for (auto item : d) {
    const std::string k = py::cast<std::string>(item.first);
    py::handle v = item.second;
    if (v.is_none()) continue;
    if (py::isinstance<py::bool_>(v)) a.set(k, py::cast<bool>(v));
    else if (py::isinstance<py::int_>(v)) a.set(k, py::cast<int64_t>(v));
    else if (py::isinstance<py::float_>(v)) a.set(k, py::cast<double>(v));
    else if (py::isinstance<py::str>(v)) a.set(k, py::cast<std::string>(v));
    else if (py::isinstance<py::sequence>(v)) {
      py::sequence s = py::reinterpret_borrow<py::sequence>(v);
      bool all_int = true;
      for (auto e : s) if (!py::isinstance<py::int_>(e) || py::isinstance<py::bool_>(e)) all_int = false;
      if (all_int) a.set(k, py::cast<std::vector<int64_t>>(v));
      else a.set(k, py::cast<std::vector<double>>(v));
    } else throw std::runtime_error("unsupported attribute type for key " + k);
  }
  return a;
}
py::dict attrs_to_dict(const AttrMap& a) {
  py::dict d;
  for (auto& kv : a.raw()) std::visit([&](auto&& v) { d[py::str(kv.first)] = py::cast(v); }, kv.second);
  return d;
}
DeviceGroupHierarchy dgh_from_py(const py::object& o) {
  DeviceGroupHierarchy h;
  if (o.is_none()) return h;
  for (auto u : py::reinterpret_borrow<py::sequence>(o)) {
    if (py::isinstance<DeviceGroupUnion>(u)) { h.add(py::cast<DeviceGroupUnion>(u)); continue; }
    if (py::isinstance<DeviceGroup>(u)) { h.add(DeviceGroupUnion({py::cast<DeviceGroup>(u)})); continue; }
    DeviceGroupUnion un;
    for (auto g : py::reinterpret_borrow<py::sequence>(u)) un.add(py::cast<DeviceGroup>(g));
    h.add(un);
  }
  return h;
}
DistributedStatesHierarchy dsh_from_py(const py::object& o) {
  DistributedStatesHierarchy h;
  if (o.is_none()) return h;
  for (auto u : py::reinterpret_borrow<py::sequence>(o)) {
    if (py::isinstance<DistributedStatesUnion>(u)) h.add(py::cast<DistributedStatesUnion>(u));
    else if (py::isinstance<DistributedStates>(u)) h.add(DistributedStatesUnion({py::cast<DistributedStates>(u)}));
    else {
      DistributedStatesUnion un;
      for (auto d : py::reinterpret_borrow<py::sequence>(u)) un.add(py::cast<DistributedStates>(d));
      h.add(un);
    }
  }
  return h;
}

}  // namespace

PYBIND11_MODULE(_C, m) {
  m.doc() = "hetu_b200 native core";
  py::register_exception<hb::Error>(m, "HetuError");

  // ---------------------------------------------------------------- core
  py::class_<Device>(m, "device")
      .def(py::init<const std::string&>())
      .def_property_readonly("index", &Device::index)
      .def_property_readonly("hostname", &Device::hostname)
      .def_property_readonly("multiplex", &Device::multiplex)
      .def_property_readonly("is_cpu", &Device::is_cpu)
      .def_property_readonly("is_cuda", &Device::is_cuda)
      .def_property_readonly("local", &Device::local)
      .def("__str__", &Device::str)
      .def("__repr__", [](const Device& d) { return "device(" + d.str() + ")"; })
      .def("__eq__", [](const Device& a, const Device& b) { return a == b; })
      .def("__lt__", [](const Device& a, const Device& b) { return a < b; })
      .def("__hash__", &Device::hash);

  py::class_<DeviceGroup>(m, "DeviceGroup")
      .def(py::init<>())
      .def(py::init<std::vector<std::string>>())
      .def(py::init<std::vector<Device>>())
      .def_property_readonly("num_devices", &DeviceGroup::num_devices)
      .def_property_readonly("empty", &DeviceGroup::empty)
      .def("contains", &DeviceGroup::contains)
      .def("get", &DeviceGroup::get)
      .def("get_index", &DeviceGroup::get_index)
      .def_property_readonly("devices", &DeviceGroup::devices)
      .def("indices", [](const DeviceGroup& g) {
        std::vector<int> v;
        for (auto& d : g.devices()) v.push_back(d.index());
        return v;
      })
      .def("__len__", &DeviceGroup::num_devices)
      .def("__eq__", [](const DeviceGroup& a, const DeviceGroup& b) { return a == b; })
      .def("__repr__", &DeviceGroup::str);

  py::class_<DeviceGroupUnion>(m, "DeviceGroupUnion")
      .def(py::init<std::vector<DeviceGroup>>())
      .def("size", &DeviceGroupUnion::size)
      .def("get", &DeviceGroupUnion::get)
      .def("all", &DeviceGroupUnion::all)
      .def("get_index", &DeviceGroupUnion::get_index)
      .def("raw", &DeviceGroupUnion::raw);

  py::class_<IntSymbol>(m, "IntSymbol")
      .def(py::init<>())
      .def(py::init<int64_t>())
      .def_property("data", &IntSymbol::get_val, &IntSymbol::set_val)
      .def("get_data", &IntSymbol::get_val)
      .def("set_data", &IntSymbol::set_val)
      .def("reset_data", &IntSymbol::reset)
      .def("is_leaf", &IntSymbol::is_leaf)
      .def("is_instantiated", &IntSymbol::is_instantiated)
      .def("__add__", [](const IntSymbol& a, const IntSymbol& b) { return a + b; })
      .def("__add__", [](const IntSymbol& a, int64_t b) { return a + IntSymbol(b); })
      .def("__radd__", [](const IntSymbol& a, int64_t b) { return IntSymbol(b) + a; })
      .def("__sub__", [](const IntSymbol& a, const IntSymbol& b) { return a - b; })
      .def("__sub__", [](const IntSymbol& a, int64_t b) { return a - IntSymbol(b); })
      .def("__rsub__", [](const IntSymbol& a, int64_t b) { return IntSymbol(b) - a; })
      .def("__mul__", [](const IntSymbol& a, const IntSymbol& b) { return a * b; })
      .def("__mul__", [](const IntSymbol& a, int64_t b) { return a * IntSymbol(b); })
      .def("__rmul__", [](const IntSymbol& a, int64_t b) { return IntSymbol(b) * a; })
      .def("__truediv__", [](const IntSymbol& a, const IntSymbol& b) { return a / b; })
      .def("__truediv__", [](const IntSymbol& a, int64_t b) { return a / IntSymbol(b); })
      .def("__floordiv__", [](const IntSymbol& a, const IntSymbol& b) { return a / b; })
      .def("__floordiv__", [](const IntSymbol& a, int64_t b) { return a / IntSymbol(b); })
      .def("__mod__", [](const IntSymbol& a, const IntSymbol& b) { return a % b; })
      .def("__mod__", [](const IntSymbol& a, int64_t b) { return a % IntSymbol(b); })
      .def("__repr__", [](const IntSymbol& s) {
        return s.is_instantiated() ? "IntSymbol(" + std::to_string(s.get_val()) + ")" : std::string("IntSymbol(?)");
      });

  py::class_<DistributedStates>(m, "DistributedStates")
      .def(py::init<>())
      .def(py::init<int, std::map<int, int>, std::vector<int>, bool>(), py::arg("device_num"), py::arg("states"),
           py::arg("order") = std::vector<int>{}, py::arg("zero") = false)
      .def_property_readonly("device_num", &DistributedStates::device_num)
      .def_property_readonly("states", [](const DistributedStates& d) { return d.states(); })
      .def_property_readonly("order", &DistributedStates::order)
      .def_property("zero", &DistributedStates::zero, &DistributedStates::set_zero)
      .def_property_readonly("is_none", &DistributedStates::is_none)
      .def_property_readonly("is_valid", &DistributedStates::is_valid)
      .def("get_dim", &DistributedStates::get_dim)
      .def("check_equal", &DistributedStates::check_equal)
      .def("check_pure_duplicate", &DistributedStates::check_pure_duplicate)
      .def("check_max_dim", &DistributedStates::check_max_dim)
      .def("combine_states", [](const DistributedStates& d, const std::vector<int>& src, int dst) { return d.combine_states(src, dst); })
      .def("combine_order", [](const DistributedStates& d, const std::vector<int>& src, int dst) { return d.combine_order(src, dst); })
      .def("check_combine", &DistributedStates::check_combine)
      .def("reduce_states", &DistributedStates::reduce_states)
      .def("reduce_order", &DistributedStates::reduce_order)
      .def("check_split", &DistributedStates::check_split)
      .def("check_scatter", &DistributedStates::check_scatter)
      .def("check_allreduce", &DistributedStates::check_allreduce)
      .def("check_allgather", &DistributedStates::check_allgather)
      .def("check_reducescatter", &DistributedStates::check_reducescatter)
      .def("check_broadcast", &DistributedStates::check_broadcast)
      .def("check_reduce", &DistributedStates::check_reduce)
      .def("get_split_dim", &DistributedStates::get_split_dim)
      .def("get_loop_sizes", &DistributedStates::get_loop_sizes)
      .def("map_device_to_state_index", &DistributedStates::map_device_to_state_index)
      .def("get_dup_group_index", &DistributedStates::get_dup_group_index)
      .def("get_device_indices_by_dim", &DistributedStates::get_device_indices_by_dim)
      .def("get_devices_by_dim", &DistributedStates::get_devices_by_dim)
      .def("local_shape", &DistributedStates::local_shape)
      .def("global_shape", &DistributedStates::global_shape)
      .def("local_slice", [](const DistributedStates& d, const std::vector<int64_t>& g, int idx) {
        std::vector<int64_t> b, s;
        d.local_slice(g, idx, &b, &s);
        return std::make_pair(b, s);
      })
      .def("__eq__", [](const DistributedStates& a, const DistributedStates& b) { return a.check_equal(b); })
      .def("__repr__", &DistributedStates::str);

  py::class_<DistributedStatesUnion>(m, "DistributedStatesUnion")
      .def(py::init<>())
      .def(py::init<std::vector<DistributedStates>, int, bool>(), py::arg("ds_list"), py::arg("hetero_dim") = kNullHeteroDim,
           py::arg("contiguous") = true)
      .def("size", &DistributedStatesUnion::size)
      .def("get", &DistributedStatesUnion::get)
      .def("get_local", &DistributedStatesUnion::get_local)
      .def("is_hetero", &DistributedStatesUnion::is_hetero)
      .def_property("hetero_dim", &DistributedStatesUnion::hetero_dim, &DistributedStatesUnion::set_hetero_dim)
      .def_property_readonly("ds_list", &DistributedStatesUnion::raw)
      .def("check_equal", &DistributedStatesUnion::check_equal)
      .def("to_hetero", &DistributedStatesUnion::to_hetero)
      .def("__repr__", &DistributedStatesUnion::str);

  py::enum_<CommType>(m, "CommType")
      .value("UNUSED", CommType::UNUSED).value("P2P", CommType::P2P).value("COMM_SPLIT", CommType::COMM_SPLIT)
      .value("SCATTER", CommType::SCATTER).value("ALL_REDUCE", CommType::ALL_REDUCE).value("ALL_GATHER", CommType::ALL_GATHER)
      .value("REDUCE_SCATTER", CommType::REDUCE_SCATTER).value("BROADCAST", CommType::BROADCAST).value("REDUCE", CommType::REDUCE)
      .value("SPLIT_ALL_REDUCE", CommType::SPLIT_ALL_REDUCE).value("SPLIT_REDUCE_SCATTER", CommType::SPLIT_REDUCE_SCATTER)
      .value("SPLIT_ALL_GATHER", CommType::SPLIT_ALL_GATHER).value("BATCHED_ISEND_IRECV", CommType::BATCHED_ISEND_IRECV)
      .value("ALL_TO_ALL", CommType::ALL_TO_ALL);
  m.def("classify_comm", &classify_comm);
  m.def("classify_comm_union", &classify_comm_union);
  m.def("plan_comm", [](const DistributedStates& s, const DistributedStates& d, const DeviceGroup& g, int idx) {
    CommPlan p = plan_comm(s, d, g, idx);
    return py::make_tuple(p.type, p.dim, p.group);
  });
  m.def("plan_resharding", [](const std::vector<int64_t>& gshape, const DistributedStates& s, const std::vector<int>& sr,
                              const DistributedStates& d, const std::vector<int>& dr, const std::string& algo, int per_node) {
    SwitchAlgorithm a = SwitchAlgorithm::NEW_GREEDY;
    if (algo == "FCFS") a = SwitchAlgorithm::FCFS;
    else if (algo == "ROUND_ROBIN") a = SwitchAlgorithm::ROUND_ROBIN;
    else if (algo == "MULTI_NODE_ROUND_ROBIN") a = SwitchAlgorithm::MULTI_NODE_ROUND_ROBIN;
    else if (algo == "GREEDY") a = SwitchAlgorithm::GREEDY;
    std::vector<int64_t> load;
    auto plan = plan_resharding(gshape, s, sr, d, dr, a, &load, per_node);
    py::list items;
    for (auto& t : plan) items.append(py::make_tuple(t.src_device, t.dst_device, t.global.begin, t.global.size));
    return py::make_tuple(items, load);
  }, py::arg("global_shape"), py::arg("src_ds"), py::arg("src_ranks"), py::arg("dst_ds"), py::arg("dst_ranks"),
     py::arg("algorithm") = "NEW_GREEDY", py::arg("devices_per_node") = 8);

  // ---------------------------------------------------------------- graph
  py::class_<TensorDef, Tensor>(m, "Tensor")
      .def_readonly("id", &TensorDef::id)
      .def_readwrite("name", &TensorDef::name)
      .def_property_readonly("shape", [](const TensorDef& t) { return t.shape; })
      .def_property_readonly("dtype", [](const TensorDef& t) { return std::string(dtype_name(t.dtype)); })
      .def_readwrite("requires_grad", &TensorDef::requires_grad)
      .def_readonly("is_grad", &TensorDef::is_grad)
      .def_property_readonly("ndim", &TensorDef::ndim)
      .def_property_readonly("global_shape", [](const TensorDef& t) { return t.global_shape(t.graph ? t.graph->cur_strategy() : 0); })
      .def_property_readonly("symbolic_shape", [](const TensorDef& t) { return t.symbolic_shape; })
      .def("set_symbolic_shape", [](TensorDef& t, const SyShape& s) { t.symbolic_shape = s; })
      .def_property_readonly("ds_hierarchy", [](const TensorDef& t) { return t.ds_hierarchy.raw(); })
      .def_property_readonly("distributed_states", [](const TensorDef& t) -> py::object {
        const size_t s = t.graph ? t.graph->cur_strategy() : 0;
        if (!t.has_ds(s)) return py::none();
        return py::cast(t.ds(s));
      })
      .def("get_ds", [](const TensorDef& t, size_t s) -> py::object { return t.has_ds(s) ? py::cast(t.ds(s)) : py::none(); })
      .def("get_ds_union", [](const TensorDef& t, size_t s) { return t.ds_hierarchy.get(s); })
      .def_property_readonly("producer_type", [](const TensorDef& t) { return t.producer ? t.producer->type : std::string(); })
      .def_property_readonly("producer_name", [](const TensorDef& t) { return t.producer ? t.producer->type + ":" + t.producer->name() : std::string(); })
      .def_property_readonly("producer_id", [](const TensorDef& t) { return t.producer ? t.producer->id : (OpId)-1; })
      .def_property_readonly("device_group", [](const TensorDef& t) {
        return t.producer ? t.producer->placement(t.graph ? t.graph->cur_strategy() : 0) : DeviceGroup();
      })
      .def_property_readonly("grad", [](const TensorDef& t) { return t.grad; })
      .def_property_readonly("graph_id", [](const TensorDef& t) { return (uintptr_t)t.graph; })
      .def("eager_data", [](const TensorDef& t) { return t.eager_data; })
      .def("set_eager_data", [](TensorDef& t, const at::Tensor& v) { t.eager_data = v; t.shape = v.sizes().vec(); })
      .def("__hash__", [](const TensorDef& t) { return (size_t)t.id ^ ((size_t)(uintptr_t)t.graph << 20); })
      .def("__eq__", [](const Tensor& a, const py::object& b) {
        if (!py::isinstance<TensorDef>(b)) return false;
        return a.get() == py::cast<Tensor>(b).get();
      })
      .def("__repr__", [](const TensorDef& t) {
        std::ostringstream os;
        os << "Tensor(" << t.name << ", shape=" << t.shape << ", dtype=" << dtype_name(t.dtype) << ")";
        return os.str();
      });

  py::enum_<GraphKind>(m, "GraphKind")
      .value("EAGER", GraphKind::EAGER).value("DEFINE_BY_RUN", GraphKind::DEFINE_BY_RUN)
      .value("DEFINE_AND_RUN", GraphKind::DEFINE_AND_RUN).value("EXECUTABLE", GraphKind::EXECUTABLE);

  py::class_<Graph, std::shared_ptr<Graph>>(m, "Graph")
      .def(py::init([](GraphKind k, const std::string& name, int ns) { return Graph::make(k, name, ns); }))
      .def_static("default_eager", &Graph::default_eager)
      .def_property_readonly("kind", &Graph::kind)
      .def_property_readonly("id", [](const Graph& g) { return (uintptr_t)&g; })
      .def_property_readonly("name", &Graph::name)
      .def_property("num_strategy", &Graph::num_strategy, &Graph::set_num_strategy)
      .def_property("cur_strategy", &Graph::cur_strategy, &Graph::set_cur_strategy)
      .def_property_readonly("num_ops", &Graph::num_ops)
      .def("op_types", [](const Graph& g) {
        std::vector<std::string> v;
        for (auto& op : g.ops()) v.push_back(op->type);
        return v;
      })
      .def("set_op_attrs", [](Graph& g, OpId id, const py::dict& attrs) {
        // attributes are read by the op bodies at execution time (e.g. the optimizer's lr / weight decay schedule)
        AttrMap upd = attrs_from_dict(attrs);
        auto op = g.op(id);
        for (auto& kv : upd.raw()) op->attrs.set(kv.first, kv.second);
      })
      .def("op_info", [](const Graph& g, OpId id) {
        auto op = g.op(id);
        py::dict d;
        d["type"] = op->type;
        d["name"] = op->name();
        d["attrs"] = attrs_to_dict(op->attrs);
        d["inputs"] = op->inputs;
        d["outputs"] = op->outputs;
        d["is_bwd"] = op->is_bwd;
        d["fw_op_id"] = op->fw_op_id;
        d["subgraph"] = op->meta.subgraph;
        d["placement"] = op->placement(g.cur_strategy());
        return d;
      })
      .def("parameters", &Graph::parameters)
      .def("push_subgraph", &Graph::push_subgraph)
      .def("pop_subgraph", &Graph::pop_subgraph)
      .def("subgraphs", [](Graph& g) {
        py::dict d;
        for (auto& kv : g.subgraphs()) {
          py::dict e;
          e["module_type"] = kv.second.module_type;
          e["parent"] = kv.second.parent;
          e["fwd_ops"] = kv.second.fwd_ops;
          e["bwd_ops"] = kv.second.bwd_ops;
          e["update_ops"] = kv.second.update_ops;
          d[py::str(kv.first)] = e;
        }
        return d;
      })
      .def("push_ctx", [](Graph& g, const py::object& dgh, int stream_index, const TensorList& deps, const std::vector<bool>& rec,
                          const std::vector<bool>& off) {
        g.ctx_stack().push_back(g.ctx());
        if (!dgh.is_none()) g.ctx().dg_hierarchy = dgh_from_py(dgh);
        if (stream_index >= 0) g.ctx().stream_index = stream_index;
        for (auto& d : deps) g.ctx().extra_deps.push_back(d);
        if (!rec.empty()) g.ctx().recompute = rec;
        if (!off.empty()) g.ctx().cpu_offload = off;
      }, py::arg("device_group_hierarchy") = py::none(), py::arg("stream_index") = -1, py::arg("extra_deps") = TensorList{},
         py::arg("recompute") = std::vector<bool>{}, py::arg("cpu_offload") = std::vector<bool>{})
      .def("pop_ctx", [](Graph& g) {
        HB_CHECK(!g.ctx_stack().empty()) << "context stack underflow";
        g.ctx() = g.ctx_stack().back();
        g.ctx_stack().pop_back();
      })
      .def("gradients", &Graph::gradients, py::arg("ys"), py::arg("xs"), py::arg("grad_ys") = TensorList{},
           py::call_guard<py::gil_scoped_release>())
      .def("backward", &Graph::eager_backward, py::arg("loss"), py::arg("grad") = nullptr, py::call_guard<py::gil_scoped_release>())
      .def("make_op", [](Graph& g, const std::string& type, const TensorList& inputs, const py::dict& attrs, const std::string& name,
                         const py::object& dgh, const py::object& dst_ds, const SyShape& sy_shape, const py::object& const_data,
                         int stream_index, const TensorList& extra_deps) {
        OpMeta meta;
        meta.name = name;
        meta.dg_hierarchy = dgh_from_py(dgh);
        meta.stream_index = stream_index;
        meta.extra_deps = extra_deps;
        DistributedStatesHierarchy ds = dsh_from_py(dst_ds);
        at::Tensor cd;
        if (!const_data.is_none()) cd = py::cast<at::Tensor>(const_data);
        AttrMap amap = attrs_from_dict(attrs);
        py::gil_scoped_release nogil;
        if (ds.size() == 0 && sy_shape.empty() && !cd.defined()) return g.make_op(type, inputs, amap, meta);   // plain op (reusable)
        return g.make_op(type, inputs, amap, meta, [&](OpDef& op) {
          op.dst_ds = ds;
          op.sy_shape = sy_shape;
          if (cd.defined()) op.const_data = cd;
        });
      }, py::arg("type"), py::arg("inputs"), py::arg("attrs") = py::dict(), py::arg("name") = "",
         py::arg("device_group_hierarchy") = py::none(), py::arg("dst_ds") = py::none(), py::arg("sy_shape") = SyShape{},
         py::arg("const_data") = py::none(), py::arg("stream_index") = -1, py::arg("extra_deps") = TensorList{})
      .def("run", [](Graph& g, const Tensor& loss, const TensorList& fetches, const py::dict& feed, int num_micro_batches,
                     int strategy, int run_level, double grad_scale, bool save_checkpoint,
                     const std::vector<std::pair<IntSymbol, std::vector<int64_t>>>& symbols) {
        std::unordered_map<TensorId, std::vector<at::Tensor>> f;
        for (auto item : feed) {
          Tensor t = py::cast<Tensor>(item.first);
          std::vector<at::Tensor> vs;
          if (py::isinstance<py::list>(item.second) || py::isinstance<py::tuple>(item.second))
            for (auto e : py::reinterpret_borrow<py::sequence>(item.second)) vs.push_back(py::cast<at::Tensor>(e));
          else vs.push_back(py::cast<at::Tensor>(item.second));
          f[t->id] = vs;
        }
        RunOptions o;
        o.num_micro_batches = num_micro_batches;
        o.strategy = strategy;
        o.run_level = (RunLevel)run_level;
        o.grad_scale = grad_scale;
        o.save_checkpoint = save_checkpoint;
        o.symbols = symbols;
        py::gil_scoped_release nogil;
        return g.executor()->run(loss, fetches, f, o);
      }, py::arg("loss"), py::arg("fetches"), py::arg("feed_dict") = py::dict(), py::arg("num_micro_batches") = 1,
         py::arg("strategy") = 0, py::arg("run_level") = 0, py::arg("grad_scale") = 1.0, py::arg("save_checkpoint") = false,
         py::arg("symbols") = std::vector<std::pair<IntSymbol, std::vector<int64_t>>>{})
      .def("get_param", [](Graph& g, const Tensor& t) { return g.executor()->get_param(t); })
      .def("set_param", [](Graph& g, const Tensor& t, const at::Tensor& v) { g.executor()->set_param(t, v); })
      .def("has_param", [](Graph& g, const Tensor& t) { return g.has_param_data(t->id); })
      .def("switch_strategy", [](Graph& g, int from, int to) { g.executor()->switch_strategy(from, to); })
      .def("reinfer_shapes", &Graph::reinfer_shapes)
      .def("materialize", [](Graph& g, const Tensor& t) { return g.materialize(t); }, py::call_guard<py::gil_scoped_release>())
      .def("prune", &Graph::prune)
      .def("num_live_ops", &Graph::num_live_ops)
      .def("reuse_hits", &Graph::reuse_hits)
      .def("set_profile", [](Graph& g, bool on) { g.executor()->set_profile(on); })
      .def("set_loss_scaler", [](Graph& g, const Tensor& var, double init_scale, double growth, double backoff, int64_t interval) {
        g.executor()->set_loss_scaler(var, init_scale, growth, backoff, interval);
      })
      .def("loss_scale", [](Graph& g) { return g.executor()->loss_scaler().scale; })
      .def("set_loss_scale", [](Graph& g, double v) {
        auto& s = g.executor()->loss_scaler();
        s.scale = v;
        if (s.scale_var) g.executor()->get_param(s.scale_var).fill_(v);
      })
      .def("loss_scaler_state", [](Graph& g) {
        auto& s = g.executor()->loss_scaler();
        return py::make_tuple(s.enabled, s.scale, s.tracker, s.skipped, s.last_found_inf);
      })
      .def("op_times", [](Graph& g) { return g.executor()->op_times(); })
      .def("step_breakdown", [](Graph& g) { return g.executor()->step_breakdown(); })
      .def("accumulated_grad", [](Graph& g, const Tensor& param) -> py::object {
        auto& m2 = g.executor()->accumulated_grads();
        auto it = m2.find(param->id);
        if (it == m2.end()) return py::none();
        return py::cast(it->second);
      });

  m.def("list_ops", [] { return OpRegistry::get().list(); });
  m.def("has_op", [](const std::string& t) { return OpRegistry::get().find(t) != nullptr; });

  // ---------------------------------------------------------------- communication
  m.def("init_comm", [](int rank, int world, const PG& world_pg, py::function factory) {
    auto fac = [factory](const std::vector<int>& ranks) {
      py::gil_scoped_acquire gil;
      py::object o = factory(ranks);
      if (o.is_none()) return PG();   // this rank is not a member of the group
      return py::cast<PG>(o);
    };
    CommRuntime::get().init(rank, world, world_pg, fac);
  });
  m.def("comm_initialized", [] { return CommRuntime::get().initialized(); });
  m.def("comm_create_group", [](const std::vector<int>& ranks) { return (bool)CommRuntime::get().group(ranks); },
        "create (or look up) the process group over `ranks`; collective among its members");
  m.def("comm_rank", [] { return CommRuntime::get().rank(); });
  m.def("comm_world", [] { return CommRuntime::get().world(); });
  m.def("comm_barrier", [] { CommRuntime::get().barrier(); });
  m.def("comm_stats", [] { return py::make_tuple(CommRuntime::get().bytes(), CommRuntime::get().calls()); });
  m.def("comm_all_reduce", [](const at::Tensor& x, const std::vector<int>& r, const std::string& red) {
    return CommRuntime::get().all_reduce(x, r, reduction_from_name(red));
  }, py::arg("x"), py::arg("ranks"), py::arg("reduction") = "sum");
  m.def("comm_all_gather", [](const at::Tensor& x, const std::vector<int>& r, int dim) { return CommRuntime::get().all_gather(x, r, dim); });
  m.def("comm_reduce_scatter", [](const at::Tensor& x, const std::vector<int>& r, int dim) { return CommRuntime::get().reduce_scatter(x, r, dim); });
  m.def("comm_all_to_all", [](const at::Tensor& x, const std::vector<int>& r, int sd, int cd) { return CommRuntime::get().all_to_all(x, r, sd, cd); });
  m.def("comm_broadcast", [](const at::Tensor& x, const std::vector<int>& r, int root) { return CommRuntime::get().broadcast(x, r, root); });
  m.def("comm_send", [](const at::Tensor& x, int dst, int channel) {
    py::gil_scoped_release nogil;
    CommRuntime::get().send(x.contiguous(), dst, channel);
    CommRuntime::get().flush_sends();
  }, py::arg("x"), py::arg("dst"), py::arg("channel") = 0);
  m.def("comm_recv", [](const std::vector<int64_t>& shape, const std::string& dtype, int src, int channel) {
    py::gil_scoped_release nogil;
    return CommRuntime::get().recv(shape, to_aten_dtype(dtype_from_name(dtype)), aten_device(), src, channel);
  }, py::arg("shape"), py::arg("dtype"), py::arg("src"), py::arg("channel") = 0);
  m.def("comm_all_reduce_coalesce", [](const std::vector<at::Tensor>& xs, const std::vector<int>& r, const std::string& red) {
    return CommRuntime::get().all_reduce_coalesce(xs, r, reduction_from_name(red));
  }, py::arg("tensors"), py::arg("ranks"), py::arg("reduction") = "sum");
  m.def("comm_reduce", [](const at::Tensor& x, const std::vector<int>& r, int root, const std::string& red) {
    return CommRuntime::get().reduce(x, r, root, reduction_from_name(red));
  }, py::arg("x"), py::arg("ranks"), py::arg("root"), py::arg("reduction") = "sum");
  m.def("comm_gather", [](const at::Tensor& x, const std::vector<int>& r, int root) { return CommRuntime::get().gather(x, r, root); });
  m.def("comm_scatter", [](const at::Tensor& x, const std::vector<int>& r, int root) { return CommRuntime::get().scatter(x, r, root); });

  // ---------------------------------------------------------------- schedules / planner / cache
  m.def("generate_gpipe_schedule", [](int S, int M, bool inf) {
    std::vector<std::vector<std::pair<int, int>>> out;
    for (auto& st : generate_gpipe_schedule(S, M, inf)) {
      out.emplace_back();
      for (auto& t : st) out.back().push_back({(int)t.kind, t.micro_batch});
    }
    return out;
  }, py::arg("num_stages"), py::arg("num_micro_batches"), py::arg("inference") = false);
  m.def("generate_1f1b_schedule", [](int S, int M, bool inf) {
    std::vector<std::vector<std::pair<int, int>>> out;
    for (auto& st : generate_1f1b_schedule(S, M, inf)) {
      out.emplace_back();
      for (auto& t : st) out.back().push_back({(int)t.kind, t.micro_batch});
    }
    return out;
  }, py::arg("num_stages"), py::arg("num_micro_batches"), py::arg("inference") = false);
  m.def("galvatron_dp", [](int layers, int max_mem, int S, const std::vector<int>& mem, const std::vector<double>& intra,
                           const std::vector<double>& inter) {
    DpResult r = galvatron_dp(layers, max_mem, S, mem, intra, inter);
    return py::make_tuple(r.cost, r.strategies, r.mem_remaining);
  });

  py::enum_<CachePolicy>(m, "CachePolicy").value("LRU", CachePolicy::LRU).value("LFU", CachePolicy::LFU).value("LFUOPT", CachePolicy::LFUOPT);
  py::class_<EmbeddingCache>(m, "EmbeddingCache")
      .def(py::init<int64_t, int, CachePolicy, int64_t, int64_t>(), py::arg("capacity"), py::arg("width"),
           py::arg("policy") = CachePolicy::LRU, py::arg("pull_bound") = 0, py::arg("push_bound") = 0)
      .def_property_readonly("size", &EmbeddingCache::size)
      .def_property_readonly("capacity", &EmbeddingCache::capacity)
      .def("contains", &EmbeddingCache::contains)
      .def("stats", [](const EmbeddingCache& c) {
        py::dict d;
        d["lookups"] = c.stats().lookups; d["hits"] = c.stats().hits; d["evictions"] = c.stats().evictions;
        d["pushes"] = c.stats().pushes; d["pulls"] = c.stats().pulls;
        return d;
      })
      .def("lookup", [](EmbeddingCache& c, const std::vector<int64_t>& keys, const std::vector<int64_t>& versions) {
        at::Tensor out = at::zeros({(int64_t)keys.size(), c.width()}, at::kFloat);
        auto miss = c.lookup(keys, versions, out.data_ptr<float>());
        return py::make_tuple(out, miss);
      })
      .def("insert", [](EmbeddingCache& c, const std::vector<int64_t>& keys, const at::Tensor& rows, const std::vector<int64_t>& versions) {
        at::Tensor r = rows.to(at::kFloat).contiguous();
        std::vector<int64_t> ek;
        std::vector<float> eg;
        c.insert(keys, r.data_ptr<float>(), versions, &ek, &eg);
        at::Tensor g = at::from_blob(eg.data(), {(int64_t)ek.size(), c.width()}, at::kFloat).clone();
        return py::make_tuple(ek, g);
      })
      .def("update", [](EmbeddingCache& c, const std::vector<int64_t>& keys, const at::Tensor& grads, float lr) {
        at::Tensor g = grads.to(at::kFloat).contiguous();
        std::vector<int64_t> pk;
        std::vector<float> pg;
        c.update(keys, g.data_ptr<float>(), lr, &pk, &pg);
        at::Tensor out = at::from_blob(pg.data(), {(int64_t)pk.size(), c.width()}, at::kFloat).clone();
        return py::make_tuple(pk, out);
      })
      .def("flush", [](EmbeddingCache& c) {
        std::vector<int64_t> pk;
        std::vector<float> pg;
        c.flush(&pk, &pg);
        at::Tensor out = at::from_blob(pg.data(), {(int64_t)pk.size(), c.width()}, at::kFloat).clone();
        return py::make_tuple(pk, out);
      });

  // ---------------------------------------------------------------- diagnostics
  m.def("kernel_launch_count", [] { return kernel_launch_count() + gemm_launch_count() + attn_launch_count(); });
  m.def("gemm_launch_count", &gemm_launch_count);
  m.def("attn_launch_count", &attn_launch_count);
  m.def("fallback_count", &fallback_count);
  m.def("set_log_level", [](const std::string& l) {
    std::string s = l;
    for (auto& c : s) c = toupper(c);
    set_log_level(s == "TRACE" ? LogLevel::TRACE : s == "DEBUG" ? LogLevel::DEBUG : s == "INFO" ? LogLevel::INFO
                  : s == "ERROR" ? LogLevel::ERROR : s == "FATAL" ? LogLevel::FATAL : LogLevel::WARN);
  });
  m.def("dtype_size", [](const std::string& n) { return dtype_size(dtype_from_name(n)); });

  // ---------------------------------------------------------------- caching memory pool on the tensor path
  // Replaces PyTorch's CUDA caching allocator by this framework's CachingMemoryPool (csrc/runtime/memory_pool.cc) through
  // the pluggable-allocator interface: every at::Tensor the executor and the ops allocate then comes from the pool --
  // per-stream free lists, split / merge, event-based cross-stream reuse (record_stream -> mark_used_by_stream), the
  // HETU_MAX_SPLIT_SIZE_MB / HETU_MAX_INTERNAL_FRAGMENT_SIZE_MB / HETU_PRE_ALLOCATE_SIZE_MB knobs and the pool statistics.
  // Must run before the first CUDA allocation of the process (hetu_b200/__init__.py does it when HETU_NATIVE_ALLOCATOR=1).
  // (ref: hetu/impl/memory/CUDACachingMemoryPool.cu on the NDArray allocation path, hetu/core/memory_pool.h)
  m.def("use_native_allocator", [] {
    static bool installed = false;
    if (installed) return true;
    auto pool_of = [](int device) { return MemoryPoolRegistry::instance().tensor_allocator("cuda:" + std::to_string(device)); };
    auto alloc = torch::cuda::CUDAPluggableAllocator::createCustomAllocator(
        [pool_of](size_t size, int device, cudaStream_t stream) -> void* {
          if (size == 0) return nullptr;
          void* p = pool_of(device)->alloc(size, (int64_t)(uintptr_t)stream);
          if (p == nullptr) {                       // out of memory: give cached segments back and retry once
            MemoryPoolRegistry::instance().empty_all_caches();
            p = pool_of(device)->alloc(size, (int64_t)(uintptr_t)stream);
          }
          HB_CHECK(p != nullptr) << "CUDA out of memory: native pool could not serve " << size << " bytes on device " << device << "\n"
                                 << pool_of(device)->summary();
          return p;
        },
        [pool_of](void* ptr, size_t, int device, cudaStream_t) {
          if (ptr != nullptr) pool_of(device)->free(ptr);
        });
    auto* plug = dynamic_cast<torch::cuda::CUDAPluggableAllocator::CUDAPluggableAllocator*>(alloc.get());
    HB_CHECK(plug != nullptr) << "unexpected allocator type";
    plug->set_record_stream_fn([](void* ptr, cudaStream_t stream) {
      int dev = 0;
      cudaGetDevice(&dev);
      MemoryPoolRegistry::instance().tensor_allocator("cuda:" + std::to_string(dev))->mark_used_by_stream(ptr, (int64_t)(uintptr_t)stream);
    });
    plug->set_reset_fn([] {
      MemoryPoolRegistry::instance().empty_all_caches();
      int n = 0;
      if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
      for (int d = 0; d < n; ++d) MemoryPoolRegistry::instance().tensor_allocator("cuda:" + std::to_string(d))->empty_cache();
    });
    torch::cuda::CUDAPluggableAllocator::changeCurrentAllocator(alloc);
    installed = true;
    return true;
  });

  // ---------------------------------------------------------------- symmetric memory (NVLink peer access)
  m.def("symm_alloc", [](const std::string& name, size_t bytes, int rank, int world) {
    return py::bytes(SymmMem::get().alloc(name, bytes, rank, world));
  });
  m.def("symm_open", [](const std::string& name, const std::vector<py::bytes>& handles) {
    std::vector<std::string> hs;
    for (auto& h : handles) hs.push_back(std::string(h));
    SymmMem::get().open(name, hs);
  });
  m.def("symm_has", [](const std::string& name) { return SymmMem::get().has(name); });
  m.def("symm_free_all", [] { SymmMem::get().free_all(); });
  m.def("symm_tensor", [](const std::string& name, size_t byte_offset, const std::vector<int64_t>& shape, const std::string& dtype) {
    SymmBuffer& b = SymmMem::get().buffer(name);
    int dev = 0;
    cudaGetDevice(&dev);
    auto opts = at::TensorOptions().dtype(to_aten_dtype(dtype_from_name(dtype))).device(at::kCUDA, dev);
    return at::from_blob(reinterpret_cast<char*>(b.local) + byte_offset, shape, opts);
  });
  m.def("symm_barrier", [](const std::string& name) { cuda_ok(symm_barrier(SymmMem::get().buffer(name), cur_stream()), "symm_barrier"); });
  m.def("symm_all_gather", [](const std::string& name, size_t src_off, at::Tensor out, size_t bytes_per_rank) {
    cuda_ok(symm_all_gather(SymmMem::get().buffer(name), src_off, out.data_ptr(), bytes_per_rank, cur_stream()), "symm_all_gather");
  });
  m.def("symm_reduce_scatter", [](const std::string& name, size_t src_off, at::Tensor out, size_t elems_per_rank) {
    cuda_ok(symm_reduce_scatter(SymmMem::get().buffer(name), src_off, out.data_ptr(), elems_per_rank,
                                out.scalar_type() == at::kBFloat16, cur_stream()), "symm_reduce_scatter");
  });
  m.def("symm_all_reduce", [](const std::string& name, size_t src_off, size_t elems, bool bf16) {
    cuda_ok(symm_all_reduce(SymmMem::get().buffer(name), src_off, elems, bf16, cur_stream()), "symm_all_reduce");
  });
  m.def("symm_all_to_all", [](const std::string& name, size_t src_off, at::Tensor out, size_t bytes_per_chunk) {
    cuda_ok(symm_all_to_all(SymmMem::get().buffer(name), src_off, out.data_ptr(), bytes_per_chunk, cur_stream()), "symm_all_to_all");
  });
  m.def("symm_launch_count", &symm_launch_count);
  // VMM allocation + NVLS multicast mapping (csrc/runtime/symm_vmm.cc): alloc -> exchange descriptors -> open -> barrier -> bind
  m.def("symm_multicast_supported", [] { return SymmMem::multicast_supported(); });
  m.def("symm_alloc_vmm", [](const std::string& name, size_t bytes, int rank, int world) {
    return py::bytes(SymmMem::get().alloc_vmm(name, bytes, rank, world));
  });
  m.def("symm_open_vmm", [](const std::string& name, const std::vector<py::bytes>& descs) {
    std::vector<std::string> ds;
    for (auto& d : descs) ds.push_back(std::string(d));
    SymmMem::get().open_vmm(name, ds);
  });
  m.def("symm_bind_multicast", [](const std::string& name) { SymmMem::get().bind_multicast(name); });
  m.def("symm_has_multicast", [](const std::string& name) { return SymmMem::get().buffer(name).mc != nullptr; });
  m.def("symm_mc_all_reduce", [](const std::string& name, size_t src_off, size_t elems, bool bf16) {
    cuda_ok(symm_mc_all_reduce(SymmMem::get().buffer(name), src_off, elems, bf16, cur_stream()), "symm_mc_all_reduce");
  });
  m.def("symm_mc_reduce_scatter", [](const std::string& name, size_t src_off, at::Tensor out, size_t elems_per_rank) {
    cuda_ok(symm_mc_reduce_scatter(SymmMem::get().buffer(name), src_off, out.data_ptr(), elems_per_rank,
                                   out.scalar_type() == at::kBFloat16, cur_stream()), "symm_mc_reduce_scatter");
  });
  m.def("symm_mc_all_gather", [](const std::string& name, const at::Tensor& src, size_t dst_off) {
    cuda_ok(symm_mc_all_gather(SymmMem::get().buffer(name), src.data_ptr(), dst_off, (size_t)src.numel() * src.element_size(),
                               cur_stream()), "symm_mc_all_gather");
  });
  // fused row-parallel GEMM -> reduce-scatter: y[T/world, N] = sum_ranks(x_r[T, K_local] * w_r[N, K_local]^T) (+bias +residual).
  // Output tiles are stored from the GEMM epilogue straight into the owner rank's staging slots over NVLink.
  m.def("gemm_reduce_scatter", [](const at::Tensor& x, const at::Tensor& w, const std::string& staging, const py::object& bias,
                                  const py::object& residual) {
    SymmBuffer& b = SymmMem::get().buffer(staging);
    const int64_t T = x.size(0), K = x.size(1), N = w.size(0);
    HB_CHECK(T % b.world == 0 && (T / b.world) % 128 == 0) << "gemm_reduce_scatter: rows per rank must be a multiple of 128";
    HB_CHECK(size_t(T) * N * 2 <= b.bytes) << "staging buffer too small";
    const int64_t rpr = T / b.world;
    GemmCall g;
    g.A = x.data_ptr(); g.B = w.data_ptr(); g.C = b.local;
    g.M = (int)T; g.N = (int)N; g.K = (int)K;
    g.lda = x.stride(0); g.ldb = w.stride(0); g.ldc = N;
    g.peer_c = b.peer; g.world = b.world; g.my_rank = b.rank; g.rows_per_rank = (int)rpr;
    cuda_ok(symm_barrier(b, cur_stream()), "barrier");           // previous consumers of the staging slots are done
    cuda_ok(gemm_bf16(g, cur_stream()), "gemm (peer epilogue)");
    cuda_ok(symm_barrier(b, cur_stream()), "barrier");           // every rank's partial tiles have landed
    at::Tensor out = at::empty({rpr, N}, x.options());
    at::Tensor bt, rt;
    if (!bias.is_none()) bt = py::cast<at::Tensor>(bias);
    if (!residual.is_none()) rt = py::cast<at::Tensor>(residual);
    cuda_ok(symm_reduce_slots(b.local, b.world, out.data_ptr(), bt.defined() ? bt.data_ptr() : nullptr,
                              rt.defined() ? rt.data_ptr() : nullptr, rpr, (int)N, cur_stream()), "reduce_slots");
    return out;
  }, py::arg("x"), py::arg("w"), py::arg("staging"), py::arg("bias") = py::none(), py::arg("residual") = py::none());

  // ---------------------------------------------------------------- native rendezvous client
  {
    using R = RpcClient;
    auto nogil = py::call_guard<py::gil_scoped_release>();
    py::class_<R, std::shared_ptr<R>>(m, "RpcClient")
        .def(py::init<std::string, int, std::string, double, double>(), py::arg("host"), py::arg("port"), py::arg("hostname") = "",
             py::arg("heartbeat_interval") = 2.0, py::arg("connect_timeout") = 60.0, nogil)
        .def("call", &R::call, py::arg("method"), py::arg("args_json") = "{}", nogil)
        .def("connect", &R::connect, py::arg("start_heartbeat") = true, nogil)
        .def_property_readonly("rank", &R::rank)
        .def_property_readonly("local_device", &R::local_device)
        .def_property_readonly("world_size", &R::world_size)
        .def_property_readonly("client_id", &R::client_id)
        .def_property_readonly("heartbeats_sent", &R::heartbeats_sent)
        .def("put_int", &R::put_int, nogil).def("get_int", &R::get_int, nogil)
        .def("put_double", &R::put_double, nogil).def("get_double", &R::get_double, nogil)
        .def("put_string", &R::put_string, nogil).def("get_string", &R::get_string, nogil)
        .def("put_bytes", [](R& c, const std::string& k, const py::bytes& v) { std::string b = v; py::gil_scoped_release g; c.put_bytes(k, b); })
        .def("get_bytes", [](R& c, const std::string& k) { std::string b; { py::gil_scoped_release g; b = c.get_bytes(k); } return py::bytes(b); })
        .def("put_json", &R::put_json, nogil).def("get_json", &R::get_json, nogil)
        .def("remove", &R::remove, py::arg("key"), py::arg("kind") = "json", nogil)
        .def("commit_hostname", &R::commit_hostname, nogil).def("get_hostname", &R::get_hostname, nogil)
        .def("commit_nccl_id", [](R& c, std::vector<int> r, int st, const py::bytes& id) { std::string b = id; py::gil_scoped_release g; c.commit_nccl_id(std::move(r), st, b); })
        .def("get_nccl_id", [](R& c, std::vector<int> r, int st) { std::string b; { py::gil_scoped_release g; b = c.get_nccl_id(std::move(r), st); } return py::bytes(b); })
        .def("barrier", &R::barrier, py::arg("ranks") = std::vector<int>{}, py::arg("tag") = "", nogil)
        .def("consistent", &R::consistent, py::arg("raw_json_value"), py::arg("ranks") = std::vector<int>{}, py::arg("tag") = "", nogil)
        .def("worker_stop", &R::worker_stop, nogil).def("already_stop", &R::already_stop, nogil).def("exit", &R::exit, nogil);
    m.def("json_field", [](const std::string& obj, const std::string& key) -> py::object {
      std::string raw;
      if (!json_field(obj, key, &raw)) return py::none();
      return py::str(raw);
    });
    m.def("json_quote", &json_quote);
    m.def("json_unquote", &json_unquote);
  }

  // ---------------------------------------------------------------- v1 parameter server
  py::enum_<PsOptimizer>(m, "PsOptimizer").value("NONE", PsOptimizer::NONE).value("SGD", PsOptimizer::SGD)
      .value("MOMENTUM", PsOptimizer::MOMENTUM).value("ADAGRAD", PsOptimizer::ADAGRAD).value("ADAM", PsOptimizer::ADAM);
  py::class_<ParameterServer, std::shared_ptr<ParameterServer>>(m, "ParameterServer")
      .def(py::init<int>(), py::arg("num_workers"))
      .def("init_dense", [](ParameterServer& ps, int64_t key, const std::vector<float>& v, PsOptimizer opt, float lr, float momentum) {
        PsParamConfig c; c.opt = opt; c.lr = lr; c.momentum = momentum;
        ps.init_dense(key, v, c);
      }, py::arg("key"), py::arg("value"), py::arg("opt") = PsOptimizer::SGD, py::arg("lr") = 0.01f, py::arg("momentum") = 0.9f)
      .def("init_sparse", [](ParameterServer& ps, int64_t key, int64_t rows, int width, const std::vector<float>& v, PsOptimizer opt, float lr) {
        PsParamConfig c; c.opt = opt; c.lr = lr;
        ps.init_sparse(key, rows, width, v, c);
      }, py::arg("key"), py::arg("rows"), py::arg("width"), py::arg("value"), py::arg("opt") = PsOptimizer::SGD, py::arg("lr") = 0.01f)
      .def("push_dense", &ParameterServer::push_dense, py::call_guard<py::gil_scoped_release>())
      .def("pull_dense", &ParameterServer::pull_dense, py::call_guard<py::gil_scoped_release>())
      .def("push_pull_dense", &ParameterServer::push_pull_dense, py::call_guard<py::gil_scoped_release>())
      .def("push_sparse", &ParameterServer::push_sparse, py::call_guard<py::gil_scoped_release>())
      .def("pull_sparse", &ParameterServer::pull_sparse, py::call_guard<py::gil_scoped_release>())
      .def("row_versions", &ParameterServer::row_versions)
      .def("sync_cache", [](ParameterServer& ps, int64_t key, const std::vector<int64_t>& rows, const std::vector<int64_t>& vers, int64_t bound) {
        std::vector<int64_t> stale, fv;
        std::vector<float> vals;
        ps.sync_cache(key, rows, vers, bound, &stale, &vals, &fv);
        return py::make_tuple(stale, vals, fv);
      })
      .def("barrier", &ParameterServer::barrier, py::call_guard<py::gil_scoped_release>())
      .def("ssp_init", &ParameterServer::ssp_init)
      .def("ssp_sync", &ParameterServer::ssp_sync, py::call_guard<py::gil_scoped_release>())
      .def("preduce", [](ParameterServer& ps, int worker, int64_t key, const std::vector<float>& v, int min_workers, int wait_ms) {
        std::vector<int> partners;
        std::vector<float> out;
        {
          py::gil_scoped_release rel;
          out = ps.preduce(worker, key, v, min_workers, wait_ms, &partners);
        }
        return py::make_tuple(out, partners);
      }, py::arg("worker"), py::arg("key"), py::arg("value"), py::arg("min_workers") = 2, py::arg("wait_ms") = 50)
      .def("stats", &ParameterServer::stats);

  // network transport of the parameter server (multi-process PS / Hybrid jobs)
  py::class_<PsNetServer, std::shared_ptr<PsNetServer>>(m, "PsNetServer")
      .def(py::init<std::shared_ptr<ParameterServer>, int, const std::string&>(), py::arg("ps"), py::arg("port") = 0, py::arg("bind_addr") = "0.0.0.0")
      .def_property_readonly("port", &PsNetServer::port)
      .def_property_readonly("requests", &PsNetServer::requests)
      .def("stop", &PsNetServer::stop, py::call_guard<py::gil_scoped_release>());
  {
    auto nogil = py::call_guard<py::gil_scoped_release>();
    py::class_<PsNodeInfo>(m, "PsNodeInfo")
        .def_readonly("role", &PsNodeInfo::role)
        .def_readonly("rank", &PsNodeInfo::rank)
        .def_readonly("host", &PsNodeInfo::host)
        .def_readonly("port", &PsNodeInfo::port)
        .def("__repr__", [](const PsNodeInfo& n) {
          return std::string(n.role == 0 ? "server" : "worker") + "[" + std::to_string(n.rank) + "]@" + n.host + ":" + std::to_string(n.port);
        });
    py::class_<PsScheduler, std::shared_ptr<PsScheduler>>(m, "PsScheduler")
        .def(py::init<int, int, int, const std::string&>(), py::arg("num_servers"), py::arg("num_workers"), py::arg("port") = 0,
             py::arg("bind_addr") = "0.0.0.0")
        .def_property_readonly("port", &PsScheduler::port)
        .def("stop", &PsScheduler::stop, nogil)
        .def("dead_nodes", &PsScheduler::dead_nodes, py::arg("timeout_s"))
        .def_property_readonly("registered", &PsScheduler::registered)
        .def_property_readonly("finalized", &PsScheduler::finalized)
        .def("wait_finalized", &PsScheduler::wait_finalized, py::arg("timeout_s") = 60.0, nogil);
    py::class_<PsSchedulerClient, std::shared_ptr<PsSchedulerClient>>(m, "PsSchedulerClient")
        .def(py::init<const std::string&, int, int, const std::string&, int, double>(), py::arg("host"), py::arg("port"), py::arg("role"),
             py::arg("my_host") = "127.0.0.1", py::arg("my_port") = 0, py::arg("connect_timeout") = 60.0, nogil)
        .def_property_readonly("role", &PsSchedulerClient::role)
        .def_property_readonly("rank", &PsSchedulerClient::rank)
        .def_property_readonly("num_servers", &PsSchedulerClient::num_servers)
        .def_property_readonly("num_workers", &PsSchedulerClient::num_workers)
        .def_property_readonly("servers", &PsSchedulerClient::servers)
        .def("barrier", &PsSchedulerClient::barrier, py::arg("group") = (int)kAllGroup, nogil)
        .def("preduce_partners", &PsSchedulerClient::preduce_partners, py::arg("key"), py::arg("rank"), py::arg("max_worker"), py::arg("wait_ms"), nogil)
        .def("heartbeat", &PsSchedulerClient::heartbeat, nogil)
        .def("dead_nodes", &PsSchedulerClient::dead_nodes, py::arg("timeout_s"), nogil)
        .def("key_ranges", &PsSchedulerClient::key_ranges, py::arg("total"))
        .def("finalize", &PsSchedulerClient::finalize, nogil)
        .def("start_heartbeat", &PsSchedulerClient::start_heartbeat, py::arg("interval_s") = 1.0);
    py::class_<PsNetClient, std::shared_ptr<PsNetClient>>(m, "PsNetClient")
        .def(py::init<const std::string&, int, double>(), py::arg("host"), py::arg("port"), py::arg("connect_timeout") = 60.0, nogil)
        .def("init_dense", [](PsNetClient& ps, int64_t key, const std::vector<float>& v, PsOptimizer opt, float lr, float momentum) {
          PsParamConfig c; c.opt = opt; c.lr = lr; c.momentum = momentum;
          py::gil_scoped_release rel;
          ps.init_dense(key, v, c);
        }, py::arg("key"), py::arg("value"), py::arg("opt") = PsOptimizer::SGD, py::arg("lr") = 0.01f, py::arg("momentum") = 0.9f)
        .def("init_sparse", [](PsNetClient& ps, int64_t key, int64_t rows, int width, const std::vector<float>& v, PsOptimizer opt, float lr) {
          PsParamConfig c; c.opt = opt; c.lr = lr;
          py::gil_scoped_release rel;
          ps.init_sparse(key, rows, width, v, c);
        }, py::arg("key"), py::arg("rows"), py::arg("width"), py::arg("value"), py::arg("opt") = PsOptimizer::SGD, py::arg("lr") = 0.01f)
        .def("push_dense", &PsNetClient::push_dense, nogil).def("pull_dense", &PsNetClient::pull_dense, nogil)
        .def("push_pull_dense", &PsNetClient::push_pull_dense, nogil)
        .def("push_sparse", &PsNetClient::push_sparse, nogil).def("pull_sparse", &PsNetClient::pull_sparse, nogil)
        .def("row_versions", &PsNetClient::row_versions, nogil)
        .def("sync_cache", [](PsNetClient& ps, int64_t key, const std::vector<int64_t>& rows, const std::vector<int64_t>& vers, int64_t bound) {
          std::vector<int64_t> stale, fv;
          std::vector<float> vals;
          { py::gil_scoped_release rel; ps.sync_cache(key, rows, vers, bound, &stale, &vals, &fv); }
          return py::make_tuple(stale, vals, fv);
        })
        .def("barrier", &PsNetClient::barrier, nogil).def("ssp_init", &PsNetClient::ssp_init, nogil).def("ssp_sync", &PsNetClient::ssp_sync, nogil)
        .def("preduce", [](PsNetClient& ps, int worker, int64_t key, const std::vector<float>& v, int min_workers, int wait_ms) {
          std::vector<int> partners;
          std::vector<float> out;
          { py::gil_scoped_release rel; out = ps.preduce(worker, key, v, min_workers, wait_ms, &partners); }
          return py::make_tuple(out, partners);
        }, py::arg("worker"), py::arg("key"), py::arg("value"), py::arg("min_workers") = 2, py::arg("wait_ms") = 50)
        .def("stats", &PsNetClient::stats, nogil)
        .def("num_workers", &PsNetClient::num_workers, nogil);
  }

  // ---------------------------------------------------------------- native runtime: memory pool, streams, RNG state, data loader
  m.def("get_memory_pool", [](const std::string& device) { return MemoryPoolRegistry::instance().get(device); }, py::arg("device") = "cpu",
        "the process-wide pool of a device: 'cuda:<i>', 'cpu', 'pinned', 'shm'");
  m.def("memory_pool_devices", [] { return MemoryPoolRegistry::instance().devices(); });
  m.def("empty_all_memory_caches", [] { return (uint64_t)MemoryPoolRegistry::instance().empty_all_caches(); });
  auto stats_dict = [](const PoolStats& s) {
    py::dict d;
    d["reserved"] = s.reserved; d["allocated"] = s.allocated; d["peak_reserved"] = s.peak_reserved; d["peak_allocated"] = s.peak_allocated;
    d["num_alloc"] = s.num_alloc; d["num_free"] = s.num_free; d["num_segment_alloc"] = s.num_segment_alloc; d["num_split"] = s.num_split;
    d["num_merge"] = s.num_merge; d["cache_hits"] = s.cache_hits;
    return d;
  };
  py::class_<DeviceAllocator, std::shared_ptr<DeviceAllocator>>(m, "DeviceAllocator")
      .def("alloc", [](DeviceAllocator& p, int64_t bytes, int64_t stream) { return (uint64_t)(uintptr_t)p.alloc((size_t)bytes, stream); },
           py::arg("bytes"), py::arg("stream") = 0)
      .def("free", [](DeviceAllocator& p, uint64_t ptr) { p.free((void*)(uintptr_t)ptr); })
      .def("mark_used_by_stream", [](DeviceAllocator& p, uint64_t ptr, int64_t stream) { p.mark_used_by_stream((void*)(uintptr_t)ptr, stream); })
      .def("wait", [](DeviceAllocator& p, uint64_t ptr) { p.wait((void*)(uintptr_t)ptr); })
      .def("empty_cache", &DeviceAllocator::empty_cache)
      .def("stats", [stats_dict](DeviceAllocator& p) { return stats_dict(p.stats()); })
      .def("summary", &DeviceAllocator::summary)
      .def_property_readonly("kind", [](DeviceAllocator& p) { return std::string(p.kind()); });
  py::class_<BFCMemoryPool, DeviceAllocator, std::shared_ptr<BFCMemoryPool>>(m, "BFCMemoryPool")
      .def(py::init([](const std::string& backend, int device, int64_t initial_region_mb, int64_t limit_mb, int64_t min_chunk) {
        BFCMemoryPool::Options o;
        if (initial_region_mb > 0) o.initial_region = (size_t)initial_region_mb << 20;
        if (limit_mb > 0) o.limit = (size_t)limit_mb << 20;
        if (min_chunk > 0) o.min_chunk = (size_t)min_chunk;
        return std::make_shared<BFCMemoryPool>(backend == "cuda" ? make_cuda_backend(device) : make_host_backend(backend == "pinned"), o);
      }), py::arg("backend") = "host", py::arg("device") = 0, py::arg("initial_region_mb") = 0, py::arg("limit_mb") = 0, py::arg("min_chunk") = 0)
      .def_property_readonly("num_regions", &BFCMemoryPool::num_regions)
      .def("bin_occupancy", &BFCMemoryPool::bin_occupancy)
      .def_property_readonly("largest_free_chunk", &BFCMemoryPool::largest_free_chunk)
      .def_property_readonly("fragmentation", &BFCMemoryPool::fragmentation);
  py::class_<StreamOrderedMemoryPool, DeviceAllocator, std::shared_ptr<StreamOrderedMemoryPool>>(m, "StreamOrderedMemoryPool")
      .def(py::init([](int device, int64_t release_threshold_mb) {
        return std::make_shared<StreamOrderedMemoryPool>(device, release_threshold_mb < 0 ? SIZE_MAX : (size_t)release_threshold_mb << 20);
      }), py::arg("device") = 0, py::arg("release_threshold_mb") = -1)
      .def_property_readonly("on_device", &StreamOrderedMemoryPool::on_device);
  m.def("tensor_allocator", [](const std::string& device) { return MemoryPoolRegistry::instance().tensor_allocator(device); },
        "the allocator that backs CUDA tensors of `device` under HETU_NATIVE_ALLOCATOR=1 (kind: HETU_MEMORY_POOL)");
  py::class_<CachingMemoryPool, DeviceAllocator, std::shared_ptr<CachingMemoryPool>>(m, "MemoryPool")
      .def(py::init([](const std::string& backend, int device, int64_t limit_mb, int64_t max_split_mb, int64_t pre_allocate_mb) {
        CachingMemoryPool::Options o = CachingMemoryPool::options_from_env();
        if (limit_mb > 0) o.limit = (size_t)limit_mb << 20;
        if (max_split_mb > 0) o.max_split_size = (size_t)max_split_mb << 20;
        if (pre_allocate_mb > 0) o.pre_allocate = (size_t)pre_allocate_mb << 20;
        std::unique_ptr<MemoryBackend> be = backend == "cuda" ? make_cuda_backend(device)
                                            : backend == "shm" ? make_shm_backend("hetu_b200") : make_host_backend(backend == "pinned");
        return std::make_shared<CachingMemoryPool>(std::move(be), o);
      }), py::arg("backend") = "host", py::arg("device") = 0, py::arg("limit_mb") = 0, py::arg("max_split_mb") = 0,
           py::arg("pre_allocate_mb") = 0)
      .def("alloc", [](CachingMemoryPool& p, int64_t bytes, int64_t stream) { return (uint64_t)(uintptr_t)p.alloc((size_t)bytes, stream); },
           py::arg("bytes"), py::arg("stream") = 0)
      .def("free", [](CachingMemoryPool& p, uint64_t ptr) { p.free((void*)(uintptr_t)ptr); })
      .def("mark_used_by_stream", [](CachingMemoryPool& p, uint64_t ptr, int64_t stream) { p.mark_used_by_stream((void*)(uintptr_t)ptr, stream); })
      .def("shm_locate", [](CachingMemoryPool& p, uint64_t ptr) -> py::object {
        std::string name;
        size_t off = 0;
        if (!shm_locate(p.backend(), (void*)(uintptr_t)ptr, &name, &off)) return py::none();
        return py::make_tuple(name, (uint64_t)off);
      }, "(shm segment name, offset) of a block of a shared-memory pool: another process maps it with mmap")
      .def("as_tensor", [](std::shared_ptr<CachingMemoryPool> p, uint64_t ptr, std::vector<int64_t> shape, const std::string& dtype) {
        // a host tensor viewing pool memory (no copy); the block stays allocated until the caller frees it
        return at::from_blob((void*)(uintptr_t)ptr, shape, at::TensorOptions().dtype(to_aten_dtype(dtype_from_name(dtype))));
      }, py::arg("ptr"), py::arg("shape"), py::arg("dtype") = "float32")
      .def("wait", [](CachingMemoryPool& p, uint64_t ptr) { p.wait((void*)(uintptr_t)ptr); })
      .def("empty_cache", [](CachingMemoryPool& p) { return (int64_t)p.empty_cache(); })
      .def("summary", &CachingMemoryPool::summary)
      .def("stats", [](const CachingMemoryPool& p) {
        PoolStats s = p.stats();
        py::dict d;
        d["reserved"] = s.reserved; d["allocated"] = s.allocated; d["peak_reserved"] = s.peak_reserved; d["peak_allocated"] = s.peak_allocated;
        d["num_alloc"] = s.num_alloc; d["num_free"] = s.num_free; d["num_segment_alloc"] = s.num_segment_alloc; d["num_split"] = s.num_split;
        d["num_merge"] = s.num_merge; d["cache_hits"] = s.cache_hits;
        return d;
      });
  m.def("stream_role_name", &stream_role_name);
  m.def("logical_stream", [](int device, int index) { return (uint64_t)(uintptr_t)logical_stream(device, index); });
  m.def("sync_logical_stream", &sync_logical_stream);
  m.def("random_seed", [] { return RandomState::get().seed(); });
  m.def("random_set_seed", [](uint64_t s) { RandomState::get().set_seed(s); });
  m.def("random_next_offset", [](uint64_t n) { return RandomState::get().next_offset(n); });
  m.def("random_offset", [] { return RandomState::get().offset(); });
  m.def("random_set_offset", [](uint64_t o) { RandomState::get().set_offset(o); });
  py::class_<NativeDataloader, std::shared_ptr<NativeDataloader>>(m, "Dataloader")
      .def(py::init([](const at::Tensor& data, int64_t batch_size, bool shuffle, bool drop_last, int dp_rank, int dp_size, uint64_t seed,
                       int prefetch, bool pin_memory) {
        return std::make_shared<NativeDataloader>(data, batch_size, shuffle, drop_last, dp_rank, dp_size, seed, prefetch, pin_memory);
      }), py::arg("data"), py::arg("batch_size"), py::arg("shuffle") = false, py::arg("drop_last") = true, py::arg("dp_rank") = 0,
           py::arg("dp_size") = 1, py::arg("seed") = 0, py::arg("prefetch") = 2, py::arg("pin_memory") = false)
      .def("next", [](NativeDataloader& d) { py::gil_scoped_release rel; return d.next(); })
      .def("reset", &NativeDataloader::reset, py::arg("start_batch") = 0)
      .def_property_readonly("num_batches", &NativeDataloader::num_batches);

  // fp8 building blocks (numerics tests / benchmarks)
  m.def("quantize_rowwise_e4m3", [](const at::Tensor& x) {
    HB_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && x.dim() == 2 && x.is_contiguous()) << "expects contiguous 2-D CUDA bf16";
    at::Tensor q = at::empty(x.sizes(), x.options().dtype(at::kByte));
    at::Tensor sc = at::empty({x.size(0)}, x.options().dtype(at::kFloat));
    cuda_ok(quantize_rowwise_e4m3(x.data_ptr(), q.data_ptr(), sc.data_ptr<float>(), x.size(0), (int)x.size(1), x.stride(0), q.stride(0),
                                  cur_stream()), "quantize_rowwise_e4m3");
    return py::make_tuple(q, sc);
  });
  m.def("quantize_transpose_e4m3", [](const at::Tensor& x) {
    HB_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && x.dim() == 2 && x.is_contiguous()) << "expects contiguous 2-D CUDA bf16";
    at::Tensor q = at::empty({x.size(1), x.size(0)}, x.options().dtype(at::kByte));
    at::Tensor sc = at::empty({x.size(1)}, x.options().dtype(at::kFloat));
    cuda_ok(quantize_transpose_e4m3(x.data_ptr(), q.data_ptr(), sc.data_ptr<float>(), x.size(0), (int)x.size(1), x.size(0), cur_stream()),
            "quantize_transpose_e4m3");
    return py::make_tuple(q, sc);
  });
  m.def("gemm_fp8", [](const at::Tensor& qa, const at::Tensor& sa, const at::Tensor& qb, const at::Tensor& sb, bool out_fp32, int cta_group) {
    const int64_t M = qa.size(0), K = qa.size(1), N = qb.size(0);
    at::Tensor c = at::empty({M, N}, qa.options().dtype(out_fp32 ? at::kFloat : at::kBFloat16));
    GemmCall g;
    g.A = qa.data_ptr(); g.B = qb.data_ptr(); g.C = c.data_ptr();
    g.M = (int)M; g.N = (int)N; g.K = (int)K;
    g.lda = qa.stride(0); g.ldb = qb.stride(0); g.ldc = N;
    g.fp8 = true; g.row_scale = sa.data_ptr<float>(); g.col_scale = sb.data_ptr<float>();
    g.out = out_fp32 ? GemmOut::FP32 : GemmOut::BF16;
    g.cta_group = cta_group;
    cuda_ok(gemm_bf16(g, cur_stream()), "gemm_fp8");
    return c;
  }, py::arg("qa"), py::arg("sa"), py::arg("qb"), py::arg("sb"), py::arg("out_fp32") = false, py::arg("cta_group") = 0);

  // ------------------------------------------------------------------ host utilities
  py::class_<ContextStore, std::shared_ptr<ContextStore>>(m, "ContextStore")
      .def(py::init<>())
      .def("put", [](ContextStore& c, const std::string& k, const py::object& v) {
        if (py::isinstance<py::bool_>(v)) c.put<bool>(k, v.cast<bool>());
        else if (py::isinstance<py::int_>(v)) c.put<int64_t>(k, v.cast<int64_t>());
        else if (py::isinstance<py::float_>(v)) c.put<double>(k, v.cast<double>());
        else if (py::isinstance<py::str>(v)) c.put<std::string>(k, v.cast<std::string>());
        else if (THPVariable_Check(v.ptr())) c.put<at::Tensor>(k, v.cast<at::Tensor>());
        else if (py::isinstance<py::sequence>(v)) {
          bool all_int = true;
          for (auto e : v) all_int = all_int && py::isinstance<py::int_>(e) && !py::isinstance<py::bool_>(e);
          if (all_int) c.put<std::vector<int64_t>>(k, v.cast<std::vector<int64_t>>());
          else c.put<std::vector<double>>(k, v.cast<std::vector<double>>());
        } else throw Error("ContextStore.put: unsupported value type for '" + k + "'");
      })
      .def("get", [](const ContextStore& c, const std::string& k, const py::object& dflt) -> py::object {
        if (!c.has(k)) return dflt;
        return std::visit([](const auto& x) -> py::object { return py::cast(x); }, c.raw(k));
      }, py::arg("key"), py::arg("default") = py::none())
      .def("pop", [](ContextStore& c, const std::string& k) -> py::object {
        py::object o = std::visit([](const auto& x) -> py::object { return py::cast(x); }, c.raw(k));
        c.erase(k);
        return o;
      })
      .def("contains", &ContextStore::has)
      .def("__contains__", &ContextStore::has)
      .def("erase", &ContextStore::erase)
      .def("migrate_from", &ContextStore::migrate_from, py::arg("src"), py::arg("key"), py::arg("new_key") = "")
      .def("keys", &ContextStore::keys)
      .def("clear", &ContextStore::clear)
      .def("__len__", &ContextStore::size);
  py::class_<TaskQueue, std::shared_ptr<TaskQueue>>(m, "TaskQueue")
      .def(py::init([](const std::string& name, int workers, size_t max_pending) {
        // the destructor joins the workers, which may still need the GIL for queued Python callables: drop it while joining
        return std::shared_ptr<TaskQueue>(new TaskQueue(name, workers, max_pending), [](TaskQueue* q) {
          py::gil_scoped_release rel;
          delete q;
        });
      }), py::arg("name"), py::arg("num_workers") = 1, py::arg("max_pending") = 1024)
      .def("add", [](TaskQueue& q, py::function fn) {
        // the callable is kept alive by the task and runs with the GIL held; submission may block on a full queue
        auto holder = std::make_shared<py::function>(std::move(fn));
        py::gil_scoped_release rel;
        q.add([holder]() mutable {
          py::gil_scoped_acquire gil;
          try {
            (*holder)();
          } catch (py::error_already_set& e) {
            std::string msg = e.what();
            holder.reset();
            throw Error(msg);
          }
          holder.reset();
        });
      })
      .def("wait", [](TaskQueue& q) { py::gil_scoped_release rel; q.wait(); })
      .def("shutdown", [](TaskQueue& q) { py::gil_scoped_release rel; q.shutdown(); })
      .def_property_readonly("num_workers", &TaskQueue::num_workers)
      .def_property_readonly("running", &TaskQueue::running)
      .def_property_readonly("pending", &TaskQueue::pending)
      .def_property_readonly("completed", &TaskQueue::completed)
      .def_property_readonly("name", &TaskQueue::name);

  // generic long-tail kernels (csrc/kernels/generic.cu) behind at::Tensor; None = not applicable (the ops then use ATen)
  {
    auto opt = [](const at::Tensor& t) -> py::object { return t.defined() ? py::cast(t) : py::none(); };
    py::dict u;
    const char* un[] = {"neg", "reciprocal", "abs", "ceil", "floor", "round", "exp", "log", "sqrt", "rsqrt", "sin", "cos", "clamp", "sigmoid",
                        "tanh", "leakyrelu", "elu", "hardshrink", "hardsigmoid", "hardtanh", "hardswish", "logsigmoid", "softplus", "mish",
                        "softshrink", "pow", "add_scalar", "mul_scalar", "rsub_scalar", "rdiv_scalar", "div_scalar"};
    for (int i = 0; i < (int)(sizeof(un) / sizeof(un[0])); ++i) u[un[i]] = i;
    m.attr("GENERIC_UNARY") = u;
    py::dict b;
    const char* bn[] = {"add", "sub", "mul", "div", "max", "min", "pow"};
    for (int i = 0; i < 7; ++i) b[bn[i]] = i;
    m.attr("GENERIC_BINARY") = b;
    py::dict r;
    const char* rn[] = {"sum", "mean", "max", "min", "prod"};
    for (int i = 0; i < 5; ++i) r[rn[i]] = i;
    m.attr("GENERIC_REDUCE") = r;
    m.def("g_unary", [opt](int op, const at::Tensor& x, float p0, float p1) { return opt(g_unary(op, x, p0, p1)); }, py::arg("op"), py::arg("x"),
          py::arg("p0") = 0.f, py::arg("p1") = 0.f);
    m.def("g_binary", [opt](int op, const at::Tensor& a, const at::Tensor& b2) { return opt(g_binary(op, a, b2)); });
    m.def("g_reduce", [opt](int mode, const at::Tensor& x, std::vector<int64_t> axes, bool keep) { return opt(g_reduce(mode, x, axes, keep)); },
          py::arg("mode"), py::arg("x"), py::arg("axes") = std::vector<int64_t>{}, py::arg("keepdims") = false);
    m.def("g_softmax", [opt](bool log, const at::Tensor& x, int64_t dim) { return opt(g_softmax(log, x, dim)); });
    m.def("g_concat", [opt](const std::vector<at::Tensor>& in, int64_t dim) { return opt(g_concat(in, dim)); });
    m.def("g_contiguous", [opt](const at::Tensor& x) { return opt(g_contiguous(x)); });
    m.def("g_cast", [opt](const at::Tensor& x, const std::string& dtype) { return opt(g_cast(x, to_aten_dtype(dtype_from_name(dtype)))); });
    m.def("g_full", [opt](const std::vector<int64_t>& shape, const std::string& dtype, double value, const std::string& device) {
      return opt(g_full(shape, at::TensorOptions().dtype(to_aten_dtype(dtype_from_name(dtype))).device(c10::Device(device)), value));
    }, py::arg("shape"), py::arg("dtype"), py::arg("value"), py::arg("device") = "cuda");
  }

  // direct kernel entry points (benchmarks / numerics tests)
  m.def("gemm", [](const at::Tensor& a, const at::Tensor& b, bool a_mn, bool b_mn, const py::object& bias, const std::string& act,
                   bool out_fp32, int cta_group, int block_n) {
    HB_CHECK(a.is_cuda() && a.scalar_type() == at::kBFloat16 && a.dim() == 2 && b.dim() == 2) << "gemm expects 2-D CUDA bf16";
    const int64_t M = a_mn ? a.size(1) : a.size(0), K = a_mn ? a.size(0) : a.size(1), N = b_mn ? b.size(1) : b.size(0);
    at::Tensor c = at::empty({M, N}, a.options().dtype(out_fp32 ? at::kFloat : at::kBFloat16));
    GemmCall g;
    g.A = a.data_ptr(); g.B = b.data_ptr(); g.C = c.data_ptr();
    g.M = (int)M; g.N = (int)N; g.K = (int)K;
    g.lda = a.stride(0); g.ldb = b.stride(0); g.ldc = N;
    g.a_mn_major = a_mn; g.b_mn_major = b_mn;
    g.out = out_fp32 ? GemmOut::FP32 : GemmOut::BF16;
    at::Tensor bt;
    if (!bias.is_none()) { bt = py::cast<at::Tensor>(bias); g.bias = bt.data_ptr(); }
    g.act = act == "gelu" ? 1 : act == "relu" ? 2 : act == "gelu_tanh" ? 3 : act == "silu" ? 4 : 0;
    g.cta_group = cta_group;
    g.block_n = block_n;
    cuda_ok(gemm_bf16(g, cur_stream()), "gemm");
    return c;
  }, py::arg("a"), py::arg("b"), py::arg("a_mn_major") = false, py::arg("b_mn_major") = false, py::arg("bias") = py::none(),
     py::arg("act") = "none", py::arg("out_fp32") = false, py::arg("cta_group") = 0, py::arg("block_n") = 0);
}
