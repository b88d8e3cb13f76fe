#include "zero_fused.h"

#include <ATen/cuda/CUDAContext.h>

#include "../runtime/symm_mem.h"
#include "op_utils.h"

namespace hb {

void wgrad_to_peer_slots(const at::Tensor& dy, const at::Tensor& x, bool trans_b, void* local_c, void* const* peer_c, int world,
                         int my_rank, int64_t rows_per_rank);   // ops_nn.cc

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

std::shared_ptr<ZeroFusedState> Executor::zero_fused_prepare(ExecPlan& plan) {
  auto st = std::make_shared<ZeroFusedState>();
  auto& comm = CommRuntime::get();
  if (!comm.initialized() || comm.world() < 2 || !aten_device().is_cuda() || env_int("HETU_ZERO_FUSED", 1) == 0) return st;
  // ---- classify the optimizer updates
  for (OpDef* u : plan.update_ops) {
    if (!u->has_flag(kFlagOptimizerUpdate) || u->type != "adam_update" || u->inputs.size() < 6) continue;
    const Tensor& param = u->inputs[0];
    const Tensor& grad = u->inputs[1];
    if (!grad->producer || grad->producer->type != "comm") continue;
    auto cit = plan.comm.find(grad->producer->id);
    if (cit == plan.comm.end() || cit->second.type != CommType::REDUCE_SCATTER || cit->second.dim != 0) continue;
    const CommStep& cs = cit->second;
    if (cs.ranks.size() < 2 || cs.ranks.size() > 8) continue;
    if (st->ranks.empty()) st->ranks = cs.ranks;
    if (cs.ranks != st->ranks) continue;     // one data-parallel group per plan on this path
    at::Tensor p = get_param(param);
    at::Tensor master = get_param(u->inputs[5]);
    const int n = (int)cs.ranks.size();
    if (p.scalar_type() != at::kBFloat16 || master.scalar_type() != at::kFloat || master.numel() * n != p.numel()) continue;
    if ((p.numel() / n) % 8 != 0) continue;
    ZeroEntry e;
    e.update = u; e.comm = grad->producer; e.param = param->id; e.numel = p.numel();
    const Tensor& raw = e.comm->inputs[0];
    e.raw_grad = raw->id;
    if (p.dim() == 2 && raw->producer && raw->producer->type == "linear_wgrad" && raw->producer->attrs.b("trans_b", true) &&
        raw->consumers.size() == 1 && p.size(0) % (128 * n) == 0 && p.size(1) % 8 == 0) {
      e.fused = true;
      e.wgrad = raw->producer;
      e.rows = p.size(0); e.cols = p.size(1);
    }
    st->entries.push_back(e);
  }
  if (st->entries.empty()) return st;
  st->world = (int)st->ranks.size();
  st->pos = -1;
  for (int i = 0; i < st->world; ++i) if (st->ranks[i] == comm.rank()) st->pos = i;
  if (st->pos < 0) return st;
  // ---- arena layout (identical on every rank: entries are in graph order)
  size_t off = 0;
  for (auto& e : st->entries) {
    if (e.fused) { e.slots_off = off; off = align_up(off + (size_t)e.numel * 2, 256); }
  }
  for (auto& e : st->entries) { e.param_off = off; off = align_up(off + (size_t)e.numel * 2, 256); }
  size_t flat = 0;
  for (auto& e : st->entries)
    if (!e.fused) { e.flat_off = flat; flat += (size_t)e.numel; flat = align_up(flat, 8 * (size_t)st->world); }
  st->flat_off = off;
  st->flat_elems = flat;
  off = align_up(off + flat * 2 * 2, 256);   // bf16, 2x for the two-shot all-reduce scratch
  // ---- allocate + exchange IPC handles over the data-parallel group
  static int arena_seq = 0;
  st->arena_name = "zero_arena_" + std::to_string(arena_seq++);
  auto& sm = SymmMem::get();
  std::string handle = sm.alloc(st->arena_name, off, st->pos, st->world);
  at::Tensor h = at::empty({(int64_t)handle.size()}, at::TensorOptions().dtype(at::kByte));
  std::memcpy(h.data_ptr(), handle.data(), handle.size());
  at::Tensor all = comm.all_gather(h.to(aten_device()), st->ranks, 0).cpu();
  std::vector<std::string> handles;
  for (int r = 0; r < st->world; ++r)
    handles.emplace_back(reinterpret_cast<const char*>(all.data_ptr()) + (size_t)r * handle.size(), handle.size());
  sm.open(st->arena_name, handles);
  SymmBuffer& buf = sm.buffer(st->arena_name);
  // ---- move the parameters into the arena (views share the symmetric allocation)
  auto& store = g_->param_data();
  std::vector<int64_t*> step_ptrs;
  for (auto& e : st->entries) {
    at::Tensor old = store[e.param];
    at::Tensor view = at::from_blob(static_cast<char*>(buf.local) + e.param_off, old.sizes(), old.options());
    view.copy_(old);
    store[e.param] = view;
    at::Tensor step = get_param(e.update->inputs[4]);
    if (!step.is_cuda()) { step = step.to(aten_device()); store[e.update->inputs[4]->id] = step; }
    step_ptrs.push_back(step.data_ptr<int64_t>());
    if (e.fused) st->by_wgrad[e.wgrad->id] = &e - &st->entries[0];
  }
  at::Tensor tab = at::empty({(int64_t)step_ptrs.size()}, at::TensorOptions().dtype(at::kLong));
  std::memcpy(tab.data_ptr(), step_ptrs.data(), step_ptrs.size() * sizeof(int64_t*));
  st->step_table = tab.to(aten_device());
  cuda_ok(cudaStreamSynchronize(cur_stream()), "zero arena setup");
  comm.barrier();
  st->ok = true;
  return st;
}

// backward hook: run the weight-gradient GEMM with the peer-store epilogue (returns true when it took the op)
bool Executor::zero_fused_wgrad(ZeroFusedState& st, OpDef* op, const std::vector<at::Tensor>& ins) {
  auto it = st.by_wgrad.find(op->id);
  if (it == st.by_wgrad.end() || !st.epilogue_this_run) return false;
  ZeroEntry& e = st.entries[it->second];
  SymmBuffer& buf = SymmMem::get().buffer(st.arena_name);
  void* peers[kMaxPeers];
  for (int r = 0; r < st.world; ++r) peers[r] = static_cast<char*>(buf.peer[r]) + e.slots_off;
  wgrad_to_peer_slots(ins[0], ins[1], true, static_cast<char*>(buf.local) + e.slots_off, peers, st.world, st.pos, e.rows / st.world);
  return true;
}

void Executor::zero_fused_update(ExecPlan& plan, ZeroFusedState& st, double scale) {
  (void)plan;
  st.handled_ops.clear();
  SymmBuffer& buf = SymmMem::get().buffer(st.arena_name);
  cudaStream_t s = cur_stream();
  auto& store = g_->param_data();
  char* base = static_cast<char*>(buf.local);
  // ---- leftovers (and everything when the epilogue path was not used): gather the gradients into the flat buffer
  bool any_flat = false;
  at::Tensor flat = at::from_blob(base + st.flat_off, {(int64_t)st.flat_elems}, at::TensorOptions().dtype(at::kBFloat16).device(aten_device()));
  std::vector<char> used_slots(st.entries.size(), 0);
  for (size_t i = 0; i < st.entries.size(); ++i) {
    ZeroEntry& e = st.entries[i];
    auto acc = accum_grads_.find(e.param);
    if (e.fused && st.epilogue_this_run && acc == accum_grads_.end()) { used_slots[i] = 1; continue; }
    if (e.fused) continue;   // fused entry with a locally accumulated gradient: handled by the generic path below
    if (acc == accum_grads_.end()) continue;
    flat.narrow(0, (int64_t)e.flat_off, e.numel).copy_(acc->second.reshape({-1}));
    any_flat = true;
  }
  if (any_flat)
    cuda_ok(symm_all_reduce(buf, st.flat_off, st.flat_elems, /*bf16=*/true, s), "zero flat all-reduce");
  else
    cuda_ok(symm_barrier(buf, s), "zero barrier");     // every rank's weight-gradient tiles have landed in my slots
  // ---- one fused kernel per parameter: slot reduce + AdamW on the shard + all-gather by peer stores
  for (size_t i = 0; i < st.entries.size(); ++i) {
    ZeroEntry& e = st.entries[i];
    const bool from_slots = used_slots[i];
    if (e.fused && !from_slots) continue;
    if (!e.fused && accum_grads_.find(e.param) == accum_grads_.end()) continue;
    OpDef* u = e.update;
    at::Tensor m = get_param(u->inputs[2]), v = get_param(u->inputs[3]), step = get_param(u->inputs[4]), master = get_param(u->inputs[5]);
    at::Tensor p = store[e.param];
    if (p.data_ptr() != base + e.param_off) {   // the parameter was replaced (checkpoint load): re-home it in the arena
      at::Tensor view = at::from_blob(base + e.param_off, p.sizes(), p.options());
      view.copy_(p);
      store[e.param] = view;
    }
    const int64_t shard = e.numel / st.world;
    AdamZeroArgs a;
    a.master = master.data_ptr<float>(); a.m = m.data_ptr<float>(); a.v = v.data_ptr<float>();
    if (from_slots) {
      a.slots = base + e.slots_off; a.nslots = st.world; a.slot_stride = shard;
    } else {
      a.slots = base + st.flat_off + ((size_t)e.flat_off + (size_t)st.pos * shard) * 2; a.nslots = 1; a.slot_stride = 0;
    }
    for (int r = 0; r < st.world; ++r) a.peer_param[r] = static_cast<char*>(buf.peer[r]) + e.param_off + (size_t)st.pos * shard * 2;
    a.world = st.world; a.n = shard;
    a.lr = (float)u->attrs.f("lr", 1e-3); a.beta1 = (float)u->attrs.f("beta1", 0.9); a.beta2 = (float)u->attrs.f("beta2", 0.999);
    a.eps = (float)u->attrs.f("eps", 1e-8); a.weight_decay = (float)u->attrs.f("weight_decay", 0.0);
    a.grad_scale = (float)scale;
    a.step_ptr = step.data_ptr<int64_t>(); a.step_add = 1;
    cuda_ok(adam_zero_fused(a, s), "adam_zero_fused");
    st.handled_ops.insert(u->id);
    st.handled_ops.insert(e.comm->id);
  }
  if (st.handled_ops.size() == 2 * st.entries.size()) {
    cuda_ok(increment_many_i64(reinterpret_cast<int64_t* const*>(st.step_table.data_ptr()), (int)st.entries.size(), s), "step counters");
  } else {
    for (auto& e : st.entries)
      if (st.handled_ops.count(e.update->id)) cuda_ok(increment_step(get_param(e.update->inputs[4]).data_ptr<int64_t>(), s), "step counter");
  }
  // nobody may start the next forward (or overwrite gradient slots) before every peer finished writing parameters
  cuda_ok(symm_barrier(buf, s), "zero barrier");
}

}  // namespace hb
