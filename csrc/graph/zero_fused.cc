#include "zero_fused.h"

#include <ATen/cuda/CUDAContext.h>

#include "../runtime/runtime.h"
#include "../runtime/symm_mem.h"
#include "op_utils.h"

namespace hb {

void wgrad_to_peer_slots(const at::Tensor& dy, const at::Tensor& x, bool trans_b, void* local_c, void* const* peer_c, int world,
                         int my_rank, int64_t rows_per_rank);   // ops_nn.cc

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// barrier among the ranks of ONE ZeRO group.  (A world barrier here dead-locks pipeline-parallel jobs: the stages prepare
// their groups at different times -- or not at all when a stage has nothing to shard -- while the other stage already
// waits in a p2p transfer; observed as the tp2 x pp2 hang on 4 GPUs.)
static void group_barrier(const std::vector<int>& ranks) {
  auto& comm = CommRuntime::get();
  at::Tensor t = at::zeros({1}, at::TensorOptions().dtype(at::kFloat).device(aten_device()));
  t = comm.all_reduce(t, ranks, ReductionType::SUM);
  if (t.is_cuda()) cuda_ok(cudaStreamSynchronize(cur_stream()), "group barrier");
}

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
  // NVLS: allocate through the VMM API and add a multicast mapping when the devices support it (csrc/runtime/symm_vmm.cc)
  const bool want_vmm = SymmMem::multicast_supported() && env_int("HETU_ZERO_NVLS", 1) != 0;
  // ---- arena layout (identical on every rank: entries are in graph order)
  // gradient regions: [world, rows / world, cols] staging slots of the peer-store path, or (NVLS) the full local gradient
  size_t off = 0;
  for (auto& e : st->entries) {
    if (e.fused || want_vmm) { e.slots_off = off; off = align_up(off + (size_t)e.numel * 2, 256); }
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
  std::string handle = want_vmm ? sm.alloc_vmm(st->arena_name, off, st->pos, st->world) : sm.alloc(st->arena_name, off, st->pos, st->world);
  at::Tensor h = at::empty({(int64_t)handle.size()}, at::TensorOptions().dtype(at::kByte));
  std::memcpy(h.data_ptr(), handle.data(), handle.size());
  at::Tensor all = comm.all_gather(h.to(aten_device()), st->ranks, 0).cpu();
  std::vector<std::string> handles;
  for (int r = 0; r < st->world; ++r)
    handles.emplace_back(reinterpret_cast<const char*>(all.data_ptr()) + (size_t)r * handle.size(), handle.size());
  if (want_vmm) {
    sm.open_vmm(st->arena_name, handles);
    group_barrier(st->ranks);                     // every device joined the multicast object before memory is bound
    sm.bind_multicast(st->arena_name);
    cuda_ok(cudaDeviceSynchronize(), "multicast bind");
  } else sm.open(st->arena_name, handles);
  SymmBuffer& buf = sm.buffer(st->arena_name);
  st->nvls = buf.mc != nullptr;
  if (st->nvls) {
    // when is backward done with a parameter?  after its weight-gradient GEMM and after its last reader (dgrad); layers
    // that are recomputed in backward read their parameters at unknown positions: those runs update at the end instead
    std::unordered_map<OpId, int> bw_pos;
    for (size_t i = 0; i < plan.bw_ops.size(); ++i) bw_pos[plan.bw_ops[i]->id] = (int)(plan.fw_ops.size() + i);
    const bool overlap = plan.recompute_ops.empty() && env_int("HETU_ZERO_OVERLAP", 1) != 0;
    for (size_t i = 0; i < st->entries.size(); ++i) {
      ZeroEntry& e = st->entries[i];
      if (!e.fused || !overlap) continue;
      auto wp = bw_pos.find(e.wgrad->id);
      if (wp == bw_pos.end()) continue;
      int pos = wp->second;
      auto lu = plan.last_use_fw.find(e.param);
      if (lu != plan.last_use_fw.end()) pos = std::max(pos, lu->second);
      e.ready_pos = pos;
      st->ready_at.emplace(pos, i);
    }
    // the ZeRO bridge work (gradient reduce + optimizer + parameter broadcast) runs on the device's logical bridge stream
    // (ref: hetu/core/stream.h kBridgeStream -- "ZeRO param-gather / grad-reduce")
    int cur_dev = 0;
    cuda_ok(cudaGetDevice(&cur_dev), "cudaGetDevice");
    st->side = logical_stream(cur_dev, kBridgeStream);
    cuda_ok(cudaEventCreateWithFlags(&st->ev_ready, cudaEventDisableTiming), "event");
    cuda_ok(cudaEventCreateWithFlags(&st->ev_done, cudaEventDisableTiming), "event");
  }
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
  group_barrier(st->ranks);
  st->ok = true;
  return st;
}

// backward hook: run the weight-gradient GEMM with the peer-store epilogue (returns true when it took the op)
bool Executor::zero_fused_wgrad(ZeroFusedState& st, OpDef* op, const std::vector<at::Tensor>& ins) {
  auto it = st.by_wgrad.find(op->id);
  if (it == st.by_wgrad.end() || !st.epilogue_this_run) return false;
  ZeroEntry& e = st.entries[it->second];
  SymmBuffer& buf = SymmMem::get().buffer(st.arena_name);
  if (st.nvls) {
    // plain epilogue into this rank's own gradient region; the switch sums the ranks' regions when the shard is loaded
    wgrad_to_peer_slots(ins[0], ins[1], true, static_cast<char*>(buf.local) + e.slots_off, nullptr, 1, 0, e.rows);
    return true;
  }
  void* peers[kMaxPeers];
  for (int r = 0; r < st.world; ++r) peers[r] = static_cast<char*>(buf.peer[r]) + e.slots_off;
  wgrad_to_peer_slots(ins[0], ins[1], true, static_cast<char*>(buf.local) + e.slots_off, peers, st.world, st.pos, e.rows / st.world);
  return true;
}

// NVLS: one entry = barrier-free kernel on stream `s` (the caller has synchronised the ranks for this entry's gradient)
void Executor::zero_nvls_launch(ZeroFusedState& st, size_t idx, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ZeroEntry& e = st.entries[idx];
  SymmBuffer& buf = SymmMem::get().buffer(st.arena_name);
  auto& store = g_->param_data();
  char* base = static_cast<char*>(buf.local);
  char* mc = static_cast<char*>(buf.mc);
  OpDef* u = e.update;
  at::Tensor m = get_param(u->inputs[2]), v = get_param(u->inputs[3]), step = get_param(u->inputs[4]), master = get_param(u->inputs[5]);
  at::Tensor p = store[e.param];
  if (p.data_ptr() != base + e.param_off) {   // the parameter was replaced (checkpoint load): re-home it in the arena
    at::Tensor view = at::from_blob(base + e.param_off, p.sizes(), p.options());
    view.copy_(p);
    store[e.param] = view;
  }
  const int64_t shard = e.numel / st.world;
  AdamNvlsArgs a;
  a.master = master.data_ptr<float>(); a.m = m.data_ptr<float>(); a.v = v.data_ptr<float>();
  a.grad_mc = mc + e.slots_off + (size_t)st.pos * shard * 2;
  a.param_mc = mc + e.param_off + (size_t)st.pos * shard * 2;
  a.n = shard;
  a.lr = (float)u->attrs.f("lr", 1e-3); a.beta1 = (float)u->attrs.f("beta1", 0.9); a.beta2 = (float)u->attrs.f("beta2", 0.999);
  a.eps = (float)u->attrs.f("eps", 1e-8); a.weight_decay = (float)u->attrs.f("weight_decay", 0.0);
  a.grad_scale = (float)st.scale_this_run;
  a.step_ptr = step.data_ptr<int64_t>(); a.step_add = 1;
  cuda_ok(adam_zero_nvls(a, s), "adam_zero_nvls");
  st.launched[idx] = 1;
}

// called by run_ops after every backward op: parameters whose gradient is complete and that backward no longer reads get
// their reduce + AdamW + broadcast kernel NOW, on the side stream, behind a cross-rank barrier of that stream
void Executor::zero_nvls_after_op(ExecPlan& plan, ZeroFusedState& st, int pos) {
  (void)plan;
  auto range = st.ready_at.equal_range(pos);
  if (range.first == range.second || !st.epilogue_this_run) return;
  for (auto it = range.first; it != range.second; ++it)
    if (!st.launched[it->second]) st.ready_queue.push_back(it->second);
  // Buckets: the cross-rank barrier of the bridge stream is a spinning one-CTA kernel; while it waits for the slowest
  // rank it holds a CTA slot and the next persistent GEMM of the compute stream cannot place its CTA on that SM.  One
  // barrier per parameter (97 per step) cost ~4 ms of such stalls at 4 GPUs; with a bucket of parameters per barrier the
  // barrier is also issued well after the first member became ready, so the peers have usually passed it already.
  static const int bucket = std::max(1, (int)env_int("HETU_ZERO_BUCKET", 24));
  if ((int)st.ready_queue.size() < bucket) return;
  SymmBuffer& buf = SymmMem::get().buffer(st.arena_name);
  cuda_ok(cudaEventRecord(st.ev_ready, cur_stream()), "event record");
  cuda_ok(cudaStreamWaitEvent(st.side, st.ev_ready, 0), "stream wait");
  cuda_ok(symm_barrier_slot(buf, 1, ++st.side_epoch, st.side), "side barrier");      // all ranks are past this point
  for (size_t idx : st.ready_queue) zero_nvls_launch(st, idx, st.side);
  st.ready_queue.clear();
}

void Executor::zero_fused_update(ExecPlan& plan, ZeroFusedState& st, double scale) {
  (void)plan;
  st.handled_ops.clear();
  SymmBuffer& buf = SymmMem::get().buffer(st.arena_name);
  cudaStream_t s = cur_stream();
  auto& store = g_->param_data();
  char* base = static_cast<char*>(buf.local);
  if (st.nvls) {
    // entries that were not launched from inside backward: weight gradients produced without the epilogue hook and the
    // non-GEMM parameters (embeddings, norms, biases) -- copy the accumulated gradient into the arena, then one barrier
    // for all of them and the same kernel
    st.ready_queue.clear();          // what is still queued joins the final batch below
    std::vector<size_t> todo;
    for (size_t i = 0; i < st.entries.size(); ++i) {
      ZeroEntry& e = st.entries[i];
      if (st.launched[i]) continue;
      auto acc = accum_grads_.find(e.param);
      if (acc != accum_grads_.end()) {
        at::Tensor dst = at::from_blob(base + e.slots_off, {e.numel}, at::TensorOptions().dtype(at::kBFloat16).device(aten_device()));
        dst.copy_(acc->second.reshape({-1}));
      } else if (!(e.fused && st.epilogue_this_run)) continue;       // no gradient in this run
      todo.push_back(i);
    }
    cuda_ok(symm_barrier(buf, s), "zero barrier");               // every rank's remaining gradients are in place
    for (size_t i : todo) zero_nvls_launch(st, i, s);
    // join the side stream (updates issued during backward), then nobody may start the next forward before every peer
    // finished multicasting parameters
    cuda_ok(cudaEventRecord(st.ev_done, st.side), "event record");
    cuda_ok(cudaStreamWaitEvent(s, st.ev_done, 0), "stream wait");
    size_t done = 0;
    for (size_t i = 0; i < st.entries.size(); ++i) {
      if (!st.launched[i]) continue;
      ++done;
      st.handled_ops.insert(st.entries[i].update->id);
      st.handled_ops.insert(st.entries[i].comm->id);
    }
    if (done == st.entries.size()) {
      cuda_ok(increment_many_i64(reinterpret_cast<int64_t* const*>(st.step_table.data_ptr()), (int)st.entries.size(), s), "step counters");
    } else {
      for (size_t i = 0; i < st.entries.size(); ++i)
        if (st.launched[i]) cuda_ok(increment_step(get_param(st.entries[i].update->inputs[4]).data_ptr<int64_t>(), s), "step counter");
    }
    cuda_ok(symm_barrier(buf, s), "zero barrier");
    return;
  }
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
