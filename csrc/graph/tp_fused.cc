// Tensor-parallel GEMM -> reduce-scatter fused over NVLink symmetric memory inside the executor.
// Pattern: comm[REDUCE_SCATTER dim 0 over the tp group](linear | linear_dgrad) -- the row-parallel forward GEMM and the
// column-parallel input gradient under sequence parallelism.  The GEMM epilogue stores every output tile into the
// staging slot of the rank that owns those rows (peer stores, overlapped with the MMA mainloop tile by tile); the comm
// op then only sums its `world` local slots.  Two staging buffers alternate so that a fast rank's next GEMM never
// overwrites slots a slow rank is still reducing.
// (reference: separate cuBLAS GEMM + ncclReduceScatter, hetu/graph/ops/Communication.cc + nccl_comm_group.cu)
#include <ATen/cuda/CUDAContext.h>

#include "../runtime/symm_mem.h"
#include "exec.h"
#include "op_utils.h"

namespace hb {

thread_local GemmSink* tls_gemm_sink = nullptr;
thread_local GemmGather* tls_gemm_gather = nullptr;

struct TpFusedState {
  std::unordered_map<OpId, OpId> comm_of_gemm;     // producer GEMM op -> its reduce-scatter comm op
  std::unordered_map<OpId, OpId> gemm_of_comm;
  struct Staging { std::string name[2]; size_t bytes = 0; int next = 0; };
  std::map<std::vector<int>, Staging> staging;     // per tp group
  std::unordered_map<OpId, std::pair<std::string, std::vector<int64_t>>> pending;   // comm op -> (buffer, [rows_per_rank, cols])
  // fused all-gather -> GEMM: all-gather comm op -> the forward linear that consumes it first
  std::unordered_map<OpId, OpId> gemm_of_gather;
  std::map<std::vector<int>, Staging> ag_staging;
  std::unordered_map<OpId, GemmGather> pending_gather;    // linear op -> gather descriptor prepared by its comm op
  std::unordered_map<OpId, at::Tensor> gather_flags;
};

void symm_exchange_and_open(const std::string& name, size_t bytes, const std::vector<int>& ranks, int pos) {
  auto& sm = SymmMem::get();
  auto& comm = CommRuntime::get();
  std::string handle = sm.alloc(name, bytes, pos, (int)ranks.size());
  at::Tensor h = at::empty({(int64_t)handle.size()}, at::TensorOptions().dtype(at::kByte));
  std::memcpy(h.data_ptr(), handle.data(), handle.size());
  at::Tensor all = comm.all_gather(h.to(aten_device()), ranks, 0).cpu();
  std::vector<std::string> handles;
  for (size_t r = 0; r < ranks.size(); ++r)
    handles.emplace_back(reinterpret_cast<const char*>(all.data_ptr()) + r * handle.size(), handle.size());
  sm.open(name, handles);
}

void Executor::tp_fused_scan(ExecPlan& plan) {
  auto st = std::make_shared<TpFusedState>();
  tp_fused_[&plan] = st;
  auto& comm = CommRuntime::get();
  if (!comm.initialized() || comm.world() < 2 || !aten_device().is_cuda() || env_int("HETU_TP_FUSED", 1) == 0) return;
  auto scan = [&](const std::vector<OpDef*>& ops) {
    for (OpDef* c : ops) {
      if (c->type != "comm") continue;
      auto it = plan.comm.find(c->id);
      if (it == plan.comm.end() || it->second.type != CommType::REDUCE_SCATTER || it->second.dim != 0) continue;
      if (it->second.ranks.size() < 2 || it->second.ranks.size() > 8) continue;
      const Tensor& x = c->inputs[0];
      OpDef* p = x->producer;
      if (p == nullptr || x->consumers.size() != 1) continue;
      const bool fw = p->type == "linear" && !p->attrs.b("has_bias") && !p->attrs.b("has_residual") &&
                      (p->attrs.s("act").empty() || p->attrs.s("act") == "none");
      const bool bw = p->type == "linear_dgrad" && p->inputs.size() == 2;
      if (!fw && !bw) continue;
      if (x->dtype != DataType::BFLOAT16) continue;
      st->comm_of_gemm[p->id] = c->id;
      st->gemm_of_comm[c->id] = p->id;
    }
  };
  scan(plan.fw_ops);
  scan(plan.bw_ops);
  // all-gather (dim 0) whose first consumer in execution order is a forward linear taking it as the activation operand
  if (env_int("HETU_TP_FUSED_AG", 1) != 0) {
    std::unordered_map<OpId, int> pos;
    for (size_t i = 0; i < plan.fw_ops.size(); ++i) pos[plan.fw_ops[i]->id] = (int)i;
    for (OpDef* c : plan.fw_ops) {
      if (c->type != "comm") continue;
      auto it = plan.comm.find(c->id);
      if (it == plan.comm.end() || it->second.type != CommType::ALL_GATHER || it->second.dim != 0) continue;
      if (it->second.ranks.size() < 2 || it->second.ranks.size() > 8) continue;
      const Tensor& y = c->outputs[0];
      if (y->dtype != DataType::BFLOAT16 || y->shape.size() != 2) continue;
      OpDef* first = nullptr;
      int best = INT32_MAX;
      for (OpDef* u : y->consumers) {
        auto p = pos.find(u->id);
        if (p != pos.end() && p->second < best) { best = p->second; first = u; }
      }
      if (first == nullptr || first->type != "linear" || first->inputs[0]->id != y->id) continue;
      st->gemm_of_gather[c->id] = first->id;
    }
  }
}

// called instead of the plain compute for a GEMM op whose result feeds a fused reduce-scatter; returns false to decline
bool Executor::tp_fused_gemm(ExecPlan& plan, OpDef* op, const std::vector<at::Tensor>& ins, RunCtx& rc, std::vector<at::Tensor>& outs) {
  auto sit = tp_fused_.find(&plan);
  if (sit == tp_fused_.end() || !sit->second) return false;
  TpFusedState& st = *sit->second;
  auto git = st.pending_gather.find(op->id);
  if (git != st.pending_gather.end()) {
    // this linear's activation operand is gathered by the GEMM kernel itself (fused all-gather -> GEMM)
    GemmGather ag = git->second;
    st.pending_gather.erase(git);
    tls_gemm_gather = &ag;
    try {
      outs = op->kernel->compute(*op, ins, &rc);
    } catch (...) {
      tls_gemm_gather = nullptr;
      throw;
    }
    tls_gemm_gather = nullptr;
    HB_CHECK(ag.used) << "fused all-gather: " << op->name() << " did not consume its gather descriptor";
    return true;
  }
  auto pit = st.comm_of_gemm.find(op->id);
  if (pit == st.comm_of_gemm.end()) return false;
  const CommStep& cs = plan.comm[pit->second];
  const int world = (int)cs.ranks.size();
  // output geometry: [T, N] with T split into `world` row blocks of a multiple of 128 rows
  const at::Tensor& a = ins[0];
  if (!is_native(a)) return false;
  const int64_t T = a.numel() / a.size(-1);
  const Tensor& w = op->inputs[1];
  const bool trans_b = op->attrs.b("trans_b", true);
  const at::Tensor& wt = ins[1];
  const int64_t N = op->type == "linear" ? (trans_b ? wt.size(0) : wt.size(1)) : (trans_b ? wt.size(1) : wt.size(0));
  (void)w;
  if (T % world != 0 || (T / world) % 128 != 0 || N % 8 != 0) return false;
  int pos = -1;
  for (int i = 0; i < world; ++i) if (cs.ranks[i] == CommRuntime::get().rank()) pos = i;
  if (pos < 0) return false;
  auto& sg = st.staging[cs.ranks];
  const size_t need = (size_t)T * N * 2;
  if (sg.bytes < need) {
    // (re)allocate both staging buffers; collective over the tp group, every member reaches this op in the same order
    static int seq = 0;
    for (int k = 0; k < 2; ++k) {
      sg.name[k] = "tp_stage_" + std::to_string(seq++);
      symm_exchange_and_open(sg.name[k], need, cs.ranks, pos);
    }
    sg.bytes = need;
  }
  const std::string& name = sg.name[sg.next];
  sg.next ^= 1;
  SymmBuffer& buf = SymmMem::get().buffer(name);
  GemmSink sink;
  sink.local = buf.local;
  for (int r = 0; r < world; ++r) sink.peers[r] = buf.peer[r];
  sink.world = world; sink.my_rank = pos; sink.rows_per_rank = T / world; sink.cols = N; sink.used = false;
  tls_gemm_sink = &sink;
  try {
    outs = op->kernel->compute(*op, ins, &rc);
  } catch (...) {
    tls_gemm_sink = nullptr;
    throw;
  }
  tls_gemm_sink = nullptr;
  if (!sink.used) return true;          // shapes did not qualify inside the op: it ran normally, the comm op falls back
  st.pending[pit->second] = {name, {T / world, N}};
  return true;
}

// the reduce-scatter side: sum this rank's `world` staging slots (returns false when the producer did not use the sink)
bool Executor::tp_fused_comm(ExecPlan& plan, OpDef* op, const std::vector<at::Tensor>& ins, std::vector<at::Tensor>& outs) {
  auto sit = tp_fused_.find(&plan);
  if (sit == tp_fused_.end() || !sit->second) return false;
  TpFusedState& st = *sit->second;
  auto gt = st.gemm_of_gather.find(op->id);
  if (gt != st.gemm_of_gather.end() && !ins.empty() && is_native(ins[0]) && ins[0].dim() == 2) {
    const CommStep& cs = plan.comm[op->id];
    const int world = (int)cs.ranks.size();
    const at::Tensor x = ins[0].contiguous();
    const int64_t rows = x.size(0), K = x.size(1);
    int pos = -1;
    for (int i = 0; i < world; ++i) if (cs.ranks[i] == CommRuntime::get().rank()) pos = i;
    if (pos >= 0 && rows % 256 == 0 && K % 8 == 0) {
      auto& sg = st.ag_staging[cs.ranks];
      const size_t need = (size_t)rows * K * 2;
      if (sg.bytes < need) {
        static int seq = 0;
        for (int k = 0; k < 2; ++k) {
          sg.name[k] = "tp_ag_" + std::to_string(seq++);
          symm_exchange_and_open(sg.name[k], need, cs.ranks, pos);
        }
        sg.bytes = need;
      }
      SymmBuffer& buf = SymmMem::get().buffer(sg.name[sg.next]);
      sg.next ^= 1;
      cudaStream_t s = cur_stream();
      cuda_ok(cudaMemcpyAsync(buf.local, x.data_ptr(), need, cudaMemcpyDeviceToDevice, s), "publish shard");
      cuda_ok(symm_barrier(buf, s), "ag barrier");     // every rank's shard is readable
      at::Tensor out = at::empty({rows * world, K}, x.options());
      at::Tensor flags = at::zeros({rows * world / 128}, x.options().dtype(at::kInt));
      GemmGather g;
      for (int r = 0; r < world; ++r) g.src[r] = buf.peer[r];
      g.dst = out.data_ptr(); g.flags = reinterpret_cast<uint32_t*>(flags.data_ptr<int>());
      g.world = world; g.my_rank = pos; g.rows_per_rank = rows;
      st.pending_gather[gt->second] = g;
      st.gather_flags[gt->second] = flags;     // keeps the flag array alive until the next use of this linear
      outs = {out};
      return true;
    }
  }
  auto it = st.pending.find(op->id);
  if (it == st.pending.end()) return false;
  SymmBuffer& buf = SymmMem::get().buffer(it->second.first);
  const int64_t rows = it->second.second[0], cols = it->second.second[1];
  st.pending.erase(it);
  cudaStream_t s = cur_stream();
  cuda_ok(symm_barrier(buf, s), "tp fused barrier");     // every rank's partial tiles have landed in my slots
  at::Tensor out = at::empty({rows, cols}, at::TensorOptions().dtype(at::kBFloat16).device(aten_device()));
  cuda_ok(symm_reduce_slots(buf.local, buf.world, out.data_ptr(), nullptr, nullptr, rows, (int)cols, s), "reduce_slots");
  std::vector<int64_t> shape = op->outputs[0]->shape;
  outs = {shape.empty() ? out : out.reshape(shape)};
  return true;
}

}  // namespace hb
