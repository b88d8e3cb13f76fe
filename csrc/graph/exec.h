// Executor: lowers a define-and-run graph into a per-rank plan (placement, comm
// substitution, pipeline schedule, gradient bridges) and runs it over micro-batches.
//
// B200-first choices: the plan is compiled once per (strategy, fetch set) and
// replayed (no per-step graph interpretation of placement / comm decisions);
// tensors are freed by static last-use analysis; collectives go through a
// CommRuntime that can route to NCCL process groups or to the symmetric-memory
// (peer load/store over NVLink) kernels; parameters / gradients of the sharded-DP
// path live in flat buffers so ZeRO all-gather / reduce-scatter are single bucketed
// collectives overlapped with compute.
// (capability parity: hetu/graph/executable_graph.{h,cc}, define_and_run_graph.cc,
//  pre_post_exec_graph.cc)
#pragma once
#include <ATen/ATen.h>
#include <torch/csrc/distributed/c10d/ProcessGroup.hpp>

#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../core/symbol.h"
#include "ir.h"

namespace hb {

// ------------------------------------------------------------------ communication runtime
using PG = c10::intrusive_ptr<c10d::ProcessGroup>;

class CommRuntime {
 public:
  static CommRuntime& get();
  void init(int rank, int world, PG world_pg, std::function<PG(const std::vector<int>&)> group_factory);
  bool initialized() const { return world_ > 0 && world_pg_; }
  int rank() const { return rank_; }
  int world() const { return world_ <= 0 ? 1 : world_; }
  // process group over `ranks` (must be requested in the same order on every rank)
  PG group(const std::vector<int>& ranks);
  void barrier();

  // collectives on rank-local tensors; `ranks` lists the participating global ranks
  at::Tensor all_reduce(const at::Tensor& x, const std::vector<int>& ranks, ReductionType red = ReductionType::SUM,
                        bool fp32_reduce = false);
  at::Tensor all_gather(const at::Tensor& x, const std::vector<int>& ranks, int dim);
  at::Tensor reduce_scatter(const at::Tensor& x, const std::vector<int>& ranks, int dim,
                            ReductionType red = ReductionType::SUM, bool fp32_reduce = false);
  at::Tensor broadcast(const at::Tensor& x, const std::vector<int>& ranks, int root_rank);
  at::Tensor all_to_all(const at::Tensor& x, const std::vector<int>& ranks, int split_dim = 0, int concat_dim = 0);
  // bucketed all-reduce: the tensors are packed into one flat buffer, reduced in ONE collective and unpacked -- many small
  // gradients cost one launch latency instead of one each (ref: NCCLCommunicationGroupDef::AllReduceCoalesce,
  // hetu/impl/communication/nccl_comm_group.cu:246-313)
  std::vector<at::Tensor> all_reduce_coalesce(const std::vector<at::Tensor>& xs, const std::vector<int>& ranks,
                                              ReductionType red = ReductionType::SUM);
  // rooted collectives (ref: nccl_comm_group.cu Reduce :354, Gather / Scatter :465-563 -- send/recv fans there)
  at::Tensor reduce(const at::Tensor& x, const std::vector<int>& ranks, int root_rank, ReductionType red = ReductionType::SUM);
  at::Tensor gather(const at::Tensor& x, const std::vector<int>& ranks, int root_rank);      // root: [n, ...]; others: empty
  at::Tensor scatter(const at::Tensor& x, const std::vector<int>& ranks, int root_rank);     // root passes [n, ...], all get [...]
  // asynchronous gradient synchronisation: the collective is enqueued on the communication stream of the process group
  // (behind everything issued so far on the compute stream) and completed later by finish() -- the backward pass keeps
  // computing in between (ref: executable_graph.cc:1137-1150 overlapped grad reduce on kBridgeStream)
  struct AsyncResult {
    c10::intrusive_ptr<c10d::Work> work;
    at::Tensor out, keep;                 // result buffer, input kept alive until completion
    ReductionType red = ReductionType::SUM;
    int64_t n = 1;
    at::ScalarType want = at::kFloat;
    bool valid() const { return out.defined(); }
  };
  AsyncResult all_reduce_async(const at::Tensor& x, const std::vector<int>& ranks, ReductionType red = ReductionType::SUM,
                               bool fp32_reduce = false);
  AsyncResult reduce_scatter_async(const at::Tensor& x, const std::vector<int>& ranks, int dim,
                                   ReductionType red = ReductionType::SUM, bool fp32_reduce = false);
  at::Tensor finish(AsyncResult& r);
  // pipeline P2P: sends are asynchronous (completed by flush_sends()), the forward and backward directions use
  // separate channels (own communicator / FIFO) so a stage sending activations never blocks behind a gradient receive
  void send(const at::Tensor& x, int dst_rank, int channel = 0);
  at::Tensor recv(const std::vector<int64_t>& shape, at::ScalarType dtype, const at::Device& dev, int src_rank, int channel = 0);
  void flush_sends();
  PG p2p_group(int channel);
  // grouped point-to-point (ring attention, re-sharding): all sends / recvs are issued as one batch
  void batched_send_recv(const std::vector<std::pair<at::Tensor, int>>& sends,
                         std::vector<std::pair<at::Tensor, int>>& recvs);

  // statistics (bytes per collective kind) for the step profiler
  std::map<std::string, int64_t>& bytes() { return bytes_; }
  std::map<std::string, int64_t>& calls() { return calls_; }

 private:
  int rank_ = 0, world_ = 0;
  PG world_pg_;
  std::function<PG(const std::vector<int>&)> factory_;
  std::map<std::vector<int>, PG> groups_;
  std::map<std::string, int64_t> bytes_, calls_;
  PG p2p_pg_[2];
  std::vector<std::pair<c10::intrusive_ptr<c10d::Work>, at::Tensor>> pending_sends_;
};

// ------------------------------------------------------------------ run context
class Executor;
struct RunCtx {
  Graph* graph = nullptr;
  Executor* exec = nullptr;
  int micro_batch = 0;
  int num_micro_batches = 1;
  int strategy = 0;
  bool training = true;
  uint64_t seed = 0;
  std::unordered_map<std::string, at::Tensor>* workspace = nullptr;  // persistent scratch keyed by name
  // optimizer step counters whose increment is deferred to ONE batched kernel at the end of the update phase
  std::vector<at::Tensor>* deferred_steps = nullptr;
  at::Tensor scratch(const std::string& key, std::vector<int64_t> shape, at::ScalarType dt, const at::Device& dev);
};

// ------------------------------------------------------------------ pipeline schedules
struct PipeTask {
  enum Kind : int { FORWARD = 0, BACKWARD = 1, FLUSH = -1 } kind;
  int micro_batch;
};
// (ref: executable_graph.cc:803 GenerateGpipeSchedule, :836 GeneratePipedreamFlushSchedule)
std::vector<std::vector<PipeTask>> generate_gpipe_schedule(int num_stages, int num_micro_batches, bool inference);
std::vector<std::vector<PipeTask>> generate_1f1b_schedule(int num_stages, int num_micro_batches, bool inference);

// ------------------------------------------------------------------ executor
struct CommStep {
  CommType type = CommType::UNUSED;
  std::vector<int> ranks;      // participating global ranks (collectives)
  int dim = 0;                 // gather / scatter dim
  int peer = -1;               // P2P peer rank
  bool is_sender = false, is_receiver = false;
  std::vector<TransferItem> transfers;   // BATCHED_ISEND_IRECV
  std::vector<int64_t> global_shape;
  int split_index = 0, split_num = 1;    // COMM_SPLIT / SCATTER
};

struct ExecPlan {
  int strategy = 0;
  std::vector<OpDef*> fw_ops, bw_ops, update_ops;   // local ops in execution order
  std::unordered_map<OpId, CommStep> comm;          // lowered comm ops
  int num_stages = 1, stage = 0;
  std::vector<DeviceGroup> stage_groups;
  std::unordered_map<TensorId, int> last_use_fw, last_use_bw;   // position of last consumer in fw/bw lists
  std::set<OpId> recompute_ops;                                  // forward ops re-executed on demand in backward
  std::vector<TensorId> fetch_ids;
  std::unordered_map<TensorId, TensorId> param_of_grad;    // grad tensor id -> param tensor id
  std::unordered_map<TensorId, OpDef*> update_of_param;
  std::unordered_map<TensorId, OpDef*> deferred_comm_of_raw;   // raw gradient tensor -> its deferred (update-phase) comm op
  bool built = false;
};

struct RunOptions {
  int num_micro_batches = 1;
  int strategy = 0;
  RunLevel run_level = RunLevel::UPDATE;
  double grad_scale = 1.0;
  bool save_checkpoint = false;
  // per-micro-batch values of symbolic ints (sequence length of each micro-batch, ...): symbols[i].second[mb]
  std::vector<std::pair<IntSymbol, std::vector<int64_t>>> symbols;
};

struct ZeroFusedState;
struct TpFusedState;
at::Device aten_device();   // device this process computes on (cuda:LOCAL_RANK, or cpu)

class Executor {
 public:
  explicit Executor(Graph* g) : g_(g) {}
  // feed: tensor id -> one tensor per micro-batch (or a single tensor that is split along dim 0)
  std::vector<at::Tensor> run(const Tensor& loss, const TensorList& fetches,
                              const std::unordered_map<TensorId, std::vector<at::Tensor>>& feed, const RunOptions& opt);
  // parameter / optimizer-state access (checkpointing, hot switching)
  at::Tensor get_param(const Tensor& t);
  void set_param(const Tensor& t, const at::Tensor& v);
  void ensure_param(OpDef* var_op, int strategy);
  std::unordered_map<std::string, at::Tensor>& workspace() { return workspace_; }
  // accumulated gradients (RunLevel::GRAD keeps them across run() calls until an UPDATE run)
  std::unordered_map<TensorId, at::Tensor>& accumulated_grads() { return accum_grads_; }
  // per-op timing of the last run (ms), filled when profiling is enabled
  // HETU_EVENT_TIMING=OFF vetoes per-op timing (it synchronises after every op)
  void set_profile(bool on) { profile_ = on && env_str("HETU_EVENT_TIMING", "ON") != "OFF"; }
  const std::vector<std::pair<std::string, double>>& op_times() const { return op_times_; }
  std::map<std::string, double> step_breakdown() const { return breakdown_; }
  int local_device_index(const DeviceGroup& g) const;
  // same, but a rank outside the group is an error that names the op and the group (instead of an index of -1 travelling on)
  int require_device_index(const DeviceGroup& g, const OpDef* op, const char* what) const;
  Device local_device() const;
  // hot switch: re-shard every parameter / optimizer state from strategy a to b
  void switch_strategy(int from, int to);
  int active_strategy() const { return active_strategy_; }
  // dynamic loss scaling (ref: hetu/graph/autocast/gradscaler.h, optimizer_update.cc SGDUpdateWithGradScaler): the loss is
  // multiplied by the value of `scale_var` inside the graph; the update phase un-scales the accumulated gradients, checks
  // them for inf/nan on every rank (MAX all-reduce), skips the optimizer step and backs the scale off when one is found,
  // and grows the scale after `growth_interval` clean steps.
  struct LossScaler {
    bool enabled = false;
    Tensor scale_var;
    double scale = 65536.0, growth = 2.0, backoff = 0.5;
    int64_t interval = 2000, tracker = 0, skipped = 0;
    bool last_found_inf = false;
  };
  void set_loss_scaler(const Tensor& scale_var, double init_scale, double growth, double backoff, int64_t interval);
  LossScaler& loss_scaler() { return scaler_; }

 private:
  ExecPlan& get_plan(const Tensor& loss, const TensorList& fetches, int strategy);
  void build_plan(ExecPlan& plan, const Tensor& loss, const TensorList& fetches, int strategy);
  void lower_comm(ExecPlan& plan, OpDef* op, int strategy);
  void run_ops(ExecPlan& plan, const std::vector<OpDef*>& ops, bool backward, int mb, RunCtx& rc,
               std::unordered_map<TensorId, at::Tensor>& vals);
  std::vector<at::Tensor> exec_comm(const CommStep& cs, OpDef* op, const std::vector<at::Tensor>& in, RunCtx& rc);
  void recompute_tensor(ExecPlan& plan, const Tensor& t, RunCtx& rc, std::unordered_map<TensorId, at::Tensor>& vals);
  void offload_activations(ExecPlan& plan, std::unordered_map<TensorId, at::Tensor>& vals);
  // ZeRO over symmetric memory fused with wgrad GEMMs and the optimizer (zero_fused.cc)
  std::shared_ptr<ZeroFusedState> zero_fused_prepare(ExecPlan& plan);
  bool zero_fused_wgrad(ZeroFusedState& st, OpDef* op, const std::vector<at::Tensor>& ins);
  void zero_fused_update(ExecPlan& plan, ZeroFusedState& st, double scale);
  void zero_nvls_after_op(ExecPlan& plan, ZeroFusedState& st, int pos);   // launches the updates that became ready at `pos`
  void zero_nvls_launch(ZeroFusedState& st, size_t entry, void* stream);   // stream: cudaStream_t
  // tensor-parallel GEMM -> reduce-scatter over symmetric memory (tp_fused.cc)
  void tp_fused_scan(ExecPlan& plan);
  bool tp_fused_gemm(ExecPlan& plan, OpDef* op, const std::vector<at::Tensor>& ins, RunCtx& rc, std::vector<at::Tensor>& outs);
  bool tp_fused_comm(ExecPlan& plan, OpDef* op, const std::vector<at::Tensor>& ins, std::vector<at::Tensor>& outs);

  Graph* g_;
  std::map<std::pair<int, std::vector<TensorId>>, ExecPlan> plans_;
  std::unordered_map<std::string, at::Tensor> workspace_;
  std::unordered_map<TensorId, at::Tensor> accum_grads_;
  std::map<const ExecPlan*, std::shared_ptr<ZeroFusedState>> zero_fused_;
  std::map<const ExecPlan*, std::shared_ptr<TpFusedState>> tp_fused_;
  bool single_shot_grads_ = false;
  bool overlap_grad_reduce_ = false;
  std::map<const ExecPlan*, std::vector<int64_t>> step_tables_host_;
  ZeroFusedState* zf_active_ = nullptr;   // set while run() executes a plan on the fused ZeRO path
  std::vector<std::pair<std::string, double>> op_times_;
  std::map<std::string, double> breakdown_;
  bool profile_ = false;
  int active_strategy_ = -1;
  int shapes_strategy_ = -1;   // strategy whose local shapes are currently stored in the tensors
  uint64_t step_ = 0;
  LossScaler scaler_;
  // gradient collectives launched from inside backward (HETU_OVERLAP_GRAD_REDUCE), keyed by the comm op's output tensor
  std::unordered_map<TensorId, CommRuntime::AsyncResult> async_grad_comm_;
};

}  // namespace hb
