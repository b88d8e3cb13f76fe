// ZeRO ("OSDP" sharded data parallelism) over NVLink symmetric memory, fused with the compute it follows:
//   * weight-gradient GEMMs store every output tile straight into the staging slot of the rank that owns those rows
//     (peer stores from the tcgen05 GEMM epilogue): the gradient reduce-scatter rides inside backward, tile by tile;
//   * one kernel per parameter then sums the slots, applies AdamW to this rank's fp32 master shard and stores the
//     new bf16 shard into EVERY rank's parameter tensor (peer stores): reduce + optimizer + all-gather in one pass;
//   * small / irregular parameters share ONE flat in-kernel all-reduce instead of a collective each.
// The reference runs these as separate NCCL reduce-scatter / all-gather calls around the optimizer
// (hetu/graph/executable_graph.cc: grad reduce comm ops + optimize-compute bridge, hetu/graph/ops/Optimizer*.cc).
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <memory>
#include <set>
#include <unordered_map>
#include <vector>

#include "exec.h"

namespace hb {

struct SymmBuffer;

struct ZeroEntry {
  OpDef* update = nullptr;      // adam_update op
  OpDef* comm = nullptr;        // deferred reduce-scatter comm op feeding it
  OpDef* wgrad = nullptr;       // linear_wgrad producing the raw gradient (fused entries only)
  TensorId param = -1, raw_grad = -1;
  int64_t rows = 0, cols = 0, numel = 0;
  size_t slots_off = 0;         // fused: [world, rows / world, cols] bf16 staging slots in the arena
  size_t param_off = 0;         // bf16 parameter inside the arena
  size_t flat_off = 0;          // leftover: element offset inside the flat gradient buffer
  bool fused = false;
  int ready_pos = -1;           // NVLS: executor position (fw ops + bw index) after which gradient AND last use of the parameter are done
};

struct ZeroFusedState {
  bool ok = false;
  std::vector<int> ranks;
  int world = 1, pos = 0;
  std::string arena_name;
  std::vector<ZeroEntry> entries;
  std::unordered_map<OpId, size_t> by_wgrad;    // wgrad op id -> entry
  std::set<OpId> handled_ops;                    // update + comm ops executed by this path in the CURRENT run
  size_t flat_off = 0, flat_elems = 0;           // leftover flat gradient buffer (2x: all-reduce scratch)
  at::Tensor step_table;                         // device int64*[] of all step counters
  bool epilogue_this_run = false;
  // ---- NVLS mode (multicast mapping of the arena): every entry has a full local bf16 gradient region at `slots_off`;
  // the update of an entry (switch-side reduce + AdamW + multicast store, optim.cu adam_nvls_kernel) is launched on a side
  // stream as soon as backward is done with its parameter, overlapping the rest of backward
  bool nvls = false;
  std::multimap<int, size_t> ready_at;           // executor position -> entries that become ready there
  std::vector<char> launched;                    // per entry: update already issued in this run
  std::vector<size_t> ready_queue;               // ready entries waiting for their bucket's barrier
  cudaStream_t side = nullptr;
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
  uint32_t side_epoch = 0;                       // generation of the side-stream barrier (flag slot 1)
  double scale_this_run = 1.0;
};

}  // namespace hb
