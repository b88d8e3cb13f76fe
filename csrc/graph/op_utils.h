// Glue between at::Tensor and the raw-pointer kernel APIs, plus DS helpers shared by op definitions.
#pragma once
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <cuda_runtime.h>

#include "../kernels/attention_sm100.h"
#include "../kernels/gemm_sm100.h"
#include "../kernels/kernels.h"
#include "ir.h"

namespace hb {

// the hand-written sm_100a path is taken for CUDA bf16 tensors; anything else (CPU tests, fp32
// unit tests) goes through ATen with identical semantics
inline bool is_native(const at::Tensor& t) { return t.is_cuda() && t.scalar_type() == at::kBFloat16; }
inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }
inline void cuda_ok(cudaError_t e, const char* what) {
  HB_CHECK(e == cudaSuccess) << what << " failed: " << cudaGetErrorString(e);
}
// set HETU_B200_STRICT=1 (the GPU test suite does) to make any silent ATen fallback on a CUDA bf16 hot op an error
bool strict_native();
void note_fallback(const char* op);
int64_t fallback_count();

at::Tensor native_add(const at::Tensor& a, const at::Tensor& b);

// generic device kernels behind at::Tensor (native_generic.cc); an UNDEFINED result means "not applicable, use ATen"
at::Tensor g_contiguous(const at::Tensor& x);
at::Tensor g_unary(int op, const at::Tensor& x, float p0 = 0.f, float p1 = 0.f);
at::Tensor g_binary(int op, const at::Tensor& a, const at::Tensor& b);
at::Tensor g_reduce(int mode, const at::Tensor& x, std::vector<int64_t> axes, bool keepdims);
at::Tensor g_softmax(bool log, const at::Tensor& x, int64_t dim);
at::Tensor g_concat(const std::vector<at::Tensor>& in, int64_t dim);
at::Tensor g_cast(const at::Tensor& x, at::ScalarType to);
at::Tensor g_full(at::IntArrayRef shape, const at::TensorOptions& opt, double value);

// Destination override for the next GEMM issued on this thread: when set (by the executor's tensor-parallel fusion),
// a GEMM whose output is [world * rows_per_rank, cols] stores its tiles into the owner ranks' symmetric staging slots
// (peer stores from the epilogue) instead of its own output tensor.
struct GemmSink {
  void* local = nullptr;
  void* peers[8] = {nullptr};
  int world = 1, my_rank = 0;
  int64_t rows_per_rank = 0, cols = 0;
  bool used = false;
};
extern thread_local GemmSink* tls_gemm_sink;
// Source override for the next GEMM: its A operand (`dst`, a local [world * rows_per_rank, K] buffer) is all-gathered from
// the ranks' symmetric shards by the GEMM kernel itself while it computes (fused all-gather -> GEMM).
struct GemmGather {
  const void* src[8] = {nullptr};
  void* dst = nullptr;
  uint32_t* flags = nullptr;
  int world = 1, my_rank = 0;
  int64_t rows_per_rank = 0;
  bool used = false;
};
extern thread_local GemmGather* tls_gemm_gather;
// collective over `ranks`: allocate a symmetric buffer, exchange the CUDA IPC handles and map every peer (tp_fused.cc)
void symm_exchange_and_open(const std::string& name, size_t bytes, const std::vector<int>& ranks, int pos);

inline at::Tensor flatten_rows(const at::Tensor& x) { return x.reshape({-1, x.size(-1)}); }

inline void set_out_ds(OpDef& op, size_t out_idx, size_t strategy, const DistributedStates& ds) {
  auto& out = op.outputs[out_idx];
  while (out->ds_hierarchy.size() <= strategy) out->ds_hierarchy.add(DistributedStatesUnion());
  out->ds_hierarchy.get_mut(strategy) = DistributedStatesUnion({ds});
}
inline void copy_out_ds(OpDef& op, size_t out_idx, size_t strategy, const Tensor& from) {
  if (!from->has_ds(strategy)) return;
  auto& out = op.outputs[out_idx];
  while (out->ds_hierarchy.size() <= strategy) out->ds_hierarchy.add(DistributedStatesUnion());
  out->ds_hierarchy.get_mut(strategy) = from->ds_hierarchy.get(strategy);
}

// Layout of Y = A x B for a contraction: A has `a_nd` dims and contracts dim `a_k`; B is 2-D and contracts
// dim `b_k`, its free dim `b_n` becomes Y's last dim.  Contracted shards become partial sums.
DistributedStates matmul_ds(const DistributedStates& a, int a_nd, int a_k, const DistributedStates& b, int b_k, int b_n,
                            int out_n_dim, const std::vector<int>& a_dim_to_out);

inline void make_out(OpDef& op, size_t i, const std::vector<int64_t>& shape, DataType dt) {
  if (op.outputs.size() <= i) op.outputs.resize(i + 1);
  if (!op.outputs[i]) op.outputs[i] = std::make_shared<TensorDef>();
  op.outputs[i]->shape = shape;
  op.outputs[i]->dtype = dt;
}

}  // namespace hb
