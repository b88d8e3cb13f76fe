// C++ host API for the hand-written sm_100a kernels (everything except the GEMM /
// attention families, which have their own headers).  No torch dependency: raw
// pointers + sizes + stream.  bf16 is the activation dtype; statistics, losses,
// optimizer state and gradient accumulators are fp32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hb {

int64_t kernel_launch_count();   // total launches issued through this API

// ---------------------------------------------------------------- norms
// y = (x - mean) * rstd * gamma + beta      (ref: hetu/impl/kernel/FusedLayerNorm.cu:456-1004)
cudaError_t layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd,
                          int64_t rows, int cols, float eps, cudaStream_t s);
// dgamma/dbeta are fp32 [cols]; accumulate=true adds into them. workspace: fp32 [2 * ln_bwd_parts() * cols]
int ln_bwd_parts();
cudaError_t layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
                          void* dx, float* dgamma, float* dbeta, float* workspace, int64_t rows, int cols,
                          bool accumulate, cudaStream_t s, const void* dx_add = nullptr,   // dx_add: bf16 [rows, cols] added into dx
                          void* dgamma_bf16 = nullptr, void* dbeta_bf16 = nullptr);         // when set: results written as bf16 here
// y = x * rstd * gamma                       (ref: hetu/impl/kernel/FusedLayerNorm.cu:1004, RMSNorm.cu)
cudaError_t rmsnorm_fwd(const void* x, const void* gamma, void* y, float* rstd, int64_t rows, int cols, float eps,
                        cudaStream_t s);
cudaError_t rmsnorm_bwd(const void* dy, const void* x, const void* gamma, const float* rstd, void* dx, float* dgamma,
                        float* workspace, int64_t rows, int cols, bool accumulate, cudaStream_t s, const void* dx_add = nullptr,
                        void* dgamma_bf16 = nullptr);

// ---------------------------------------------------------------- generic long-tail kernels (generic.cu)
// dtype codes of the generic kernels
enum GenericDtype : int { GD_F32 = 0, GD_BF16, GD_F16, GD_I64, GD_I32, GD_U8 };
enum GenericUnary : int {
  G_NEG = 0, G_RECIPROCAL, G_ABS, G_CEIL, G_FLOOR, G_ROUND, G_EXP, G_LOG, G_SQRT, G_RSQRT, G_SIN, G_COS, G_CLAMP /*p0 min, p1 max*/, G_SIGMOID,
  G_TANH, G_LEAKYRELU /*p0 slope*/, G_ELU /*p0 alpha, p1 scale*/, G_HARDSHRINK /*p0 lambda*/, G_HARDSIGMOID, G_HARDTANH /*p0 min, p1 max*/,
  G_HARDSWISH, G_LOGSIGMOID, G_SOFTPLUS /*p0 beta, p1 threshold*/, G_MISH, G_SOFTSHRINK /*p0 lambda*/, G_POW /*p0 exponent*/,
  G_ADD_SCALAR, G_MUL_SCALAR, G_RSUB_SCALAR /*p0 - x*/, G_RDIV_SCALAR /*p0 / x*/, G_DIV_SCALAR
};
enum GenericBinary : int { B_ADD = 0, B_SUB, B_MUL, B_DIV, B_MAX, B_MIN, B_POW };
enum GenericReduce : int { R_SUM = 0, R_MEAN, R_MAX, R_MIN, R_PROD };
// dst[...] = src[...] for any-rank (<= 8) strided operands of one shape; strides in elements
cudaError_t strided_copy(int elem_bytes, const void* src, void* dst, int ndim, const int64_t* shape, const int64_t* src_strides,
                         const int64_t* dst_strides, cudaStream_t s);
// y = f(x), contiguous, fp32 / bf16 / fp16 (fp32 arithmetic); cudaErrorMisalignedAddress when a pointer is not 16-byte aligned
cudaError_t generic_unary(int op, int dtype, const void* x, void* y, int64_t n, float p0, float p1, cudaStream_t s);
// out = a (op) b with broadcasting: `out` is contiguous of `out_shape`, operand strides in elements (0 on broadcast dims)
cudaError_t generic_binary(int op, int dtype, const void* a, const void* b, void* out, int ndim, const int64_t* out_shape,
                           const int64_t* a_strides, const int64_t* b_strides, cudaStream_t s);
// x viewed as contiguous [outer, red, inner] -> y [outer, inner]; workspace (fp32, generic_reduce_workspace_floats) enables the
// two-pass path for long reductions with little outer parallelism
int generic_reduce_chunks(int64_t outer, int64_t red, int64_t inner);
int64_t generic_reduce_workspace_floats(int64_t outer, int64_t red, int64_t inner);     // 0: one pass, no workspace needed
cudaError_t generic_reduce(int mode, int dtype, const void* x, void* y, float* workspace, int64_t outer, int64_t red, int64_t inner,
                           cudaStream_t s);
// softmax / log-softmax along the middle extent of a contiguous [outer, dim, inner] view
cudaError_t generic_softmax(bool log, int dtype, const void* x, void* y, int64_t outer, int64_t dim, int64_t inner, cudaStream_t s);
cudaError_t generic_cast(int src_dtype, int dst_dtype, const void* src, void* dst, int64_t n, cudaStream_t s);
cudaError_t generic_fill(int dtype, void* dst, int64_t n, double value, cudaStream_t s);

// ---------------------------------------------------------------- elementwise (bf16 in/out)
enum UnaryOp : int { U_GELU = 0, U_RELU, U_SILU, U_SIGMOID, U_TANH, U_GELU_TANH, U_EXP, U_NEG, U_SQRT, U_RSQRT, U_ABS };
cudaError_t unary_fwd(int op, const void* x, void* y, int64_t n, cudaStream_t s);
// dx = dy * f'(x)
cudaError_t unary_bwd(int op, const void* dy, const void* x, void* dx, int64_t n, cudaStream_t s);
// y[r, :] = silu(x[r, :d]) * x[r, d:2d]        (ref: hetu/impl/kernel/SwiGLU.cu:79,116)
cudaError_t swiglu_fwd(const void* x, void* y, int64_t rows, int d, cudaStream_t s);
cudaError_t swiglu_bwd(const void* dy, const void* x, void* dx, int64_t rows, int d, cudaStream_t s);
// interleaved layout x[r, 2i] = gate_i, x[r, 2i+1] = up_i (pairs stay together under any tensor-parallel split)
cudaError_t swiglu_interleaved_fwd(const void* x, void* y, int64_t rows, int d, cudaStream_t s);
cudaError_t swiglu_interleaved_bwd(const void* dy, const void* x, void* dx, int64_t rows, int d, cudaStream_t s);
// out = a + b (bf16), out may alias a
cudaError_t add_bf16(const void* a, const void* b, void* out, int64_t n, cudaStream_t s);
// dst(fp32) (+)= src(bf16)
cudaError_t accum_bf16_into_fp32(const void* src, float* dst, int64_t n, bool accumulate, cudaStream_t s);
cudaError_t cast_fp32_to_bf16(const float* src, void* dst, int64_t n, cudaStream_t s);
cudaError_t cast_bf16_to_fp32(const void* src, float* dst, int64_t n, cudaStream_t s);
// fused z = residual + dropout(x); y = norm(z): one pass (ref: hetu/impl/kernel/RMSNorm.cu:90,257 DropoutAddLn*);
// residual may be null, p may be 0; cols % 8 == 0 and cols <= 8192
cudaError_t dropout_add_layernorm_fwd(const void* x, const void* residual, const void* gamma, const void* beta, void* y, void* z,
                                      float* mean, float* rstd, int64_t rows, int cols, float eps, float p, uint64_t seed,
                                      uint64_t offset, cudaStream_t s);
cudaError_t dropout_add_rmsnorm_fwd(const void* x, const void* residual, const void* gamma, void* y, void* z, float* rstd,
                                    int64_t rows, int cols, float eps, float p, uint64_t seed, uint64_t offset, cudaStream_t s);
// blockwise absmax quantisation (quant_block.cu); kind: 0 int8, 1 nf4, 2 fp4; 4-bit codes packed two per byte
// q: int8 [n] (kind 0) or uint8 [ceil(n / blocksize) * blocksize / 2]; absmax: fp32 [ceil(n / blocksize)]
cudaError_t quantize_blockwise(const void* x, bool x_is_bf16, void* q, float* absmax, int64_t n, int blocksize, int kind,
                               cudaStream_t s);
cudaError_t dequantize_blockwise(const void* q, const float* absmax, void* y, bool y_is_bf16, int64_t n, int blocksize, int kind,
                                 cudaStream_t s);
// Philox dropout: y = x * mask / (1-p); the mask is recomputed in bwd from (seed, offset)
cudaError_t dropout_fwd(const void* x, void* y, int64_t n, float p, uint64_t seed, uint64_t offset, cudaStream_t s);
// out[c] (+)= sum_r x[r, c]   (bias gradient)  fp32 out
cudaError_t colsum_bf16(const void* x, float* out, int64_t rows, int cols, bool accumulate, cudaStream_t s);

// ---------------------------------------------------------------- rotary (ref: hetu/impl/kernel/rotary.cu:249-400)
// x: [tokens, heads, head_dim] bf16, half-split convention; pos[token] gives the position id.
// inverse=true applies the transposed rotation (backward).
cudaError_t rotary_apply(const void* x, void* y, const int32_t* pos, int64_t tokens, int heads, int head_dim,
                         int rot_dim, float base, bool inverse, int64_t token_stride, cudaStream_t s, int group_size = 0,
                         int group_stride = 0);   // grouped layouts: head h -> (h / group_size) * group_stride + (h % group_size) * head_dim

// ---------------------------------------------------------------- scaled masked softmax (softmax.cu)
// y = softmax(scale * x) over the last dim of [rows, cols] bf16; mode 0 plain, 1 boolean mask [b, 1, sq, cols] (1 = masked),
// 2 causal (bottom-right aligned); rows are ordered [b, h, sq], rows_per_batch = h * sq
cudaError_t scaled_softmax_fwd(const void* x, const uint8_t* mask, void* y, int64_t rows, int cols, int rows_per_batch, int sq, float scale,
                               int mode, cudaStream_t s);
cudaError_t scaled_softmax_bwd(const void* dy, const void* y, void* dx, int64_t rows, int cols, float scale, cudaStream_t s);

// ---------------------------------------------------------------- fp8 quantisation (quant_fp8.cu)
// q[r, :] = e4m3(x[r, :] / scale[r]) with scale[r] = amax(row r) / 448
cudaError_t quantize_rowwise_e4m3(const void* x, void* q, float* scale, int64_t rows, int cols, int64_t ldx, int64_t ldq,
                                  cudaStream_t s);
// q[c, r] = e4m3(x[r, c] / scale[c]) with one scale per column of x (transposed, K-major operand of the dgrad GEMM)
cudaError_t quantize_transpose_e4m3(const void* x, void* q, float* scale, int64_t rows, int cols, int64_t ldq, cudaStream_t s);

// ---------------------------------------------------------------- embedding (ref: hetu/impl/kernel/EmbeddingLookup.cu:91,140)
// y[t, :] = wte[ids[t], :] (+ wpe[pos[t], :])
cudaError_t embedding_fwd(const int64_t* ids, const int32_t* pos, const void* wte, const void* wpe, void* y,
                          int64_t tokens, int hidden, int64_t vocab, cudaStream_t s);
// dwte[ids[t], :] += dy[t, :]   (fp32 accumulators), dwpe likewise when non-null
cudaError_t embedding_bwd(const int64_t* ids, const int32_t* pos, const void* dy, float* dwte, float* dwpe,
                          int64_t tokens, int hidden, int64_t vocab, cudaStream_t s);

// ---------------------------------------------------------------- softmax cross-entropy
// Fused fwd+bwd over bf16 logits [rows, ld] (only the first `cols` are valid):
//   loss[r] = logsumexp(logits[r]) - logits[r, label]  (0 when label == ignore_index)
//   logits[r, :] <- (softmax - onehot) * grad_scale    (in place, when write_grad)
// (ref: hetu/impl/kernel/SoftmaxCrossEntropySparse.cu, VocabParallelCrossEntropyLoss.cu:57,114)
cudaError_t softmax_ce_fwd_bwd(void* logits, const int64_t* labels, float* loss, float* lse, int64_t rows, int cols,
                               int64_t ld, int64_t ignore_index, float grad_scale, bool write_grad, cudaStream_t s,
                               const float* grad_scale_ptr = nullptr);   // device scalar multiplied into grad_scale
// y[r, :] = x[r, :] * scale[per_row ? r : 0]   (scale lives on the device; bf16 rows of `cols` elements)
cudaError_t scale_rows_bf16(const void* x, void* y, const float* scale, bool per_row, int64_t rows, int64_t cols,
                            cudaStream_t s);
// Vocab-parallel pieces: local max / local sum-exp & target logit, with a collective between them.
cudaError_t vp_ce_local_max(const void* logits, float* row_max, int64_t rows, int cols, int64_t ld, cudaStream_t s);
cudaError_t vp_ce_local_sum(const void* logits, const int64_t* labels, const float* row_max, float* sum_exp,
                            float* target_logit, int64_t rows, int cols, int64_t ld, int64_t vocab_start,
                            cudaStream_t s);
cudaError_t vp_ce_finish(void* logits, const int64_t* labels, const float* row_max, const float* sum_exp,
                         const float* target_logit, float* loss, int64_t rows, int cols, int64_t ld,
                         int64_t vocab_start, int64_t ignore_index, float grad_scale, bool write_grad, cudaStream_t s);

// ---------------------------------------------------------------- optimizers (ref: hetu/impl/kernel/Optimizers.cu:13-188)
// Flat fused Adam(W) over a contiguous shard: fp32 master params / m / v, fp32 or bf16 grads, writes the
// bf16 compute copy in the same pass.  step/bias-correction are computed on device from *step_ptr (int64)
// so the launch is CUDA-graph capturable; grad_scale_ptr (optional) multiplies grads (loss-scale / 1/N).
struct AdamArgs {
  float* master = nullptr;     // fp32 [n]
  float* m = nullptr;          // fp32 [n]
  float* v = nullptr;          // fp32 [n]
  const void* grad = nullptr;  // fp32 or bf16 [n]
  bool grad_is_bf16 = false;
  void* param_bf16 = nullptr;  // optional bf16 [n] compute copy
  int64_t n = 0;
  float lr = 1e-3f, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f, weight_decay = 0.0f;
  const int64_t* step_ptr = nullptr;     // device step counter
  int step_add = 0;                      // bias correction uses *step_ptr + step_add (counter bumped later, in one batch)
  const float* grad_scale_ptr = nullptr; // device pointer or null
  const float* lr_ptr = nullptr;         // device pointer overriding lr (scheduler) or null
  bool zero_grad = false;                // clear the grad buffer in the same pass
};
cudaError_t adam_update(const AdamArgs& a, cudaStream_t s);
// ZeRO fused update over symmetric memory: ONE kernel per parameter
//   grad   = grad_scale * sum over `nslots` bf16 gradient slots (the GEMM->peer-store epilogue of the weight-gradient
//            GEMMs filled slot s with rank s's partial sum of this rank's rows)          -> the reduce-scatter
//   m, v, fp32 master of this rank's shard are updated (AdamW)                            -> the optimizer
//   the new bf16 shard is stored into EVERY rank's parameter tensor through peer pointers -> the all-gather
struct AdamZeroArgs {
  float* master = nullptr; float* m = nullptr; float* v = nullptr;   // fp32 shard [n]
  const void* slots = nullptr;      // bf16 [nslots, slot_stride] (this rank's symmetric staging area)
  int nslots = 1;
  int64_t slot_stride = 0;          // elements between slots
  void* peer_param[8] = {nullptr};  // bf16 destination of this shard inside each rank's (symmetric) parameter tensor
  int world = 1;
  int64_t n = 0;
  float lr = 1e-3f, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f, weight_decay = 0.0f, grad_scale = 1.0f;
  const int64_t* step_ptr = nullptr;   // device step counter
  int step_add = 0;                    // bias correction uses *step_ptr + step_add (counter bumped later in one batch)
};
cudaError_t adam_zero_fused(const AdamZeroArgs& a, cudaStream_t s);
// the same step over an NVLS multicast mapping: gradient shard = multimem.ld_reduce over all ranks' copies, new bf16
// shard = one multimem.st into every rank's parameter tensor (see optim.cu)
struct AdamNvlsArgs {
  float* master = nullptr; float* m = nullptr; float* v = nullptr;   // fp32 shard [n]
  const void* grad_mc = nullptr;    // multicast address of the shard inside the symmetric bf16 gradient region
  void* param_mc = nullptr;         // multicast address of the shard inside the symmetric bf16 parameter tensor
  int64_t n = 0;                    // shard elements (multiple of 8)
  float lr = 1e-3f, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f, weight_decay = 0.0f, grad_scale = 1.0f;
  const int64_t* step_ptr = nullptr;
  int step_add = 0;
};
cudaError_t adam_zero_nvls(const AdamNvlsArgs& a, cudaStream_t s);
// counters[i] += 1 for a device table of int64 pointers (one launch for all optimizer step counters)
cudaError_t increment_many_i64(int64_t* const* table, int count, cudaStream_t s);
cudaError_t sgd_update(float* master, float* momentum_buf, const void* grad, bool grad_is_bf16, void* param_bf16,
                       int64_t n, float lr, float momentum, bool nesterov, float weight_decay, cudaStream_t s);
cudaError_t increment_step(int64_t* step_ptr, cudaStream_t s);
// found_inf[0] = 1 if any non-finite value in x (fp32)
cudaError_t check_finite(const float* x, int64_t n, float* found_inf, cudaStream_t s);
// sum of squares into out[0] (fp32) for grad-norm clipping
cudaError_t sumsq_fp32(const float* x, int64_t n, float* out, bool accumulate, cudaStream_t s);

// ---------------------------------------------------------------- MoE (ref: hetu/v1/src/ops/{TopKIdx,LayoutTransform,...}.cu)
// probs: softmax(logits) fp32 [tokens, experts]; top-k indices/values.
cudaError_t moe_gate_topk(const void* logits_bf16, float* probs, int32_t* topk_idx, float* topk_val, int64_t tokens,
                          int experts, int k, cudaStream_t s);
// capacity-based slot assignment: location[t,k] = position of token t inside expert e's buffer (or -1 when dropped)
cudaError_t moe_assign_slots(const int32_t* topk_idx, int32_t* location, int32_t* expert_count, int64_t tokens,
                             int experts, int k, int capacity, cudaStream_t s);
// y[b][a][chunk] = x[a][b][chunk] (16-byte aligned chunks): layout transform of the hierarchical all-to-all
cudaError_t chunk_transpose(const void* x, void* y, int a, int b, int64_t chunk_bytes, cudaStream_t s);
// BASE balanced assignment (ref: hetu/v1/python/hetu/gpu_ops/BalanceAssignment.py): idx[t] = expert, loc[t] = slot, every
// expert gets at most `capacity` tokens (exactly T/E when E divides T).  scores fp32 [tokens, experts];
// filled [experts] and choice [tokens] are int32 scratch.
cudaError_t moe_balance_assign(const float* scores, int32_t* idx, int32_t* loc, int32_t* filled, int32_t* choice,
                               int64_t tokens, int experts, int capacity, cudaStream_t s);
// dispatched[e, slot, :] = (scale ? scale[t,k] : 1) * x[t, :]; unassigned slots are zero-filled.
// (with scale = gates this is also the backward of moe_combine w.r.t. expert_out)
cudaError_t moe_dispatch(const void* x, const int32_t* topk_idx, const int32_t* location, const float* scale,
                         void* dispatched, int64_t tokens, int hidden, int experts, int k, int capacity,
                         cudaStream_t s);
// y[t, :] = sum_k (gate ? gate[t,k] : 1) * expert_out[e_k, slot_k, :]
// (with gate = null this is also the backward of moe_dispatch w.r.t. x)
cudaError_t moe_combine(const void* expert_out, const int32_t* topk_idx, const int32_t* location, const float* gate,
                        void* y, int64_t tokens, int hidden, int experts, int k, int capacity, cudaStream_t s);
// Expert-parallel variants over NVLink peer memory (the dispatch / combine all-to-all fused into the layout transform):
// expert e lives on rank e / experts_per_rank; its buffer there is [experts_per_rank, ep * capacity, hidden] and tokens of
// source rank r occupy rows [r * capacity, (r + 1) * capacity).  base[r] = rank r's (symmetric) buffer.
struct MoePeers {
  void* base[8] = {nullptr};
  int ep = 1, experts_per_rank = 1, src_rank = 0;
};
// scatter: peer[e / epr][(e % epr), src_rank * C + slot, :] = scale * x[t, :]   (buffers must have been zero-filled)
cudaError_t moe_dispatch_peers(const void* x, const int32_t* topk_idx, const int32_t* location, const float* scale,
                               const MoePeers& peers, int64_t tokens, int hidden, int k, int capacity, cudaStream_t s);
// gather: y[t, :] = sum_k gate * peer[e / epr][(e % epr), src_rank * C + slot, :]
cudaError_t moe_combine_peers(const MoePeers& peers, const int32_t* topk_idx, const int32_t* location, const float* gate,
                              void* y, int64_t tokens, int hidden, int k, int capacity, cudaStream_t s);
cudaError_t moe_combine_bwd_gate_peers(const void* dy, const MoePeers& peers, const int32_t* topk_idx, const int32_t* location,
                                       float* dgate, int64_t tokens, int hidden, int k, int capacity, cudaStream_t s);
// backward of combine wrt gates: dgate[t,k] = <dy[t], expert_out[e_k, slot_k]>
cudaError_t moe_combine_bwd_gate(const void* dy, const void* expert_out, const int32_t* topk_idx,
                                 const int32_t* location, float* dgate, int64_t tokens, int hidden, int experts, int k,
                                 int capacity, cudaStream_t s);

}  // namespace hb
