// Host API of the hand-written sm_100a flash-attention family (tcgen05 + TMEM + TMA).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hb {

// All tensors are bf16 with head_dim contiguous; strides are in elements.
// q: [B, Sq, Hq, D]   k, v: [B, Sk, Hkv, D]   o: [B, Sq, Hq, D]   lse: fp32 [B, Hq, Sq]
// (a packed qkv projection output is expressed through the strides, no copies)
struct AttnTensor {
  const void* ptr = nullptr;
  int64_t stride_b = 0, stride_s = 0, stride_h = 0;
  // Grouped (GQA, kv-head-major) packed layouts: logical head h lives in slot (h / h_div) * h_mul + h % h_div of a
  // row made of `h_slots` head-sized slots (stride_h apart).  h_div == 0: plain layout, slot == h.
  int h_div = 0, h_mul = 0, h_slots = 0;
};

struct AttnFwdCall {
  AttnTensor q, k, v;
  AttnTensor o;   // written
  float* lse = nullptr;
  int B = 0, Hq = 0, Hkv = 0, Sq = 0, Sk = 0, D = 0;
  float softmax_scale = 1.0f;
  bool causal = true;   // bottom-right aligned when Sq != Sk (FlashAttention-2 convention)
  // packed variable-length rows (B == 1, causal, Sq == Sk): device int32 [Sq], first token of every token's document.
  // One launch covers all documents (block-diagonal causal mask, tiles outside a document are skipped).
  const int* row_start = nullptr;
};

struct AttnBwdCall {
  AttnTensor q, k, v, o, d_o;
  AttnTensor dq, dk, dv;   // written
  const float* lse = nullptr;
  float* delta = nullptr;  // workspace fp32 [B, Hq, Sq]
  int B = 0, Hq = 0, Hkv = 0, Sq = 0, Sk = 0, D = 0;
  float softmax_scale = 1.0f;
  bool causal = true;
  const int* row_start = nullptr;   // varlen (see AttnFwdCall): document start and end (exclusive) of every token
  const int* row_end = nullptr;
};

cudaError_t attn_fwd(const AttnFwdCall& c, cudaStream_t s);
cudaError_t attn_bwd(const AttnBwdCall& c, cudaStream_t s);
int64_t attn_launch_count();

}  // namespace hb
