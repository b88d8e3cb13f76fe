// Flash-attention backward for sm_100a, split in two tcgen05 kernels that share
// the forward's structure (score GEMMs -> elementwise in registers -> bf16 operand
// written back into TMEM -> accumulate GEMM whose A operand is read from TMEM):
//
//   dQ kernel  (CTA per 128 query rows, loops over KV blocks)
//       S  = Q K^T, dP = dO V^T          -> P = exp2(S*c - lse), dS = P o (dP - delta) * scale
//       dQ += dS K                        (A = dS from TMEM, B = K tile read MN-major)
//   dKV kernel (CTA per 128 key rows, loops over query tiles and the q-heads of a GQA group)
//       S^T = K Q^T, dP^T = V dO^T       -> P^T, dS^T (row statistics broadcast from smem)
//       dV += P^T dO, dK += dS^T Q       (A from TMEM, B = dO / Q tiles read MN-major)
//
// No atomics, no transposes through shared memory: the same 128B-swizzled TMA tile
// is consumed K-major by the score GEMM and MN-major by the accumulate GEMM.
//
// Capability parity: hetu/impl/kernel/FlashAttention.cu:592 (FlashAttnGradientCuda ->
// run_mha_bwd_), causal, GQA.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>

#include "attention_fwd_sm100.cuh"
#include "attention_sm100.h"
#include "ptx.cuh"
#include "tma_host.h"

namespace hb {

struct AttnBwdParams {
  int q_div, q_mul;   // grouped-layout slot mapping of Q / dQ heads (q_div == 0: identity)
  int B, Hq, Hkv, Sq, Sk;
  float scale_log2, scale;
  int causal, causal_off;
  const float* LSE;    // [B, Hq, Sq] natural log
  const float* DELTA;  // [B, Hq, Sq]
  __nv_bfloat16 *dQ, *dK, *dV;
  int64_t dq_sb, dq_ss, dq_sh, dk_sb, dk_ss, dk_sh, dv_sb, dv_ss, dv_sh;
  // packed variable-length rows in one launch (B == 1, causal): document start / end (exclusive) of every token
  const int* row_start = nullptr;
  const int* row_end = nullptr;
};

// delta[b,h,s] = sum_d dO * O: D/8 lanes per row, one 16-byte load of each tensor per lane (2 or 4 rows per warp)
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                                  float* __restrict__ delta, int B, int H, int S, int D, int64_t o_sb, int64_t o_ss,
                                  int64_t o_sh, int64_t do_sb, int64_t do_ss, int64_t do_sh) {
  const int lpr = D >> 3;                                   // lanes per row: 8 (D = 64) or 16 (D = 128)
  const int64_t gt = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  const int64_t row = gt / lpr;
  const int c = int(gt % lpr) * 8;
  float acc = 0.f;
  const bool live = row < int64_t(B) * H * S;
  int s = 0, h = 0, b = 0;
  if (live) {
    s = int(row % S);
    h = int((row / S) % H);
    b = int(row / (int64_t(S) * H));
    const uint4 ra = __ldg(reinterpret_cast<const uint4*>(o + b * o_sb + s * o_ss + h * o_sh + c));
    const uint4 rg = __ldg(reinterpret_cast<const uint4*>(d_o + b * do_sb + s * do_ss + h * do_sh + c));
    const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&ra);
    const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&rg);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = __bfloat1622float2(a2[j]), g = __bfloat1622float2(g2[j]);
      acc += a.x * g.x + a.y * g.y;
    }
  }
  for (int o2 = lpr >> 1; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o2);
  if (live && c == 0) delta[(int64_t(b) * H + h) * S + s] = acc;
}

// ---------------------------------------------------------------------------------------------
// dQ kernel
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256, 1)
attn_bwd_dq_sm100_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                         const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                         const AttnBwdParams p) {
  using namespace attn_detail;
  constexpr int NBOX = D / 64;
  constexpr int T = 128 * D * 2;
  constexpr int DP_COL = 256, DQ_COL = 384;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sDO = smem + T;
  uint8_t* sK = smem + 2 * T;  // 2 stages
  uint8_t* sV = smem + 4 * T;  // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * T);
  uint64_t* qdo_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 3;
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 7;
  uint64_t* s_full = bars + 9;
  uint64_t* ds_ready = bars + 11;
  uint64_t* dp_full = bars + 13;
  uint64_t* dq_done = bars + 14;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = gridDim.x - 1 - blockIdx.x;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int hk = h / (p.Hq / p.Hkv);
  const int hslot = p.q_div ? (h / p.q_div) * p.q_mul + (h % p.q_div) : h;
  const int q0 = qt * 128;
  int n_kv_end = (p.Sk + 127) / 128;
  if (p.causal) {
    const int last = q0 + 127 + p.causal_off;
    const int lim = last < 0 ? 0 : last / 128 + 1;
    n_kv_end = min(n_kv_end, lim);
  }
  const int j_begin = (p.row_start != nullptr) ? min(p.row_start[min(q0, p.Sq - 1)] / 128, n_kv_end) : 0;
  const int n_kv = n_kv_end - j_begin;          // KV tiles visited: tile index = j_begin + j

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_q); ptx::prefetch_tmap(&tmap_do); ptx::prefetch_tmap(&tmap_k); ptx::prefetch_tmap(&tmap_v);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(qdo_full, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&k_full[i], 1); ptx::mbar_init(&k_empty[i], 1);
      ptx::mbar_init(&v_full[i], 1); ptx::mbar_init(&v_empty[i], 1);
      ptx::mbar_init(&s_full[i], 1); ptx::mbar_init(&ds_ready[i], 4);
    }
    ptx::mbar_init(dp_full, 1);
    ptx::mbar_init(dq_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<1>(tmem_ptr_smem, 512);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0 && n_kv > 0) {
      ptx::mbar_arrive_expect_tx(qdo_full, 2 * T);
      for (int bx = 0; bx < NBOX; ++bx) {
        ptx::tma_load_4d(sQ + bx * kBoxBytes, &tmap_q, qdo_full, bx * 64, hslot, q0, b);
        ptx::tma_load_4d(sDO + bx * kBoxBytes, &tmap_do, qdo_full, bx * 64, h, q0, b);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        ptx::mbar_wait(&k_empty[st], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&k_full[st], T);
        for (int bx = 0; bx < NBOX; ++bx)
          ptx::tma_load_4d(sK + st * T + bx * kBoxBytes, &tmap_k, &k_full[st], bx * 64, hk, (j_begin + j) * 128, b);
        ptx::mbar_wait(&v_empty[st], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&v_full[st], T);
        for (int bx = 0; bx < NBOX; ++bx)
          ptx::tma_load_4d(sV + st * T + bx * kBoxBytes, &tmap_v, &v_full[st], bx * 64, hk, (j_begin + j) * 128, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && n_kv > 0) {
      constexpr uint32_t idesc_s = ptx::make_idesc(128, 128, 1, 1, false, false);
      constexpr uint32_t idesc_dq = ptx::make_idesc(128, D, 1, 1, false, true);
      const uint32_t q_addr = ptx::smem_u32(sQ), do_addr = ptx::smem_u32(sDO);
      const uint32_t k_addr = ptx::smem_u32(sK), v_addr = ptx::smem_u32(sV);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        ptx::mbar_wait(&k_full[st], (j >> 1) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk >> 2) * kBoxBytes + (kk & 3) * 32;
          ptx::mma_f16_ss<1>(tmem_base + st * 128, ptx::make_smem_desc_sw128(q_addr + off, 0, 1024),
                             ptx::make_smem_desc_sw128(k_addr + st * T + off, 0, 1024), idesc_s, kk > 0 ? 1u : 0u);
        }
        ptx::mma_commit<1>(&s_full[st]);
      };
      auto issue_dp = [&](int j) {
        const int st = j & 1;
        ptx::mbar_wait(&v_full[st], (j >> 1) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk >> 2) * kBoxBytes + (kk & 3) * 32;
          ptx::mma_f16_ss<1>(tmem_base + DP_COL, ptx::make_smem_desc_sw128(do_addr + off, 0, 1024),
                             ptx::make_smem_desc_sw128(v_addr + st * T + off, 0, 1024), idesc_s, kk > 0 ? 1u : 0u);
        }
        ptx::mma_commit<1>(&v_empty[st]);
        ptx::mma_commit<1>(dp_full);
      };
      ptx::mbar_wait(qdo_full, 0);
      issue_s(0);
      issue_dp(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) issue_s(j + 1);
        const int st = j & 1;
        ptx::mbar_wait(&ds_ready[st], (j >> 1) & 1);
        ptx::tc_fence_after();
        if (j + 1 < n_kv) issue_dp(j + 1);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t bd = ptx::make_smem_desc_sw128(k_addr + st * T + kk * 2048, kBoxBytes, 1024);
          ptx::mma_f16_ts<1>(tmem_base + DQ_COL, tmem_base + st * 128 + kk * 8, bd, idesc_dq, (j > 0 || kk > 0) ? 1u : 0u);
        }
        ptx::mma_commit<1>(&k_empty[st]);
      }
      ptx::mma_commit<1>(dq_done);
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const int grow = q0 + row;
    const bool row_ok = grow < p.Sq;
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    const int64_t stat_idx = (int64_t(b) * p.Hq + h) * p.Sq + grow;
    float lse2 = row_ok ? p.LSE[stat_idx] * 1.4426950408889634f : INFINITY;
    if (lse2 == -INFINITY) lse2 = INFINITY;  // fully masked row: P = 0
    const float delta = row_ok ? p.DELTA[stat_idx] : 0.f;
    // visible keys of this row: [doc_lo, lim] (doc_lo = 0 without packing); one unsigned compare per element tests both ends
    const int lim = p.causal ? min(p.Sk - 1, grow + p.causal_off) : p.Sk - 1;
    const int doc_lo = (p.row_start != nullptr && row_ok) ? p.row_start[grow] : 0;
    const unsigned span = lim >= doc_lo ? unsigned(lim - doc_lo) : 0u;
    const int vis_lo = lim >= doc_lo ? doc_lo : -(1 << 29);          // nothing visible: every column compares above `span`
    for (int j = 0; j < n_kv; ++j) {
      const int st = j & 1;
      ptx::mbar_wait(&s_full[st], (j >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + lane_base + st * 128;
      uint32_t sr[128];
      ptx::tmem_ld_32x32b_x32(taddr, sr);
      ptx::tmem_ld_32x32b_x32(taddr + 32, sr + 32);
      ptx::tmem_ld_32x32b_x32(taddr + 64, sr + 64);
      ptx::tmem_ld_32x32b_x32(taddr + 96, sr + 96);
      ptx::tmem_ld_wait();
      const int kb = (j_begin + j) * 128 - vis_lo;
#pragma unroll
      for (int c = 0; c < 128; ++c) {
        float pv = ex2(fmaf(__uint_as_float(sr[c]), p.scale_log2, -lse2));
        if (unsigned(kb + c) > span) pv = 0.f;
        sr[c] = __float_as_uint(pv);
      }
      ptx::mbar_wait(dp_full, j & 1);
      ptx::tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t dpr[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + lane_base + DP_COL + c * 32, dpr);
        ptx::tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const float d0 = __uint_as_float(sr[c * 32 + 2 * t]) * (__uint_as_float(dpr[2 * t]) - delta) * p.scale;
          const float d1 = __uint_as_float(sr[c * 32 + 2 * t + 1]) * (__uint_as_float(dpr[2 * t + 1]) - delta) * p.scale;
          const __nv_bfloat162 v2 = __floats2bfloat162_rn(d0, d1);
          pk[t] = *reinterpret_cast<const uint32_t*>(&v2);
        }
        ptx::tmem_st_32x32b_x16(taddr + c * 16, pk);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&ds_ready[st]);
    }
    __nv_bfloat16* orow = p.dQ + int64_t(b) * p.dq_sb + int64_t(grow) * p.dq_ss + int64_t(hslot) * p.dq_sh;
    if (n_kv > 0) {
      ptx::mbar_wait(dq_done, 0);
      ptx::tc_fence_after();
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint32_t orr[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + lane_base + DQ_COL + c * 32, orr);
        ptx::tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int t4 = 0; t4 < 4; ++t4) {
            uint4 o;
            __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int t = 0; t < 4; ++t)
              o2[t] = __floats2bfloat162_rn(__uint_as_float(orr[t4 * 8 + t * 2]), __uint_as_float(orr[t4 * 8 + t * 2 + 1]));
            reinterpret_cast<uint4*>(orow + c * 32)[t4] = o;
          }
        }
      }
    } else if (row_ok) {
      for (int c = 0; c < D / 8; ++c) reinterpret_cast<uint4*>(orow)[c] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// dK/dV kernel
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256, 1)
attn_bwd_dkv_sm100_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                          const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                          const AttnBwdParams p) {
  using namespace attn_detail;
  constexpr int NBOX = D / 64;
  constexpr int T = 128 * D * 2;
  constexpr int ST_COL = 0, DPT_COL = 128, DV_COL = 256, DK_COL = 256 + D;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + T;
  uint8_t* sQ = smem + 2 * T;   // 2 stages
  uint8_t* sDO = smem + 4 * T;  // 2 stages
  float* sStat = reinterpret_cast<float*>(smem + 6 * T);  // [2 stages][3][128]: lse, delta, document start of the tile's queries
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * T + 3072);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;
  uint64_t* q_empty = bars + 3;
  uint64_t* do_full = bars + 5;
  uint64_t* do_empty = bars + 7;
  uint64_t* st_full = bars + 9;
  uint64_t* dpt_full = bars + 10;
  uint64_t* pt_ready = bars + 11;
  uint64_t* dst_ready = bars + 12;
  uint64_t* acc_done = bars + 13;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kt = blockIdx.x;
  const int hk = blockIdx.y;
  const int b = blockIdx.z;
  const int group = p.Hq / p.Hkv;
  const int k0 = kt * 128;
  const int n_q = (p.Sq + 127) / 128;
  int i_start = 0;
  if (p.causal) {
    const int first_row = k0 - p.causal_off;  // first query row that can see key k0
    i_start = first_row <= 0 ? 0 : first_row / 128;
    if (i_start > n_q) i_start = n_q;
  }
  // varlen: queries past the end of the last document that starts inside this key tile never see these keys
  int i_end = n_q;
  if (p.row_end != nullptr) i_end = min(n_q, (p.row_end[min(k0 + 127, p.Sk - 1)] + 127) / 128);
  const int nq_iters = max(i_end - i_start, 0);
  const int n_it = nq_iters * group;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_q); ptx::prefetch_tmap(&tmap_do); ptx::prefetch_tmap(&tmap_k); ptx::prefetch_tmap(&tmap_v);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&q_full[i], 1); ptx::mbar_init(&q_empty[i], 1);
      ptx::mbar_init(&do_full[i], 1); ptx::mbar_init(&do_empty[i], 1);
    }
    ptx::mbar_init(st_full, 1); ptx::mbar_init(dpt_full, 1);
    ptx::mbar_init(pt_ready, 4); ptx::mbar_init(dst_ready, 4);
    ptx::mbar_init(acc_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<1>(tmem_ptr_smem, 512);
    ptx::tmem_relinquish<1>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0 && n_it > 0) {
      ptx::mbar_arrive_expect_tx(kv_full, 2 * T);
      for (int bx = 0; bx < NBOX; ++bx) {
        ptx::tma_load_4d(sK + bx * kBoxBytes, &tmap_k, kv_full, bx * 64, hk, k0, b);
        ptx::tma_load_4d(sV + bx * kBoxBytes, &tmap_v, kv_full, bx * 64, hk, k0, b);
      }
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        const int hq = hk * group + it / nq_iters;
        const int q0 = (i_start + it % nq_iters) * 128;
        ptx::mbar_wait(&q_empty[st], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&q_full[st], T);
        for (int bx = 0; bx < NBOX; ++bx)
          ptx::tma_load_4d(sQ + st * T + bx * kBoxBytes, &tmap_q, &q_full[st], bx * 64,
                           p.q_div ? (hq / p.q_div) * p.q_mul + (hq % p.q_div) : hq, q0, b);
        ptx::mbar_wait(&do_empty[st], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&do_full[st], T);
        for (int bx = 0; bx < NBOX; ++bx)
          ptx::tma_load_4d(sDO + st * T + bx * kBoxBytes, &tmap_do, &do_full[st], bx * 64, hq, q0, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && n_it > 0) {
      constexpr uint32_t idesc_s = ptx::make_idesc(128, 128, 1, 1, false, false);
      constexpr uint32_t idesc_acc = ptx::make_idesc(128, D, 1, 1, false, true);
      const uint32_t q_addr = ptx::smem_u32(sQ), do_addr = ptx::smem_u32(sDO);
      const uint32_t k_addr = ptx::smem_u32(sK), v_addr = ptx::smem_u32(sV);
      ptx::mbar_wait(kv_full, 0);
      for (int it = 0; it < n_it; ++it) {
        const int st = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        // S^T = K Q^T
        ptx::mbar_wait(&q_full[st], ph);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk >> 2) * kBoxBytes + (kk & 3) * 32;
          ptx::mma_f16_ss<1>(tmem_base + ST_COL, ptx::make_smem_desc_sw128(k_addr + off, 0, 1024),
                             ptx::make_smem_desc_sw128(q_addr + st * T + off, 0, 1024), idesc_s, kk > 0 ? 1u : 0u);
        }
        ptx::mma_commit<1>(st_full);
        // dP^T = V dO^T
        ptx::mbar_wait(&do_full[st], ph);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk >> 2) * kBoxBytes + (kk & 3) * 32;
          ptx::mma_f16_ss<1>(tmem_base + DPT_COL, ptx::make_smem_desc_sw128(v_addr + off, 0, 1024),
                             ptx::make_smem_desc_sw128(do_addr + st * T + off, 0, 1024), idesc_s, kk > 0 ? 1u : 0u);
        }
        ptx::mma_commit<1>(dpt_full);
        // dV += P^T dO
        ptx::mbar_wait(pt_ready, it & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t bd = ptx::make_smem_desc_sw128(do_addr + st * T + kk * 2048, kBoxBytes, 1024);
          ptx::mma_f16_ts<1>(tmem_base + DV_COL, tmem_base + ST_COL + kk * 8, bd, idesc_acc, (it > 0 || kk > 0) ? 1u : 0u);
        }
        ptx::mma_commit<1>(&do_empty[st]);
        // dK += dS^T Q
        ptx::mbar_wait(dst_ready, it & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t bd = ptx::make_smem_desc_sw128(q_addr + st * T + kk * 2048, kBoxBytes, 1024);
          ptx::mma_f16_ts<1>(tmem_base + DK_COL, tmem_base + DPT_COL + kk * 8, bd, idesc_acc, (it > 0 || kk > 0) ? 1u : 0u);
        }
        ptx::mma_commit<1>(&q_empty[st]);
      }
      ptx::mma_commit<1>(acc_done);
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;   // key row inside the tile
    const int gk = k0 + row;
    const int tid = threadIdx.x - 128;  // 0..127
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    for (int it = 0; it < n_it; ++it) {
      const int hq = hk * group + it / nq_iters;
      const int q0 = (i_start + it % nq_iters) * 128;
      float* stat = sStat + (it & 1) * 384;
      {
        const int gq = q0 + tid;
        const bool ok = gq < p.Sq;
        const int64_t si = (int64_t(b) * p.Hq + hq) * p.Sq + gq;
        float l2 = ok ? p.LSE[si] * 1.4426950408889634f : INFINITY;
        if (l2 == -INFINITY) l2 = INFINITY;  // fully masked row: P = 0
        stat[tid] = l2;
        stat[128 + tid] = ok ? p.DELTA[si] : 0.f;
        reinterpret_cast<int*>(stat)[256 + tid] = (p.row_start != nullptr && ok) ? p.row_start[gq] : 0;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      ptx::mbar_wait(st_full, it & 1);
      ptx::tc_fence_after();
      uint32_t sr[128];
      const uint32_t ta = tmem_base + lane_base + ST_COL;
      ptx::tmem_ld_32x32b_x32(ta, sr);
      ptx::tmem_ld_32x32b_x32(ta + 32, sr + 32);
      ptx::tmem_ld_32x32b_x32(ta + 64, sr + 64);
      ptx::tmem_ld_32x32b_x32(ta + 96, sr + 96);
      ptx::tmem_ld_wait();
      // visible iff key gk <= min(Sk-1, qrow + off)  <=>  qrow >= gk - off (causal) and gk < Sk
      const int first_q = p.causal ? gk - p.causal_off : -0x3fffffff;
      const bool key_ok = gk < p.Sk;
#pragma unroll
      for (int c4 = 0; c4 < 32; ++c4) {
        const float4 l4 = *reinterpret_cast<const float4*>(stat + c4 * 4);
        const float ls[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int c = c4 * 4 + t;
          float pv = ex2(fmaf(__uint_as_float(sr[c]), p.scale_log2, -ls[t]));
          if (!key_ok || (q0 + c) < first_q || gk < reinterpret_cast<const int*>(stat)[256 + c]) pv = 0.f;   // (other document)
          sr[c] = __float_as_uint(pv);
        }
      }
      {
        uint32_t pk[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          const __nv_bfloat162 v2 = __floats2bfloat162_rn(__uint_as_float(sr[2 * c]), __uint_as_float(sr[2 * c + 1]));
          pk[c] = *reinterpret_cast<const uint32_t*>(&v2);
        }
        ptx::tmem_st_32x32b_x32(ta, pk);
        ptx::tmem_st_32x32b_x32(ta + 32, pk + 32);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(pt_ready);
      }
      ptx::mbar_wait(dpt_full, it & 1);
      ptx::tc_fence_after();
      const uint32_t td = tmem_base + lane_base + DPT_COL;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t dpr[32];
        ptx::tmem_ld_32x32b_x32(td + c * 32, dpr);
        ptx::tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int t4 = 0; t4 < 8; ++t4) {
          const float4 d4 = *reinterpret_cast<const float4*>(stat + 128 + c * 32 + t4 * 4);
          const float d0 = __uint_as_float(sr[c * 32 + t4 * 4 + 0]) * (__uint_as_float(dpr[t4 * 4 + 0]) - d4.x) * p.scale;
          const float d1 = __uint_as_float(sr[c * 32 + t4 * 4 + 1]) * (__uint_as_float(dpr[t4 * 4 + 1]) - d4.y) * p.scale;
          const float d2 = __uint_as_float(sr[c * 32 + t4 * 4 + 2]) * (__uint_as_float(dpr[t4 * 4 + 2]) - d4.z) * p.scale;
          const float d3 = __uint_as_float(sr[c * 32 + t4 * 4 + 3]) * (__uint_as_float(dpr[t4 * 4 + 3]) - d4.w) * p.scale;
          const __nv_bfloat162 a2 = __floats2bfloat162_rn(d0, d1);
          const __nv_bfloat162 b2 = __floats2bfloat162_rn(d2, d3);
          pk[t4 * 2] = *reinterpret_cast<const uint32_t*>(&a2);
          pk[t4 * 2 + 1] = *reinterpret_cast<const uint32_t*>(&b2);
        }
        ptx::tmem_st_32x32b_x16(td + c * 16, pk);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(dst_ready);
    }
    const bool row_ok = gk < p.Sk;
    __nv_bfloat16* dvrow = p.dV + int64_t(b) * p.dv_sb + int64_t(gk) * p.dv_ss + int64_t(hk) * p.dv_sh;
    __nv_bfloat16* dkrow = p.dK + int64_t(b) * p.dk_sb + int64_t(gk) * p.dk_ss + int64_t(hk) * p.dk_sh;
    if (n_it > 0) {
      ptx::mbar_wait(acc_done, 0);
      ptx::tc_fence_after();
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        __nv_bfloat16* orow = which == 0 ? dvrow : dkrow;
        const uint32_t col = which == 0 ? DV_COL : DK_COL;
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
          uint32_t orr[32];
          ptx::tmem_ld_32x32b_x32(tmem_base + lane_base + col + c * 32, orr);
          ptx::tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int t4 = 0; t4 < 4; ++t4) {
              uint4 o;
              __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int t = 0; t < 4; ++t)
                o2[t] = __floats2bfloat162_rn(__uint_as_float(orr[t4 * 8 + t * 2]), __uint_as_float(orr[t4 * 8 + t * 2 + 1]));
              reinterpret_cast<uint4*>(orow + c * 32)[t4] = o;
            }
          }
        }
      }
    } else if (row_ok) {
      for (int c = 0; c < D / 8; ++c) {
        reinterpret_cast<uint4*>(dvrow)[c] = make_uint4(0u, 0u, 0u, 0u);
        reinterpret_cast<uint4*>(dkrow)[c] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------
inline bool make_attn_tmap_bwd(CUtensorMap* out, const AttnTensor& t, int D, int H, int S, int B) {
  uint64_t dims[4] = {(uint64_t)D, (uint64_t)(t.h_div ? t.h_slots : H), (uint64_t)S, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)t.stride_h * 2, (uint64_t)t.stride_s * 2, (uint64_t)t.stride_b * 2};
  uint32_t box[4] = {64, 1, 128, 1};
  return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t.ptr, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int D>
cudaError_t attn_bwd_launch(const AttnBwdCall& c, cudaStream_t s, std::atomic<int64_t>* counter) {
  CUtensorMap tq, tdo, tk, tv;
  if (!make_attn_tmap_bwd(&tq, c.q, D, c.Hq, c.Sq, c.B)) return cudaErrorInvalidValue;
  if (!make_attn_tmap_bwd(&tdo, c.d_o, D, c.Hq, c.Sq, c.B)) return cudaErrorInvalidValue;
  if (!make_attn_tmap_bwd(&tk, c.k, D, c.Hkv, c.Sk, c.B)) return cudaErrorInvalidValue;
  if (!make_attn_tmap_bwd(&tv, c.v, D, c.Hkv, c.Sk, c.B)) return cudaErrorInvalidValue;
  {
    const int64_t rows = int64_t(c.B) * c.Hq * c.Sq;
    const int64_t threads = rows * (D / 8);
    attn_delta_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(
        (const __nv_bfloat16*)c.o.ptr, (const __nv_bfloat16*)c.d_o.ptr, c.delta, c.B, c.Hq, c.Sq, D, c.o.stride_b,
        c.o.stride_s, c.o.stride_h, c.d_o.stride_b, c.d_o.stride_s, c.d_o.stride_h);
  }
  AttnBwdParams p;
  p.B = c.B; p.Hq = c.Hq; p.Hkv = c.Hkv; p.Sq = c.Sq; p.Sk = c.Sk;
  p.scale = c.softmax_scale;
  p.scale_log2 = c.softmax_scale * 1.4426950408889634f;
  p.causal = c.causal ? 1 : 0;
  p.causal_off = c.Sk - c.Sq;
  p.LSE = c.lse; p.DELTA = c.delta;
  p.row_start = c.row_start; p.row_end = c.row_end;
  p.dQ = (__nv_bfloat16*)c.dq.ptr; p.dK = (__nv_bfloat16*)c.dk.ptr; p.dV = (__nv_bfloat16*)c.dv.ptr;
  p.q_div = c.q.h_div; p.q_mul = c.q.h_mul;
  p.dq_sb = c.dq.stride_b; p.dq_ss = c.dq.stride_s; p.dq_sh = c.dq.stride_h;
  p.dk_sb = c.dk.stride_b; p.dk_ss = c.dk.stride_s; p.dk_sh = c.dk.stride_h;
  p.dv_sb = c.dv.stride_b; p.dv_ss = c.dv.stride_s; p.dv_sh = c.dv.stride_h;
  constexpr int smem = 6 * 128 * D * 2 + 3072 + 1024 + 256;
  auto kq = attn_bwd_dq_sm100_kernel<D>;
  auto kkv = attn_bwd_dkv_sm100_kernel<D>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  kq<<<dim3((c.Sq + 127) / 128, c.Hq, c.B), 256, smem, s>>>(tq, tdo, tk, tv, p);
  kkv<<<dim3((c.Sk + 127) / 128, c.Hkv, c.B), 256, smem, s>>>(tq, tdo, tk, tv, p);
  counter->fetch_add(3);
  return cudaGetLastError();
}

}  // namespace hb
