// Flash-attention forward for sm_100a: S = Q K^T and O += P V both run on
// tcgen05 tensor cores with accumulators in TMEM; P never leaves tensor memory
// (the softmax warps overwrite the S columns with bf16 P and the PV MMA reads its
// A operand straight from TMEM).  Q/K/V tiles arrive by TMA into 128B-swizzled
// shared memory; K and V are double buffered.
//
// Warp roles (256 threads): warp 0 TMA producer, warp 1 MMA issuer (one thread),
// warp 2 TMEM allocator, warps 4..7 softmax + lazy O rescale + epilogue (thread i
// owns query row i, matching the 32x32b TMEM access pattern).
//
// Online softmax uses a *lazy* reference maximum: the running max is only
// refreshed (and O rescaled in TMEM) when the new row max exceeds it by more
// than 2^8, so in steady state the O accumulator is never touched by CUDA cores.
//
// Capability parity: hetu/impl/kernel/FlashAttention.cu:215 (FlashAttnCuda ->
// run_mha_fwd_), causal, GQA, LSE output.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "ptx.cuh"

namespace hb {

struct AttnFwdParams {
  int q_div, q_mul;   // grouped-layout slot mapping of Q heads (q_div == 0: identity)
  int B, Hq, Hkv, Sq, Sk;
  float scale_log2;  // softmax_scale * log2(e)
  int causal;
  int causal_off;    // Sk - Sq
  __nv_bfloat16* O;
  int64_t o_sb, o_ss, o_sh;
  float* LSE;        // [B, Hq, Sq]
  // packed variable-length rows in ONE launch (B == 1, causal): row_start[t] = first token of the document that token t
  // belongs to.  A query attends to keys in [row_start[q], q]; KV tiles wholly before the first row's document are skipped.
  const int* row_start = nullptr;
};

namespace attn_detail {
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr int kBoxBytes = 128 * 128;  // 128 rows x 64 bf16 (one 128B-swizzled TMA box)
}  // namespace attn_detail

template <int D>
__global__ void __launch_bounds__(256, 1)
attn_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                      const __grid_constant__ CUtensorMap tmap_v, const AttnFwdParams p) {
  using namespace attn_detail;
  constexpr int NBOX = D / 64;
  constexpr int TILE_BYTES = 128 * D * 2;
  constexpr int O_COL = 256;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE_BYTES;          // 2 stages
  uint8_t* sV = smem + 3 * TILE_BYTES;      // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* p_ready = bars + 11;  // [2]
  uint64_t* o_done = bars + 13;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = gridDim.x - 1 - blockIdx.x;  // heavy (late) causal tiles first
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int hk = h / (p.Hq / p.Hkv);
  const int hslot = p.q_div ? (h / p.q_div) * p.q_mul + (h % p.q_div) : h;
  const int q0 = qt * 128;

  int n_kv_end = (p.Sk + 127) / 128;
  if (p.causal) {
    const int last = q0 + 127 + p.causal_off;  // last visible key column for this tile
    const int lim = last < 0 ? 0 : last / 128 + 1;
    n_kv_end = min(n_kv_end, lim);
  }
  // varlen: documents are contiguous, so the first row of the tile has the smallest document start
  const int j_begin = (p.row_start != nullptr) ? min(p.row_start[min(q0, p.Sq - 1)] / 128, n_kv_end) : 0;
  const int n_kv = n_kv_end - j_begin;          // number of KV tiles this CTA visits: tile index = j_begin + j

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_q);
    ptx::prefetch_tmap(&tmap_k);
    ptx::prefetch_tmap(&tmap_v);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&k_full[i], 1);
      ptx::mbar_init(&k_empty[i], 1);
      ptx::mbar_init(&v_full[i], 1);
      ptx::mbar_init(&v_empty[i], 1);
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&p_ready[i], 4);
    }
    ptx::mbar_init(o_done, 1);
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
      ptx::mbar_arrive_expect_tx(q_full, TILE_BYTES);
      for (int bx = 0; bx < NBOX; ++bx) ptx::tma_load_4d(sQ + bx * kBoxBytes, &tmap_q, q_full, bx * 64, hslot, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        ptx::mbar_wait(&k_empty[st], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
        for (int bx = 0; bx < NBOX; ++bx)
          ptx::tma_load_4d(sK + st * TILE_BYTES + bx * kBoxBytes, &tmap_k, &k_full[st], bx * 64, hk, (j_begin + j) * 128, b);
        ptx::mbar_wait(&v_empty[st], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
        for (int bx = 0; bx < NBOX; ++bx)
          ptx::tma_load_4d(sV + st * TILE_BYTES + bx * kBoxBytes, &tmap_v, &v_full[st], bx * 64, hk, (j_begin + j) * 128, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && n_kv > 0) {
      constexpr uint32_t idesc_s = ptx::make_idesc(128, 128, 1, 1, false, false);
      constexpr uint32_t idesc_pv = ptx::make_idesc(128, D, 1, 1, false, true);
      const uint32_t q_addr = ptx::smem_u32(sQ);
      const uint32_t k_addr = ptx::smem_u32(sK);
      const uint32_t v_addr = ptx::smem_u32(sV);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        ptx::mbar_wait(&k_full[st], (j >> 1) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk >> 2) * kBoxBytes + (kk & 3) * 32;
          const uint64_t a = ptx::make_smem_desc_sw128(q_addr + off, 0, 1024);
          const uint64_t bd = ptx::make_smem_desc_sw128(k_addr + st * TILE_BYTES + off, 0, 1024);
          ptx::mma_f16_ss<1>(tmem_base + st * 128, a, bd, idesc_s, kk > 0 ? 1u : 0u);
        }
        ptx::mma_commit<1>(&k_empty[st]);
        ptx::mma_commit<1>(&s_full[st]);
      };
      ptx::mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) issue_s(j + 1);
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        ptx::mbar_wait(&v_full[st], ph);
        ptx::mbar_wait(&p_ready[st], ph);
        ptx::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t bd = ptx::make_smem_desc_sw128(v_addr + st * TILE_BYTES + kk * 2048, kBoxBytes, 1024);
          ptx::mma_f16_ts<1>(tmem_base + O_COL, tmem_base + st * 128 + kk * 8, bd, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
        }
        ptx::mma_commit<1>(&v_empty[st]);
        ptx::mma_commit<1>(o_done);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const int grow = q0 + row;
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    float m_ref = -INFINITY;
    float l = 0.f;
    const int doc_lo = (p.row_start != nullptr) ? p.row_start[min(grow, p.Sq - 1)] : 0;   // first visible key of this row
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
      const int k0 = (j_begin + j) * 128;
      const bool need_mask = (k0 + 128 > p.Sk) || (p.causal && (k0 + 127 > q0 + p.causal_off)) || (k0 < doc_lo);
      if (need_mask) {
        const int lim = p.causal ? min(p.Sk - 1, grow + p.causal_off) : p.Sk - 1;  // last visible column
#pragma unroll
        for (int c = 0; c < 128; ++c)
          if (k0 + c > lim || k0 + c < doc_lo) sr[c] = 0xff800000u;  // -inf
      }
      float mx = __uint_as_float(sr[0]);
#pragma unroll
      for (int c = 1; c < 128; ++c) mx = fmaxf(mx, __uint_as_float(sr[c]));
      mx *= p.scale_log2;
      if (j == 0) {
        m_ref = mx;
      } else {
        const bool need = mx > m_ref + 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          // PV(j-1) must have retired before O is touched
          ptx::mbar_wait(o_done, (j - 1) & 1);
          ptx::tc_fence_after();
          const float alpha = need ? ex2(m_ref - mx) : 1.0f;
          if (need) m_ref = mx;
          l *= alpha;
#pragma unroll
          for (int c = 0; c < D / 32; ++c) {
            uint32_t orr[32];
            ptx::tmem_ld_32x32b_x32(tmem_base + lane_base + O_COL + c * 32, orr);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 32; ++t) orr[t] = __float_as_uint(__uint_as_float(orr[t]) * alpha);
            ptx::tmem_st_32x32b_x32(tmem_base + lane_base + O_COL + c * 32, orr);
          }
          ptx::tmem_st_wait();
        }
      }
      const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
      uint32_t pk[64];
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        const float p0 = ex2(fmaf(__uint_as_float(sr[2 * c]), p.scale_log2, -m_use));
        const float p1 = ex2(fmaf(__uint_as_float(sr[2 * c + 1]), p.scale_log2, -m_use));
        lsum += p0 + p1;
        const __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
        pk[c] = *reinterpret_cast<const uint32_t*>(&pb);
      }
      l += lsum;
      ptx::tmem_st_32x32b_x32(taddr, pk);
      ptx::tmem_st_32x32b_x32(taddr + 32, pk + 32);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_ready[st]);
    }
    // epilogue: O / l -> bf16 -> global, LSE
    const bool row_ok = grow < p.Sq;
    __nv_bfloat16* orow = p.O + int64_t(b) * p.o_sb + int64_t(grow) * p.o_ss + int64_t(h) * p.o_sh;
    if (n_kv > 0) {
      ptx::mbar_wait(o_done, (n_kv - 1) & 1);
      ptx::tc_fence_after();
      const float inv_l = l > 0.f ? 1.0f / l : 0.f;
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint32_t orr[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + lane_base + O_COL + c * 32, orr);
        ptx::tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int t4 = 0; t4 < 4; ++t4) {
            uint4 o;
            __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int t = 0; t < 4; ++t)
              o2[t] = __floats2bfloat162_rn(__uint_as_float(orr[t4 * 8 + t * 2]) * inv_l,
                                            __uint_as_float(orr[t4 * 8 + t * 2 + 1]) * inv_l);
            reinterpret_cast<uint4*>(orow + c * 32)[t4] = o;
          }
        }
      }
      if (row_ok && p.LSE)
        p.LSE[(int64_t(b) * p.Hq + h) * p.Sq + grow] =
            (l > 0.f) ? (m_ref + log2f(l)) * 0.6931471805599453f : -INFINITY;
    } else if (row_ok) {
      for (int c = 0; c < D / 8; ++c) reinterpret_cast<uint4*>(orow)[c] = make_uint4(0u, 0u, 0u, 0u);
      if (p.LSE) p.LSE[(int64_t(b) * p.Hq + h) * p.Sq + grow] = -INFINITY;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, 512);
  }
}

}  // namespace hb
