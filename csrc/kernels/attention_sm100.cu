// Host launchers for the tcgen05 flash-attention kernels.
#include "attention_sm100.h"

#include <atomic>

#include "attention_bwd_sm100.cuh"
#include "attention_fwd_sm100.cuh"
#include "tma_host.h"

namespace hb {

static std::atomic<int64_t> g_attn_launches{0};
int64_t attn_launch_count() { return g_attn_launches.load(); }

namespace {

// 4-D map (D, H, S, B) with a {64, 1, rows, 1} box; strides in elements.
bool make_attn_tmap(CUtensorMap* out, const AttnTensor& t, int D, int H, int S, int B, int box_rows) {
  uint64_t dims[4] = {(uint64_t)D, (uint64_t)(t.h_div ? t.h_slots : H), (uint64_t)S, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)t.stride_h * 2, (uint64_t)t.stride_s * 2, (uint64_t)t.stride_b * 2};
  uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};
  return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t.ptr, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

bool strides_ok(const AttnTensor& t) {
  return (reinterpret_cast<uintptr_t>(t.ptr) & 15) == 0 && (t.stride_b % 8) == 0 && (t.stride_s % 8) == 0 &&
         (t.stride_h % 8) == 0;
}

template <int D>
cudaError_t launch_fwd(const AttnFwdCall& c, cudaStream_t s) {
  CUtensorMap tq, tk, tv;
  if (!make_attn_tmap(&tq, c.q, D, c.Hq, c.Sq, c.B, 128)) return cudaErrorInvalidValue;
  if (!make_attn_tmap(&tk, c.k, D, c.Hkv, c.Sk, c.B, 128)) return cudaErrorInvalidValue;
  if (!make_attn_tmap(&tv, c.v, D, c.Hkv, c.Sk, c.B, 128)) return cudaErrorInvalidValue;
  AttnFwdParams p;
  p.B = c.B; p.Hq = c.Hq; p.Hkv = c.Hkv; p.Sq = c.Sq; p.Sk = c.Sk;
  p.q_div = c.q.h_div; p.q_mul = c.q.h_mul;
  p.scale_log2 = c.softmax_scale * 1.4426950408889634f;
  p.causal = c.causal ? 1 : 0;
  p.causal_off = c.Sk - c.Sq;
  p.O = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(c.o.ptr));
  p.o_sb = c.o.stride_b; p.o_ss = c.o.stride_s; p.o_sh = c.o.stride_h;
  p.LSE = c.lse;
  p.row_start = c.row_start;
  constexpr int smem = 5 * 128 * D * 2 + 1024 + 256;
  auto kern = attn_fwd_sm100_kernel<D>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((c.Sq + 127) / 128, c.Hq, c.B);
  kern<<<grid, 256, smem, s>>>(tq, tk, tv, p);
  g_attn_launches.fetch_add(1);
  return cudaGetLastError();
}

}  // namespace

cudaError_t attn_fwd(const AttnFwdCall& c, cudaStream_t s) {
  if (c.B == 0 || c.Sq == 0) return cudaSuccess;
  if (!strides_ok(c.q) || !strides_ok(c.k) || !strides_ok(c.v) || !strides_ok(c.o)) return cudaErrorMisalignedAddress;
  if (c.Hkv <= 0 || c.Hq % c.Hkv != 0) return cudaErrorInvalidValue;
  if (c.row_start != nullptr && (c.B != 1 || !c.causal || c.Sq != c.Sk)) return cudaErrorInvalidValue;
  if (c.D == 64) return launch_fwd<64>(c, s);
  if (c.D == 128) return launch_fwd<128>(c, s);
  return cudaErrorInvalidValue;
}

cudaError_t attn_bwd(const AttnBwdCall& c, cudaStream_t s) {
  if (c.B == 0 || c.Sq == 0) return cudaSuccess;
  if (!strides_ok(c.q) || !strides_ok(c.k) || !strides_ok(c.v) || !strides_ok(c.o) || !strides_ok(c.d_o) ||
      !strides_ok(c.dq) || !strides_ok(c.dk) || !strides_ok(c.dv))
    return cudaErrorMisalignedAddress;
  if (c.Hkv <= 0 || c.Hq % c.Hkv != 0) return cudaErrorInvalidValue;
  if ((c.row_start != nullptr) != (c.row_end != nullptr)) return cudaErrorInvalidValue;
  if (c.row_start != nullptr && (c.B != 1 || !c.causal || c.Sq != c.Sk)) return cudaErrorInvalidValue;
  if (c.D == 64) return attn_bwd_launch<64>(c, s, &g_attn_launches);
  if (c.D == 128) return attn_bwd_launch<128>(c, s, &g_attn_launches);
  return cudaErrorInvalidValue;
}

}  // namespace hb
