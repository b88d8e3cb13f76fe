// Host launcher for the tcgen05 bf16 GEMM (see gemm_sm100.cuh for the design).
#include "gemm_sm100.h"

#include <atomic>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "gemm_sm100.cuh"
#include "tma_host.h"

namespace hb {

static std::atomic<int64_t> g_gemm_launches{0};
int64_t gemm_launch_count() { return g_gemm_launches.load(); }

namespace {

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

template <int G, bool AMN, bool BMN, int ST, typename OutT, bool FP8 = false, int BN = 256>
cudaError_t launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  auto kern = gemm_bf16_sm100_kernel<G, AMN, BMN, ST, OutT, FP8, BN>;
  constexpr int smem = gemm_detail::smem_bytes(G, ST, BN);
  static_assert(smem <= 227 * 1024, "shared memory per CTA");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles_m = (p.M + 128 * G - 1) / (128 * G);
  const int tiles_n = (p.N + BN - 1) / BN;
  int clusters = num_sms() / G;
  if (clusters > tiles_m * tiles_n) clusters = tiles_m * tiles_n;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * G);
  cfg.blockDim = dim3(gemm_detail::kNumThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = G;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  g_gemm_launches.fetch_add(1);
  return cudaLaunchKernelEx(&cfg, kern, ta, tb, p);
}

template <int G, int ST, typename OutT, int BN = 256>
cudaError_t dispatch_major(bool amn, bool bmn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                           cudaStream_t s) {
  if (!amn && !bmn) return launch_cfg<G, false, false, ST, OutT, false, BN>(ta, tb, p, s);
  if (!amn && bmn) return launch_cfg<G, false, true, ST, OutT, false, BN>(ta, tb, p, s);
  if (amn && !bmn) return launch_cfg<G, true, false, ST, OutT, false, BN>(ta, tb, p, s);
  return launch_cfg<G, true, true, ST, OutT, false, BN>(ta, tb, p, s);
}

// Persistent grid of `clusters` CTA pairs: fraction of the last wave that does work, for a given tile width.
double wave_efficiency(int M, int N, int bn, int clusters) {
  const int64_t tiles = int64_t((M + 255) / 256) * ((N + bn - 1) / bn);
  const int64_t waves = (tiles + clusters - 1) / clusters;
  return double(tiles) / double(waves * clusters);
}

// 128-column tiles when the 256-column tiling leaves a large part of the last wave idle and the narrower one does not
// (HETU_GEMM_BN = 128 / 256 forces a width, 0 = this heuristic)
int pick_block_n(const GemmCall& c) {
  static const int forced = [] { const char* e = getenv("HETU_GEMM_BN"); return e ? atoi(e) : -1; }();
  if (c.block_n == 128 || c.block_n == 256) return c.block_n;
  if (forced == 128 || forced == 256) return forced;
  if (forced != 0) return 256;            // heuristic is opt-in until measured: HETU_GEMM_BN=0
  const int clusters = num_sms() / 2;
  const double e256 = wave_efficiency(c.M, c.N, 256, clusters), e128 = wave_efficiency(c.M, c.N, 128, clusters);
  return (e256 < 0.92 && e128 > e256 + 0.06) ? 128 : 256;
}

}  // namespace

cudaError_t gemm_bf16(const GemmCall& c, cudaStream_t stream) {
  if (c.M <= 0 || c.N <= 0) return cudaSuccess;
  if (c.K <= 0) return cudaErrorInvalidValue;
  // TMA needs 16-byte aligned bases and row strides.
  if ((reinterpret_cast<uintptr_t>(c.A) & 15) || (reinterpret_cast<uintptr_t>(c.B) & 15)) return cudaErrorMisalignedAddress;
  if (c.fp8 ? ((c.lda & 15) || (c.ldb & 15) || c.a_mn_major || c.b_mn_major) : ((c.lda & 7) || (c.ldb & 7)))
    return cudaErrorMisalignedAddress;
  const int out_elem = c.out == GemmOut::BF16 ? 2 : 4;
  if ((reinterpret_cast<uintptr_t>(c.C) & 15) || ((c.ldc * out_elem) & 15)) return cudaErrorMisalignedAddress;
  if ((c.aux_in || c.aux_out) && (c.ld_aux & 7)) return cudaErrorMisalignedAddress;

  int G = c.cta_group == 0 ? 2 : c.cta_group;
  // the narrow tile exists for the plain bf16 pair kernel only (not fp8, not the fused all-gather / reduce-scatter variants)
  const bool narrow_ok = G == 2 && !c.fp8 && c.ag_src == nullptr && !(c.peer_c != nullptr && c.rows_per_rank > 0);
  const int BN = narrow_ok ? pick_block_n(c) : 256;
  const int load_n = BN / G;

  CUtensorMap ta, tb;
  bool ok;
  if (c.fp8) {
    ok = make_tmap_2d_u8(&ta, c.A, (uint64_t)c.K, (uint64_t)c.M, (uint64_t)c.lda, 128, 128) &&
         make_tmap_2d_u8(&tb, c.B, (uint64_t)c.K, (uint64_t)c.N, (uint64_t)c.ldb, 128, load_n);
    if (!ok) return cudaErrorInvalidValue;
  } else {
  if (!c.a_mn_major) ok = make_tmap_2d_bf16(&ta, c.A, (uint64_t)c.K, (uint64_t)c.M, (uint64_t)c.lda, 64, 128);
  else ok = make_tmap_2d_bf16(&ta, c.A, (uint64_t)c.M, (uint64_t)c.K, (uint64_t)c.lda, 64, 64);
  if (!ok) return cudaErrorInvalidValue;
  if (!c.b_mn_major) ok = make_tmap_2d_bf16(&tb, c.B, (uint64_t)c.K, (uint64_t)c.N, (uint64_t)c.ldb, 64, load_n);
  else ok = make_tmap_2d_bf16(&tb, c.B, (uint64_t)c.N, (uint64_t)c.K, (uint64_t)c.ldb, 64, 64);
  if (!ok) return cudaErrorInvalidValue;
  }

  GemmParams p;
  p.M = c.M; p.N = c.N; p.K = c.K;
  p.C = c.C; p.ldc = c.ldc;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(c.bias);
  p.aux_in = reinterpret_cast<const __nv_bfloat16*>(c.aux_in);
  p.aux_out = reinterpret_cast<__nv_bfloat16*>(c.aux_out);
  p.ld_aux = c.ld_aux;
  p.act = c.act;
  p.aux_mode = c.aux_in ? c.aux_mode : 0;
  p.accumulate = c.accumulate ? 1 : 0;
  p.alpha = c.alpha;
  p.rows_per_rank = 0;
  p.my_rank = c.my_rank;
  for (int i = 0; i < 8; ++i) p.peer_c[i] = nullptr;
  if (c.peer_c != nullptr && c.rows_per_rank > 0) {
    if (c.world > 8 || (c.rows_per_rank % 128) != 0) return cudaErrorInvalidValue;   // a CTA's 128 rows have one owner
    p.rows_per_rank = c.rows_per_rank;
    for (int i = 0; i < c.world; ++i) p.peer_c[i] = c.peer_c[i];
  }

  p.row_scale = c.row_scale; p.col_scale = c.col_scale;
  for (int i = 0; i < 8; ++i) p.ag_src[i] = nullptr;
  p.ag_dst = nullptr; p.ag_flags = nullptr; p.ag_world = 1; p.ag_rank = 0; p.ag_rows_per_rank = 0;
  if (c.ag_src != nullptr && c.ag_world > 1) {
    // A must be contiguous [M, K] K-major; a rank's shard is a whole number of M tiles of the widest config
    if (c.a_mn_major || c.fp8 || c.lda != c.K || c.ag_world > 8 || c.ag_flags == nullptr || (c.ag_rows_per_rank % 256) != 0 ||
        c.ag_rows_per_rank * c.ag_world != c.M || (c.K % 8) != 0)
      return cudaErrorInvalidValue;
    for (int i = 0; i < c.ag_world; ++i) p.ag_src[i] = c.ag_src[i];
    p.ag_dst = const_cast<void*>(c.A); p.ag_flags = c.ag_flags; p.ag_world = c.ag_world; p.ag_rank = c.ag_rank;
    p.ag_rows_per_rank = c.ag_rows_per_rank;
  }
  if (c.fp8) {
    if (G == 2) {
      if (c.out == GemmOut::BF16) return launch_cfg<2, false, false, 6, __nv_bfloat16, true>(ta, tb, p, stream);
      return launch_cfg<2, false, false, 6, float, true>(ta, tb, p, stream);
    }
    if (c.out == GemmOut::BF16) return launch_cfg<1, false, false, 4, __nv_bfloat16, true>(ta, tb, p, stream);
    return launch_cfg<1, false, false, 4, float, true>(ta, tb, p, stream);
  }
  if (G == 2 && BN == 128) {
    if (c.out == GemmOut::BF16) return dispatch_major<2, 8, __nv_bfloat16, 128>(c.a_mn_major, c.b_mn_major, ta, tb, p, stream);
    return dispatch_major<2, 8, float, 128>(c.a_mn_major, c.b_mn_major, ta, tb, p, stream);
  }
  if (G == 2) {
    if (c.out == GemmOut::BF16) return dispatch_major<2, 6, __nv_bfloat16>(c.a_mn_major, c.b_mn_major, ta, tb, p, stream);
    return dispatch_major<2, 6, float>(c.a_mn_major, c.b_mn_major, ta, tb, p, stream);
  } else {
    if (c.out == GemmOut::BF16) return dispatch_major<1, 4, __nv_bfloat16>(c.a_mn_major, c.b_mn_major, ta, tb, p, stream);
    return dispatch_major<1, 4, float>(c.a_mn_major, c.b_mn_major, ta, tb, p, stream);
  }
}

}  // namespace hb
