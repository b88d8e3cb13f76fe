// LayerNorm / RMSNorm forward + backward for bf16 activations (fp32 statistics).
//
// Forward: one 128-thread CTA per row, the row lives in registers between the
// statistics pass and the normalise pass (one HBM read, one write).
// Backward: persistent CTAs stride over rows, keep per-column dgamma/dbeta
// partials in registers, and a tiny second kernel folds the per-CTA partials,
// so dx needs exactly one read of (x, dy) and one write.
//
// Capability parity: hetu/impl/kernel/FusedLayerNorm.cu:456-1004 (cuApplyLayerNorm,
// cuApplyRMSNorm, cuComputePartGradGammaBeta, cuComputeGradInput).
#include "common.cuh"
#include "kernels.h"

namespace hb {

std::atomic<int64_t> g_kernel_launches{0};
int64_t kernel_launch_count() { return g_kernel_launches.load(); }

namespace {

constexpr int kFwdThreads = 128;
constexpr int kFwdMaxV = 8;   // up to 8192 columns
constexpr int kBwdThreads = 256;

template <bool kRMS>
__global__ void __launch_bounds__(kFwdThreads) norm_fwd_kernel(const void* __restrict__ x, const void* __restrict__ gamma,
                                                               const void* __restrict__ beta, void* __restrict__ y,
                                                               float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                               int cols, float eps) {
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const int nvec = cols >> 3;
  const int tid = threadIdx.x;
  const char* xr = reinterpret_cast<const char*>(x) + row * int64_t(cols) * 2;
  char* yr = reinterpret_cast<char*>(y) + row * int64_t(cols) * 2;
  float xv[kFwdMaxV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kFwdMaxV; ++i) {
    const int v = tid + i * kFwdThreads;
    if (v < nvec) {
      unpack8(ld8_stream(xr, v), xv[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += kRMS ? xv[i][j] * xv[i][j] : xv[i][j];
    }
  }
  float mean = 0.f, rstd;
  if constexpr (kRMS) {
    const float ms = block_sum(sum, red) / cols;
    rstd = rsqrtf(ms + eps);
  } else {
    mean = block_sum(sum, red) / cols;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < kFwdMaxV; ++i) {
      const int v = tid + i * kFwdThreads;
      if (v < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = xv[i][j] - mean; var += d * d; }
      }
    }
    var = block_sum(var, red) / cols;
    rstd = rsqrtf(var + eps);
  }
  if (tid == 0) {
    if (mean_out) mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
#pragma unroll
  for (int i = 0; i < kFwdMaxV; ++i) {
    const int v = tid + i * kFwdThreads;
    if (v < nvec) {
      float g[8], b[8], o[8];
      unpack8(ld8(gamma, v), g);
      if (!kRMS && beta != nullptr) unpack8(ld8(beta, v), b);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (xv[i][j] - mean) * rstd * g[j] + b[j];
      st8(yr, v, pack8(o));
    }
  }
}

// Warp-per-row forward for cols <= 4096: no block barrier at all, the row lives in registers (kV vectors of 8 per lane),
// 4 rows per CTA and many CTAs per SM keep enough loads in flight to saturate HBM.
template <bool kRMS, int kV>
__global__ void __launch_bounds__(128) norm_fwd_warp_kernel(const void* __restrict__ x, const void* __restrict__ gamma,
                                                            const void* __restrict__ beta, void* __restrict__ y,
                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                            int64_t rows, int cols, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = cols >> 3;
  const char* xr = reinterpret_cast<const char*>(x) + row * int64_t(cols) * 2;
  char* yr = reinterpret_cast<char*>(y) + row * int64_t(cols) * 2;
  float xv[kV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      unpack8(ld8_stream(xr, v), xv[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += kRMS ? xv[i][j] * xv[i][j] : xv[i][j];
    }
  }
  float mean = 0.f, rstd;
  const float inv = 1.0f / cols;
  if constexpr (kRMS) {
    rstd = rsqrtf(warp_sum(sum) * inv + eps);
  } else {
    mean = warp_sum(sum) * inv;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = xv[i][j] - mean; var += d * d; }
      }
    }
    rstd = rsqrtf(warp_sum(var) * inv + eps);
  }
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      float g[8], b[8], o[8];
      unpack8(ld8(gamma, v), g);
      if (!kRMS && beta != nullptr) unpack8(ld8(beta, v), b);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (xv[i][j] - mean) * rstd * g[j] + b[j];
      st8(yr, v, pack8(o));
    }
  }
}

// Generic fallback (any column count): block per row, scalar loops.
template <bool kRMS>
__global__ void norm_fwd_generic_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gamma,
                                        const __nv_bfloat16* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                        float* mean_out, float* rstd_out, int cols, float eps) {
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const __nv_bfloat16* xr = x + row * cols;
  float s = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float v = __bfloat162float(xr[c]);
    s += kRMS ? v * v : v;
  }
  float mean = 0.f, rstd;
  if constexpr (kRMS) {
    rstd = rsqrtf(block_sum(s, red) / cols + eps);
  } else {
    mean = block_sum(s, red) / cols;
    float var = 0.f;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
      const float d = __bfloat162float(xr[c]) - mean;
      var += d * d;
    }
    rstd = rsqrtf(block_sum(var, red) / cols + eps);
  }
  if (threadIdx.x == 0) {
    if (mean_out) mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float v = (__bfloat162float(xr[c]) - mean) * rstd * __bfloat162float(gamma[c]) +
                    ((!kRMS && beta) ? __bfloat162float(beta[c]) : 0.f);
    y[row * cols + c] = __float2bfloat16(v);
  }
}

__device__ __forceinline__ float2 block_sum2(float2 v, float2* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v.x = warp_sum(v.x);
  v.y = warp_sum(v.y);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float2 r = (lane < nw) ? red[lane] : make_float2(0.f, 0.f);
  r.x = warp_sum(r.x);
  r.y = warp_sum(r.y);
  return r;
}

// kVPT vectors (of 8 columns) per thread, kR rows per iteration: kR*kVPT*2 16-byte loads are in flight per thread and
// the 2*kR row statistics are reduced with ONE block barrier per iteration (double-buffered scratch).
template <bool kRMS, int kVPT, int kR>
__global__ void __launch_bounds__(kBwdThreads) norm_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ x,
                                                               const void* __restrict__ gamma, const float* __restrict__ mean,
                                                               const float* __restrict__ rstd, void* __restrict__ dx,
                                                               float* __restrict__ part_dg, float* __restrict__ part_db,
                                                               int64_t rows, int cols, const void* __restrict__ dx_add) {
  constexpr int kWarps = kBwdThreads / 32;
  __shared__ float red[2][kWarps][2 * kR];
  const int nvec = cols >> 3;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float g[kVPT][8], dg[kVPT][8], db[kVPT][8];
#pragma unroll
  for (int i = 0; i < kVPT; ++i) {
    const int v = tid + i * kBwdThreads;
#pragma unroll
    for (int j = 0; j < 8; ++j) { dg[i][j] = 0.f; db[i][j] = 0.f; g[i][j] = 0.f; }
    if (v < nvec) unpack8(ld8(gamma, v), g[i]);
  }
  const float inv_cols = 1.0f / cols;
  int buf = 0;
  for (int64_t row0 = int64_t(blockIdx.x) * kR; row0 < rows; row0 += int64_t(gridDim.x) * kR) {
    float xh[kR][kVPT][8], dyv[kR][kVPT][8];
    float s1[kR], s2[kR], rs[kR];
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      const int64_t row = row0 + r;
      const bool live = row < rows;
      const char* xr = reinterpret_cast<const char*>(x) + row * int64_t(cols) * 2;
      const char* dyr = reinterpret_cast<const char*>(dy) + row * int64_t(cols) * 2;
#pragma unroll
      for (int i = 0; i < kVPT; ++i) {
        const int v = tid + i * kBwdThreads;
        if (live && v < nvec) {
          unpack8(ld8_stream(xr, v), xh[r][i]);
          unpack8(ld8_stream(dyr, v), dyv[r][i]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) { xh[r][i][j] = 0.f; dyv[r][i][j] = 0.f; }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      const int64_t row = row0 + r;
      const bool live = row < rows;
      const float mu = (kRMS || !live) ? 0.f : mean[row];
      rs[r] = live ? rstd[row] : 0.f;
      s1[r] = 0.f; s2[r] = 0.f;
#pragma unroll
      for (int i = 0; i < kVPT; ++i) {
        const int v = tid + i * kBwdThreads;
        if (v < nvec) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xn = (xh[r][i][j] - mu) * rs[r];
            xh[r][i][j] = xn;
            const float dyg = dyv[r][i][j] * g[i][j];
            s1[r] += dyg;
            s2[r] += dyg * xn;
            dg[i][j] += dyv[r][i][j] * xn;
            db[i][j] += dyv[r][i][j];
          }
        }
      }
      s1[r] = warp_sum(s1[r]);
      s2[r] = warp_sum(s2[r]);
    }
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < kR; ++r) { red[buf][warp][2 * r] = s1[r]; red[buf][warp][2 * r + 1] = s2[r]; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) { a += red[buf][w][2 * r]; b += red[buf][w][2 * r + 1]; }
      const float m1 = kRMS ? 0.f : a * inv_cols;
      const float m2 = b * inv_cols;
      const int64_t row = row0 + r;
      if (row < rows) {
        char* dxr = reinterpret_cast<char*>(dx) + row * int64_t(cols) * 2;
#pragma unroll
        for (int i = 0; i < kVPT; ++i) {
          const int v = tid + i * kBwdThreads;
          if (v < nvec) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = rs[r] * (dyv[r][i][j] * g[i][j] - m1 - xh[r][i][j] * m2);
            if (dx_add != nullptr) {
              // gradient arriving through the residual stream: d(input) = d(norm branch) + d(skip), added here instead of a
              // separate elementwise pass over the activation
              float e[8];
              unpack8(ld8_stream(reinterpret_cast<const char*>(dx_add) + row * int64_t(cols) * 2, v), e);
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] += e[j];
            }
            st8(dxr, v, pack8(o));
          }
        }
      }
    }
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < kVPT; ++i) {
    const int v = tid + i * kBwdThreads;
    if (v < nvec) {
      float* pg = part_dg + int64_t(blockIdx.x) * cols + v * 8;
      reinterpret_cast<float4*>(pg)[0] = make_float4(dg[i][0], dg[i][1], dg[i][2], dg[i][3]);
      reinterpret_cast<float4*>(pg)[1] = make_float4(dg[i][4], dg[i][5], dg[i][6], dg[i][7]);
      if (!kRMS) {
        float* pb = part_db + int64_t(blockIdx.x) * cols + v * 8;
        reinterpret_cast<float4*>(pb)[0] = make_float4(db[i][0], db[i][1], db[i][2], db[i][3]);
        reinterpret_cast<float4*>(pb)[1] = make_float4(db[i][4], db[i][5], db[i][6], db[i][7]);
      }
    }
  }
}

template <bool kRMS>
__global__ void norm_bwd_generic_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                        const __nv_bfloat16* __restrict__ gamma, const float* __restrict__ mean,
                                        const float* __restrict__ rstd, __nv_bfloat16* __restrict__ dx,
                                        float* part_dg, float* part_db, int64_t rows, int cols, const __nv_bfloat16* __restrict__ dx_add) {
  // one block per "part"; columns strided over threads; rows strided over blocks
  __shared__ float2 red[32];
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    part_dg[int64_t(blockIdx.x) * cols + c] = 0.f;
    if (!kRMS) part_db[int64_t(blockIdx.x) * cols + c] = 0.f;
  }
  __syncthreads();
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const float mu = kRMS ? 0.f : mean[row];
    const float rs = rstd[row];
    float2 s = make_float2(0.f, 0.f);
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
      const float xh = (__bfloat162float(x[row * cols + c]) - mu) * rs;
      const float d = __bfloat162float(dy[row * cols + c]);
      const float dyg = d * __bfloat162float(gamma[c]);
      s.x += dyg;
      s.y += dyg * xh;
      part_dg[int64_t(blockIdx.x) * cols + c] += d * xh;
      if (!kRMS) part_db[int64_t(blockIdx.x) * cols + c] += d;
    }
    s = block_sum2(s, red);
    const float m1 = kRMS ? 0.f : s.x / cols, m2 = s.y / cols;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
      const float xh = (__bfloat162float(x[row * cols + c]) - mu) * rs;
      const float dyg = __bfloat162float(dy[row * cols + c]) * __bfloat162float(gamma[c]);
      dx[row * cols + c] = __float2bfloat16(rs * (dyg - m1 - xh * m2) + (dx_add ? __bfloat162float(dx_add[row * cols + c]) : 0.f));
    }
  }
}

// Folds the per-CTA partials: block (32 columns, 8 part-lanes), coalesced 128-byte rows, 8 independent chains per
// column and BOTH outputs (dgamma, dbeta) in one launch (blockIdx.y).
__global__ void __launch_bounds__(256) fold_parts_kernel(const float* __restrict__ parts0, float* __restrict__ out0,
                                                         const float* __restrict__ parts1, float* __restrict__ out1,
                                                         int nparts, int cols, int accumulate,
                                                         __nv_bfloat16* __restrict__ outb0, __nv_bfloat16* __restrict__ outb1) {
  __shared__ float sm[8][33];
  const float* parts = blockIdx.y == 0 ? parts0 : parts1;
  float* out = blockIdx.y == 0 ? out0 : out1;
  __nv_bfloat16* outb = blockIdx.y == 0 ? outb0 : outb1;      // bf16 result written directly (parameter-gradient dtype)
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < cols) {
    int p = ty;
    for (; p + 24 < nparts; p += 32) {
      s0 += parts[int64_t(p) * cols + c];
      s1 += parts[int64_t(p + 8) * cols + c];
      s2 += parts[int64_t(p + 16) * cols + c];
      s3 += parts[int64_t(p + 24) * cols + c];
    }
    for (; p < nparts; p += 8) s0 += parts[int64_t(p) * cols + c];
  }
  sm[ty][tx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (ty == 0 && c < cols) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sm[w][tx];
    if (outb != nullptr) outb[c] = __float2bfloat16(s);
    else out[c] = accumulate ? out[c] + s : s;
  }
}

}  // namespace

int ln_bwd_parts() { return sm_count() * 2; }

template <bool kRMS>
static cudaError_t norm_fwd_impl(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd,
                                 int64_t rows, int cols, float eps, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  const unsigned wgrid = (unsigned)((rows + 3) / 4);
  if ((cols & 7) == 0 && cols <= 512) {
    norm_fwd_warp_kernel<kRMS, 2><<<wgrid, 128, 0, s>>>(x, gamma, beta, y, mean, rstd, rows, cols, eps);
  } else if ((cols & 7) == 0 && cols <= 1024) {
    norm_fwd_warp_kernel<kRMS, 4><<<wgrid, 128, 0, s>>>(x, gamma, beta, y, mean, rstd, rows, cols, eps);
  } else if ((cols & 7) == 0 && cols <= 2048) {
    norm_fwd_warp_kernel<kRMS, 8><<<wgrid, 128, 0, s>>>(x, gamma, beta, y, mean, rstd, rows, cols, eps);
  } else if ((cols & 7) == 0 && cols <= 4096) {
    norm_fwd_warp_kernel<kRMS, 16><<<wgrid, 128, 0, s>>>(x, gamma, beta, y, mean, rstd, rows, cols, eps);
  } else if ((cols & 7) == 0 && cols <= kFwdMaxV * kFwdThreads * 8) {
    norm_fwd_kernel<kRMS><<<(unsigned)rows, kFwdThreads, 0, s>>>(x, gamma, beta, y, mean, rstd, cols, eps);
  } else {
    norm_fwd_generic_kernel<kRMS><<<(unsigned)rows, 256, 0, s>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)gamma, (const __nv_bfloat16*)beta, (__nv_bfloat16*)y, mean, rstd,
        cols, eps);
  }
  count_launch();
  return cudaGetLastError();
}

template <bool kRMS>
static cudaError_t norm_bwd_impl(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
                                 void* dx, float* dgamma, float* dbeta, float* ws, int64_t rows, int cols,
                                 bool accumulate, cudaStream_t s, const void* dx_add, void* dgamma_bf16 = nullptr,
                                 void* dbeta_bf16 = nullptr) {
  if (rows == 0) return cudaSuccess;
  int parts = ln_bwd_parts();
  if (parts > rows) parts = (int)rows;
  float* part_dg = ws;
  float* part_db = ws + int64_t(ln_bwd_parts()) * cols;
  if ((cols & 7) == 0 && cols <= 2048) {
    norm_bwd_kernel<kRMS, 1, 4><<<parts, kBwdThreads, 0, s>>>(dy, x, gamma, mean, rstd, dx, part_dg, part_db, rows, cols, dx_add);
  } else if ((cols & 7) == 0 && cols <= 4096) {
    norm_bwd_kernel<kRMS, 2, 2><<<parts, kBwdThreads, 0, s>>>(dy, x, gamma, mean, rstd, dx, part_dg, part_db, rows, cols, dx_add);
  } else if ((cols & 7) == 0 && cols <= 8192) {
    norm_bwd_kernel<kRMS, 4, 1><<<parts, kBwdThreads, 0, s>>>(dy, x, gamma, mean, rstd, dx, part_dg, part_db, rows, cols, dx_add);
  } else {
    norm_bwd_generic_kernel<kRMS><<<parts, 256, 0, s>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x,
                                                        (const __nv_bfloat16*)gamma, mean, rstd, (__nv_bfloat16*)dx,
                                                        part_dg, part_db, rows, cols, (const __nv_bfloat16*)dx_add);
  }
  const bool two = !kRMS && (dbeta != nullptr || dbeta_bf16 != nullptr);
  fold_parts_kernel<<<dim3((cols + 31) / 32, two ? 2 : 1), 256, 0, s>>>(part_dg, dgamma, part_db, dbeta, parts, cols,
                                                                        accumulate ? 1 : 0, (__nv_bfloat16*)dgamma_bf16,
                                                                        (__nv_bfloat16*)dbeta_bf16);
  count_launch(2);
  return cudaGetLastError();
}

// Fused dropout + residual add + LayerNorm / RMSNorm forward (pre-norm transformer block glue):
//     z = residual + dropout(x)          (written out: it is the next block's residual stream)
//     y = norm(z) * gamma (+ beta)
// one pass over x / residual instead of three kernels; warp per row, the row stays in registers.
// The dropout mask uses the same Philox counters as the standalone dropout kernel (vector index of the flattened
// tensor + offset), so the backward re-creates it with that kernel.
// (ref: hetu/impl/kernel/RMSNorm.cu:90 DropoutAddLnFwdCuda, :257 DropoutAddLnBwd, FlashAttention's layer_norm library)
template <bool kRMS, int kV>
__global__ void __launch_bounds__(128) dropout_add_norm_fwd_kernel(const void* __restrict__ x, const void* __restrict__ residual,
                                                                   const void* __restrict__ gamma, const void* __restrict__ beta,
                                                                   void* __restrict__ y, void* __restrict__ z,
                                                                   float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                                   int64_t rows, int cols, float eps, float p, uint64_t seed,
                                                                   uint64_t offset) {
  const int lane = threadIdx.x & 31;
  const int64_t row = int64_t(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = cols >> 3;
  const int64_t rbytes = row * int64_t(cols) * 2;
  const char* xr = reinterpret_cast<const char*>(x) + rbytes;
  const char* rr = residual ? reinterpret_cast<const char*>(residual) + rbytes : nullptr;
  char* yr = reinterpret_cast<char*>(y) + rbytes;
  char* zr = reinterpret_cast<char*>(z) + rbytes;
  const float scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  const uint32_t thresh = (uint32_t)(p * 65536.0f);
  float zv[kV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      unpack8(ld8_stream(xr, v), zv[i]);
      if (p > 0.f) {
        bool keep[8];
        dropout_keep8(uint64_t(row) * nvec + v + offset, seed, thresh, keep);
#pragma unroll
        for (int j = 0; j < 8; ++j) zv[i][j] = keep[j] ? zv[i][j] * scale : 0.f;
      }
      if (rr) {
        float r8[8];
        unpack8(ld8_stream(rr, v), r8);
#pragma unroll
        for (int j = 0; j < 8; ++j) zv[i][j] += r8[j];
      }
      // the residual stream is stored in bf16: normalise exactly what the next block will read
      const bf16x8 packed = pack8(zv[i]);
      st8(zr, v, packed);
      unpack8(packed, zv[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += kRMS ? zv[i][j] * zv[i][j] : zv[i][j];
    }
  }
  float mean = 0.f, rstd;
  const float inv = 1.0f / cols;
  if constexpr (kRMS) {
    rstd = rsqrtf(warp_sum(sum) * inv + eps);
  } else {
    mean = warp_sum(sum) * inv;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const int v = lane + i * 32;
      if (v < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = zv[i][j] - mean; var += d * d; }
      }
    }
    rstd = rsqrtf(warp_sum(var) * inv + eps);
  }
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      float g[8], b[8], o[8];
      unpack8(ld8(gamma, v), g);
      if (!kRMS && beta != nullptr) unpack8(ld8(beta, v), b);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (zv[i][j] - mean) * rstd * g[j] + b[j];
      st8(yr, v, pack8(o));
    }
  }
}

template <bool kRMS>
cudaError_t dropout_add_norm_impl(const void* x, const void* residual, const void* gamma, const void* beta, void* y, void* z,
                                  float* mean, float* rstd, int64_t rows, int cols, float eps, float p, uint64_t seed,
                                  uint64_t offset, cudaStream_t s) {
  if (rows == 0) return cudaSuccess;
  if ((cols & 7) || cols > 8192) return cudaErrorInvalidValue;
  const unsigned grid = (unsigned)((rows + 3) / 4);
  const int nvec = cols >> 3;
#define HB_DAN(V) dropout_add_norm_fwd_kernel<kRMS, V><<<grid, 128, 0, s>>>(x, residual, gamma, beta, y, z, mean, rstd, rows, cols, eps, p, seed, offset)
  if (nvec <= 32) HB_DAN(1);
  else if (nvec <= 64) HB_DAN(2);
  else if (nvec <= 128) HB_DAN(4);
  else if (nvec <= 256) HB_DAN(8);
  else if (nvec <= 512) HB_DAN(16);
  else HB_DAN(32);
#undef HB_DAN
  count_launch();
  return cudaGetLastError();
}

cudaError_t layernorm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean, float* rstd,
                          int64_t rows, int cols, float eps, cudaStream_t s) {
  return norm_fwd_impl<false>(x, gamma, beta, y, mean, rstd, rows, cols, eps, s);
}
cudaError_t layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd,
                          void* dx, float* dgamma, float* dbeta, float* workspace, int64_t rows, int cols,
                          bool accumulate, cudaStream_t s, const void* dx_add, void* dgamma_bf16, void* dbeta_bf16) {
  return norm_bwd_impl<false>(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, workspace, rows, cols, accumulate, s, dx_add, dgamma_bf16,
                              dbeta_bf16);
}
cudaError_t rmsnorm_fwd(const void* x, const void* gamma, void* y, float* rstd, int64_t rows, int cols, float eps,
                        cudaStream_t s) {
  return norm_fwd_impl<true>(x, gamma, nullptr, y, nullptr, rstd, rows, cols, eps, s);
}
cudaError_t rmsnorm_bwd(const void* dy, const void* x, const void* gamma, const float* rstd, void* dx, float* dgamma,
                        float* workspace, int64_t rows, int cols, bool accumulate, cudaStream_t s, const void* dx_add,
                        void* dgamma_bf16) {
  return norm_bwd_impl<true>(dy, x, gamma, nullptr, rstd, dx, dgamma, nullptr, workspace, rows, cols, accumulate, s, dx_add, dgamma_bf16,
                             nullptr);
}

cudaError_t dropout_add_layernorm_fwd(const void* x, const void* residual, const void* gamma, const void* beta, void* y, void* z,
                                      float* mean, float* rstd, int64_t rows, int cols, float eps, float p, uint64_t seed,
                                      uint64_t offset, cudaStream_t s) {
  return dropout_add_norm_impl<false>(x, residual, gamma, beta, y, z, mean, rstd, rows, cols, eps, p, seed, offset, s);
}
cudaError_t dropout_add_rmsnorm_fwd(const void* x, const void* residual, const void* gamma, void* y, void* z, float* rstd,
                                    int64_t rows, int cols, float eps, float p, uint64_t seed, uint64_t offset, cudaStream_t s) {
  return dropout_add_norm_impl<true>(x, residual, gamma, nullptr, y, z, nullptr, rstd, rows, cols, eps, p, seed, offset, s);
}

}  // namespace hb
