// Generic (long-tail) device kernels: any-rank strided copy, broadcasting binary arithmetic, the unary function family,
// reductions over [outer, reduce, inner] views, softmax / log-softmax along any dimension, dtype casts and fills.
// They are bandwidth kernels: 16-byte accesses where the layout allows, fp32 arithmetic for 16-bit types, two-pass
// reductions when one pass would leave most SMs idle.
// (capability parity: hetu/impl/kernel/{Arithmetics,Reduce,Softmax,Concat,Slice,DataTransfer,...}.cu and
//  hetu/impl/utils/{offset_calculator,cuda_math}.h -- own design, one file)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace hb {
namespace {

constexpr int kMaxDims = 8;

// division by a runtime constant through a 64-bit multiply (valid for n < 2^31, d < 2^31)
struct FastDiv {
  uint32_t d = 1, shift = 0;
  uint64_t magic = 1;
  __host__ void init(uint32_t div) {
    d = div;
    if (div == 1) { magic = 1; shift = 0; return; }
    uint32_t l = 0;
    while ((1ull << l) < div) ++l;
    shift = 31 + l;
    magic = ((1ull << shift) + div - 1) / div;
  }
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1 ? n : (uint32_t)(((uint64_t)n * magic) >> shift); }   // magic <= 2^32, n < 2^31

};

// element index -> offsets into up to NARG strided operands (innermost dimension first)
template <int NARG>
struct OffsetCalc {
  int ndim;
  FastDiv sizes[kMaxDims];
  int64_t strides[NARG][kMaxDims];
  __device__ __forceinline__ void get(uint32_t linear, int64_t* off) const {
#pragma unroll
    for (int a = 0; a < NARG; ++a) off[a] = 0;
#pragma unroll
    for (int d = 0; d < kMaxDims; ++d) {
      if (d == ndim) break;
      const uint32_t q = sizes[d].div(linear);
      const uint32_t r = linear - q * sizes[d].d;
      linear = q;
#pragma unroll
      for (int a = 0; a < NARG; ++a) off[a] += (int64_t)r * strides[a][d];
    }
  }
};

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<int64_t>(int64_t v) { return (float)v; }
template <> __device__ __forceinline__ float to_f<int32_t>(int32_t v) { return (float)v; }
template <> __device__ __forceinline__ float to_f<uint8_t>(uint8_t v) { return (float)v; }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ int64_t from_f<int64_t>(float v) { return (int64_t)v; }
template <> __device__ __forceinline__ int32_t from_f<int32_t>(float v) { return (int32_t)v; }
template <> __device__ __forceinline__ uint8_t from_f<uint8_t>(float v) { return (uint8_t)v; }

inline int grid_for(int64_t work, int block, int per_sm = 8) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int64_t blocks = (work + block - 1) / block;
  return (int)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)sms * per_sm));
}

// ------------------------------------------------------------------ strided copy
template <typename V>
__global__ void strided_copy_kernel(const V* __restrict__ src, V* __restrict__ dst, OffsetCalc<2> oc, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int64_t off[2];
    oc.get(i, off);
    dst[off[1]] = src[off[0]];
  }
}

// ------------------------------------------------------------------ unary family
template <int OP>
__device__ __forceinline__ float unary_fn(float x, float p0, float p1) {
  if constexpr (OP == G_NEG) return -x;
  else if constexpr (OP == G_RECIPROCAL) return 1.0f / x;
  else if constexpr (OP == G_ABS) return fabsf(x);
  else if constexpr (OP == G_CEIL) return ceilf(x);
  else if constexpr (OP == G_FLOOR) return floorf(x);
  else if constexpr (OP == G_ROUND) return nearbyintf(x);
  else if constexpr (OP == G_EXP) return expf(x);
  else if constexpr (OP == G_LOG) return logf(x);
  else if constexpr (OP == G_SQRT) return sqrtf(x);
  else if constexpr (OP == G_RSQRT) return 1.0f / sqrtf(x);
  else if constexpr (OP == G_SIN) return sinf(x);
  else if constexpr (OP == G_COS) return cosf(x);
  else if constexpr (OP == G_CLAMP) return fminf(fmaxf(x, p0), p1);
  else if constexpr (OP == G_SIGMOID) return 1.0f / (1.0f + expf(-x));
  else if constexpr (OP == G_TANH) return tanhf(x);
  else if constexpr (OP == G_LEAKYRELU) return x > 0.f ? x : x * p0;
  else if constexpr (OP == G_ELU) return x > 0.f ? x * p1 : (expf(x) - 1.0f) * p0 * p1;
  else if constexpr (OP == G_HARDSHRINK) return (x >= -p0 && x <= p0) ? 0.f : x;
  else if constexpr (OP == G_HARDSIGMOID) return fminf(fmaxf(x + 3.0f, 0.f), 6.0f) / 6.0f;
  else if constexpr (OP == G_HARDTANH) return fminf(fmaxf(x, p0), p1);
  else if constexpr (OP == G_HARDSWISH) return x * fminf(fmaxf(x + 3.0f, 0.f), 6.0f) / 6.0f;
  else if constexpr (OP == G_LOGSIGMOID) return fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
  else if constexpr (OP == G_SOFTPLUS) return x * p0 > p1 ? x : log1pf(expf(x * p0)) / p0;
  else if constexpr (OP == G_MISH) return x * tanhf(x > 20.f ? x : log1pf(expf(x)));
  else if constexpr (OP == G_SOFTSHRINK) return x > p0 ? x - p0 : (x < -p0 ? x + p0 : 0.f);
  else if constexpr (OP == G_POW) return powf(x, p0);
  else if constexpr (OP == G_ADD_SCALAR) return x + p0;
  else if constexpr (OP == G_MUL_SCALAR) return x * p0;
  else if constexpr (OP == G_RSUB_SCALAR) return p0 - x;
  else if constexpr (OP == G_RDIV_SCALAR) return p0 / x;
  else if constexpr (OP == G_DIV_SCALAR) return x / p0;
  else return x;
}

template <int OP, typename T>
__global__ void unary_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n, float p0, float p1) {
  constexpr int V = 16 / sizeof(T);
  const int64_t nv = n / V;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint4* x4 = reinterpret_cast<const uint4*>(x);
  uint4* y4 = reinterpret_cast<uint4*>(y);
  for (int64_t i = tid; i < nv; i += stride) {
    uint4 in = x4[i], out;
    const T* a = reinterpret_cast<const T*>(&in);
    T* b = reinterpret_cast<T*>(&out);
#pragma unroll
    for (int j = 0; j < V; ++j) b[j] = from_f<T>(unary_fn<OP>(to_f<T>(a[j]), p0, p1));
    y4[i] = out;
  }
  for (int64_t i = nv * V + tid; i < n; i += stride) y[i] = from_f<T>(unary_fn<OP>(to_f<T>(x[i]), p0, p1));
}

// ------------------------------------------------------------------ broadcasting binary
template <int OP>
__device__ __forceinline__ float binary_fn(float a, float b) {
  if constexpr (OP == B_ADD) return a + b;
  else if constexpr (OP == B_SUB) return a - b;
  else if constexpr (OP == B_MUL) return a * b;
  else if constexpr (OP == B_DIV) return a / b;
  else if constexpr (OP == B_MAX) return fmaxf(a, b);
  else if constexpr (OP == B_MIN) return fminf(a, b);
  else return powf(a, b);
}
template <int OP, typename T>
__global__ void binary_bcast_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, OffsetCalc<2> oc, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int64_t off[2];
    oc.get(i, off);
    out[i] = from_f<T>(binary_fn<OP>(to_f<T>(a[off[0]]), to_f<T>(b[off[1]])));
  }
}
// same-shape contiguous operands: 16-byte path
template <int OP, typename T>
__global__ void binary_flat_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t n) {
  constexpr int V = 16 / sizeof(T);
  const int64_t nv = n / V;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = tid; i < nv; i += stride) {
    uint4 va = reinterpret_cast<const uint4*>(a)[i], vb = reinterpret_cast<const uint4*>(b)[i], vo;
    const T* pa = reinterpret_cast<const T*>(&va);
    const T* pb = reinterpret_cast<const T*>(&vb);
    T* po = reinterpret_cast<T*>(&vo);
#pragma unroll
    for (int j = 0; j < V; ++j) po[j] = from_f<T>(binary_fn<OP>(to_f<T>(pa[j]), to_f<T>(pb[j])));
    reinterpret_cast<uint4*>(out)[i] = vo;
  }
  for (int64_t i = nv * V + tid; i < n; i += stride) out[i] = from_f<T>(binary_fn<OP>(to_f<T>(a[i]), to_f<T>(b[i])));
}

// ------------------------------------------------------------------ reductions
template <int MODE> __device__ __forceinline__ float red_init() {
  if constexpr (MODE == R_MAX) return -INFINITY;
  else if constexpr (MODE == R_MIN) return INFINITY;
  else if constexpr (MODE == R_PROD) return 1.0f;
  else return 0.0f;
}
template <int MODE> __device__ __forceinline__ float red_op(float a, float b) {
  if constexpr (MODE == R_MAX) return (a != a || b != b) ? NAN : fmaxf(a, b);      // NaN propagates like the library reduction
  else if constexpr (MODE == R_MIN) return (a != a || b != b) ? NAN : fminf(a, b);
  else if constexpr (MODE == R_PROD) return a * b;
  else return a + b;
}
template <int MODE> __device__ __forceinline__ float block_reduce(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = red_op<MODE>(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (l == 0) sh[w] = v;
  __syncthreads();
  v = l < nw ? sh[l] : red_init<MODE>();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = red_op<MODE>(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  return v;
}
// x viewed as [rows, len] contiguous; block (row, chunk) reduces x[row, chunk * chunk_len : ...) -> out[row * chunks + chunk]
template <int MODE, typename TI, typename TO>
__global__ void reduce_rows_kernel(const TI* __restrict__ x, TO* __restrict__ out, int64_t len, int64_t chunk_len, int chunks, float scale) {
  __shared__ float sh[32];
  const int64_t row = blockIdx.x;
  const int chunk = blockIdx.y;
  const int64_t lo = (int64_t)chunk * chunk_len, hi = min(len, lo + chunk_len);
  const TI* p = x + row * len;
  float acc = red_init<MODE>();
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) acc = red_op<MODE>(acc, to_f<TI>(p[i]));
  acc = block_reduce<MODE>(acc, sh);
  if (threadIdx.x == 0) out[row * chunks + chunk] = from_f<TO>(acc * scale);
}
// x viewed as [outer, red, inner]; thread (o, i) of chunk c reduces x[o, c * chunk_len : ..., i] -> out[(o * chunks + c) * inner + i]
template <int MODE, typename TI, typename TO>
__global__ void reduce_cols_kernel(const TI* __restrict__ x, TO* __restrict__ out, int64_t red, int64_t inner, int64_t chunk_len, int chunks,
                                   float scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= inner) return;
  const int64_t o = blockIdx.y;
  const int c = blockIdx.z;
  const int64_t lo = (int64_t)c * chunk_len, hi = min(red, lo + chunk_len);
  const TI* p = x + (o * red + lo) * inner + i;
  float acc = red_init<MODE>();
  for (int64_t r = lo; r < hi; ++r, p += inner) acc = red_op<MODE>(acc, to_f<TI>(*p));
  out[(o * chunks + c) * inner + i] = from_f<TO>(acc * scale);
}

// ------------------------------------------------------------------ softmax
// rows of `dim` contiguous elements; one block per row, three passes (the row stays in L1 / L2 between them)
template <bool LOG, typename T>
__global__ void softmax_rows_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t dim) {
  __shared__ float sh[32];
  const T* p = x + (int64_t)blockIdx.x * dim;
  T* q = y + (int64_t)blockIdx.x * dim;
  float m = -INFINITY;
  for (int64_t i = threadIdx.x; i < dim; i += blockDim.x) m = fmaxf(m, to_f<T>(p[i]));
  m = block_reduce<R_MAX>(m, sh);
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < dim; i += blockDim.x) s += expf(to_f<T>(p[i]) - m);
  s = block_reduce<R_SUM>(s, sh);
  const float inv = 1.0f / s, ls = logf(s);
  for (int64_t i = threadIdx.x; i < dim; i += blockDim.x) {
    const float v = to_f<T>(p[i]) - m;
    q[i] = from_f<T>(LOG ? v - ls : expf(v) * inv);
  }
}
// [outer, dim, inner] with inner > 1: thread per (outer, inner) column, coalesced across inner
template <bool LOG, typename T>
__global__ void softmax_cols_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t dim, int64_t inner) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= inner) return;
  const T* p = x + (int64_t)blockIdx.y * dim * inner + i;
  T* q = y + (int64_t)blockIdx.y * dim * inner + i;
  float m = -INFINITY;
  for (int64_t d = 0; d < dim; ++d) m = fmaxf(m, to_f<T>(p[d * inner]));
  float s = 0.f;
  for (int64_t d = 0; d < dim; ++d) s += expf(to_f<T>(p[d * inner]) - m);
  const float inv = 1.0f / s, ls = logf(s);
  for (int64_t d = 0; d < dim; ++d) {
    const float v = to_f<T>(p[d * inner]) - m;
    q[d * inner] = from_f<T>(LOG ? v - ls : expf(v) * inv);
  }
}

// ------------------------------------------------------------------ cast / fill
template <typename S, typename D> __device__ __forceinline__ D cast_one(S v) { return from_f<D>(to_f<S>(v)); }
template <> __device__ __forceinline__ int64_t cast_one<int32_t, int64_t>(int32_t v) { return (int64_t)v; }
template <> __device__ __forceinline__ int32_t cast_one<int64_t, int32_t>(int64_t v) { return (int32_t)v; }
template <> __device__ __forceinline__ int64_t cast_one<uint8_t, int64_t>(uint8_t v) { return (int64_t)v; }
template <> __device__ __forceinline__ int32_t cast_one<uint8_t, int32_t>(uint8_t v) { return (int32_t)v; }
template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ src, D* __restrict__ dst, int64_t n) {
  constexpr int V = 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * V;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V; i < n; i += stride) {
    if (i + V <= n) {
      S in[V];
#pragma unroll
      for (int j = 0; j < V; ++j) in[j] = src[i + j];
#pragma unroll
      for (int j = 0; j < V; ++j) dst[i + j] = cast_one<S, D>(in[j]);
    } else {
      for (int64_t j = i; j < n; ++j) dst[j] = cast_one<S, D>(src[j]);
    }
  }
}
template <typename T>
__global__ void fill_kernel(T* __restrict__ dst, int64_t n, T v) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = v;
}

// collapse adjacent dimensions that are contiguous in every operand; returns the new rank (innermost first)
template <int NARG>
int collapse_dims(int ndim, const int64_t* shape, const int64_t* const* strides, int64_t* oshape, int64_t (*ostrides)[kMaxDims]) {
  // inputs are outermost-first (torch order); build innermost-first while dropping size-1 dims
  int n = 0;
  for (int d = ndim - 1; d >= 0; --d) {
    if (shape[d] == 1) continue;
    bool merge = n > 0;
    if (merge)
      for (int a = 0; a < NARG; ++a) merge = merge && strides[a][d] == ostrides[a][n - 1] * oshape[n - 1];
    if (merge) oshape[n - 1] *= shape[d];
    else {
      if (n == kMaxDims) return -1;
      oshape[n] = shape[d];
      for (int a = 0; a < NARG; ++a) ostrides[a][n] = strides[a][d];
      ++n;
    }
  }
  if (n == 0) {
    oshape[0] = 1;
    for (int a = 0; a < NARG; ++a) ostrides[a][0] = 1;
    n = 1;
  }
  return n;
}

template <typename V>
cudaError_t launch_strided_copy(const void* src, void* dst, int n, const int64_t* shape, int64_t (*st)[kMaxDims], int64_t total, cudaStream_t s) {
  OffsetCalc<2> oc;
  oc.ndim = n;
  for (int d = 0; d < n; ++d) {
    oc.sizes[d].init((uint32_t)shape[d]);
    oc.strides[0][d] = st[0][d];
    oc.strides[1][d] = st[1][d];
  }
  strided_copy_kernel<V><<<grid_for(total, 256), 256, 0, s>>>(reinterpret_cast<const V*>(src), reinterpret_cast<V*>(dst), oc, (uint32_t)total);
  count_launch();
  return cudaGetLastError();
}

}  // namespace

// ------------------------------------------------------------------ host API
cudaError_t strided_copy(int elem_bytes, const void* src, void* dst, int ndim, const int64_t* shape, const int64_t* src_strides,
                         const int64_t* dst_strides, cudaStream_t s) {
  if (ndim > kMaxDims) return cudaErrorInvalidValue;
  int64_t total = 1;
  for (int d = 0; d < ndim; ++d) total *= shape[d];
  if (total == 0) return cudaSuccess;
  const int64_t* strides[2] = {src_strides, dst_strides};
  int64_t oshape[kMaxDims], ost[2][kMaxDims];
  int n = collapse_dims<2>(ndim, shape, strides, oshape, ost);
  if (n < 0) return cudaErrorInvalidValue;
  // widen the element to 16 bytes when the innermost run is contiguous in both operands and everything is 16-byte aligned
  int64_t unit = elem_bytes;
  if (ost[0][0] == 1 && ost[1][0] == 1) {
    for (int64_t w : {16, 8, 4, 2}) {
      if (w <= elem_bytes || w % elem_bytes != 0) continue;
      const int64_t k = w / elem_bytes;
      bool ok = oshape[0] % k == 0 && (reinterpret_cast<uintptr_t>(src) % w) == 0 && (reinterpret_cast<uintptr_t>(dst) % w) == 0;
      for (int d = 1; d < n && ok; ++d) ok = ost[0][d] % k == 0 && ost[1][d] % k == 0;
      if (ok) {
        oshape[0] /= k;
        for (int d = 1; d < n; ++d) { ost[0][d] /= k; ost[1][d] /= k; }
        total /= k;
        unit = w;
        break;
      }
    }
  }
  if (total >= (1ll << 31)) return cudaErrorInvalidValue;
  switch (unit) {
    case 16: return launch_strided_copy<uint4>(src, dst, n, oshape, ost, total, s);
    case 8: return launch_strided_copy<uint64_t>(src, dst, n, oshape, ost, total, s);
    case 4: return launch_strided_copy<uint32_t>(src, dst, n, oshape, ost, total, s);
    case 2: return launch_strided_copy<uint16_t>(src, dst, n, oshape, ost, total, s);
    case 1: return launch_strided_copy<uint8_t>(src, dst, n, oshape, ost, total, s);
    default: return cudaErrorInvalidValue;
  }
}

namespace {
template <int OP>
cudaError_t unary_dispatch_dtype(int dtype, const void* x, void* y, int64_t n, float p0, float p1, cudaStream_t s) {
  const bool vec_ok = (reinterpret_cast<uintptr_t>(x) % 16) == 0 && (reinterpret_cast<uintptr_t>(y) % 16) == 0;
  const int64_t nn = n;
  const int grid = grid_for(n / 4 + 1, 256);
  if (!vec_ok) return cudaErrorMisalignedAddress;
  switch (dtype) {
    case GD_F32: unary_kernel<OP, float><<<grid, 256, 0, s>>>((const float*)x, (float*)y, nn, p0, p1); break;
    case GD_BF16: unary_kernel<OP, __nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, nn, p0, p1); break;
    case GD_F16: unary_kernel<OP, __half><<<grid, 256, 0, s>>>((const __half*)x, (__half*)y, nn, p0, p1); break;
    default: return cudaErrorInvalidValue;
  }
  count_launch();
  return cudaGetLastError();
}
}  // namespace

cudaError_t generic_unary(int op, int dtype, const void* x, void* y, int64_t n, float p0, float p1, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
#define HB_U(OP) case OP: return unary_dispatch_dtype<OP>(dtype, x, y, n, p0, p1, s);
  switch (op) {
    HB_U(G_NEG) HB_U(G_RECIPROCAL) HB_U(G_ABS) HB_U(G_CEIL) HB_U(G_FLOOR) HB_U(G_ROUND) HB_U(G_EXP) HB_U(G_LOG) HB_U(G_SQRT)
    HB_U(G_RSQRT) HB_U(G_SIN) HB_U(G_COS) HB_U(G_CLAMP) HB_U(G_SIGMOID) HB_U(G_TANH) HB_U(G_LEAKYRELU) HB_U(G_ELU) HB_U(G_HARDSHRINK)
    HB_U(G_HARDSIGMOID) HB_U(G_HARDTANH) HB_U(G_HARDSWISH) HB_U(G_LOGSIGMOID) HB_U(G_SOFTPLUS) HB_U(G_MISH) HB_U(G_SOFTSHRINK)
    HB_U(G_POW) HB_U(G_ADD_SCALAR) HB_U(G_MUL_SCALAR) HB_U(G_RSUB_SCALAR) HB_U(G_RDIV_SCALAR) HB_U(G_DIV_SCALAR)
    default: return cudaErrorInvalidValue;
  }
#undef HB_U
}

namespace {
template <int OP, typename T>
cudaError_t binary_launch(const void* a, const void* b, void* out, int ndim, const int64_t* shape, const int64_t* as, const int64_t* bs,
                          cudaStream_t s) {
  int64_t total = 1;
  for (int d = 0; d < ndim; ++d) total *= shape[d];
  if (total == 0) return cudaSuccess;
  const int64_t* strides[2] = {as, bs};
  int64_t oshape[kMaxDims], ost[2][kMaxDims];
  const int n = collapse_dims<2>(ndim, shape, strides, oshape, ost);
  if (n < 0) return cudaErrorInvalidValue;
  const bool flat = n == 1 && ost[0][0] == 1 && ost[1][0] == 1;
  const bool aligned = (reinterpret_cast<uintptr_t>(a) % 16) == 0 && (reinterpret_cast<uintptr_t>(b) % 16) == 0 &&
                       (reinterpret_cast<uintptr_t>(out) % 16) == 0;
  if (flat && aligned) {
    binary_flat_kernel<OP, T><<<grid_for(total / 4 + 1, 256), 256, 0, s>>>((const T*)a, (const T*)b, (T*)out, total);
  } else {
    if (total >= (1ll << 31)) return cudaErrorInvalidValue;
    OffsetCalc<2> oc;
    oc.ndim = n;
    for (int d = 0; d < n; ++d) {
      oc.sizes[d].init((uint32_t)oshape[d]);
      oc.strides[0][d] = ost[0][d];
      oc.strides[1][d] = ost[1][d];
    }
    binary_bcast_kernel<OP, T><<<grid_for(total, 256), 256, 0, s>>>((const T*)a, (const T*)b, (T*)out, oc, (uint32_t)total);
  }
  count_launch();
  return cudaGetLastError();
}
template <int OP>
cudaError_t binary_dtype(int dtype, const void* a, const void* b, void* out, int ndim, const int64_t* shape, const int64_t* as,
                         const int64_t* bs, cudaStream_t s) {
  switch (dtype) {
    case GD_F32: return binary_launch<OP, float>(a, b, out, ndim, shape, as, bs, s);
    case GD_BF16: return binary_launch<OP, __nv_bfloat16>(a, b, out, ndim, shape, as, bs, s);
    case GD_F16: return binary_launch<OP, __half>(a, b, out, ndim, shape, as, bs, s);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace

cudaError_t generic_binary(int op, int dtype, const void* a, const void* b, void* out, int ndim, const int64_t* out_shape,
                           const int64_t* a_strides, const int64_t* b_strides, cudaStream_t s) {
  if (ndim > kMaxDims) return cudaErrorInvalidValue;
  switch (op) {
    case B_ADD: return binary_dtype<B_ADD>(dtype, a, b, out, ndim, out_shape, a_strides, b_strides, s);
    case B_SUB: return binary_dtype<B_SUB>(dtype, a, b, out, ndim, out_shape, a_strides, b_strides, s);
    case B_MUL: return binary_dtype<B_MUL>(dtype, a, b, out, ndim, out_shape, a_strides, b_strides, s);
    case B_DIV: return binary_dtype<B_DIV>(dtype, a, b, out, ndim, out_shape, a_strides, b_strides, s);
    case B_MAX: return binary_dtype<B_MAX>(dtype, a, b, out, ndim, out_shape, a_strides, b_strides, s);
    case B_MIN: return binary_dtype<B_MIN>(dtype, a, b, out, ndim, out_shape, a_strides, b_strides, s);
    case B_POW: return binary_dtype<B_POW>(dtype, a, b, out, ndim, out_shape, a_strides, b_strides, s);
    default: return cudaErrorInvalidValue;
  }
}

namespace {
template <int MODE, typename T>
cudaError_t reduce_launch(const T* x, T* y, float* workspace, int64_t outer, int64_t red, int64_t inner, float scale, cudaStream_t s) {
  // one pass when it fills the machine; otherwise split the reduced extent into chunks -> fp32 partials -> second pass
  int chunks = generic_reduce_chunks(outer, red, inner);
  if (chunks > 1 && workspace == nullptr) chunks = 1;
  const int64_t chunk_len = (red + chunks - 1) / chunks;
  if (inner == 1) {
    if (outer > 2147483647ll) return cudaErrorInvalidValue;
    const int block = red >= 1024 ? 256 : (red >= 128 ? 128 : 32);
    if (chunks == 1) reduce_rows_kernel<MODE, T, T><<<dim3((unsigned)outer, 1), block, 0, s>>>(x, y, red, chunk_len, 1, scale);
    else reduce_rows_kernel<MODE, T, float><<<dim3((unsigned)outer, chunks), block, 0, s>>>(x, workspace, red, chunk_len, chunks, 1.0f);
  } else {
    if (outer > 65535) return cudaErrorInvalidValue;
    const dim3 grid((unsigned)((inner + 255) / 256), (unsigned)outer, chunks);
    if (chunks == 1) reduce_cols_kernel<MODE, T, T><<<grid, 256, 0, s>>>(x, y, red, inner, chunk_len, 1, scale);
    else reduce_cols_kernel<MODE, T, float><<<grid, 256, 0, s>>>(x, workspace, red, inner, chunk_len, chunks, 1.0f);
  }
  count_launch();
  if (chunks > 1) {
    if (outer > 65535) return cudaErrorInvalidValue;
    const dim3 grid((unsigned)((inner + 255) / 256), (unsigned)outer, 1);
    reduce_cols_kernel<MODE, float, T><<<grid, 256, 0, s>>>(workspace, y, chunks, inner, chunks, 1, scale);
    count_launch();
  }
  return cudaGetLastError();
}
template <int MODE>
cudaError_t reduce_dtype(int dtype, const void* x, void* y, float* ws, int64_t outer, int64_t red, int64_t inner, float scale, cudaStream_t s) {
  switch (dtype) {
    case GD_F32: return reduce_launch<MODE, float>((const float*)x, (float*)y, ws, outer, red, inner, scale, s);
    case GD_BF16: return reduce_launch<MODE, __nv_bfloat16>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, ws, outer, red, inner, scale, s);
    case GD_F16: return reduce_launch<MODE, __half>((const __half*)x, (__half*)y, ws, outer, red, inner, scale, s);
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace

// number of slices the reduced extent is cut into: enough (row, slice) blocks / (column, slice) threads to cover ~2 waves of the
// machine, at least 1024 elements per slice, at most 1024 slices
int generic_reduce_chunks(int64_t outer, int64_t red, int64_t inner) {
  const int64_t parallel = inner == 1 ? outer * 256 : outer * inner;      // threads one pass would run
  if (parallel >= 148 * 1024 || red < 4096) return 1;
  const int64_t want = (148 * 2048 + parallel - 1) / std::max<int64_t>(parallel, 1);
  const int64_t chunks = std::min<int64_t>(std::min<int64_t>(1024, red / 1024), want);
  return chunks < 2 ? 1 : (int)chunks;
}
int64_t generic_reduce_workspace_floats(int64_t outer, int64_t red, int64_t inner) {
  const int chunks = generic_reduce_chunks(outer, red, inner);
  return chunks > 1 ? (int64_t)chunks * outer * inner : 0;
}

cudaError_t generic_reduce(int mode, int dtype, const void* x, void* y, float* workspace, int64_t outer, int64_t red, int64_t inner,
                           cudaStream_t s) {
  if (outer * inner == 0) return cudaSuccess;
  if (red == 0) return cudaErrorInvalidValue;
  switch (mode) {
    case R_SUM: return reduce_dtype<R_SUM>(dtype, x, y, workspace, outer, red, inner, 1.0f, s);
    case R_MEAN: return reduce_dtype<R_SUM>(dtype, x, y, workspace, outer, red, inner, 1.0f / (float)red, s);
    case R_MAX: return reduce_dtype<R_MAX>(dtype, x, y, workspace, outer, red, inner, 1.0f, s);
    case R_MIN: return reduce_dtype<R_MIN>(dtype, x, y, workspace, outer, red, inner, 1.0f, s);
    case R_PROD: return reduce_dtype<R_PROD>(dtype, x, y, workspace, outer, red, inner, 1.0f, s);
    default: return cudaErrorInvalidValue;
  }
}

namespace {
template <bool LOG, typename T>
cudaError_t softmax_launch(const T* x, T* y, int64_t outer, int64_t dim, int64_t inner, cudaStream_t s) {
  if (inner == 1) {
    if (outer > 2147483647ll) return cudaErrorInvalidValue;
    const int block = dim >= 2048 ? 512 : (dim >= 512 ? 256 : (dim >= 128 ? 128 : 32));
    softmax_rows_kernel<LOG, T><<<(unsigned)outer, block, 0, s>>>(x, y, dim);
  } else {
    if (outer > 65535) return cudaErrorInvalidValue;
    softmax_cols_kernel<LOG, T><<<dim3((unsigned)((inner + 127) / 128), (unsigned)outer), 128, 0, s>>>(x, y, dim, inner);
  }
  count_launch();
  return cudaGetLastError();
}
}  // namespace

cudaError_t generic_softmax(bool log, int dtype, const void* x, void* y, int64_t outer, int64_t dim, int64_t inner, cudaStream_t s) {
  if (outer * dim * inner == 0) return cudaSuccess;
#define HB_SM(T) (log ? softmax_launch<true, T>((const T*)x, (T*)y, outer, dim, inner, s) : softmax_launch<false, T>((const T*)x, (T*)y, outer, dim, inner, s))
  switch (dtype) {
    case GD_F32: return HB_SM(float);
    case GD_BF16: return HB_SM(__nv_bfloat16);
    case GD_F16: return HB_SM(__half);
    default: return cudaErrorInvalidValue;
  }
#undef HB_SM
}

namespace {
template <typename S>
cudaError_t cast_from(int dst_dtype, const S* src, void* dst, int64_t n, cudaStream_t s) {
  const int grid = grid_for(n / 4 + 1, 256);
  switch (dst_dtype) {
    case GD_F32: cast_kernel<S, float><<<grid, 256, 0, s>>>(src, (float*)dst, n); break;
    case GD_BF16: cast_kernel<S, __nv_bfloat16><<<grid, 256, 0, s>>>(src, (__nv_bfloat16*)dst, n); break;
    case GD_F16: cast_kernel<S, __half><<<grid, 256, 0, s>>>(src, (__half*)dst, n); break;
    case GD_I64: cast_kernel<S, int64_t><<<grid, 256, 0, s>>>(src, (int64_t*)dst, n); break;
    case GD_I32: cast_kernel<S, int32_t><<<grid, 256, 0, s>>>(src, (int32_t*)dst, n); break;
    default: return cudaErrorInvalidValue;
  }
  count_launch();
  return cudaGetLastError();
}
}  // namespace

cudaError_t generic_cast(int src_dtype, int dst_dtype, const void* src, void* dst, int64_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  switch (src_dtype) {
    case GD_F32: return cast_from<float>(dst_dtype, (const float*)src, dst, n, s);
    case GD_BF16: return cast_from<__nv_bfloat16>(dst_dtype, (const __nv_bfloat16*)src, dst, n, s);
    case GD_F16: return cast_from<__half>(dst_dtype, (const __half*)src, dst, n, s);
    case GD_I64: return cast_from<int64_t>(dst_dtype, (const int64_t*)src, dst, n, s);
    case GD_I32: return cast_from<int32_t>(dst_dtype, (const int32_t*)src, dst, n, s);
    case GD_U8: return cast_from<uint8_t>(dst_dtype, (const uint8_t*)src, dst, n, s);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t generic_fill(int dtype, void* dst, int64_t n, double value, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  const int grid = grid_for(n, 256);
  switch (dtype) {
    case GD_F32: fill_kernel<float><<<grid, 256, 0, s>>>((float*)dst, n, (float)value); break;
    case GD_BF16: fill_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((__nv_bfloat16*)dst, n, __float2bfloat16_rn((float)value)); break;
    case GD_F16: fill_kernel<__half><<<grid, 256, 0, s>>>((__half*)dst, n, __float2half_rn((float)value)); break;
    case GD_I64: fill_kernel<int64_t><<<grid, 256, 0, s>>>((int64_t*)dst, n, (int64_t)value); break;
    case GD_I32: fill_kernel<int32_t><<<grid, 256, 0, s>>>((int32_t*)dst, n, (int32_t)value); break;
    case GD_U8: fill_kernel<uint8_t><<<grid, 256, 0, s>>>((uint8_t*)dst, n, (uint8_t)value); break;
    default: return cudaErrorInvalidValue;
  }
  count_launch();
  return cudaGetLastError();
}

}  // namespace hb
