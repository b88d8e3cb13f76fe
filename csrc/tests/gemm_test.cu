// Standalone correctness + timing harness for the tcgen05 GEMM.
//   gemm_test G amn bmn M N K [epi] [out_fp32] [time_iters]
// Verifies sampled outputs against an fp32 reference computed on the GPU with a
// trivially-correct kernel, then (optionally) times it against cuBLAS.
#include <cublas_v2.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../kernels/gemm_sm100.h"

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

__global__ void fill_kernel(__nv_bfloat16* p, size_t n, uint32_t seed, float scale) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x = (uint32_t)i * 2654435761u + seed;
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  float f = ((x & 0xFFFF) / 65536.0f - 0.5f) * 2.0f * scale;
  p[i] = __float2bfloat16(f);
}

__device__ float gelu_ref(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678f)); }

// sampled reference: out[s] = sum_k A(i,k) * B(k,j) for sample s = (i, j)
__global__ void ref_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B, int M, int N, int K, int64_t lda, int64_t ldb,
                           int amn, int bmn, const int* si, const int* sj, int ns, float* out) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ns) return;
  int i = si[s], j = sj[s];
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    float a = __bfloat162float(amn ? A[(int64_t)k * lda + i] : A[(int64_t)i * lda + k]);
    float b = __bfloat162float(bmn ? B[(int64_t)k * ldb + j] : B[(int64_t)j * ldb + k]);
    acc += a * b;
  }
  out[s] = acc;
}

int main(int argc, char** argv) {
  if (argc < 7) { printf("usage: gemm_test G amn bmn M N K [epi] [out_fp32] [time_iters]\n"); return 1; }
  int G = atoi(argv[1]), amn = atoi(argv[2]), bmn = atoi(argv[3]);
  int M = atoi(argv[4]), N = atoi(argv[5]), K = atoi(argv[6]);
  int epi = argc > 7 ? atoi(argv[7]) : 0;  // 0 none, 1 bias+gelu+aux_out, 2 bias+residual, 3 accumulate
  int out_fp32 = argc > 8 ? atoi(argv[8]) : 0;
  int iters = argc > 9 ? atoi(argv[9]) : 0;

  int64_t lda = amn ? M : K, ldb = bmn ? N : K, ldc = N;
  size_t na = (size_t)M * K, nb = (size_t)N * K, nc = (size_t)M * N;
  __nv_bfloat16 *A, *B, *bias, *aux, *auxo;
  void* C;
  CK(cudaMalloc(&A, na * 2)); CK(cudaMalloc(&B, nb * 2)); CK(cudaMalloc(&C, nc * 4));
  CK(cudaMalloc(&bias, N * 2)); CK(cudaMalloc(&aux, nc * 2)); CK(cudaMalloc(&auxo, nc * 2));
  fill_kernel<<<(na + 255) / 256, 256>>>(A, na, 1u, 1.0f);
  fill_kernel<<<(nb + 255) / 256, 256>>>(B, nb, 7u, 1.0f);
  fill_kernel<<<(N + 255) / 256, 256>>>(bias, N, 13u, 1.0f);
  fill_kernel<<<(nc + 255) / 256, 256>>>(aux, nc, 17u, 1.0f);
  if (epi == 3) {
    if (out_fp32) CK(cudaMemset(C, 0, nc * 4));
    else fill_kernel<<<(nc + 255) / 256, 256>>>((__nv_bfloat16*)C, nc, 23u, 1.0f);
  } else CK(cudaMemset(C, 0xFF, nc * 4));
  CK(cudaDeviceSynchronize());

  std::vector<__nv_bfloat16> c_before;
  if (epi == 3 && !out_fp32) { c_before.resize(nc); CK(cudaMemcpy(c_before.data(), C, nc * 2, cudaMemcpyDeviceToHost)); }

  hb::GemmCall call;
  call.A = A; call.B = B; call.C = C; call.M = M; call.N = N; call.K = K;
  call.lda = lda; call.ldb = ldb; call.ldc = ldc; call.a_mn_major = amn; call.b_mn_major = bmn;
  call.out = out_fp32 ? hb::GemmOut::FP32 : hb::GemmOut::BF16;
  call.cta_group = G;
  if (epi == 1) { call.bias = bias; call.act = 1; call.aux_out = auxo; call.ld_aux = N; }
  if (epi == 2) { call.bias = bias; call.aux_in = aux; call.aux_mode = 1; call.ld_aux = N; }
  if (epi == 3) { call.accumulate = true; }
  float alpha = 1.0f / sqrtf((float)K);
  call.alpha = alpha;

  CK(hb::gemm_bf16(call, 0));
  CK(cudaDeviceSynchronize());

  // sampled verification
  const int ns = 8192;
  std::vector<int> hi(ns), hj(ns);
  srand(123);
  for (int s = 0; s < ns; ++s) { hi[s] = rand() % M; hj[s] = rand() % N; }
  // make sure corners / edges are covered
  hi[0] = 0; hj[0] = 0; hi[1] = M - 1; hj[1] = N - 1; hi[2] = M - 1; hj[2] = 0; hi[3] = 0; hj[3] = N - 1;
  int *di, *dj; float* dref;
  CK(cudaMalloc(&di, ns * 4)); CK(cudaMalloc(&dj, ns * 4)); CK(cudaMalloc(&dref, ns * 4));
  CK(cudaMemcpy(di, hi.data(), ns * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dj, hj.data(), ns * 4, cudaMemcpyHostToDevice));
  ref_kernel<<<(ns + 127) / 128, 128>>>(A, B, M, N, K, lda, ldb, amn, bmn, di, dj, ns, dref);
  CK(cudaDeviceSynchronize());
  std::vector<float> href(ns);
  CK(cudaMemcpy(href.data(), dref, ns * 4, cudaMemcpyDeviceToHost));
  std::vector<__nv_bfloat16> hc16, hbias(N), haux, hauxo;
  std::vector<float> hc32;
  if (out_fp32) { hc32.resize(nc); CK(cudaMemcpy(hc32.data(), C, nc * 4, cudaMemcpyDeviceToHost)); }
  else { hc16.resize(nc); CK(cudaMemcpy(hc16.data(), C, nc * 2, cudaMemcpyDeviceToHost)); }
  CK(cudaMemcpy(hbias.data(), bias, N * 2, cudaMemcpyDeviceToHost));
  if (epi == 2) { haux.resize(nc); CK(cudaMemcpy(haux.data(), aux, nc * 2, cudaMemcpyDeviceToHost)); }
  if (epi == 1) { hauxo.resize(nc); CK(cudaMemcpy(hauxo.data(), auxo, nc * 2, cudaMemcpyDeviceToHost)); }
  double max_err = 0, max_ref = 0; int bad = 0;
  for (int s = 0; s < ns; ++s) {
    size_t idx = (size_t)hi[s] * N + hj[s];
    float want = href[s] * alpha;
    if (epi == 1) {
      want += __bfloat162float(hbias[hj[s]]);
      float pre = __bfloat162float(hauxo[idx]);
      if (fabsf(pre - want) > 0.02f + 0.01f * fabsf(want)) { if (bad < 5) printf("  aux_out mismatch (%d,%d): got %f want %f\n", hi[s], hj[s], pre, want); bad++; }
      want = 0.5f * want * (1.0f + erff(want * 0.70710678f));
    }
    if (epi == 2) want += __bfloat162float(hbias[hj[s]]) + __bfloat162float(haux[idx]);
    if (epi == 3 && !out_fp32) want += __bfloat162float(c_before[idx]);
    float got = out_fp32 ? hc32[idx] : __bfloat162float(hc16[idx]);
    double err = fabs((double)got - want);
    if (err > max_err) max_err = err;
    if (fabs(want) > max_ref) max_ref = fabs(want);
    if (!(err <= 0.02 + 0.01 * fabs(want))) { if (bad < 5) printf("  mismatch (%d,%d): got %f want %f\n", hi[s], hj[s], got, want); bad++; }
  }
  printf("G=%d amn=%d bmn=%d M=%d N=%d K=%d epi=%d fp32=%d : max_err=%.5f max_ref=%.3f bad=%d %s\n", G, amn, bmn, M, N, K,
         epi, out_fp32, max_err, max_ref, bad, bad == 0 ? "PASS" : "FAIL");

  if (iters > 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // L2 flush buffer
    void* flush; size_t fb = 256u << 20; CK(cudaMalloc(&flush, fb));
    for (int w = 0; w < 3; ++w) CK(hb::gemm_bf16(call, 0));
    CK(cudaDeviceSynchronize());
    float total = 0, best = 1e9;
    for (int it = 0; it < iters; ++it) {
      CK(cudaMemsetAsync(flush, it, fb, 0));
      CK(cudaEventRecord(e0, 0));
      CK(hb::gemm_bf16(call, 0));
      CK(cudaEventRecord(e1, 0));
      CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      total += ms; if (ms < best) best = ms;
    }
    double flops = 2.0 * M * N * (double)K;
    printf("  ours  : avg %.4f ms (%.1f TFLOP/s)  best %.4f ms (%.1f TFLOP/s)\n", total / iters,
           flops / (total / iters) * 1e-9, best, flops / best * 1e-9);
    // cuBLAS reference timing (same math: C = A * B)
    cublasHandle_t h; cublasCreate(&h);
    float one = 1.f, zero = 0.f;
    // row-major C[M,N] = A*B  <=> col-major C^T[N,M] = B^T * A^T
    cublasOperation_t opB = bmn ? CUBLAS_OP_N : CUBLAS_OP_T;   // first operand (B)
    cublasOperation_t opA = amn ? CUBLAS_OP_T : CUBLAS_OP_N;   // second operand (A)
    auto run_cublas = [&]() {
      return cublasGemmEx(h, opB, opA, N, M, K, &one, B, CUDA_R_16BF, (int)ldb, A, CUDA_R_16BF, (int)lda, &zero, C,
                          out_fp32 ? CUDA_R_32F : CUDA_R_16BF, (int)ldc, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT);
    };
    for (int w = 0; w < 3; ++w) run_cublas();
    CK(cudaDeviceSynchronize());
    total = 0; best = 1e9;
    for (int it = 0; it < iters; ++it) {
      CK(cudaMemsetAsync(flush, it, fb, 0));
      CK(cudaEventRecord(e0, 0));
      cublasStatus_t st = run_cublas();
      CK(cudaEventRecord(e1, 0));
      CK(cudaEventSynchronize(e1));
      if (st != CUBLAS_STATUS_SUCCESS) { printf("  cublas failed %d\n", (int)st); break; }
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      total += ms; if (ms < best) best = ms;
    }
    printf("  cublas: avg %.4f ms (%.1f TFLOP/s)  best %.4f ms (%.1f TFLOP/s)\n", total / iters,
           flops / (total / iters) * 1e-9, best, flops / best * 1e-9);
  }
  return bad == 0 ? 0 : 1;
}
