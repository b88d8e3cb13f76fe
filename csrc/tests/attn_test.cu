// Standalone correctness + timing harness for the tcgen05 flash-attention kernels.
//   attn_test B Hq Hkv Sq Sk D causal [time_iters]
// Reference: fp32 math on the same bf16 inputs, materialising P (small shapes only).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../kernels/attention_sm100.h"

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
  p[i] = __float2bfloat16(((x & 0xFFFF) / 65536.0f - 0.5f) * 2.0f * scale);
}

// layout [B, S, H, D] contiguous
__device__ __forceinline__ size_t idx4(int b, int s, int h, int d, int S, int H, int D) {
  return (((size_t)b * S + s) * H + h) * D + d;
}

// P[b,h,i,j] (fp32) and LSE
__global__ void ref_scores(const __nv_bfloat16* q, const __nv_bfloat16* k, float* P, float* lse, int B, int Hq, int Hkv,
                           int Sq, int Sk, int D, float scale, int causal) {
  int i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  int hk = h / (Hq / Hkv);
  extern __shared__ float sh[];
  float* row = P + (((size_t)b * Hq + h) * Sq + i) * Sk;
  int off = Sk - Sq;
  for (int j = threadIdx.x; j < Sk; j += blockDim.x) {
    float acc = 0.f;
    for (int d = 0; d < D; ++d)
      acc += __bfloat162float(q[idx4(b, i, h, d, Sq, Hq, D)]) * __bfloat162float(k[idx4(b, j, hk, d, Sk, Hkv, D)]);
    acc *= scale;
    if (causal && j > i + off) acc = -INFINITY;
    row[j] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = -INFINITY;
    for (int j = 0; j < Sk; ++j) m = fmaxf(m, row[j]);
    float s = 0.f;
    for (int j = 0; j < Sk; ++j) s += expf(row[j] - m);
    sh[0] = m; sh[1] = s;
    lse[((size_t)b * Hq + h) * Sq + i] = m + logf(s);
  }
  __syncthreads();
  float m = sh[0], s = sh[1];
  for (int j = threadIdx.x; j < Sk; j += blockDim.x) row[j] = expf(row[j] - m) / s;
}
__global__ void ref_out(const float* P, const __nv_bfloat16* v, float* O, int B, int Hq, int Hkv, int Sq, int Sk, int D) {
  int i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  int hk = h / (Hq / Hkv);
  const float* row = P + (((size_t)b * Hq + h) * Sq + i) * Sk;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < Sk; ++j) acc += row[j] * __bfloat162float(v[idx4(b, j, hk, d, Sk, Hkv, D)]);
    O[idx4(b, i, h, d, Sq, Hq, D)] = acc;
  }
}
// dS[b,h,i,j] = P * (dP - delta) ; dP = dO . V ; delta = sum_d dO * O(ref fp32)
__global__ void ref_ds(const float* P, const __nv_bfloat16* d_o, const __nv_bfloat16* v, const float* Oref, float* dS,
                       int B, int Hq, int Hkv, int Sq, int Sk, int D) {
  int i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  int hk = h / (Hq / Hkv);
  __shared__ float delta;
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int d = 0; d < D; ++d) a += __bfloat162float(d_o[idx4(b, i, h, d, Sq, Hq, D)]) * Oref[idx4(b, i, h, d, Sq, Hq, D)];
    delta = a;
  }
  __syncthreads();
  const float* prow = P + (((size_t)b * Hq + h) * Sq + i) * Sk;
  float* drow = dS + (((size_t)b * Hq + h) * Sq + i) * Sk;
  for (int j = threadIdx.x; j < Sk; j += blockDim.x) {
    float dp = 0.f;
    for (int d = 0; d < D; ++d)
      dp += __bfloat162float(d_o[idx4(b, i, h, d, Sq, Hq, D)]) * __bfloat162float(v[idx4(b, j, hk, d, Sk, Hkv, D)]);
    drow[j] = prow[j] * (dp - delta);
  }
}
__global__ void ref_dq(const float* dS, const __nv_bfloat16* k, float* dQ, int B, int Hq, int Hkv, int Sq, int Sk, int D,
                       float scale) {
  int i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  int hk = h / (Hq / Hkv);
  const float* drow = dS + (((size_t)b * Hq + h) * Sq + i) * Sk;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < Sk; ++j) acc += drow[j] * __bfloat162float(k[idx4(b, j, hk, d, Sk, Hkv, D)]);
    dQ[idx4(b, i, h, d, Sq, Hq, D)] = acc * scale;
  }
}
__global__ void ref_dkv(const float* P, const float* dS, const __nv_bfloat16* q, const __nv_bfloat16* d_o, float* dK,
                        float* dV, int B, int Hq, int Hkv, int Sq, int Sk, int D, float scale) {
  int j = blockIdx.x, hk = blockIdx.y, b = blockIdx.z;
  int g = Hq / Hkv;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float ak = 0.f, av = 0.f;
    for (int hh = 0; hh < g; ++hh) {
      int h = hk * g + hh;
      for (int i = 0; i < Sq; ++i) {
        size_t pi = (((size_t)b * Hq + h) * Sq + i) * Sk + j;
        ak += dS[pi] * __bfloat162float(q[idx4(b, i, h, d, Sq, Hq, D)]);
        av += P[pi] * __bfloat162float(d_o[idx4(b, i, h, d, Sq, Hq, D)]);
      }
    }
    dK[idx4(b, j, hk, d, Sk, Hkv, D)] = ak * scale;
    dV[idx4(b, j, hk, d, Sk, Hkv, D)] = av;
  }
}

static bool compare(const char* name, const std::vector<__nv_bfloat16>& got, const std::vector<float>& want) {
  double max_err = 0, max_ref = 0; size_t bad = 0;
  for (size_t i = 0; i < want.size(); ++i) {
    float g = __bfloat162float(got[i]);
    double e = fabs((double)g - want[i]);
    if (e > max_err) max_err = e;
    if (fabs(want[i]) > max_ref) max_ref = fabs(want[i]);
    if (!(e <= 0.02 * 1.0 + 0.03 * fabs(want[i]))) { if (bad < 3) printf("   %s[%zu] got %f want %f\n", name, i, g, want[i]); bad++; }
  }
  printf("  %-4s max_err=%.5f max_ref=%.4f bad=%zu %s\n", name, max_err, max_ref, bad, bad == 0 ? "PASS" : "FAIL");
  return bad == 0;
}

int main(int argc, char** argv) {
  if (argc < 8) { printf("usage: attn_test B Hq Hkv Sq Sk D causal [iters] [skip_check]\n"); return 1; }
  int B = atoi(argv[1]), Hq = atoi(argv[2]), Hkv = atoi(argv[3]), Sq = atoi(argv[4]), Sk = atoi(argv[5]), D = atoi(argv[6]);
  int causal = atoi(argv[7]);
  int iters = argc > 8 ? atoi(argv[8]) : 0;
  int skip = argc > 9 ? atoi(argv[9]) : 0;
  float scale = 1.0f / sqrtf((float)D);
  size_t nq = (size_t)B * Sq * Hq * D, nk = (size_t)B * Sk * Hkv * D;
  __nv_bfloat16 *q, *k, *v, *o, *d_o, *dq, *dk, *dv;
  float *lse, *delta;
  CK(cudaMalloc(&q, nq * 2)); CK(cudaMalloc(&k, nk * 2)); CK(cudaMalloc(&v, nk * 2)); CK(cudaMalloc(&o, nq * 2));
  CK(cudaMalloc(&d_o, nq * 2)); CK(cudaMalloc(&dq, nq * 2)); CK(cudaMalloc(&dk, nk * 2)); CK(cudaMalloc(&dv, nk * 2));
  CK(cudaMalloc(&lse, (size_t)B * Hq * Sq * 4)); CK(cudaMalloc(&delta, (size_t)B * Hq * Sq * 4));
  fill_kernel<<<(nq + 255) / 256, 256>>>(q, nq, 1u, 1.0f);
  fill_kernel<<<(nk + 255) / 256, 256>>>(k, nk, 5u, 1.0f);
  fill_kernel<<<(nk + 255) / 256, 256>>>(v, nk, 9u, 1.0f);
  fill_kernel<<<(nq + 255) / 256, 256>>>(d_o, nq, 11u, 1.0f);
  CK(cudaMemset(o, 0xFF, nq * 2)); CK(cudaMemset(dq, 0xFF, nq * 2)); CK(cudaMemset(dk, 0xFF, nk * 2)); CK(cudaMemset(dv, 0xFF, nk * 2));
  CK(cudaDeviceSynchronize());

  auto mk = [&](const void* p, int S, int H) { hb::AttnTensor t; t.ptr = p; t.stride_h = D; t.stride_s = (int64_t)H * D; t.stride_b = (int64_t)S * H * D; return t; };
  hb::AttnFwdCall f;
  f.q = mk(q, Sq, Hq); f.k = mk(k, Sk, Hkv); f.v = mk(v, Sk, Hkv); f.o = mk(o, Sq, Hq); f.lse = lse;
  f.B = B; f.Hq = Hq; f.Hkv = Hkv; f.Sq = Sq; f.Sk = Sk; f.D = D; f.softmax_scale = scale; f.causal = causal;
  CK(hb::attn_fwd(f, 0));
  CK(cudaDeviceSynchronize());
  hb::AttnBwdCall g;
  g.q = f.q; g.k = f.k; g.v = f.v; g.o = f.o; g.d_o = mk(d_o, Sq, Hq); g.dq = mk(dq, Sq, Hq); g.dk = mk(dk, Sk, Hkv); g.dv = mk(dv, Sk, Hkv);
  g.lse = lse; g.delta = delta; g.B = B; g.Hq = Hq; g.Hkv = Hkv; g.Sq = Sq; g.Sk = Sk; g.D = D; g.softmax_scale = scale; g.causal = causal;
  CK(hb::attn_bwd(g, 0));
  CK(cudaDeviceSynchronize());

  bool ok = true;
  if (!skip) {
    float *P, *dS, *Oref, *lse_ref, *dQr, *dKr, *dVr;
    size_t np = (size_t)B * Hq * Sq * Sk;
    CK(cudaMalloc(&P, np * 4)); CK(cudaMalloc(&dS, np * 4)); CK(cudaMalloc(&Oref, nq * 4)); CK(cudaMalloc(&lse_ref, (size_t)B * Hq * Sq * 4));
    CK(cudaMalloc(&dQr, nq * 4)); CK(cudaMalloc(&dKr, nk * 4)); CK(cudaMalloc(&dVr, nk * 4));
    dim3 gq(Sq, Hq, B), gk(Sk, Hkv, B);
    ref_scores<<<gq, 128, 16>>>(q, k, P, lse_ref, B, Hq, Hkv, Sq, Sk, D, scale, causal);
    ref_out<<<gq, 128>>>(P, v, Oref, B, Hq, Hkv, Sq, Sk, D);
    ref_ds<<<gq, 128>>>(P, d_o, v, Oref, dS, B, Hq, Hkv, Sq, Sk, D);
    ref_dq<<<gq, 128>>>(dS, k, dQr, B, Hq, Hkv, Sq, Sk, D, scale);
    ref_dkv<<<gk, 128>>>(P, dS, q, d_o, dKr, dVr, B, Hq, Hkv, Sq, Sk, D, scale);
    CK(cudaDeviceSynchronize());
    std::vector<__nv_bfloat16> ho(nq), hdq(nq), hdk(nk), hdv(nk);
    std::vector<float> ro(nq), rdq(nq), rdk(nk), rdv(nk), hl((size_t)B * Hq * Sq), rl((size_t)B * Hq * Sq);
    CK(cudaMemcpy(ho.data(), o, nq * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ro.data(), Oref, nq * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hdq.data(), dq, nq * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(rdq.data(), dQr, nq * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hdk.data(), dk, nk * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(rdk.data(), dKr, nk * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hdv.data(), dv, nk * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(rdv.data(), dVr, nk * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hl.data(), lse, hl.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(rl.data(), lse_ref, rl.size() * 4, cudaMemcpyDeviceToHost));
    printf("B=%d Hq=%d Hkv=%d Sq=%d Sk=%d D=%d causal=%d\n", B, Hq, Hkv, Sq, Sk, D, causal);
    ok &= compare("O", ho, ro);
    double le = 0; for (size_t i = 0; i < hl.size(); ++i) if (isfinite(rl[i])) le = fmax(le, fabs((double)hl[i] - rl[i]));
    printf("  LSE  max_err=%.5f %s\n", le, le < 0.02 ? "PASS" : "FAIL"); ok &= le < 0.02;
    ok &= compare("dQ", hdq, rdq);
    ok &= compare("dK", hdk, rdk);
    ok &= compare("dV", hdv, rdv);
  }
  if (iters > 0) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) { CK(hb::attn_fwd(f, 0)); CK(hb::attn_bwd(g, 0)); }
    CK(cudaDeviceSynchronize());
    double fl = 4.0 * B * Hq * (double)Sq * Sk * D * (causal ? 0.5 : 1.0);
    float ms;
    CK(cudaEventRecord(e0)); for (int i = 0; i < iters; ++i) CK(hb::attn_fwd(f, 0)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= iters;
    printf("  fwd: %.4f ms  %.1f TFLOP/s\n", ms, fl / ms * 1e-9);
    CK(cudaEventRecord(e0)); for (int i = 0; i < iters; ++i) CK(hb::attn_bwd(g, 0)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= iters;
    printf("  bwd: %.4f ms  %.1f TFLOP/s (2.5x fwd flops)\n", ms, 2.5 * fl / ms * 1e-9);
  }
  printf("%s\n", ok ? "ALL PASS" : "SOME FAIL");
  return ok ? 0 : 1;
}
