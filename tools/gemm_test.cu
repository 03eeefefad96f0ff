// Standalone check + timing of the tcgen05 GEMM against a CPU double-precision
// reference (sampled rows for big problems).  Dev tool, not part of the library.
//   gemm_test <cta_group> <m> <n> <k> <mode> [iters]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "tdc_gemm.cuh"
#include "tdc_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main(int argc, char** argv) {
  if (argc < 6) { printf("usage: gemm_test cg m n k mode [iters]\n"); return 1; }
  const int cg = atoi(argv[1]), m = atoi(argv[2]), n = atoi(argv[3]), k = atoi(argv[4]), mode = atoi(argv[5]);
  const int iters = argc > 6 ? atoi(argv[6]) : 5;
  printf("gemm_test cg=%d m=%d n=%d k=%d mode=%d\n", cg, m, n, k, mode);
  std::mt19937 rng(1234);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<__nv_bfloat16> ha((size_t)m * k), hw((size_t)n * k);
  std::vector<float> hbias(n), hres;
  for (auto& v : ha) v = __float2bfloat16(nd(rng));
  for (auto& v : hw) v = __float2bfloat16(nd(rng) * 0.05f);
  for (auto& v : hbias) v = nd(rng);
  const bool f32out = (mode == tdc::EPI_BIAS_F32);

  __nv_bfloat16 *da, *dw; float *dbias; void* dout;
  CK(cudaMalloc(&da, ha.size() * 2)); CK(cudaMalloc(&dw, hw.size() * 2)); CK(cudaMalloc(&dbias, n * 4));
  CK(cudaMalloc(&dout, (size_t)m * n * (f32out ? 4 : 2)));
  CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, hbias.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xFF, (size_t)m * n * (f32out ? 4 : 2)));

  tdc::GemmProblem p;
  p.a = da; p.w = dw; p.lda = k; p.ldw = k; p.m = m; p.n = n; p.k = k; p.out = dout; p.ldo = n;
  p.bias = dbias; p.mode = mode; p.cta_group = cg;
  const char* err = nullptr;
  int rc = tdc::gemm_launch(p, 0, &err);
  if (rc != 0) { printf("launch failed rc=%d: %s\n", rc, err ? err : "?"); return 3; }
  CK(cudaDeviceSynchronize());

  // verify
  std::vector<uint8_t> hout((size_t)m * n * (f32out ? 4 : 2));
  CK(cudaMemcpy(hout.data(), dout, hout.size(), cudaMemcpyDeviceToHost));
  std::vector<int> rows;
  if ((double)m * n * k < 4e9) { for (int i = 0; i < m; ++i) rows.push_back(i); }
  else { for (int i = 0; i < 64; ++i) rows.push_back((int)(((long long)i * 7919 * 131) % m)); rows.push_back(m - 1); rows.push_back(0); rows.push_back(127); rows.push_back(128); rows.push_back(m / 2 + 129); }
  double max_err = 0, max_ref = 0; long long bad = 0;
  for (int r : rows) {
    for (int c = 0; c < n; ++c) {
      double acc = 0;
      const __nv_bfloat16* ar = &ha[(size_t)r * k]; const __nv_bfloat16* wr = &hw[(size_t)c * k];
      for (int i = 0; i < k; ++i) acc += (double)__bfloat162float(ar[i]) * (double)__bfloat162float(wr[i]);
      acc += hbias[c];
      if (mode == tdc::EPI_BIAS_GELU_BF16) acc = 0.5 * acc * (1.0 + erf(acc * 0.70710678118654752440));
      double got = f32out ? (double)((float*)hout.data())[(size_t)r * n + c]
                          : (double)__bfloat162float(((__nv_bfloat16*)hout.data())[(size_t)r * n + c]);
      double e = fabs(got - acc); double tol = f32out ? 2e-3 + 1e-4 * fabs(acc) : 2e-2 + 8e-3 * fabs(acc);
      if (!(e <= tol)) { if (bad < 10) printf("  mismatch r=%d c=%d got=%f ref=%f\n", r, c, got, acc); ++bad; }
      if (e > max_err) max_err = e; if (fabs(acc) > max_ref) max_ref = fabs(acc);
    }
  }
  printf("verify: rows=%zu max_abs_err=%.5f max_ref=%.3f bad=%lld => %s\n", rows.size(), max_err, max_ref, bad, bad ? "FAIL" : "PASS");

  // timing
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 2; ++i) tdc::gemm_launch(p, 0, &err);
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) tdc::gemm_launch(p, 0, &err);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= iters;
  printf("time: %.3f ms  %.1f TFLOP/s\n", ms, 2.0 * m * n * (double)k / (ms * 1e-3) / 1e12);
  return bad ? 4 : 0;
}
