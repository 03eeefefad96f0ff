// Microbenchmark: how fast can every SM pull tiles that are resident in L2 into shared memory with TMA?
// (the weight-streaming term of the cost model in DESIGN Appendix A).  Every CTA walks the same `mb`-megabyte
// bf16 matrix (L2-resident after the first pass) in 16 KB boxes through an N-deep mbarrier ring; nothing consumes the
// data.  Dev tool.   l2_tma_rate [mb=8] [passes=40] [stages=6, max 13]
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "tdc_ptx.cuh"
using namespace tdc;

constexpr int kMaxStages = 13, kBox = 128 * 64 * 2;

__global__ void __launch_bounds__(64) pull_kernel(const __grid_constant__ CUtensorMap map, int rows, int cols, int passes,
                                                  long long* cycles, int kStages) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[kMaxStages];
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int tiles_r = rows / 128, tiles_c = cols / 64, tiles = tiles_r * tiles_c;
    const int total = tiles * passes;
    const long long t0 = clock64();
    // every CTA starts at a different tile so that the L2 slices are hit evenly; all counters are 32-bit and the
    // ring / tile indices are advanced incrementally (no divisions in the issue loop)
    int issued = 0, done = 0, s_issue = 0, s_done = 0, par_done = 0;
    int t = (blockIdx.x * 37) % tiles, tc = t % tiles_c, tr = t / tiles_c;
    while (done < total) {
      while (issued < total && issued - done < kStages) {
        mbar_arrive_expect_tx(&full[s_issue], kBox);
        tma_load_2d(base + s_issue * kBox, &map, &full[s_issue], tc * 64, tr * 128);
        ++issued;
        if (++s_issue == kStages) s_issue = 0;
        if (++tc == tiles_c) { tc = 0; if (++tr == tiles_r) tr = 0; }
      }
      mbar_wait(&full[s_done], par_done);
      ++done;
      if (++s_done == kStages) { s_done = 0; par_done ^= 1; }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int mb = argc > 1 ? atoi(argv[1]) : 8, passes = argc > 2 ? atoi(argv[2]) : 40;
  const int kStages = argc > 3 ? atoi(argv[3]) : 6;
  const int cols = 1024, rows = mb * (1 << 20) / (cols * 2);
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  __nv_bfloat16* buf; long long* cyc;
  cudaMalloc(&buf, (size_t)rows * cols * 2); cudaMemset(buf, 0, (size_t)rows * cols * 2); cudaMalloc(&cyc, sms * 8);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no encode fn\n"); return 2; }
  CUtensorMap map;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, gstride[1] = {(cuuint64_t)cols * 2};
  const cuuint32_t box[2] = {64, 128}, estr[2] = {1, 1};
  if (reinterpret_cast<EncodeTiledFn>(p)(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gdim, gstride, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("map failed\n"); return 3; }
  const int smem = kStages * kBox + 1024;
  cudaFuncSetAttribute(pull_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  pull_kernel<<<sms, 64, smem>>>(map, rows, cols, 2, cyc, kStages);   // warm the L2
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  pull_kernel<<<sms, 64, smem>>>(map, rows, cols, passes, cyc, kStages);
  cudaEventRecord(e1);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 4; }
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  long long c0 = 0; cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
  const double bytes_per_cta = (double)rows * cols * 2 * passes;
  printf("L2->smem by TMA, %d x 16 KB in flight per SM, %d MB matrix x %d passes per CTA, %d CTAs: %.3f ms, %.1f TB/s aggregate, %.1f GB/s per SM, %.1f B/clk/SM\n",
         kStages, mb, passes, sms, ms, bytes_per_cta * sms / (ms * 1e-3) / 1e12, bytes_per_cta / (ms * 1e-3) / 1e9, bytes_per_cta / (double)c0);
  return 0;
}
