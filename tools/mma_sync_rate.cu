// Microbenchmark: issue rate of the legacy tensor path (mma.sync.m16n8k16 bf16 -> fp32) on sm_100a, as dense
// MAC/clk/SM and TFLOP/s -- input of the cost model in DESIGN Appendix A (the per-head M = 16 products of the
// weight-absorbed cross-attention cannot use tcgen05, whose smallest tile is M = 64).  Dev tool.
//   mma_sync_rate [warps_per_block=8] [blocks_per_sm=2]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024) rate_kernel(float* out, int iters, long long* cycles) {
  unsigned a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};   // bf16 1.0 pairs
  unsigned b[2] = {0x3f803f80u, 0x3f803f80u};
  float c[8][4];
  for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main(int argc, char** argv) {
  const int warps = argc > 1 ? atoi(argv[1]) : 8, bps = argc > 2 ? atoi(argv[2]) : 2;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * bps, threads = warps * 32, iters = 20000;
  float* out; long long* cyc;
  cudaMalloc(&out, (size_t)blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
  rate_kernel<<<blocks, threads>>>(out, 100, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  rate_kernel<<<blocks, threads>>>(out, iters, cyc);
  cudaEventRecord(e1);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 2; }
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  long long c0 = 0; cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
  const double mmas = (double)blocks * warps * iters * 8, macs = mmas * 16 * 8 * 16;
  printf("mma.sync m16n8k16 bf16: %d SMs x %d blocks x %d warps: %.3f ms, %.1f TFLOP/s, %.0f MAC/clk/SM (block 0: %lld cycles)\n",
         sms, bps, warps, ms, 2 * macs / (ms * 1e-3) / 1e12, (double)warps * bps * iters * 8 * 2048 / (double)c0, c0);
  return 0;
}
