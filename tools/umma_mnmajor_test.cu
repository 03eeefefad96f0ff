// Unit test of an MN-major B operand for tcgen05.mma (stepping stone (a) of DESIGN Appendix A):
//   D[128, n] = P[128, 128] * X[128, n]      P K-major (row-major, K contiguous), X row-major [tokens, n] = MN-major B
// Both operands arrive by TMA with the 128-byte swizzle; the B tile is used exactly as TMA delivers it, with the
// B-transpose bit of the instruction descriptor and an MN-major shared-memory descriptor
// (SBO = 1024 B between 8-token groups, LBO = distance between 64-column blocks).  Dev tool, not part of the library.
//   umma_mnmajor_test <n: 64|128> [1 = A tile written by threads + proxy fence instead of TMA]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "tdc_ptx.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

using namespace tdc;

constexpr int kM = 128, kK = 128, kTile = 128 * 64 * 2;  // one [128 x 64] bf16 box = 16 KB

__global__ void __launch_bounds__(128) mn_test_kernel(const __grid_constant__ CUtensorMap map_p,
                                                      const __grid_constant__ CUtensorMap map_x, float* d, int n,
                                                      const __nv_bfloat16* p_manual, int mma_m) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
  uint8_t* a_tile = base;               // 2 x 16 KB: K chunks 0..63, 64..127
  uint8_t* b_tile = base + 2 * kTile;   // n/64 x 16 KB: column blocks
  __shared__ uint64_t bar_full, bar_done;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<1>(&tmem_base_s, 128);
  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_done, 1);
    fence_mbar_init();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (p_manual != nullptr) {
    // stepping stone (c): the A tile is WRITTEN BY THREADS (as the mma.sync warps of the fused kernel would) in the
    // K-major 128B-swizzled layout -- row r at r*128 B, its 16-byte piece j at ((j ^ (r & 7)) * 16) -- and handed to
    // the tensor core after a generic->async proxy fence, instead of arriving by TMA.
    const int r = threadIdx.x;
    for (int kc = 0; kc < 2; ++kc)
      for (int j = 0; j < 8; ++j) {
        const uint4 v = *reinterpret_cast<const uint4*>(p_manual + r * kK + kc * 64 + j * 8);
        *reinterpret_cast<uint4*>(a_tile + kc * kTile + r * 128 + ((j ^ (r & 7)) * 16)) = v;
      }
    fence_proxy_async_smem();
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_full, ((p_manual ? 0 : 2) + n / 64) * kTile);
    if (p_manual == nullptr) {
      tma_load_2d(a_tile, &map_p, &bar_full, 0, 0);
      tma_load_2d(a_tile + kTile, &map_p, &bar_full, 64, 0);
    }
    for (int nb = 0; nb < n / 64; ++nb) tma_load_2d(b_tile + nb * kTile, &map_x, &bar_full, nb * 64, 0);
    mbar_wait(&bar_full, 0);
    tc_fence_after_sync();
    const uint32_t idesc = make_idesc_bf16_f32(mma_m, n) | kIdescBMnMajor;
    for (int s = 0; s < kK / 16; ++s) {
      const uint64_t da = make_kmajor_sw128_desc(smem_u32(a_tile + (s / 4) * kTile)) + 2 * (s % 4);
      const uint64_t db = make_mnmajor_sw128_desc(smem_u32(b_tile) + 2048u * s, kTile, 1024);
      umma_f16<1>(tmem, da, db, idesc, s > 0 ? 1u : 0u);
    }
    umma_commit<1>(&bar_done);
  }
  mbar_wait(&bar_done, 0);
  tc_fence_after_sync();
  for (int c0 = 0; c0 < n; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(32 * warp) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) d[(32 * warp + lane) * n + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmem, 128);
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_map(CUtensorMap* map, const void* ptr, int rows, int cols) {   // bf16 row-major, boxes 128 rows x 64 cols
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
  const cuuint32_t box[2] = {64, 128}, estr[2] = {1, 1};
  return reinterpret_cast<EncodeTiledFn>(p)(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box,
                                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 64;
  const bool manual_a = argc > 2 && atoi(argv[2]) == 1;
  const int mma_m = argc > 3 ? atoi(argv[3]) : 128;   // 64: discover which TMEM lanes hold the 64 rows of D
  if (n != 64 && n != 128) { printf("n must be 64 or 128\n"); return 1; }
  std::mt19937 rng(7);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<__nv_bfloat16> hp(kM * kK), hx(kK * n);
  for (auto& v : hp) v = __float2bfloat16(nd(rng));
  for (auto& v : hx) v = __float2bfloat16(nd(rng));
  __nv_bfloat16 *dp, *dx; float* dd;
  CK(cudaMalloc(&dp, hp.size() * 2)); CK(cudaMalloc(&dx, hx.size() * 2)); CK(cudaMalloc(&dd, kM * n * 4));
  CK(cudaMemcpy(dp, hp.data(), hp.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0xFF, kM * n * 4));
  CUtensorMap mp, mx;
  if (!make_map(&mp, dp, kM, kK) || !make_map(&mx, dx, kK, n)) { printf("tensor map failed\n"); return 3; }
  const int smem = (2 + n / 64) * kTile + 1024;
  CK(cudaFuncSetAttribute(mn_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mn_test_kernel<<<1, 128, smem>>>(mp, mx, dd, n, manual_a ? dp : nullptr, mma_m);
  CK(cudaDeviceSynchronize());
  std::vector<float> hd(kM * n);
  CK(cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost));
  if (mma_m == 64) {
    // match every TMEM lane against the 64 reference rows
    std::vector<int> lane_row(kM, -1);
    for (int l = 0; l < kM; ++l)
      for (int i = 0; i < 64 && lane_row[l] < 0; ++i) {
        bool ok = true;
        for (int j = 0; j < n && ok; ++j) {
          double acc = 0;
          for (int t = 0; t < kK; ++t) acc += (double)__bfloat162float(hp[i * kK + t]) * (double)__bfloat162float(hx[t * n + j]);
          ok = std::fabs(acc - hd[l * n + j]) <= 1e-3 * (1.0 + std::fabs(acc));
        }
        if (ok) lane_row[l] = i;
      }
    printf("M=64 accumulator layout (TMEM lane -> D row, '-' = unused):\n");
    for (int l = 0; l < kM; ++l) { if (lane_row[l] >= 0) printf("%d:%d ", l, lane_row[l]); else printf("%d:- ", l); if (l % 16 == 15) printf("\n"); }
    return 0;
  }
  double max_err = 0, max_ref = 0;
  for (int i = 0; i < kM; ++i)
    for (int j = 0; j < n; ++j) {
      double acc = 0;
      for (int t = 0; t < kK; ++t) acc += (double)__bfloat162float(hp[i * kK + t]) * (double)__bfloat162float(hx[t * n + j]);
      max_err = std::fmax(max_err, std::fabs(acc - hd[i * n + j]));
      max_ref = std::fmax(max_ref, std::fabs(acc));
    }
  printf("umma_mnmajor_test n=%d A=%s: max_abs_err=%.6f max_ref=%.3f => %s\n", n, manual_a ? "thread-written" : "TMA", max_err, max_ref,
         max_err <= 1e-3 * max_ref ? "PASS" : "FAIL");
  return max_err <= 1e-3 * max_ref ? 0 : 4;
}
