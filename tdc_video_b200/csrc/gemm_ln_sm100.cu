// tcgen05 GEMM with the post-LN block's tail fused into its epilogue (sm_100a):
//
//     h = LayerNorm( A . W^T + bias + h ) * gamma + beta          (fp32 residual stream, updated in place)
//     h_bf16 = bf16(h)                                            (the next GEMM's operand)
//
// i.e. BertSelfOutput / BertOutput of the reference (tdc/Qformer.py:285-289, 371-375:
// `LayerNorm(dense(x) + input_tensor)`) in ONE kernel instead of GEMM -> fp32 `pre` -> LayerNorm kernel.
// What disappears: the fp32 `pre` write + read (8 of the 18 bytes per element those sub-layers moved) and
// one kernel launch per sub-layer (30 per forward pass).
//
// A LayerNorm row spans the whole output width N (768), more than one CTA's TMEM can hold next to a second,
// double-buffered accumulator (2 x 768 fp32 columns > 512).  So one 128-row tile is computed by a CLUSTER of
// CL = ceil(N / BLOCK_N) CTAs (3 for N = 768), each owning BLOCK_N = 256 output columns, and the row
// statistics are combined through distributed shared memory:
//
//   mainloop   as in gemm_sm100.cu (cta_group::1): warp 0 = TMA producer, warp 1 = tcgen05.mma issuer,
//              fp32 accumulator [128 x 256] in TMEM, double buffered;
//   pass 1     epilogue warps (lane quadrant x column half): x = acc + bias + resid, written back into the
//              accumulator's own TMEM columns (tcgen05.st); per-thread shifted sums -> (mean, M2) of its
//              128 columns -> st.async (remote store + mbarrier complete_tx) into EVERY CTA of the cluster;
//   combine    each thread merges the 2 * CL partials of its row (Chan et al. parallel variance);
//   pass 2     y = (x - mean) * rstd * gamma + beta from TMEM -> fp32 and bf16 staging tiles -> TMA stores.
//
// Statistics stay fp32 end to end (eps = 1e-12 vanishes in bf16), merge is exact up to fp32 rounding.
#include "tdc_gemm.cuh"
#include "tdc_ptx.cuh"
#include "tdc_b200.h"

#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace tdc {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kNumEpilogueWarps = 8;
constexpr int kNumThreads = 64 + 32 * kNumEpilogueWarps;
constexpr int kAccStages = 2;
constexpr int kStoreTileBytes = 32 * 128;
constexpr int kMaxCluster = 3;  // N <= 3 * BLOCK_N (768: the Q-Former width); wider rows use the two-kernel form

struct LnArgs {
  const float* bias;
  const float* resid;  // fp32 [M, N], pitch ldr (may alias the fp32 output)
  long long ldr;
  const float* gamma;
  const float* beta;
  float eps;
};

template <int BLOCK_N, int STAGES>
struct LnSmem {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingOffset = STAGES * kStageBytes;
  static constexpr int kStagingBytes = kNumEpilogueWarps * 2 * kStoreTileBytes;  // per warp: one fp32 + one bf16 tile
  // stats[buf][src cta][half][row] = (mean, M2) of 128-row x (BLOCK_N / 2)-column pieces
  static constexpr int kStatsOffset = kStagingOffset + kStagingBytes;
  static constexpr int kStatsBytes = kAccStages * kMaxCluster * 2 * kBlockM * 8;
  // this CTA's BLOCK_N columns of bias / gamma / beta (read by every row of every tile)
  static constexpr int kVecOffset = kStatsOffset + kStatsBytes;
  static constexpr int kVecBytes = 3 * BLOCK_N * 4;
  static constexpr int kBarrierOffset = kVecOffset + kVecBytes;
  // full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], stats[2][4 quads], tmem base ptr
  static constexpr int kBarrierBytes = (2 * STAGES + 2 * kAccStages + kAccStages * 4) * 8 + 16;
  static constexpr int kTotalBytes = kBarrierOffset + kBarrierBytes + 1024;
};

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Store 8 bytes into the shared memory of CTA `cta` of the cluster (same offset as `local`) and credit them to
// the mbarrier at the same offset as `bar` in that CTA — remote store and signal in one instruction.
__device__ __forceinline__ void st_async_f32x2(void* local, uint64_t* bar, uint32_t cta, float a, float b) {
  asm volatile(
      "{\n\t.reg .b32 ra, rb;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %2;\n\t"
      "mapa.shared::cluster.u32 rb, %1, %2;\n\t"
      "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [ra], {%3, %4}, [rb];\n\t}" ::"r"(
          smem_u32(local)),
      "r"(smem_u32(bar)), "r"(cta), "r"(__float_as_uint(a)), "r"(__float_as_uint(b))
      : "memory");
}

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kNumThreads, 1)
tdc_gemm_ln_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                   const __grid_constant__ CUtensorMap map_f32, const __grid_constant__ CUtensorMap map_bf16, int m,
                   int n, int k, int cluster_size, LnArgs ln) {
  using L = LnSmem<BLOCK_N, STAGES>;
  constexpr uint32_t kTmemCols = kAccStages * BLOCK_N;
  constexpr uint32_t kIdesc = make_idesc_bf16_f32(kBlockM, BLOCK_N);
  constexpr int kHalfCols = BLOCK_N / 2;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float2* stats = reinterpret_cast<float2*>(smem + L::kStatsOffset);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarrierOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + kAccStages;
  uint64_t* stats_bar = tmem_empty_bar + kAccStages;  // [kAccStages][4 quads]
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(stats_bar + kAccStages * 4);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();         // = this CTA's N tile
  const int num_m_tiles = (m + kBlockM - 1) / kBlockM;
  const int num_kb = (k + kBlockK - 1) / kBlockK;
  const int first_tile = blockIdx.x / cluster_size;
  const int tile_stride = gridDim.x / cluster_size;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_f32);
    tma_prefetch_desc(&map_bf16);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], kNumEpilogueWarps);
      // cluster: one expect_tx arrive per use, the partials arrive as transaction bytes (st.async);
      // single CTA: plain shared-memory stores, every thread of the quadrant's two warps arrives
      for (int q = 0; q < 4; ++q) mbar_init(&stats_bar[s * 4 + q], cluster_size > 1 ? 1 : 64);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    __syncwarp();
    tmem_alloc<1>(tmem_base_smem, kTmemCols);
    tmem_relinquish<1>();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // every CTA's stats barriers exist before any peer's st.async can land
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int row_w = static_cast<int>(cta) * BLOCK_N;
      for (int tile = first_tile; tile < num_m_tiles; tile += tile_stride) {
        const int row_a = tile * kBlockM;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
          tma_load_2d(sa, &map_a, &full_bar[stage], kb * kBlockK, row_a, kL2EvictNormal);
          tma_load_2d(sa + L::kABytes, &map_w, &full_bar[stage], kb * kBlockK, row_w, kL2EvictLast);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = first_tile; tile < num_m_tiles; tile += tile_stride) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint64_t desc_a = make_kmajor_sw128_desc(sa);
          const uint64_t desc_b = make_kmajor_sw128_desc(sa + L::kABytes);
#pragma unroll
          for (int kk = 0; kk < kBlockK / kUmmaK; ++kk)
            umma_f16<1>(d_tmem, desc_a + 2u * kk, desc_b + 2u * kk, kIdesc, (kb | kk) != 0 ? 1u : 0u);
          umma_commit<1>(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit<1>(&tmem_full_bar[acc]);
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const uint32_t quad = warp & 3;
    const uint32_t half = (warp - 2) >> 2;
    uint8_t* stage_f32 = smem + L::kStagingOffset + (warp - 2) * (2 * kStoreTileBytes);
    uint8_t* stage_bf16 = stage_f32 + kStoreTileBytes;
    const int col_base = static_cast<int>(cta) * BLOCK_N + static_cast<int>(half) * kHalfCols;
    // valid columns of this thread's piece and of every piece of the row (the same for all rows)
    const int my_cnt = max(0, min(kHalfCols, n - col_base));
    const float inv_n = 1.0f / static_cast<float>(n);
    const uint32_t sw = lane & 7;
    // bias / gamma / beta of this CTA's columns -> shared memory once (zeros beyond n)
    float* vec_s = reinterpret_cast<float*>(smem + L::kVecOffset);
    for (int i = static_cast<int>(threadIdx.x) - 64; i < 3 * BLOCK_N; i += 32 * kNumEpilogueWarps) {
      const int which = i / BLOCK_N, col = static_cast<int>(cta) * BLOCK_N + i % BLOCK_N;
      const float* src = which == 0 ? ln.bias : (which == 1 ? ln.gamma : ln.beta);
      vec_s[i] = col < n ? __ldg(src + col) : 0.f;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kNumEpilogueWarps) : "memory");  // the epilogue warps only
    const float4* bias_s = reinterpret_cast<const float4*>(vec_s + half * kHalfCols);
    const float4* gamma_s = reinterpret_cast<const float4*>(vec_s + BLOCK_N + half * kHalfCols);
    const float4* beta_s = reinterpret_cast<const float4*>(vec_s + 2 * BLOCK_N + half * kHalfCols);
    constexpr int kChunks = kHalfCols / 32;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = first_tile; tile < num_m_tiles; tile += tile_stride) {
      const int row_in_tile = static_cast<int>(quad) * 32 + static_cast<int>(lane);
      const int row = tile * kBlockM + row_in_tile;
      const bool row_ok = row < m;
      uint64_t* sbar = &stats_bar[acc * 4 + quad];
      // the partials of this (tile, quad): 2 halves x CL CTAs x 32 rows x 8 bytes, announced once per use
      if (cluster_size > 1 && half == 0 && lane == 0)
        mbar_arrive_expect_tx(sbar, static_cast<uint32_t>(2 * cluster_size * 32 * 8));
      // The residual does not depend on the MMAs: fetch this thread's row piece (kHalfCols fp32, one 128-byte line
      // per 32 columns) while the mainloop of this tile is still running; two 32-column chunks stay in flight.
      const float4* rrow = reinterpret_cast<const float4*>(ln.resid + static_cast<long long>(row) * ln.ldr + col_base);
      float4 rbuf[2][8];
      auto fetch = [&](int c, float4 (&dst)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          dst[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok && col_base + c * 32 + 4 * j < n) dst[j] = __ldg(rrow + c * 8 + j);
        }
      };
      fetch(0, rbuf[0]);
      if (kChunks > 1) fetch(1, rbuf[1]);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t t_addr = tmem_base + ((quad * 32u) << 16) + static_cast<uint32_t>(acc * BLOCK_N) + half * kHalfCols;

      // ---- pass 1: x = acc + bias + resid -> back into TMEM; shifted sums of the valid columns
      float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        uint32_t v[32];
        const int col0 = col_base + c * 32;
        tmem_ld_32x32(t_addr + c * 32, v);
        float r[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {   // resid + bias (zeros beyond n)
          const float4 t = rbuf[c & 1][j], b = bias_s[c * 8 + j];
          r[4 * j] = t.x + b.x; r[4 * j + 1] = t.y + b.y; r[4 * j + 2] = t.z + b.z; r[4 * j + 3] = t.w + b.w;
        }
        if (c + 2 < kChunks) fetch(c + 2, rbuf[c & 1]);
        tmem_ld_wait();
        if (c == 0) shift = __uint_as_float(v[0]) + r[0];   // column col_base: valid whenever my_cnt > 0
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = __uint_as_float(v[j]) + r[j];
          v[j] = __float_as_uint(x);
          if (col0 + j < n) {
            const float d = x - shift;
            s1 += d;
            s2 = fmaf(d, d, s2);
          }
        }
        tmem_st_32x32(t_addr + c * 32, v);
      }
      tmem_st_wait();
      {
        float mean_l = 0.f, m2_l = 0.f;
        if (my_cnt > 0) {
          const float inv = 1.0f / static_cast<float>(my_cnt);
          mean_l = shift + s1 * inv;
          m2_l = fmaxf(s2 - s1 * s1 * inv, 0.f);
        }
        // every CTA of the cluster (this one included) receives this thread's partial
        float2* slot = stats + ((acc * kMaxCluster + cta) * 2 + half) * kBlockM + row_in_tile;
        if (cluster_size > 1) {
          for (int dst = 0; dst < cluster_size; ++dst) st_async_f32x2(slot, sbar, static_cast<uint32_t>(dst), mean_l, m2_l);
        } else {  // narrow rows (N <= BLOCK_N): the tile lives in one CTA, no distributed shared memory involved
          *slot = make_float2(mean_l, m2_l);
          mbar_arrive(sbar);
        }
      }
      // ---- combine the 2 * CL partials of this thread's row
      mbar_wait(sbar, acc_phase);
      float mean = 0.f;
      for (int j = 0; j < cluster_size; ++j)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int cnt = max(0, min(kHalfCols, n - (j * BLOCK_N + hh * kHalfCols)));
          mean += static_cast<float>(cnt) * stats[((acc * kMaxCluster + j) * 2 + hh) * kBlockM + row_in_tile].x;
        }
      mean *= inv_n;
      float m2 = 0.f;
      for (int j = 0; j < cluster_size; ++j)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int cnt = max(0, min(kHalfCols, n - (j * BLOCK_N + hh * kHalfCols)));
          const float2 p = stats[((acc * kMaxCluster + j) * 2 + hh) * kBlockM + row_in_tile];
          const float d = p.x - mean;
          m2 += p.y + static_cast<float>(cnt) * d * d;
        }
      const float rstd = 1.0f / sqrtf(m2 * inv_n + ln.eps);  // biased variance, as nn.LayerNorm

      // ---- pass 2: y = (x - mean) * rstd * gamma + beta -> fp32 + bf16 tiles -> TMA stores
      const bool live = tile * kBlockM + static_cast<int>(quad) * 32 < m;  // warp-uniform
#pragma unroll 1
      for (int c = 0; c < kHalfCols / 64; ++c) {
        const int col0 = col_base + c * 64;
        const bool store_ok = live && col0 < n;  // warp-uniform
        if (store_ok) {
          if (lane == 0) tma_store_wait_read<0>();  // both staging tiles free again
          __syncwarp();
        }
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          uint32_t v[32];
          const int cc = col0 + hlf * 32;
          tmem_ld_32x32(t_addr + c * 64 + hlf * 32, v);
          tmem_ld_wait();
          if (store_ok && cc < n) {
            if (hlf == 1) {  // the fp32 tile is reused for the second 32 columns
              if (lane == 0) tma_store_wait_read<0>();
              __syncwarp();
            }
            uint8_t* rowp = stage_f32 + lane * 128;
            uint8_t* rowb = stage_bf16 + lane * 128;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float y[8], gm[8], bt[8];
#pragma unroll
              for (int q4 = 0; q4 < 2; ++q4) {
                const float4 g4 = gamma_s[(c * 64 + hlf * 32 + g * 8) / 4 + q4];
                const float4 b4 = beta_s[(c * 64 + hlf * 32 + g * 8) / 4 + q4];
                gm[q4 * 4] = g4.x; gm[q4 * 4 + 1] = g4.y; gm[q4 * 4 + 2] = g4.z; gm[q4 * 4 + 3] = g4.w;
                bt[q4 * 4] = b4.x; bt[q4 * 4 + 1] = b4.y; bt[q4 * 4 + 2] = b4.z; bt[q4 * 4 + 3] = b4.w;
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) y[j] = (__uint_as_float(v[g * 8 + j]) - mean) * rstd * gm[j] + bt[j];
              *reinterpret_cast<float4*>(rowp + (((2 * g) ^ sw) << 4)) = make_float4(y[0], y[1], y[2], y[3]);
              *reinterpret_cast<float4*>(rowp + (((2 * g + 1) ^ sw) << 4)) = make_float4(y[4], y[5], y[6], y[7]);
              uint4 pk;
              pk.x = pack_bf16x2(y[0], y[1]);
              pk.y = pack_bf16x2(y[2], y[3]);
              pk.z = pack_bf16x2(y[4], y[5]);
              pk.w = pack_bf16x2(y[6], y[7]);
              *reinterpret_cast<uint4*>(rowb + (((hlf * 4 + g) ^ sw) << 4)) = pk;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&map_f32, stage_f32, cc, tile * kBlockM + static_cast<int>(quad) * 32);
              tma_store_commit();
            }
          }
        }
        if (store_ok) {
          // (the second 32 columns may lie beyond n: TMA clips the box; their staging bytes are stale but unread)
          if (lane == 0) {
            tma_store_2d(&map_bf16, stage_bf16, col0, tile * kBlockM + static_cast<int>(quad) * 32);
            tma_store_commit();
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  // ===================== teardown =====================
  tc_fence_before_sync();
  cluster_sync_all();  // no CTA exits (and frees its shared memory) while a peer may still st.async into it
  if (warp == 1) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc<1>(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

bool make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows, bool f32) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return false;
  const int esz = f32 ? 4 : 2;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * esz};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esz), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
            gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BLOCK_N, int STAGES>
int launch_ln(const GemmLnProblem& p, cudaStream_t stream, const char** err) {
  using L = LnSmem<BLOCK_N, STAGES>;
  const int cluster = (p.n + BLOCK_N - 1) / BLOCK_N;
  CUtensorMap map_a, map_w, map_f32, map_bf16;
  if (!make_map(&map_a, p.a, p.m, p.k, p.lda, kBlockM, false) || !make_map(&map_w, p.w, p.n, p.k, p.ldw, BLOCK_N, false) ||
      !make_map(&map_f32, p.out_f32, p.m, p.n, p.ldo, 32, true) || !make_map(&map_bf16, p.out_bf16, p.m, p.n, p.ldo, 32, false)) {
    if (err) *err = "gemm_ln: cuTensorMapEncodeTiled failed (pointer/pitch alignment?)";
    return TDC_ECUDA;
  }
  auto kernel = tdc_gemm_ln_kernel<BLOCK_N, STAGES>;
  int dev = 0;
  cudaGetDevice(&dev);
  dev = (dev >= 0 && dev < 64) ? dev : 0;
  static std::mutex mu;
  static bool attr_set[64] = {};
  static int sms[64] = {};
  static int max_clusters[64][kMaxCluster + 1] = {};
  int clusters;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!attr_set[dev]) {
      if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotalBytes) != cudaSuccess ||
          cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        if (err) *err = "gemm_ln: cudaFuncSetAttribute failed";
        return TDC_ECUDA;
      }
      cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
      attr_set[dev] = true;
    }
    if (max_clusters[dev][cluster] == 0) {
      int nmax = sms[dev] / cluster;
      cudaLaunchConfig_t q{};
      q.gridDim = dim3(static_cast<unsigned>(nmax * cluster));
      q.blockDim = dim3(kNumThreads);
      q.dynamicSmemBytes = L::kTotalBytes;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = cluster; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa; q.numAttrs = 1;
      int nq = 0;
      if (cluster > 1 && cudaOccupancyMaxActiveClusters(&nq, kernel, &q) == cudaSuccess && nq > 0 && nq < nmax) nmax = nq;
      max_clusters[dev][cluster] = nmax > 0 ? nmax : 1;
    }
    clusters = max_clusters[dev][cluster];
  }
  const int m_tiles = (p.m + kBlockM - 1) / kBlockM;
  if (m_tiles < clusters) clusters = m_tiles;
  static const bool info = [] { const char* e = getenv("TDC_GEMM_LN_INFO"); return e != nullptr && atoi(e) == 1; }();
  if (info)
    fprintf(stderr, "tdc_gemm_ln: m %d n %d k %d -> %d clusters of %d CTAs (%d M tiles), smem %d B\n", p.m, p.n, p.k,
            clusters, cluster, m_tiles, L::kTotalBytes);
  LnArgs ln{p.bias, p.resid, p.ldr, p.gamma, p.beta, p.eps};
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(clusters * cluster));
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = L::kTotalBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t rc = cudaLaunchKernelEx(&cfg, kernel, map_a, map_w, map_f32, map_bf16, p.m, p.n, p.k, cluster, ln);
  if (rc != cudaSuccess) {
    if (err) *err = cudaGetErrorString(rc);
    return TDC_ECUDA;
  }
  return TDC_OK;
}

}  // namespace

bool gemm_ln_supported(int n) { return n > 0 && n % 8 == 0 && n <= kMaxCluster * 256; }  // <= 768

int gemm_ln_launch(const GemmLnProblem& p, cudaStream_t stream, const char** err) {
  if (p.m <= 0 || p.n <= 0 || p.k <= 0) {
    if (err) *err = "gemm_ln: empty problem";
    return TDC_EINVAL;
  }
  if (!gemm_ln_supported(p.n) || (p.k % 8) != 0 || (p.lda % 8) != 0 || (p.ldw % 8) != 0 || (p.ldo % 8) != 0 ||
      (p.ldr % 4) != 0 || (reinterpret_cast<uintptr_t>(p.a) & 15) || (reinterpret_cast<uintptr_t>(p.w) & 15) ||
      (reinterpret_cast<uintptr_t>(p.out_f32) & 15) || (reinterpret_cast<uintptr_t>(p.out_bf16) & 15) ||
      (reinterpret_cast<uintptr_t>(p.resid) & 15)) {
    if (err) *err = "gemm_ln: N <= 768, K / N / pitches multiples of 8 elements, pointers 16-byte aligned";
    return TDC_EINVAL;
  }
  if (p.bias == nullptr || p.resid == nullptr || p.gamma == nullptr || p.beta == nullptr || p.out_f32 == nullptr ||
      p.out_bf16 == nullptr) {
    if (err) *err = "gemm_ln: null pointer";
    return TDC_EINVAL;
  }
  if (p.n <= 128) return launch_ln<128, 4>(p, stream, err);
  return launch_ln<256, 3>(p, stream, err);
}

}  // namespace tdc
