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
// double-buffered accumulator (2 x 768 fp32 columns > 512).  So one 256-row tile is computed by a CLUSTER of
// ceil(N / BLOCK_N) CTA PAIRS (3 pairs = 6 CTAs for N = 768): every pair owns BLOCK_N = 256 output columns and runs
// cta_group::2 UMMAs (256 x 256 x 16; each CTA holds 128 rows of the accumulator and streams half of the W tile,
// 32 KB of operands per k-block and SM instead of 48 KB — these GEMMs run against the L2 -> SM ceiling), and the row
// statistics are combined through distributed shared memory:
//
//   mainloop   as in gemm_sm100.cu (cta_group::2): warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (pair leader),
//              fp32 accumulator [128 x 256] per CTA in TMEM, double buffered;
//   pass 1     epilogue warps (lane quadrant x column half): x = acc + bias + resid, written back into the
//              accumulator's own TMEM columns (tcgen05.st); per-thread shifted sums -> (mean, M2) of its
//              128 columns -> st.async (remote store + mbarrier complete_tx) into the CTA holding the same rows in
//              EVERY pair of the cluster;
//   combine    each thread merges the 2 * pairs partials of its row (Chan et al. parallel variance);
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
  int debug;  // dev knob TDC_GEMM_LN_DEBUG (wrong results by design): 1 = no residual loads, 2 = no output stores,
              // 4 = no statistics exchange (local partials only)
  const float* bias;
  const float* resid;  // fp32 [M, N], pitch ldr (may alias the fp32 output)
  long long ldr;
  const float* gamma;
  const float* beta;
  float eps;
};

template <int BLOCK_N, int STAGES, int TRANS_ROWS>
struct LnSmem {
  static constexpr int kBRows = BLOCK_N / 2;                 // each CTA of a pair holds half of the W tile
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = kBRows * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingOffset = STAGES * kStageBytes;
  static constexpr int kStagingBytes = kNumEpilogueWarps * 2 * kStoreTileBytes;  // per warp: one fp32 + one bf16 tile
  // per warp: one TRANS_ROWS x 128-byte tile through which the coalesced residual loads are transposed to
  // thread-owns-row.  32 rows = one step per 32-column chunk (short-K GEMMs, where the epilogue is the critical path,
  // 3-stage operand ring); 16 rows = two half steps, which leaves room for a 4-stage ring (long-K GEMMs).
  static constexpr int kTransOffset = kStagingOffset + kStagingBytes;
  static constexpr int kTransBytes = kNumEpilogueWarps * TRANS_ROWS * 128;
  // stats[buf][src pair][half][row] = (mean, M2) of 128-row x (BLOCK_N / 2)-column pieces
  static constexpr int kStatsOffset = kTransOffset + kTransBytes;
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

template <int BLOCK_N, int STAGES, int TRANS_ROWS>
__global__ void __launch_bounds__(kNumThreads, 1)
tdc_gemm_ln_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                   const __grid_constant__ CUtensorMap map_f32, const __grid_constant__ CUtensorMap map_bf16, int m,
                   int n, int k, int cluster_size, LnArgs ln) {
  using L = LnSmem<BLOCK_N, STAGES, TRANS_ROWS>;
  constexpr uint32_t kTmemCols = kAccStages * BLOCK_N;
  constexpr uint32_t kIdesc = make_idesc_bf16_f32(2 * kBlockM, BLOCK_N);   // cta_group::2: UMMA 256 x BLOCK_N
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
  const uint32_t cluster_rank = cluster_ctarank();
  const uint32_t cta_rank = cluster_rank & 1u;     // position inside the CTA pair: rows [cta_rank * 128, +128) of the tile
  const uint32_t cta = cluster_rank >> 1;          // the pair = this CTA's N tile
  const bool is_leader = cta_rank == 0;
  const int num_pairs = cluster_size / 2;
  constexpr int kTileRows = 2 * kBlockM;           // rows of one pair's tile
  const int num_m_tiles = (m + kTileRows - 1) / kTileRows;
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
      mbar_init(&tmem_empty_bar[s], kNumEpilogueWarps * 2);   // the epilogue warps of both CTAs of the pair
      // cluster: one expect_tx arrive per use, the partials arrive as transaction bytes (st.async);
      // single CTA: plain shared-memory stores, every thread of the quadrant's two warps arrives
      for (int q = 0; q < 4; ++q) mbar_init(&stats_bar[s * 4 + q], num_pairs > 1 ? 1 : 64);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    __syncwarp();
    tmem_alloc<2>(tmem_base_smem, kTmemCols);
    tmem_relinquish<2>();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // every CTA's stats barriers exist before any peer's st.async can land
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_smem;
  // everything above (barriers, TMEM, descriptor prefetch) may overlap the previous kernel's tail
  grid_dependency_wait();
  grid_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs of the pair; bytes are credited to the leader's barrier) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int row_w = static_cast<int>(cta) * BLOCK_N + static_cast<int>(cta_rank) * L::kBRows;
      for (int tile = first_tile; tile < num_m_tiles; tile += tile_stride) {
        const int row_a = tile * kTileRows + static_cast<int>(cta_rank) * kBlockM;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
          tma_load_2d_pair(sa, &map_a, &full_bar[stage], kb * kBlockK, row_a, kL2EvictNormal);
          tma_load_2d_pair(sa + L::kABytes, &map_w, &full_bar[stage], kb * kBlockK, row_w, kL2EvictLast);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread, the pair's leader CTA) =====================
    if (lane == 0 && is_leader) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint16_t pair_mask = static_cast<uint16_t>(0x3u << (2 * cta));   // both CTAs of this pair
      for (int tile = first_tile; tile < num_m_tiles; tile += tile_stride) {
        mbar_wait<true>(&tmem_empty_bar[acc], acc_phase ^ 1);
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
            umma_f16<2>(d_tmem, desc_a + 2u * kk, desc_b + 2u * kk, kIdesc, (kb | kk) != 0 ? 1u : 0u);
          umma_commit<2>(&empty_bar[stage], pair_mask);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit<2>(&tmem_full_bar[acc], pair_mask);
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const uint32_t quad = warp & 3;
    const uint32_t half = (warp - 2) >> 2;
    uint8_t* stage_f32 = smem + L::kStagingOffset + (warp - 2) * (2 * kStoreTileBytes);  // two 4 KB staging tiles
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
    // Residual loads are COALESCED — per instruction a warp reads 4 rows x 128 bytes (lane -> row 4i + lane/8, 16-byte
    // piece lane%8) instead of 32 rows x 16 bytes, 8x fewer L1 wavefronts — and are turned into the thread-owns-row
    // layout of the TMEM accumulator through a swizzled per-warp shared-memory tile when they are consumed.
    float4 rbuf[2][8];
    uint8_t* trans = smem + L::kTransOffset + (warp - 2) * (TRANS_ROWS * 128);
    const int ld_row = static_cast<int>(lane >> 3), ld_j = static_cast<int>(lane & 7);
    auto fetch = [&](long long tile_row0, int c, float4 (&dst)[8]) {   // rows tile_row0 .. +32 of chunk c
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long r_ = tile_row0 + 4 * i + ld_row;
        dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r_ < m && col_base + c * 32 + 4 * ld_j < n && !(ln.debug & 1))
          dst[i] = __ldg(reinterpret_cast<const float4*>(ln.resid + r_ * ln.ldr + col_base + c * 32) + ld_j);
      }
    };
    auto transpose_in = [&](const float4 (&src)[8], float (&r)[32]) {   // coalesced pieces -> this lane's row
      constexpr int kSteps = 32 / TRANS_ROWS;                            // 1 (whole chunk) or 2 (halves of 16 rows)
#pragma unroll
      for (int hb = 0; hb < kSteps; ++hb) {
        __syncwarp();
#pragma unroll
        for (int i = 0; i < TRANS_ROWS / 4; ++i) {
          const int row_ = 4 * i + ld_row;                               // row inside the step
          *reinterpret_cast<float4*>(trans + row_ * 128 + ((ld_j ^ (row_ & 7)) << 4)) = src[(TRANS_ROWS / 4) * hb + i];
        }
        __syncwarp();
        if (kSteps == 1 || static_cast<int>(lane >> 4) == hb) {
          const int row_ = static_cast<int>(lane) & (TRANS_ROWS - 1);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 t = *reinterpret_cast<const float4*>(trans + row_ * 128 + ((g ^ (row_ & 7)) << 4));
            r[4 * g] = t.x; r[4 * g + 1] = t.y; r[4 * g + 2] = t.z; r[4 * g + 3] = t.w;
          }
        }
      }
    };
    auto fetch_head = [&](int t) {  // chunks 0 and 1 of tile t
      const long long r0_ = static_cast<long long>(t) * kTileRows + static_cast<int>(cta_rank) * kBlockM +
                            static_cast<int>(quad) * 32;
      const long long lim = t < num_m_tiles ? r0_ : static_cast<long long>(m);   // beyond the last tile: nothing
      fetch(lim, 0, rbuf[0]);
      if (kChunks > 1) fetch(lim, 1, rbuf[1]);
    };
    fetch_head(first_tile);
    uint32_t flip = 0;  // which of the warp's two staging tiles takes the next fp32 store
    for (int tile = first_tile; tile < num_m_tiles; tile += tile_stride) {
      const int row_in_tile = static_cast<int>(quad) * 32 + static_cast<int>(lane);
      const int row = tile * kTileRows + static_cast<int>(cta_rank) * kBlockM + row_in_tile;
      uint64_t* sbar = &stats_bar[acc * 4 + quad];
      // the partials of this (tile, quad): 2 halves x CL CTAs x 32 rows x 8 bytes, announced once per use
      if (num_pairs > 1 && half == 0 && lane == 0 && !(ln.debug & 4))
        mbar_arrive_expect_tx(sbar, static_cast<uint32_t>(2 * num_pairs * 32 * 8));
      // The residual does not depend on the MMAs.  Chunks 0 and 1 of this tile were requested before pass 2 of the
      // PREVIOUS tile (before the loop for the first one), so their DRAM latency is hidden behind that pass; chunks 2
      // and 3 follow while chunks 0 and 1 are consumed.
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
        transpose_in(rbuf[c & 1], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {   // + bias (zeros beyond n)
          const float4 b = bias_s[c * 8 + j];
          r[4 * j] += b.x; r[4 * j + 1] += b.y; r[4 * j + 2] += b.z; r[4 * j + 3] += b.w;
        }
        if (c + 2 < kChunks) fetch(row - static_cast<int>(lane), c + 2, rbuf[c & 1]);
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
        if (ln.debug & 4) {
          *slot = make_float2(mean_l, m2_l);
        } else if (num_pairs > 1) {
          // the CTAs that hold the same 128 rows: rank cta_rank of every pair
          for (int dst = 0; dst < num_pairs; ++dst)
            st_async_f32x2(slot, sbar, static_cast<uint32_t>(2 * dst) + cta_rank, mean_l, m2_l);
        } else {  // narrow rows (N <= BLOCK_N): the tile lives in one CTA, no distributed shared memory involved
          *slot = make_float2(mean_l, m2_l);
          mbar_arrive(sbar);
        }
      }
      // ---- combine the 2 * CL partials of this thread's row
      if (!(ln.debug & 4)) mbar_wait(sbar, acc_phase);
      float mean = 0.f;
      for (int j = 0; j < num_pairs; ++j)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int cnt = max(0, min(kHalfCols, n - (j * BLOCK_N + hh * kHalfCols)));
          mean += static_cast<float>(cnt) * stats[((acc * kMaxCluster + j) * 2 + hh) * kBlockM + row_in_tile].x;
        }
      mean *= inv_n;
      float m2 = 0.f;
      for (int j = 0; j < num_pairs; ++j)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int cnt = max(0, min(kHalfCols, n - (j * BLOCK_N + hh * kHalfCols)));
          const float2 p = stats[((acc * kMaxCluster + j) * 2 + hh) * kBlockM + row_in_tile];
          const float d = p.x - mean;
          m2 += p.y + static_cast<float>(cnt) * d * d;
        }
      const float rstd = 1.0f / sqrtf(m2 * inv_n + ln.eps);  // biased variance, as nn.LayerNorm

      // the next tile's first residual chunks: in flight during pass 2
      fetch_head(tile + tile_stride);

      // ---- pass 2: y = (x - mean) * rstd * gamma + beta -> fp32 + bf16 tiles -> TMA stores.
      // Two 4 KB staging tiles per warp, three stores per 64 columns (fp32 cols 0-31, fp32 cols 32-63, bf16 cols
      // 0-63) that alternate between the tiles so that a tile is only rewritten two stores after it was handed to the
      // TMA engine: every reuse waits with wait_group.read 1, i.e. for a store that has had a whole step to drain.
      const int row0 = tile * kTileRows + static_cast<int>(cta_rank) * kBlockM + static_cast<int>(quad) * 32;
      const bool live = row0 < m;  // warp-uniform
      const float rs = rstd;
#pragma unroll 1
      for (int c = 0; c < kHalfCols / 64; ++c) {
        const int col0 = col_base + c * 64;
        const bool store_ok = live && col0 < n && !(ln.debug & 2);  // warp-uniform
        uint8_t* tile_x = stage_f32 + (flip ? kStoreTileBytes : 0);
        uint8_t* tile_y = stage_f32 + (flip ? 0 : kStoreTileBytes);
        uint32_t pk[2][16];  // bf16 copy of the 64 columns, staged after both fp32 stores
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          uint32_t v[32];
          const int cc = col0 + hlf * 32;
          tmem_ld_32x32(t_addr + c * 64 + hlf * 32, v);
          tmem_ld_wait();
          float y[32];
#pragma unroll
          for (int g4 = 0; g4 < 8; ++g4) {
            const float4 gm = gamma_s[(c * 64 + hlf * 32) / 4 + g4];
            const float4 bt = beta_s[(c * 64 + hlf * 32) / 4 + g4];
            y[4 * g4 + 0] = fmaf((__uint_as_float(v[4 * g4 + 0]) - mean) * rs, gm.x, bt.x);
            y[4 * g4 + 1] = fmaf((__uint_as_float(v[4 * g4 + 1]) - mean) * rs, gm.y, bt.y);
            y[4 * g4 + 2] = fmaf((__uint_as_float(v[4 * g4 + 2]) - mean) * rs, gm.z, bt.z);
            y[4 * g4 + 3] = fmaf((__uint_as_float(v[4 * g4 + 3]) - mean) * rs, gm.w, bt.w);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[hlf][j] = pack_bf16x2(y[2 * j], y[2 * j + 1]);
          if (store_ok && cc < n) {
            uint8_t* dst = hlf == 0 ? tile_x : tile_y;
            if (lane == 0) tma_store_wait_read<1>();  // the store that last read this tile is two commits old
            __syncwarp();
            uint8_t* rowp = dst + lane * 128;
#pragma unroll
            for (int g = 0; g < 8; ++g)
              *reinterpret_cast<float4*>(rowp + ((g ^ sw) << 4)) = make_float4(y[4 * g], y[4 * g + 1], y[4 * g + 2], y[4 * g + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&map_f32, dst, cc, row0);
              tma_store_commit();
            }
          }
        }
        if (store_ok) {
          if (lane == 0) tma_store_wait_read<1>();  // tile_x's fp32 store has been read
          __syncwarp();
          uint8_t* rowb = tile_x + lane * 128;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const uint32_t* q = &pk[g >> 2][(g & 3) * 4];
            *reinterpret_cast<uint4*>(rowb + ((g ^ sw) << 4)) = make_uint4(q[0], q[1], q[2], q[3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          // (columns beyond n: TMA clips the box; their staging bytes are stale but never leave shared memory)
          if (lane == 0) {
            tma_store_2d(&map_bf16, tile_x, col0, row0);
            tma_store_commit();
          }
          flip ^= 1u;
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[acc], 2 * cta);   // the leader of this pair
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
    tmem_dealloc<2>(tmem_base, kTmemCols);
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

template <int BLOCK_N, int STAGES, int TRANS_ROWS>
int launch_ln(const GemmLnProblem& p, cudaStream_t stream, const char** err) {
  using L = LnSmem<BLOCK_N, STAGES, TRANS_ROWS>;
  const int cluster = 2 * ((p.n + BLOCK_N - 1) / BLOCK_N);   // CTA pairs, one per N tile
  CUtensorMap map_a, map_w, map_f32, map_bf16;
  if (!make_map(&map_a, p.a, p.m, p.k, p.lda, kBlockM, false) || !make_map(&map_w, p.w, p.n, p.k, p.ldw, L::kBRows, false) ||
      !make_map(&map_f32, p.out_f32, p.m, p.n, p.ldo, 32, true) || !make_map(&map_bf16, p.out_bf16, p.m, p.n, p.ldo, 32, false)) {
    if (err) *err = "gemm_ln: cuTensorMapEncodeTiled failed (pointer/pitch alignment?)";
    return TDC_ECUDA;
  }
  auto kernel = tdc_gemm_ln_kernel<BLOCK_N, STAGES, TRANS_ROWS>;
  int dev = 0;
  cudaGetDevice(&dev);
  dev = (dev >= 0 && dev < 64) ? dev : 0;
  static std::mutex mu;
  static bool attr_set[64] = {};
  static int sms[64] = {};
  static int max_clusters[64][2 * kMaxCluster + 1] = {};
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
  const int m_tiles = (p.m + 2 * kBlockM - 1) / (2 * kBlockM);
  if (m_tiles < clusters) clusters = m_tiles;
  static const bool info = [] { const char* e = getenv("TDC_GEMM_LN_INFO"); return e != nullptr && atoi(e) == 1; }();
  if (info)
    fprintf(stderr, "tdc_gemm_ln: m %d n %d k %d -> %d clusters of %d CTAs (%d M tiles), smem %d B\n", p.m, p.n, p.k,
            clusters, cluster, m_tiles, L::kTotalBytes);
  static const int debug = [] { const char* e = getenv("TDC_GEMM_LN_DEBUG"); return e ? atoi(e) : 0; }();
  LnArgs ln{debug, p.bias, p.resid, p.ldr, p.gamma, p.beta, p.eps};
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(clusters * cluster));
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = L::kTotalBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (pdl_enabled()) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
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
  if (p.n <= 128) return launch_ln<128, 4, 32>(p, stream, err);
  // short K (out-projections): the epilogue is the critical path -> one-step transposes, 3-stage ring;
  // long K (FFN down): the mainloop matters -> 4-stage ring, half-step transposes (profiles/r02_gemm_ln_fusion.txt)
  if (p.k <= 1024) return launch_ln<256, 3, 32>(p, stream, err);
  return launch_ln<256, 4, 16>(p, stream, err);
}

}  // namespace tdc
