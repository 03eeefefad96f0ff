// tcgen05 / TMEM / TMA GEMM for sm_100a:  C = epilogue(A . W^T), bf16 in, fp32 accumulate.
//
// This is the kernel behind every dense contraction on the TDC path — the
// cross-attention K/V projections of all Q-Former layers at once (reference:
// tdc/Qformer.py:129-130,186-187 — 74-87 % of the path's FLOPs), the self-attention
// QKV / output projections (Qformer.py:125,285-289), the FFN (Qformer.py:349-375),
// vision_proj / query_proj / audio_proj (tdc/cambrian_arch.py:483-484,180-181) and
// the GELU-MLP projector (cambrian_arch.py:65-69).
//
// Design (B200-first, not a translation of anything in the reference, which only
// calls cuBLAS through nn.Linear):
//   * persistent grid: one CTA (CG=1) or one CTA pair (CG=2, cta_group::2) per SM / TPC,
//     static round-robin over output tiles, N fastest so that the concurrently
//     running tiles share a handful of A row-panels while W stays L2-resident;
//   * warp-specialised: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread
//     tcgen05.mma issuer, warps 2-9 = epilogue (TMEM lane quadrant x column half each);
//   * STAGES-deep smem ring of 128B-swizzled K-major tiles (A 128x64, W BLOCK_N/CG x64)
//     with full/empty mbarriers; tcgen05.commit releases a slot as soon as the MMAs
//     that read it have retired;
//   * fp32 accumulators in TMEM, double buffered (2 x BLOCK_N columns) so the
//     epilogue of tile i overlaps the mainloop of tile i+1;
//   * epilogue fused into the drain: +bias, exact-erf GELU, bf16 / fp32 conversion; each
//     thread owns one output row, writes 128-byte row pieces into a 128B-swizzled smem
//     staging tile (bank-conflict free) and one lane hands the 32-row tile to the TMA
//     store engine — fully coalesced global writes, clipped at the M/N edges by hardware,
//     double-buffered so the next TMEM drain overlaps the store.
#include "tdc_gemm.cuh"
#include "tdc_ptx.cuh"
#include "tdc_b200.h"

#include <array>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

namespace tdc {

namespace {

constexpr int kBlockM = 128;  // rows per CTA (UMMA M = 128 * CG)
constexpr int kBlockK = 64;   // one 128-byte swizzle span of bf16
constexpr int kUmmaK = 16;
constexpr int kNumEpilogueWarps = 8;
constexpr int kNumThreads = 64 + 32 * kNumEpilogueWarps;
constexpr int kAccStages = 2;
constexpr int kStoreTileBytes = 32 * 128;  // 32 rows x 128 B, one TMA store box
constexpr int kStoreBufs = 2;              // per epilogue warp

struct EpilogueArgs {
  const float* bias;
  int mode;
  int slab_cols;  // output columns per slab (== N for a plain matrix)
  unsigned long long hint_a, hint_w, hint_c;  // L2 eviction policies of the A / W loads and the C stores
  int debug_skip;  // dev knob (TDC_GEMM_DEBUG): 1 = drain TMEM but skip the epilogue math + stores, 2 = skip the TMA
                   // stores only, 3 = skip the math / shared-memory staging only
};

template <int CG, int BLOCK_N, int STAGES>
struct SmemLayout {
  static constexpr int kBRows = BLOCK_N / CG;
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = kBRows * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingOffset = STAGES * kStageBytes;
  static constexpr int kStagingBytes = kNumEpilogueWarps * kStoreBufs * kStoreTileBytes;
  static constexpr int kBarrierOffset = kStagingOffset + kStagingBytes;
  // full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], tmem base ptr
  static constexpr int kBarrierBytes = (2 * STAGES + 2 * kAccStages) * 8 + 16;
  static constexpr int kTotalBytes = kBarrierOffset + kBarrierBytes + 1024;  // + manual 1024 B alignment slack
};

// acc (+bias) for 8 consecutive columns starting at `col` (bias may be null; columns >= n read no bias)
__device__ __forceinline__ void load_bias8(const float* bias, int col, int n, float (&b)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) b[j] = 0.f;
  if (bias != nullptr && col < n) {  // n % 8 == 0: a group of 8 is either fully inside or fully outside
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col + 4));
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
  }
}

// Write one thread's 32 accumulator columns (already in registers) into its row of the
// 128B-swizzled staging tile: 16-byte chunk j of row r lives at r*128 + ((j ^ (r & 7)) * 16).
// bf16 modes: the 32 columns are 64 B = chunks [chunk0, chunk0+4); fp32: 128 B = chunks [0, 8).
// bias of 32 consecutive columns into registers (issued before the TMEM load is waited for, so the global
// load latency hides behind it)
__device__ __forceinline__ void load_bias32(const float* bias, int col0, int n, float (&b)[32]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float t[8];
    load_bias8(bias, col0 + g * 8, n, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) b[g * 8 + j] = t[j];
  }
}

template <int MODE>
__device__ __forceinline__ void stage_32_columns(const uint32_t (&v)[32], uint8_t* tile, uint32_t lane, int chunk0,
                                                 const float (&bias)[32]) {
  uint8_t* row = tile + lane * 128;
  const uint32_t sw = lane & 7;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[g * 8 + j]) + bias[g * 8 + j];
    if (MODE == EPI_BIAS_GELU_BF16) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) gelu_erf_fast_x2(x[j], x[j + 1]);
    }
    if (MODE == EPI_BIAS_F32) {
      *reinterpret_cast<float4*>(row + (((2 * g) ^ sw) << 4)) = make_float4(x[0], x[1], x[2], x[3]);
      *reinterpret_cast<float4*>(row + (((2 * g + 1) ^ sw) << 4)) = make_float4(x[4], x[5], x[6], x[7]);
    } else {
      uint4 pk;
      pk.x = pack_bf16x2(x[0], x[1]);
      pk.y = pack_bf16x2(x[2], x[3]);
      pk.z = pack_bf16x2(x[4], x[5]);
      pk.w = pack_bf16x2(x[6], x[7]);
      *reinterpret_cast<uint4*>(row + (((chunk0 + g) ^ sw) << 4)) = pk;
    }
  }
}

// Tile rasterisation.  N tiles are walked in groups of `n_group` columns: within a group the
// order is N-fastest (the ~74-148 concurrently running tiles share a few A row-panels), and a
// whole group's W slice (n_group * BLOCK_N * K * 2 B, sized by the host to ~1/4 of L2) stays
// L2-resident while ALL M tiles stream past it, so A is read from HBM once per group and W
// once per kernel instead of W being re-fetched every wave (measured: 24 GB of DRAM reads for
// 2.7 GB of operands before this).
struct TileCoord { int tm, tn; };
__device__ __forceinline__ TileCoord decode_tile(long long tile, int num_m_tiles, int num_n_tiles, int n_group) {
  const long long per_group = static_cast<long long>(num_m_tiles) * n_group;
  const int g = static_cast<int>(tile / per_group);
  const int rem = static_cast<int>(tile - g * per_group);
  const int n0 = g * n_group;
  const int width = (num_n_tiles - n0) < n_group ? (num_n_tiles - n0) : n_group;
  return TileCoord{rem / width, n0 + rem % width};
}

// MC = CTA pairs per cluster (CG == 2 only).  MC == 2: a 4-CTA cluster works on two vertically adjacent
// 256-row tiles of the same N tile; the W tile they share is fetched once per cluster — every CTA loads a
// quarter of it and TMA-multicasts it to the CTA holding the same W half in the other pair — which cuts the
// L2 -> SM operand traffic per FLOP by 25 %.  The smem ring is then shared state of the cluster: a slot is
// free when BOTH pairs' MMAs have retired it (empty barriers count MC commits, multicast to all CTAs).
template <int CG, int BLOCK_N, int STAGES, int MC>
__global__ void __launch_bounds__(kNumThreads, 1)
tdc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                const __grid_constant__ CUtensorMap map_c, int m, int n, int k, int n_group, EpilogueArgs epi) {
  using L = SmemLayout<CG, BLOCK_N, STAGES>;
  constexpr uint32_t kTmemCols = kAccStages * BLOCK_N;  // 512 (BLOCK_N=256) or 256
  constexpr uint32_t kIdesc = make_idesc_bf16_f32(kBlockM * CG, BLOCK_N);

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment; dynamic smem only promises 16.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarrierOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + kAccStages;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + kAccStages);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  static_assert(MC == 1 || CG == 2, "multi-pair clusters need cta_group::2");
  const uint32_t cluster_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const uint32_t cta_rank = cluster_rank & 1u;   // position inside the CTA pair
  const uint32_t pair = cluster_rank >> 1;       // pair inside the cluster (0 when MC == 1)
  const bool is_leader = (cta_rank == 0);

  const int tile_m_rows = kBlockM * CG;          // rows of one pair's tile
  const int num_m_tiles = ((m + tile_m_rows - 1) / tile_m_rows + MC - 1) / MC;  // in units of MC stacked tiles
  const int num_n_tiles = (n + BLOCK_N - 1) / BLOCK_N;
  const long long num_tiles = static_cast<long long>(num_m_tiles) * num_n_tiles;
  const int num_kb = (k + kBlockK - 1) / kBlockK;
  const long long first_tile = blockIdx.x / (CG * MC);
  const long long tile_stride = gridDim.x / (CG * MC);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_c);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], MC);  // one tcgen05.commit per pair of the cluster
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], kNumEpilogueWarps * CG);  // one lane per epilogue warp (of both CTAs)
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    __syncwarp();
    tmem_alloc<CG>(tmem_base_smem, kTmemCols);
    tmem_relinquish<CG>();
  }
  tc_fence_before_sync();
  __syncthreads();                   // CTA-level: barrier inits + the TMEM address written by tcgen05.alloc
  if (CG == 2) cluster_sync_all();   // pair-level: the peer's barriers are initialised before any remote arrive
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_smem;
  // everything above (barriers, TMEM, descriptor prefetch) may overlap the previous kernel's tail
  grid_dependency_wait();
  grid_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = first_tile; tile < num_tiles; tile += tile_stride) {
        const TileCoord tc = decode_tile(tile, num_m_tiles, num_n_tiles, n_group);
        const int row_a = (tc.tm * MC + static_cast<int>(pair)) * tile_m_rows + static_cast<int>(cta_rank) * kBlockM;
        const int row_w = tc.tn * BLOCK_N + static_cast<int>(cta_rank) * L::kBRows;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          if (CG == 1) {
            mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
            tma_load_2d(sa, &map_a, &full_bar[stage], kb * kBlockK, row_a, epi.hint_a);
            tma_load_2d(sb, &map_w, &full_bar[stage], kb * kBlockK, row_w, epi.hint_w);
          } else {
            // both CTAs' bytes are credited to the leader's barrier
            if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
            tma_load_2d_pair(sa, &map_a, &full_bar[stage], kb * kBlockK, row_a, epi.hint_a);
            if (MC == 1) {
              tma_load_2d_pair(sb, &map_w, &full_bar[stage], kb * kBlockK, row_w, epi.hint_w);  // W: keep in L2
            } else {
              // this CTA fetches 1/MC of its W half and multicasts it to the same-half CTA of every pair
              constexpr int kPiece = L::kBRows / MC;
              const uint16_t mask = static_cast<uint16_t>(0x5u << cta_rank);  // ranks {r, r + 2}
              tma_load_2d_pair_multicast(sb + pair * (kPiece * kBlockK * 2), &map_w, &full_bar[stage], kb * kBlockK,
                                         row_w + static_cast<int>(pair) * kPiece, mask, epi.hint_w);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread, leader CTA only) =====================
    if (lane == 0 && is_leader) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (long long tile = first_tile; tile < num_tiles; tile += tile_stride) {
        mbar_wait<CG == 2>(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint64_t desc_a = make_kmajor_sw128_desc(sa);
          const uint64_t desc_b = make_kmajor_sw128_desc(sa + L::kABytes);
#pragma unroll
          for (int kk = 0; kk < kBlockK / kUmmaK; ++kk) {
            // advance 16 bf16 = 32 B along K inside the swizzle span: +2 in the (addr >> 4) field
            umma_f16<CG>(d_tmem, desc_a + 2u * kk, desc_b + 2u * kk, kIdesc, (kb | kk) != 0 ? 1u : 0u);
          }
          // slot reusable once these MMAs retire (every CTA of the cluster is told)
          umma_commit<CG>(&empty_bar[stage], static_cast<uint16_t>((1u << (CG * MC)) - 1u));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit<CG>(&tmem_full_bar[acc], static_cast<uint16_t>(0x3u << (2 * pair)));  // -> this pair's epilogues
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    // warp -> (TMEM lane quadrant = warp % 4 [hardware rule], column half of the tile)
    const uint32_t quad = warp & 3;
    const uint32_t half = (warp - 2) >> 2;
    constexpr int kHalfCols = BLOCK_N / 2;
    uint8_t* staging = smem + L::kStagingOffset + (warp - 2) * (kStoreBufs * kStoreTileBytes);
    const bool f32_out = (epi.mode == EPI_BIAS_F32);
    int sbuf = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long tile = first_tile; tile < num_tiles; tile += tile_stride) {
      const TileCoord tc = decode_tile(tile, num_m_tiles, num_n_tiles, n_group);
      const int row0 = (tc.tm * MC + static_cast<int>(pair)) * tile_m_rows + static_cast<int>(cta_rank) * kBlockM +
                       static_cast<int>(quad) * 32;
      const int col_base = tc.tn * BLOCK_N + static_cast<int>(half) * kHalfCols;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t t_addr = tmem_base + ((quad * 32u) << 16) + static_cast<uint32_t>(acc * BLOCK_N) + half * kHalfCols;
      const bool live = row0 < m;  // warp-uniform: this 32-row slab has at least one real row
      if (f32_out) {
        // 32 fp32 columns = one 128-byte staging row per store
#pragma unroll 1
        for (int c = 0; c < kHalfCols / 32; ++c) {
          uint32_t v[32];
          float bias[32];
          const int col0 = col_base + c * 32;
          tmem_ld_32x32(t_addr + c * 32, v);
          load_bias32(epi.bias, col0, n, bias);
          tmem_ld_wait();
          if (live && col0 < n && epi.debug_skip != 1) {
            if (lane == 0) tma_store_wait_read<kStoreBufs - 1>();  // staging[sbuf] no longer being read
            __syncwarp();
            uint8_t* tile_buf = staging + sbuf * kStoreTileBytes;
            if (epi.debug_skip != 3) stage_32_columns<EPI_BIAS_F32>(v, tile_buf, lane, 0, bias);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && epi.debug_skip != 2) {
              tma_store_3d(&map_c, tile_buf, col0 % epi.slab_cols, row0, col0 / epi.slab_cols, epi.hint_c);
              tma_store_commit();
            }
            sbuf ^= 1;
          }
        }
      } else {
        // 64 bf16 columns (two TMEM loads) = one 128-byte staging row per store
#pragma unroll 1
        for (int c = 0; c < kHalfCols / 64; ++c) {
          uint32_t v0[32], v1[32];
          float bias0[32], bias1[32];
          const int col0 = col_base + c * 64;
          tmem_ld_32x32(t_addr + c * 64, v0);
          tmem_ld_32x32(t_addr + c * 64 + 32, v1);
          load_bias32(epi.bias, col0, n, bias0);
          load_bias32(epi.bias, col0 + 32, n, bias1);
          tmem_ld_wait();
          if (live && col0 < n && epi.debug_skip != 1) {
            if (lane == 0) tma_store_wait_read<kStoreBufs - 1>();
            __syncwarp();
            uint8_t* tile_buf = staging + sbuf * kStoreTileBytes;
            if (epi.debug_skip == 3) {
              // dev knock-out: stores only (stale staging contents)
            } else if (epi.mode == EPI_BIAS_GELU_BF16) {
              stage_32_columns<EPI_BIAS_GELU_BF16>(v0, tile_buf, lane, 0, bias0);
              stage_32_columns<EPI_BIAS_GELU_BF16>(v1, tile_buf, lane, 4, bias1);
            } else {
              stage_32_columns<EPI_BIAS_BF16>(v0, tile_buf, lane, 0, bias0);
              stage_32_columns<EPI_BIAS_BF16>(v1, tile_buf, lane, 4, bias1);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && epi.debug_skip != 2) {
              tma_store_3d(&map_c, tile_buf, col0 % epi.slab_cols, row0, col0 / epi.slab_cols, epi.hint_c);
              tma_store_commit();
            }
            sbuf ^= 1;
          }
        }
      }
      // all TMEM reads of this accumulator are complete (wait::ld above): hand it back to the MMA warp
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (CG == 1) mbar_arrive(&tmem_empty_bar[acc]);
        else mbar_arrive_cluster(&tmem_empty_bar[acc], pair * 2);  // leader of this pair
      }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all<0>();  // smem must outlive the last bulk stores
  }

  // ===================== teardown =====================
  tc_fence_before_sync();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after_sync();
    tmem_dealloc<CG>(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
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

// 2-D row-major [rows, cols] (pitch ld elements, bf16 or fp32) -> boxes of box_rows x 128 bytes with
// 128B swizzle (64 bf16 / 32 fp32 columns per box row).
// cuTensorMapEncodeTiled costs ~1 us on the host and a forward pass issues ~190 of them per row batch with
// a handful of distinct (pointer, shape) combinations (the workspace is reused call after call), so encoded
// descriptors are memoised by their full argument list.  A descriptor is a pure function of those arguments,
// hence a hit is always valid even if the allocation behind the pointer changed owner in between.
using MapKey = std::array<long long, 8>;
std::mutex g_map_mutex;
std::map<MapKey, CUtensorMap> g_map_cache;

template <typename Encode>
bool cached_map(CUtensorMap* map, const MapKey& key, Encode&& encode) {
  {
    std::lock_guard<std::mutex> lock(g_map_mutex);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) { *map = it->second; return true; }
  }
  if (!encode(map)) return false;
  std::lock_guard<std::mutex> lock(g_map_mutex);
  if (g_map_cache.size() >= 4096) g_map_cache.clear();
  g_map_cache.emplace(key, *map);
  return true;
}

bool encode_tensor_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows,
                       bool f32) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return false;
  const int esz = f32 ? 4 : 2;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * esz};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esz), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                        const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Output map: [slabs][rows][slab_cols] (a plain matrix is one slab), boxes of 32 rows x 128 bytes.
bool make_tensor_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows,
                     bool f32 = false) {
  const MapKey key{0, reinterpret_cast<long long>(base), rows, cols, ld, box_rows, f32 ? 1 : 0, 0};
  return cached_map(map, key, [&](CUtensorMap* m) { return encode_tensor_map(m, base, rows, cols, ld, box_rows, f32); });
}

bool encode_output_map(CUtensorMap* map, const void* base, long long rows, long long slab_cols, long long ld,
                       long long slabs, long long slab_stride, bool f32) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return false;
  const int esz = f32 ? 4 : 2;
  const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(slab_cols), static_cast<cuuint64_t>(rows),
                              static_cast<cuuint64_t>(slabs)};
  const cuuint64_t gstride[2] = {static_cast<cuuint64_t>(ld) * esz,
                                 static_cast<cuuint64_t>(slabs > 1 ? slab_stride : rows * ld) * esz};
  const cuuint32_t box[3] = {static_cast<cuuint32_t>(128 / esz), 32u, 1u};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                        const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Per-device launch state.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize), the SM count and the number of
// co-resident clusters are properties of ONE device: a process that drives several GPUs (one engine per device)
// needs them per device ordinal, not per process.
constexpr int kMaxDevices = 64;
struct DeviceState {
  bool attr_set = false;
  long long max_clusters = 0;
};
std::mutex g_device_mutex;

bool make_output_map(CUtensorMap* map, const void* base, long long rows, long long slab_cols, long long ld,
                     long long slabs, long long slab_stride, bool f32) {
  const MapKey key{1, reinterpret_cast<long long>(base), rows, slab_cols, ld, slabs, slab_stride, f32 ? 1 : 0};
  return cached_map(map, key, [&](CUtensorMap* m) {
    return encode_output_map(m, base, rows, slab_cols, ld, slabs, slab_stride, f32);
  });
}

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

int num_sms(int dev) {
  static int sms[kMaxDevices] = {};
  std::lock_guard<std::mutex> lock(g_device_mutex);
  if (sms[dev] == 0) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
  return sms[dev];
}

template <int CG, int BLOCK_N, int STAGES, int MC = 1>
int launch_variant(const GemmProblem& p, cudaStream_t stream, const char** err) {
  using L = SmemLayout<CG, BLOCK_N, STAGES>;
  CUtensorMap map_a, map_w, map_c;
  if (!make_tensor_map(&map_a, p.a, p.m, p.k, p.lda, kBlockM) ||
      !make_tensor_map(&map_w, p.w, p.n, p.k, p.ldw, L::kBRows / MC) ||
      !make_output_map(&map_c, p.out, p.m, p.slab_cols > 0 ? p.slab_cols : p.n, p.ldo,
                       p.slab_cols > 0 ? (p.n + p.slab_cols - 1) / p.slab_cols : 1, p.slab_stride,
                       p.mode == EPI_BIAS_F32)) {
    if (err) *err = "cuTensorMapEncodeTiled failed (pointer/pitch alignment?)";
    return TDC_ECUDA;
  }
  auto kernel = tdc_gemm_kernel<CG, BLOCK_N, STAGES, MC>;
  const int dev = current_device();
  const int sms = num_sms(dev);
  static DeviceState state[kMaxDevices];  // per template instantiation AND per device
  long long max_clusters;
  {
    std::lock_guard<std::mutex> lock(g_device_mutex);
    DeviceState& st = state[dev];
    if (!st.attr_set) {
      if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotalBytes) != cudaSuccess) {
        if (err) *err = "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed";
        return TDC_ECUDA;
      }
      st.attr_set = true;
    }
    // persistent grid = as many clusters as can be co-resident (clusters of 4 do not tile the 148 SMs
    // perfectly: GPCs have 16-20 SMs), otherwise the static tile striding would serialise the leftovers
    if (st.max_clusters == 0) {
      st.max_clusters = sms / (CG * MC);
      if (CG * MC > 2) {
        cudaLaunchConfig_t q{};
        q.gridDim = dim3(static_cast<unsigned>(sms / (CG * MC) * CG * MC));
        q.blockDim = dim3(kNumThreads);
        q.dynamicSmemBytes = L::kTotalBytes;
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = CG * MC; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
        q.attrs = qa; q.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kernel, &q) == cudaSuccess && n > 0 && n < st.max_clusters)
          st.max_clusters = n;
      }
    }
    max_clusters = st.max_clusters;
  }
  const long long m_tiles = (p.m + kBlockM * CG - 1) / (kBlockM * CG);
  const long long tiles = ((m_tiles + MC - 1) / MC) * ((p.n + BLOCK_N - 1) / BLOCK_N);
  long long clusters = max_clusters;
  if (tiles < clusters) clusters = tiles;
  // L2 policy: W is re-read by every M tile -> evict_last; A and C stream.  (dev knob TDC_GEMM_HINTS=awc, one
  // letter each for the A loads, W loads and C stores: n|f|l = normal / evict_first / evict_last)
  static const char* hints_env = getenv("TDC_GEMM_HINTS");
  auto pick = [](char ch, unsigned long long dflt) -> unsigned long long {
    return ch == 'n' ? kL2EvictNormal : ch == 'f' ? kL2EvictFirst : ch == 'l' ? kL2EvictLast : dflt;
  };
  // dev knock-outs (wrong results by design): 1 = drain TMEM only, 2 = no TMA stores, 3 = stores without the math
  static const int debug_skip_epilogue = [] { const char* e = getenv("TDC_GEMM_DEBUG"); const int v = e ? atoi(e) : 0; return (v >= 1 && v <= 3) ? v : 0; }();
  const bool he = hints_env != nullptr && strlen(hints_env) == 3;
  EpilogueArgs e{p.bias, p.mode, p.slab_cols > 0 ? p.slab_cols : p.n,
                 pick(he ? hints_env[0] : 0, kL2EvictNormal), pick(he ? hints_env[1] : 0, kL2EvictLast),
                 pick(he ? hints_env[2] : 0, kL2EvictNormal),
                 debug_skip_epilogue};
  // N-tile group whose W slice fits a quarter of the 126 MB L2 (the L2 is two ~63 MB halves and
  // read-shared lines end up in both), balanced over the groups.
  const int num_n_tiles = (p.n + BLOCK_N - 1) / BLOCK_N;
  const long long w_tile_bytes = static_cast<long long>(BLOCK_N) * p.k * 2;
  static const long long group_budget = [] {  // dev knob: TDC_GEMM_NGROUP_MB overrides the 36 MB W-slice budget
    const char* e = getenv("TDC_GEMM_NGROUP_MB");
    return (e != nullptr && atoi(e) > 0) ? (static_cast<long long>(atoi(e)) << 20) : (36ll << 20);
  }();
  int n_group = static_cast<int>(group_budget / (w_tile_bytes > 0 ? w_tile_bytes : 1));
  if (n_group < 1) n_group = 1;
  if (n_group > num_n_tiles) n_group = num_n_tiles;
  const int groups = (num_n_tiles + n_group - 1) / n_group;
  n_group = (num_n_tiles + groups - 1) / groups;
  // Wave alignment (dev knob TDC_GEMM_ALIGN_WAVES=1): with N-fastest order inside a group, the `clusters` tiles in
  // flight cover clusters / n_group A row-panels; when that is not an integer a panel's tiles are split over two
  // waves and the panel is fetched from DRAM twice (the K/V output streaming through L2 evicts it in between).
  // Rounding the persistent grid down to a multiple of n_group (74 -> 72 pairs for 18-tile groups) keeps every
  // panel inside one wave at the price of 2 idle pairs.
  static const bool align_waves = [] { const char* e = getenv("TDC_GEMM_ALIGN_WAVES"); return e != nullptr && atoi(e) == 1; }();
  if (align_waves && groups > 1 && n_group >= 8 && clusters > n_group && tiles > 4 * clusters)
    clusters = clusters / n_group * n_group;

  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(clusters * CG * MC));
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = L::kTotalBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG * MC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (pdl_enabled()) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  const cudaError_t rc = cudaLaunchKernelEx(&cfg, kernel, map_a, map_w, map_c, p.m, p.n, p.k, n_group, e);
  if (rc != cudaSuccess) {
    if (err) *err = cudaGetErrorString(rc);
    return TDC_ECUDA;
  }
  return TDC_OK;
}

}  // namespace

int gemm_launch(const GemmProblem& p, cudaStream_t stream, const char** err) {
  if (p.m <= 0 || p.n <= 0 || p.k <= 0) {
    if (err) *err = "gemm: empty problem";
    return TDC_EINVAL;
  }
  if ((p.k % 8) != 0 || (p.n % 8) != 0 || (p.lda % 8) != 0 || (p.ldw % 8) != 0 || (p.ldo % 4) != 0 ||
      (reinterpret_cast<uintptr_t>(p.a) & 15) || (reinterpret_cast<uintptr_t>(p.w) & 15) ||
      (reinterpret_cast<uintptr_t>(p.out) & 15)) {
    if (err) *err = "gemm: K, N and pitches must be multiples of 8 elements and pointers 16-byte aligned";
    return TDC_EINVAL;
  }
  if (p.mode != EPI_BIAS_F32 && (p.ldo % 8) != 0) {
    if (err) *err = "gemm: bf16 output pitch must be a multiple of 8 elements";
    return TDC_EINVAL;
  }
  // a store box is 128 bytes wide: 64 bf16 or 32 fp32 columns; slabs must be whole boxes
  const int box_cols = p.mode == EPI_BIAS_F32 ? 32 : 64;
  if (p.slab_cols < 0 || (p.slab_cols > 0 && ((p.slab_cols % box_cols) != 0 || (p.n % p.slab_cols) != 0 ||
                                              (p.slab_stride % 8) != 0))) {
    if (err) *err = "gemm: slab_cols must be a multiple of the 128-byte store box dividing N, slab_stride a multiple of 8";
    return TDC_EINVAL;
  }
  if (p.mode < 0 || p.mode > EPI_BIAS_F32) {
    if (err) *err = "gemm: unknown epilogue mode";
    return TDC_EINVAL;
  }
  int cg = p.cta_group == 0 ? 1 : p.cta_group;
  if (p.n <= 128) {
    // narrow outputs (small test geometries): single-CTA 128x128 tiles
    return launch_variant<1, 128, 4>(p, stream, err);
  }
  if (cg == 2) {
    // dev knob TDC_GEMM_MC=1|2: CTA pairs per cluster (2 = W multicast across two stacked tiles)
    static const int mc = [] { const char* e = getenv("TDC_GEMM_MC"); return (e && atoi(e) == 2) ? 2 : 1; }();
    if (mc == 2 && p.m > 2 * kBlockM * 2) return launch_variant<2, 256, 5, 2>(p, stream, err);
    return launch_variant<2, 256, 5>(p, stream, err);
  }
  return launch_variant<1, 256, 3>(p, stream, err);
}

}  // namespace tdc
