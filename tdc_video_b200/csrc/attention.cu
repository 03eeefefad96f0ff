// Short-query attention for the Q-Former: K (<= a few dozen) queries per row against
// n key/value tokens, head size 64, no mask other than an optional per-row KV length.
//
// Reference semantics (tdc/Qformer.py:205-268): S = Q K^T / sqrt(64); P = softmax(S, -1);
// ctx = P V; heads merged back to [tokens, hidden].  The additive masks of the reference
// are identically zero on the TDC path (all-ones attention masks, cambrian_arch.py:1648-1650
// and Qformer.py:879-882), so they are not materialised.
//
// This op is bandwidth-bound (16 queries re-use each K/V byte only 16x), so the design is
// about reading K/V exactly once at full sector efficiency, not about tensor throughput:
//   * one warp per (row, 16-query block, head); no block-level sync;
//   * K/V fragments are fetched directly in mma.sync operand order.  The
//     contraction index of an MMA may be permuted freely as long as both operands agree,
//     so each thread fetches 16-byte chunks (a full 32 B sector per thread pair) and the
//     permutation is absorbed into which 16 B of Q the thread holds (for S = Q K^T) and
//     into which 16 output columns it owns (for O = P V);
//   * online softmax in fp32 (exp2 with the 1/sqrt(d)*log2(e) scale folded in), P rounded
//     to bf16 only as the MMA operand, O accumulated in fp32;
//   * K/V stream through a per-warp cp.async ring in shared memory (3 stages x 4 KB, every thread
//     writes and later reads only its own 16-byte slots, so no barrier is needed): two 16-token
//     groups are in flight per warp at zero register cost, 16 warps/SM -> 128 KB in flight per SM.
// mma.sync m16n8k16 (legacy HMMA path) is deliberate: a 16-row tile is 1/8 of the
// smallest tcgen05 tile and the kernel sits on the HBM roofline, not the tensor roofline.
#include "tdc_kernels.cuh"
#include "tdc_ptx.cuh"

#include <mutex>

namespace tdc {

namespace {

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ long long seg_row(int r, int i, int seg1, int seg2, long long base1, long long base2) {
  return (i < seg1) ? base1 + static_cast<long long>(r) * seg1 + i
                    : base2 + static_cast<long long>(r) * seg2 + (i - seg1);
}

constexpr int kStages = 3;            // cp.async ring depth per warp
constexpr int kSlotsPerStage = 8 * 32;  // 8 x 16-byte chunks per lane

__device__ __forceinline__ void cp_async_16(uint4* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct KVRegs {
  uint4 k[4];  // [tile 0: dh c*8.., dh 32+c*8..] [tile 1: same]
  uint4 v[4];  // tokens t0+2c, t0+2c+1, t0+8+2c, t0+8+2c+1 at dh g*8..g*8+7
};

// kTwoSeg = false: all KV (and Q) tokens of a row are contiguous (cross-attention, text-less
// self-attention) so the per-token slab-row lookup collapses to base + token.
template <bool kTwoSeg>
__global__ void __launch_bounds__(128, 4) tdc_attention_kernel(AttentionArgs a) {
  grid_dependency_wait();
  grid_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2;  // MMA "group" id: fragment row / column owner
  const int c = lane & 3;   // thread within group
  const int nqb = (a.nq + 15) >> 4;
  const long long total = static_cast<long long>(a.rows) * nqb * a.heads;
  const long long wid = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (wid >= total) return;
  const int h = static_cast<int>(wid % a.heads);
  const int qb = static_cast<int>((wid / a.heads) % nqb);
  const int r = static_cast<int>(wid / (static_cast<long long>(a.heads) * nqb));

  const int kv_total = a.kv_seg1 + a.kv_seg2 + a.kv_seg3;
  const uint32_t mbits = a.kv_mask != nullptr ? a.kv_mask[r] : 0xffffffffu;
  int kvn = kv_total;
  if (a.kv_len != nullptr) {
    kvn = a.kv_len[r];
    kvn = kvn < 1 ? 1 : (kvn > kv_total ? kv_total : kvn);
  }

  // ---- Q fragments: rows g and g+8 of the block, dh chunks {c*8..+7} and {32+c*8..+7}
  const int qi0 = qb * 16 + g, qi1 = qi0 + 8;
  const int qc0 = qi0 < a.nq ? qi0 : a.nq - 1, qc1 = qi1 < a.nq ? qi1 : a.nq - 1;
  const long long qrow0 = kTwoSeg ? seg_row(r, qc0, a.q_seg1, a.q_seg2, a.q_base1, a.q_base2)
                                  : a.q_base1 + static_cast<long long>(r) * a.q_seg1 + qc0;
  const long long qrow1 = kTwoSeg ? seg_row(r, qc1, a.q_seg1, a.q_seg2, a.q_base1, a.q_base2)
                                  : a.q_base1 + static_cast<long long>(r) * a.q_seg1 + qc1;
  uint32_t qa[8], qb_[8];
  {
    const __nv_bfloat16* p0 = a.q + qrow0 * a.ldq + h * 64 + c * 8;
    const __nv_bfloat16* p1 = a.q + qrow1 * a.ldq + h * 64 + c * 8;
    const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(p0));
    const uint4 x1 = __ldg(reinterpret_cast<const uint4*>(p0 + 32));
    const uint4 y0 = __ldg(reinterpret_cast<const uint4*>(p1));
    const uint4 y1 = __ldg(reinterpret_cast<const uint4*>(p1 + 32));
    qa[0] = x0.x; qa[1] = x0.y; qa[2] = x0.z; qa[3] = x0.w; qa[4] = x1.x; qa[5] = x1.y; qa[6] = x1.z; qa[7] = x1.w;
    qb_[0] = y0.x; qb_[1] = y0.y; qb_[2] = y0.z; qb_[3] = y0.w; qb_[4] = y1.x; qb_[5] = y1.y; qb_[6] = y1.z; qb_[7] = y1.w;
  }

  const __nv_bfloat16* kbase = a.k + h * a.k_head_stride + c * 8;
  const __nv_bfloat16* vbase = a.v + h * a.v_head_stride + g * 8;
  if (!kTwoSeg) {  // fold the row's first KV token into the base pointers
    const long long first = a.kv_base1 + static_cast<long long>(r) * a.kv_seg1;
    kbase += first * a.ldk;
    vbase += first * a.ldv;
  }
  auto kv_row = [&](int t) -> long long {
    if (!kTwoSeg) return static_cast<long long>(t);
    if (t >= a.kv_seg1 + a.kv_seg2) return a.kv_base3 + (t - a.kv_seg1 - a.kv_seg2);  // shared by all rows
    return seg_row(r, t, a.kv_seg1, a.kv_seg2, a.kv_base1, a.kv_base2);
  };
  extern __shared__ uint4 kv_ring[];  // [4 warps][kStages][8 chunks][32 lanes]
  uint4* ring = kv_ring + (threadIdx.x >> 5) * (kStages * kSlotsPerStage) + lane;
  // queue the 8 x 16 B this thread needs from token group t0 into ring stage `stage`
  auto issue_group = [&](int t0, int stage) {
    // clamp to the last valid token: out-of-range columns are masked to P = 0 below
    int tk0 = t0 + g, tk1 = t0 + 8 + g;
    tk0 = tk0 < kvn ? tk0 : kvn - 1;
    tk1 = tk1 < kvn ? tk1 : kvn - 1;
    const __nv_bfloat16* pk0 = kbase + kv_row(tk0) * a.ldk;
    const __nv_bfloat16* pk1 = kbase + kv_row(tk1) * a.ldk;
    uint4* dst = ring + stage * kSlotsPerStage;
    cp_async_16(dst + 0 * 32, pk0);
    cp_async_16(dst + 1 * 32, pk0 + 32);
    cp_async_16(dst + 2 * 32, pk1);
    cp_async_16(dst + 3 * 32, pk1 + 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int tv = t0 + (j >> 1) * 8 + c * 2 + (j & 1);
      tv = tv < kvn ? tv : kvn - 1;
      cp_async_16(dst + (4 + j) * 32, vbase + kv_row(tv) * a.ldv);
    }
  };

  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  const int ngroups = (kvn + 15) >> 4;
#pragma unroll
  for (int st = 0; st < kStages - 1; ++st) {  // prologue: fill all but one stage
    if (st < ngroups) issue_group(st * 16, st);
    cp_async_commit();                        // (empty groups keep the wait count uniform)
  }
  for (int gi = 0; gi < ngroups; ++gi) {
    const int t0 = gi * 16;
    {
      const int gn = gi + kStages - 1;        // refill the stage consumed in the previous iteration
      if (gn < ngroups) issue_group(gn * 16, gn % kStages);
      cp_async_commit();
    }
    cp_async_wait<kStages - 1>();             // group gi has landed (this thread's own slots only)
    KVRegs cur;
    {
      const uint4* src = ring + (gi % kStages) * kSlotsPerStage;
#pragma unroll
      for (int j = 0; j < 4; ++j) { cur.k[j] = src[j * 32]; cur.v[j] = src[(4 + j) * 32]; }
    }

    // ---- S = Q K^T for 16 tokens: two n-tiles of 8, four k-steps of 16 over dh
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
    {
      const uint32_t k0[8] = {cur.k[0].x, cur.k[0].y, cur.k[0].z, cur.k[0].w,
                              cur.k[1].x, cur.k[1].y, cur.k[1].z, cur.k[1].w};
      const uint32_t k1[8] = {cur.k[2].x, cur.k[2].y, cur.k[2].z, cur.k[2].w,
                              cur.k[3].x, cur.k[3].y, cur.k[3].z, cur.k[3].w};
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        mma_bf16_16816(s0, qa[2 * s], qb_[2 * s], qa[2 * s + 1], qb_[2 * s + 1], k0[2 * s], k0[2 * s + 1]);
        mma_bf16_16816(s1, qa[2 * s], qb_[2 * s], qa[2 * s + 1], qb_[2 * s + 1], k1[2 * s], k1[2 * s + 1]);
      }
    }
    // ---- scale, mask the tail, online softmax (rows g and g+8)
    const int tc = t0 + c * 2;
    const bool v00 = tc < kvn && ((mbits >> (tc & 31)) & 1u), v01 = tc + 1 < kvn && ((mbits >> ((tc + 1) & 31)) & 1u);
    const bool v10 = tc + 8 < kvn && ((mbits >> ((tc + 8) & 31)) & 1u);
    const bool v11 = tc + 9 < kvn && ((mbits >> ((tc + 9) & 31)) & 1u);
    s0[0] = v00 ? s0[0] * a.scale_log2 : -INFINITY;
    s0[1] = v01 ? s0[1] * a.scale_log2 : -INFINITY;
    s0[2] = v00 ? s0[2] * a.scale_log2 : -INFINITY;
    s0[3] = v01 ? s0[3] * a.scale_log2 : -INFINITY;
    s1[0] = v10 ? s1[0] * a.scale_log2 : -INFINITY;
    s1[1] = v11 ? s1[1] * a.scale_log2 : -INFINITY;
    s1[2] = v10 ? s1[2] * a.scale_log2 : -INFINITY;
    s1[3] = v11 ? s1[3] * a.scale_log2 : -INFINITY;
    float mx0 = fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1]));
    float mx1 = fmaxf(fmaxf(s0[2], s0[3]), fmaxf(s1[2], s1[3]));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mp0 = m0, mp1 = m1;  // running maxima before this group
    m0 = fmaxf(m0, mx0);
    m1 = fmaxf(m1, mx1);
    // safe subtrahend: with a kv_mask a whole group (or everything so far) can be masked; exp2(-inf - (-inf))
    // must not enter the arithmetic — such rows simply keep P = 0, O = 0, l = 0 until a valid token arrives
    const float mn0 = (m0 == -INFINITY) ? 0.f : m0;
    const float mn1 = (m1 == -INFINITY) ? 0.f : m1;
    const float al0 = exp2f(mp0 - mn0), al1 = exp2f(mp1 - mn1);
    const float p00 = exp2f(s0[0] - mn0), p01 = exp2f(s0[1] - mn0), p02 = exp2f(s1[0] - mn0), p03 = exp2f(s1[1] - mn0);
    const float p10 = exp2f(s0[2] - mn1), p11 = exp2f(s0[3] - mn1), p12 = exp2f(s1[2] - mn1), p13 = exp2f(s1[3] - mn1);
    l0 = l0 * al0 + (p00 + p01 + p02 + p03);
    l1 = l1 * al1 + (p10 + p11 + p12 + p13);
    const uint32_t pa0 = pack_bf16x2(p00, p01);  // row g,   tokens t0+2c, +1
    const uint32_t pa1 = pack_bf16x2(p10, p11);  // row g+8, tokens t0+2c, +1
    const uint32_t pa2 = pack_bf16x2(p02, p03);  // row g,   tokens t0+8+2c, +1
    const uint32_t pa3 = pack_bf16x2(p12, p13);  // row g+8, tokens t0+8+2c, +1

    // ---- O = alpha * O + P V.  n-tile i <-> dh g*8+i (as B column owner) / dh {2c,2c+1}*8+i (as D owner)
    {
      const uint32_t va[4] = {cur.v[0].x, cur.v[0].y, cur.v[0].z, cur.v[0].w};
      const uint32_t vb[4] = {cur.v[1].x, cur.v[1].y, cur.v[1].z, cur.v[1].w};
      const uint32_t vc[4] = {cur.v[2].x, cur.v[2].y, cur.v[2].z, cur.v[2].w};
      const uint32_t vd[4] = {cur.v[3].x, cur.v[3].y, cur.v[3].z, cur.v[3].w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        o[i][0] *= al0; o[i][1] *= al0; o[i][2] *= al1; o[i][3] *= al1;
        const uint32_t sel = (i & 1) ? 0x7632u : 0x5410u;
        const uint32_t b0 = __byte_perm(va[i >> 1], vb[i >> 1], sel);
        const uint32_t b1 = __byte_perm(vc[i >> 1], vd[i >> 1], sel);
        mma_bf16_16816(o[i], pa0, pa1, pa2, pa3, b0, b1);
      }
    }
  }

  // ---- finalise: row sums across the 4 threads of a group, normalise, store
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
  // thread owns dh [c*16, c*16+16) of rows g and g+8: o[i][0] -> dh c*16+i, o[i][1] -> dh c*16+8+i
  if (qi0 < a.nq) {
    uint4 w0, w1;
    w0.x = pack_bf16x2(o[0][0] * inv0, o[1][0] * inv0); w0.y = pack_bf16x2(o[2][0] * inv0, o[3][0] * inv0);
    w0.z = pack_bf16x2(o[4][0] * inv0, o[5][0] * inv0); w0.w = pack_bf16x2(o[6][0] * inv0, o[7][0] * inv0);
    w1.x = pack_bf16x2(o[0][1] * inv0, o[1][1] * inv0); w1.y = pack_bf16x2(o[2][1] * inv0, o[3][1] * inv0);
    w1.z = pack_bf16x2(o[4][1] * inv0, o[5][1] * inv0); w1.w = pack_bf16x2(o[6][1] * inv0, o[7][1] * inv0);
    __nv_bfloat16* po = a.out + qrow0 * a.ldo + h * 64 + c * 16;
    *reinterpret_cast<uint4*>(po) = w0;
    *reinterpret_cast<uint4*>(po + 8) = w1;
  }
  if (qi1 < a.nq) {
    uint4 w0, w1;
    w0.x = pack_bf16x2(o[0][2] * inv1, o[1][2] * inv1); w0.y = pack_bf16x2(o[2][2] * inv1, o[3][2] * inv1);
    w0.z = pack_bf16x2(o[4][2] * inv1, o[5][2] * inv1); w0.w = pack_bf16x2(o[6][2] * inv1, o[7][2] * inv1);
    w1.x = pack_bf16x2(o[0][3] * inv1, o[1][3] * inv1); w1.y = pack_bf16x2(o[2][3] * inv1, o[3][3] * inv1);
    w1.z = pack_bf16x2(o[4][3] * inv1, o[5][3] * inv1); w1.w = pack_bf16x2(o[6][3] * inv1, o[7][3] * inv1);
    __nv_bfloat16* po = a.out + qrow1 * a.ldo + h * 64 + c * 16;
    *reinterpret_cast<uint4*>(po) = w0;
    *reinterpret_cast<uint4*>(po + 8) = w1;
  }
}

}  // namespace

int attention_launch(const AttentionArgs& a, cudaStream_t stream, const char** err) {
  if (a.rows <= 0 || a.nq <= 0 || a.heads <= 0 || a.kv_seg1 + a.kv_seg2 + a.kv_seg3 <= 0) {
    if (err) *err = "attention: empty problem";
    return TDC_EINVAL;
  }
  if (a.kv_mask != nullptr && a.kv_seg1 + a.kv_seg2 + a.kv_seg3 > 32) {
    if (err) *err = "attention: kv_mask supports at most 32 KV tokens per row";
    return TDC_EINVAL;
  }
  if ((a.ldq % 8) || (a.ldk % 8) || (a.ldv % 8) || (a.ldo % 8)) {
    if (err) *err = "attention: pitches must be multiples of 8 elements";
    return TDC_EINVAL;
  }
  const long long nqb = (a.nq + 15) / 16;
  const long long warps = static_cast<long long>(a.rows) * nqb * a.heads;
  const long long blocks = (warps + 3) / 4;
  if (blocks > 0x7fffffffLL) {
    if (err) *err = "attention: grid too large";
    return TDC_EINVAL;
  }
  constexpr int kSmem = 4 * kStages * kSlotsPerStage * 16;  // 48 KB per 4-warp block -> 4 blocks / SM
  {  // the opt-in is a per-device property: one flag per device ordinal
    static std::mutex mu;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev = (dev >= 0 && dev < 64) ? dev : 0;
    std::lock_guard<std::mutex> lock(mu);
    if (!attr_set[dev]) {
      cudaFuncSetAttribute(tdc_attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
      cudaFuncSetAttribute(tdc_attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
      attr_set[dev] = true;
    }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(blocks));
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = kSmem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  const bool segs = a.q_seg2 > 0 || a.kv_seg2 > 0 || a.kv_seg3 > 0;
  const cudaError_t rc = segs ? cudaLaunchKernelEx(&cfg, tdc_attention_kernel<true>, a)
                              : cudaLaunchKernelEx(&cfg, tdc_attention_kernel<false>, a);
  if (rc != cudaSuccess) {
    if (err) *err = cudaGetErrorString(rc);
    return TDC_ECUDA;
  }
  return TDC_OK;
}

}  // namespace tdc
