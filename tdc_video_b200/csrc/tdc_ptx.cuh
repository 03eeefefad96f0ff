// Thin inline-PTX wrappers for the sm_100a features the TDC kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and
// thread-block-cluster helpers.  Nothing in here is a library call; every
// wrapper is one PTX instruction (or a spin loop around one).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>

namespace tdc {

#ifndef TDC_SPIN_TIMEOUT_CYCLES
// A wedged mbarrier wait becomes a trap (-> cudaErrorLaunchFailure) after this
// many SM cycles instead of a hung GPU.  ~4 s at 1.9 GHz.
#define TDC_SPIN_TIMEOUT_CYCLES (8000000000ll)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() {
  uint32_t l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// ----------------------------------------------------------------------------
// programmatic dependent launch (the kernel may start while its predecessor in the stream drains; it must not touch
// the predecessor's data before grid_dependency_wait() returns).  No-ops when launched without the attribute.
// ----------------------------------------------------------------------------
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// host side: TDC_PDL=1 (dev knob, read once) adds cudaLaunchAttributeProgrammaticStreamSerialization to the launches
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("TDC_PDL"); return e != nullptr && atoi(e) == 1; }();
  return on;
}

// ----------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Arrive on the barrier at the same smem offset in CTA `cta` of the cluster.  Default (.release.cta)
// semantics on purpose: the callers publish no generic-proxy data through this arrive (it hands a TMEM
// accumulator back after tcgen05.wait::ld + tcgen05.fence), and a .release.cluster arrive costs a
// MEMBAR.ALL + ERRBAR per call — 28 % of the epilogue warps' stall samples in the K=768 GEMMs.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// acquire.cluster flavour for barriers that peers in the cluster arrive on
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
template <bool kCluster = false>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (kCluster ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!(kCluster ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity))) {
    if (clock64() - t0 > TDC_SPIN_TIMEOUT_CYCLES) {
      printf("tdc: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// ----------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// L2 eviction-priority policies for TMA (.L2::cache_hint operand; fixed encodings of
// createpolicy.fractional.L2::evict_{normal,first,last} with fraction 1.0)
constexpr uint64_t kL2EvictNormal = 0x1000000000000000ull;
constexpr uint64_t kL2EvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kL2EvictLast = 0x14F0000000000000ull;

// 2-D tiled load, completion counted in bytes on `bar` (this CTA's barrier).
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0,
                                            int32_t c1, uint64_t hint = kL2EvictNormal) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
// Same, issued by either CTA of a cta_group::2 pair; the bytes are credited to
// the barrier at the same offset in the pair's leader (even-ranked) CTA.
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0,
                                                 int32_t c1, uint64_t hint = kL2EvictNormal) {
  const uint32_t bar_leader = smem_u32(bar) & 0xFEFFFFFFu;  // clear the peer-CTA bit
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(bar_leader), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}

// cta_group::2 load multicast to the CTAs in `cta_mask` (cluster ranks): the box lands at the same smem offset
// in every destination CTA and its bytes are credited to the barrier at the same offset in the leader of the
// destination CTA's pair.
__device__ __forceinline__ void tma_load_2d_pair_multicast(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                                           int32_t c0, int32_t c1, uint16_t cta_mask,
                                                           uint64_t hint = kL2EvictNormal) {
  const uint32_t bar_leader = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      ".L2::cache_hint [%0], [%1, {%4, %5}], [%2], %3, %6;" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(bar_leader), "h"(cta_mask), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}

// 2-D tiled store smem -> global (bulk async-group completion; out-of-bounds part is clipped).
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// 3-D variant: coordinates (column, row, slab)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1,
                                             int32_t c2, uint64_t hint) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still have to READ their smem source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// wait until at most N of this thread's bulk groups are incomplete (writes performed)
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  if (CG == 1)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
  else
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_relinquish() {
  if (CG == 1)
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  else
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if (CG == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, bf16/fp16 inputs, fp32 accumulate.
template <int CG>
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  if (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Arrive on `bar` once every tcgen05 op this thread issued so far has retired.
// CG==2: the arrive is multicast to the barrier at the same offset in both CTAs.
template <int CG>
__device__ __forceinline__ void umma_commit(uint64_t* bar, uint16_t mask = 0x3) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
  else {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(mask)
        : "memory");
  }
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
// (thread t of the warp reads TMEM lane base+t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------
// UMMA descriptors (bit layouts: SM100 matrix descriptor / instruction descriptor)
// ----------------------------------------------------------------------------
// K-major operand tile whose rows are exactly one 128-byte swizzle span
// (64 bf16): rows at 128 B pitch, 8-row swizzle atoms at 1024 B pitch.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);  // [0,14)  start address >> 4
  d |= static_cast<uint64_t>(1) << 16;                      // [16,30) leading byte offset (unused for SW128 K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // [32,46) stride byte offset = 8 rows * 128 B
  d |= static_cast<uint64_t>(1) << 46;                      // [46,48) descriptor version 1 (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                      // [61,64) layout = SWIZZLE_128B
  return d;
}

// MN-major operand tile (the operand's M/N index is the contiguous one, e.g. a row-major [K, N] matrix used as B):
// exactly what a 128B-swizzled TMA box of [k rows x 64 columns] delivers.  SBO = bytes between groups of 8 K rows
// (1024 when the rows are dense), LBO = bytes between successive 64-element blocks along M/N.  Needs the
// corresponding transpose bit of the instruction descriptor (bit 15 for A, bit 16 for B).
// Verified on sm_100a by tools/umma_mnmajor_test.cu.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);       // [0,14)  start address >> 4
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;  // [16,30) leading byte offset
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;  // [32,46) stride byte offset
  d |= static_cast<uint64_t>(1) << 46;                           // [46,48) descriptor version 1 (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                           // [61,64) layout = SWIZZLE_128B
  return d;
}
constexpr uint32_t kIdescBMnMajor = 1u << 16;  // OR into make_idesc_bf16_f32(...) when B is MN-major
constexpr uint32_t kIdescAMnMajor = 1u << 15;

// kind::f16 instruction descriptor: A/B bf16 K-major, D fp32.
__host__ __device__ constexpr uint32_t make_idesc_bf16_f32(int m, int n) {
  return (1u << 4)                             // [4,6)   D format = F32
         | (1u << 7)                           // [7,10)  A format = BF16
         | (1u << 10)                          // [10,13) B format = BF16
         | (0u << 15) | (0u << 16)             // A, B K-major
         | (static_cast<uint32_t>(n >> 3) << 17)  // [17,23) N >> 3
         | (static_cast<uint32_t>(m >> 4) << 24); // [24,29) M >> 4
}

// ----------------------------------------------------------------------------
// small math helpers
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// Exact-erf GELU, gelu(x) = x * Phi(x), evaluated for the GEMM epilogue (bf16 output) with ONE MUFU op:
//   Phi(-|x|) = 0.5 * erfc(z),  z = |x| / sqrt(2),  erfc(z) = erfcx(z) * exp(-z^2)
// and 0.5*erfcx(z) as a degree-8 polynomial in s = z / 3.5 on [0, 1] (Chebyshev fit; z is clamped at 3.5
// where Phi(-x) < 4e-7).  Max |error| vs the erf form: 1.5e-5 absolute (fp32 evaluation, checked over
// [-9, 9]), i.e. below half a bf16 ulp of the result wherever |gelu(x)| > 4e-3 — the tanh "approximate GELU"
// is ~1e-3 off by comparison.  libdevice erff (~40 instructions) and the A&S 7.1.26 form (rcp + ex2 = 2 MUFU)
// both left the GELU GEMM epilogue-bound: the MUFU pipe issues 16 lanes/clk/SM.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float s = fminf(fabsf(x) * (0.70710678118654752440f / 3.5f), 1.0f);
  float g = 2.612654719e+00f;
  g = fmaf(g, s, -1.273282320e+01f);
  g = fmaf(g, s, 2.688603589e+01f);
  g = fmaf(g, s, -3.264060733e+01f);
  g = fmaf(g, s, 2.571818791e+01f);
  g = fmaf(g, s, -1.426494436e+01f);
  g = fmaf(g, s, 5.968592886e+00f);
  g = fmaf(g, s, -1.969402482e+00f);
  g = fmaf(g, s, 4.999701200e-01f);
  float e;  // exp(-z^2) = 2^(-(3.5 s)^2 * log2 e)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(s * s * (-3.5f * 3.5f * 1.4426950408889634f)));
  e *= g;                                   // Phi(-|x|)
  return x * (x >= 0.f ? 1.0f - e : e);
}

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 issue two fp32 lanes per instruction) ----------
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// gelu_erf_fast on two values at once: the degree-8 Horner chain and the surrounding multiplies run as
// FFMA2 / FMUL2 (half the FMA-pipe instructions of the scalar form); abs / min / select / ex2 stay scalar.
__device__ __forceinline__ void gelu_erf_fast_x2(float& x0, float& x1) {
  const float s0 = fminf(fabsf(x0) * (0.70710678118654752440f / 3.5f), 1.0f);
  const float s1 = fminf(fabsf(x1) * (0.70710678118654752440f / 3.5f), 1.0f);
  const uint64_t s = pack_f32x2(s0, s1);
#define TDC_C2(v) pack_f32x2(v, v)
  uint64_t g = TDC_C2(2.612654719e+00f);
  g = fma_f32x2(g, s, TDC_C2(-1.273282320e+01f));
  g = fma_f32x2(g, s, TDC_C2(2.688603589e+01f));
  g = fma_f32x2(g, s, TDC_C2(-3.264060733e+01f));
  g = fma_f32x2(g, s, TDC_C2(2.571818791e+01f));
  g = fma_f32x2(g, s, TDC_C2(-1.426494436e+01f));
  g = fma_f32x2(g, s, TDC_C2(5.968592886e+00f));
  g = fma_f32x2(g, s, TDC_C2(-1.969402482e+00f));
  g = fma_f32x2(g, s, TDC_C2(4.999701200e-01f));
  const uint64_t w = mul_f32x2(mul_f32x2(s, s), TDC_C2(-3.5f * 3.5f * 1.4426950408889634f));
#undef TDC_C2
  float w0, w1, e0, e1;
  unpack_f32x2(w, w0, w1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(w0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(w1));
  const uint64_t p = mul_f32x2(pack_f32x2(e0, e1), g);   // Phi(-|x|) for both
  float p0, p1;
  unpack_f32x2(p, p0, p1);
  p0 = x0 >= 0.f ? 1.0f - p0 : p0;
  p1 = x1 >= 0.f ? 1.0f - p1 : p1;
  const uint64_t y = mul_f32x2(pack_f32x2(x0, x1), pack_f32x2(p0, p1));
  unpack_f32x2(y, x0, x1);
}

}  // namespace tdc
