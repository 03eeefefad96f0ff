// Host-side interface of the tcgen05 GEMM (implementation: gemm_sm100.cu).
//
//   C[M,N] = epilogue( A[M,K] . W[N,K]^T )        A, W bf16 row-major (K contiguous)
//
// which is exactly nn.Linear's layout (weight [out_features, in_features]), so
// reference weights are used as stored, only cast to bf16.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tdc {

enum GemmEpilogueMode : int {
  EPI_BIAS_BF16 = 0,       // out bf16  = acc + bias
  EPI_BIAS_GELU_BF16 = 1,  // out bf16  = gelu_erf(acc + bias)
  EPI_BIAS_F32 = 2,        // out fp32  = acc + bias (the residual add lives in the LayerNorm kernel)
};

struct GemmProblem {
  const void* a = nullptr;  // bf16 [M, K], row pitch lda elements
  const void* w = nullptr;  // bf16 [N, K], row pitch ldw elements
  long long lda = 0, ldw = 0;
  int m = 0, n = 0, k = 0;
  void* out = nullptr;  // bf16 or fp32 [M, N], row pitch ldo elements
  long long ldo = 0;
  const float* bias = nullptr;  // [N] or null
  // Optional slab layout of the output: column n is stored in slab n / slab_cols at column n % slab_cols,
  // slabs `slab_stride` elements apart (each slab a dense [M, slab_cols] matrix of pitch ldo).  Used to give
  // every cross-attention layer its own contiguous K/V matrix.  slab_cols == 0: plain [M, N] output.
  int slab_cols = 0;
  long long slab_stride = 0;
  int mode = EPI_BIAS_BF16;
  int cta_group = 0;  // 0 = library default, 1 = single-CTA tiles, 2 = CTA-pair (cta_group::2) tiles
};

// Returns 0 on success, otherwise a negative tdc_status; on failure *err (if
// non-null) points at a static description.
int gemm_launch(const GemmProblem& p, cudaStream_t stream, const char** err);

// GEMM + residual + LayerNorm in one kernel (gemm_ln_sm100.cu):
//   out_f32 = LayerNorm(A . W^T + bias + resid) * gamma + beta,  out_bf16 = bf16(out_f32)
// — BertSelfOutput / BertOutput of the reference (tdc/Qformer.py:285-289, 371-375).  resid may alias out_f32.
struct GemmLnProblem {
  const void* a = nullptr;  // bf16 [M, K]
  const void* w = nullptr;  // bf16 [N, K]
  long long lda = 0, ldw = 0;
  int m = 0, n = 0, k = 0;
  const float* bias = nullptr;   // [N]
  const float* resid = nullptr;  // fp32 [M, N], pitch ldr
  long long ldr = 0;
  const float* gamma = nullptr;  // [N]
  const float* beta = nullptr;   // [N]
  float eps = 1e-12f;
  float* out_f32 = nullptr;      // fp32 [M, N], pitch ldo
  void* out_bf16 = nullptr;      // bf16 [M, N], pitch ldo
  long long ldo = 0;
};
bool gemm_ln_supported(int n);   // N <= 768 (a cluster of up to 3 CTAs x 256 columns holds one LayerNorm row)
int gemm_ln_launch(const GemmLnProblem& p, cudaStream_t stream, const char** err);

// Number of kernels gemm_launch enqueues (always 1) — kept for launch accounting.
inline int gemm_launch_count() { return 1; }

}  // namespace tdc
