// Adaptive segmentation helpers — the step right before the TDC path:
// cosine similarity of consecutive frames' DINO features and selection of the
// `max_num_segments` lowest similarities as segment boundaries
// (reference: tdc/cambrian_arch.py:832-849, CambrianMetaForCausalLM.adapt_segment).
//
// HBM-bound reduction over [n_frames, dim] (dim = 576*1536 for DINOv2-giant): every frame is
// needed twice (as "previous" and as "next"); blocks of neighbouring pairs working on the same
// feature slice are adjacent in launch order, so the second use is an L2 hit and DRAM sees each
// byte once.  Deterministic: per-slice partial sums are written out and reduced in a fixed order
// (no floating-point atomics).
#include "tdc_kernels.cuh"
#include "tdc_ptx.cuh"

#include <cuda_fp16.h>

namespace tdc {

namespace {

__device__ __forceinline__ void unpack8(const uint4& raw, int dtype, float (&f)[8]) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (dtype == TDC_BF16) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      f[2 * i] = __low2float(v);
      f[2 * i + 1] = __high2float(v);
    } else {
      const __half2 v = *reinterpret_cast<const __half2*>(&w[i]);
      f[2 * i] = __low2float(v);
      f[2 * i + 1] = __high2float(v);
    }
  }
}

// grid (pairs, slices); partial[pair][slice] = {dot(f_i, f_{i+1}), |f_i|^2, |f_{i+1}|^2} over the slice
__global__ void __launch_bounds__(256) frame_pair_partial_kernel(const void* __restrict__ feats, int dtype,
                                                                 long long dim, int slices,
                                                                 float* __restrict__ partial) {
  const int pair = blockIdx.x, slice = blockIdx.y;
  const int per = dtype == TDC_F32 ? 4 : 8;  // elements per 16-byte vector
  const long long nvec = dim / per;
  const long long v0 = nvec * slice / slices, v1 = nvec * (slice + 1) / slices;
  const size_t esz = dtype == TDC_F32 ? 4 : 2;
  const uint4* a = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(feats) + static_cast<size_t>(pair) * dim * esz);
  const uint4* b = reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(feats) + static_cast<size_t>(pair + 1) * dim * esz);
  float dot = 0.f, na = 0.f, nb = 0.f;
  for (long long i = v0 + threadIdx.x; i < v1; i += 256) {
    const uint4 ra = __ldg(a + i), rb = __ldg(b + i);
    if (dtype == TDC_F32) {
      const float* x = reinterpret_cast<const float*>(&ra);
      const float* y = reinterpret_cast<const float*>(&rb);
#pragma unroll
      for (int j = 0; j < 4; ++j) { dot = fmaf(x[j], y[j], dot); na = fmaf(x[j], x[j], na); nb = fmaf(y[j], y[j], nb); }
    } else {
      float x[8], y[8];
      unpack8(ra, dtype, x);
      unpack8(rb, dtype, y);
#pragma unroll
      for (int j = 0; j < 8; ++j) { dot = fmaf(x[j], y[j], dot); na = fmaf(x[j], x[j], na); nb = fmaf(y[j], y[j], nb); }
    }
  }
  __shared__ float red[3][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
    na += __shfl_xor_sync(0xffffffffu, na, o);
    nb += __shfl_xor_sync(0xffffffffu, nb, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = dot; red[1][threadIdx.x >> 5] = na; red[2][threadIdx.x >> 5] = nb; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    partial[(static_cast<size_t>(pair) * slices + slice) * 3 + threadIdx.x] = s;
  }
}

// cos[pair] = dot / (max(|a|, eps) * max(|b|, eps))   (F.cosine_similarity, eps = 1e-8), accumulated in fp32 and
// then rounded to the features' own dtype: the reference evaluates F.cosine_similarity in the DINO feature dtype
// (bf16 / fp16 under the model dtype) and argsorts those rounded values, so ties created by the rounding must
// exist here too (they are broken by index, like a stable sort).  The intermediate roundings of torch's bf16
// composition are not reproduced — only the final value's precision.
__global__ void cosine_finish_kernel(const float* __restrict__ partial, int pairs, int slices, int dtype,
                                     float* __restrict__ cos) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pairs) return;
  float dot = 0.f, na = 0.f, nb = 0.f;
  for (int s = 0; s < slices; ++s) {
    const float* q = partial + (static_cast<size_t>(p) * slices + s) * 3;
    dot += q[0]; na += q[1]; nb += q[2];
  }
  float c = dot / (fmaxf(sqrtf(na), 1e-8f) * fmaxf(sqrtf(nb), 1e-8f));
  if (dtype == TDC_BF16) c = __bfloat162float(__float2bfloat16_rn(c));
  else if (dtype == TDC_F16) c = __half2float(__float2half_rn(c));
  cos[p] = c;
}

// The k smallest values' indices, ascending by index (== sort(argsort(x)[:k])); ties broken by index.
// NaN (inf/inf from overflowed fp16 features) ranks last, as in torch.argsort: the key is +inf, so at most k
// flags are ever set and the store below cannot run past the k-element output.
__global__ void __launch_bounds__(256) select_smallest_kernel(const float* __restrict__ x, int n, int k,
                                                              long long* __restrict__ out) {
  extern __shared__ unsigned char sel[];  // n flags
  for (int i = threadIdx.x; i < n; i += 256) {
    const float xi = isnan(x[i]) ? INFINITY : x[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const float xj = isnan(x[j]) ? INFINITY : x[j];
      rank += (xj < xi) || (xj == xi && j < i);
    }
    sel[i] = rank < k;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += 256) {
    if (!sel[i]) continue;
    int pos = 0;
    for (int j = 0; j < i; ++j) pos += sel[j];
    if (pos < k) out[pos] = i;
  }
}

}  // namespace

int frame_cosine_launch(const void* feats, int dtype, int n_frames, long long dim, float* partial_ws, float* cos,
                        cudaStream_t stream, const char** err) {
  if (n_frames < 2) return TDC_OK;
  const int per = dtype == TDC_F32 ? 4 : 8;
  if (dim <= 0 || dim % per != 0) {
    if (err) *err = "frame_cosine: feature dim must be a multiple of 16 bytes";
    return TDC_EINVAL;
  }
  const int pairs = n_frames - 1;
  const int slices = frame_cosine_slices(dim);
  frame_pair_partial_kernel<<<dim3(pairs, slices), 256, 0, stream>>>(feats, dtype, dim, slices, partial_ws);
  cosine_finish_kernel<<<(pairs + 127) / 128, 128, 0, stream>>>(partial_ws, pairs, slices, dtype, cos);
  const cudaError_t rc = cudaGetLastError();
  if (rc != cudaSuccess) {
    if (err) *err = cudaGetErrorString(rc);
    return TDC_ECUDA;
  }
  return TDC_OK;
}

int select_smallest_launch(const float* x, int n, int k, long long* out, cudaStream_t stream, const char** err) {
  if (n <= 0 || k <= 0) return TDC_OK;
  if (n > 40000) {
    if (err) *err = "select_smallest: n too large";
    return TDC_EINVAL;
  }
  select_smallest_kernel<<<1, 256, n, stream>>>(x, n, k, out);
  const cudaError_t rc = cudaGetLastError();
  if (rc != cudaSuccess) {
    if (err) *err = cudaGetErrorString(rc);
    return TDC_ECUDA;
  }
  return TDC_OK;
}

}  // namespace tdc
