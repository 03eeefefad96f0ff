// Kernels of the upstream ("frames") entry of the TDC path: everything between the towers' outputs and the
// Q-Former that is not a GEMM.  Reference: tdc/cambrian_arch.py:1149-1150 (mm_projector on every frame
// token), :1269-1281 (image_newline closes every row of the token grid), :1611-1614 (audio_proj tokens
// appended to every frame), :1629-1640 (queries = avg-pool of the key frame through query_proj).
//
// All of them are plain HBM-bound data movement (16-byte vectors, one read + one write of their operand);
// the fold helpers (transpose, matvec) run once per weight load.
#include "tdc_kernels.cuh"
#include "tdc_ptx.cuh"

#include <cuda_fp16.h>

namespace tdc {

namespace {

// dst block i = src block idx[i]; a block is `vecs` 16-byte vectors (one frame's tokens).
__global__ void __launch_bounds__(256) gather_blocks_kernel(const uint4* __restrict__ src, const int32_t* __restrict__ idx,
                                                            uint4* __restrict__ dst, long long vecs, int n_src) {
  const long long item = blockIdx.y;
  const int i = min(max(idx[item], 0), n_src - 1);   // a bad index must not read outside the caller's buffer
  const uint4* s = src + static_cast<long long>(i) * vecs;
  uint4* d = dst + item * vecs;
  for (long long v = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; v < vecs;
       v += static_cast<long long>(gridDim.x) * 256)
    d[v] = __ldg(s + v);
}

// out[c, r] = in[r, c] (bf16), 32x32 tiles through shared memory.
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, int rows, int cols,
                                                             __nv_bfloat16* __restrict__ out) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    if (r < rows && c < cols) tile[j][tx] = in[static_cast<long long>(r) * cols + c];
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (r < rows && c < cols) out[static_cast<long long>(c) * rows + r] = tile[tx][j];
  }
}

// y[n] = sum_k W[n, k] * x[k] + b[n]   (W bf16 [n, k], x / b / y fp32); one warp per output.
__global__ void __launch_bounds__(256) matvec_bias_kernel(const __nv_bfloat16* __restrict__ w, const float* __restrict__ x,
                                                          const float* __restrict__ b, float* __restrict__ y, int n,
                                                          int k) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const __nv_bfloat16* wr = w + static_cast<long long>(row) * k;
  float acc = 0.f;
  for (int j = lane * 2; j < k; j += 64) {
    const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(wr + j);
    acc = fmaf(__low2float(v), x[j], acc);
    acc = fmaf(__high2float(v), x[j + 1], acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[row] = acc + (b != nullptr ? b[row] : 0.f);
}

// Token u of a frame's sequence with newline tokens: grid row g = u / (side + 1), column j = u % (side + 1);
// j < side -> visual token g * side + j, j == side -> the image_newline vector.
// adaptive_avg_pool1d over that sequence (bins [floor(i*L/K), ceil((i+1)*L/K)), L = side * (side + 1)).
__global__ void __launch_bounds__(256) pool_static_queries_kernel(const __nv_bfloat16* __restrict__ xv,
                                                                  const float* __restrict__ newline, int side, int d,
                                                                  int num_query, __nv_bfloat16* __restrict__ out) {
  const int c = blockIdx.x / num_query, i = blockIdx.x % num_query;
  const int L = side * (side + 1);
  const int start = static_cast<int>((static_cast<long long>(i) * L) / num_query);
  const int end = static_cast<int>((static_cast<long long>(i + 1) * L + num_query - 1) / num_query);
  const float inv = 1.0f / static_cast<float>(end - start);
  const __nv_bfloat16* frame = xv + static_cast<long long>(c) * side * side * d;
  // 8 columns (one 16-byte vector) per thread; the bin's tokens are independent loads, all issued before the adds
  for (int j = threadIdx.x; j < d / 8; j += 256) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int u = start; u < end; ++u) {
      const int g = u / (side + 1), col = u % (side + 1);
      if (col < side) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(frame + static_cast<long long>(g * side + col) * d) + j);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
          acc[2 * e] += __low2float(v);
          acc[2 * e + 1] += __high2float(v);
        }
      } else {
        // the reference pools the model-dtype (bf16) copy of the parameter
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += __bfloat162float(__float2bfloat16_rn(__ldg(newline + j * 8 + e)));
      }
    }
    uint4 pk;
    pk.x = pack_bf16x2(acc[0] * inv, acc[1] * inv);
    pk.y = pack_bf16x2(acc[2] * inv, acc[3] * inv);
    pk.z = pack_bf16x2(acc[4] * inv, acc[5] * inv);
    pk.w = pack_bf16x2(acc[6] * inv, acc[7] * inv);
    reinterpret_cast<uint4*>(out + (static_cast<long long>(c) * num_query + i) * d)[j] = pk;
  }
}

// 16-byte store; kMulticast: `p` is an NVSwitch multicast address (multimem.st: every GPU of the group gets the bytes)
template <bool kMulticast>
__device__ __forceinline__ void store16_mc(void* p, uint4 v) {
  if (kMulticast)
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
  else
    *reinterpret_cast<uint4*>(p) = v;
}

// local memory -> multicast address, 16 bytes per thread and iteration
__global__ void __launch_bounds__(512) multicast_copy_kernel(const uint4* __restrict__ src, uint4* dst, size_t n16) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    store16_mc<true>(dst + i, __ldg(src + i));
}

// static_out[c] = [ (side visual tokens, newline) x side | Ta audio tokens ]  in out_dtype (bf16 / fp16 / fp32).
// One warp per token, 16-byte stores; kMulticast: stored through an NVSwitch multicast address, so the key frames'
// pass-through tokens land on every GPU of the group while the kernel runs (the all-gather fused into the producer).
template <bool kMulticast>
__global__ void __launch_bounds__(256) assemble_static_kernel(const __nv_bfloat16* __restrict__ xv,
                                                              const __nv_bfloat16* __restrict__ xa,
                                                              const float* __restrict__ newline, int chunks, int side,
                                                              int ta, int d, void* __restrict__ out, int out_dtype) {
  const int ls = side * (side + 1) + ta;
  const long long tok = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const long long c = tok / ls;
  const int u = static_cast<int>(tok % ls);
  if (c >= chunks) return;
  const int lane = threadIdx.x & 31;
  const __nv_bfloat16* src = nullptr;
  if (u < side * (side + 1)) {
    const int g = u / (side + 1), col = u % (side + 1);
    if (col < side) src = xv + (c * side * side + g * side + col) * d;
  } else {
    src = xa + (c * ta + (u - side * (side + 1))) * d;
  }
  const size_t esz = out_dtype == TDC_F32 ? 4 : 2;
  uint8_t* dst = static_cast<uint8_t*>(out) + static_cast<size_t>(tok) * d * esz;
  for (int j = lane; j < d / 8; j += 32) {   // 8 columns per lane and iteration
    float v[8];
    if (src != nullptr) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(src) + j);
      const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 p2 = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
        v[2 * e] = __low2float(p2);
        v[2 * e + 1] = __high2float(p2);
      }
    } else {
      const float4 a = __ldg(reinterpret_cast<const float4*>(newline) + 2 * j);
      const float4 b = __ldg(reinterpret_cast<const float4*>(newline) + 2 * j + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    if (out_dtype == TDC_F32) {
      store16_mc<kMulticast>(dst + j * 32, make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]),
                                                      __float_as_uint(v[3])));
      store16_mc<kMulticast>(dst + j * 32 + 16, make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]),
                                                           __float_as_uint(v[6]), __float_as_uint(v[7])));
    } else if (out_dtype == TDC_BF16) {
      store16_mc<kMulticast>(dst + j * 16, make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                      pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7])));
    } else {
      uint32_t h[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __half2 hh = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
        h[e] = *reinterpret_cast<const uint32_t*>(&hh);
      }
      store16_mc<kMulticast>(dst + j * 16, make_uint4(h[0], h[1], h[2], h[3]));
    }
  }
}

// rows [row0, row0 + count) of every slab = the slab's fp32 source vector (bf16): the K/V of the newline tokens
__global__ void __launch_bounds__(256) broadcast_rows_kernel(const float* __restrict__ src, int width, int slabs,
                                                             __nv_bfloat16* __restrict__ dst, long long slab_stride,
                                                             long long row0, int count) {
  const int slab = blockIdx.y;
  const float* s = src + static_cast<long long>(slab) * width;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < count * width; e += gridDim.x * 256) {
    const int r = e / width, c = e % width;
    dst[slab * slab_stride + (row0 + r) * width + c] = __float2bfloat16_rn(s[c]);
  }
}

// F.interpolate(mode="bilinear", align_corners=False) on a token grid kept token-major: in [bs, s_in^2, d] ->
// out [bs, s_out^2, d] (the reference permutes to [bs, d, s, s], interpolates in fp32 and permutes back,
// cambrian_arch.py:1107-1131).  Source coordinate of output cell o: max((o + 0.5) * s_in / s_out - 0.5, 0).
__global__ void __launch_bounds__(256) resize_tokens_bilinear_kernel(const void* __restrict__ in, int in_dtype, int s_in,
                                                                     int s_out, int d, void* __restrict__ out,
                                                                     int out_dtype, long long tokens_out) {
  const long long tok = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (tok >= tokens_out) return;
  const int lane = threadIdx.x & 31;
  const long long b = tok / (s_out * s_out);
  const int oy = static_cast<int>(tok % (s_out * s_out)) / s_out, ox = static_cast<int>(tok % s_out);
  const float scale = static_cast<float>(s_in) / static_cast<float>(s_out);
  const float fy = fmaxf((oy + 0.5f) * scale - 0.5f, 0.f), fx = fmaxf((ox + 0.5f) * scale - 0.5f, 0.f);
  const int y0 = min(static_cast<int>(fy), s_in - 1), x0 = min(static_cast<int>(fx), s_in - 1);
  const int y1 = min(y0 + 1, s_in - 1), x1 = min(x0 + 1, s_in - 1);
  const float ly = fy - y0, lx = fx - x0;
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
  const size_t isz = in_dtype == TDC_F32 ? 4 : 2, osz = out_dtype == TDC_F32 ? 4 : 2;
  const uint8_t* base = static_cast<const uint8_t*>(in) + static_cast<size_t>(b) * s_in * s_in * d * isz;
  auto load4 = [&](int y, int x, int j) -> float4 {
    const uint8_t* p = base + (static_cast<size_t>(y) * s_in + x) * d * isz;
    if (in_dtype == TDC_F32) return __ldg(reinterpret_cast<const float4*>(p) + j);
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p) + j);
    if (in_dtype == TDC_BF16) {
      const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
      const __nv_bfloat162 c = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
      return make_float4(__low2float(a), __high2float(a), __low2float(c), __high2float(c));
    }
    const __half2 a = *reinterpret_cast<const __half2*>(&raw.x);
    const __half2 c = *reinterpret_cast<const __half2*>(&raw.y);
    return make_float4(__low2float(a), __high2float(a), __low2float(c), __high2float(c));
  };
  uint8_t* dst = static_cast<uint8_t*>(out) + static_cast<size_t>(tok) * d * osz;
  for (int j = lane; j < d / 4; j += 32) {
    const float4 a = load4(y0, x0, j), bq = load4(y0, x1, j), c = load4(y1, x0, j), e = load4(y1, x1, j);
    float4 v;
    v.x = w00 * a.x + w01 * bq.x + w10 * c.x + w11 * e.x;
    v.y = w00 * a.y + w01 * bq.y + w10 * c.y + w11 * e.y;
    v.z = w00 * a.z + w01 * bq.z + w10 * c.z + w11 * e.z;
    v.w = w00 * a.w + w01 * bq.w + w10 * c.w + w11 * e.w;
    if (out_dtype == TDC_F32) {
      reinterpret_cast<float4*>(dst)[j] = v;
    } else if (out_dtype == TDC_BF16) {
      uint2 pk;
      pk.x = pack_bf16x2(v.x, v.y);
      pk.y = pack_bf16x2(v.z, v.w);
      reinterpret_cast<uint2*>(dst)[j] = pk;
    } else {
      const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&h0);
      pk.y = *reinterpret_cast<const uint32_t*>(&h1);
      reinterpret_cast<uint2*>(dst)[j] = pk;
    }
  }
}

// Tokens under every query of a Q x Q query grid, window-major (tdc/cambrian_arch.py:624-645): in [bs, (Q r)^2, d]
// (row-major token grid) -> out [bs, Q, Q, r, r, d] bf16.  One warp per output token.
__global__ void __launch_bounds__(256) window_rearrange_kernel(const void* __restrict__ in, int in_dtype, int q, int r, int d,
                                                               __nv_bfloat16* __restrict__ out, long long tokens) {
  const long long tok = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (tok >= tokens) return;
  const int lane = threadIdx.x & 31;
  const int grid = q * r;
  long long t = tok;
  const int wx = static_cast<int>(t % r); t /= r;
  const int wy = static_cast<int>(t % r); t /= r;
  const int qx = static_cast<int>(t % q); t /= q;
  const int qy = static_cast<int>(t % q); t /= q;
  const long long src = (t * grid + (qy * r + wy)) * grid + (qx * r + wx);
  const size_t isz = in_dtype == TDC_F32 ? 4 : 2;
  const uint8_t* sp = static_cast<const uint8_t*>(in) + static_cast<size_t>(src) * d * isz;
  __nv_bfloat16* dp = out + tok * d;
  for (int j = lane; j < d / 4; j += 32) {
    uint2 pk;
    if (in_dtype == TDC_BF16) {
      pk = __ldg(reinterpret_cast<const uint2*>(sp) + j);
    } else if (in_dtype == TDC_F32) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(sp) + j);
      pk.x = pack_bf16x2(v.x, v.y);
      pk.y = pack_bf16x2(v.z, v.w);
    } else {
      const uint2 raw = __ldg(reinterpret_cast<const uint2*>(sp) + j);
      const __half2 a = *reinterpret_cast<const __half2*>(&raw.x);
      const __half2 b = *reinterpret_cast<const __half2*>(&raw.y);
      pk.x = pack_bf16x2(__low2float(a), __high2float(a));
      pk.y = pack_bf16x2(__low2float(b), __high2float(b));
    }
    reinterpret_cast<uint2*>(dp)[j] = pk;
  }
}

// out[r, :] = base[r, :] + sum_t softmax(logits[r, :T])[t] * parts[t][r, :]  — VisionAggregationLayer's combination of the
// per-tower aggregates (tdc/vision_sampler.py:468-474, 505-507).  One warp per row, T <= 8.
__global__ void __launch_bounds__(256) combine_parts_kernel(const float* __restrict__ base, const float* __restrict__ parts,
                                                            const float* __restrict__ logits, int ld_logits, int num_parts,
                                                            long long rows, int width, float* __restrict__ out) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float w[8];
  float mx = -INFINITY;
  for (int t = 0; t < num_parts; ++t) { w[t] = logits[row * ld_logits + t]; mx = fmaxf(mx, w[t]); }
  float sum = 0.f;
  for (int t = 0; t < num_parts; ++t) { w[t] = expf(w[t] - mx); sum += w[t]; }
  const float inv = 1.0f / sum;
  for (int j = lane; j < width / 4; j += 32) {
    float4 acc = reinterpret_cast<const float4*>(base + row * width)[j];
    for (int t = 0; t < num_parts; ++t) {
      const float4 p = reinterpret_cast<const float4*>(parts + (static_cast<long long>(t) * rows + row) * width)[j];
      const float wt = w[t] * inv;
      acc.x = fmaf(wt, p.x, acc.x); acc.y = fmaf(wt, p.y, acc.y); acc.z = fmaf(wt, p.z, acc.z); acc.w = fmaf(wt, p.w, acc.w);
    }
    reinterpret_cast<float4*>(out + row * width)[j] = acc;
  }
}

// out[r] = a[b[r]]  (row -> chunk -> prompt)
__global__ void compose_index_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int32_t* __restrict__ out,
                                     long long n, int n_a) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[min(max(b[i], 0), n_a - 1)];
}

int launched(const char** err) {
  const cudaError_t rc = cudaGetLastError();
  if (rc != cudaSuccess) {
    if (err) *err = cudaGetErrorString(rc);
    return TDC_ECUDA;
  }
  return TDC_OK;
}

}  // namespace

int gather_blocks_launch(const void* src, const int32_t* idx, void* dst, long long items, long long block_bytes,
                         int n_src, cudaStream_t stream, const char** err) {
  if (items <= 0) return TDC_OK;
  if (block_bytes % 16 != 0 || items > 65535) {
    if (err) *err = "gather_blocks: block size must be a multiple of 16 bytes and at most 65535 blocks per call";
    return TDC_EINVAL;
  }
  const long long vecs = block_bytes / 16;
  int bx = static_cast<int>((vecs + 255) / 256);
  if (bx > 32) bx = 32;
  gather_blocks_kernel<<<dim3(bx, static_cast<unsigned>(items)), 256, 0, stream>>>(
      static_cast<const uint4*>(src), idx, static_cast<uint4*>(dst), vecs, n_src);
  return launched(err);
}

int transpose_bf16_launch(const __nv_bfloat16* in, int rows, int cols, __nv_bfloat16* out, cudaStream_t stream,
                          const char** err) {
  transpose_bf16_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), 256, 0, stream>>>(in, rows, cols, out);
  return launched(err);
}

int matvec_bias_launch(const __nv_bfloat16* w, const float* x, const float* b, float* y, int n, int k,
                       cudaStream_t stream, const char** err) {
  if (k % 2 != 0) {
    if (err) *err = "matvec: k must be even";
    return TDC_EINVAL;
  }
  matvec_bias_kernel<<<(n + 7) / 8, 256, 0, stream>>>(w, x, b, y, n, k);
  return launched(err);
}

int pool_static_queries_launch(const __nv_bfloat16* xv, const float* newline, int chunks, int side, int d,
                               int num_query, __nv_bfloat16* out, cudaStream_t stream, const char** err) {
  if (chunks <= 0) return TDC_OK;
  if (d % 8 != 0) {
    if (err) *err = "pool_static_queries: d must be a multiple of 8";
    return TDC_EINVAL;
  }
  pool_static_queries_kernel<<<static_cast<unsigned>(chunks * num_query), 256, 0, stream>>>(xv, newline, side, d,
                                                                                           num_query, out);
  return launched(err);
}

int multicast_copy_launch(const void* src, void* dst, size_t bytes, int ctas, cudaStream_t stream, const char** err) {
  if (bytes == 0) return TDC_OK;
  if (bytes % 16 != 0 || ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15)) {
    if (err) *err = "multicast_copy: bytes and both pointers must be multiples of 16";
    return TDC_EINVAL;
  }
  const size_t n16 = bytes / 16;
  const unsigned blocks = static_cast<unsigned>(std::min<size_t>(ctas > 0 ? ctas : 16, (n16 + 511) / 512));
  multicast_copy_kernel<<<blocks, 512, 0, stream>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), n16);
  if (cudaGetLastError() != cudaSuccess) {
    if (err) *err = "multicast_copy launch failed";
    return TDC_ECUDA;
  }
  return TDC_OK;
}

int assemble_static_launch(const __nv_bfloat16* xv, const __nv_bfloat16* xa, const float* newline, int chunks, int side,
                           int ta, int d, void* out, int out_dtype, bool multicast, cudaStream_t stream,
                           const char** err) {
  if (chunks <= 0) return TDC_OK;
  const long long toks = static_cast<long long>(chunks) * (side * (side + 1) + ta);
  if (d % 8 != 0 || (reinterpret_cast<uintptr_t>(out) & 15)) {
    if (err) *err = "assemble_static: d must be a multiple of 8 and the output 16-byte aligned";
    return TDC_EINVAL;
  }
  const unsigned blocks = static_cast<unsigned>((toks + 7) / 8);
  if (multicast)
    assemble_static_kernel<true><<<blocks, 256, 0, stream>>>(xv, xa, newline, chunks, side, ta, d, out, out_dtype);
  else
    assemble_static_kernel<false><<<blocks, 256, 0, stream>>>(xv, xa, newline, chunks, side, ta, d, out, out_dtype);
  return launched(err);
}

int resize_tokens_bilinear_launch(const void* in, int in_dtype, int bs, int s_in, int s_out, int d, void* out,
                                  int out_dtype, cudaStream_t stream, const char** err) {
  const long long toks = static_cast<long long>(bs) * s_out * s_out;
  if (toks <= 0) return TDC_OK;
  if (d % 4 != 0 || s_in <= 0 || s_out <= 0) {
    if (err) *err = "resize_tokens_bilinear: d must be a multiple of 4, sides positive";
    return TDC_EINVAL;
  }
  resize_tokens_bilinear_kernel<<<static_cast<unsigned>((toks + 7) / 8), 256, 0, stream>>>(in, in_dtype, s_in, s_out, d,
                                                                                          out, out_dtype, toks);
  return launched(err);
}

int window_rearrange_launch(const void* in, int in_dtype, int bs, int q, int r, int d, __nv_bfloat16* out,
                            cudaStream_t stream, const char** err) {
  const long long toks = static_cast<long long>(bs) * q * q * r * r;
  if (toks <= 0) return TDC_OK;
  if (d % 4 != 0) {
    if (err) *err = "window_rearrange: d must be a multiple of 4";
    return TDC_EINVAL;
  }
  window_rearrange_kernel<<<static_cast<unsigned>((toks + 7) / 8), 256, 0, stream>>>(in, in_dtype, q, r, d, out, toks);
  return launched(err);
}

int combine_parts_launch(const float* base, const float* parts, const float* logits, int ld_logits, int num_parts,
                         long long rows, int width, float* out, cudaStream_t stream, const char** err) {
  if (rows <= 0) return TDC_OK;
  if (num_parts < 1 || num_parts > 8 || width % 4 != 0) {
    if (err) *err = "combine_parts: 1..8 parts, width a multiple of 4";
    return TDC_EINVAL;
  }
  combine_parts_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(base, parts, logits, ld_logits, num_parts,
                                                                                 rows, width, out);
  return launched(err);
}

int compose_index_launch(const int32_t* a, int n_a, const int32_t* b, int32_t* out, long long n, cudaStream_t stream,
                         const char** err) {
  if (n <= 0) return TDC_OK;
  compose_index_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(a, b, out, n, n_a);
  return launched(err);
}

int broadcast_rows_launch(const float* src, int width, int slabs, __nv_bfloat16* dst, long long slab_stride,
                          long long row0, int count, cudaStream_t stream, const char** err) {
  if (count <= 0 || slabs <= 0) return TDC_OK;
  broadcast_rows_kernel<<<dim3(8, slabs), 256, 0, stream>>>(src, width, slabs, dst, slab_stride, row0, count);
  return launched(err);
}

}  // namespace tdc
