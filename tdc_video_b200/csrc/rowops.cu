// Row-wise (HBM-bound) kernels of the TDC path: LayerNorm, embeddings + LayerNorm,
// L2-normalise, adaptive average pooling over tokens, dtype conversion, gathers.
// All statistics are fp32; all global accesses are 8- or 16-byte vectors, one warp per
// row so a row's bytes are read exactly once and coalesced.
#include "tdc_kernels.cuh"
#include "tdc_ptx.cuh"

#include <cuda_fp16.h>

namespace tdc {

namespace {

constexpr int kMaxVec = 12;  // float4 per lane cached in registers: width <= 1536 (Whisper-large features are 1280 wide)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float4 load4_as_f32(const void* base, int dtype, long long idx4) {
  // idx4 counts groups of 4 elements
  if (dtype == TDC_F32) return __ldg(reinterpret_cast<const float4*>(base) + idx4);
  const uint2 raw = __ldg(reinterpret_cast<const uint2*>(base) + idx4);
  float4 r;
  if (dtype == TDC_BF16) {
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
    r.x = __low2float(a); r.y = __high2float(a); r.z = __low2float(b); r.w = __high2float(b);
  } else {
    const __half2 a = *reinterpret_cast<const __half2*>(&raw.x);
    const __half2 b = *reinterpret_cast<const __half2*>(&raw.y);
    r.x = __low2float(a); r.y = __high2float(a); r.z = __low2float(b); r.w = __high2float(b);
  }
  return r;
}

__device__ __forceinline__ void store4_from_f32(void* base, int dtype, long long idx4, float4 v) {
  if (dtype == TDC_F32) {
    reinterpret_cast<float4*>(base)[idx4] = v;
  } else if (dtype == TDC_BF16) {
    uint2 raw;
    raw.x = pack_bf16x2(v.x, v.y);
    raw.y = pack_bf16x2(v.z, v.w);
    reinterpret_cast<uint2*>(base)[idx4] = raw;
  } else {
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 raw;
    raw.x = *reinterpret_cast<const uint32_t*>(&a);
    raw.y = *reinterpret_cast<const uint32_t*>(&b);
    reinterpret_cast<uint2*>(base)[idx4] = raw;
  }
}

// Normalise the row held in x[] (nv float4 per lane), write fp32 and/or bf16.
__device__ __forceinline__ void ln_finish(float4 (&x)[kMaxVec], int nv, int lane, int width, const float* gamma,
                                          const float* beta, float eps, float* y_f32, __nv_bfloat16* y_bf16) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j)
    if (j < nv && lane + 32 * j < width / 4) s += x[j].x + x[j].y + x[j].z + x[j].w;
  const float mean = warp_sum(s) / static_cast<float>(width);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j)
    if (j < nv && lane + 32 * j < width / 4) {
      const float a = x[j].x - mean, b = x[j].y - mean, c = x[j].z - mean, d = x[j].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
  const float var = warp_sum(q) / static_cast<float>(width);  // biased, as nn.LayerNorm
  const float rstd = 1.0f / sqrtf(var + eps);
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j)
    if (j < nv && lane + 32 * j < width / 4) {
      const int i4 = lane + 32 * j;
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + i4);
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + i4);
      float4 y;
      y.x = (x[j].x - mean) * rstd * gm.x + bt.x;
      y.y = (x[j].y - mean) * rstd * gm.y + bt.y;
      y.z = (x[j].z - mean) * rstd * gm.z + bt.z;
      y.w = (x[j].w - mean) * rstd * gm.w + bt.w;
      if (y_f32 != nullptr) reinterpret_cast<float4*>(y_f32)[i4] = y;
      if (y_bf16 != nullptr) {
        uint2 raw;
        raw.x = pack_bf16x2(y.x, y.y);
        raw.y = pack_bf16x2(y.z, y.w);
        reinterpret_cast<uint2*>(y_bf16)[i4] = raw;
      }
    }
}

__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, long long ldx,
                                                        const float* resid, long long ldr,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps,
                                                        float* y_f32, __nv_bfloat16* __restrict__ y_bf16,
                                                        long long ldy, long long rows, int width, int resid_period) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nv = (width / 4 + 31) / 32;
  float4 v[kMaxVec];
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j)
    if (j < nv && lane + 32 * j < width / 4) v[j] = xr[lane + 32 * j];
  if (resid != nullptr) {  // post-LN residual: LN(dense(x) + input), Qformer.py:285-289 / 371-375
    const long long rrow = resid_period > 0 ? row % resid_period : row;
    const float4* rr = reinterpret_cast<const float4*>(resid + rrow * ldr);
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j)
      if (j < nv && lane + 32 * j < width / 4) {
        const float4 r = rr[lane + 32 * j];
        v[j].x += r.x; v[j].y += r.y; v[j].z += r.z; v[j].w += r.w;
      }
  }
  ln_finish(v, nv, lane, width, gamma, beta, eps, y_f32 ? y_f32 + row * ldy : nullptr,
            y_bf16 ? y_bf16 + row * ldy : nullptr);
}

__global__ void __launch_bounds__(256) embed_layernorm_kernel(EmbedArgs a) {
  const int n = a.num_query + a.num_text;
  const long long tok = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (tok >= static_cast<long long>(a.rows) * n) return;
  const int lane = threadIdx.x & 31;
  const int r = static_cast<int>(tok / n), i = static_cast<int>(tok % n);
  const int width = a.hidden;
  const int nv = (width / 4 + 31) / 32;
  float4 v[kMaxVec];
  long long dst_row;
  if (i < a.num_query) {
    long long set = a.query_set ? a.query_set[r] : r;
    if (a.num_query_sets > 0) set = set < 0 ? 0 : (set >= a.num_query_sets ? a.num_query_sets - 1 : set);
    const long long base4 = (set * a.num_query + i) * (width / 4);
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j)
      if (j < nv && lane + 32 * j < width / 4) v[j] = load4_as_f32(a.query_embeds, a.query_dtype, base4 + lane + 32 * j);
    dst_row = static_cast<long long>(r) * a.num_query + i;
  } else {
    const int t = i - a.num_query;
    long long set = a.text_set ? a.text_set[r] : r;
    if (a.num_text_sets > 0) set = set < 0 ? 0 : (set >= a.num_text_sets ? a.num_text_sets - 1 : set);
    long long id = a.input_ids[set * a.num_text + t];
    id = id < 0 ? 0 : (id >= a.vocab ? a.vocab - 1 : id);
    const float4* we = reinterpret_cast<const float4*>(a.word_emb + id * width);
    const float4* pe = reinterpret_cast<const float4*>(a.pos_emb + static_cast<long long>(t) * width);
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j)
      if (j < nv && lane + 32 * j < width / 4) {
        const float4 w = __ldg(we + lane + 32 * j), p = __ldg(pe + lane + 32 * j);
        v[j] = make_float4(w.x + p.x, w.y + p.y, w.z + p.z, w.w + p.w);
      }
    dst_row = static_cast<long long>(a.rows) * a.num_query + static_cast<long long>(r) * a.num_text + t;
  }
  ln_finish(v, nv, lane, width, a.gamma, a.beta, a.eps, a.h_f32 + dst_row * width, a.h_bf16 + dst_row * width);
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ h, int hidden, int rows,
                                                          int num_query, int num_text, int tokens_out,
                                                          void* __restrict__ out, int out_dtype) {
  const long long tok = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (tok >= static_cast<long long>(rows) * tokens_out) return;
  const int lane = threadIdx.x & 31;
  const int r = static_cast<int>(tok / tokens_out), i = static_cast<int>(tok % tokens_out);
  const long long src = (i < num_query) ? static_cast<long long>(r) * num_query + i
                                        : static_cast<long long>(rows) * num_query +
                                              static_cast<long long>(r) * num_text + (i - num_query);
  const float4* s = reinterpret_cast<const float4*>(h + src * hidden);
  for (int j = lane; j < hidden / 4; j += 32) store4_from_f32(out, out_dtype, tok * (hidden / 4) + j, s[j]);
}

// h (slab layout of `rows` rows) <- sets[set_map[r]] (set-major [set][K + T][hidden], fp32): the layer-0 state
// computed once per (query set, prompt) broadcast to every row that shares it; also writes the bf16 copy.
__global__ void __launch_bounds__(256) broadcast_sets_kernel(const float* __restrict__ sets,
                                                             const int32_t* __restrict__ set_map, int num_sets, int rows,
                                                             int num_query, int num_text, int hidden,
                                                             float* __restrict__ h_f32,
                                                             __nv_bfloat16* __restrict__ h_bf16) {
  const int n = num_query + num_text;
  const long long tok = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (tok >= static_cast<long long>(rows) * n) return;
  const int lane = threadIdx.x & 31;
  const int r = static_cast<int>(tok / n), i = static_cast<int>(tok % n);
  long long set = set_map ? set_map[r] : 0;
  set = set < 0 ? 0 : (set >= num_sets ? num_sets - 1 : set);
  const float4* src = reinterpret_cast<const float4*>(sets + (set * n + i) * hidden);
  const long long dst = (i < num_query) ? static_cast<long long>(r) * num_query + i
                                        : static_cast<long long>(rows) * num_query +
                                              static_cast<long long>(r) * num_text + (i - num_query);
  for (int j = lane; j < hidden / 4; j += 32) {
    const float4 v = __ldg(src + j);
    reinterpret_cast<float4*>(h_f32 + dst * hidden)[j] = v;
    uint2 raw;
    raw.x = pack_bf16x2(v.x, v.y);
    raw.y = pack_bf16x2(v.z, v.w);
    reinterpret_cast<uint2*>(h_bf16 + dst * hidden)[j] = raw;
  }
}

// 16-byte store; kMulticast: `p` is an NVSwitch multicast address (all GPUs of the group receive the
// bytes with one store — the all-gather of the compressed tokens rides on this kernel's output).
template <bool kMulticast>
__device__ __forceinline__ void store16(void* p, uint4 v) {
  if (kMulticast)
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
  else
    *reinterpret_cast<uint4*>(p) = v;
}

template <bool kMulticast>
__global__ void __launch_bounds__(256) l2_normalize_kernel(const float* __restrict__ x, long long ldx,
                                                           void* __restrict__ out, int out_dtype, long long rows,
                                                           int width) {
  // one 256-thread block per row (width is the LLM hidden size, 3-4 K); each thread keeps its 8-element
  // pieces in registers (up to kCache pieces = width <= 8192) so the row is read from HBM once
  constexpr int kCache = 4;
  __shared__ float red[8];
  const long long row = blockIdx.x;
  const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
  const int pieces = width / 8;
  float4 ca[kCache], cb[kCache];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kCache; ++i) {
    const int j = threadIdx.x + i * 256;
    if (j < pieces) {
      ca[i] = xr[2 * j];
      cb[i] = xr[2 * j + 1];
      s += ca[i].x * ca[i].x + ca[i].y * ca[i].y + ca[i].z * ca[i].z + ca[i].w * ca[i].w;
      s += cb[i].x * cb[i].x + cb[i].y * cb[i].y + cb[i].z * cb[i].z + cb[i].w * cb[i].w;
    }
  }
  for (int j = threadIdx.x + kCache * 256; j < pieces; j += 256) {  // very wide rows: not cached
    const float4 a = xr[2 * j], b = xr[2 * j + 1];
    s += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float inv = 1.0f / fmaxf(sqrtf(tot), 1e-12f);  // F.normalize: x / max(||x||, eps)
  uint8_t* orow = static_cast<uint8_t*>(out) + row * width * (out_dtype == TDC_F32 ? 4 : 2);
  auto emit = [&](int j, float4 a, float4 b) {
    a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
    b.x *= inv; b.y *= inv; b.z *= inv; b.w *= inv;
    if (out_dtype == TDC_F32) {
      store16<kMulticast>(orow + j * 32, *reinterpret_cast<uint4*>(&a));
      store16<kMulticast>(orow + j * 32 + 16, *reinterpret_cast<uint4*>(&b));
    } else {
      uint4 pk;
      if (out_dtype == TDC_BF16) {
        pk.x = pack_bf16x2(a.x, a.y); pk.y = pack_bf16x2(a.z, a.w);
        pk.z = pack_bf16x2(b.x, b.y); pk.w = pack_bf16x2(b.z, b.w);
      } else {
        const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w);
        const __half2 h2 = __floats2half2_rn(b.x, b.y), h3 = __floats2half2_rn(b.z, b.w);
        pk.x = *reinterpret_cast<const uint32_t*>(&h0); pk.y = *reinterpret_cast<const uint32_t*>(&h1);
        pk.z = *reinterpret_cast<const uint32_t*>(&h2); pk.w = *reinterpret_cast<const uint32_t*>(&h3);
      }
      store16<kMulticast>(orow + j * 16, pk);
    }
  };
#pragma unroll
  for (int i = 0; i < kCache; ++i) {
    const int j = threadIdx.x + i * 256;
    if (j < pieces) emit(j, ca[i], cb[i]);
  }
  for (int j = threadIdx.x + kCache * 256; j < pieces; j += 256) emit(j, xr[2 * j], xr[2 * j + 1]);
}

__global__ void __launch_bounds__(256) convert_kernel(const void* __restrict__ src, int src_dtype,
                                                      void* __restrict__ dst, int dst_dtype, long long count4) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < count4; i += stride)
    store4_from_f32(dst, dst_dtype, i, load4_as_f32(src, src_dtype, i));
}

__global__ void __launch_bounds__(256) residual_add_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                           float* __restrict__ out_f32,
                                                           __nv_bfloat16* __restrict__ out_bf16, long long count4) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < count4; i += stride) {
    const float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
    const float4 z = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    if (out_f32 != nullptr) reinterpret_cast<float4*>(out_f32)[i] = z;
    if (out_bf16 != nullptr) store4_from_f32(out_bf16, TDC_BF16, i, z);
  }
}

__global__ void __launch_bounds__(256) take_query_tokens_kernel(const void* __restrict__ hidden, int dtype, int rows,
                                                                int tokens_per_row, int num_query, int width,
                                                                __nv_bfloat16* __restrict__ out) {
  const long long tok = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (tok >= static_cast<long long>(rows) * num_query) return;
  const int lane = threadIdx.x & 31;
  const long long r = tok / num_query, i = tok % num_query;
  const long long src4 = (r * tokens_per_row + i) * (width / 4);
  for (int j = lane; j < width / 4; j += 32)
    store4_from_f32(out, TDC_BF16, tok * (width / 4) + j, load4_as_f32(hidden, dtype, src4 + j));
}

__global__ void __launch_bounds__(256) avg_pool_tokens_kernel(const void* __restrict__ frames, int dtype, int tokens,
                                                              int d, int num_query, __nv_bfloat16* __restrict__ out) {
  // block = (frame, bin); threads sweep the feature axis
  const int f = blockIdx.x / num_query, i = blockIdx.x % num_query;
  const int start = static_cast<int>((static_cast<long long>(i) * tokens) / num_query);
  const int end = static_cast<int>((static_cast<long long>(i + 1) * tokens + num_query - 1) / num_query);
  const float inv = 1.0f / static_cast<float>(end - start);
  for (int j = threadIdx.x; j < d / 4; j += 256) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = start; t < end; ++t) {
      const float4 v = load4_as_f32(frames, dtype, (static_cast<long long>(f) * tokens + t) * (d / 4) + j);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
    store4_from_f32(out, TDC_BF16, (static_cast<long long>(f) * num_query + i) * (d / 4) + j, acc);
  }
}

int check_launch(const char** err) {
  const cudaError_t rc = cudaGetLastError();
  if (rc != cudaSuccess) {
    if (err) *err = cudaGetErrorString(rc);
    return TDC_ECUDA;
  }
  return TDC_OK;
}

}  // namespace

int layernorm_launch(const float* x, long long ldx, const float* resid, long long ldr, const float* gamma,
                     const float* beta, float eps, float* y_f32, __nv_bfloat16* y_bf16, long long ldy, long long rows,
                     int width, cudaStream_t stream, const char** err, int resid_period) {
  if (rows <= 0) return TDC_OK;
  if (width % 4 != 0 || width > kMaxVec * 128 || ldx % 4 != 0 || ldy % 4 != 0 || ldr % 4 != 0) {
    if (err) *err = "layernorm: width must be a multiple of 4 and <= 1536";
    return TDC_EINVAL;
  }
  layernorm_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(x, ldx, resid, ldr, gamma, beta, eps,
                                                                              y_f32, y_bf16, ldy, rows, width,
                                                                              resid_period);
  return check_launch(err);
}

int embed_layernorm_launch(const EmbedArgs& a, cudaStream_t stream, const char** err) {
  if (a.rows <= 0) return TDC_OK;
  if (a.hidden % 4 != 0 || a.hidden > kMaxVec * 128) {
    if (err) *err = "embeddings: hidden must be a multiple of 4 and <= 1536";
    return TDC_EINVAL;
  }
  const long long toks = static_cast<long long>(a.rows) * (a.num_query + a.num_text);
  embed_layernorm_kernel<<<static_cast<unsigned>((toks + 7) / 8), 256, 0, stream>>>(a);
  return check_launch(err);
}

int gather_rows_launch(const float* h_f32, int hidden, int rows, int num_query, int num_text, int tokens_out,
                       void* out, int out_dtype, cudaStream_t stream, const char** err) {
  const long long toks = static_cast<long long>(rows) * tokens_out;
  if (toks <= 0) return TDC_OK;
  gather_rows_kernel<<<static_cast<unsigned>((toks + 7) / 8), 256, 0, stream>>>(h_f32, hidden, rows, num_query,
                                                                                num_text, tokens_out, out, out_dtype);
  return check_launch(err);
}

int broadcast_sets_launch(const float* sets, const int32_t* set_map, int num_sets, int rows, int num_query, int num_text,
                          int hidden, float* h_f32, __nv_bfloat16* h_bf16, cudaStream_t stream, const char** err) {
  const long long toks = static_cast<long long>(rows) * (num_query + num_text);
  if (toks <= 0) return TDC_OK;
  broadcast_sets_kernel<<<static_cast<unsigned>((toks + 7) / 8), 256, 0, stream>>>(sets, set_map, num_sets, rows,
                                                                                   num_query, num_text, hidden, h_f32,
                                                                                   h_bf16);
  return check_launch(err);
}

int l2_normalize_launch(const float* x, long long ldx, void* out, int out_dtype, long long rows, int width,
                        bool multicast, cudaStream_t stream, const char** err) {
  if (rows <= 0) return TDC_OK;
  if (width % 8 != 0 || ldx % 4 != 0 || (reinterpret_cast<uintptr_t>(out) & 15)) {
    if (err) *err = "l2_normalize: width must be a multiple of 8 and the output 16-byte aligned";
    return TDC_EINVAL;
  }
  if (multicast)
    l2_normalize_kernel<true><<<static_cast<unsigned>(rows), 256, 0, stream>>>(x, ldx, out, out_dtype, rows, width);
  else
    l2_normalize_kernel<false><<<static_cast<unsigned>(rows), 256, 0, stream>>>(x, ldx, out, out_dtype, rows, width);
  return check_launch(err);
}

int convert_launch(const void* src, int src_dtype, void* dst, int dst_dtype, long long count, cudaStream_t stream,
                   const char** err) {
  if (count <= 0) return TDC_OK;
  if (count % 4 != 0) {
    if (err) *err = "convert: element count must be a multiple of 4";
    return TDC_EINVAL;
  }
  long long blocks = (count / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  convert_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(src, src_dtype, dst, dst_dtype, count / 4);
  return check_launch(err);
}

int residual_add_launch(const float* a, const float* b, float* out_f32, __nv_bfloat16* out_bf16, long long count,
                        cudaStream_t stream, const char** err) {
  if (count <= 0) return TDC_OK;
  if (count % 4 != 0) {
    if (err) *err = "residual_add: element count must be a multiple of 4";
    return TDC_EINVAL;
  }
  long long blocks = (count / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  residual_add_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(a, b, out_f32, out_bf16, count / 4);
  return check_launch(err);
}

int take_query_tokens_launch(const void* hidden, int dtype, int rows, int tokens_per_row, int num_query, int width,
                             __nv_bfloat16* out, cudaStream_t stream, const char** err) {
  const long long toks = static_cast<long long>(rows) * num_query;
  if (toks <= 0) return TDC_OK;
  if (width % 4 != 0) {
    if (err) *err = "take_query_tokens: width must be a multiple of 4";
    return TDC_EINVAL;
  }
  take_query_tokens_kernel<<<static_cast<unsigned>((toks + 7) / 8), 256, 0, stream>>>(hidden, dtype, rows,
                                                                                      tokens_per_row, num_query, width,
                                                                                      out);
  return check_launch(err);
}

int avg_pool_tokens_launch(const void* frames, int dtype, int n, int tokens, int d, int num_query,
                           __nv_bfloat16* out, cudaStream_t stream, const char** err) {
  if (n <= 0) return TDC_OK;
  if (d % 4 != 0 || num_query <= 0 || tokens <= 0) {
    if (err) *err = "avg_pool_tokens: d must be a multiple of 4";
    return TDC_EINVAL;
  }
  avg_pool_tokens_kernel<<<static_cast<unsigned>(n * num_query), 256, 0, stream>>>(frames, dtype, tokens, d,
                                                                                   num_query, out);
  return check_launch(err);
}

}  // namespace tdc
