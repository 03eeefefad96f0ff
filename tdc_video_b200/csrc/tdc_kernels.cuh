// Launch interfaces of the non-GEMM kernels (attention.cu, rowops.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "tdc_b200.h"

namespace tdc {

// Token addressing shared by the kernels: the Q-Former's sequence of one row is
// [K query tokens | T text tokens] (tdc/Qformer.py:95-102).  The library stores the two
// kinds in two contiguous slabs — all query tokens of all rows first, then all text
// tokens — so that the query-only stages (cross-attention, query FFN; Qformer.py:430-454)
// and the text-only FFN (:455-462) are plain dense GEMMs over contiguous rows.
// Token i of row r lives at slab row
//     i <  seg1 ?  base1 + r*seg1 + i  :  base2 + r*seg2 + (i - seg1).
// KV tokens may have a third segment of kv_seg3 tokens SHARED by all rows (slab rows kv_base3 ...): the
// image_newline tokens of a frame, whose K/V are the same for every frame (cambrian_arch.py:1269-1281).
struct AttentionArgs {
  const __nv_bfloat16* q = nullptr;  // head h of a token at q + row*ldq + h*64
  const __nv_bfloat16* k = nullptr;
  const __nv_bfloat16* v = nullptr;
  __nv_bfloat16* out = nullptr;  // same token addressing as q
  long long ldq = 0, ldk = 0, ldv = 0, ldo = 0;
  // elements between the K (V) of consecutive heads: 64 = heads side by side inside a token row (the default), or a
  // whole [tokens, 64] matrix per head (the Q-Former's cross-attention K/V slabs: a (row, head) pair streams one
  // contiguous run of 128-byte lines)
  long long k_head_stride = 64, v_head_stride = 64;
  int rows = 0, heads = 0;
  int nq = 0;  // queries per row (= q_seg1 + q_seg2)
  int q_seg1 = 0, q_seg2 = 0;
  long long q_base1 = 0, q_base2 = 0;
  int kv_seg1 = 0, kv_seg2 = 0, kv_seg3 = 0;
  long long kv_base1 = 0, kv_base2 = 0, kv_base3 = 0;
  const int32_t* kv_len = nullptr;  // [rows] or null
  const uint32_t* kv_mask = nullptr;  // [rows] or null: bit j = KV token j may be attended (needs <= 32 KV tokens)
  float scale_log2 = 0.f;           // log2(e) / sqrt(head size)
};
int attention_launch(const AttentionArgs& a, cudaStream_t stream, const char** err);

// y = LayerNorm(x + resid) * gamma + beta, fp32 statistics (tdc/Qformer.py:285-289, 371-375).
// x, resid (nullable) fp32 [rows, width]; writes y as fp32 (nullable) and bf16 (nullable).
// y_f32 may alias resid (each warp reads its whole row before writing it).  resid_period > 0: the residual is a
// [resid_period, width] table added cyclically (row r uses table row r % resid_period) — the SVA position
// embeddings of the KV windows (tdc/vision_sampler.py:376-386).
int layernorm_launch(const float* x, long long ldx, const float* resid, long long ldr, const float* gamma,
                     const float* beta, float eps, float* y_f32, __nv_bfloat16* y_bf16, long long ldy, long long rows,
                     int width, cudaStream_t stream, const char** err, int resid_period = 0);

// out = a + b (fp32), optionally also as bf16 — the outer residual of the SVA layer (vision_sampler.py:399).
int residual_add_launch(const float* a, const float* b, float* out_f32, __nv_bfloat16* out_bf16, long long count,
                        cudaStream_t stream, const char** err);

// BertEmbeddings.forward (tdc/Qformer.py:78-108): query tokens = query_embeds (no position
// embedding), text tokens = word_emb[id] + pos_emb[t]; LayerNorm over everything.
struct EmbedArgs {
  const void* query_embeds = nullptr;  // [n_sets, K, H]
  int query_dtype = TDC_F32;
  const int32_t* query_set = nullptr;  // [rows] or null
  const int64_t* input_ids = nullptr;  // [n_text_sets, T] or null
  const int32_t* text_set = nullptr;   // [rows] or null
  int num_query_sets = 0, num_text_sets = 0;   // > 0: set indices are clamped into range (0: trusted)
  const float* word_emb = nullptr;     // [vocab, H]
  const float* pos_emb = nullptr;      // [max_pos, H]
  int vocab = 0;
  const float* gamma = nullptr;
  const float* beta = nullptr;
  float eps = 1e-12f;
  float* h_f32 = nullptr;  // [rows*K + rows*T, H] slab layout
  __nv_bfloat16* h_bf16 = nullptr;
  int rows = 0, num_query = 0, num_text = 0, hidden = 0;
};
int embed_layernorm_launch(const EmbedArgs& a, cudaStream_t stream, const char** err);

// out[r, i, :] = h[slab row of (r, i)] for i < tokens_out, converted to out_dtype.
int gather_rows_launch(const float* h_f32, int hidden, int rows, int num_query, int num_text, int tokens_out,
                       void* out, int out_dtype, cudaStream_t stream, const char** err);

// h (slab layout) <- sets[set_map[r]] (set-major [set][K + T][hidden] fp32; set_map NULL = set 0 for every row)
int broadcast_sets_launch(const float* sets, const int32_t* set_map, int num_sets, int rows, int num_query, int num_text,
                          int hidden, float* h_f32, __nv_bfloat16* h_bf16, cudaStream_t stream, const char** err);

// F.normalize(x, dim=-1) (eps 1e-12; tdc/cambrian_arch.py:1664-1667): x fp32 [rows, width] -> out_dtype.
// multicast: `out` is an NVSwitch multicast address (multimem.st): every GPU of the group gets the rows.
int l2_normalize_launch(const float* x, long long ldx, void* out, int out_dtype, long long rows, int width,
                        bool multicast, cudaStream_t stream, const char** err);

// Elementwise dtype conversion.
int convert_launch(const void* src, int src_dtype, void* dst, int dst_dtype, long long count, cudaStream_t stream,
                   const char** err);

// hidden [rows, tokens_per_row, H] (dtype) -> bf16 [rows*num_query, H] (first num_query tokens of every row)
int take_query_tokens_launch(const void* hidden, int dtype, int rows, int tokens_per_row, int num_query, int width,
                             __nv_bfloat16* out, cudaStream_t stream, const char** err);

// adaptive_avg_pool1d over the token axis: bins [floor(i*L/K), ceil((i+1)*L/K)) (tdc/cambrian_arch.py:1633-1637)
int avg_pool_tokens_launch(const void* frames, int dtype, int n, int tokens, int d, int num_query,
                           __nv_bfloat16* out, cudaStream_t stream, const char** err);

// ---- upstream ("frames") entry: frontend.cu -------------------------------------------------------------
// dst block i = src block idx[i] (blocks of block_bytes, a multiple of 16): frames gathered by role (static / dynamic)
// (indices are clamped to [0, n_src))
int gather_blocks_launch(const void* src, const int32_t* idx, void* dst, long long items, long long block_bytes,
                         int n_src, cudaStream_t stream, const char** err);
int transpose_bf16_launch(const __nv_bfloat16* in, int rows, int cols, __nv_bfloat16* out, cudaStream_t stream,
                          const char** err);
// y = W x + b (W bf16 [n, k]; x, b, y fp32) — weight folding at load time
int matvec_bias_launch(const __nv_bfloat16* w, const float* x, const float* b, float* y, int n, int k,
                       cudaStream_t stream, const char** err);
// adaptive_avg_pool1d over [side x (side visual tokens + newline)] of every key frame (cambrian_arch.py:1633-1637
// on the newline-extended frame of :1269-1281): xv [chunks, side*side, d] bf16 -> out [chunks, K, d] bf16
int pool_static_queries_launch(const __nv_bfloat16* xv, const float* newline, int chunks, int side, int d,
                               int num_query, __nv_bfloat16* out, cudaStream_t stream, const char** err);
// static_out[c] = [(side visual tokens, newline) x side | ta audio tokens] (the key frame as it passes through)
int multicast_copy_launch(const void* src, void* dst, size_t bytes, int ctas, cudaStream_t stream, const char** err);
// (multicast: `out` is an NVSwitch multicast address, see l2_normalize_launch)
int assemble_static_launch(const __nv_bfloat16* xv, const __nv_bfloat16* xa, const float* newline, int chunks, int side,
                           int ta, int d, void* out, int out_dtype, bool multicast, cudaStream_t stream,
                           const char** err);
// bilinear (align_corners = False) resize of a token grid, token-major: [bs, s_in^2, d] -> [bs, s_out^2, d]
int resize_tokens_bilinear_launch(const void* in, int in_dtype, int bs, int s_in, int s_out, int d, void* out,
                                  int out_dtype, cudaStream_t stream, const char** err);
// SVA connector: tokens under every query, window-major: in [bs, (q r)^2, d] -> out [bs, q, q, r, r, d] bf16
int window_rearrange_launch(const void* in, int in_dtype, int bs, int q, int r, int d, __nv_bfloat16* out,
                            cudaStream_t stream, const char** err);
// out = base + sum_t softmax(logits[:, :T])[t] * parts[t]   (fp32 [rows, width]; parts [T, rows, width])
int combine_parts_launch(const float* base, const float* parts, const float* logits, int ld_logits, int num_parts,
                         long long rows, int width, float* out, cudaStream_t stream, const char** err);
// out[i] = a[clamp(b[i], 0, n_a - 1)]
int compose_index_launch(const int32_t* a, int n_a, const int32_t* b, int32_t* out, long long n, cudaStream_t stream,
                         const char** err);
// rows [row0, row0 + count) of each of `slabs` matrices [*, width] (slab_stride elements apart) = src[slab] as bf16
int broadcast_rows_launch(const float* src, int width, int slabs, __nv_bfloat16* dst, long long slab_stride,
                          long long row0, int count, cudaStream_t stream, const char** err);

// Adaptive segmentation (tdc/cambrian_arch.py:832-849): cos[i] = cosine_similarity(frame i, frame i+1) over
// the flattened feature dim; partial_ws holds (n_frames-1) * frame_cosine_slices(dim) * 3 floats.
inline int frame_cosine_slices(long long dim) { return dim >= (1 << 16) ? 8 : 1; }
int frame_cosine_launch(const void* feats, int dtype, int n_frames, long long dim, float* partial_ws, float* cos,
                        cudaStream_t stream, const char** err);
// out[0..min(k,n)) = indices of the k smallest x, ascending by index (== sort(argsort(x)[:k]))
int select_smallest_launch(const float* x, int n, int k, long long* out, cudaStream_t stream, const char** err);

}  // namespace tdc
