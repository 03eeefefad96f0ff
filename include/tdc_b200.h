/*
 * tdc_b200.h — C ABI of libtdc_b200.so: hand-written sm_100a CUDA kernels for the
 * Temporal Dynamic Context (TDC) compression path of Hoar012/TDC-Video.
 *
 * The reference has no FFI layer of its own: its boundary for this path is a set of
 * nn.Module attributes called from tdc/cambrian_arch.py.  Each entry point below
 * names the reference call it stands in for (file:line under the reference repo);
 * INTEGRATION.md shows the Python-side binding (ctypes) a maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - all calls are stream-ordered and asynchronous, never synchronise, and never
 *     allocate after tdc_create / tdc_load_weights (graph-capturable);
 *   - the caller owns inputs, outputs and the workspace; the handle owns only its
 *     re-packed copy of the weights;
 *   - return value: TDC_OK (0) or a negative tdc_status; tdc_last_error() gives text.
 *   - eval-mode only (dropout = identity), exactly like every reference call site
 *     on the inference path.
 */
#ifndef TDC_B200_H_
#define TDC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDC_B200_ABI_VERSION 2

typedef enum tdc_status {
  TDC_OK = 0,
  TDC_EINVAL = -1,   /* bad shape / alignment / null pointer / unsupported geometry */
  TDC_ECUDA = -2,    /* a CUDA runtime or driver call failed */
  TDC_ENOMEM = -3,   /* device allocation failed (create / load_weights only) */
  TDC_ESTATE = -4,   /* call made in the wrong state (e.g. forward before load_weights) */
  TDC_EWORKSPACE = -5 /* workspace smaller than tdc_workspace_bytes() */
} tdc_status;

typedef enum tdc_dtype {
  TDC_BF16 = 0,
  TDC_F16 = 1,
  TDC_F32 = 2
} tdc_dtype;

typedef void* tdc_stream_t; /* cudaStream_t */
typedef struct tdc_handle tdc_handle;

/* Geometry of the Q-Former + projections.  Mirrors the BertConfig fields the
 * reference sets in tdc/cambrian_arch.py:403-424 (init_Qformer) plus the widths of
 * vision_proj / query_proj (cambrian_arch.py:483-484). */
typedef struct tdc_config {
  int32_t hidden;            /* BertConfig.hidden_size (768); must be heads*64 */
  int32_t heads;             /* num_attention_heads (12); head size is fixed at 64 */
  int32_t intermediate;      /* intermediate_size (3072) */
  int32_t layers;            /* num_hidden_layers (12) */
  int32_t cross_freq;        /* cross_attention_freq (2): cross-attention in layers i % cross_freq == 0 */
  int32_t d_enc;             /* encoder_width: width of the frame / segment tokens */
  int32_t d_out;             /* vision_proj out_features (LLM hidden size); 0 = no vision_proj */
  int32_t vocab;             /* word-embedding rows (30522); 0 = text input unsupported */
  int32_t max_pos;           /* position-embedding rows (512) */
  float ln_eps;              /* layer_norm_eps (1e-12) */
  int32_t gemm_cta_group;    /* 0 = default, 1 = single-CTA tcgen05 tiles, 2 = CTA-pair tiles */
  int32_t d_frame_in;        /* mm_projector.0 in_features (1024 * towers); 0 = no upstream ("frames") entry */
  int32_t d_audio;           /* audio_proj in_features (768, BEATs); 0 = no audio */
  int32_t reserved[3];
} tdc_config;

/* One reference tensor (as stored in the reference state_dict, row-major). */
typedef struct tdc_tensor {
  const char* name;   /* state_dict key relative to `Qformer.bert.` or a sibling name, see DESIGN.md */
  const void* data;   /* device pointer */
  int32_t dtype;      /* tdc_dtype */
  int32_t ndim;
  int64_t shape[4];
} tdc_tensor;

/* ---- lifecycle ------------------------------------------------------------------ */
/* replaces: BertLMHeadModel(config) construction, cambrian_arch.py:403-424 */
int tdc_create(tdc_handle** out, const tdc_config* cfg);
int tdc_destroy(tdc_handle* h);
const char* tdc_last_error(const tdc_handle* h); /* h may be NULL: last error of failed create */
int tdc_abi_version(void);

/* replaces: load_state_dict of `model.Qformer.bert.*`, `model.vision_proj.*`
 * (reference checkpoint layout).  Tensors are converted to the
 * library's own bf16 / fp32 buffers on `stream`; missing optional tensors
 * (text FFN, embeddings, vision_proj) disable the corresponding feature. */
int tdc_load_weights(tdc_handle* h, const tdc_tensor* tensors, int32_t count, tdc_stream_t stream);

/* ---- the hot path --------------------------------------------------------------- */
/* Bytes of workspace tdc_qformer_forward / tdc_compress need for `rows` rows. */
size_t tdc_workspace_bytes(const tdc_handle* h, int32_t rows, int32_t kv_len, int32_t num_query, int32_t num_text);

/* replaces: Qformer.bert(input_ids=, query_embeds=, encoder_hidden_states=,
 *           encoder_attention_mask=, use_cache=False, return_dict=True).last_hidden_state
 *           — tdc/cambrian_arch.py:1653-1662, tdc/Qformer.py:804-965.
 *   query_embeds   [n_query_sets, K, hidden]   (query_dtype)
 *   query_set      [rows] int32 or NULL        row -> query set (NULL: row r uses set r)
 *   input_ids      [n_text_sets, T] int64 or NULL (T == 0)
 *   text_set       [rows] int32 or NULL        row -> text set (NULL: row r uses set r)
 *   enc            [rows, L, d_enc]            (enc_dtype) encoder_hidden_states
 *   kv_len         [rows] int32 or NULL        valid KV tokens per row (NULL: all L; the
 *                                              reference always passes all-ones masks)
 *   out_hidden     [rows, K+T, hidden]         (out_dtype) last_hidden_state
 */
int tdc_qformer_forward(tdc_handle* h, const void* query_embeds, int32_t query_dtype, const int32_t* query_set,
                        const int64_t* input_ids, const int32_t* text_set, const void* enc, int32_t enc_dtype,
                        const int32_t* kv_len, int32_t rows, int32_t kv_tokens, int32_t num_query, int32_t num_text,
                        void* out_hidden, int32_t out_dtype, void* workspace, size_t workspace_bytes,
                        tdc_stream_t stream);

/* replaces: F.normalize(vision_proj(last_hidden_state[:, :K]), dim=-1)
 *           — tdc/cambrian_arch.py:1664-1667.
 *   hidden [rows, tokens_per_row, hidden] (hidden_dtype); the first num_query tokens of
 *   every row are projected; out [rows, num_query, d_out] (out_dtype), unit L2 norm. */
int tdc_proj_norm(tdc_handle* h, const void* hidden, int32_t hidden_dtype, int32_t rows, int32_t tokens_per_row,
                  int32_t num_query, void* out, int32_t out_dtype, void* workspace, size_t workspace_bytes,
                  tdc_stream_t stream);

/* Fused convenience entry: Q-Former + vision_proj + L2-normalise for all rows of a
 * video at once (the reference runs <= 7 rows per call from a Python loop,
 * cambrian_arch.py:1603-1692).  Arguments as tdc_qformer_forward; out is
 * [rows, K, d_out] (out_dtype).  Only the query tokens of the last layer are read
 * (`[:, :K]`, :1665), so the text tokens' out-projection / feed-forward of that layer,
 * which nothing consumes, are not computed. */
int tdc_compress(tdc_handle* h, const void* query_embeds, int32_t query_dtype, const int32_t* query_set,
                 const int64_t* input_ids, const int32_t* text_set, const void* enc, int32_t enc_dtype,
                 const int32_t* kv_len, int32_t rows, int32_t kv_tokens, int32_t num_query, int32_t num_text,
                 void* out, int32_t out_dtype, void* workspace, size_t workspace_bytes, tdc_stream_t stream);

/* tdc_compress whose result is written through an NVSwitch MULTICAST address (multimem.st): `out_multicast`
 * is this rank's slot inside a symmetric buffer mapped with a multicast VA (e.g. torch symmetric memory's
 * multicast_ptr + rank offset), so the final L2-normalise kernel delivers every row to all GPUs of the group
 * at once — the path's only exchange step (all-gather of the compressed tokens) fused into
 * the producing kernel.  The caller synchronises the group (barrier) before reading peers' rows. */
int tdc_compress_multicast(tdc_handle* h, const void* query_embeds, int32_t query_dtype, const int32_t* query_set,
                           const int64_t* input_ids, const int32_t* text_set, const void* enc, int32_t enc_dtype,
                           const int32_t* kv_len, int32_t rows, int32_t kv_tokens, int32_t num_query,
                           int32_t num_text, void* out_multicast, int32_t out_dtype, void* workspace,
                           size_t workspace_bytes, tdc_stream_t stream);

/* ---- the upstream entry: from the towers' outputs --------------------------------- */
/* replaces, for ALL chunks of a video in one call (tdc/cambrian_arch.py):
 *   :1149-1150  image_features = mm_projector(cat(tower features))            (Linear . GELU . Linear)
 *   :1269-1281  image_newline appended to every row of the token grid
 *   :1611-1614  audio_proj(audio_chunk_feature) concatenated to every frame's tokens
 *   :1629-1640  query_tokens = query_proj(adaptive_avg_pool1d(key_frame)) | the learned query_tokens
 *   :1653-1667  Qformer.bert(...) -> F.normalize(vision_proj(h[:, :K]))
 * needs the handle created with d_frame_in > 0 (and d_audio > 0 for audio), d_enc == d_out, and the extra
 * tensors `mm_projector.{0,2}.{weight,bias}`, `image_newline`, `query_proj.{weight,bias}`,
 * [`audio_proj.{weight,bias}`], [`query_tokens`] passed to tdc_load_weights.
 *
 * A chunk is <= 8 consecutive frames of one segment (:1606): its first frame is the key ("static") frame, the
 * others are rows.  The caller passes the integer plan (the reference builds it in its Python loop).
 *   fold = 1: dynamic frames use weights folded at load time (K/V weights x mm_projector.2 / audio_proj; the
 *             newline tokens' K/V are constants), so their d_llm-wide tokens are never materialised;
 *   fold = 0: every projection runs as the reference orders them.
 * Cross-attention does not depend on the order of its KV tokens, so both modes keep a row's tokens as
 * [visual | audio | newline] instead of the reference's row-interleaved newlines. */
typedef struct tdc_frames_args {
  const void* frames;            /* [n_frames, Tv, d_frame_in] bf16: input of mm_projector; Tv = side*side */
  const void* audio;             /* [n_frames, Ta, d_audio] bf16 per-frame audio tokens, or NULL (Ta = 0) */
  const int32_t* static_frames;  /* [n_chunks] frame index of every chunk's key frame */
  const int32_t* row_frames;     /* [rows]     frame index of every dynamic frame */
  const int32_t* row_chunk;      /* [rows]     chunk of every row (its query set) */
  const int64_t* input_ids;      /* [n_prompts, num_text] BERT ids of the prompt(s), or NULL */
  int32_t n_frames, n_chunks, rows;
  int32_t visual_tokens;         /* Tv (144) */
  int32_t audio_tokens;          /* Ta (50) or 0 */
  int32_t num_query;             /* K */
  int32_t num_text;              /* T */
  int32_t learned_queries;       /* 1: queries = loaded `query_tokens` (query_type "learned") */
  int32_t fold;
  int32_t multicast;             /* 1: `out` is an NVSwitch multicast address (see tdc_compress_multicast) */
  int32_t out_dtype;             /* dtype of out and static_out */
  int32_t no_layer0_dedup;       /* 0 (default): all rows of a chunk share their queries and the prompt
                                  * (cambrian_arch.py:1629-1646), so embeddings + the self-attention block of layer 0
                                  * are computed once per chunk and broadcast to its rows (bit-identical);
                                  * 1: computed per row */
  void* static_out;              /* [n_chunks, side*(side+1) + Ta, d_out] the key frames as they pass through, or NULL */
  void* out;                     /* [rows, K, d_out] compressed tokens */
  const int32_t* chunk_prompt;   /* [n_chunks] prompt (row of input_ids) of every chunk — several videos with their own
                                  * questions in one call (the eval loops run many clips concurrently); NULL: prompt 0 */
  int32_t n_prompts;             /* rows of input_ids (0 or 1: a single prompt).  All index arrays are DEVICE data the
                                  * library cannot validate on the host: frame / chunk / prompt indices are clamped into
                                  * range on the device instead of being trusted */
  int32_t static_multicast;      /* 1: `static_out` is an NVSwitch multicast address: the key frames' pass-through tokens are
                                  * delivered to every GPU of the group by the kernel that assembles them */
  void* static_ready_event;      /* cudaEvent_t or NULL: recorded on `stream` as soon as `static_out` is complete (before the
                                  * dynamic frames are processed), so that a caller can ship the key frames' tokens on another
                                  * stream — e.g. tdc_multicast_copy — while the rows are being compressed */
} tdc_frames_args;

/* Workspace for tdc_compress_frames processing `batch` rows (and key frames) at a time; any size from
 * batch = 1 upwards works — the call sizes its internal batches to what it is given. */
size_t tdc_frames_workspace_bytes(const tdc_handle* h, int32_t n_chunks, int32_t rows, int32_t batch,
                                  int32_t visual_tokens, int32_t audio_tokens, int32_t num_query, int32_t num_text);
int tdc_compress_frames(tdc_handle* h, const tdc_frames_args* args, void* workspace, size_t workspace_bytes,
                        tdc_stream_t stream);

/* Copy `bytes` (a multiple of 16; both pointers 16-byte aligned) from local device memory to an NVSwitch multicast
 * address with multimem.st: every GPU of the multicast group receives the bytes at the same offset of its buffer.
 * `ctas` bounds the grid (0: 16) so that the copy can share the GPU with compute on another stream.  This is the
 * exchange step of the sharded path (the key frames' tokens "ride the same all-gather") without a
 * collective library call; the caller closes it with its group barrier.  No reference counterpart (the reference has no
 * intra-video parallelism, eval/eval_mlvu.py:129-156). */
int tdc_multicast_copy(const void* src, void* dst_multicast, size_t bytes, int32_t ctas, tdc_stream_t stream);

/* Stream-ordered copy by the GPU's copy engines (cudaMemcpyAsync, device to device / peer): `dst` may be a peer GPU's
 * mapping of a symmetric buffer.  The bulk variant of the exchange: it takes no SM from the persistent compute kernels
 * that run beside it, where tdc_multicast_copy's CTAs would have to wait for one. */
int tdc_peer_copy(const void* src, void* dst, size_t bytes, tdc_stream_t stream);

/* ---- small dense helpers on the same path ---------------------------------------- */
/* replaces: nn.Linear forward — query_proj / audio_proj / vision_proj as plain callables
 * (cambrian_arch.py:1613,1638,1665).  y[m, n] = x[m, k] . w[n, k]^T + bias.
 * x, w bf16; bias fp32 or NULL; y bf16 (out_dtype TDC_BF16) or fp32 (TDC_F32). */
int tdc_linear(const void* x, const void* w, const float* bias, void* y, int32_t m, int32_t n, int32_t k,
               int32_t out_dtype, int32_t gelu, int32_t cta_group, tdc_stream_t stream);

/* replaces: BertSelfOutput / BertOutput forward — LayerNorm(dense(x) + input_tensor), tdc/Qformer.py:285-289, 371-375 —
 * as ONE kernel: y = LayerNorm(x[m, k] . w[n, k]^T + bias + resid) * gamma + beta, written as fp32 and as bf16.
 * x, w bf16; bias, resid [m, n], gamma, beta fp32; resid may alias y_f32; n <= 768 (a thread-block cluster of
 * ceil(n / 256) CTAs holds one LayerNorm row; row statistics are exchanged through distributed shared memory). */
int tdc_linear_layernorm(const void* x, const void* w, const float* bias, const float* resid, const float* gamma,
                         const float* beta, float eps, float* y_f32, void* y_bf16, int32_t m, int32_t n, int32_t k,
                         tdc_stream_t stream);

/* replaces: mm_projector = Linear -> GELU(erf) -> Linear (cambrian_arch.py:65-69,1149-1150;
 * tdc/multimodal_projector/builder.py:40-47).  x [m, d_in] bf16, w0 [d_mid, d_in], w1 [d_out, d_mid] bf16,
 * biases fp32, mid = caller scratch [m, d_mid] bf16, y [m, d_out] bf16. */
int tdc_gelu_mlp(const void* x, const void* w0, const float* b0, const void* w1, const float* b1, void* mid, void* y,
                 int32_t m, int32_t d_in, int32_t d_mid, int32_t d_out, tdc_stream_t stream);

/* replaces: F.adaptive_avg_pool1d(key_frame.permute(2,0,1), K).permute(1,2,0)
 *           — tdc/cambrian_arch.py:1633-1637.  frames [n, tokens, d] (dtype) -> out [n, K, d] bf16. */
int tdc_avg_pool_tokens(const void* frames, int32_t dtype, int32_t n, int32_t tokens, int32_t d, int32_t num_query,
                        void* out_bf16, tdc_stream_t stream);

/* replaces: the similarity + boundary selection of adapt_segment — tdc/cambrian_arch.py:832-849.
 *   feats [n_frames, dim] (dtype; dim = tokens * channels of the DINO features, flattened)
 *   cos_out [n_frames - 1] fp32: F.cosine_similarity(frame i, frame i+1)
 *   boundaries_out [min(max_segments, n_frames - 1)] int64: sort(argsort(cos)[:max_segments])
 *   workspace: tdc_segment_workspace_bytes(n_frames, dim) bytes */
size_t tdc_segment_workspace_bytes(int32_t n_frames, int64_t dim);
int tdc_segment_boundaries(const void* feats, int32_t dtype, int32_t n_frames, int64_t dim, int32_t max_segments,
                           float* cos_out, int64_t* boundaries_out, void* workspace, size_t workspace_bytes,
                           tdc_stream_t stream);

/* ---- building blocks shared with the SVA connector (tdc/vision_sampler.py) ------------------ */
/* replaces: nn.LayerNorm forward (optionally on x + resid).  x, resid fp32 [rows, width]; resid_period > 0 makes
 * resid a [resid_period, width] table added cyclically (vision_sampler.py:376-386 position embeddings of the KV
 * windows).  Writes y as fp32 and/or bf16 (either may be NULL). */
int tdc_layernorm(const float* x, const float* resid, int32_t resid_period, const float* gamma, const float* beta,
                  float eps, float* y_f32, void* y_bf16, int64_t rows, int32_t width, tdc_stream_t stream);

/* replaces: torch.nn.functional.scaled_dot_product_attention with a boolean key mask and head size 64
 * (vision_sampler.py:272-276) — and BertSelfAttention's softmax(QK^T/8)V (Qformer.py:205-268).
 *   q / k / v / out: bf16; head h of a token at ptr + token_row * ld + h * 64
 *   row r, query i lives at token row  q_base[i < q_seg1 ? 0 : 1] + r * q_seg{1,2} + i (- q_seg1), same for KV
 *   kv_len  [rows] int32 or NULL; kv_mask [rows] uint32 or NULL (bit j = KV token j allowed; <= 32 KV tokens) */
int tdc_attention(const void* q, const void* k, const void* v, void* out, int64_t ldq, int64_t ldk, int64_t ldv,
                  int64_t ldo, int32_t rows, int32_t heads, int32_t q_seg1, int32_t q_seg2, int64_t q_base1,
                  int64_t q_base2, int32_t kv_seg1, int32_t kv_seg2, int64_t kv_base1, int64_t kv_base2,
                  const int32_t* kv_len, const uint32_t* kv_mask, tdc_stream_t stream);

/* replaces: F.interpolate(x.permute(0,2,1).view(bs,-1,s,s).float(), size=(S,S), mode="bilinear", align_corners=False)
 * and the permute back — the query groups whose grid differs from the final token grid, cambrian_arch.py:1107-1131.
 * Token-major in and out: in [bs, side_in^2, d] -> out [bs, side_out^2, d]; fp32 arithmetic. */
int tdc_resize_tokens_bilinear(const void* in, int32_t in_dtype, int32_t bs, int32_t side_in, int32_t side_out,
                               int32_t d, void* out, int32_t out_dtype, tdc_stream_t stream);

/* replaces: the window regrouping of rearrange_vision_tower_features_inference (cambrian_arch.py:624-645) — the
 * tokens under every query of a q x q grid, window-major: in [bs, (q r)^2, d] (dtype) -> out [bs, q, q, r, r, d] bf16. */
int tdc_window_rearrange(const void* in, int32_t in_dtype, int32_t bs, int32_t q, int32_t r, int32_t d, void* out_bf16,
                         tdc_stream_t stream);

/* replaces: queries + (stack(aggregated) * weight_mlp(...).softmax(-1).unsqueeze(-1)).sum(2) of VisionAggregationLayer
 * (tdc/vision_sampler.py:468-474, 505-507): out = base + sum_t softmax(logits[:, :num_parts])[t] * parts[t].
 * base, out fp32 [rows, width]; parts fp32 [num_parts, rows, width]; logits fp32 [rows, ld_logits]. */
int tdc_combine_parts(const float* base, const float* parts, const float* logits, int32_t ld_logits, int32_t num_parts,
                      int64_t rows, int32_t width, float* out, tdc_stream_t stream);

/* out = a + b in fp32 (out_f32 and/or out_bf16 may be NULL) — the outer residual of an SVA layer (:399). */
int tdc_residual_add(const float* a, const float* b, float* out_f32, void* out_bf16, int64_t count,
                     tdc_stream_t stream);

/* dtype conversion used by the host wrapper (fp32 / fp16 -> bf16 and back) */
int tdc_convert(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, int64_t count, tdc_stream_t stream);

/* ---- instrumentation ------------------------------------------------------------- */
/* Per-kernel-class device time, measured with CUDA events on the launching stream when
 * profiling is on.  Classes: see TDC_K_* below. */
enum {
  TDC_K_KV_GEMM = 0,    /* cross-attention K/V projection GEMM (dominant kernel) */
  TDC_K_QUERY_GEMM = 1, /* all query-side GEMMs */
  TDC_K_ATTENTION = 2,  /* self + cross short-query attention */
  TDC_K_ROWOPS = 3,     /* LayerNorm / embeddings / L2-normalise / converts */
  TDC_K_FRONTEND = 4,   /* upstream entry: mm_projector / audio_proj / query_proj GEMMs, gathers, pooling, assembly */
  TDC_K_COUNT = 5
};
int tdc_set_profiling(tdc_handle* h, int32_t enabled);
/* Sums (ms) and launch counts since the last reset; synchronises the recorded events. */
int tdc_get_profile(tdc_handle* h, int32_t kernel_class, double* total_ms, int64_t* launches);
int tdc_reset_profile(tdc_handle* h);
/* Total kernels launched by this handle since creation (for launch accounting). */
int64_t tdc_launch_count(const tdc_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* TDC_B200_H_ */
