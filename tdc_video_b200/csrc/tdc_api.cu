// C ABI of libtdc_b200.so (see include/tdc_b200.h): weight re-packing and the launch
// sequence of the batched Q-Former forward.  No torch types, no hidden allocations in the
// forward calls, no host<->device synchronisation.
//
// Launch sequence for `rows` rows (one row = one dynamic frame = one Q-Former batch element
// of the reference, tdc/cambrian_arch.py:1625-1662), K queries, T text tokens, L KV tokens:
//
//   [convert enc -> bf16]                                                   (only if not bf16)
//   KV   = enc . Wkv^T + b        one GEMM for all cross layers, N = 2*H*n_cross     (Qformer.py:186-187)
//   h    = LN(embeddings)                                                            (Qformer.py:78-108)
//   for each layer:
//     qkv  = h . Wqkv^T + b                                                          (Qformer.py:125-198)
//     ctx  = softmax(q k^T / 8) v        over the row's K+T tokens                   (Qformer.py:205-268)
//     h    = LN(ctx . Wo^T + b + h)                                                  (Qformer.py:285-289)
//     even layers, query tokens only:                                                (Qformer.py:430-447)
//       qc = h . Wq^T + b ; ctx = softmax(qc KV_k^T / 8) KV_v ; h = LN(ctx . Wo^T + b + h)
//     query tokens: h = LN(gelu(h W1^T + b1) W2^T + b2 + h)   (intermediate_query/output_query, :449-454)
//     text  tokens: same with intermediate/output                                    (:455-462)
//   out  = h  |  F.normalize(h[:, :K] . Wvp^T + b)                          (cambrian_arch.py:1664-1667)
#include "tdc_b200.h"
#include "tdc_gemm.cuh"
#include "tdc_kernels.cuh"

#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <string>
#include <vector>
#include <algorithm>

using namespace tdc;

namespace {

thread_local std::string g_create_error;

struct LayerW {
  __nv_bfloat16* w_qkv = nullptr; float* b_qkv = nullptr;
  __nv_bfloat16* w_ao = nullptr; float* b_ao = nullptr; float* ln_a_g = nullptr; float* ln_a_b = nullptr;
  int cross_index = -1;  // >= 0: this layer has cross-attention; slot in the fused K/V weight
  __nv_bfloat16* w_cq = nullptr; float* b_cq = nullptr;
  __nv_bfloat16* w_co = nullptr; float* b_co = nullptr; float* ln_c_g = nullptr; float* ln_c_b = nullptr;
  __nv_bfloat16* w_fq1 = nullptr; float* b_fq1 = nullptr; __nv_bfloat16* w_fq2 = nullptr; float* b_fq2 = nullptr;
  float* ln_fq_g = nullptr; float* ln_fq_b = nullptr;
  __nv_bfloat16* w_ft1 = nullptr; float* b_ft1 = nullptr; __nv_bfloat16* w_ft2 = nullptr; float* b_ft2 = nullptr;
  float* ln_ft_g = nullptr; float* ln_ft_b = nullptr;
};

struct ProfileClass {
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
  size_t used = 0;
};

}  // namespace

struct tdc_handle {
  tdc_config cfg{};
  int n_cross = 0;
  void* arena = nullptr;
  size_t arena_bytes = 0;
  std::vector<LayerW> layers;
  __nv_bfloat16* w_ckv = nullptr; float* b_ckv = nullptr;  // [n_cross*2H, d_enc]
  float* word_emb = nullptr; float* pos_emb = nullptr; float* ln_e_g = nullptr; float* ln_e_b = nullptr;
  __nv_bfloat16* w_vp = nullptr; float* b_vp = nullptr;
  // upstream ("frames") entry: mm_projector, image_newline, query_proj, audio_proj and the folded K/V weights
  __nv_bfloat16* w_p0 = nullptr; float* b_p0 = nullptr;      // mm_projector.0 [d, d_in]
  __nv_bfloat16* w_p2 = nullptr; float* b_p2 = nullptr;      // mm_projector.2 [d, d]
  __nv_bfloat16* w_p2t = nullptr;                            // its transpose (fold scratch)
  float* newline = nullptr;                                  // image_newline [d]
  __nv_bfloat16* w_qp = nullptr; float* b_qp = nullptr;      // query_proj [hidden, d]
  float* query_tokens = nullptr;                             // learned queries [K_max, hidden] (optional)
  int query_tokens_rows = 0;
  __nv_bfloat16* w_kvf = nullptr; float* b_kvf = nullptr;    // Wckv . W2 [kvw, d], Wckv b2 + b_ckv
  float* kv_newline = nullptr;                               // Wckv newline + b_ckv [kvw]
  __nv_bfloat16* w_ap = nullptr; float* b_ap = nullptr;      // audio_proj [d, d_audio]
  __nv_bfloat16* w_apt = nullptr;                            // transpose (fold scratch)
  __nv_bfloat16* w_kva = nullptr; float* b_kva = nullptr;    // Wckv . Wap [kvw, d_audio], Wckv b_ap + b_ckv
  bool loaded = false, have_text_ffn = false, have_embeddings = false, have_vp = false;
  bool have_frontend = false, have_audio = false, have_query_tokens = false;
  std::vector<uint8_t> seen;  // per expected tensor
  std::string last_error;
  bool profiling = false;
  ProfileClass prof[TDC_K_COUNT];
  int64_t launches = 0;
};

namespace {

constexpr size_t kAlign = 256;
constexpr int kMaxLearnedQueries = 256;  // rows reserved for `query_tokens` (context_token_num <= 256)
inline size_t align_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

int fail(tdc_handle* h, int code, const std::string& msg) {
  if (h) h->last_error = msg;
  return code;
}

// ---- weight arena -------------------------------------------------------------------
struct Carver {
  uint8_t* base;
  size_t off = 0;
  template <typename T> T* take(size_t count) {
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += align_up(count * sizeof(T));
    return p;
  }
};

void carve_weights(tdc_handle* h, uint8_t* base, size_t* total) {
  const tdc_config& c = h->cfg;
  const size_t H = c.hidden, I = c.intermediate, E = c.d_enc;
  Carver cv{base};
  h->w_ckv = cv.take<__nv_bfloat16>(static_cast<size_t>(h->n_cross) * 2 * H * E);
  h->b_ckv = cv.take<float>(static_cast<size_t>(h->n_cross) * 2 * H);
  h->ln_e_g = cv.take<float>(H);
  h->ln_e_b = cv.take<float>(H);
  if (c.vocab > 0) {
    h->word_emb = cv.take<float>(static_cast<size_t>(c.vocab) * H);
    h->pos_emb = cv.take<float>(static_cast<size_t>(c.max_pos) * H);
  }
  if (c.d_out > 0) {
    h->w_vp = cv.take<__nv_bfloat16>(static_cast<size_t>(c.d_out) * H);
    h->b_vp = cv.take<float>(c.d_out);
  }
  if (c.d_frame_in > 0) {
    const size_t D = c.d_enc, kvw = static_cast<size_t>(h->n_cross) * 2 * H;
    h->w_p0 = cv.take<__nv_bfloat16>(D * c.d_frame_in); h->b_p0 = cv.take<float>(D);
    h->w_p2 = cv.take<__nv_bfloat16>(D * D); h->b_p2 = cv.take<float>(D);
    h->w_p2t = cv.take<__nv_bfloat16>(D * D);
    h->newline = cv.take<float>(D);
    h->w_qp = cv.take<__nv_bfloat16>(H * D); h->b_qp = cv.take<float>(H);
    h->query_tokens = cv.take<float>(static_cast<size_t>(kMaxLearnedQueries) * H);
    h->w_kvf = cv.take<__nv_bfloat16>(kvw * D); h->b_kvf = cv.take<float>(kvw);
    h->kv_newline = cv.take<float>(kvw);
    if (c.d_audio > 0) {
      h->w_ap = cv.take<__nv_bfloat16>(D * c.d_audio); h->b_ap = cv.take<float>(D);
      h->w_apt = cv.take<__nv_bfloat16>(D * c.d_audio);
      h->w_kva = cv.take<__nv_bfloat16>(kvw * c.d_audio); h->b_kva = cv.take<float>(kvw);
    }
  }
  int cross = 0;
  for (int l = 0; l < c.layers; ++l) {
    LayerW& w = h->layers[l];
    w.w_qkv = cv.take<__nv_bfloat16>(3 * H * H); w.b_qkv = cv.take<float>(3 * H);
    w.w_ao = cv.take<__nv_bfloat16>(H * H); w.b_ao = cv.take<float>(H);
    w.ln_a_g = cv.take<float>(H); w.ln_a_b = cv.take<float>(H);
    if (l % c.cross_freq == 0) {
      w.cross_index = cross++;
      w.w_cq = cv.take<__nv_bfloat16>(H * H); w.b_cq = cv.take<float>(H);
      w.w_co = cv.take<__nv_bfloat16>(H * H); w.b_co = cv.take<float>(H);
      w.ln_c_g = cv.take<float>(H); w.ln_c_b = cv.take<float>(H);
    }
    w.w_fq1 = cv.take<__nv_bfloat16>(I * H); w.b_fq1 = cv.take<float>(I);
    w.w_fq2 = cv.take<__nv_bfloat16>(H * I); w.b_fq2 = cv.take<float>(H);
    w.ln_fq_g = cv.take<float>(H); w.ln_fq_b = cv.take<float>(H);
    if (c.vocab > 0) {
      w.w_ft1 = cv.take<__nv_bfloat16>(I * H); w.b_ft1 = cv.take<float>(I);
      w.w_ft2 = cv.take<__nv_bfloat16>(H * I); w.b_ft2 = cv.take<float>(H);
      w.ln_ft_g = cv.take<float>(H); w.ln_ft_b = cv.take<float>(H);
    }
  }
  *total = cv.off;
}

// A destination inside the arena for one named reference tensor.
struct Slot {
  void* dst = nullptr;
  bool as_bf16 = false;
  int64_t rows = 0, cols = 0;  // expected shape (cols == 0: 1-D of `rows`)
  int group = 0;               // 0 core, 1 text FFN, 2 embeddings, 3 vision_proj
};

bool resolve_slot(tdc_handle* h, const std::string& name, Slot* s) {
  const tdc_config& c = h->cfg;
  const int64_t H = c.hidden, I = c.intermediate, E = c.d_enc;
  auto W = [&](__nv_bfloat16* p, int64_t r, int64_t cc, int grp = 0) { *s = Slot{p, true, r, cc, grp}; return p != nullptr; };
  auto V = [&](float* p, int64_t r, int grp = 0) { *s = Slot{p, false, r, 0, grp}; return p != nullptr; };
  if (name == "vision_proj.weight") return W(h->w_vp, c.d_out, H, 3);
  if (name == "vision_proj.bias") return V(h->b_vp, c.d_out, 3);
  if (c.d_frame_in > 0) {
    const int64_t D = c.d_enc;
    if (name == "mm_projector.0.weight") return W(h->w_p0, D, c.d_frame_in, 4);
    if (name == "mm_projector.0.bias") return V(h->b_p0, D, 4);
    if (name == "mm_projector.2.weight") return W(h->w_p2, D, D, 4);
    if (name == "mm_projector.2.bias") return V(h->b_p2, D, 4);
    if (name == "image_newline") return V(h->newline, D, 4);
    if (name == "query_proj.weight") return W(h->w_qp, H, D, 4);
    if (name == "query_proj.bias") return V(h->b_qp, H, 4);
    if (name == "query_tokens") { *s = Slot{h->query_tokens, false, -1, H, 6}; return true; }  // [K, hidden], K free
    if (c.d_audio > 0) {
      if (name == "audio_proj.weight") return W(h->w_ap, D, c.d_audio, 5);
      if (name == "audio_proj.bias") return V(h->b_ap, D, 5);
    }
  }
  if (name == "embeddings.LayerNorm.weight") return V(h->ln_e_g, H);
  if (name == "embeddings.LayerNorm.bias") return V(h->ln_e_b, H);
  if (name == "embeddings.word_embeddings.weight") { *s = Slot{h->word_emb, false, c.vocab, H, 2}; return h->word_emb != nullptr; }
  if (name == "embeddings.position_embeddings.weight") { *s = Slot{h->pos_emb, false, c.max_pos, H, 2}; return h->pos_emb != nullptr; }
  const std::string pre = "encoder.layer.";
  if (name.compare(0, pre.size(), pre) != 0) return false;
  size_t dot = name.find('.', pre.size());
  if (dot == std::string::npos) return false;
  int l = 0;
  for (size_t i = pre.size(); i < dot; ++i) {
    if (name[i] < '0' || name[i] > '9') return false;
    l = l * 10 + (name[i] - '0');
  }
  if (l < 0 || l >= c.layers) return false;
  LayerW& w = h->layers[l];
  const std::string rest = name.substr(dot + 1);
  if (rest == "attention.self.query.weight") return W(w.w_qkv, H, H);
  if (rest == "attention.self.key.weight") return W(w.w_qkv + H * H, H, H);
  if (rest == "attention.self.value.weight") return W(w.w_qkv + 2 * H * H, H, H);
  if (rest == "attention.self.query.bias") return V(w.b_qkv, H);
  if (rest == "attention.self.key.bias") return V(w.b_qkv + H, H);
  if (rest == "attention.self.value.bias") return V(w.b_qkv + 2 * H, H);
  if (rest == "attention.output.dense.weight") return W(w.w_ao, H, H);
  if (rest == "attention.output.dense.bias") return V(w.b_ao, H);
  if (rest == "attention.output.LayerNorm.weight") return V(w.ln_a_g, H);
  if (rest == "attention.output.LayerNorm.bias") return V(w.ln_a_b, H);
  if (w.cross_index >= 0) {
    const int64_t j = w.cross_index;
    if (rest == "crossattention.self.query.weight") return W(w.w_cq, H, H);
    if (rest == "crossattention.self.query.bias") return V(w.b_cq, H);
    if (rest == "crossattention.self.key.weight") return W(h->w_ckv + (j * 2) * H * E, H, E);
    if (rest == "crossattention.self.value.weight") return W(h->w_ckv + (j * 2 + 1) * H * E, H, E);
    if (rest == "crossattention.self.key.bias") return V(h->b_ckv + (j * 2) * H, H);
    if (rest == "crossattention.self.value.bias") return V(h->b_ckv + (j * 2 + 1) * H, H);
    if (rest == "crossattention.output.dense.weight") return W(w.w_co, H, H);
    if (rest == "crossattention.output.dense.bias") return V(w.b_co, H);
    if (rest == "crossattention.output.LayerNorm.weight") return V(w.ln_c_g, H);
    if (rest == "crossattention.output.LayerNorm.bias") return V(w.ln_c_b, H);
  }
  if (rest == "intermediate_query.dense.weight") return W(w.w_fq1, I, H);
  if (rest == "intermediate_query.dense.bias") return V(w.b_fq1, I);
  if (rest == "output_query.dense.weight") return W(w.w_fq2, H, I);
  if (rest == "output_query.dense.bias") return V(w.b_fq2, H);
  if (rest == "output_query.LayerNorm.weight") return V(w.ln_fq_g, H);
  if (rest == "output_query.LayerNorm.bias") return V(w.ln_fq_b, H);
  if (rest == "intermediate.dense.weight") return W(w.w_ft1, I, H, 1);
  if (rest == "intermediate.dense.bias") return V(w.b_ft1, I, 1);
  if (rest == "output.dense.weight") return W(w.w_ft2, H, I, 1);
  if (rest == "output.dense.bias") return V(w.b_ft2, H, 1);
  if (rest == "output.LayerNorm.weight") return V(w.ln_ft_g, H, 1);
  if (rest == "output.LayerNorm.bias") return V(w.ln_ft_b, H, 1);
  return false;
}

// Names every handle must receive (core) — used to detect an incomplete load.
std::vector<std::string> expected_names(const tdc_handle* h, int group) {
  std::vector<std::string> out;
  const tdc_config& c = h->cfg;
  auto wb = [&](const std::string& p) { out.push_back(p + ".weight"); out.push_back(p + ".bias"); };
  if (group == 0) wb("embeddings.LayerNorm");
  if (group == 2) { out.push_back("embeddings.word_embeddings.weight"); out.push_back("embeddings.position_embeddings.weight"); }
  if (group == 3) wb("vision_proj");
  if (group == 4) { wb("mm_projector.0"); wb("mm_projector.2"); wb("query_proj"); out.push_back("image_newline"); }
  if (group == 5) wb("audio_proj");
  if (group >= 4) return out;
  for (int l = 0; l < c.layers; ++l) {
    const std::string p = "encoder.layer." + std::to_string(l) + ".";
    if (group == 0) {
      wb(p + "attention.self.query"); wb(p + "attention.self.key"); wb(p + "attention.self.value");
      wb(p + "attention.output.dense"); wb(p + "attention.output.LayerNorm");
      if (l % c.cross_freq == 0) {
        wb(p + "crossattention.self.query"); wb(p + "crossattention.self.key"); wb(p + "crossattention.self.value");
        wb(p + "crossattention.output.dense"); wb(p + "crossattention.output.LayerNorm");
      }
      wb(p + "intermediate_query.dense"); wb(p + "output_query.dense"); wb(p + "output_query.LayerNorm");
    }
    if (group == 1) { wb(p + "intermediate.dense"); wb(p + "output.dense"); wb(p + "output.LayerNorm"); }
  }
  return out;
}

// ---- profiling ----------------------------------------------------------------------
struct KernelScope {
  tdc_handle* h;
  cudaStream_t s;
  cudaEvent_t stop = nullptr;
  KernelScope(tdc_handle* h_, int cls, cudaStream_t s_, int n_launches = 1) : h(h_), s(s_) {
    h->launches += n_launches;
    if (!h->profiling) return;
    ProfileClass& pc = h->prof[cls];
    if (pc.used == pc.events.size()) {
      cudaEvent_t a, b;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      pc.events.emplace_back(a, b);
    }
    cudaEventRecord(pc.events[pc.used].first, s);
    stop = pc.events[pc.used].second;
    ++pc.used;
  }
  ~KernelScope() {
    if (stop) cudaEventRecord(stop, s);
  }
};

// ---- forward ------------------------------------------------------------------------
struct Workspace {
  __nv_bfloat16* enc_bf16;
  __nv_bfloat16* kv;
  float* h_f32;
  __nv_bfloat16* h_bf16;
  __nv_bfloat16* qkv;
  __nv_bfloat16* ctx;
  float* pre;
  __nv_bfloat16* qc;
  __nv_bfloat16* mid;
  float* proj;
  size_t bytes;
};

Workspace carve_workspace(const tdc_handle* h, uint8_t* base, long long rows, int L, int K, int T, bool enc_convert,
                          bool with_proj) {
  const tdc_config& c = h->cfg;
  const size_t H = c.hidden, I = c.intermediate, n = static_cast<size_t>(K) + T, R = static_cast<size_t>(rows);
  Carver cv{base};
  Workspace w{};
  w.enc_bf16 = enc_convert ? cv.take<__nv_bfloat16>(R * L * c.d_enc) : nullptr;
  w.kv = cv.take<__nv_bfloat16>(R * L * 2 * H * h->n_cross);
  w.h_f32 = cv.take<float>(R * n * H);
  w.h_bf16 = cv.take<__nv_bfloat16>(R * n * H);
  w.qkv = cv.take<__nv_bfloat16>(R * n * 3 * H);
  w.ctx = cv.take<__nv_bfloat16>(R * n * H);
  w.pre = cv.take<float>(R * n * H);
  w.qc = cv.take<__nv_bfloat16>(R * K * H);
  w.mid = cv.take<__nv_bfloat16>(R * std::max<size_t>(K, T) * I);
  w.proj = with_proj ? cv.take<float>(R * K * c.d_out) : nullptr;
  w.bytes = cv.off;
  return w;
}

struct ForwardCall {
  const void* query_embeds; int query_dtype; const int32_t* query_set;
  const int64_t* input_ids; const int32_t* text_set;
  const void* enc; int enc_dtype; const int32_t* kv_len;
  long long rows; int L, K, T;
  void* out; int out_dtype;   // hidden [rows, K+T, H] or compressed [rows, K, d_out]
  bool compress;
  bool multicast = false;     // compress only: `out` is an NVSwitch multicast address
  // layer-0 de-duplication (rows that share queries and prompt, cambrian_arch.py:1629-1646):
  float* l0_out = nullptr;            // pre-pass: stop after the self-attention block of layer 0 and write the
                                      // state of every "row" (= one per set) set-major [rows, K + T, hidden] fp32
  const float* l0_sets = nullptr;     // main pass: start from these states instead of recomputing them per row
  const int32_t* l0_map = nullptr;    // [rows] row -> set (NULL: set 0)
  int l0_num_sets = 1;
  int num_query_sets = 0, num_text_sets = 0;   // > 0: device-side set indices are clamped into range
};

#define TDC_TRY(expr)                                         \
  do {                                                        \
    const int rc_ = (expr);                                   \
    if (rc_ != TDC_OK) return fail(h, rc_, err ? err : "?");  \
  } while (0)

int gemm(tdc_handle* h, int cls, cudaStream_t s, const void* a, long long lda, const void* w, long long ldw,
         const float* bias, void* out, long long ldo, long long m, int n, int k, int mode, const char** err,
         int slab_cols = 0, long long slab_stride = 0) {
  GemmProblem p;
  p.a = a; p.lda = lda; p.w = w; p.ldw = ldw; p.bias = bias; p.out = out; p.ldo = ldo;
  p.m = static_cast<int>(m); p.n = n; p.k = k; p.mode = mode; p.slab_cols = slab_cols; p.slab_stride = slab_stride;
  p.cta_group = h->cfg.gemm_cta_group == 0 ? 2 : h->cfg.gemm_cta_group;
  KernelScope ks(h, cls, s);
  return gemm_launch(p, s, err);
}

// dense + residual + LayerNorm of a post-LN sub-layer (Qformer.py:285-289, 371-375) on slab rows [first, first + count):
// ONE kernel (gemm_ln_sm100.cu) when the width fits a cluster (N <= 768), otherwise GEMM -> fp32 `pre` -> LayerNorm.
// TDC_NO_FUSED_LN=1 (dev knob, read once) forces the two-kernel form for A/B measurements.
bool fused_ln_enabled(int n) {
  static const bool off = [] { const char* e = getenv("TDC_NO_FUSED_LN"); return e != nullptr && atoi(e) == 1; }();
  return !off && gemm_ln_supported(n);
}

// Where the cross-attention K/V of a row batch live: one dense [slab_rows, 64] matrix per (cross layer, K | V, head),
// `head_stride` elements apart in the column order of the fused K/V weight (layer, K | V, head) — a (row, head) pair
// of an attention launch streams one contiguous run of 128-byte lines, and every TMA store box of the K/V GEMM's
// epilogue (32 tokens x 64 columns) is one contiguous 4 KB write.  Row r's KV tokens: seg1 tokens at slab row r*seg1 + t, then
// seg2 tokens at base2 + r*seg2 + t, then seg3 tokens at base3 + t shared by ALL rows (the image_newline tokens).
struct KvView {
  const __nv_bfloat16* base = nullptr;
  long long head_stride = 0;   // = slab rows * 64
  int seg1 = 0, seg2 = 0, seg3 = 0;
  long long base2 = 0, base3 = 0;
};

// Embeddings + the 12-layer stack + output stage for rows [row0, row0 + rows) whose cross-attention K/V are in `kv`.
int qformer_layers(tdc_handle* h, const ForwardCall& f, long long row0, long long rows, const Workspace& w,
                   const KvView& kv, const int* kv_len, cudaStream_t s) {
  const tdc_config& c = h->cfg;
  const int H = c.hidden, I = c.intermediate, K = f.K, T = f.T, n = K + T;
  const char* err = nullptr;
  const float scale_log2 = 1.4426950408889634f / 8.0f;  // log2(e) / sqrt(64)

  // embeddings + LayerNorm into the [query slab | text slab] layout
  if (f.l0_sets != nullptr) {
    KernelScope ks(h, TDC_K_ROWOPS, s);
    TDC_TRY(broadcast_sets_launch(f.l0_sets, f.l0_map ? f.l0_map + row0 : nullptr, f.l0_num_sets,
                                  static_cast<int>(rows), K, T, H,
                                  w.h_f32, w.h_bf16, s, &err));
  } else {
    EmbedArgs e;
    const size_t qsz = f.query_dtype == TDC_F32 ? 4 : 2;
    e.query_embeds = f.query_set ? f.query_embeds
                                 : static_cast<const uint8_t*>(f.query_embeds) + static_cast<size_t>(row0) * K * H * qsz;
    e.query_dtype = f.query_dtype;
    e.query_set = f.query_set ? f.query_set + row0 : nullptr;
    e.input_ids = (T > 0) ? (f.text_set ? f.input_ids : f.input_ids + row0 * T) : nullptr;
    e.text_set = f.text_set ? f.text_set + row0 : nullptr;
    e.num_query_sets = f.query_set ? f.num_query_sets : 0;
    e.num_text_sets = f.text_set ? f.num_text_sets : 0;
    e.word_emb = h->word_emb; e.pos_emb = h->pos_emb; e.vocab = c.vocab;
    e.gamma = h->ln_e_g; e.beta = h->ln_e_b; e.eps = c.ln_eps;
    e.h_f32 = w.h_f32; e.h_bf16 = w.h_bf16;
    e.rows = static_cast<int>(rows); e.num_query = K; e.num_text = T; e.hidden = H;
    KernelScope ks(h, TDC_K_ROWOPS, s);
    TDC_TRY(embed_layernorm_launch(e, s, &err));
  }

  const long long MQ = rows * K, MT = rows * T, MA = rows * n;
  const bool fused = fused_ln_enabled(H);
  // h[first .. first + count) = LN(a . wt^T + bias + h) * g + b   (a: bf16 [count, kdim])
  auto dense_ln = [&](const __nv_bfloat16* a, int kdim, const __nv_bfloat16* wt, const float* bias, const float* g,
                      const float* b, long long first, long long count) -> int {
    if (fused) {
      GemmLnProblem p;
      p.a = a; p.lda = kdim; p.w = wt; p.ldw = kdim; p.m = static_cast<int>(count); p.n = H; p.k = kdim;
      p.bias = bias; p.resid = w.h_f32 + first * H; p.ldr = H; p.gamma = g; p.beta = b; p.eps = c.ln_eps;
      p.out_f32 = w.h_f32 + first * H; p.out_bf16 = w.h_bf16 + first * H; p.ldo = H;
      KernelScope ks(h, TDC_K_QUERY_GEMM, s);
      return gemm_ln_launch(p, s, &err);
    }
    const int rc = gemm(h, TDC_K_QUERY_GEMM, s, a, kdim, wt, kdim, bias, w.pre + first * H, H, count, H, kdim,
                        EPI_BIAS_F32, &err);
    if (rc != TDC_OK) return rc;
    KernelScope ks(h, TDC_K_ROWOPS, s);
    // h = LN(pre + h): the residual add of the post-LN block lives here, not in the GEMM epilogue
    return layernorm_launch(w.pre + first * H, H, w.h_f32 + first * H, H, g, b, c.ln_eps, w.h_f32 + first * H,
                            w.h_bf16 + first * H, H, count, H, s, &err);
  };

  for (int l = 0; l < c.layers; ++l) {
    const LayerW& lw = h->layers[l];
    // tdc_compress reads only the query tokens of the last layer (cambrian_arch.py:1665 `[:, :K]`): there the
    // text tokens still feed the queries' self-attention as keys / values, but their own attention output,
    // out-projection, LayerNorms and feed-forward are dead and are skipped (the query tokens' bits do not change).
    const bool text_live = T > 0 && !(f.compress && l == c.layers - 1 && f.l0_out == nullptr);
    const int nq_self = text_live ? n : K;
    const long long M_self = text_live ? MA : MQ;
    // ---- self-attention over all K+T tokens of the row
    const bool self_done = (l == 0 && f.l0_sets != nullptr);   // layer 0's self block came from the per-set pre-pass
    if (!self_done) {
    TDC_TRY(gemm(h, TDC_K_QUERY_GEMM, s, w.h_bf16, H, lw.w_qkv, H, lw.b_qkv, w.qkv, 3 * H, MA, 3 * H, H,
                 EPI_BIAS_BF16, &err));
    {
      AttentionArgs a;
      a.q = w.qkv; a.k = w.qkv + H; a.v = w.qkv + 2 * H; a.out = w.ctx;
      a.ldq = a.ldk = a.ldv = 3 * H; a.ldo = H;
      a.rows = static_cast<int>(rows); a.heads = c.heads; a.nq = nq_self;
      a.q_seg1 = K; a.q_seg2 = text_live ? T : 0; a.q_base1 = 0; a.q_base2 = MQ;
      a.kv_seg1 = K; a.kv_seg2 = T; a.kv_base1 = 0; a.kv_base2 = MQ;
      a.scale_log2 = scale_log2;
      KernelScope ks(h, TDC_K_ATTENTION, s);
      TDC_TRY(attention_launch(a, s, &err));
    }
    TDC_TRY(dense_ln(w.ctx, H, lw.w_ao, lw.b_ao, lw.ln_a_g, lw.ln_a_b, 0, M_self));
    }
    if (f.l0_out != nullptr) {   // pre-pass: the state after layer 0's self-attention block, one "row" per set
      KernelScope ks(h, TDC_K_ROWOPS, s);
      TDC_TRY(gather_rows_launch(w.h_f32, H, static_cast<int>(rows), K, T, n,
                                 f.l0_out + static_cast<size_t>(row0) * n * H, TDC_F32, s, &err));
      return TDC_OK;
    }

    // ---- cross-attention: query tokens only
    if (lw.cross_index >= 0) {
      TDC_TRY(gemm(h, TDC_K_QUERY_GEMM, s, w.h_bf16, H, lw.w_cq, H, lw.b_cq, w.qc, H, MQ, H, H, EPI_BIAS_BF16, &err));
      AttentionArgs a;
      a.q = w.qc; a.out = w.ctx; a.ldq = H; a.ldo = H;
      a.k = kv.base + static_cast<size_t>(lw.cross_index) * 2 * c.heads * kv.head_stride;
      a.v = a.k + static_cast<size_t>(c.heads) * kv.head_stride;
      a.ldk = a.ldv = 64;
      a.k_head_stride = a.v_head_stride = kv.head_stride;
      a.rows = static_cast<int>(rows); a.heads = c.heads; a.nq = K;
      a.q_seg1 = K; a.q_seg2 = 0;
      a.kv_seg1 = kv.seg1; a.kv_seg2 = kv.seg2; a.kv_seg3 = kv.seg3;
      a.kv_base2 = kv.base2; a.kv_base3 = kv.base3;
      a.kv_len = kv_len;
      a.scale_log2 = scale_log2;
      {
        KernelScope ks(h, TDC_K_ATTENTION, s);
        TDC_TRY(attention_launch(a, s, &err));
      }
      TDC_TRY(dense_ln(w.ctx, H, lw.w_co, lw.b_co, lw.ln_c_g, lw.ln_c_b, 0, MQ));
    }

    // ---- feed-forward: query tokens and text tokens use different weights
    TDC_TRY(gemm(h, TDC_K_QUERY_GEMM, s, w.h_bf16, H, lw.w_fq1, H, lw.b_fq1, w.mid, I, MQ, I, H, EPI_BIAS_GELU_BF16,
                 &err));
    TDC_TRY(dense_ln(w.mid, I, lw.w_fq2, lw.b_fq2, lw.ln_fq_g, lw.ln_fq_b, 0, MQ));
    if (text_live) {
      TDC_TRY(gemm(h, TDC_K_QUERY_GEMM, s, w.h_bf16 + MQ * H, H, lw.w_ft1, H, lw.b_ft1, w.mid, I, MT, I, H,
                   EPI_BIAS_GELU_BF16, &err));
      TDC_TRY(dense_ln(w.mid, I, lw.w_ft2, lw.b_ft2, lw.ln_ft_g, lw.ln_ft_b, MQ, MT));
    }
  }

  if (!f.compress) {
    const size_t osz = f.out_dtype == TDC_F32 ? 4 : 2;
    void* out = static_cast<uint8_t*>(f.out) + static_cast<size_t>(row0) * n * H * osz;
    KernelScope ks(h, TDC_K_ROWOPS, s);
    TDC_TRY(gather_rows_launch(w.h_f32, H, static_cast<int>(rows), K, T, n, out, f.out_dtype, s, &err));
  } else {
    // vision_proj on the (contiguous) query slab, then unit-normalise every token
    TDC_TRY(gemm(h, TDC_K_QUERY_GEMM, s, w.h_bf16, H, h->w_vp, H, h->b_vp, w.proj, c.d_out, MQ, c.d_out, H,
                 EPI_BIAS_F32, &err));
    const size_t osz = f.out_dtype == TDC_F32 ? 4 : 2;
    void* out = static_cast<uint8_t*>(f.out) + static_cast<size_t>(row0) * K * c.d_out * osz;
    KernelScope ks(h, TDC_K_ROWOPS, s);
    TDC_TRY(l2_normalize_launch(w.proj, c.d_out, out, f.out_dtype, MQ, c.d_out, f.multicast, s, &err));
  }
  return TDC_OK;
}

// One batch of rows [row0, row0 + rows) of the call; all pointers in `f` are for the whole call.
int forward_batch(tdc_handle* h, const ForwardCall& f, long long row0, long long rows, uint8_t* ws_base,
                  cudaStream_t s) {
  const tdc_config& c = h->cfg;
  const int H = c.hidden, K = f.K, T = f.T, L = f.L;
  const char* err = nullptr;
  const bool enc_convert = f.enc_dtype != TDC_BF16;
  Workspace w = carve_workspace(h, ws_base, rows, L, K, T, enc_convert, f.compress);
  const size_t esz = f.enc_dtype == TDC_F32 ? 4 : 2;
  const uint8_t* enc_in = static_cast<const uint8_t*>(f.enc) + static_cast<size_t>(row0) * L * c.d_enc * esz;
  const __nv_bfloat16* enc = reinterpret_cast<const __nv_bfloat16*>(enc_in);
  if (enc_convert) {
    KernelScope ks(h, TDC_K_ROWOPS, s);
    TDC_TRY(convert_launch(enc_in, f.enc_dtype, w.enc_bf16, TDC_BF16, rows * L * c.d_enc, s, &err));
    enc = w.enc_bf16;
  }
  const int kvw = 2 * H * h->n_cross;  // K/V columns per KV token over all cross layers

  // every cross layer's K and V for every KV token of every row: the dominant GEMM.  Output layout: one dense
  // [rows*L, 64] matrix per (layer, K | V, head) — see KvView.
  KvView kv;
  kv.base = w.kv;
  kv.head_stride = rows * L * 64ll;
  kv.seg1 = L;
  if (h->n_cross > 0)
    TDC_TRY(gemm(h, TDC_K_KV_GEMM, s, enc, c.d_enc, h->w_ckv, c.d_enc, h->b_ckv, w.kv, 64, rows * L, kvw,
                 c.d_enc, EPI_BIAS_BF16, &err, 64, kv.head_stride));
  return qformer_layers(h, f, row0, rows, w, kv, f.kv_len ? f.kv_len + row0 : nullptr, s);
}

// ---- upstream ("frames") entry -----------------------------------------------------------------------------
// Weight folding at load time (exact algebra, one extra bf16 rounding of the folded matrices):
//   K/V of a visual token  = (gelu(x W0^T + b0) W2^T + b2) Wckv^T + b_ckv = gelu(..) (Wckv W2)^T + (Wckv b2 + b_ckv)
//   K/V of an audio token  = (a Wap^T + b_ap) Wckv^T + b_ckv            = a (Wckv Wap)^T + (Wckv b_ap + b_ckv)
//   K/V of a newline token = Wckv newline + b_ckv                         (the same for every frame)
// so that a dynamic frame never materialises its d_llm-wide tokens: mm_projector.2 and audio_proj only run on the
// key frames (which pass through to the LLM, cambrian_arch.py:1617-1623).
int fold_frontend_weights(tdc_handle* h, cudaStream_t s) {
  const tdc_config& c = h->cfg;
  const int D = c.d_enc, kvw = 2 * c.hidden * h->n_cross;
  const char* err = nullptr;
  auto fold = [&](const __nv_bfloat16* w_t, int n, __nv_bfloat16* out) -> int {  // out[kvw, n] = Wckv . w_t[n, D]^T
    GemmProblem p;
    p.a = h->w_ckv; p.lda = D; p.w = w_t; p.ldw = D; p.bias = nullptr; p.out = out; p.ldo = n;
    p.m = kvw; p.n = n; p.k = D; p.mode = EPI_BIAS_BF16; p.cta_group = 2;
    ++h->launches;
    return gemm_launch(p, s, &err);
  };
  if (h->n_cross > 0) {
    TDC_TRY(transpose_bf16_launch(h->w_p2, D, D, h->w_p2t, s, &err));
    TDC_TRY(fold(h->w_p2t, D, h->w_kvf));
    TDC_TRY(matvec_bias_launch(h->w_ckv, h->b_p2, h->b_ckv, h->b_kvf, kvw, D, s, &err));
    TDC_TRY(matvec_bias_launch(h->w_ckv, h->newline, h->b_ckv, h->kv_newline, kvw, D, s, &err));
    h->launches += 3;
    if (h->have_audio) {
      TDC_TRY(transpose_bf16_launch(h->w_ap, D, c.d_audio, h->w_apt, s, &err));
      TDC_TRY(fold(h->w_apt, c.d_audio, h->w_kva));
      TDC_TRY(matvec_bias_launch(h->w_ckv, h->b_ap, h->b_ckv, h->b_kva, kvw, D, s, &err));
      h->launches += 2;
    }
  }
  return TDC_OK;
}

struct FramesCall {
  const tdc_frames_args* a;
  int side;        // token grid side: Tv = side * side
  int Tv, Ta, K, T;
  bool learned;    // queries = the loaded query_tokens instead of avg-pool(key frame) -> query_proj
};

struct FramesWorkspace {
  // whole call
  float* q_sets;       // [n_chunks, K, H] fp32 (Avg_pool queries, one set per chunk)
  int32_t* zero_map;   // [max(rows, n_chunks)] zeros: "every row uses set 0" (shared prompt / learned queries)
  int32_t* row_prompt; // [rows] prompt of every row when several videos share the call
  float* l0_sets;      // [n_chunks, K + T, H] fp32: state after layer 0's self-attention block, one per chunk
  // per batch of nb items (nb key frames in the static pass, nb rows in the dynamic pass)
  __nv_bfloat16* fin;     // [nb*Tv, d_in]   gathered tower features
  __nv_bfloat16* pmid;    // [nb*Tv, d]      gelu(mm_projector.0)
  __nv_bfloat16* xv;      // [nb*Tv, d]      projected visual tokens (key frames; dynamic frames only if !fold)
  __nv_bfloat16* ain;     // [nb*Ta, d_audio]
  __nv_bfloat16* xa;      // [nb*Ta, d]
  __nv_bfloat16* pooled;  // [nb*K, d]
  __nv_bfloat16* kv;      // n_cross * 2 * heads slabs of [nb*(Tv+Ta) + side, 64]
  uint8_t* qws;           // Q-Former workspace of nb rows (carve_workspace without the kv / enc parts)
  size_t bytes;
};

FramesWorkspace carve_frames(const tdc_handle* h, uint8_t* base, long long n_chunks, long long rows, long long nb,
                             int Tv, int Ta, int side, int K, int T, bool fold = false) {
  const tdc_config& c = h->cfg;
  const size_t D = c.d_enc, H = c.hidden, N = static_cast<size_t>(nb);
  Carver cv{base};
  FramesWorkspace w{};
  w.q_sets = cv.take<float>(static_cast<size_t>(n_chunks) * K * H);
  w.zero_map = cv.take<int32_t>(static_cast<size_t>(std::max(rows, n_chunks)));
  w.row_prompt = cv.take<int32_t>(static_cast<size_t>(rows));
  w.l0_sets = cv.take<float>(static_cast<size_t>(n_chunks) * (K + T) * H);
  w.fin = cv.take<__nv_bfloat16>(N * Tv * c.d_frame_in);
  w.pmid = cv.take<__nv_bfloat16>(N * Tv * D);
  w.ain = cv.take<__nv_bfloat16>(N * Ta * std::max(c.d_audio, 8));
  // K/V slabs of the dynamic pass.  With folded weights the projected visual / audio tokens and the pooled queries
  // are needed by the key-frame pass only, which is over before the first K/V GEMM runs: they alias the slabs
  // (1.5 MB per item saved -> larger row batches in the same workspace).  With fold = 0 the dynamic pass reads xv / xa
  // as the INPUT of the K/V GEMMs, so they get their own space behind the slabs.
  const size_t kv_elems = (N * (Tv + Ta) + side) * 2 * H * h->n_cross;
  const size_t alias_elems = N * Tv * D + N * Ta * D + N * K * D;
  w.kv = cv.take<__nv_bfloat16>(std::max(kv_elems, fold ? alias_elems : size_t{0}));
  __nv_bfloat16* alias_base = fold ? w.kv : cv.take<__nv_bfloat16>(alias_elems);
  w.xv = alias_base;
  w.xa = alias_base ? alias_base + N * Tv * D : nullptr;
  w.pooled = alias_base ? alias_base + N * Tv * D + N * Ta * D : nullptr;
  w.qws = base ? base + cv.off : nullptr;
  // the Q-Former's own buffers: carve_workspace with L = 0 drops its enc / kv parts
  cv.off += carve_workspace(h, nullptr, nb, 0, K, T, false, true).bytes;
  w.bytes = cv.off;
  return w;
}

int frames_gemm(tdc_handle* h, int cls, cudaStream_t s, const void* a, long long k, const void* w, const float* bias,
                void* out, long long m, int n, int mode, const char** err, int slab_cols = 0, long long slab_stride = 0,
                long long ldo = 0) {
  return gemm(h, cls, s, a, k, w, k, bias, out, ldo > 0 ? ldo : n, m, n, static_cast<int>(k), mode, err, slab_cols,
              slab_stride);
}

// Key frames [c0, c0 + cb): mm_projector -> (pool -> query_proj -> q_sets) and (assemble static_out).
int frames_static_batch(tdc_handle* h, const FramesCall& fc, const FramesWorkspace& w, long long c0, long long cb,
                        cudaStream_t s) {
  const tdc_config& c = h->cfg;
  const tdc_frames_args& a = *fc.a;
  const int D = c.d_enc, H = c.hidden, Tv = fc.Tv, Ta = fc.Ta, K = fc.K;
  const char* err = nullptr;
  {
    KernelScope ks(h, TDC_K_FRONTEND, s);
    TDC_TRY(gather_blocks_launch(a.frames, a.static_frames + c0, w.fin, cb, static_cast<long long>(Tv) * c.d_frame_in * 2,
                                 a.n_frames, s, &err));
  }
  TDC_TRY(frames_gemm(h, TDC_K_FRONTEND, s, w.fin, c.d_frame_in, h->w_p0, h->b_p0, w.pmid, cb * Tv, D,
                      EPI_BIAS_GELU_BF16, &err));
  TDC_TRY(frames_gemm(h, TDC_K_FRONTEND, s, w.pmid, D, h->w_p2, h->b_p2, w.xv, cb * Tv, D, EPI_BIAS_BF16, &err));
  if (!fc.learned) {
    {
      KernelScope ks(h, TDC_K_FRONTEND, s);
      TDC_TRY(pool_static_queries_launch(w.xv, h->newline, static_cast<int>(cb), fc.side, D, K, w.pooled, s, &err));
    }
    TDC_TRY(frames_gemm(h, TDC_K_FRONTEND, s, w.pooled, D, h->w_qp, h->b_qp, w.q_sets + c0 * K * H, cb * K, H,
                        EPI_BIAS_F32, &err));
  }
  if (a.static_out != nullptr) {
    if (Ta > 0) {
      {
        KernelScope ks(h, TDC_K_FRONTEND, s);
        TDC_TRY(gather_blocks_launch(a.audio, a.static_frames + c0, w.ain, cb, static_cast<long long>(Ta) * c.d_audio * 2,
                                     a.n_frames, s, &err));
      }
      TDC_TRY(frames_gemm(h, TDC_K_FRONTEND, s, w.ain, c.d_audio, h->w_ap, h->b_ap, w.xa, cb * Ta, D, EPI_BIAS_BF16,
                          &err));
    }
    const size_t osz = a.out_dtype == TDC_F32 ? 4 : 2;
    const long long ls = static_cast<long long>(fc.side) * (fc.side + 1) + Ta;
    KernelScope ks(h, TDC_K_FRONTEND, s);
    TDC_TRY(assemble_static_launch(w.xv, w.xa, h->newline, static_cast<int>(cb), fc.side, Ta, D,
                                   static_cast<uint8_t*>(a.static_out) + static_cast<size_t>(c0) * ls * D * osz,
                                   a.out_dtype, a.static_multicast != 0, s, &err));
  }
  return TDC_OK;
}

// Dynamic frames (rows) [r0, r0 + rb): mm_projector.0 -> K/V of the frame's visual, audio and newline tokens ->
// Q-Former + vision_proj + L2-normalise.
int frames_dynamic_batch(tdc_handle* h, const FramesCall& fc, const ForwardCall& f, const FramesWorkspace& w,
                         long long r0, long long rb, cudaStream_t s) {
  const tdc_config& c = h->cfg;
  const tdc_frames_args& a = *fc.a;
  const int D = c.d_enc, H = c.hidden, Tv = fc.Tv, Ta = fc.Ta;
  const int kvw = 2 * H * h->n_cross;
  const char* err = nullptr;
  KvView kv;
  kv.base = w.kv;
  kv.seg1 = Tv; kv.seg2 = Ta; kv.seg3 = fc.side;
  kv.base2 = rb * Tv;
  kv.base3 = rb * (Tv + Ta);
  kv.head_stride = (rb * (Tv + Ta) + fc.side) * 64ll;
  {
    KernelScope ks(h, TDC_K_FRONTEND, s);
    TDC_TRY(gather_blocks_launch(a.frames, a.row_frames + r0, w.fin, rb, static_cast<long long>(Tv) * c.d_frame_in * 2,
                                 a.n_frames, s, &err));
  }
  TDC_TRY(frames_gemm(h, TDC_K_FRONTEND, s, w.fin, c.d_frame_in, h->w_p0, h->b_p0, w.pmid, rb * Tv, D,
                      EPI_BIAS_GELU_BF16, &err));
  if (h->n_cross > 0) {
    if (a.fold) {
      TDC_TRY(frames_gemm(h, TDC_K_KV_GEMM, s, w.pmid, D, h->w_kvf, h->b_kvf, w.kv, rb * Tv, kvw, EPI_BIAS_BF16, &err,
                          64, kv.head_stride, 64));
    } else {
      TDC_TRY(frames_gemm(h, TDC_K_FRONTEND, s, w.pmid, D, h->w_p2, h->b_p2, w.xv, rb * Tv, D, EPI_BIAS_BF16, &err));
      TDC_TRY(frames_gemm(h, TDC_K_KV_GEMM, s, w.xv, D, h->w_ckv, h->b_ckv, w.kv, rb * Tv, kvw, EPI_BIAS_BF16, &err,
                          64, kv.head_stride, 64));
    }
    if (Ta > 0) {
      {
        KernelScope ks(h, TDC_K_FRONTEND, s);
        TDC_TRY(gather_blocks_launch(a.audio, a.row_frames + r0, w.ain, rb, static_cast<long long>(Ta) * c.d_audio * 2,
                                     a.n_frames, s, &err));
      }
      __nv_bfloat16* kv_aud = w.kv + kv.base2 * 64;
      if (a.fold) {
        TDC_TRY(frames_gemm(h, TDC_K_KV_GEMM, s, w.ain, c.d_audio, h->w_kva, h->b_kva, kv_aud, rb * Ta, kvw,
                            EPI_BIAS_BF16, &err, 64, kv.head_stride, 64));
      } else {
        TDC_TRY(frames_gemm(h, TDC_K_FRONTEND, s, w.ain, c.d_audio, h->w_ap, h->b_ap, w.xa, rb * Ta, D, EPI_BIAS_BF16,
                            &err));
        TDC_TRY(frames_gemm(h, TDC_K_KV_GEMM, s, w.xa, D, h->w_ckv, h->b_ckv, kv_aud, rb * Ta, kvw, EPI_BIAS_BF16, &err,
                            64, kv.head_stride, 64));
      }
    }
    KernelScope ks(h, TDC_K_FRONTEND, s);
    TDC_TRY(broadcast_rows_launch(h->kv_newline, 64, kvw / 64, w.kv, kv.head_stride, kv.base3, fc.side, s, &err));
  }
  Workspace qw = carve_workspace(h, w.qws, rb, 0, fc.K, fc.T, false, true);
  return qformer_layers(h, f, r0, rb, qw, kv, nullptr, s);
}

int validate_call(tdc_handle* h, const ForwardCall& f) {
  const tdc_config& c = h->cfg;
  if (!h->loaded) return fail(h, TDC_ESTATE, "forward called before tdc_load_weights");
  if (f.rows < 0 || f.L <= 0 || f.K <= 0 || f.T < 0) return fail(h, TDC_EINVAL, "rows/kv_tokens/num_query/num_text out of range");
  if (f.rows > 0 && (f.query_embeds == nullptr || f.enc == nullptr || f.out == nullptr))
    return fail(h, TDC_EINVAL, "null query_embeds / enc / out pointer");
  if (f.T > 0) {
    if (f.input_ids == nullptr) return fail(h, TDC_EINVAL, "num_text > 0 needs input_ids");
    if (!h->have_text_ffn || !h->have_embeddings)
      return fail(h, TDC_ESTATE, "text input needs the word/position embeddings and the text FFN weights");
    if (f.T > c.max_pos) return fail(h, TDC_EINVAL, "num_text exceeds max_position_embeddings");
  }
  if (f.compress && !h->have_vp) return fail(h, TDC_ESTATE, "tdc_compress needs vision_proj weights (d_out > 0)");
  for (int d : {f.query_dtype, f.enc_dtype, f.out_dtype})
    if (d < TDC_BF16 || d > TDC_F32) return fail(h, TDC_EINVAL, "unknown dtype");
  if (f.rows * static_cast<long long>(f.L) >= (1ll << 31))
    return fail(h, TDC_EINVAL, "rows * kv_tokens must stay below 2^31 per call");
  return TDC_OK;
}

int run_call(tdc_handle* h, const ForwardCall& f, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  const int rc = validate_call(h, f);
  if (rc != TDC_OK) return rc;
  if (f.rows == 0) return TDC_OK;
  if (workspace == nullptr || (reinterpret_cast<uintptr_t>(workspace) & (kAlign - 1)))
    return fail(h, TDC_EINVAL, "workspace must be non-null and 256-byte aligned");
  const bool conv = f.enc_dtype != TDC_BF16;
  // largest row batch that fits the workspace (row cost is linear up to alignment padding)
  const size_t one = carve_workspace(h, nullptr, 1, f.L, f.K, f.T, conv, f.compress).bytes;
  const size_t all = carve_workspace(h, nullptr, f.rows, f.L, f.K, f.T, conv, f.compress).bytes;
  long long batch = f.rows;
  if (all > workspace_bytes) {
    batch = static_cast<long long>(workspace_bytes / one);
    while (batch > 0 && carve_workspace(h, nullptr, batch, f.L, f.K, f.T, conv, f.compress).bytes > workspace_bytes) --batch;
    if (batch <= 0) return fail(h, TDC_EWORKSPACE, "workspace too small for a single row; see tdc_workspace_bytes");
  }
  for (long long r0 = 0; r0 < f.rows; r0 += batch) {
    const int brc = forward_batch(h, f, r0, std::min(batch, f.rows - r0), static_cast<uint8_t*>(workspace), s);
    if (brc != TDC_OK) return brc;
  }
  return TDC_OK;
}

}  // namespace

// =====================================================================================
extern "C" {

int tdc_abi_version(void) { return TDC_B200_ABI_VERSION; }

const char* tdc_last_error(const tdc_handle* h) { return h ? h->last_error.c_str() : g_create_error.c_str(); }

int tdc_create(tdc_handle** out, const tdc_config* cfg) {
  if (out == nullptr || cfg == nullptr) { g_create_error = "null argument"; return TDC_EINVAL; }
  *out = nullptr;
  const tdc_config& c = *cfg;
  if (c.hidden <= 0 || c.heads <= 0 || c.hidden != c.heads * 64) { g_create_error = "hidden must equal heads * 64 (head size is fixed at 64)"; return TDC_EINVAL; }
  if (c.hidden > 1024) { g_create_error = "hidden must be <= 1024"; return TDC_EINVAL; }
  if (c.layers <= 0 || c.cross_freq <= 0 || c.intermediate <= 0 || c.intermediate % 8) { g_create_error = "bad layers / cross_freq / intermediate"; return TDC_EINVAL; }
  if (c.d_enc <= 0 || c.d_enc % 8 || c.d_out < 0 || c.d_out % 8) { g_create_error = "d_enc and d_out must be multiples of 8"; return TDC_EINVAL; }
  if (c.vocab < 0 || (c.vocab > 0 && c.max_pos <= 0)) { g_create_error = "bad vocab / max_pos"; return TDC_EINVAL; }
  if (c.gemm_cta_group < 0 || c.gemm_cta_group > 2) { g_create_error = "gemm_cta_group must be 0, 1 or 2"; return TDC_EINVAL; }
  if (c.d_frame_in < 0 || c.d_frame_in % 8 || c.d_audio < 0 || c.d_audio % 8) { g_create_error = "d_frame_in and d_audio must be multiples of 8"; return TDC_EINVAL; }
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) { g_create_error = "no CUDA device: libtdc_b200 has no CPU fallback"; return TDC_ECUDA; }
  int dev = 0, major = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) { g_create_error = "libtdc_b200 needs an sm_100 (Blackwell) device"; return TDC_ECUDA; }
  tdc_handle* h = new tdc_handle();
  h->cfg = c;
  if (h->cfg.ln_eps <= 0.f) h->cfg.ln_eps = 1e-12f;
  h->layers.resize(c.layers);
  for (int l = 0; l < c.layers; ++l) if (l % c.cross_freq == 0) ++h->n_cross;
  carve_weights(h, nullptr, &h->arena_bytes);
  if (cudaMalloc(&h->arena, h->arena_bytes) != cudaSuccess) {
    g_create_error = "cudaMalloc of the weight arena failed";
    delete h;
    return TDC_ENOMEM;
  }
  size_t dummy = 0;
  carve_weights(h, static_cast<uint8_t*>(h->arena), &dummy);
  *out = h;
  return TDC_OK;
}

int tdc_destroy(tdc_handle* h) {
  if (h == nullptr) return TDC_OK;
  for (auto& pc : h->prof)
    for (auto& e : pc.events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  if (h->arena) cudaFree(h->arena);
  delete h;
  return TDC_OK;
}

int tdc_load_weights(tdc_handle* h, const tdc_tensor* tensors, int32_t count, tdc_stream_t stream) {
  if (h == nullptr) return TDC_EINVAL;
  if (tensors == nullptr || count <= 0) return fail(h, TDC_EINVAL, "empty tensor table");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  std::vector<std::string> got;
  for (int i = 0; i < count; ++i) {
    const tdc_tensor& t = tensors[i];
    if (t.name == nullptr || t.data == nullptr) return fail(h, TDC_EINVAL, "tensor with null name or data");
    Slot slot;
    std::string name = t.name;
    if (!resolve_slot(h, name, &slot)) continue;  // tensors the path does not use (cls.*, position_ids, ...) are ignored
    int64_t numel = 1;
    for (int d = 0; d < t.ndim; ++d) numel *= t.shape[d];
    bool shape_ok = slot.cols == 0 ? (t.ndim == 1 && t.shape[0] == slot.rows)
                                   : (t.ndim == 2 && t.shape[0] == slot.rows && t.shape[1] == slot.cols);
    if (slot.rows == -1) {  // query_tokens [1, K, hidden] or [K, hidden]: K is free (<= kMaxLearnedQueries)
      const int64_t kq = t.ndim == 3 ? t.shape[1] : t.shape[0];
      shape_ok = ((t.ndim == 3 && t.shape[0] == 1 && t.shape[2] == slot.cols) || (t.ndim == 2 && t.shape[1] == slot.cols)) &&
                 kq >= 1 && kq <= kMaxLearnedQueries;
      if (shape_ok) { h->query_tokens_rows = static_cast<int>(kq); h->have_query_tokens = true; }
    }
    if (!shape_ok) return fail(h, TDC_EINVAL, "shape mismatch for tensor " + name);
    const char* err = nullptr;
    const int rc = convert_launch(t.data, t.dtype, slot.dst, slot.as_bf16 ? TDC_BF16 : TDC_F32, numel, s, &err);
    if (rc != TDC_OK) return fail(h, rc, std::string("convert failed for ") + name + ": " + (err ? err : "?"));
    ++h->launches;
    got.push_back(name);
  }
  std::sort(got.begin(), got.end());
  auto have_all = [&](int group, std::string* missing) {
    for (const std::string& n : expected_names(h, group))
      if (!std::binary_search(got.begin(), got.end(), n)) { if (missing) *missing = n; return false; }
    return true;
  };
  std::string missing;
  if (!have_all(0, &missing)) return fail(h, TDC_EINVAL, "missing tensor " + missing);
  h->have_text_ffn = h->cfg.vocab > 0 && have_all(1, nullptr);
  h->have_embeddings = h->cfg.vocab > 0 && have_all(2, nullptr);
  h->have_vp = h->cfg.d_out > 0 && have_all(3, nullptr);
  h->have_frontend = h->cfg.d_frame_in > 0 && h->have_vp && have_all(4, nullptr);
  h->have_audio = h->have_frontend && h->cfg.d_audio > 0 && have_all(5, nullptr);
  if (h->have_frontend) {
    const int rc = fold_frontend_weights(h, s);
    if (rc != TDC_OK) return rc;
  }
  h->loaded = true;
  return TDC_OK;
}

size_t tdc_workspace_bytes(const tdc_handle* h, int32_t rows, int32_t kv_len, int32_t num_query, int32_t num_text) {
  if (h == nullptr || rows <= 0 || kv_len <= 0 || num_query <= 0 || num_text < 0) return 0;
  // sized for the most demanding variant: non-bf16 encoder input + projection output
  return carve_workspace(h, nullptr, rows, kv_len, num_query, num_text, true, h->cfg.d_out > 0).bytes;
}

int tdc_qformer_forward(tdc_handle* h, const void* query_embeds, int32_t query_dtype, const int32_t* query_set,
                        const int64_t* input_ids, const int32_t* text_set, const void* enc, int32_t enc_dtype,
                        const int32_t* kv_len, int32_t rows, int32_t kv_tokens, int32_t num_query, int32_t num_text,
                        void* out_hidden, int32_t out_dtype, void* workspace, size_t workspace_bytes,
                        tdc_stream_t stream) {
  if (h == nullptr) return TDC_EINVAL;
  ForwardCall f{query_embeds, query_dtype, query_set, input_ids, text_set, enc, enc_dtype, kv_len,
                rows, kv_tokens, num_query, num_text, out_hidden, out_dtype, false};
  return run_call(h, f, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int tdc_compress(tdc_handle* h, const void* query_embeds, int32_t query_dtype, const int32_t* query_set,
                 const int64_t* input_ids, const int32_t* text_set, const void* enc, int32_t enc_dtype,
                 const int32_t* kv_len, int32_t rows, int32_t kv_tokens, int32_t num_query, int32_t num_text,
                 void* out, int32_t out_dtype, void* workspace, size_t workspace_bytes, tdc_stream_t stream) {
  if (h == nullptr) return TDC_EINVAL;
  ForwardCall f{query_embeds, query_dtype, query_set, input_ids, text_set, enc, enc_dtype, kv_len,
                rows, kv_tokens, num_query, num_text, out, out_dtype, true};
  return run_call(h, f, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int tdc_compress_multicast(tdc_handle* h, const void* query_embeds, int32_t query_dtype, const int32_t* query_set,
                           const int64_t* input_ids, const int32_t* text_set, const void* enc, int32_t enc_dtype,
                           const int32_t* kv_len, int32_t rows, int32_t kv_tokens, int32_t num_query,
                           int32_t num_text, void* out_multicast, int32_t out_dtype, void* workspace,
                           size_t workspace_bytes, tdc_stream_t stream) {
  if (h == nullptr) return TDC_EINVAL;
  ForwardCall f{query_embeds, query_dtype, query_set, input_ids, text_set, enc, enc_dtype, kv_len,
                rows, kv_tokens, num_query, num_text, out_multicast, out_dtype, true};
  f.multicast = true;
  return run_call(h, f, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t tdc_frames_workspace_bytes(const tdc_handle* h, int32_t n_chunks, int32_t rows, int32_t batch,
                                  int32_t visual_tokens, int32_t audio_tokens, int32_t num_query, int32_t num_text) {
  if (h == nullptr || n_chunks < 0 || rows < 0 || batch <= 0 || visual_tokens <= 0 || audio_tokens < 0 ||
      num_query <= 0 || num_text < 0)
    return 0;
  int side = 1;
  while (side * side < visual_tokens) ++side;
  return carve_frames(h, nullptr, n_chunks, rows, batch, visual_tokens, audio_tokens, side, num_query, num_text).bytes;
}

int tdc_compress_frames(tdc_handle* h, const tdc_frames_args* args, void* workspace, size_t workspace_bytes,
                        tdc_stream_t stream) {
  if (h == nullptr) return TDC_EINVAL;
  if (args == nullptr) return fail(h, TDC_EINVAL, "null tdc_frames_args");
  const tdc_frames_args& a = *args;
  const tdc_config& c = h->cfg;
  if (!h->loaded || !h->have_frontend)
    return fail(h, TDC_ESTATE, "tdc_compress_frames needs d_frame_in > 0 and the mm_projector / image_newline / "
                               "query_proj / vision_proj weights");
  if (c.d_enc != c.d_out) return fail(h, TDC_EINVAL, "tdc_compress_frames needs d_enc == d_out (the LLM width)");
  FramesCall fc{};
  fc.a = args;
  fc.Tv = a.visual_tokens; fc.Ta = a.audio_tokens; fc.K = a.num_query; fc.T = a.num_text;
  fc.learned = a.learned_queries != 0;
  if (a.n_frames < 0 || a.n_chunks < 0 || a.rows < 0 || fc.Tv <= 0 || fc.Ta < 0 || fc.K <= 0 || fc.T < 0)
    return fail(h, TDC_EINVAL, "bad frame / chunk / row / token counts");
  fc.side = 1;
  while (fc.side * fc.side < fc.Tv) ++fc.side;
  if (fc.side * fc.side != fc.Tv) return fail(h, TDC_EINVAL, "visual_tokens must be a square grid (side * side)");
  if (fc.Ta > 0 && (!h->have_audio || a.audio == nullptr))
    return fail(h, TDC_ESTATE, "audio_tokens > 0 needs d_audio > 0, the audio_proj weights and an audio pointer");
  if (fc.learned && (!h->have_query_tokens || h->query_tokens_rows != fc.K))
    return fail(h, TDC_ESTATE, "learned_queries needs a loaded `query_tokens` with num_query rows");
  if (fc.T > 0) {
    if (a.input_ids == nullptr) return fail(h, TDC_EINVAL, "num_text > 0 needs input_ids");
    if (!h->have_text_ffn || !h->have_embeddings)
      return fail(h, TDC_ESTATE, "text input needs the word/position embeddings and the text FFN weights");
    if (fc.T > c.max_pos) return fail(h, TDC_EINVAL, "num_text exceeds max_position_embeddings");
  }
  if (a.out_dtype < TDC_BF16 || a.out_dtype > TDC_F32) return fail(h, TDC_EINVAL, "unknown dtype");
  if (a.n_chunks == 0 && a.rows == 0) return TDC_OK;
  if (a.frames == nullptr || (a.n_chunks > 0 && a.static_frames == nullptr) ||
      (a.rows > 0 && (a.row_frames == nullptr || a.row_chunk == nullptr || a.out == nullptr)))
    return fail(h, TDC_EINVAL, "null frames / index / out pointer");
  if (workspace == nullptr || (reinterpret_cast<uintptr_t>(workspace) & (kAlign - 1)))
    return fail(h, TDC_EINVAL, "workspace must be non-null and 256-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);

  // internal batch: the largest item count (<= 65535, the gather grid) whose buffers fit the workspace
  const long long most = std::max<long long>(a.rows, a.n_chunks);
  long long nb = std::min<long long>(most, 65535);
  const bool fold = a.fold != 0;
  auto need = [&](long long n) { return carve_frames(h, nullptr, a.n_chunks, a.rows, n, fc.Tv, fc.Ta, fc.side, fc.K, fc.T, fold).bytes; };
  if (need(nb) > workspace_bytes) {
    const size_t fixed = need(0), one = need(1) - fixed;
    if (workspace_bytes < fixed + one) return fail(h, TDC_EWORKSPACE, "workspace too small for a single row; see tdc_frames_workspace_bytes");
    nb = std::min<long long>(nb, static_cast<long long>((workspace_bytes - fixed) / one));
    while (nb > 0 && need(nb) > workspace_bytes) --nb;
    if (nb <= 0) return fail(h, TDC_EWORKSPACE, "workspace too small for a single row; see tdc_frames_workspace_bytes");
  }
  if (const char* cap = std::getenv("TDC_FRAMES_BATCH")) {   // dev knob: cap the internal batch (rows / key frames)
    const long long v = std::atoll(cap);
    if (v > 0) nb = std::min(nb, v);
  }
  FramesWorkspace w = carve_frames(h, static_cast<uint8_t*>(workspace), a.n_chunks, a.rows, nb, fc.Tv, fc.Ta, fc.side,
                                   fc.K, fc.T, fold);

  // pass 1: key frames (queries of every chunk, pass-through tokens)
  if (a.n_chunks > 0 && (!fc.learned || a.static_out != nullptr))
    for (long long c0 = 0; c0 < a.n_chunks; c0 += nb) {
      const int rc = frames_static_batch(h, fc, w, c0, std::min<long long>(nb, a.n_chunks - c0), s);
      if (rc != TDC_OK) return rc;
    }
  if (a.static_ready_event != nullptr && cudaEventRecord(static_cast<cudaEvent_t>(a.static_ready_event), s) != cudaSuccess)
    return fail(h, TDC_ECUDA, "cudaEventRecord(static_ready_event) failed");
  if (a.rows == 0) return TDC_OK;
  // pass 2: dynamic frames
  const bool zero_map = fc.learned || fc.T > 0;
  if (zero_map && cudaMemsetAsync(w.zero_map, 0, static_cast<size_t>(std::max(a.rows, a.n_chunks)) * sizeof(int32_t),
                                  s) != cudaSuccess)
    return fail(h, TDC_ECUDA, "cudaMemsetAsync failed");
  // several prompts: row -> chunk -> prompt (the chunk's video)
  const bool multi_prompt = fc.T > 0 && a.chunk_prompt != nullptr && a.n_prompts > 1;
  if (multi_prompt) {
    const char* err = nullptr;
    const int rc = compose_index_launch(a.chunk_prompt, a.n_chunks, a.row_chunk, w.row_prompt, a.rows, s, &err);
    if (rc != TDC_OK) return fail(h, rc, err ? err : "compose_index failed");
  }
  ForwardCall f{};
  f.query_embeds = fc.learned ? h->query_tokens : w.q_sets;
  f.query_dtype = TDC_F32;
  f.query_set = fc.learned ? w.zero_map : a.row_chunk;
  f.input_ids = fc.T > 0 ? a.input_ids : nullptr;
  f.text_set = fc.T > 0 ? (multi_prompt ? w.row_prompt : w.zero_map) : nullptr;
  f.enc = nullptr; f.enc_dtype = TDC_BF16; f.kv_len = nullptr;
  f.rows = a.rows; f.L = fc.Tv + fc.Ta + fc.side; f.K = fc.K; f.T = fc.T;
  f.out = a.out; f.out_dtype = a.out_dtype; f.compress = true; f.multicast = a.multicast != 0;
  f.num_query_sets = fc.learned ? 1 : a.n_chunks;
  f.num_text_sets = std::max(a.n_prompts, 1);
  // layer-0 de-duplication: every row of a chunk has the same queries and the same prompt, hence the same state up
  // to and including layer 0's self-attention block — compute it once per chunk (once in total for learned queries)
  // and broadcast it.  Same kernels on the same values: bit-identical to the per-row computation.
  if (!a.no_layer0_dedup && c.layers > 1) {
    // one state per chunk — or ONE in total when nothing distinguishes the chunks (learned queries, one prompt)
    const bool per_chunk = !fc.learned || multi_prompt;
    const long long n_sets = per_chunk ? a.n_chunks : 1;
    ForwardCall p = f;
    p.rows = n_sets;
    p.query_set = fc.learned ? w.zero_map : nullptr;   // "row" i of the pre-pass uses query set i (set 0 if learned)
    if (multi_prompt) p.text_set = a.chunk_prompt;      // ... and its chunk's prompt
    p.l0_out = w.l0_sets;
    Workspace qw0{};
    for (long long s0 = 0; s0 < n_sets; s0 += nb) {
      const long long sb = std::min<long long>(nb, n_sets - s0);
      qw0 = carve_workspace(h, w.qws, sb, 0, fc.K, fc.T, false, true);
      const int rc = qformer_layers(h, p, s0, sb, qw0, KvView{}, nullptr, s);
      if (rc != TDC_OK) return rc;
    }
    f.l0_sets = w.l0_sets;
    f.l0_map = per_chunk ? a.row_chunk : nullptr;
    f.l0_num_sets = static_cast<int>(n_sets);
  }
  for (long long r0 = 0; r0 < a.rows; r0 += nb) {
    const int rc = frames_dynamic_batch(h, fc, f, w, r0, std::min<long long>(nb, a.rows - r0), s);
    if (rc != TDC_OK) return rc;
  }
  return TDC_OK;
}

int tdc_proj_norm(tdc_handle* h, const void* hidden, int32_t hidden_dtype, int32_t rows, int32_t tokens_per_row,
                  int32_t num_query, void* out, int32_t out_dtype, void* workspace, size_t workspace_bytes,
                  tdc_stream_t stream) {
  if (h == nullptr) return TDC_EINVAL;
  if (!h->loaded || !h->have_vp) return fail(h, TDC_ESTATE, "tdc_proj_norm needs loaded vision_proj weights");
  if (rows < 0 || num_query <= 0 || tokens_per_row < num_query) return fail(h, TDC_EINVAL, "bad rows / tokens");
  if (rows == 0) return TDC_OK;
  if (hidden == nullptr || out == nullptr || workspace == nullptr) return fail(h, TDC_EINVAL, "null pointer");
  const tdc_config& c = h->cfg;
  const size_t M = static_cast<size_t>(rows) * num_query;
  Carver cv{static_cast<uint8_t*>(workspace)};
  __nv_bfloat16* x = cv.take<__nv_bfloat16>(M * c.hidden);
  float* y = cv.take<float>(M * c.d_out);
  if (cv.off > workspace_bytes) return fail(h, TDC_EWORKSPACE, "workspace too small for tdc_proj_norm");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const char* err = nullptr;
  {
    KernelScope ks(h, TDC_K_ROWOPS, s);
    TDC_TRY(take_query_tokens_launch(hidden, hidden_dtype, rows, tokens_per_row, num_query, c.hidden, x, s, &err));
  }
  TDC_TRY(gemm(h, TDC_K_QUERY_GEMM, s, x, c.hidden, h->w_vp, c.hidden, h->b_vp, y, c.d_out, M, c.d_out, c.hidden,
               EPI_BIAS_F32, &err));
  KernelScope ks(h, TDC_K_ROWOPS, s);
  TDC_TRY(l2_normalize_launch(y, c.d_out, out, out_dtype, M, c.d_out, false, s, &err));
  return TDC_OK;
}

int tdc_linear(const void* x, const void* w, const float* bias, void* y, int32_t m, int32_t n, int32_t k,
               int32_t out_dtype, int32_t gelu, int32_t cta_group, tdc_stream_t stream) {
  if (m == 0) return TDC_OK;
  if (x == nullptr || w == nullptr || y == nullptr || m < 0) { g_create_error = "tdc_linear: null pointer"; return TDC_EINVAL; }
  GemmProblem p;
  p.a = x; p.lda = k; p.w = w; p.ldw = k; p.bias = bias; p.out = y; p.ldo = n; p.m = m; p.n = n; p.k = k;
  if (out_dtype == TDC_F32) {
    if (gelu) { g_create_error = "tdc_linear: gelu needs bf16 output"; return TDC_EINVAL; }
    p.mode = EPI_BIAS_F32;
  } else if (out_dtype == TDC_BF16) {
    p.mode = gelu ? EPI_BIAS_GELU_BF16 : EPI_BIAS_BF16;
  } else { g_create_error = "tdc_linear: out_dtype must be bf16 or fp32"; return TDC_EINVAL; }
  p.cta_group = cta_group == 0 ? 2 : cta_group;
  const char* err = nullptr;
  const int rc = gemm_launch(p, static_cast<cudaStream_t>(stream), &err);
  if (rc != TDC_OK) g_create_error = err ? err : "tdc_linear failed";
  return rc;
}

int tdc_linear_layernorm(const void* x, const void* w, const float* bias, const float* resid, const float* gamma,
                         const float* beta, float eps, float* y_f32, void* y_bf16, int32_t m, int32_t n, int32_t k,
                         tdc_stream_t stream) {
  if (m == 0) return TDC_OK;
  GemmLnProblem p;
  p.a = x; p.lda = k; p.w = w; p.ldw = k; p.m = m; p.n = n; p.k = k; p.bias = bias; p.resid = resid; p.ldr = n;
  p.gamma = gamma; p.beta = beta; p.eps = eps > 0.f ? eps : 1e-12f; p.out_f32 = y_f32; p.out_bf16 = y_bf16; p.ldo = n;
  const char* err = nullptr;
  const int rc = gemm_ln_launch(p, static_cast<cudaStream_t>(stream), &err);
  if (rc != TDC_OK) g_create_error = err ? err : "tdc_linear_layernorm failed";
  return rc;
}

int tdc_gelu_mlp(const void* x, const void* w0, const float* b0, const void* w1, const float* b1, void* mid, void* y,
                 int32_t m, int32_t d_in, int32_t d_mid, int32_t d_out, tdc_stream_t stream) {
  int rc = tdc_linear(x, w0, b0, mid, m, d_mid, d_in, TDC_BF16, 1, 0, stream);
  if (rc != TDC_OK) return rc;
  return tdc_linear(mid, w1, b1, y, m, d_out, d_mid, TDC_BF16, 0, 0, stream);
}

int tdc_avg_pool_tokens(const void* frames, int32_t dtype, int32_t n, int32_t tokens, int32_t d, int32_t num_query,
                        void* out_bf16, tdc_stream_t stream) {
  const char* err = nullptr;
  const int rc = avg_pool_tokens_launch(frames, dtype, n, tokens, d, num_query, static_cast<__nv_bfloat16*>(out_bf16),
                                        static_cast<cudaStream_t>(stream), &err);
  if (rc != TDC_OK) g_create_error = err ? err : "tdc_avg_pool_tokens failed";
  return rc;
}

int tdc_layernorm(const float* x, const float* resid, int32_t resid_period, const float* gamma, const float* beta,
                  float eps, float* y_f32, void* y_bf16, int64_t rows, int32_t width, tdc_stream_t stream) {
  if (rows == 0) return TDC_OK;
  if (x == nullptr || gamma == nullptr || beta == nullptr || (y_f32 == nullptr && y_bf16 == nullptr) || rows < 0 ||
      resid_period < 0) {
    g_create_error = "tdc_layernorm: null pointer / bad argument";
    return TDC_EINVAL;
  }
  const char* err = nullptr;
  const int rc = layernorm_launch(x, width, resid, width, gamma, beta, eps, y_f32, static_cast<__nv_bfloat16*>(y_bf16),
                                  width, rows, width, static_cast<cudaStream_t>(stream), &err, resid_period);
  if (rc != TDC_OK) g_create_error = err ? err : "tdc_layernorm failed";
  return rc;
}

int tdc_attention(const void* q, const void* k, const void* v, void* out, int64_t ldq, int64_t ldk, int64_t ldv,
                  int64_t ldo, int32_t rows, int32_t heads, int32_t q_seg1, int32_t q_seg2, int64_t q_base1,
                  int64_t q_base2, int32_t kv_seg1, int32_t kv_seg2, int64_t kv_base1, int64_t kv_base2,
                  const int32_t* kv_len, const uint32_t* kv_mask, tdc_stream_t stream) {
  if (rows == 0) return TDC_OK;
  if (q == nullptr || k == nullptr || v == nullptr || out == nullptr || q_seg1 < 0 || q_seg2 < 0 || kv_seg1 < 0 ||
      kv_seg2 < 0) {
    g_create_error = "tdc_attention: null pointer / bad argument";
    return TDC_EINVAL;
  }
  AttentionArgs a;
  a.q = static_cast<const __nv_bfloat16*>(q); a.k = static_cast<const __nv_bfloat16*>(k);
  a.v = static_cast<const __nv_bfloat16*>(v); a.out = static_cast<__nv_bfloat16*>(out);
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo;
  a.rows = rows; a.heads = heads; a.nq = q_seg1 + q_seg2;
  a.q_seg1 = q_seg1; a.q_seg2 = q_seg2; a.q_base1 = q_base1; a.q_base2 = q_base2;
  a.kv_seg1 = kv_seg1; a.kv_seg2 = kv_seg2; a.kv_base1 = kv_base1; a.kv_base2 = kv_base2;
  a.kv_len = kv_len; a.kv_mask = kv_mask;
  a.scale_log2 = 1.4426950408889634f / 8.0f;
  const char* err = nullptr;
  const int rc = attention_launch(a, static_cast<cudaStream_t>(stream), &err);
  if (rc != TDC_OK) g_create_error = err ? err : "tdc_attention failed";
  return rc;
}

int tdc_resize_tokens_bilinear(const void* in, int32_t in_dtype, int32_t bs, int32_t side_in, int32_t side_out,
                               int32_t d, void* out, int32_t out_dtype, tdc_stream_t stream) {
  if (bs == 0) return TDC_OK;
  if (in == nullptr || out == nullptr || bs < 0 || in_dtype < TDC_BF16 || in_dtype > TDC_F32 || out_dtype < TDC_BF16 ||
      out_dtype > TDC_F32) {
    g_create_error = "tdc_resize_tokens_bilinear: null pointer / bad argument";
    return TDC_EINVAL;
  }
  const char* err = nullptr;
  const int rc = resize_tokens_bilinear_launch(in, in_dtype, bs, side_in, side_out, d, out, out_dtype,
                                               static_cast<cudaStream_t>(stream), &err);
  if (rc != TDC_OK) g_create_error = err ? err : "tdc_resize_tokens_bilinear failed";
  return rc;
}

int tdc_window_rearrange(const void* in, int32_t in_dtype, int32_t bs, int32_t q, int32_t r, int32_t d, void* out_bf16,
                         tdc_stream_t stream) {
  if (bs == 0) return TDC_OK;
  if (in == nullptr || out_bf16 == nullptr || bs < 0 || q <= 0 || r <= 0 || in_dtype < TDC_BF16 || in_dtype > TDC_F32) {
    g_create_error = "tdc_window_rearrange: null pointer / bad argument";
    return TDC_EINVAL;
  }
  const char* err = nullptr;
  const int rc = window_rearrange_launch(in, in_dtype, bs, q, r, d, static_cast<__nv_bfloat16*>(out_bf16),
                                         static_cast<cudaStream_t>(stream), &err);
  if (rc != TDC_OK) g_create_error = err ? err : "tdc_window_rearrange failed";
  return rc;
}

int tdc_combine_parts(const float* base, const float* parts, const float* logits, int32_t ld_logits, int32_t num_parts,
                      int64_t rows, int32_t width, float* out, tdc_stream_t stream) {
  if (rows == 0) return TDC_OK;
  if (base == nullptr || parts == nullptr || logits == nullptr || out == nullptr || rows < 0) {
    g_create_error = "tdc_combine_parts: null pointer / bad argument";
    return TDC_EINVAL;
  }
  const char* err = nullptr;
  const int rc = combine_parts_launch(base, parts, logits, ld_logits, num_parts, rows, width, out,
                                      static_cast<cudaStream_t>(stream), &err);
  if (rc != TDC_OK) g_create_error = err ? err : "tdc_combine_parts failed";
  return rc;
}

int tdc_multicast_copy(const void* src, void* dst_multicast, size_t bytes, int32_t ctas, tdc_stream_t stream) {
  if (bytes == 0) return TDC_OK;
  if (src == nullptr || dst_multicast == nullptr) {
    g_create_error = "tdc_multicast_copy: null pointer";
    return TDC_EINVAL;
  }
  const char* err = nullptr;
  const int rc = multicast_copy_launch(src, dst_multicast, bytes, ctas, static_cast<cudaStream_t>(stream), &err);
  if (rc != TDC_OK) g_create_error = err ? err : "tdc_multicast_copy failed";
  return rc;
}

int tdc_peer_copy(const void* src, void* dst, size_t bytes, tdc_stream_t stream) {
  if (bytes == 0) return TDC_OK;
  if (src == nullptr || dst == nullptr) {
    g_create_error = "tdc_peer_copy: null pointer";
    return TDC_EINVAL;
  }
  if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)) != cudaSuccess) {
    g_create_error = "tdc_peer_copy: cudaMemcpyAsync failed";
    return TDC_ECUDA;
  }
  return TDC_OK;
}

int tdc_residual_add(const float* a, const float* b, float* out_f32, void* out_bf16, int64_t count,
                     tdc_stream_t stream) {
  if (count == 0) return TDC_OK;
  if (a == nullptr || b == nullptr || (out_f32 == nullptr && out_bf16 == nullptr) || count < 0) {
    g_create_error = "tdc_residual_add: null pointer / bad argument";
    return TDC_EINVAL;
  }
  const char* err = nullptr;
  const int rc = residual_add_launch(a, b, out_f32, static_cast<__nv_bfloat16*>(out_bf16), count,
                                     static_cast<cudaStream_t>(stream), &err);
  if (rc != TDC_OK) g_create_error = err ? err : "tdc_residual_add failed";
  return rc;
}

size_t tdc_segment_workspace_bytes(int32_t n_frames, int64_t dim) {
  if (n_frames < 2 || dim <= 0) return 0;
  return static_cast<size_t>(n_frames - 1) * frame_cosine_slices(dim) * 3 * sizeof(float);
}

int tdc_segment_boundaries(const void* feats, int32_t dtype, int32_t n_frames, int64_t dim, int32_t max_segments,
                           float* cos_out, int64_t* boundaries_out, void* workspace, size_t workspace_bytes,
                           tdc_stream_t stream) {
  if (n_frames < 2) return TDC_OK;
  if (feats == nullptr || cos_out == nullptr || boundaries_out == nullptr || workspace == nullptr) {
    g_create_error = "tdc_segment_boundaries: null pointer";
    return TDC_EINVAL;
  }
  if (dtype < TDC_BF16 || dtype > TDC_F32 || max_segments <= 0) { g_create_error = "tdc_segment_boundaries: bad dtype / max_segments"; return TDC_EINVAL; }
  if (workspace_bytes < tdc_segment_workspace_bytes(n_frames, dim)) { g_create_error = "tdc_segment_boundaries: workspace too small"; return TDC_EWORKSPACE; }
  const char* err = nullptr;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = frame_cosine_launch(feats, dtype, n_frames, dim, static_cast<float*>(workspace), cos_out, s, &err);
  if (rc == TDC_OK)
    rc = select_smallest_launch(cos_out, n_frames - 1, max_segments, reinterpret_cast<long long*>(boundaries_out), s, &err);
  if (rc != TDC_OK) g_create_error = err ? err : "tdc_segment_boundaries failed";
  return rc;
}

int tdc_convert(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, int64_t count, tdc_stream_t stream) {
  const char* err = nullptr;
  const int rc = convert_launch(src, src_dtype, dst, dst_dtype, count, static_cast<cudaStream_t>(stream), &err);
  if (rc != TDC_OK) g_create_error = err ? err : "tdc_convert failed";
  return rc;
}

int tdc_set_profiling(tdc_handle* h, int32_t enabled) {
  if (h == nullptr) return TDC_EINVAL;
  h->profiling = enabled != 0;
  return TDC_OK;
}

int tdc_reset_profile(tdc_handle* h) {
  if (h == nullptr) return TDC_EINVAL;
  for (auto& pc : h->prof) pc.used = 0;
  return TDC_OK;
}

int tdc_get_profile(tdc_handle* h, int32_t kernel_class, double* total_ms, int64_t* launches) {
  if (h == nullptr || kernel_class < 0 || kernel_class >= TDC_K_COUNT) return TDC_EINVAL;
  ProfileClass& pc = h->prof[kernel_class];
  double tot = 0;
  for (size_t i = 0; i < pc.used; ++i) {
    if (cudaEventSynchronize(pc.events[i].second) != cudaSuccess) return fail(h, TDC_ECUDA, "event synchronize failed");
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, pc.events[i].first, pc.events[i].second) != cudaSuccess)
      return fail(h, TDC_ECUDA, "event elapsed time failed");
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = static_cast<int64_t>(pc.used);
  return TDC_OK;
}

int64_t tdc_launch_count(const tdc_handle* h) { return h ? h->launches : 0; }

}  // extern "C"
