"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the reference TDC compression path.

This module is the *checker*: `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it; the product (`tdc_video_b200/`) never
does and fails loudly when libtdc_b200.so or a GPU is missing.

It restates, function by function, what the reference computes on this path, as plain
torch fp32 ops on explicit weight dictionaries (no nn.Module of the reference is copied):

    qformer_forward      <- tdc/Qformer.py:804-965 (BertModel.forward) with
                            :78-108 embeddings, :402-474 layer, :169-275 attention,
                            :285-289 / :371-375 output blocks, :358-361 intermediate
    proj_norm            <- tdc/cambrian_arch.py:1664-1667
    avg_pool_queries     <- tdc/cambrian_arch.py:1633-1638

PINNING: the restatement is checked against the reference's own modules (imported unmodified
through oracle/ref_shim.py) by tests/test_oracle_pinning.py in the build container, and
against the committed golden outputs tests/golden/*.npz (generated from the reference by
oracle/make_golden.py) everywhere else.  The reference itself ships no golden vectors or
tests for this path, so outputs of the reference run here are the anchor.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F


def _t(x) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.detach().to(torch.float32).cpu()
    return torch.as_tensor(x, dtype=torch.float32)


def _linear(sd, prefix, x):
    return F.linear(x, _t(sd[prefix + ".weight"]), _t(sd[prefix + ".bias"]))


def _layer_norm(sd, prefix, x, eps):
    return F.layer_norm(x, (x.shape[-1],), _t(sd[prefix + ".weight"]), _t(sd[prefix + ".bias"]), eps)


def _attention(sd, prefix, hidden, heads, kv_source=None, kv_len=None):
    """BertSelfAttention.forward (tdc/Qformer.py:169-275), absolute position embeddings,
    eval mode.  `kv_source` None = self-attention.  All-ones masks on the TDC path add an
    exact 0 to the scores (Qformer.py:798-802, 920-926); `kv_len` is the one masking
    generalisation offered (scores of padded keys get the same finfo.min additive mask
    `invert_attention_mask` would produce)."""
    B, n, H = hidden.shape
    dh = H // heads
    src = hidden if kv_source is None else kv_source
    q = _linear(sd, prefix + ".query", hidden).view(B, n, heads, dh).permute(0, 2, 1, 3)
    k = _linear(sd, prefix + ".key", src).view(B, src.shape[1], heads, dh).permute(0, 2, 1, 3)
    v = _linear(sd, prefix + ".value", src).view(B, src.shape[1], heads, dh).permute(0, 2, 1, 3)
    scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dh)
    if kv_len is not None:
        pos = torch.arange(src.shape[1])[None, :]
        pad = (pos >= torch.as_tensor(kv_len)[:, None]).to(torch.float32)
        scores = scores + (pad * torch.finfo(torch.float32).min)[:, None, None, :]
    probs = torch.softmax(scores, dim=-1)
    ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).reshape(B, n, H)
    return ctx


def qformer_forward(sd: Dict[str, object], geom, query_embeds, encoder_hidden_states, input_ids=None,
                    kv_len=None) -> torch.Tensor:
    """last_hidden_state [B, K+T, H] of `Qformer.bert(input_ids=, query_embeds=,
    encoder_hidden_states=, encoder_attention_mask=ones, use_cache=False, return_dict=True)`
    (call site tdc/cambrian_arch.py:1653-1662)."""
    q = _t(query_embeds)
    enc = _t(encoder_hidden_states)
    B, K, H = q.shape
    eps = geom.ln_eps
    # --- BertEmbeddings.forward, Qformer.py:78-108: queries carry no position embedding
    if input_ids is not None and input_ids.shape[1] > 0:
        ids = torch.as_tensor(input_ids, dtype=torch.long)
        T = ids.shape[1]
        text = F.embedding(ids, _t(sd["embeddings.word_embeddings.weight"])) + \
            _t(sd["embeddings.position_embeddings.weight"])[:T][None]
        x = torch.cat([q, text], dim=1)
    else:
        T = 0
        x = q
    x = _layer_norm(sd, "embeddings.LayerNorm", x, eps)
    # --- BertEncoder / BertLayer, Qformer.py:495-589, 402-474
    for l in range(geom.layers):
        p = f"encoder.layer.{l}."
        # self-attention over all K+T tokens + BertSelfOutput (post-LN), :417-423, :285-289
        ctx = _attention(sd, p + "attention.self", x, geom.heads)
        x = _layer_norm(sd, p + "attention.output.LayerNorm", _linear(sd, p + "attention.output.dense", ctx) + x, eps)
        xq = x[:, :K]
        # cross-attention for the query tokens in layers l % cross_freq == 0, :430-447
        if l % geom.cross_freq == 0:
            ctx = _attention(sd, p + "crossattention.self", xq, geom.heads, kv_source=enc, kv_len=kv_len)
            xq = _layer_norm(sd, p + "crossattention.output.LayerNorm",
                             _linear(sd, p + "crossattention.output.dense", ctx) + xq, eps)
        # query FFN (intermediate_query/output_query), :449-454, :481-484; exact-erf GELU
        mid = F.gelu(_linear(sd, p + "intermediate_query.dense", xq))
        xq = _layer_norm(sd, p + "output_query.LayerNorm", _linear(sd, p + "output_query.dense", mid) + xq, eps)
        if T > 0:
            # text tokens skip cross-attention and use intermediate/output, :455-462, :476-479
            xt = x[:, K:]
            mid = F.gelu(_linear(sd, p + "intermediate.dense", xt))
            xt = _layer_norm(sd, p + "output.LayerNorm", _linear(sd, p + "output.dense", mid) + xt, eps)
            x = torch.cat([xq, xt], dim=1)
        else:
            x = xq
    return x


def proj_norm(sd: Dict[str, object], hidden, num_query: int) -> torch.Tensor:
    """F.normalize(vision_proj(last_hidden_state[:, :K]), dim=-1) — cambrian_arch.py:1664-1667."""
    h = _t(hidden)[:, :num_query]
    return F.normalize(_linear(sd, "vision_proj", h), dim=-1)


def compress(sd, geom, query_embeds, encoder_hidden_states, input_ids=None, kv_len=None) -> torch.Tensor:
    K = _t(query_embeds).shape[1]
    return proj_norm(sd, qformer_forward(sd, geom, query_embeds, encoder_hidden_states, input_ids, kv_len), K)


def avg_pool_queries(key_frame, num_query: int) -> torch.Tensor:
    """adaptive_avg_pool1d over the token axis of a static frame [.., tokens, d] -> [.., K, d]
    (cambrian_arch.py:1633-1637; bins [floor(i*L/K), ceil((i+1)*L/K)))."""
    x = _t(key_frame)
    return F.adaptive_avg_pool1d(x.transpose(-1, -2), num_query).transpose(-1, -2)


def gelu_mlp(w0, b0, w1, b1, x) -> torch.Tensor:
    """mm_projector = Linear -> GELU(erf) -> Linear (cambrian_arch.py:65-69)."""
    return F.linear(F.gelu(F.linear(_t(x), _t(w0), _t(b0))), _t(w1), _t(b1))


# ---- parity metrics -------------------------------------------------------
def parity_metrics(test, ref) -> Dict[str, float]:
    """per-token cosine (min), per-token normalised L2 error (max), max|err|/max|ref|."""
    a = _t(test).reshape(-1, _t(test).shape[-1]).double()
    b = _t(ref).reshape(-1, _t(ref).shape[-1]).double()
    cos = F.cosine_similarity(a, b, dim=-1)
    rel_tok = (a - b).norm(dim=-1) / b.norm(dim=-1).clamp_min(1e-30)
    return {
        "min_cos": float(cos.min()),
        "max_tok_rel_l2": float(rel_tok.max()),
        "max_abs_over_max_ref": float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)),
    }
