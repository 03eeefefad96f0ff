"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the Spatial Vision Aggregator (SVA) connector,
the learned-query cross-attention that produces the frame tokens the TDC path consumes.

Restated, as plain torch ops on a dict of tensors keyed by the reference's parameter names:

    mm_projector_aux         <- tdc/cambrian_arch.py:83-93, applied at :1002-1013
                                (Linear . GELU . Linear . LayerNorm per vision tower)
    window rearrangement     <- cambrian_arch.py:601-690 rearrange_vision_tower_features_inference
    attention masks          <- cambrian_arch.py:487-509 unmask_attention_mask + :647-669
    VisionCrossAttentionLayer<- tdc/vision_sampler.py:305-401
    MultiKVCrossAttention    <- tdc/vision_sampler.py:170-291 (SDPA with a boolean mask)
    VisionTokenSampler       <- tdc/vision_sampler.py:519-566 ("joint" layers)
    sva_frames               <- cambrian_arch.py:1002-1053 for one query group whose side equals the
                                final grid side (query_num_list == [image_token_len], the shipped config)

PINNING: tests/test_sva_pinning.py compares every function with the reference's own modules
(`tdc/vision_sampler.py` imports only torch/numpy, so it is loaded unmodified) and with the real
`rearrange_vision_tower_features_inference`; tests/golden/sva_*.npz hold reference outputs for the GPU box.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
import torch.nn.functional as F

from .qformer_oracle import _t

LN_EPS = 1e-5  # nn.LayerNorm default (vision_sampler.py / cambrian_arch.py never override it)


def _lin(sd, prefix, x, bias=True):
    return F.linear(x, _t(sd[prefix + ".weight"]), _t(sd[prefix + ".bias"]) if bias else None)


def _ln(sd, prefix, x):
    return F.layer_norm(x, (x.shape[-1],), _t(sd[prefix + ".weight"]), _t(sd[prefix + ".bias"]), LN_EPS)


def mm_projector_aux(sd, prefix, x):
    """nn.Sequential(Linear, GELU, Linear, LayerNorm) — cambrian_arch.py:83-93."""
    h = F.gelu(_lin(sd, prefix + ".0", _t(x)))
    return _ln(sd, prefix + ".3", _lin(sd, prefix + ".2", h))


def window_masks(image_size: Tuple[int, int], grid: int, query_side: int) -> torch.Tensor:
    """Boolean [query_side^2, (grid/query_side)^2] mask of one frame: True = real image content
    (cambrian_arch.py:487-509 and :619-669; rows that would be all-False are set all-True)."""
    w, h = image_size
    mask = torch.ones((1, grid, grid), dtype=torch.bool)
    if (w / h) > 1.0:                       # original_aspect_ratio > current (square grid)
        new_h = int(h * (grid / w))
        pad = (grid - new_h) // 2
        if pad > 0:
            mask[:, :pad, :] = 0
            mask[:, -pad:, :] = 0
    else:
        new_w = int(w * (grid / h))
        pad = (grid - new_w) // 2
        if pad > 0:
            mask[:, :, :pad] = 0
            mask[:, :, -pad:] = 0
    r = grid // query_side
    m = mask.view(1, query_side, r, query_side, r).permute(0, 1, 3, 2, 4).contiguous().flatten(0, 2).flatten(1, 2)
    m[m.sum(-1) == 0] = True
    return m


def rearrange_windows(feat: torch.Tensor, query_side: int) -> torch.Tensor:
    """[bs, grid^2, C] -> [bs * query_side^2, (grid/query_side)^2, C]: the tokens under every query
    (cambrian_arch.py:624-645, unpad=False)."""
    bs, n, c = feat.shape
    grid = int(n ** 0.5)
    r = grid // query_side
    assert r * query_side == grid
    x = feat.view(bs, query_side, r, query_side, r, c).permute(0, 1, 3, 2, 4, 5).contiguous()
    return x.view(bs * query_side * query_side, r * r, c)


def cross_attention_layer(sd, p, queries, context, latents: Sequence[torch.Tensor], masks: Sequence[torch.Tensor],
                          num_heads: int = 16) -> torch.Tensor:
    """VisionCrossAttentionLayer.forward (vision_sampler.py:343-401) with MultiKVCrossAttention (:219-291).
    queries [R, q_len, q_dim], context [R, q_len, ctx_dim], latents[t] [R, n_t, kv_dim], masks[t] bool [R, n_t]."""
    residual = queries
    ctx = _lin(sd, p + "proj_context", context, bias=False)
    q = _lin(sd, p + "proj_in", torch.cat([queries, ctx], -1), bias=False)
    lat = []
    for t, v in enumerate(latents):
        if v.shape[1] > 1:
            v = v + _t(sd[p + f"pos_embed_{t}"])[None]
        lat.append(v)
    # --- MultiKVCrossAttention
    R, q_len, hidden = q.shape
    dh = hidden // num_heads
    qs = _lin(sd, p + "cross_attn.q_proj.1", _ln(sd, p + "cross_attn.q_proj.0", q), bias=False)
    ks = torch.cat([_lin(sd, p + f"cross_attn.k_proj_{t}.1", _ln(sd, p + f"cross_attn.k_proj_{t}.0", v), bias=False)
                    for t, v in enumerate(lat)], dim=1)
    vs = torch.cat([_lin(sd, p + f"cross_attn.v_proj_{t}.1", _ln(sd, p + f"cross_attn.v_proj_{t}.0", v), bias=False)
                    for t, v in enumerate(lat)], dim=1)
    n_kv = ks.shape[1]
    qh = qs.view(R, q_len, num_heads, dh).transpose(1, 2)
    kh = ks.view(R, n_kv, num_heads, dh).transpose(1, 2)
    vh = vs.view(R, n_kv, num_heads, dh).transpose(1, 2)
    mask = torch.cat([m.view(R, 1, 1, -1).expand(-1, -1, q_len, -1) for m in masks], dim=-1)
    att = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=mask)
    att = att.transpose(1, 2).reshape(R, q_len, hidden)
    att = _lin(sd, p + "cross_attn.o_proj", att, bias=False)
    # --- back in the layer
    q = _ln(sd, p + "norm", q + att)
    q = _lin(sd, p + "proj_out.linear_2", F.gelu(_lin(sd, p + "proj_out.linear_1", q, bias=False)), bias=False)
    return q + residual


def aggregation_layer(sd, p, queries, context, latents: Sequence[torch.Tensor], masks: Sequence[torch.Tensor],
                      num_heads: int = 16) -> torch.Tensor:
    """VisionAggregationLayer.forward (vision_sampler.py:457-517) with AggregationBlock (:138-167) and CrossAttention
    (:61-135): one attention (or MLP for single-token windows) per tower, mixed by softmax(weight_mlp(...))."""
    residual = queries
    ctx = _lin(sd, p + "proj_context", context, bias=False)
    cat = torch.cat([queries, ctx], -1)
    if len(latents) > 1:
        wl = _lin(sd, p + "weight_mlp.linear_2", F.gelu(_lin(sd, p + "weight_mlp.linear_1", cat, bias=False)), bias=False)
        weight = wl.softmax(-1).unsqueeze(-1)                                    # [R, q_len, T, 1]
    else:
        weight = 1
    q = _lin(sd, p + "proj_in", cat, bias=False)
    R, q_len, hidden = q.shape
    dh = hidden // num_heads
    parts = []
    for t, v in enumerate(latents):
        a = p + f"aggregate_{t}.attention_layer."
        if v.shape[1] > 1:
            v = v + _t(sd[p + f"pos_embed_{t}"])[None]
            qs = _lin(sd, a + "q_proj.1", _ln(sd, a + "q_proj.0", q), bias=False)
            ks = _lin(sd, a + "k_proj.1", _ln(sd, a + "k_proj.0", v), bias=False)
            vs = _lin(sd, a + "v_proj.1", _ln(sd, a + "v_proj.0", v), bias=False)
            n_kv = ks.shape[1]
            qh = qs.view(R, q_len, num_heads, dh).transpose(1, 2)
            kh = ks.view(R, n_kv, num_heads, dh).transpose(1, 2)
            vh = vs.view(R, n_kv, num_heads, dh).transpose(1, 2)
            m = masks[t].view(R, 1, 1, -1).expand(-1, -1, q_len, -1)
            att = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=m).transpose(1, 2).reshape(R, q_len, hidden)
            parts.append(_lin(sd, a + "o_proj", att, bias=False))
        else:
            parts.append(_lin(sd, a + "linear_2", F.gelu(_lin(sd, a + "linear_1", v, bias=False)), bias=False))
    q = q + (torch.stack(parts, 2) * weight).sum(2)
    q = _ln(sd, p + "norm", q)
    q = _lin(sd, p + "proj_out.linear_2", F.gelu(_lin(sd, p + "proj_out.linear_1", q, bias=False)), bias=False)
    return q + residual


def token_sampler(sd, prefix, queries, context, latents, masks, num_layers: int, num_heads: int = 16,
                  layer_type: str = "joint"):
    """VisionTokenSampler.forward (vision_sampler.py:560-566): "joint" or "sep" layers (:531-559)."""
    layer = cross_attention_layer if layer_type == "joint" else aggregation_layer
    for i in range(num_layers):
        queries = layer(sd, f"{prefix}layers.{i}.", queries, context, latents, masks, num_heads)
    return queries


def sva_frames_groups(sd, tower_feats: Sequence[torch.Tensor], image_sizes: Sequence[Tuple[int, int]],
                      query_sides: Sequence[int], final_side: int, num_layers: int, num_heads: int = 16) -> torch.Tensor:
    """cambrian_arch.py:1017-1148 for several query groups (`query_num_list`): group g uses `vision_query[g]` and
    `vision_sampler_{g}`; a group whose grid differs from the final one is resized with
    F.interpolate(..., mode="bilinear", align_corners=False) in fp32 (:1107-1131, input_high_res); the groups are
    concatenated on the feature axis (:1148).  -> [bs, final_side^2, groups * hidden]"""
    outs = []
    for g, q in enumerate(query_sides):
        o = sva_frames(sd, tower_feats, image_sizes, q, num_layers, num_heads, group=g)
        bs = o.shape[0]
        if q != final_side:
            x = o.permute(0, 2, 1).contiguous().view(bs, -1, q, q)
            x = F.interpolate(x.float(), size=(final_side, final_side), mode="bilinear", align_corners=False)
            o = x.permute(0, 2, 3, 1).contiguous().flatten(1, 2)
        outs.append(o)
    return torch.cat(outs, -1)


def sva_frames(sd, tower_feats: Sequence[torch.Tensor], image_sizes: Sequence[Tuple[int, int]], query_side: int,
               num_layers: int, num_heads: int = 16, group: int = 0, layer_type: str = "joint") -> torch.Tensor:
    """cambrian_arch.py:1002-1053 for one query group (group 0) at final resolution: tower features
    [bs, grid_t^2, C_t] -> query features [bs, query_side^2, hidden]."""
    feats = [mm_projector_aux(sd, f"mm_projector_aux_{t}", f) for t, f in enumerate(tower_feats)]
    bs = feats[0].shape[0]
    hidden = feats[0].shape[-1]
    nq = query_side * query_side
    context = feats[0].mean(1).view(bs, 1, 1, -1).expand(-1, nq, 1, -1).flatten(0, 1)            # :1009-1011, 1024-1026
    queries = _t(sd["vision_query"])[group].view(1, 1, 1, -1).expand(bs, nq, -1, -1).flatten(0, 1)  # :1018-1023
    latents = [rearrange_windows(f, query_side) for f in feats]
    masks = [torch.cat([window_masks(image_sizes[b], int(f.shape[1] ** 0.5), query_side) for b in range(bs)], 0)
             for f in feats]
    out = token_sampler(sd, f"vision_sampler_{group}.", queries, context, latents, masks, num_layers, num_heads,
                        layer_type)
    return out.view(bs, nq, hidden)
