"""Spatial Vision Aggregator (SVA) connector on the B200 kernels — SURVEY §8f-3.

Reference: `mm_projector_aux_{t}` (tdc/cambrian_arch.py:83-93, applied :1002-1013), the window rearrangement and
attention masks (:601-690, :487-509) and `vision_sampler_{g}` = `VisionTokenSampler` of joint
`VisionCrossAttentionLayer`s with `MultiKVCrossAttention` (tdc/vision_sampler.py:170-291, 305-401, 519-566), as driven
from cambrian_arch.py:1002-1053: every one of the Q x Q learnable queries of a frame cross-attends to the r x r
window of tokens under it in each vision tower's feature grid.

`SVAConnector` keeps the reference's attribute and parameter names (`mm_projector_aux_0`, `vision_query`,
`vision_sampler_0.layers.{i}.cross_attn.k_proj_0.1.weight`, ...), so the `model.*` entries of a reference checkpoint
load unchanged.  The forward is a sequence of libtdc_b200 calls: tcgen05 GEMMs (`tdc_linear`, GELU fused), the
fused (residual / position-embedding +) LayerNorm (`tdc_layernorm`), the short-query attention kernel with a
per-row key mask (`tdc_attention`; one query against r0^2 + r1^2 keys, two-segment KV = the two towers) and
`tdc_residual_add`, the window regrouping (`tdc_window_rearrange`), the bilinear resize of coarse query groups
(`tdc_resize_tokens_bilinear`) and the tower mix of "sep" layers (`tdc_combine_parts`).  torch does data movement only
(concatenation, broadcast).  CUDA only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch
from torch import nn

from . import _lib
from .engine import _ptr, _stream, avg_pool_tokens, linear

LN_EPS = 1e-5
_DT = {torch.bfloat16: _lib.TDC_BF16, torch.float16: _lib.TDC_F16, torch.float32: _lib.TDC_F32}


def window_mask_bits(image_sizes: Sequence[Tuple[int, int]], grids: Sequence[int], query_side: int) -> np.ndarray:
    """uint32 [bs * Q^2]: bit j set = KV token j of the row (tower 0's window tokens first, then tower 1's) is real
    image content.  Integer restatement of unmask_attention_mask + the per-window regrouping
    (cambrian_arch.py:487-509, 619-669); all-padding windows are fully enabled, as in the reference (:667-669)."""
    out = []
    for (w, h) in image_sizes:
        per_tower = []
        for grid in grids:
            m = np.ones((grid, grid), dtype=bool)
            if (w / h) > 1.0:
                pad = (grid - int(h * (grid / w))) // 2
                if pad > 0:
                    m[:pad, :] = False
                    m[-pad:, :] = False
            else:
                pad = (grid - int(w * (grid / h))) // 2
                if pad > 0:
                    m[:, :pad] = False
                    m[:, -pad:] = False
            r = grid // query_side
            win = m.reshape(query_side, r, query_side, r).transpose(0, 2, 1, 3).reshape(query_side * query_side, r * r)
            win[win.sum(-1) == 0] = True
            per_tower.append(win)
        bits = np.zeros(query_side * query_side, dtype=np.uint64)
        shift = 0
        for win in per_tower:
            for j in range(win.shape[1]):
                bits |= win[:, j].astype(np.uint64) << np.uint64(shift + j)
            shift += win.shape[1]
        out.append(bits.astype(np.uint32))
    return np.concatenate(out) if out else np.zeros(0, np.uint32)


def _layernorm(x: torch.Tensor, ln: nn.LayerNorm, *, resid: Optional[torch.Tensor] = None, resid_period: int = 0,
               want_f32: bool = False, want_bf16: bool = True):
    """LayerNorm(x [+ resid]) through tdc_layernorm: x fp32 [rows, width] -> bf16 and/or fp32 copies."""
    lib = _lib.load_library()
    rows, width = x.shape
    y16 = torch.empty((rows, width), dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    y32 = torch.empty((rows, width), dtype=torch.float32, device=x.device) if want_f32 else None
    with torch.cuda.device(x.device):
        rc = lib.tdc_layernorm(_ptr(x), _ptr(resid), resid_period, _ptr(ln.weight), _ptr(ln.bias), float(ln.eps),
                               _ptr(y32), _ptr(y16), rows, width, _stream(x.device))
    _lib.check(rc, None, "tdc_layernorm")
    if want_f32 and want_bf16:
        return y16, y32
    return y32 if want_f32 else y16


class _Params(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: the computation runs in libtdc_b200.so")


def _ln_linear(d_in: int, d_out: int) -> nn.Sequential:
    return nn.Sequential(nn.LayerNorm(d_in), nn.Linear(d_in, d_out, bias=False))


def _cross_attention_layer(hidden: int, window_sides: Sequence[int]) -> nn.Module:
    layer = _Params()
    layer.proj_context = nn.Linear(hidden, hidden, bias=False)
    layer.proj_in = nn.Linear(2 * hidden, hidden, bias=False)
    layer.proj_out = _Params()
    layer.proj_out.linear_1 = nn.Linear(hidden, hidden, bias=False)
    layer.proj_out.linear_2 = nn.Linear(hidden, hidden, bias=False)
    layer.norm = nn.LayerNorm(hidden)
    layer.cross_attn = _Params()
    layer.cross_attn.q_proj = _ln_linear(hidden, hidden)
    for t, side in enumerate(window_sides):
        setattr(layer.cross_attn, f"k_proj_{t}", _ln_linear(hidden, hidden))
        setattr(layer.cross_attn, f"v_proj_{t}", _ln_linear(hidden, hidden))
        if side > 1:
            setattr(layer, f"pos_embed_{t}", nn.Parameter(torch.randn(side * side, hidden)))
    layer.cross_attn.o_proj = nn.Linear(hidden, hidden, bias=False)
    return layer


def _aggregation_layer(hidden: int, window_sides: Sequence[int]) -> nn.Module:
    """Parameter container of `VisionAggregationLayer` (vision_sampler.py:404-455): one AggregationBlock per tower —
    a CrossAttention (own q/k/v/o projections) where the window has several tokens, an MLP where it has one — and a
    `weight_mlp` that mixes the towers."""
    layer = _Params()
    layer.proj_context = nn.Linear(hidden, hidden, bias=False)
    layer.proj_in = nn.Linear(2 * hidden, hidden, bias=False)
    layer.proj_out = _Params()
    layer.proj_out.linear_1 = nn.Linear(hidden, hidden, bias=False)
    layer.proj_out.linear_2 = nn.Linear(hidden, hidden, bias=False)
    layer.norm = nn.LayerNorm(hidden)
    if len(window_sides) > 1:
        layer.weight_mlp = _Params()
        layer.weight_mlp.linear_1 = nn.Linear(2 * hidden, hidden, bias=False)
        layer.weight_mlp.linear_2 = nn.Linear(hidden, len(window_sides), bias=False)
    for t, side in enumerate(window_sides):
        agg = _Params()
        if side > 1:
            setattr(layer, f"pos_embed_{t}", nn.Parameter(torch.randn(side * side, hidden)))
            agg.attention_layer = _Params()
            agg.attention_layer.q_proj = _ln_linear(hidden, hidden)
            agg.attention_layer.k_proj = _ln_linear(hidden, hidden)
            agg.attention_layer.v_proj = _ln_linear(hidden, hidden)
            agg.attention_layer.o_proj = nn.Linear(hidden, hidden, bias=False)
        else:
            agg.attention_layer = _Params()
            agg.attention_layer.linear_1 = nn.Linear(hidden, hidden, bias=False)
            agg.attention_layer.linear_2 = nn.Linear(hidden, hidden, bias=False)
        setattr(layer, f"aggregate_{t}", agg)
    return layer


class SVAConnector(nn.Module):
    """`mm_projector_aux_{t}` + `vision_query` + `vision_sampler_{g}` of the reference model.

    Default: one query group whose grid equals the final token grid (query_num_list == [image_token_len], the
    shipped configuration).  `query_sides` = the grid side of every query group (sqrt of `query_num_list`,
    cambrian_arch.py:1017-1053): group g gets its own `vision_query[g]` and `vision_sampler_{g}` over windows of
    side grid_t / query_sides[g]; groups whose grid differs from the final `query_side` are bilinearly resized to it
    (:1107-1131) and all groups are concatenated on the feature axis (:1148), as the reference does.
    `layer_type`: "joint" = `VisionCrossAttentionLayer` (one softmax over the windows of all towers — what the
    reference model constructs, cambrian_arch.py:101, 128, 286) or "sep" = `VisionAggregationLayer`
    (vision_sampler.py:404-517: one attention per tower, mixed by a learned softmax weight; selectable in
    `VisionTokenSampler`, never selected by the shipped model code).
    Not provided: the samplers inside the LLM layers (`vision_sampler_layers`, not connector_only), which belong to
    the language model."""

    def __init__(self, tower_dims: Sequence[int], window_sides: Sequence[int], hidden: int = 1024, query_side: int = 12,
                 num_layers: int = 3, query_sides: Optional[Sequence[int]] = None, layer_type: str = "joint"):
        super().__init__()
        if layer_type not in ("joint", "sep"):
            raise ValueError('layer_type must be "joint" or "sep" (vision_sampler.py:531)')
        self.layer_type = layer_type
        if len(tower_dims) != 2 or len(window_sides) != 2:
            raise NotImplementedError("two vision towers (SigLIP + DINOv2) as in the reference; the attention kernel "
                                      "addresses the two towers as its two KV segments")
        if hidden % 64 != 0:
            raise ValueError("libtdc_b200 attention has head size 64: hidden must be a multiple of 64 "
                             "(reference: 1024 = 16 heads x 64)")
        self.hidden, self.query_side, self.num_layers = hidden, query_side, num_layers
        self.window_sides = tuple(int(s) for s in window_sides)
        self.tower_grids = tuple(query_side * s for s in self.window_sides)
        self.query_sides = tuple(int(q) for q in (query_sides if query_sides is not None else (query_side,)))
        self.group_window_sides = []
        for q in self.query_sides:
            if any(g % q for g in self.tower_grids):
                raise ValueError(f"query grid side {q} does not divide the tower grids {self.tower_grids}")
            sides = tuple(g // q for g in self.tower_grids)
            if sum(s * s for s in sides) > 32:
                raise ValueError("at most 32 KV tokens per query (per-row key mask is 32 bits)")
            self.group_window_sides.append(sides)
        for t, c in enumerate(tower_dims):
            setattr(self, f"mm_projector_aux_{t}", nn.Sequential(nn.Linear(c, hidden), nn.GELU(),
                                                                 nn.Linear(hidden, hidden), nn.LayerNorm(hidden)))
        self.vision_query = nn.Parameter(torch.randn(len(self.query_sides), hidden))
        for g, sides in enumerate(self.group_window_sides):
            sampler = _Params()
            make = _cross_attention_layer if layer_type == "joint" else _aggregation_layer
            sampler.layers = nn.ModuleList([make(hidden, sides) for _ in range(num_layers)])
            setattr(self, f"vision_sampler_{g}", sampler)
        self._bf16 = {}
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._bf16.clear())
        self.eval()

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._bf16 = {}
        return out

    def _w(self, lin: nn.Linear) -> torch.Tensor:
        """bf16 copy of a weight, made once (cleared by load_state_dict / .to())."""
        key = id(lin)
        if key not in self._bf16:
            self._bf16[key] = lin.weight.detach().to(torch.bfloat16).contiguous()
        return self._bf16[key]

    def _kv_folded_of(self, key, k_seq, v_seq):
        """`_kv_folded` for an explicit pair of (LayerNorm, Linear) sequences."""
        if key not in self._bf16:
            ws, bs = [], []
            for ln, lin in (k_seq, v_seq):
                w = lin.weight.detach().float()
                ws.append(w * ln.weight.detach().float()[None, :])
                bs.append(w @ ln.bias.detach().float())
            dev = ws[0].device
            self._bf16[key] = (torch.cat(ws, 0).to(torch.bfloat16).contiguous(), torch.cat(bs, 0).contiguous(),
                               torch.ones(self.hidden, device=dev), torch.zeros(self.hidden, device=dev))
        return self._bf16[key]

    def _kv_folded(self, li: int, t: int, g: int = 0):
        """K and V projections of one tower as ONE GEMM over the shared normalised input: both are
        Linear(LayerNorm(x)) of the same x (vision_sampler.py:192-217), and LayerNorm(x) = xhat * gamma + beta with
        xhat = (x - mean) / std common to the two, so W (gamma * xhat + beta) = (W diag(gamma)) xhat + W beta.
        Returns ([2H, H] bf16 weight, [2H] fp32 bias, ones/zeros LayerNorm parameters); made once."""
        key = ("kv", g, li, t)
        if key not in self._bf16:
            ca = getattr(self, f"vision_sampler_{g}").layers[li].cross_attn
            ws, bs = [], []
            for name in (f"k_proj_{t}", f"v_proj_{t}"):
                ln, lin = getattr(ca, name)
                w = lin.weight.detach().float()
                ws.append(w * ln.weight.detach().float()[None, :])
                bs.append(w @ ln.bias.detach().float())
            dev = ws[0].device
            self._bf16[key] = (torch.cat(ws, 0).to(torch.bfloat16).contiguous(), torch.cat(bs, 0).contiguous(),
                               torch.ones(self.hidden, device=dev), torch.zeros(self.hidden, device=dev))
        return self._bf16[key]

    @torch.no_grad()
    def forward(self, tower_feats: Sequence[torch.Tensor], image_sizes: Sequence[Tuple[int, int]]) -> torch.Tensor:
        """tower_feats[t]: [bs, grid_t^2, C_t] (CUDA); image_sizes[b] = (width, height) of frame b before padding.
        Returns the query features [bs, Q^2, groups * hidden] (bf16) that feed `mm_projector`
        (cambrian_arch.py:1146-1150)."""
        if self.training:
            raise RuntimeError("SVAConnector is inference-only (eval mode)")
        if not tower_feats[0].is_cuda:
            raise RuntimeError("SVAConnector needs CUDA tensors: there is no CPU fallback")
        outs = []
        for g, q in enumerate(self.query_sides):
            o = self._forward_group(g, tower_feats, image_sizes)                   # [bs, q^2, H] bf16
            if q != self.query_side:                                               # :1107-1131 (input_high_res)
                lib = _lib.load_library()
                r = torch.empty((o.shape[0], self.query_side ** 2, self.hidden), dtype=torch.bfloat16, device=o.device)
                with torch.cuda.device(o.device):
                    rc = lib.tdc_resize_tokens_bilinear(_ptr(o), _lib.TDC_BF16, o.shape[0], q, self.query_side,
                                                        self.hidden, _ptr(r), _lib.TDC_BF16, _stream(o.device))
                _lib.check(rc, None, "tdc_resize_tokens_bilinear")
                o = r
            outs.append(o)
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=-1)               # :1148

    def _aggregation_forward(self, group, li, layer, q32, q16, context, latents, counts, mask, bs, Q):
        """One VisionAggregationLayer (vision_sampler.py:457-517) on the library's kernels."""
        lib = _lib.load_library()
        dev, H = q32.device, self.hidden
        R = q32.shape[0]
        T = len(latents)
        ctx = linear(context, self._w(layer.proj_context))                       # [bs, H] bf16
        cat = torch.cat([q16, ctx.repeat_interleave(Q * Q, dim=0)], dim=-1)      # [R, 2H] bf16
        q1 = linear(cat, self._w(layer.proj_in), out_dtype=torch.float32)        # [R, H] fp32
        parts = torch.empty((T, R, H), dtype=torch.float32, device=dev)
        shift = 0
        for t, (lat, n) in enumerate(zip(latents, counts)):
            agg = getattr(layer, f"aggregate_{t}").attention_layer
            if n > 1:
                pos = getattr(layer, f"pos_embed_{t}")
                qs = linear(_layernorm(q1, agg.q_proj[0]), self._w(agg.q_proj[1]))
                w_kv, b_kv, ones, zeros = self._kv_folded_of(("sep", group, li, t), agg.k_proj, agg.v_proj)
                xhat = torch.empty((R * n, H), dtype=torch.bfloat16, device=dev)
                with torch.cuda.device(dev):
                    rc = lib.tdc_layernorm(_ptr(lat), _ptr(pos.detach().float().contiguous()), n, _ptr(ones), _ptr(zeros),
                                           float(agg.k_proj[0].eps), None, _ptr(xhat), R * n, H, _stream(dev))
                _lib.check(rc, None, "tdc_layernorm")
                kv = linear(xhat, w_kv, b_kv)                                    # [R*n, 2H] bf16: K | V
                att = torch.empty((R, H), dtype=torch.bfloat16, device=dev)
                tmask = torch.from_numpy(((mask >> np.uint32(shift)) & np.uint32((1 << n) - 1)).astype(np.uint32)
                                         .view(np.int32)).to(dev)                # this tower's bits of the key mask
                with torch.cuda.device(dev):
                    rc = lib.tdc_attention(_ptr(qs), _ptr(kv), C.c_void_p(kv.data_ptr() + H * kv.element_size()), _ptr(att),
                                           H, 2 * H, 2 * H, H, R, H // 64, 1, 0, 0, 0, n, 0, 0, 0, None, _ptr(tmask),
                                           _stream(dev))
                _lib.check(rc, None, "tdc_attention")
                linear(att, self._w(agg.o_proj), out=parts[t])
            else:                                                                # one token per window: an MLP on it
                x16 = lat.to(torch.bfloat16)
                linear(linear(x16, self._w(agg.linear_1), gelu=True), self._w(agg.linear_2), out=parts[t])
            shift += n
        if T > 1:
            # weight_mlp(cat).softmax(-1): the N = T output is padded to 8 columns for the GEMM (zero rows)
            key = ("wmlp2", group, li)
            if key not in self._bf16:
                w2 = layer.weight_mlp.linear_2.weight.detach()
                pad = torch.zeros((8, H), dtype=torch.bfloat16, device=dev)
                pad[:T] = w2.to(torch.bfloat16)
                self._bf16[key] = pad
            logits = linear(linear(cat, self._w(layer.weight_mlp.linear_1), gelu=True), self._bf16[key],
                            out_dtype=torch.float32)                            # [R, 8]
        else:
            logits = torch.zeros((R, 8), dtype=torch.float32, device=dev)
        mixed = torch.empty_like(q1)
        with torch.cuda.device(dev):
            rc = lib.tdc_combine_parts(_ptr(q1), _ptr(parts), _ptr(logits), 8, T, R, H, _ptr(mixed), _stream(dev))
        _lib.check(rc, None, "tdc_combine_parts")
        q2 = _layernorm(mixed, layer.norm)
        m = linear(linear(q2, self._w(layer.proj_out.linear_1), gelu=True), self._w(layer.proj_out.linear_2),
                   out_dtype=torch.float32)
        new32, new16 = torch.empty_like(q32), torch.empty_like(q16)
        with torch.cuda.device(dev):
            rc = lib.tdc_residual_add(_ptr(m), _ptr(q32), _ptr(new32), _ptr(new16), m.numel(), _stream(dev))
        _lib.check(rc, None, "tdc_residual_add")
        return new32, new16

    def _forward_group(self, group: int, tower_feats, image_sizes) -> torch.Tensor:
        lib = _lib.load_library()
        dev = tower_feats[0].device
        bs, Q, H = tower_feats[0].shape[0], self.query_sides[group], self.hidden
        window_sides = self.group_window_sides[group]
        sampler = getattr(self, f"vision_sampler_{group}")
        R = bs * Q * Q
        grids = [int(round(f.shape[1] ** 0.5)) for f in tower_feats]
        for g, s in zip(grids, window_sides):
            if g != Q * s:
                raise ValueError(f"tower grid {g} != query_side {Q} x window side {s}")
        # --- mm_projector_aux_t: Linear . GELU . Linear . LayerNorm  (cambrian_arch.py:1002-1013)
        latents, feat0 = [], None
        for t, x in enumerate(tower_feats):
            seq = getattr(self, f"mm_projector_aux_{t}")
            r = window_sides[t]
            # tokens under every query, window-major (cambrian_arch.py:624-645).  The projector acts on each token
            # separately and the context is a mean over tokens, so the rearrangement is applied to the (narrow,
            # bf16) tower features instead of the projector's fp32 output: pure data movement, fused with the cast.
            xw = torch.empty((bs, Q, Q, r, r, x.shape[-1]), dtype=torch.bfloat16, device=dev)
            xs = x.contiguous()
            if xs.dtype not in _DT:
                xs = xs.float()
            with torch.cuda.device(dev):
                rc = lib.tdc_window_rearrange(_ptr(xs), _DT[xs.dtype], bs, Q, r, xs.shape[-1], _ptr(xw), _stream(dev))
            _lib.check(rc, None, "tdc_window_rearrange")
            h = linear(xw.view(R * r * r, -1), self._w(seq[0]), seq[0].bias, gelu=True)
            y = linear(h, self._w(seq[2]), seq[2].bias, out_dtype=torch.float32)
            f32 = _layernorm(y, seq[3], want_f32=True, want_bf16=False)              # [R * r * r, H] fp32
            if t == 0:
                feat0 = f32.view(bs, grids[t] * grids[t], H)
            latents.append(f32)
        context = avg_pool_tokens(feat0, 1).reshape(bs, H)                       # global context = token mean (:1009)
        q32 = self.vision_query.detach()[group].float().view(1, H).expand(R, H).contiguous()
        q16 = q32.to(torch.bfloat16)
        mask_np = window_mask_bits(image_sizes, grids, Q)                        # host integers (uint32 per query)
        mask = torch.from_numpy(mask_np.view(np.int32)).to(dev)
        n0, n1 = window_sides[0] ** 2, window_sides[1] ** 2
        rows0 = R * n0                                                           # K/V rows of tower 0 precede tower 1's
        kv = torch.empty((R * (n0 + n1), 2 * H), dtype=torch.bfloat16, device=dev)   # row = [K | V] of one token
        xhat = torch.empty((R * max(n0, n1), H), dtype=torch.bfloat16, device=dev)
        att = torch.empty((R, H), dtype=torch.bfloat16, device=dev)
        for li, layer in enumerate(sampler.layers):
            if self.layer_type == "sep":
                q32, q16 = self._aggregation_forward(group, li, layer, q32, q16, context, latents, (n0, n1), mask_np, bs, Q)
                continue
            ca = layer.cross_attn
            # proj_context + cat + proj_in (vision_sampler.py:346-360); the context is one vector per frame
            ctx = linear(context, self._w(layer.proj_context))                   # [bs, H] bf16
            cat = torch.cat([q16, ctx.repeat_interleave(Q * Q, dim=0)], dim=-1)  # [R, 2H] bf16
            q1 = linear(cat, self._w(layer.proj_in), out_dtype=torch.float32)    # fp32 residual of the inner block
            qs = linear(_layernorm(q1, ca.q_proj[0]), self._w(ca.q_proj[1]))
            for t, (lat, n, off) in enumerate(zip(latents, (n0, n1), (0, rows0))):
                pos = getattr(layer, f"pos_embed_{t}", None)
                w_kv, b_kv, ones, zeros = self._kv_folded(li, t, group)
                # one normalisation pass (statistics only) shared by K and V, then one N = 2H GEMM
                with torch.cuda.device(dev):
                    rc = lib.tdc_layernorm(_ptr(lat), None if pos is None else _ptr(pos.detach().float().contiguous()),
                                           0 if pos is None else n, _ptr(ones), _ptr(zeros),
                                           float(getattr(ca, f"k_proj_{t}")[0].eps), None,
                                           _ptr(xhat), R * n, H, _stream(dev))
                _lib.check(rc, None, "tdc_layernorm")
                linear(xhat[:R * n], w_kv, b_kv, out=kv[off:off + R * n])
            with torch.cuda.device(dev):
                rc = lib.tdc_attention(_ptr(qs), _ptr(kv), C.c_void_p(kv.data_ptr() + H * kv.element_size()), _ptr(att), H, 2 * H,
                                       2 * H, H, R, H // 64, 1, 0, 0, 0, n0, n1, 0, rows0, None, _ptr(mask),
                                       _stream(dev))
            _lib.check(rc, None, "tdc_attention")
            o = linear(att, self._w(ca.o_proj), out_dtype=torch.float32)
            q2 = _layernorm(o, layer.norm, resid=q1)                             # norm(queries + attention_output)
            m = linear(linear(q2, self._w(layer.proj_out.linear_1), gelu=True), self._w(layer.proj_out.linear_2),
                       out_dtype=torch.float32)
            new32 = torch.empty_like(q32)
            new16 = torch.empty_like(q16)
            with torch.cuda.device(dev):
                rc = lib.tdc_residual_add(_ptr(m), _ptr(q32), _ptr(new32), _ptr(new16), m.numel(), _stream(dev))
            _lib.check(rc, None, "tdc_residual_add")
            q32, q16 = new32, new16
        return q16.view(bs, Q * Q, H)
