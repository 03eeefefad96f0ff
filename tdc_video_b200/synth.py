"""Deterministic synthetic weights / inputs for the TDC path (benchmarks, smoke test, golden fixtures).

Data generation only — no model arithmetic lives here; `oracle/synth.py` re-exports these names for the tests.

Everything is drawn from numpy's legacy `RandomState` (bit-stable across numpy versions and
machines), so a golden fixture only has to store a geometry, a seed and the expected
outputs; tests regenerate identical weights on the GPU box where the reference is absent.

Weight statistics follow the reference initialisation (tdc/Qformer.py:664-674: N(0, 0.02)
linears/embeddings, LayerNorm (1, 0)) but with non-zero biases and perturbed LayerNorm
affine parameters, so bias / affine bugs cannot hide behind zeros.  `stress=s` scales the
attention query/key weights by s so that softmax is peaked (random-init attention is nearly
uniform and hides softmax bugs — SURVEY.md §8d).  s = 8 suits the small test geometries
(score std ~3); at the reference geometry (hidden 768, d_enc 3584) s = 2 gives the same
score spread — s = 8 there makes the scores' std ~40, i.e. a hard arg-max whose near-ties
no 16-bit implementation (fp16 reference inference included) can reproduce.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import Dict, Optional

import numpy as np


@dataclass(frozen=True)
class QFormerGeometry:
    hidden: int = 768
    heads: int = 12
    intermediate: int = 3072
    layers: int = 12
    cross_freq: int = 2
    d_enc: int = 3584
    d_out: int = 3584
    vocab: int = 30522
    max_pos: int = 512
    ln_eps: float = 1e-12

    def to_dict(self):
        return asdict(self)

    @property
    def cross_layers(self):
        return [l for l in range(self.layers) if l % self.cross_freq == 0]


def _lin(rs, out_f, in_f, scale=0.02):
    return (rs.standard_normal((out_f, in_f)) * scale).astype(np.float32), \
           (rs.standard_normal((out_f,)) * 0.02).astype(np.float32)


def _ln(rs, n):
    return (1.0 + 0.1 * rs.standard_normal((n,))).astype(np.float32), \
           (0.1 * rs.standard_normal((n,))).astype(np.float32)


def make_state_dict(geom: QFormerGeometry, seed: int, stress: float = 0.0, with_text: bool = True,
                    with_vision_proj: bool = True) -> Dict[str, np.ndarray]:
    """State dict with the reference's key names relative to `Qformer.bert.`,
    plus the sibling `vision_proj.{weight,bias}` (tdc/cambrian_arch.py:483)."""
    rs = np.random.RandomState(seed)
    H, I, E = geom.hidden, geom.intermediate, geom.d_enc
    qk = float(stress) if stress else 1.0  # True -> 1.0 is never passed; use an explicit scale
    sd: Dict[str, np.ndarray] = {}

    def put_lin(prefix, out_f, in_f, scale=0.02):
        w, b = _lin(rs, out_f, in_f, scale)
        sd[prefix + ".weight"], sd[prefix + ".bias"] = w, b

    def put_ln(prefix, n):
        g, b = _ln(rs, n)
        sd[prefix + ".weight"], sd[prefix + ".bias"] = g, b

    if with_text and geom.vocab > 0:
        sd["embeddings.word_embeddings.weight"] = (rs.standard_normal((geom.vocab, H)) * 0.02).astype(np.float32)
        sd["embeddings.position_embeddings.weight"] = (rs.standard_normal((geom.max_pos, H)) * 0.02).astype(np.float32)
    put_ln("embeddings.LayerNorm", H)
    for l in range(geom.layers):
        p = f"encoder.layer.{l}."
        put_lin(p + "attention.self.query", H, H, 0.02 * qk)
        put_lin(p + "attention.self.key", H, H, 0.02 * qk)
        put_lin(p + "attention.self.value", H, H)
        put_lin(p + "attention.output.dense", H, H)
        put_ln(p + "attention.output.LayerNorm", H)
        if l % geom.cross_freq == 0:
            put_lin(p + "crossattention.self.query", H, H, 0.02 * qk)
            put_lin(p + "crossattention.self.key", H, E, 0.02 * qk)
            put_lin(p + "crossattention.self.value", H, E)
            put_lin(p + "crossattention.output.dense", H, H)
            put_ln(p + "crossattention.output.LayerNorm", H)
        put_lin(p + "intermediate_query.dense", I, H)
        put_lin(p + "output_query.dense", H, I)
        put_ln(p + "output_query.LayerNorm", H)
        if with_text and geom.vocab > 0:
            put_lin(p + "intermediate.dense", I, H)
            put_lin(p + "output.dense", H, I)
            put_ln(p + "output.LayerNorm", H)
    if with_vision_proj and geom.d_out > 0:
        put_lin("vision_proj", geom.d_out, H)
    return sd


def make_inputs(geom: QFormerGeometry, seed: int, rows: int, kv_tokens: int, num_query: int, num_text: int = 0,
                audio_tokens: int = 0) -> Dict[str, Optional[np.ndarray]]:
    """Synthetic call inputs: query_embeds [rows,K,H], enc [rows,L,d_enc], input_ids [rows,T].

    Visual tokens ~ N(0,1); the last `audio_tokens` KV tokens of every row are "audio":
    N(0,1)*0.5 with the trailing 10 % of rows zero (the reference zero-pads missing audio,
    tdc/cambrian_arch.py:1593-1595)."""
    rs = np.random.RandomState(seed + 7919)
    q = rs.standard_normal((rows, num_query, geom.hidden)).astype(np.float32)
    enc = rs.standard_normal((rows, kv_tokens, geom.d_enc)).astype(np.float32)
    if audio_tokens > 0:
        enc[:, kv_tokens - audio_tokens:, :] *= 0.5
        nz = max(1, rows // 10)
        enc[rows - nz:, kv_tokens - audio_tokens:, :] = 0.0
    ids = None
    if num_text > 0:
        lo = min(1000, max(geom.vocab - 2, 0))
        ids = rs.randint(lo if lo < geom.vocab - 1 else 0, max(geom.vocab - 1, 1), size=(rows, num_text)).astype(np.int64)
    return {"query_embeds": q, "enc": enc, "input_ids": ids}


def make_frontend_state_dict(d_llm: int, d_frame_in: int, d_audio: int, hidden: int, seed: int,
                             num_query: int = 16) -> Dict[str, np.ndarray]:
    """The sibling tensors of the upstream ("frames") entry under the reference's names: `mm_projector.{0,2}.*`
    (cambrian_arch.py:65-69), `image_newline` (:148), `query_proj.*` (:484), `query_tokens` (:420-423) and, with
    d_audio > 0, `audio_proj.*` (:180-181).  Scales are chosen so that N(0,1) tower features give projected
    tokens of about unit variance and N(0, 0.25) audio features give audio tokens of variance 0.25 — the
    statistics `make_inputs` uses for already-projected tokens."""
    rs = np.random.RandomState(seed)
    f = lambda shape, scale: (rs.standard_normal(shape) * scale).astype(np.float32)
    sd = {
        "mm_projector.0.weight": f((d_llm, d_frame_in), 1.0 / np.sqrt(d_frame_in)), "mm_projector.0.bias": f((d_llm,), 0.02),
        "mm_projector.2.weight": f((d_llm, d_llm), 1.0 / np.sqrt(0.425 * d_llm)), "mm_projector.2.bias": f((d_llm,), 0.02),
        "image_newline": f((d_llm,), 0.5),
        "query_proj.weight": f((hidden, d_llm), 0.02), "query_proj.bias": f((hidden,), 0.02),
        "query_tokens": f((1, num_query, hidden), 0.02),
    }
    if d_audio > 0:
        sd["audio_proj.weight"] = f((d_llm, d_audio), 1.0 / np.sqrt(d_audio))
        sd["audio_proj.bias"] = f((d_llm,), 0.02)
    return sd


# ---- Spatial Vision Aggregator -----------------------------------------------------------------
def make_sva_state_dict(hidden: int, tower_dims, window_sides, num_layers: int, seed: int, stress: float = 1.0
                        ) -> Dict[str, np.ndarray]:
    """Weights of `mm_projector_aux_{t}`, `vision_query` and `vision_sampler_0` with the reference's parameter
    names (tdc/cambrian_arch.py:83-101,139-142; tdc/vision_sampler.py:305-341,170-217).  window_sides[t] =
    tower grid side / query grid side (pos_embed_t has window_sides[t]^2 rows)."""
    rs = np.random.RandomState(seed)
    sd: Dict[str, np.ndarray] = {}
    f32 = np.float32

    def lin(name, out_f, in_f, bias=False, scale=None):
        sd[name + ".weight"] = (rs.standard_normal((out_f, in_f)) * (scale or 1.0 / np.sqrt(in_f))).astype(f32)
        if bias:
            sd[name + ".bias"] = (rs.standard_normal((out_f,)) * 0.05).astype(f32)

    def ln(name, n):
        sd[name + ".weight"] = (1.0 + 0.1 * rs.standard_normal((n,))).astype(f32)
        sd[name + ".bias"] = (0.1 * rs.standard_normal((n,))).astype(f32)

    for t, c in enumerate(tower_dims):
        lin(f"mm_projector_aux_{t}.0", hidden, c, bias=True)
        lin(f"mm_projector_aux_{t}.2", hidden, hidden, bias=True)
        ln(f"mm_projector_aux_{t}.3", hidden)
    sd["vision_query"] = rs.standard_normal((1, hidden)).astype(f32)
    _sva_sampler_weights(sd, rs, "vision_sampler_0.", hidden, window_sides, num_layers, stress)
    return sd


def make_sva_sep_state_dict(hidden: int, tower_dims, window_sides, num_layers: int, seed: int,
                            stress: float = 1.0) -> Dict[str, np.ndarray]:
    """As make_sva_state_dict, with `vision_sampler_0` made of VisionAggregationLayers (layer_type "sep",
    tdc/vision_sampler.py:404-455): per tower an AggregationBlock (`aggregate_{t}.attention_layer.*`) and a
    `weight_mlp` mixing the towers."""
    rs = np.random.RandomState(seed)
    sd: Dict[str, np.ndarray] = {}
    f32 = np.float32

    def lin(name, out_f, in_f, bias=False, scale=None):
        sd[name + ".weight"] = (rs.standard_normal((out_f, in_f)) * (scale or 1.0 / np.sqrt(in_f))).astype(f32)
        if bias:
            sd[name + ".bias"] = (rs.standard_normal((out_f,)) * 0.05).astype(f32)

    def ln(name, n):
        sd[name + ".weight"] = (1.0 + 0.1 * rs.standard_normal((n,))).astype(f32)
        sd[name + ".bias"] = (0.1 * rs.standard_normal((n,))).astype(f32)

    for t, c in enumerate(tower_dims):
        lin(f"mm_projector_aux_{t}.0", hidden, c, bias=True)
        lin(f"mm_projector_aux_{t}.2", hidden, hidden, bias=True)
        ln(f"mm_projector_aux_{t}.3", hidden)
    sd["vision_query"] = rs.standard_normal((1, hidden)).astype(f32)
    for i in range(num_layers):
        p = f"vision_sampler_0.layers.{i}."
        lin(p + "proj_context", hidden, hidden)
        lin(p + "proj_in", hidden, 2 * hidden)
        lin(p + "proj_out.linear_1", hidden, hidden)
        lin(p + "proj_out.linear_2", hidden, hidden)
        ln(p + "norm", hidden)
        if len(window_sides) > 1:
            lin(p + "weight_mlp.linear_1", hidden, 2 * hidden)
            lin(p + "weight_mlp.linear_2", len(window_sides), hidden, scale=2.0 / np.sqrt(hidden))
        for t, side in enumerate(window_sides):
            a = p + f"aggregate_{t}.attention_layer."
            if side > 1:
                sd[p + f"pos_embed_{t}"] = rs.standard_normal((side * side, hidden)).astype(f32)
                for nm in ("q_proj", "k_proj", "v_proj"):
                    ln(a + nm + ".0", hidden)
                    lin(a + nm + ".1", hidden, hidden, scale=(stress if nm != "v_proj" else 1.0) / np.sqrt(hidden))
                lin(a + "o_proj", hidden, hidden)
            else:
                lin(a + "linear_1", hidden, hidden)
                lin(a + "linear_2", hidden, hidden)
    return sd


def add_sva_group(sd: Dict[str, np.ndarray], group: int, hidden: int, window_sides, num_layers: int, seed: int,
                  stress: float = 1.0) -> None:
    """One more query group (`vision_query[group]`, `vision_sampler_{group}`; cambrian_arch.py:92-110, 139-142)."""
    rs = np.random.RandomState(seed)
    assert sd["vision_query"].shape[0] == group
    sd["vision_query"] = np.concatenate([sd["vision_query"], rs.standard_normal((1, hidden)).astype(np.float32)], 0)
    _sva_sampler_weights(sd, rs, f"vision_sampler_{group}.", hidden, window_sides, num_layers, stress)


def _sva_sampler_weights(sd, rs, prefix, hidden, window_sides, num_layers, stress):
    f32 = np.float32

    def lin(name, out_f, in_f, bias=False, scale=None):
        sd[name + ".weight"] = (rs.standard_normal((out_f, in_f)) * (scale or 1.0 / np.sqrt(in_f))).astype(f32)

    def ln(name, n):
        sd[name + ".weight"] = (1.0 + 0.1 * rs.standard_normal((n,))).astype(f32)
        sd[name + ".bias"] = (0.1 * rs.standard_normal((n,))).astype(f32)

    for i in range(num_layers):
        p = f"{prefix}layers.{i}."
        lin(p + "proj_context", hidden, hidden)
        lin(p + "proj_in", hidden, 2 * hidden)
        lin(p + "proj_out.linear_1", hidden, hidden)
        lin(p + "proj_out.linear_2", hidden, hidden)
        ln(p + "norm", hidden)
        ln(p + "cross_attn.q_proj.0", hidden)
        lin(p + "cross_attn.q_proj.1", hidden, hidden, scale=stress / np.sqrt(hidden))
        for t, side in enumerate(window_sides):
            ln(p + f"cross_attn.k_proj_{t}.0", hidden)
            lin(p + f"cross_attn.k_proj_{t}.1", hidden, hidden, scale=stress / np.sqrt(hidden))
            ln(p + f"cross_attn.v_proj_{t}.0", hidden)
            lin(p + f"cross_attn.v_proj_{t}.1", hidden, hidden)
            if side > 1:
                sd[p + f"pos_embed_{t}"] = rs.standard_normal((side * side, hidden)).astype(f32)
        lin(p + "cross_attn.o_proj", hidden, hidden)
    return sd
