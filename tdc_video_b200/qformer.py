"""Drop-in for the reference Q-Former module on the TDC path.

`TDCQFormer` stands where `BertLMHeadModel` stands in the reference (`model.Qformer`,
tdc/cambrian_arch.py:403-424): `model.Qformer.bert(input_ids=, query_embeds=,
encoder_hidden_states=, encoder_attention_mask=, use_cache=False, return_dict=True)` returns an
object with `.last_hidden_state [B, K+T, hidden]` (call site cambrian_arch.py:1653-1662).

Parameter names and shapes are exactly those of tdc/Qformer.py, so a
reference checkpoint's `Qformer.*` entries load with `load_state_dict(strict=True)`.  The
modules below are parameter containers only: the forward pass is one call into
libtdc_b200.so (sm_100a kernels) through `QFormerEngine`.  Inference only — like every
reference call site on this path the module must be in eval mode; there is no CPU path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
from torch import nn

from .engine import QFormerEngine


@dataclass
class QFormerConfig:
    """The BertConfig fields the reference sets/uses (cambrian_arch.py:405-412; BertConfig
    defaults = bert-base-uncased)."""
    vocab_size: int = 30522
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    max_position_embeddings: int = 512
    layer_norm_eps: float = 1e-12
    encoder_width: int = 3584
    add_cross_attention: bool = True
    cross_attention_freq: int = 2
    query_length: int = 16
    initializer_range: float = 0.02
    pad_token_id: int = 0


@dataclass
class QFormerOutput:
    """Field names of BaseModelOutputWithPoolingAndCrossAttentions (Qformer.py:958-965)."""
    last_hidden_state: torch.Tensor
    pooler_output: Optional[torch.Tensor] = None
    past_key_values: Optional[Tuple] = None
    hidden_states: Optional[Tuple] = None
    attentions: Optional[Tuple] = None
    cross_attentions: Optional[Tuple] = None

    def __getitem__(self, i):
        return (self.last_hidden_state, self.pooler_output)[i]


class _Params(nn.Module):
    """Named parameter container (never called)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: the computation runs in libtdc_b200.so")


def _attention_block(hidden: int, kv_width: int, eps: float) -> nn.Module:
    blk = _Params()
    blk.self = _Params()
    blk.self.query = nn.Linear(hidden, hidden)
    blk.self.key = nn.Linear(kv_width, hidden)
    blk.self.value = nn.Linear(kv_width, hidden)
    blk.output = _Params()
    blk.output.dense = nn.Linear(hidden, hidden)
    blk.output.LayerNorm = nn.LayerNorm(hidden, eps=eps)
    return blk


def _ffn_blocks(hidden: int, inter: int, eps: float):
    up = _Params()
    up.dense = nn.Linear(hidden, inter)
    down = _Params()
    down.dense = nn.Linear(inter, hidden)
    down.LayerNorm = nn.LayerNorm(hidden, eps=eps)
    return up, down


class TDCBertModel(nn.Module):
    """Mirror of `BertModel` (tdc/Qformer.py:677-965) for the argument set the TDC path uses."""

    def __init__(self, config: QFormerConfig):
        super().__init__()
        self.config = config
        c = config
        if c.hidden_size != c.num_attention_heads * 64:
            raise ValueError("libtdc_b200 supports head size 64 (hidden_size == 64 * num_attention_heads)")
        self.embeddings = _Params()
        self.embeddings.word_embeddings = nn.Embedding(c.vocab_size, c.hidden_size, padding_idx=c.pad_token_id)
        self.embeddings.position_embeddings = nn.Embedding(c.max_position_embeddings, c.hidden_size)
        self.embeddings.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.embeddings.register_buffer("position_ids", torch.arange(c.max_position_embeddings).expand((1, -1)))
        self.encoder = _Params()
        layers = []
        for i in range(c.num_hidden_layers):
            layer = _Params()
            layer.attention = _attention_block(c.hidden_size, c.hidden_size, c.layer_norm_eps)
            if c.add_cross_attention and i % c.cross_attention_freq == 0:
                layer.crossattention = _attention_block(c.hidden_size, c.encoder_width, c.layer_norm_eps)
            layer.intermediate, layer.output = _ffn_blocks(c.hidden_size, c.intermediate_size, c.layer_norm_eps)
            layer.intermediate_query, layer.output_query = _ffn_blocks(c.hidden_size, c.intermediate_size,
                                                                       c.layer_norm_eps)
            layers.append(layer)
        self.encoder.layer = nn.ModuleList(layers)
        self.apply(self._init_weights)
        self._engine: Optional[QFormerEngine] = None
        self._engine_key = None
        # Verify that encoder_attention_mask is a prefix mask.  Off by default: the check reads device data on
        # the host (one sync per call, breaks CUDA-graph capture) and the reference only ever passes all-ones
        # masks (cambrian_arch.py:1648-1650).  Turn it on when feeding hand-made masks.
        self.check_masks = False
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_engine())

    def _init_weights(self, m):
        # same statistics as tdc/Qformer.py:664-674
        if isinstance(m, (nn.Linear, nn.Embedding)):
            m.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        elif isinstance(m, nn.LayerNorm):
            m.bias.data.zero_()
            m.weight.data.fill_(1.0)
        if isinstance(m, nn.Linear) and m.bias is not None:
            m.bias.data.zero_()

    # -- engine management ----------------------------------------------------------------
    def invalidate_engine(self):
        """Call after mutating parameters in place (load_state_dict / .to() do it for you)."""
        self._engine_key = None

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._engine_key = None
        return out

    def _weights_version(self) -> int:
        """Cheap fingerprint of in-place edits: every parameter's autograd version counter."""
        return sum(int(p._version) for p in self.parameters())

    def engine(self, d_out: int = 0, extra_state=None, extra_key=None, d_frame_in: int = 0,
               d_audio: int = 0) -> QFormerEngine:
        """The libtdc handle holding this module's weights (rebuilt when they move/change).
        `extra_state` adds sibling tensors (vision_proj.*); `extra_key` identifies their version.
        A handle that carries vision_proj also serves the plain `.bert(...)` forward (d_out = 0 requests reuse
        it), so alternating `Qformer.bert(...)` and `TDCCompressor.compress_video(...)` keeps ONE handle."""
        p = self.embeddings.LayerNorm.weight
        if d_out == 0 and extra_state is None and self._engine is not None and self._engine_key is not None \
                and self._engine_key[0] == p.device and self._engine_key[3] == self._weights_version():
            return self._engine
        key = (p.device, d_out, extra_key, self._weights_version(), d_frame_in, d_audio)
        if self._engine is None or self._engine_key != key:
            if p.device.type != "cuda":
                raise RuntimeError("TDCBertModel runs on a CUDA (sm_100a) device only: move the module to the GPU; "
                                   "there is no CPU fallback")
            c = self.config
            if self._engine is not None:
                self._engine.close()
            self._engine = QFormerEngine(hidden=c.hidden_size, heads=c.num_attention_heads,
                                         intermediate=c.intermediate_size, layers=c.num_hidden_layers,
                                         cross_freq=c.cross_attention_freq, d_enc=c.encoder_width, d_out=d_out,
                                         vocab=c.vocab_size, max_pos=c.max_position_embeddings,
                                         ln_eps=c.layer_norm_eps, device=p.device, d_frame_in=d_frame_in,
                                         d_audio=d_audio)
            state = {k: v for k, v in self.state_dict().items() if v.dtype.is_floating_point}
            if extra_state:
                state.update(extra_state)
            self._engine.load_weights(state)
            self._engine_key = key
        return self._engine

    # -- reference-shaped forward ---------------------------------------------------------
    def forward(self, input_ids=None, attention_mask=None, position_ids=None, head_mask=None, query_embeds=None,
                encoder_hidden_states=None, encoder_attention_mask=None, past_key_values=None, use_cache=None,
                output_attentions=None, output_hidden_states=None, return_dict=None, is_decoder=False):
        if self.training:
            raise RuntimeError("TDCBertModel is inference-only: call .eval() (the TDC call sites run under eval)")
        if query_embeds is None:
            raise ValueError("You have to specify query_embeds (the TDC path always does; Qformer.py:855-857)")
        if encoder_hidden_states is None:
            raise ValueError("encoder_hidden_states is required (cross-attention layers; Qformer.py:433-435)")
        for name, val in (("attention_mask", attention_mask), ("position_ids", position_ids),
                          ("head_mask", head_mask), ("past_key_values", past_key_values)):
            if val is not None:
                raise NotImplementedError(f"{name} is not used on the TDC path and is not supported")
        if is_decoder or output_attentions or output_hidden_states or use_cache:
            raise NotImplementedError("is_decoder / output_attentions / output_hidden_states / use_cache "
                                      "are not supported (TDC calls with use_cache=False)")
        kv_len = None
        if encoder_attention_mask is not None:
            m = encoder_attention_mask
            if m.dim() != 2 or m.shape != encoder_hidden_states.shape[:2]:
                raise ValueError("encoder_attention_mask must be [batch, kv_tokens]")
            if self.check_masks and m.shape[1] > 1 and not bool((m[:, :-1] >= m[:, 1:]).all()):
                raise NotImplementedError("only prefix (right-padded) encoder_attention_mask is supported; "
                                          "the reference passes all-ones (cambrian_arch.py:1648-1650)")
            kv_len = m.to(torch.int32).sum(-1, dtype=torch.int32)
        out = self.engine().forward(query_embeds, encoder_hidden_states, input_ids, kv_len=kv_len,
                                    out_dtype=query_embeds.dtype)
        if return_dict is False:
            return (out, None)
        return QFormerOutput(last_hidden_state=out)


class TDCQFormer(nn.Module):
    """Mirror of the `BertLMHeadModel` container the reference stores at `model.Qformer`
    (tdc/Qformer.py:968-979).  `.bert` is the compute path; `.cls` holds the (unused on this
    path, Qformer.py:607-651) LM head parameters only so that reference checkpoints load
    strictly."""

    def __init__(self, config: QFormerConfig, with_lm_head: bool = True):
        super().__init__()
        self.config = config
        self.bert = TDCBertModel(config)
        if with_lm_head:
            c = config
            self.cls = _Params()
            self.cls.predictions = _Params()
            self.cls.predictions.transform = _Params()
            self.cls.predictions.transform.dense = nn.Linear(c.hidden_size, c.hidden_size)
            self.cls.predictions.transform.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
            self.cls.predictions.decoder = nn.Linear(c.hidden_size, c.vocab_size, bias=False)
            self.cls.predictions.bias = nn.Parameter(torch.zeros(c.vocab_size))
            self.cls.predictions.decoder.bias = self.cls.predictions.bias  # tied, as in Qformer.py:627-631

    def forward(self, *args, **kwargs):
        raise NotImplementedError("the TDC path calls `.bert(...)` (cambrian_arch.py:1653); the LM head is unused")
