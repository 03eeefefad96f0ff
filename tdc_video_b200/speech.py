"""Drop-in for the reference's speech Q-Former (tdc/audio_models/audio_encoder.py).

Reference: `AudioEncoder.init_speech_Qformer` (:10-24) builds the same `BertLMHeadModel` as the TDC Q-Former at
another geometry — `num_hidden_layers = 2`, `cross_attention_freq = 1` (cross-attention in EVERY layer),
`num_query_token = 1` — and `_encode_auditory_feature` (:75-116, first definition) runs it per 0.33-second window of
the Whisper (+ BEATs) features: LayerNorm(s), pad + concat on the feature axis, `F.unfold` into windows of
`round(1500 * second_per_window / 30)` = 17 frames, `speech_Qformer.bert(query_embeds=speech_query_tokens, ...)`,
`speech_llama_proj`, reshape back to [B, windows * queries, d_llm].  (In the shipped reference file that method is
shadowed by a second definition at :118 that returns the concatenated features, and `__init__` no longer creates
the modules; the attribute names below are the ones the first definition and SALMONN-style checkpoints use.)

Everything heavy runs in libtdc_b200.so: the LayerNorms (`tdc_layernorm`), the Q-Former (`tdc_qformer_forward`
through `TDCQFormer`, generic in layers / cross_attention_freq / queries) and the projection (`tdc_linear`);
torch only reshapes.  CUDA only, eval only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .engine import _ptr, _stream, linear
from .qformer import QFormerConfig, TDCQFormer


def init_speech_Qformer(num_query_token: int, speech_width: int, num_hidden_layers: int = 2,
                        vocab_size: int = 30522) -> Tuple[TDCQFormer, nn.Parameter]:
    """`AudioEncoder.init_speech_Qformer` (:10-24): (Qformer, query_tokens) with the reference's parameter names."""
    cfg = QFormerConfig(vocab_size=vocab_size, num_hidden_layers=num_hidden_layers, encoder_width=speech_width,
                        add_cross_attention=True, cross_attention_freq=1, query_length=num_query_token)
    qformer = TDCQFormer(cfg)
    query_tokens = nn.Parameter(torch.zeros(1, num_query_token, cfg.hidden_size))
    query_tokens.data.normal_(mean=0.0, std=cfg.initializer_range)
    return qformer, query_tokens


def _layernorm_bf16(x: torch.Tensor, ln: nn.LayerNorm) -> torch.Tensor:
    """LayerNorm over the last axis on the GPU kernel (fp32 statistics) -> bf16."""
    lib = _lib.load_library()
    shape = x.shape
    x2 = x.reshape(-1, shape[-1]).float().contiguous()
    out = torch.empty(x2.shape, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.tdc_layernorm(_ptr(x2), None, 0, _ptr(ln.weight.detach().float().contiguous()),
                               _ptr(ln.bias.detach().float().contiguous()), C.c_float(ln.eps), None, _ptr(out),
                               x2.shape[0], x2.shape[1], _stream(x.device))
    _lib.check(rc, None, "tdc_layernorm")
    return out.reshape(shape)


class TDCSpeechQFormer(nn.Module):
    """The modules `_encode_auditory_feature` uses, under its attribute names: `ln_speech`, `ln_audio`,
    `speech_Qformer`, `speech_query_tokens`, `speech_llama_proj`."""

    def __init__(self, speech_width: int, audio_width: int = 0, llama_hidden_size: int = 4096,
                 num_speech_query_token: int = 1, num_hidden_layers: int = 2, window_level_Qformer: bool = True,
                 second_per_window: float = 0.333333, second_stride: float = 0.333333, vocab_size: int = 30522):
        super().__init__()
        if speech_width > 1536 or audio_width > 1536:
            raise ValueError("feature widths above 1536 are not supported by the LayerNorm kernel")
        self.window_level_Qformer = window_level_Qformer
        self.second_per_window, self.second_stride = second_per_window, second_stride
        self.ln_speech = nn.LayerNorm(speech_width)
        if audio_width:
            self.ln_audio = nn.LayerNorm(audio_width)
        self.speech_Qformer, self.speech_query_tokens = init_speech_Qformer(
            num_speech_query_token, speech_width + audio_width, num_hidden_layers, vocab_size)
        self.speech_llama_proj = nn.Linear(self.speech_Qformer.config.hidden_size, llama_hidden_size)
        self.eval()

    @torch.no_grad()
    def encode_auditory_feature(self, speech_embeds: torch.Tensor, audio_embeds: Optional[torch.Tensor] = None):
        """speech_embeds [B, T, C_s] (Whisper encoder output), audio_embeds [B, T', C_a] (BEATs) or None ->
        (speech tokens [B, windows * queries, d_llm] bf16, attention mask of ones)."""
        if self.training:
            raise RuntimeError("TDCSpeechQFormer is inference-only (eval mode)")
        if not speech_embeds.is_cuda:
            raise RuntimeError("TDCSpeechQFormer needs CUDA tensors: there is no CPU fallback")
        x = _layernorm_bf16(speech_embeds, self.ln_speech)
        if audio_embeds is not None:
            a = _layernorm_bf16(audio_embeds, self.ln_audio)
            if a.size(1) < x.size(1):
                a = F.pad(a, (0, 0, 0, x.size(1) - a.size(1)))
            elif a.size(1) > x.size(1):
                x = F.pad(x, (0, 0, 0, a.size(1) - x.size(1)))
            x = torch.cat((x, a), dim=-1)
        B, T, Cw = x.shape
        if self.window_level_Qformer:
            kernel = round(1500 * self.second_per_window / 30.0)
            stride = round(1500 * self.second_stride / 30.0)
            # F.unfold over the time axis == windows [t0, t0 + kernel) at t0 = 0, stride, 2*stride, ...
            win = x.unfold(1, kernel, stride)                       # [B, L, C, kernel] (a view)
            x = win.permute(0, 1, 3, 2).reshape(-1, kernel, Cw)     # [B*L, kernel, C]
        q = self.speech_query_tokens.detach().float()
        rows = x.shape[0]
        hidden = self.speech_Qformer.bert.engine().forward(q, x.contiguous(), None,
                                                           query_set=torch.zeros(rows, dtype=torch.int32),
                                                           out_dtype=torch.bfloat16)
        y = linear(hidden, self.speech_llama_proj.weight, self.speech_llama_proj.bias, out_dtype=torch.bfloat16)
        if self.window_level_Qformer:
            y = y.view(B, -1, y.size(2)).contiguous()
        atts = torch.ones(y.size()[:-1], dtype=torch.long, device=y.device)
        return y, atts
