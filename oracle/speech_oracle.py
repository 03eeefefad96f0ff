"""TEST INFRASTRUCTURE ONLY — CPU (fp32 torch) restatement of the speech Q-Former wrapper,
tdc/audio_models/audio_encoder.py:75-116 (`_encode_auditory_feature`, first definition) with the Q-Former built
by `init_speech_Qformer` (:10-24: 2 layers, cross_attention_freq = 1, 1 query).

Pinning: the Q-Former core (`qformer_oracle.qformer_forward`) is pinned to the reference's BertModel at exactly
this kind of geometry (tests/golden/qformer_freq1_speech_style.npz, tests/test_oracle_pinning.py).  The wrapper around it cannot
be executed from the reference: in the shipped file the method is shadowed by a second definition (:118) and
`__init__` no longer creates `ln_speech` / `speech_Qformer` / `speech_llama_proj`.  It is restated here with the
very torch calls the reference uses (nn.LayerNorm semantics, F.pad, torch.cat, F.unfold, nn.Linear semantics).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import qformer_oracle as qo


def encode_auditory_feature(sd, geom, speech_embeds, audio_embeds=None, second_per_window=0.333333,
                            second_stride=0.333333, window_level=True):
    """sd: Q-Former keys relative to `speech_Qformer.bert.` + `ln_speech.*`, `ln_audio.*`, `speech_query_tokens`,
    `speech_llama_proj.*`.  Returns speech tokens [B, windows * queries, d_llm]."""
    t = qo._t
    x = F.layer_norm(t(speech_embeds), (t(speech_embeds).shape[-1],), t(sd["ln_speech.weight"]), t(sd["ln_speech.bias"]), 1e-5)
    if audio_embeds is not None:
        a = t(audio_embeds)
        a = F.layer_norm(a, (a.shape[-1],), t(sd["ln_audio.weight"]), t(sd["ln_audio.bias"]), 1e-5)
        if a.size(1) < x.size(1):
            a = F.pad(a, (0, 0, 0, x.size(1) - a.size(1)))
        elif a.size(1) > x.size(1):
            x = F.pad(x, (0, 0, 0, a.size(1) - x.size(1)))
        x = torch.cat((x, a), dim=-1)
    B, T, C = x.shape
    if window_level:
        kernel = (1, round(1500 * second_per_window / 30.0))
        stride = (1, round(1500 * second_stride / 30.0))
        tr = x.transpose(1, 2).unsqueeze(2)
        ov = F.unfold(tr, kernel_size=kernel, dilation=1, padding=0, stride=stride)
        _, _, L = ov.shape
        ov = ov.view(B, -1, kernel[1], L)
        ov = torch.permute(ov, [0, 3, 2, 1])
        x = ov.reshape(-1, kernel[1], C)
    q = t(sd["speech_query_tokens"]).expand(x.shape[0], -1, -1)
    h = qo.qformer_forward(sd, geom, q, x, None)
    y = F.linear(h, t(sd["speech_llama_proj.weight"]), t(sd["speech_llama_proj.bias"]))
    if window_level:
        y = y.view(B, -1, y.size(2)).contiguous()
    return y
