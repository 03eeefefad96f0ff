"""CPU checks of the speech Q-Former wrapper (audio_encoder.py:75-116): the product's windowing (a strided view) is the
reference's F.unfold sequence, and the oracle's shapes follow the reference (windows * queries tokens per clip)."""
import torch
import torch.nn.functional as F

from oracle import speech_oracle
from oracle.synth import QFormerGeometry, make_state_dict


def test_strided_view_equals_the_reference_unfold_sequence():
    B, T, C, kernel, stride = 2, 53, 6, 17, 17
    x = torch.randn(B, T, C)
    tr = x.transpose(1, 2).unsqueeze(2)
    ov = F.unfold(tr, kernel_size=(1, kernel), dilation=1, padding=0, stride=(1, stride))
    ov = ov.view(B, -1, kernel, ov.shape[-1])
    ref = torch.permute(ov, [0, 3, 2, 1]).reshape(-1, kernel, C)
    got = x.unfold(1, kernel, stride).permute(0, 1, 3, 2).reshape(-1, kernel, C)
    assert torch.equal(got, ref)
    got2 = x.unfold(1, kernel, 5).permute(0, 1, 3, 2).reshape(-1, kernel, C)          # overlapping windows
    ov = F.unfold(tr, kernel_size=(1, kernel), stride=(1, 5))
    ref2 = torch.permute(ov.view(B, -1, kernel, ov.shape[-1]), [0, 3, 2, 1]).reshape(-1, kernel, C)
    assert torch.equal(got2, ref2)


def test_speech_oracle_shapes():
    geom = QFormerGeometry(hidden=64, heads=1, intermediate=128, layers=2, cross_freq=1, d_enc=40, d_out=0, vocab=0)
    sd = {k: torch.from_numpy(v) for k, v in make_state_dict(geom, 3, with_text=False, with_vision_proj=False).items()}
    sd.update({"ln_speech.weight": torch.ones(24), "ln_speech.bias": torch.zeros(24), "ln_audio.weight": torch.ones(16),
               "ln_audio.bias": torch.zeros(16), "speech_query_tokens": torch.randn(1, 1, 64) * 0.02,
               "speech_llama_proj.weight": torch.randn(32, 64) * 0.1, "speech_llama_proj.bias": torch.zeros(32)})
    y = speech_oracle.encode_auditory_feature(sd, geom, torch.randn(2, 40, 24), torch.randn(2, 37, 16))
    assert y.shape == (2, 2, 32)          # 40 frames -> 2 windows of 17, 1 query each
