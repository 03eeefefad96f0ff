"""Per-frame audio tokens for the TDC path from BEATs window features (tdc/cambrian_arch.py:1547-1598).

The BEATs encoder itself is upstream and out of scope; this module takes its outputs — one
`[1, t_w, 768]` tensor per 10-second window, 50 tokens per second — and produces the `[n_frames, 50, 768]`
tensor `TDCCompressor.compress_video(audio_frames=...)` consumes: every sampled frame gets the audio from
its own second up to the next sampled second, average-pooled over the token axis to 50 tokens
(`adaptive_avg_pool2d(x, (50, 768))` only pools tokens because the width stays 768), zero-padded at the end.

Host side: integer bookkeeping of which seconds feed which frame.  Device side: the pooling runs on the
`tdc_avg_pool_tokens` kernel (same bins as adaptive_avg_pool1d: [floor(i*L/K), ceil((i+1)*L/K))).
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from .engine import avg_pool_tokens

TOKENS_PER_SECOND = 50
WINDOW_SECONDS = 10


def _pool_to_50(x: torch.Tensor) -> torch.Tensor:
    """[1, t, 768] -> [1, 50, 768] (the kernel writes bf16; cast back to the input dtype)."""
    return avg_pool_tokens(x.contiguous(), TOKENS_PER_SECOND).to(x.dtype)


def pool_audio_per_frame(window_embeds: Sequence[torch.Tensor], sample_indices, n_frames: int) -> torch.Tensor:
    """window_embeds[w]: [1, t_w, 768] CUDA tensor of window w; sample_indices: 0/1 per second;
    returns [n_frames, 50, 768] in the dtype of the embeddings."""
    if len(window_embeds) == 0:
        raise ValueError("no audio windows")
    if not window_embeds[0].is_cuda:
        raise RuntimeError("tdc_video_b200.audio needs CUDA tensors: there is no CPU fallback")
    si = [int(v) for v in torch.as_tensor(sample_indices).tolist()]
    pieces: List[torch.Tensor] = []     # finished [1, 50, 768] per sampled frame
    pending: List[torch.Tensor] = []
    for w, embed in enumerate(window_embeds):
        k = w * WINDOW_SECONDS
        window = si[k:k + WINDOW_SECONDS]
        for idx, flag in enumerate(window):
            token = embed[:, idx * TOKENS_PER_SECOND:(idx + 1) * TOKENS_PER_SECOND, :]
            if token.shape[1] == 0:
                continue
            if token.shape[1] != TOKENS_PER_SECOND:
                token = _pool_to_50(token)
            if flag == 1:
                if pending:
                    pieces.append(_pool_to_50(torch.cat(pending, dim=1)))
                    pending = []
                pending.append(token)
                if idx + 1 < len(window) and si[k + idx + 1] == 1:
                    pieces.append(token)
                    pending = []
            elif flag == 0:
                pending.append(token)
    if pending:
        pieces.append(_pool_to_50(torch.cat(pending, dim=1)))
    out = torch.cat(pieces, dim=0)                          # [m, 50, 768]
    if out.shape[0] > n_frames:
        # a clip that starts with unsampled seconds yields one leading extra piece; the reference's
        # F.pad(..., pad_size < 0) then silently crops the END of the sequence (cambrian_arch.py:1591-1595)
        out = out[:n_frames]
    if out.shape[0] < n_frames:
        out = torch.cat([out, out.new_zeros((n_frames - out.shape[0], TOKENS_PER_SECOND, out.shape[-1]))], dim=0)
    return out
