"""TEST INFRASTRUCTURE ONLY — literal CPU restatement of the reference TDC chunk loop.

Follows tdc/cambrian_arch.py line by line, *including* its Python double loop and its
<= 7-row Q-Former calls (so it doubles as the reference-batching CPU baseline):

    :1541-1545  split frames into DINO segments
    :1603-1606  for segment: for chunk of 8 frames
    :1609       key frame = chunk[0], taken BEFORE the audio concat
    :1611-1614  chunk = cat([chunk, audio_proj(audio_chunk)], dim=1)
    :1617-1623  single-frame chunk -> static frame + frame_seg, no Q-Former
    :1625-1640  queries: adaptive_avg_pool1d(key frame -> K) -> query_proj | learned query_tokens
    :1643-1662  Qformer.bert(input_ids | None, query_embeds, encoder_hidden_states=other frames, ones mask)
    :1664-1667  F.normalize(vision_proj(last_hidden_state[:, :K]), dim=-1)
    :1668-1692  [static, frame_seg, (K tokens, frame_seg) per frame]
    :1694-1709  budget truncation: drop ceil(excess/num_chunks) trailing tokens of every chunk, hard clip

PINNING: the Q-Former inside is oracle/qformer_oracle.py (pinned to the reference modules and
golden vectors).  The loop/assembly logic itself is pinned against the reference's real
`prepare_inputs_labels_for_multimodal` by oracle/harness.py where /root/reference exists
(tests/test_driver_pinning.py); see DESIGN.md for the status of that pin.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

from . import qformer_oracle as qo


def adapt_segment(frame_features, max_num_segments: int = 24, window_size: int = 64):
    """Per-video body of CambrianMetaForCausalLM.adapt_segment (cambrian_arch.py:801-849), fp32 CPU:
    returns (selected_frame_indices, segment_frame_indices, cos_similarities | None)."""
    f = qo._t(frame_features)
    n = len(f)
    if n <= max_num_segments + 1:                                   # :803-810
        return torch.arange(n), torch.arange(n), None
    max_num_frames = 224                                            # :813
    if n > max_num_frames:                                          # :815-822
        interval = n / float(max_num_frames)
        indices = [int(interval * i) for i in range(max_num_frames)]
    else:
        indices = list(range(n))
    q = f[indices].flatten(1, 2) if f.dim() == 3 else f[indices]    # :832
    prev, nxt = q[:-1], q[1:]
    sims = []
    for s0 in range(0, len(prev), window_size):                     # :836-841 (windowing does not change values)
        sims.append(F.cosine_similarity(prev[s0:s0 + window_size], nxt[s0:s0 + window_size], dim=1))
    cos = torch.cat(sims)
    seg, _ = torch.argsort(cos)[:max_num_segments].sort()           # :849
    return torch.tensor(indices), seg, cos


def audio_frames_from_beats(window_embeds, sample_indices, n_frames: int, dist: int = 10):
    """Per-frame audio tokens from BEATs window features — cambrian_arch.py:1547-1598, with the BEATs
    forward itself (out of scope) replaced by its outputs: `window_embeds[w]` = `audio_embed` [1, t_w, 768]
    of the w-th 10-second window (50 tokens per second).  `sample_indices[s]` = 1 when second s is a sampled
    frame.  Every sampled frame receives the audio from its own second up to the next sampled second, pooled
    to 50 tokens (adaptive_avg_pool2d over the token axis); the result is zero-padded to n_frames x 50."""
    t = qo._t
    audio_embeds, seg = [], []
    si = [int(v) for v in torch.as_tensor(sample_indices).tolist()]
    for w, k in enumerate(range(0, len(window_embeds) * dist, dist)):
        audio_embed = t(window_embeds[w])
        window = si[k:k + dist]
        sample_len = len(window)
        for idx, indice in enumerate(window):
            token = audio_embed[:, idx * 50:(idx + 1) * 50, :]
            if token.shape[1] == 0:
                continue
            if token.shape[1] != 50:
                token = F.adaptive_avg_pool2d(token, (50, 768))
            if indice == 1:
                if seg:
                    audio_embeds.append(F.adaptive_avg_pool2d(torch.cat(seg, dim=1), (50, 768)))
                    seg = []
                seg.append(token)
                if idx + 1 < sample_len and si[k + idx + 1] == 1:
                    audio_embeds.append(token)
                    seg = []
            elif indice == 0:
                seg.append(token)
    if seg:
        audio_embeds.append(F.adaptive_avg_pool2d(torch.cat(seg, dim=1), (50, 768)))
    out = torch.cat(audio_embeds).flatten(0, 1).unsqueeze(0)
    pad = n_frames * 50 - out.size(1)
    out = F.pad(out, (0, 0, 0, pad, 0, 0), "constant", 0)
    return out.reshape(-1, 50, 768)


def segment_sizes_from_boundaries(segment_frame_indices, n_frames: int):
    """cambrian_arch.py:1541-1544: boundaries after frames `segment_frame_indices` ->
    frames per segment (empty segments are possible and skipped by the loop, :1604-1605)."""
    seg = (torch.as_tensor(segment_frame_indices).long() + 1).tolist()
    points = [0] + seg + [int(n_frames)]
    return [points[i + 1] - points[i] for i in range(len(points) - 1)]


def compress_video(weights: Dict[str, object], geom, visual_emb_frame, segment_sizes: Sequence[int], *,
                   context_token_num: int = 16, query_type: str = "Avg_pool", add_text: bool = True,
                   keep_static: bool = True, add_sep: bool = True, input_ids=None, audio_frames=None,
                   max_visual_len: Optional[int] = None, return_chunks: bool = False):
    """weights: Q-Former state dict (keys relative to `Qformer.bert.`) plus `vision_proj.*`,
    `query_proj.*`, `frame_seg`, optional `audio_proj.*`, `query_tokens`.  All fp32 CPU."""
    t = qo._t
    frames = t(visual_emb_frame)
    segments = torch.split(frames, [int(s) for s in segment_sizes])
    audio_segments = None
    if audio_frames is not None:
        audio_segments = torch.split(t(audio_frames), [int(s) for s in segment_sizes])
    frame_seg = t(weights["frame_seg"])
    out_chunks = []
    n_calls = 0
    for seg_i, segment in enumerate(segments):
        if len(segment) == 0:
            continue
        for start_idx in range(0, len(segment), 8):
            end_idx = min(start_idx + 8, len(segment))
            chunk_feature = segment[start_idx:end_idx]
            key_frame = chunk_feature[0]
            if audio_segments is not None:
                audio_chunk = audio_segments[seg_i][start_idx:end_idx]
                audio_chunk = F.linear(audio_chunk, t(weights["audio_proj.weight"]), t(weights["audio_proj.bias"]))
                chunk_feature = torch.cat([chunk_feature, audio_chunk], dim=1)
            other_frames = chunk_feature[1:]
            if keep_static and len(chunk_feature) == 1:
                out_chunks.append(torch.cat([chunk_feature[0], frame_seg[None]]) if add_sep else chunk_feature[0])
                continue
            visual_input = other_frames if keep_static else chunk_feature
            B = len(visual_input)
            if query_type == "Avg_pool":
                q = qo.avg_pool_queries(key_frame[None], context_token_num)
                q = F.linear(q, t(weights["query_proj.weight"]), t(weights["query_proj.bias"])).expand(B, -1, -1)
            else:
                q = t(weights["query_tokens"]).expand(B, -1, -1)
            ids = None
            if add_text and input_ids is not None:
                ids = torch.as_tensor(input_ids, dtype=torch.long).reshape(1, -1).expand(B, -1)
            hidden = qo.qformer_forward(weights, geom, q, visual_input, ids)
            n_calls += 1
            comp = qo.proj_norm(weights, hidden, q.shape[1])
            if add_sep:
                body = torch.cat([comp, frame_seg[None, None].expand(B, 1, -1)], dim=1).flatten(0, 1)
                head = torch.cat([chunk_feature[0], frame_seg[None]])
            else:
                body = comp.flatten(0, 1)
                head = chunk_feature[0]
            out_chunks.append(torch.cat([head, body], dim=0) if keep_static else body)
    reduced = sum(x.shape[0] for x in out_chunks)
    if max_visual_len is not None and reduced > max_visual_len:
        force_remove = math.ceil((reduced - max_visual_len) / len(out_chunks))
        out_chunks = [x[:-force_remove] for x in out_chunks]
    seq = torch.cat(out_chunks, dim=0)
    if max_visual_len is not None:
        seq = seq[:max_visual_len]
    if return_chunks:
        return seq, out_chunks, n_calls
    return seq
