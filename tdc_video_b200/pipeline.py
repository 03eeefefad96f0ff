"""The TDC stage of `prepare_inputs_labels_for_multimodal` for one video, device-resident end to end.

Reference flow (tdc/cambrian_arch.py), with the piece of this package that replaces each step:

    :946-960, 783-861   adapt_segment on the DINO features          -> segment.adapt_segment (CUDA)
    :1149-1150          mm_projector on the frame tokens            -> projector.GeluMLPProjector / engine.linear
    :1269-1281          image_newline appended to every token row   -> append_newline_tokens (data movement)
    :1541-1545          segment boundaries -> frames per segment    -> segment.segment_sizes (host integers)
    :1547-1598          per-frame audio tokens from BEATs windows   -> audio.pool_audio_per_frame (CUDA pooling)
    :1603-1709          chunk loop: Q-Former, vision_proj, assembly -> compressor.TDCCompressor.compress_video

`tdc_video_stage` strings them together so that a caller hands over the tower outputs of a video and gets the
visual token sequence the reference splices into the prompt (`image_features[cur_image_idx]` after :1709).
CUDA only; nothing here computes on the host except the integer bookkeeping the reference also does in Python.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from .audio import pool_audio_per_frame
from .compressor import TDCCompressor
from .segment import adapt_segment, segment_sizes


def append_newline_tokens(frame_tokens: torch.Tensor, image_newline: torch.Tensor, side: Optional[int] = None
                          ) -> torch.Tensor:
    """[n, side*side, d] -> [n, side*(side+1), d]: the learned `image_newline` vector closes every row of the
    token grid (cambrian_arch.py:1269-1281: cat along the width, then flatten)."""
    n, tokens, d = frame_tokens.shape
    side = side or int(round(tokens ** 0.5))
    if side * side != tokens:
        raise ValueError(f"{tokens} tokens per frame is not a square grid")
    out = torch.empty((n, side, side + 1, d), dtype=frame_tokens.dtype, device=frame_tokens.device)
    out[:, :, :side].copy_(frame_tokens.view(n, side, side, d))
    out[:, :, side] = image_newline.to(frame_tokens.device, frame_tokens.dtype)
    return out.view(n, side * (side + 1), d)


@torch.no_grad()
def tdc_video_stage(compressor: TDCCompressor, visual_emb_frame: Optional[torch.Tensor], dino_features: torch.Tensor, *,
                    input_ids: Optional[torch.Tensor] = None, audio_windows: Optional[Sequence[torch.Tensor]] = None,
                    sample_indices=None, max_visual_len: Optional[int] = None, max_num_segments: int = 24,
                    shard: bool = False, return_segments: bool = False,
                    tower_features: Optional[torch.Tensor] = None, fold: bool = True):
    """One video through adaptive segmentation and TDC compression.

    visual_emb_frame [n_frames, Lv, d_llm]  projected frame tokens incl. newline tokens (CUDA)
    dino_features    [n_frames, tokens, C]  the DINO tower's features of the same frames (CUDA), the input of the
                                            reference's adapt_segment (:946-960)
    input_ids        [1, T]                 BERT ids of the prompt (text_input mode)
    audio_windows / sample_indices          BEATs features [1, t, 768] of the 10-second windows and the 0/1
                                            "a frame was sampled in this second" flags (:1547-1598); None = silent
    tower_features   [n_frames, Tv, C]      ALTERNATIVE to visual_emb_frame (pass None there): the concatenated tower
                                            features, i.e. the INPUT of mm_projector (:1149); the projector, the
                                            newline tokens and everything after run in one tdc_compress_frames call
                                            (`compressor` built with mm_input_size=C; `fold` see the C header)
    Returns the visual token sequence [tokens, d_llm] (and, if asked, the selected frames / boundaries)."""
    from_towers = tower_features is not None
    if from_towers:
        if visual_emb_frame is not None:
            raise ValueError("pass either visual_emb_frame or tower_features")
        visual_emb_frame = tower_features
    if not visual_emb_frame.is_cuda or not dino_features.is_cuda:
        raise RuntimeError("tdc_video_stage needs CUDA tensors: there is no CPU fallback")
    if visual_emb_frame.shape[0] != dino_features.shape[0]:
        raise ValueError("visual_emb_frame and dino_features must describe the same frames")
    selected, boundaries, _ = adapt_segment(dino_features, max_num_segments)
    if len(selected) != visual_emb_frame.shape[0]:          # > 224 frames: the reference keeps a uniform subsample
        visual_emb_frame = visual_emb_frame.index_select(0, selected.to(visual_emb_frame.device))
    n = visual_emb_frame.shape[0]
    sizes = segment_sizes(boundaries, n)
    audio_frames = None
    if audio_windows is not None:
        # sample_indices: 0/1 per SECOND of the clip, 1 = a frame was sampled there (default: 1 fps, every second).
        # When frames were dropped by the > 224 subsample, only the kept frames' seconds stay flagged (:917-925).
        n_in = dino_features.shape[0]
        si = torch.ones(n_in, dtype=torch.int16) if sample_indices is None else torch.as_tensor(sample_indices).cpu()
        if len(selected) != n_in:
            pos = torch.where(si == 1)[0]
            kept = torch.zeros_like(si)
            kept[pos[selected]] = 1
            si = kept
        audio_frames = pool_audio_per_frame(audio_windows, si, n)
    if from_towers:
        seq = compressor.compress_video_from_towers(visual_emb_frame, sizes, input_ids=input_ids,
                                                    audio_frames=audio_frames, max_visual_len=max_visual_len, fold=fold)
    else:
        seq = compressor.compress_video(visual_emb_frame, sizes, input_ids=input_ids, audio_frames=audio_frames,
                                        max_visual_len=max_visual_len, shard=shard)
    if return_segments:
        return seq, selected, boundaries
    return seq
