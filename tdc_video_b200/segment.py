"""Adaptive segmentation on the GPU — `CambrianMetaForCausalLM.adapt_segment`
(tdc/cambrian_arch.py:783-861) for one video: the step that produces the path's
`segment_frame_indices`.

Host side (pure integers, as in the reference): <= max_num_segments+1 frames -> every frame
is its own segment (:803-810); > 224 frames -> uniform subsample to 224 (:813-822).
Device side (libtdc_b200.so, `tdc_segment_boundaries`): cosine similarity of consecutive frames
over the flattened DINO features (:832-842) and `sort(argsort(sims)[:max_num_segments])` (:849).
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import _lib
from .engine import _dt, _ptr, _stream

MAX_FRAMES = 224  # cambrian_arch.py:813


def frame_cosine_and_boundaries(features: torch.Tensor, max_num_segments: int = 24) -> Tuple[torch.Tensor, torch.Tensor]:
    """features [n, ...] (CUDA; bf16/fp16/fp32) -> (cos [n-1] fp32, boundaries [min(k, n-1)] int64), on device."""
    if not features.is_cuda:
        raise RuntimeError("tdc_video_b200.segment needs CUDA tensors: there is no CPU fallback")
    lib = _lib.load_library()
    n = features.shape[0]
    feats = features.reshape(n, -1).contiguous()
    dim = feats.shape[1]
    dev = feats.device
    cos = torch.empty(max(n - 1, 0), dtype=torch.float32, device=dev)
    k = min(max_num_segments, max(n - 1, 0))
    bounds = torch.empty(k, dtype=torch.int64, device=dev)
    if n < 2:
        return cos, bounds
    ws = torch.empty(max(int(lib.tdc_segment_workspace_bytes(n, dim)), 16), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.tdc_segment_boundaries(_ptr(feats), _dt(feats), n, dim, max_num_segments, _ptr(cos), _ptr(bounds),
                                        _ptr(ws), ws.numel(), _stream(dev))
    _lib.check(rc, None, "tdc_segment_boundaries")
    return cos, bounds


def adapt_segment(features: torch.Tensor, max_num_segments: int = 24):
    """One video's DINO features [n_frames, tokens, C] ->
    (selected_frame_indices [m] int64 (CPU), segment_frame_indices [s] int64 (device or CPU), cos or None).
    Mirrors the per-video body of the reference's loop; `selected_frame_indices` are the frames kept
    after the > 224-frame subsample, `segment_frame_indices` index into the selected frames."""
    n = features.shape[0]
    if n <= max_num_segments + 1:
        idx = torch.arange(n)
        return idx, idx.clone(), None
    if n > MAX_FRAMES:
        interval = n / float(MAX_FRAMES)
        selected = torch.tensor([int(interval * i) for i in range(MAX_FRAMES)])
        feats = features.index_select(0, selected.to(features.device))
    else:
        selected = torch.arange(n)
        feats = features
    cos, bounds = frame_cosine_and_boundaries(feats, max_num_segments)
    return selected, bounds, cos


def segment_sizes(segment_frame_indices, n_frames: int) -> List[int]:
    """cambrian_arch.py:1541-1544: boundaries after the given frames -> frames per segment."""
    seg = (torch.as_tensor(segment_frame_indices).cpu().long() + 1).tolist()
    points = [0] + seg + [int(n_frames)]
    return [points[i + 1] - points[i] for i in range(len(points) - 1)]
