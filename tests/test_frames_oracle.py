"""frames_oracle.frames_stage (all rows of all chunks in one batched Q-Former call) equals the pinned composition
driver_frames + driver_oracle.compress_video (chunk-by-chunk Python loop, checked against the real reference in
test_towers_golden.py) on the same data — CPU only."""
import numpy as np
import torch

from oracle import driver_oracle, frames_oracle
from oracle.make_golden import DRIVER_GEOM, driver_audio, driver_frames, driver_tables, driver_weights_mlp
from tdc_video_b200.compressor import plan_chunks


def test_frames_stage_equals_the_pinned_composition():
    n, K = 27, 8
    w = driver_weights_mlp(7, K)
    sig, dino = driver_tables(8, n)
    windows, flags, _, proj = driver_audio(9, n, "sparse")
    w.update(proj)
    audio = driver_oracle.audio_frames_from_beats(windows, flags, n)
    sizes = [9, 1, 12, 5]
    feats = torch.from_numpy(np.concatenate([sig, dino], -1))
    plan = plan_chunks(sizes, True)
    ids = torch.tensor([[3, 9, 4]])
    static, comp = frames_oracle.frames_stage(w, DRIVER_GEOM, feats, audio, plan.static_frames, plan.chunk_len, K, ids)
    assert torch.allclose(frames_oracle.project_frames(w, feats), driver_frames(w, sig, dino), atol=1e-6)
    seq, chunks, _ = driver_oracle.compress_video(w, DRIVER_GEOM, driver_frames(w, sig, dino), sizes, context_token_num=K,
                                                  input_ids=ids, audio_frames=audio, return_chunks=True)
    # rebuild the sequence from (static, comp) the way the reference interleaves them
    seg = torch.from_numpy(w["frame_seg"])
    out, r = [], 0
    for c in range(plan.num_chunks):
        out += [static[c], seg[None]]
        for _ in range(int(plan.rows_per_chunk[c])):
            out += [comp[r], seg[None]]
            r += 1
    assert torch.allclose(torch.cat(out), seq, atol=2e-5)
