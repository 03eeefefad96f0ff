"""The oracle of the upstream ("frames") entry — GELU-MLP mm_projector + newline tokens (make_golden.driver_frames)
followed by the chunk loop (driver_oracle.compress_video) — against committed outputs of the reference's real
`prepare_inputs_labels_for_multimodal` run with the nn.Sequential(Linear, GELU, Linear) projector
(tests/golden/towers_*.npz, written by oracle/make_golden.py).  Runs anywhere (CPU, no reference tree)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import driver_oracle
from oracle.make_golden import DRIVER_GEOM, driver_audio, driver_frames, driver_tables, driver_weights_mlp

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "towers_*.npz")))


def oracle_sequence(z, m):
    w = driver_weights_mlp(m["weight_seed"], m["num_query"])
    sig, dino = driver_tables(m["table_seed"], m["n_frames"])
    n = m["n_frames"]
    frames = driver_frames(w, sig, dino)
    sizes = driver_oracle.segment_sizes_from_boundaries(z["segment_frame_indices"], n)
    audio_frames = None
    if m.get("audio"):
        windows, flags, _, proj = driver_audio(m["audio_seed"], n, m["audio"])
        w.update(proj)
        audio_frames = driver_oracle.audio_frames_from_beats(windows, flags, n)
    return driver_oracle.compress_video(w, DRIVER_GEOM, frames, sizes, context_token_num=m["num_query"],
                                        query_type=m["query_type"], add_text=m["text"], keep_static=m["add_static"],
                                        input_ids=torch.tensor([m["prompt_ids"]]), max_visual_len=m["max_visual_len"],
                                        audio_frames=audio_frames)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[7:-4] for p in GOLDEN])
def test_frames_oracle_matches_reference_golden(path):
    z = np.load(path)
    m = json.loads(str(z["meta"]))
    assert m["projector"] == "mlp2x_gelu"
    got = oracle_sequence(z, m)
    assert got.shape == z["visual_tokens"].shape
    assert float(np.abs(got.numpy() - z["visual_tokens"]).max()) <= 2e-5


def test_tower_goldens_exist():
    assert len(GOLDEN) >= 3
