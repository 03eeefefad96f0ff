"""The driver oracle against committed outputs of the reference's real
`prepare_inputs_labels_for_multimodal` (tests/golden/driver_*.npz, written by oracle/make_golden.py).
Runs anywhere (CPU, no reference tree)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import driver_oracle
from oracle.make_golden import DRIVER_GEOM, driver_audio, driver_frames, driver_tables, driver_weights

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "driver_*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[7:-4] for p in GOLDEN])
def test_driver_oracle_matches_reference_golden(path):
    z = np.load(path)
    m = json.loads(str(z["meta"]))
    w = driver_weights(m["weight_seed"], m["num_query"])
    sig, dino = driver_tables(m["table_seed"], m["n_frames"])
    n = m["n_frames"]
    kept = list(range(n)) if n <= 224 else [int(n / 224.0 * i) for i in range(224)]   # cambrian_arch.py:908-916
    frames = driver_frames(w, sig, dino)[kept]
    sizes = driver_oracle.segment_sizes_from_boundaries(z["segment_frame_indices"], len(kept))
    audio_frames = None
    if m.get("audio"):
        windows, flags, _, proj = driver_audio(m["audio_seed"], m["n_frames"], m["audio"])
        w.update(proj)
        audio_frames = driver_oracle.audio_frames_from_beats(windows, flags, len(kept))
    got = driver_oracle.compress_video(w, DRIVER_GEOM, frames, sizes, context_token_num=m["num_query"],
                                       query_type=m["query_type"], add_text=m["text"], keep_static=m["add_static"],
                                       input_ids=torch.tensor([m["prompt_ids"]]), max_visual_len=m["max_visual_len"],
                                       audio_frames=audio_frames)
    assert got.shape == z["visual_tokens"].shape
    assert float(np.abs(got.numpy() - z["visual_tokens"]).max()) <= 2e-5


def test_driver_goldens_exist():
    assert len(GOLDEN) >= 4
