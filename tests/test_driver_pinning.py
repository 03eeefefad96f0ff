"""Pin the driver oracle (oracle/driver_oracle.py) and the product's host-side planning against the
reference's REAL `prepare_inputs_labels_for_multimodal`, run unmodified through the fake-model harness
(oracle/harness.py).  Build container only (needs /root/reference); fp32 CPU on both sides."""
import numpy as np
import pytest
import torch

from oracle import driver_oracle, ref_shim
from oracle.synth import QFormerGeometry, make_state_dict

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")

D = 48
GEOM = QFormerGeometry(hidden=64, heads=1, intermediate=96, layers=2, cross_freq=2, d_enc=D, d_out=D, vocab=40,
                       max_pos=16)


def _weights(seed, K):
    rs = np.random.RandomState(seed)
    w = make_state_dict(GEOM, seed, stress=4.0)
    f = lambda *s: (rs.standard_normal(s) * 0.3).astype(np.float32)
    w.update({"query_proj.weight": f(64, D), "query_proj.bias": f(64), "frame_seg": f(D), "image_newline": f(D),
              "query_tokens": f(1, K, 64), "mm_projector.weight": f(D, 24), "mm_projector.bias": f(D),
              "embed_tokens": f(10, D)})
    return w


def _tables(seed, n):
    rs = np.random.RandomState(seed)
    base = rs.standard_normal((n, 144, 8)).astype(np.float32)
    # dino features drift slowly with a few jumps so that segment boundaries are non-trivial
    dino = np.cumsum(rs.standard_normal((n, 1, 16)) * 0.2, axis=0) + rs.standard_normal((1, 144, 16))
    jumps = rs.choice(n, size=max(1, n // 9), replace=False)
    dino[jumps] += rs.standard_normal((len(jumps), 1, 16)) * 3
    return base, dino.astype(np.float32)


@pytest.mark.parametrize("n_frames,query_type,text,add_static,budget", [
    (60, "Avg_pool", True, True, None),
    (60, "learned", False, True, None),
    (41, "Avg_pool", True, False, None),
    (12, "Avg_pool", True, True, None),       # <= 25 frames: every frame is its own segment, nothing compressed
    (60, "Avg_pool", False, True, 3000),      # over budget -> per-chunk truncation
])
def test_driver_oracle_equals_real_reference_function(n_frames, query_type, text, add_static, budget):
    from oracle import harness
    K = 8
    w = _weights(5, K)
    sig, dino = _tables(6, n_frames)
    max_len = 100000 if budget is None else budget
    ref = harness.run_reference_driver(w, GEOM, n_frames, d_llm=D, context_token_num=K, query_type=query_type,
                                       text_input=text, add_static=add_static, tokenizer_model_max_length=max_len,
                                       prompt_ids=[[3, 9, 4, 1]], siglip_table=sig, dino_table=dino)
    sizes = driver_oracle.segment_sizes_from_boundaries(ref["segment_frame_indices"], n_frames)
    assert sum(sizes) == n_frames
    max_visual_len = max_len - 16 - 3      # tokenizer_model_max_length - inference_max_length - text_len (:1501-1505)
    got = driver_oracle.compress_video(w, GEOM, ref["frames"], sizes, context_token_num=K, query_type=query_type,
                                       add_text=text, keep_static=add_static, input_ids=ref["prompt_ids"],
                                       max_visual_len=max_visual_len)
    assert got.shape == ref["visual_tokens"].shape
    assert float((got - ref["visual_tokens"]).abs().max()) <= 2e-5

    # the product's host-side plan produces the same token count / static positions
    from tdc_video_b200.compressor import output_layout, plan_chunks, truncation_keep_index
    plan = plan_chunks(sizes, add_static)
    off, tok, _ = output_layout(plan, 156, K, add_static)
    keep = truncation_keep_index(off, tok, max_visual_len)
    assert (int(tok.sum()) if keep is None else len(keep)) == got.shape[0]


@pytest.mark.parametrize("n_frames", [12, 26, 60, 300])
def test_adapt_segment_oracle_equals_reference(n_frames):
    """oracle.adapt_segment vs the reference method itself (called on the harness object)."""
    from oracle import harness
    arch = harness._load_cambrian_arch()
    rs = np.random.RandomState(n_frames)
    dino = torch.from_numpy(_tables(n_frames, n_frames)[1])
    imgs = torch.zeros(n_frames, 1)

    class Bare(arch.CambrianMetaForCausalLM):
        def get_model(self):
            return None

    feats, split, _, sel_all, seg_all = Bare().adapt_segment(dino, [n_frames], [imgs, imgs], max_num_segments=24)
    sel, seg, cos = driver_oracle.adapt_segment(dino, 24)
    assert torch.equal(sel, sel_all[0]) and torch.equal(seg, seg_all[0])
    assert split == [len(sel)] and feats.shape[0] == len(sel)


@pytest.mark.parametrize("pattern", ["every_second", "sparse"])
def test_audio_branch_equals_real_reference_function(pattern):
    """The audio branch (cambrian_arch.py:1547-1614): BEATs window features -> per-frame tokens -> audio_proj ->
    appended to every frame's KV tokens.  BEATs itself is stubbed; everything after it is the reference's code."""
    from oracle import harness
    K, n_frames = 8, 27
    if pattern == "every_second":
        seconds = n_frames                       # 1 fps, all seconds sampled (main.py:30)
        flags = [1] * n_frames
        vi = torch.ones(n_frames, dtype=torch.int16)   # the audio branch needs explicit video_indices (:928, :1562)
    else:                                       # 61-second clip, 27 sampled seconds with gaps of 1-4 seconds
        rs = np.random.RandomState(3)
        seconds = 61
        pos = np.sort(rs.choice(seconds, size=n_frames, replace=False))
        flags = [1 if i in set(pos.tolist()) else 0 for i in range(seconds)]
        vi = torch.tensor(flags, dtype=torch.int16)
    rs = np.random.RandomState(11)
    n_win = (seconds + 9) // 10
    windows = []
    for w in range(n_win):
        secs_w = min(10, seconds - 10 * w)
        tlen = secs_w * 50 - (7 if w == n_win - 1 else 0)      # the last window ends 7 tokens short (ragged tail)
        windows.append(rs.standard_normal((1, tlen, 768)).astype(np.float32))
    w = _weights(5, K)
    w["audio_proj.weight"] = (rs.standard_normal((D, 768)) * 0.05).astype(np.float32)
    w["audio_proj.bias"] = (rs.standard_normal((D,)) * 0.05).astype(np.float32)
    sig, dino = _tables(6, n_frames)
    ref = harness.run_reference_driver(w, GEOM, n_frames, d_llm=D, context_token_num=K, prompt_ids=[[3, 9, 4, 1]],
                                       siglip_table=sig, dino_table=dino, audio_windows=windows, video_indices=vi,
                                       audio_seconds=seconds)
    sizes = driver_oracle.segment_sizes_from_boundaries(ref["segment_frame_indices"], n_frames)
    audio_frames = driver_oracle.audio_frames_from_beats(windows, flags, n_frames)
    assert audio_frames.shape == (n_frames, 50, 768)
    got = driver_oracle.compress_video(w, GEOM, ref["frames"], sizes, context_token_num=K, input_ids=ref["prompt_ids"],
                                       audio_frames=audio_frames, max_visual_len=100000 - 16 - 3)
    assert got.shape == ref["visual_tokens"].shape
    assert float((got - ref["visual_tokens"]).abs().max()) <= 2e-5
