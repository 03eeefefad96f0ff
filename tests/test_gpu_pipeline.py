"""The composed TDC stage (tdc_video_b200.pipeline.tdc_video_stage: adapt_segment -> mm_projector + newline tokens
-> chunked Q-Former compression -> assembly / budget truncation), all on the GPU, against the committed outputs of
the reference's REAL `prepare_inputs_labels_for_multimodal` (tests/golden/driver_*.npz, generated through
oracle/harness.py).  Tolerance as everywhere: cosine >= 0.999, max-normalised error <= 2e-2."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import qformer_oracle as oracle
from oracle.make_golden import DRIVER_D, DRIVER_GEOM, driver_audio, driver_frames, driver_tables, driver_weights

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "driver_*.npz")))


def _compressor(m, w):
    from tdc_video_b200.compressor import TDCCompressor
    from tdc_video_b200.qformer import QFormerConfig
    g = DRIVER_GEOM
    cfg = QFormerConfig(vocab_size=g.vocab, hidden_size=g.hidden, num_hidden_layers=g.layers,
                        num_attention_heads=g.heads, intermediate_size=g.intermediate,
                        max_position_embeddings=g.max_pos, layer_norm_eps=g.ln_eps,
                        cross_attention_freq=g.cross_freq, encoder_width=DRIVER_D, query_length=m["num_query"])
    comp = TDCCompressor(DRIVER_D, context_token_num=m["num_query"], query_type=m["query_type"], text_input=m["text"],
                         add_static=m["add_static"], audio_input=bool(m.get("audio")), qformer_config=cfg)
    sd = {}
    for k, v in w.items():
        if k.startswith(("embeddings.", "encoder.")):
            sd["Qformer.bert." + k] = torch.from_numpy(v)
        elif k.split(".")[0] in ("vision_proj", "query_proj", "frame_seg", "query_tokens", "audio_proj"):
            sd[k] = torch.from_numpy(v)
    missing, unexpected = comp.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in k or k.startswith("Qformer.cls.") for k in missing), missing
    return comp.cuda().eval()


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[7:-4] for p in GOLDEN])
def test_stage_matches_the_reference_driver(path):
    from tdc_video_b200 import linear
    from tdc_video_b200.pipeline import append_newline_tokens, tdc_video_stage
    z = np.load(path)
    m = json.loads(str(z["meta"]))
    w = driver_weights(m["weight_seed"], m["num_query"])
    sig, dino = driver_tables(m["table_seed"], m["n_frames"])
    n = m["n_frames"]
    audio_kw = {}
    if m.get("audio"):   # stub BEATs windows + sampled-second flags: the audio branch (cambrian_arch.py:1547-1614)
        windows, flags, _, proj = driver_audio(m["audio_seed"], n, m["audio"])
        w.update(proj)
        audio_kw = dict(audio_windows=[torch.from_numpy(a).cuda() for a in windows], sample_indices=flags)
    # frame tokens: mm_projector on the concatenated tower features (tcgen05 GEMM), then the newline column
    feats = torch.from_numpy(np.concatenate([sig, dino], -1)).cuda()
    proj = linear(feats, torch.from_numpy(w["mm_projector.weight"]).cuda(), torch.from_numpy(w["mm_projector.bias"]).cuda(),
                  out_dtype=torch.float32)
    frames = append_newline_tokens(proj.view(n, 144, DRIVER_D), torch.from_numpy(w["image_newline"]).cuda())
    ref_frames = driver_frames(w, sig, dino)
    assert frames.shape == ref_frames.shape == (n, 156, DRIVER_D)
    fm = oracle.parity_metrics(frames.float().cpu(), ref_frames)
    assert fm["min_cos"] >= 0.999 and fm["max_abs_over_max_ref"] <= 2e-2 and fm["max_tok_rel_l2"] <= 2e-2, fm
    assert torch.equal(frames.view(n, 12, 13, DRIVER_D)[:, :, 12].cpu(),
                       torch.from_numpy(w["image_newline"]).expand(n, 12, DRIVER_D))

    comp = _compressor(m, w)
    ids = torch.tensor([m["prompt_ids"]], device="cuda")
    seq, selected, bounds = tdc_video_stage(comp, frames, torch.from_numpy(dino).cuda(), input_ids=ids,
                                            max_visual_len=m["max_visual_len"], return_segments=True, **audio_kw)
    torch.cuda.synchronize()
    # segmentation: the same boundaries as the reference's adapt_segment chose
    kept = list(range(n)) if n <= 224 else [int(n / 224.0 * i) for i in range(224)]   # cambrian_arch.py:908-916
    assert selected.tolist() == kept
    assert bounds.cpu().tolist() == z["segment_frame_indices"].tolist()
    ref = torch.from_numpy(z["visual_tokens"])
    assert tuple(seq.shape) == tuple(ref.shape)
    mt = oracle.parity_metrics(seq.float().cpu(), ref)
    print(os.path.basename(path), mt)
    assert mt["min_cos"] >= 0.999 and mt["max_abs_over_max_ref"] <= 2e-2 and mt["max_tok_rel_l2"] <= 2e-2, mt


def test_append_newline_tokens_layout():
    from tdc_video_b200.pipeline import append_newline_tokens
    x = torch.arange(2 * 9 * 4, dtype=torch.float32, device="cuda").view(2, 9, 4)
    nl = torch.tensor([-1.0, -2.0, -3.0, -4.0], device="cuda")
    y = append_newline_tokens(x, nl)
    ref = torch.cat([x.view(2, 3, 3, 4), nl.view(1, 1, 1, 4).expand(2, 3, 1, 4)], dim=2).flatten(1, 2)
    assert torch.equal(y, ref)
    with pytest.raises(ValueError):
        append_newline_tokens(x[:, :8], nl)
