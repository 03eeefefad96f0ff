"""TEST INFRASTRUCTURE — generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python -m oracle.make_golden

For every case the reference `BertModel` (tdc/Qformer.py, imported through oracle/ref_shim.py)
is constructed, loaded with the seeded synthetic state dict of oracle/synth.py and called
exactly like the TDC call site (tdc/cambrian_arch.py:1653-1667), in fp32 on the CPU.
Only geometry + seeds + the reference outputs are stored; weights and inputs are regenerated
from the seeds by the tests.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

from oracle import ref_shim
from oracle.synth import QFormerGeometry, make_inputs, make_state_dict

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SMALL = QFormerGeometry(hidden=128, heads=2, intermediate=256, layers=4, cross_freq=2, d_enc=96, d_out=160,
                        vocab=64, max_pos=32)
FREQ1 = QFormerGeometry(hidden=64, heads=1, intermediate=128, layers=2, cross_freq=1, d_enc=40, d_out=72,
                        vocab=32, max_pos=16)
FULL_LITERAL = QFormerGeometry(d_enc=1152, d_out=3072)   # BASELINE config 1/2: SigLIP-width tokens, Llama-3.2-3B out
FULL_QWEN = QFormerGeometry(d_enc=3584, d_out=3584)      # reference order, Qwen2-7B widths

# name -> (geometry, seed, stress, rows, kv_tokens, K, T, audio_tokens, kv_len)
CASES = {
    "small_notext": (SMALL, 11, 0.0, 5, 23, 8, 0, 0, None),
    "small_text": (SMALL, 12, 0.0, 3, 37, 8, 5, 7, None),
    "small_stress_k16": (SMALL, 13, 8.0, 4, 50, 16, 3, 0, None),
    "small_kvlen": (SMALL, 14, 8.0, 4, 40, 8, 0, 0, [40, 17, 1, 33]),
    "small_k20_ragged_queries": (SMALL, 15, 0.0, 2, 19, 20, 4, 0, None),
    "freq1_speech_style": (FREQ1, 16, 8.0, 3, 30, 1, 0, 0, None),  # audio_encoder.py-style: 1 query, cross every layer
    "full_literal_l194": (FULL_LITERAL, 21, 0.0, 2, 194, 16, 0, 50, None),
    "full_qwen_l206_text": (FULL_QWEN, 22, 2.0, 2, 206, 16, 6, 50, None),
}


def run_reference(geom, sd_np, inputs, num_query, kv_len=None):
    model = ref_shim.build_reference_bert(geom, num_query)
    own = model.state_dict()
    load = {k: torch.from_numpy(v) for k, v in sd_np.items() if not k.startswith("vision_proj")}
    missing = [k for k in own if k not in load and not k.endswith("position_ids")]
    # text FFN / embeddings absent from a text-less state dict keep their (unused) init values
    model.load_state_dict(load, strict=False)
    assert all(("intermediate.dense" in k or "output.dense" in k or "output.LayerNorm" in k or "embeddings." in k)
               for k in missing), missing
    q = torch.from_numpy(inputs["query_embeds"])
    enc = torch.from_numpy(inputs["enc"])
    ids = torch.from_numpy(inputs["input_ids"]) if inputs["input_ids"] is not None else None
    if kv_len is None:
        atts = torch.ones(enc.shape[:-1], dtype=torch.long)   # cambrian_arch.py:1648-1650
    else:
        atts = (torch.arange(enc.shape[1])[None, :] < torch.tensor(kv_len)[:, None]).long()
    with torch.no_grad():
        out = model(input_ids=ids, query_embeds=q, encoder_hidden_states=enc, encoder_attention_mask=atts,
                    use_cache=False, return_dict=True)
        hidden = out.last_hidden_state
        vp = torch.nn.Linear(geom.hidden, geom.d_out)
        vp.weight.data = torch.from_numpy(sd_np["vision_proj.weight"])
        vp.bias.data = torch.from_numpy(sd_np["vision_proj.bias"])
        comp = F.normalize(vp(hidden[:, :num_query]), dim=-1)   # cambrian_arch.py:1664-1667
    return hidden.numpy(), comp.numpy()


# ---- driver goldens: the REAL prepare_inputs_labels_for_multimodal through oracle/harness.py ------------
DRIVER_D = 48
DRIVER_GEOM = QFormerGeometry(hidden=64, heads=1, intermediate=96, layers=2, cross_freq=2, d_enc=DRIVER_D,
                              d_out=DRIVER_D, vocab=40, max_pos=16)
# name -> (n_frames, K, query_type, text, add_static, tokenizer_model_max_length, audio pattern or None)
DRIVER_CASES = {
    "avgpool_text_33f": (33, 8, "Avg_pool", True, True, 100000, None),
    "learned_notext_budget_30f": (30, 8, "learned", False, True, 2600, None),
    "avgpool_text_audio_sparse_27f": (27, 8, "Avg_pool", True, True, 100000, "sparse"),
    "avgpool_notext_nostatic_230f": (230, 8, "Avg_pool", False, False, 100000, None),   # > 224 frames: subsample
}


def driver_audio(seed, n_frames, pattern):
    """Stub BEATs output for the audio branch (cambrian_arch.py:1547-1598): per 10-second window a [1, t, 768]
    feature block (the last one 7 tokens short), the 0/1 "second was sampled" flags and `audio_proj` weights.
    "sparse": a 61-second clip of which n_frames seconds were sampled, with gaps of 1-4 seconds."""
    rs = np.random.RandomState(seed)
    if pattern == "every_second":
        seconds, flags = n_frames, [1] * n_frames
    else:
        seconds = 61
        pos = set(np.sort(rs.choice(seconds, size=n_frames, replace=False)).tolist())
        flags = [1 if i in pos else 0 for i in range(seconds)]
    n_win = (seconds + 9) // 10
    windows = []
    for w in range(n_win):
        secs_w = min(10, seconds - 10 * w)
        windows.append(rs.standard_normal((1, secs_w * 50 - (7 if w == n_win - 1 else 0), 768)).astype(np.float32))
    proj = {"audio_proj.weight": (rs.standard_normal((DRIVER_D, 768)) * 0.05).astype(np.float32),
            "audio_proj.bias": (rs.standard_normal((DRIVER_D,)) * 0.05).astype(np.float32)}
    return windows, flags, seconds, proj


def driver_weights(seed, K):
    """Q-Former + the sibling tensors the driver touches, all from one numpy RandomState."""
    rs = np.random.RandomState(seed)
    w = make_state_dict(DRIVER_GEOM, seed, stress=4.0)
    f = lambda *s: (rs.standard_normal(s) * 0.3).astype(np.float32)
    w.update({"query_proj.weight": f(64, DRIVER_D), "query_proj.bias": f(64), "frame_seg": f(DRIVER_D),
              "image_newline": f(DRIVER_D), "query_tokens": f(1, K, 64), "mm_projector.weight": f(DRIVER_D, 24),
              "mm_projector.bias": f(DRIVER_D), "embed_tokens": f(10, DRIVER_D)})
    return w


def driver_tables(seed, n):
    """Stub tower outputs: SigLIP-like [n,144,8] noise, DINO-like [n,144,16] slow drift with jumps."""
    rs = np.random.RandomState(seed)
    base = rs.standard_normal((n, 144, 8)).astype(np.float32)
    dino = np.cumsum(rs.standard_normal((n, 1, 16)) * 0.2, axis=0) + rs.standard_normal((1, 144, 16))
    jumps = rs.choice(n, size=max(1, n // 9), replace=False)
    dino[jumps] += rs.standard_normal((len(jumps), 1, 16)) * 3
    return base, dino.astype(np.float32)


def driver_weights_mlp(seed, K):
    """driver_weights with the shipped GELU-MLP projector (`mm_projector.0.*`, `mm_projector.2.*`) instead of the
    single Linear: the weights of the upstream ("frames") entry."""
    w = driver_weights(seed, K)
    del w["mm_projector.weight"], w["mm_projector.bias"]
    rs = np.random.RandomState(seed + 1000)
    f = lambda *s: (rs.standard_normal(s) * 0.3).astype(np.float32)
    w.update({"mm_projector.0.weight": f(DRIVER_D, 24), "mm_projector.0.bias": f(DRIVER_D),
              "mm_projector.2.weight": f(DRIVER_D, DRIVER_D) * 0.5, "mm_projector.2.bias": f(DRIVER_D)})
    return w


def driver_frames(w, sig, dino):
    """What cambrian_arch.py:1146-1299 turns the tower features into for square frames:
    mm_projector(cat(siglip, dino)) on the 12x12 grid + one image_newline per grid row -> [n,156,d]."""
    feats = torch.from_numpy(np.concatenate([sig, dino], -1))
    if "mm_projector.0.weight" in w:
        tt = lambda k: torch.from_numpy(w[k])
        proj = F.linear(F.gelu(F.linear(feats, tt("mm_projector.0.weight"), tt("mm_projector.0.bias"))),
                        tt("mm_projector.2.weight"), tt("mm_projector.2.bias"))
    else:
        proj = F.linear(feats, torch.from_numpy(w["mm_projector.weight"]), torch.from_numpy(w["mm_projector.bias"]))
    n = proj.shape[0]
    proj = proj.view(n, 12, 12, -1)
    nl = torch.from_numpy(w["image_newline"]).view(1, 1, 1, -1).expand(n, 12, 1, -1)
    return torch.cat([proj, nl], dim=2).flatten(1, 2)


def make_driver_goldens(only=None):
    from oracle import harness
    for name, (n, K, qt, text, static, max_len, audio) in DRIVER_CASES.items():
        if only and name not in only:
            continue
        w = driver_weights(40 + n, K)
        sig, dino = driver_tables(50 + n, n)
        akw = {}
        if audio:
            windows, flags, seconds, proj = driver_audio(60 + n, n, audio)
            w.update(proj)
            akw = dict(audio_windows=windows, video_indices=torch.tensor(flags, dtype=torch.int16), audio_seconds=seconds)
        elif n > 224:   # the > 224-frame branch indexes video_indices[i] (cambrian_arch.py:918-921): it must be a tensor
            akw = dict(video_indices=torch.ones(n, dtype=torch.int16))
        ref = harness.run_reference_driver(w, DRIVER_GEOM, n, d_llm=DRIVER_D, context_token_num=K, query_type=qt,
                                           text_input=text, add_static=static, tokenizer_model_max_length=max_len,
                                           prompt_ids=[[3, 9, 4, 1]], siglip_table=sig, dino_table=dino, **akw)
        kept = list(range(n)) if n <= 224 else [int(n / 224.0 * i) for i in range(224)]   # cambrian_arch.py:908-916
        assert torch.allclose(ref["frames"], driver_frames(w, sig, dino)[kept], atol=1e-6)
        meta = dict(n_frames=n, num_query=K, query_type=qt, text=text, add_static=static,
                    tokenizer_model_max_length=max_len, max_visual_len=max_len - 16 - 3, prompt_ids=[3, 9, 4, 1],
                    weight_seed=40 + n, table_seed=50 + n, audio=audio, audio_seed=60 + n,
                    generator="oracle/make_golden.py: reference prepare_inputs_labels_for_multimodal via oracle/harness.py")
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"driver_{name}.npz"), meta=json.dumps(meta),
                            segment_frame_indices=ref["segment_frame_indices"].numpy().astype(np.int64),
                            visual_tokens=ref["visual_tokens"].numpy().astype(np.float32))
        print(f"driver {name}: tokens {tuple(ref['visual_tokens'].shape)}")


# ---- "towers" goldens: the same real driver, with the shipped GELU-MLP projector, for the upstream entry
# (tdc_compress_frames: tower features in, token sequence out).  name -> as DRIVER_CASES
TOWER_CASES = {
    "avgpool_text_audio_29f": (29, 8, "Avg_pool", True, True, 100000, "sparse"),
    "learned_notext_31f": (31, 8, "learned", False, True, 100000, None),
    "avgpool_notext_40f": (40, 8, "Avg_pool", False, True, 100000, None),
}


def make_tower_goldens(only=None):
    from oracle import harness
    for name, (n, K, qt, text, static, max_len, audio) in TOWER_CASES.items():
        if only and ("towers_" + name) not in only and name not in only:
            continue
        w = driver_weights_mlp(140 + n, K)
        sig, dino = driver_tables(150 + n, n)
        akw = {}
        if audio:
            windows, flags, seconds, proj = driver_audio(160 + n, n, audio)
            w.update(proj)
            akw = dict(audio_windows=windows, video_indices=torch.tensor(flags, dtype=torch.int16), audio_seconds=seconds)
        ref = harness.run_reference_driver(w, DRIVER_GEOM, n, d_llm=DRIVER_D, context_token_num=K, query_type=qt,
                                           text_input=text, add_static=static, tokenizer_model_max_length=max_len,
                                           prompt_ids=[[3, 9, 4, 1]], siglip_table=sig, dino_table=dino, **akw)
        assert torch.allclose(ref["frames"], driver_frames(w, sig, dino), atol=1e-5)
        meta = dict(n_frames=n, num_query=K, query_type=qt, text=text, add_static=static,
                    tokenizer_model_max_length=max_len, max_visual_len=max_len - 16 - 3, prompt_ids=[3, 9, 4, 1],
                    weight_seed=140 + n, table_seed=150 + n, audio=audio, audio_seed=160 + n, projector="mlp2x_gelu",
                    generator="oracle/make_golden.py: reference prepare_inputs_labels_for_multimodal via "
                              "oracle/harness.py with the nn.Sequential(Linear, GELU, Linear) mm_projector")
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"towers_{name}.npz"), meta=json.dumps(meta),
                            segment_frame_indices=ref["segment_frame_indices"].numpy().astype(np.int64),
                            visual_tokens=ref["visual_tokens"].numpy().astype(np.float32))
        print(f"towers {name}: tokens {tuple(ref['visual_tokens'].shape)}")


# ---- SVA goldens: reference VisionTokenSampler + mm_projector_aux + real window rearrangement ----------------
# name -> (hidden, tower_dims, window_sides, layers, query_side, image sizes, seed, stress)
SVA_CASES = {
    "h128_masks": (128, (96, 64), (2, 2), 2, 4, [(640, 360), (384, 384), (300, 500)], 31, 3.0),
    "h1024_full": (1024, (1152, 1536), (2, 2), 3, 12, [(1280, 720), (384, 384)], 32, 2.0),   # shipped geometry
}


def run_reference_sva(sd, hidden, tower_dims, sides, layers, Q, sizes, tower):
    import importlib.util
    name = "_tdc_reference_vision_sampler"
    if name not in sys.modules:
        spec = importlib.util.spec_from_file_location(name, os.path.join(ref_shim.REFERENCE_ROOT, "tdc", "vision_sampler.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    vs = sys.modules[name]
    from oracle import harness
    arch = harness._load_cambrian_arch()

    class Bare(arch.CambrianMetaForCausalLM):
        def get_model(self):
            return None

    sampler = vs.VisionTokenSampler(hidden, hidden, [hidden] * 2, list(sides), hidden, layers).eval()
    sampler.load_state_dict({k[len("vision_sampler_0."):]: torch.from_numpy(v) for k, v in sd.items()
                             if k.startswith("vision_sampler_0.")}, strict=True)
    bs, nq = len(sizes), Q * Q
    with torch.no_grad():
        feats = []
        for t, c in enumerate(tower_dims):
            m = torch.nn.Sequential(torch.nn.Linear(c, hidden), torch.nn.GELU(), torch.nn.Linear(hidden, hidden),
                                    torch.nn.LayerNorm(hidden)).eval()
            m.load_state_dict({k[len(f"mm_projector_aux_{t}."):]: torch.from_numpy(v) for k, v in sd.items()
                               if k.startswith(f"mm_projector_aux_{t}.")}, strict=True)
            feats.append(m(tower[t]))
        lat, masks = Bare().rearrange_vision_tower_features_inference(feats, Q, sizes)
        ctx = feats[0].mean(1).view(bs, 1, 1, -1).expand(-1, nq, 1, -1).flatten(0, 1)
        qry = torch.from_numpy(sd["vision_query"])[0].view(1, 1, 1, -1).expand(bs, nq, -1, -1).flatten(0, 1)
        return sampler(qry, ctx, *lat, *masks).view(bs, nq, hidden)


def sva_inputs(tower_dims, sides, Q, bs, seed):
    rs = np.random.RandomState(seed + 5)
    return [torch.from_numpy(rs.standard_normal((bs, (Q * s) ** 2, c)).astype(np.float32)) for s, c in zip(sides, tower_dims)]


def make_sva_goldens(only=None):
    from oracle.synth import make_sva_state_dict
    for name, (hidden, dims, sides, layers, Q, sizes, seed, stress) in SVA_CASES.items():
        if only and name not in only:
            continue
        sd = make_sva_state_dict(hidden, dims, sides, layers, seed, stress)
        tower = sva_inputs(dims, sides, Q, len(sizes), seed)
        out = run_reference_sva(sd, hidden, dims, sides, layers, Q, sizes, tower)
        meta = dict(hidden=hidden, tower_dims=dims, window_sides=sides, layers=layers, query_side=Q, image_sizes=sizes,
                    seed=seed, stress=stress, heads=16,
                    generator="oracle/make_golden.py: reference VisionTokenSampler / mm_projector_aux / "
                              "rearrange_vision_tower_features_inference (fp32 CPU)")
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"sva_{name}.npz"), meta=json.dumps(meta),
                            out=out.numpy().astype(np.float32))
        print(f"sva {name}: {tuple(out.shape)}")


def main(only=None):
    assert ref_shim.reference_available(), "needs /root/reference"
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.manual_seed(0)
    for name, (geom, seed, stress, rows, L, K, T, audio, kv_len) in CASES.items():
        if only and name not in only:
            continue
        sd = make_state_dict(geom, seed, stress=stress, with_text=T > 0)
        inputs = make_inputs(geom, seed, rows, L, K, T, audio_tokens=audio)
        hidden, comp = run_reference(geom, sd, inputs, K, kv_len)
        meta = dict(geometry=geom.to_dict(), seed=seed, stress=stress, rows=rows, kv_tokens=L, num_query=K,
                    num_text=T, audio_tokens=audio, kv_len=kv_len,
                    generator="oracle/make_golden.py on reference tdc/Qformer.py (fp32 CPU)",
                    torch=torch.__version__)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"qformer_{name}.npz"), meta=json.dumps(meta),
                            hidden=hidden.astype(np.float32), compressed=comp.astype(np.float32))
        print(f"{name}: hidden {hidden.shape} compressed {comp.shape}")
    make_driver_goldens(only)
    make_tower_goldens(only)
    make_sva_goldens(only)


if __name__ == "__main__":
    main(sys.argv[1:])
