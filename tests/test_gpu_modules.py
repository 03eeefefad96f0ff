"""The reference-shaped modules (TDCBertModel / TDCQFormer / TDCCompressor) on the GPU against the
oracle: same state_dict names, same forward keywords, same token sequence."""
import numpy as np
import pytest
import torch

from oracle import driver_oracle
from oracle import qformer_oracle as oracle

pytestmark = pytest.mark.gpu


def _metrics_ok(test, ref, what):
    m = oracle.parity_metrics(test.float().cpu(), ref)
    print(what, m)
    assert m["min_cos"] >= 0.999 and m["max_abs_over_max_ref"] <= 2e-2 and m["max_tok_rel_l2"] <= 2e-2, (what, m)


def _small_cfg(**kw):
    from tdc_video_b200.qformer import QFormerConfig
    base = dict(vocab_size=40, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                max_position_embeddings=16, encoder_width=64, query_length=16)
    base.update(kw)
    return QFormerConfig(**base)


def _geom_of(cfg, d_out):
    from oracle.synth import QFormerGeometry
    return QFormerGeometry(hidden=cfg.hidden_size, heads=cfg.num_attention_heads, intermediate=cfg.intermediate_size,
                           layers=cfg.num_hidden_layers, cross_freq=cfg.cross_attention_freq, d_enc=cfg.encoder_width,
                           d_out=d_out, vocab=cfg.vocab_size, max_pos=cfg.max_position_embeddings,
                           ln_eps=cfg.layer_norm_eps)


def _randomize(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith("LayerNorm.weight"):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
            elif p.dim() == 1:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(torch.randn(p.shape, generator=g) * (0.5 / p.shape[-1] ** 0.5))


def test_bert_mirror_forward_signature_and_parity():
    from tdc_video_b200.qformer import TDCQFormer
    cfg = _small_cfg()
    model = TDCQFormer(cfg)
    _randomize(model, 1)
    model = model.cuda().eval()
    sd = {k[len("bert."):]: v.detach().cpu() for k, v in model.state_dict().items() if k.startswith("bert.")}
    B, K, T, L = 5, 16, 4, 21
    q = torch.randn(B, K, 128, device="cuda")
    enc = torch.randn(B, L, 64, device="cuda")
    ids = torch.randint(0, 40, (B, T), device="cuda")
    atts = torch.ones(B, L, dtype=torch.long, device="cuda")
    out = model.bert(input_ids=ids, query_embeds=q, encoder_hidden_states=enc, encoder_attention_mask=atts,
                     use_cache=False, return_dict=True)          # the call of cambrian_arch.py:1653-1662
    assert out.last_hidden_state.shape == (B, K + T, 128) and out.last_hidden_state.dtype == q.dtype
    ref = oracle.qformer_forward(sd, _geom_of(cfg, 0), q.cpu(), enc.cpu(), ids.cpu())
    _metrics_ok(out.last_hidden_state, ref, "bert mirror")
    # weights changed in place + invalidate -> new result
    with torch.no_grad():
        model.bert.encoder.layer[0].output_query.dense.bias.add_(1.0)
    model.bert.invalidate_engine()
    out2 = model.bert(query_embeds=q, encoder_hidden_states=enc, input_ids=ids).last_hidden_state
    assert not torch.allclose(out2, out.last_hidden_state)
    with pytest.raises(NotImplementedError):
        model.bert(query_embeds=q, encoder_hidden_states=enc, attention_mask=torch.ones(B, K, device="cuda"))
    with pytest.raises(RuntimeError, match="inference-only"):
        model.train().bert(query_embeds=q, encoder_hidden_states=enc)


def test_state_dict_names_match_reference_layout():
    """Key names/shapes of SURVEY.md appendix A, incl. the unused LM head, so reference checkpoints load strictly."""
    from tdc_video_b200.qformer import TDCQFormer
    cfg = _small_cfg()
    keys = dict(TDCQFormer(cfg).state_dict())
    for k in ["bert.embeddings.word_embeddings.weight", "bert.embeddings.position_embeddings.weight",
              "bert.embeddings.LayerNorm.weight", "bert.embeddings.position_ids",
              "bert.encoder.layer.0.attention.self.query.weight", "bert.encoder.layer.0.attention.output.LayerNorm.bias",
              "bert.encoder.layer.0.crossattention.self.key.weight", "bert.encoder.layer.0.crossattention.output.dense.bias",
              "bert.encoder.layer.1.intermediate.dense.weight", "bert.encoder.layer.1.output.LayerNorm.weight",
              "bert.encoder.layer.1.intermediate_query.dense.bias", "bert.encoder.layer.1.output_query.dense.weight",
              "cls.predictions.bias", "cls.predictions.transform.dense.weight", "cls.predictions.decoder.weight"]:
        assert k in keys, k
    assert "bert.encoder.layer.1.crossattention.self.key.weight" not in keys     # cross_attention_freq = 2
    assert keys["bert.encoder.layer.0.crossattention.self.key.weight"].shape == (128, 64)


@pytest.mark.parametrize("query_type,text,audio,keep_static", [("Avg_pool", True, True, True),
                                                                 ("Avg_pool", False, False, True),
                                                                 ("learned", True, False, False)])
def test_compressor_matches_reference_loop(query_type, text, audio, keep_static):
    from tdc_video_b200.compressor import TDCCompressor
    cfg = _small_cfg()
    d = 64
    comp = TDCCompressor(d, context_token_num=16, query_type=query_type, text_input=text, add_static=keep_static,
                         audio_input=audio, qformer_config=cfg)
    _randomize(comp, 2)
    comp = comp.cuda().eval()
    sizes = [3, 1, 12, 8, 2]
    n, Lv, La = sum(sizes), 20, 6
    frames = torch.randn(n, Lv, d, device="cuda")
    audio_frames = torch.randn(n, La, 768, device="cuda") if audio else None
    ids = torch.randint(0, 40, (1, 5), device="cuda") if text else None
    weights = {k[len("Qformer.bert."):]: v.detach().cpu() for k, v in comp.state_dict().items()
               if k.startswith("Qformer.bert.")}
    for k, v in comp.state_dict().items():
        if not k.startswith("Qformer."):
            weights[k] = v.detach().cpu()
    for budget in (None, 150):
        seq = comp.compress_video(frames, sizes, input_ids=ids, audio_frames=audio_frames, max_visual_len=budget)
        ref = driver_oracle.compress_video(weights, _geom_of(cfg, d), frames.cpu(), sizes, context_token_num=16,
                                           query_type=query_type, add_text=text, keep_static=keep_static,
                                           input_ids=None if ids is None else ids.cpu(),
                                           audio_frames=None if audio_frames is None else audio_frames.cpu(),
                                           max_visual_len=budget)
        assert seq.shape == ref.shape
        _metrics_ok(seq, ref, f"compressor {query_type} text={text} audio={audio} budget={budget}")


@pytest.mark.parametrize("query_type", ["Avg_pool", "learned"])
def test_compress_videos_batches_many_videos_bit_exactly(query_type):
    """The eval-style workload (many concurrent clips, BASELINE config 5): rows of all videos with the same KV and
    prompt length go through one tdc_compress call; rows are independent, so every video's sequence must equal
    the per-video call bit for bit.  Mixed: audio / silent videos, two prompt lengths, a one-frame video (no rows)."""
    from tdc_video_b200.compressor import TDCCompressor
    cfg = _small_cfg()
    d = 64
    comp = TDCCompressor(d, context_token_num=16, query_type=query_type, text_input=True, add_static=True,
                         audio_input=True, qformer_config=cfg)
    _randomize(comp, 4)
    comp = comp.cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(9)
    spec = [([3, 1, 9, 2], 5, True, None), ([4, 4], 5, False, 120), ([1], 5, False, None), ([2, 11], 3, True, None),
            ([6, 1, 1], 5, True, 200), ([5], 3, False, None)]
    videos = []
    for sizes, T, audio, budget in spec:
        n = sum(sizes)
        videos.append(dict(visual_emb_frame=torch.randn(n, 20, d, device="cuda", generator=g), segment_sizes=sizes,
                           input_ids=torch.randint(0, 40, (1, T), device="cuda", generator=g),
                           audio_frames=torch.randn(n, 6, 768, device="cuda", generator=g) if audio else None,
                           max_visual_len=budget))
    eng = comp._engine()
    calls0 = eng.launch_count()
    batched = comp.compress_videos(videos)
    launches_batched = eng.launch_count() - calls0
    calls0 = eng.launch_count()
    single = [comp.compress_video(v["visual_emb_frame"], v["segment_sizes"], input_ids=v["input_ids"],
                                  audio_frames=v["audio_frames"], max_visual_len=v["max_visual_len"]) for v in videos]
    launches_single = eng.launch_count() - calls0
    torch.cuda.synchronize()
    assert len(batched) == len(videos)
    for i, (a, b) in enumerate(zip(batched, single)):
        assert a.shape == b.shape and torch.equal(a, b), i
    # 4 groups (audio x prompt length) instead of 5 calls with rows
    assert launches_batched < launches_single


def test_gelu_mlp_projector_and_linear():
    """mm_projector Linear-GELU(erf)-Linear (cambrian_arch.py:65-69) on the tcgen05 GEMM."""
    from tdc_video_b200 import linear
    torch.manual_seed(0)
    x = torch.randn(1000, 1024, device="cuda")
    w0, b0 = torch.randn(3584, 1024, device="cuda") * 0.03, torch.randn(3584, device="cuda") * 0.1
    w1, b1 = torch.randn(3584, 3584, device="cuda") * 0.02, torch.randn(3584, device="cuda") * 0.1
    mid = linear(x, w0, b0, gelu=True)
    y = linear(mid, w1, b1, out_dtype=torch.float32)
    ref = oracle.gelu_mlp(w0.cpu(), b0.cpu(), w1.cpu(), b1.cpu(), x.cpu())
    _metrics_ok(y, ref, "gelu mlp")
    for cg in (1, 2):
        y1 = linear(x, w0, b0, out_dtype=torch.float32, cta_group=cg)
        _metrics_ok(y1, torch.nn.functional.linear(x.cpu(), w0.cpu(), b0.cpu()), f"linear cg{cg}")


def test_avg_pool_tokens_matches_adaptive_avg_pool1d():
    from tdc_video_b200 import avg_pool_tokens
    x = torch.randn(7, 156, 64, device="cuda")
    got = avg_pool_tokens(x, 16).float().cpu()
    ref = oracle.avg_pool_queries(x.cpu(), 16)
    assert torch.allclose(got, ref, atol=1e-2, rtol=1e-2)


@pytest.mark.parametrize("n_frames,dtype", [(12, torch.bfloat16), (60, torch.bfloat16), (224, torch.float16),
                                            (300, torch.float32)])
def test_adapt_segment_matches_oracle(n_frames, dtype):
    """tdc_segment_boundaries vs the restated adapt_segment (cambrian_arch.py:783-861): identical boundary
    indices, cosine similarities within 1e-3 (features drift slowly with a few well-separated jumps)."""
    from tdc_video_b200.segment import adapt_segment, segment_sizes
    rs = np.random.RandomState(n_frames)
    dino = np.cumsum(rs.standard_normal((n_frames, 1, 64)) * 0.05, axis=0) + rs.standard_normal((1, 576, 64))
    if n_frames > 25:   # exactly 24 scene cuts: their similarities sit far below the slow drift of the rest
        jumps = np.sort(rs.choice(np.arange(1, n_frames), size=24, replace=False))
        for j, scale in zip(jumps, np.linspace(1.5, 4.0, 24)):
            dino[j:] += rs.standard_normal((1, 1, 64)) * scale
    feats = torch.from_numpy(dino.astype(np.float32)).to(dtype)
    sel, seg, cos = adapt_segment(feats.cuda(), 24)
    sel_o, seg_o, cos_o = driver_oracle.adapt_segment(feats.float(), 24)
    assert torch.equal(sel, sel_o)
    if cos_o is None:
        assert cos is None and torch.equal(seg.cpu(), seg_o)
    else:
        # the library rounds the similarities to the feature dtype, as the reference's F.cosine_similarity on
        # bf16 / fp16 features does (half a bf16 ulp below 1.0 = 2e-3)
        tol = 1e-3 if dtype == torch.float32 else 3e-3
        assert torch.allclose(cos.cpu(), cos_o, atol=tol)
        if dtype != torch.float32:
            assert torch.equal(cos.cpu(), cos.cpu().to(dtype).float())
        # the 24 chosen similarities must be separated from the rest by more than the tolerance for the
        # index comparison to be meaningful
        srt = torch.sort(cos_o).values
        if len(srt) > 24:
            assert float(srt[24] - srt[23]) > 2 * tol
        assert torch.equal(seg.cpu(), seg_o)
    assert sum(segment_sizes(seg, len(sel))) == len(sel)


def test_compress_is_cuda_graph_capturable():
    """The C ABI promises stream-ordered, allocation-free, sync-free calls: capture one tdc_compress in a CUDA
    graph, replay it on new input data in the same buffers, and compare with the eager call."""
    from oracle.synth import QFormerGeometry, make_state_dict
    from tdc_video_b200 import QFormerEngine
    geom = QFormerGeometry(hidden=128, heads=2, intermediate=256, layers=2, cross_freq=2, d_enc=64, d_out=96, vocab=0)
    eng = QFormerEngine(hidden=128, heads=2, intermediate=256, layers=2, cross_freq=2, d_enc=64, d_out=96)
    eng.load_weights(make_state_dict(geom, 4, with_text=False))
    q = torch.randn(9, 16, 128, device="cuda")
    enc = torch.randn(9, 40, 64, device="cuda").bfloat16()
    eager = eng.compress(q, enc)          # also sizes the workspace outside the capture
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            out = eng.compress(q, enc)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager)
    enc.copy_(torch.randn(9, 40, 64, device="cuda").bfloat16())   # new data, same addresses
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eng.compress(q, enc))
    assert not torch.equal(out, eager)


def test_gelu_mlp_projector_module_matches_reference_sequential():
    """`mm_projector` drop-in: same state_dict keys as nn.Sequential(Linear, GELU, Linear) (cambrian_arch.py:65-69),
    output of tdc_gelu_mlp vs the fp32 torch modules."""
    from tdc_video_b200 import GeluMLPProjector
    torch.manual_seed(3)
    ref = torch.nn.Sequential(torch.nn.Linear(1024, 3584), torch.nn.GELU(), torch.nn.Linear(3584, 3584)).eval()
    mod = GeluMLPProjector(1024, 3584)
    assert set(mod.state_dict()) == set(ref.state_dict())
    mod.load_state_dict(ref.state_dict(), strict=True)
    mod = mod.cuda().eval()
    x = torch.randn(7, 144, 1024)
    with torch.no_grad():
        y_ref = ref(x)
        y = mod(x.cuda().bfloat16())
    assert y.shape == y_ref.shape and y.dtype == torch.bfloat16
    _metrics_ok(y, y_ref, "mm_projector module")


def test_speech_qformer_geometry():
    """the speech Q-Former of audio_models/audio_encoder.py:11-24,98-105 is the same module at another
    geometry (2 layers, cross-attention in every layer, ONE query token, BEATs/Whisper-width KV)."""
    from tdc_video_b200.qformer import QFormerConfig, TDCQFormer
    cfg = QFormerConfig(vocab_size=32, hidden_size=768, num_hidden_layers=2, num_attention_heads=12,
                        intermediate_size=3072, max_position_embeddings=16, encoder_width=2048,
                        cross_attention_freq=1, query_length=1)
    model = TDCQFormer(cfg, with_lm_head=False)
    _randomize(model, 9)
    model = model.cuda().eval()
    sd = {k[len("bert."):]: v.detach().cpu() for k, v in model.state_dict().items() if k.startswith("bert.")}
    assert "encoder.layer.1.crossattention.self.key.weight" in sd        # cross_attention_freq = 1
    B, L = 40, 17                                                         # 0.33 s windows of ~17 frames
    q = torch.randn(B, 1, 768, device="cuda")
    enc = torch.randn(B, L, 2048, device="cuda")
    out = model.bert(query_embeds=q, encoder_hidden_states=enc, return_dict=True).last_hidden_state
    ref = oracle.qformer_forward(sd, _geom_of(cfg, 0), q.cpu(), enc.cpu())
    assert out.shape == (B, 1, 768)
    _metrics_ok(out, ref, "speech qformer geometry")


def test_speech_qformer_module_matches_oracle():
    """The drop-in for audio_encoder.py:10-24, 75-116: LayerNorms, pad + concat, 17-frame windows, the 2-layer
    cross-every-layer Q-Former with ONE query, speech_llama_proj — Whisper-large (1280) + BEATs (768) widths."""
    from oracle import speech_oracle
    from tdc_video_b200.speech import TDCSpeechQFormer
    mod = TDCSpeechQFormer(1280, 768, llama_hidden_size=512, vocab_size=32)
    _randomize(mod, 21)
    with torch.no_grad():
        for ln in (mod.ln_speech, mod.ln_audio):
            ln.weight.add_(1.0)
    mod = mod.cuda().eval()
    sd = {}
    for k, v in mod.state_dict().items():
        k = k[len("speech_Qformer.bert."):] if k.startswith("speech_Qformer.bert.") else k
        sd[k] = v.detach().cpu()
    assert "encoder.layer.1.crossattention.self.key.weight" in sd and sd["speech_query_tokens"].shape == (1, 1, 768)
    B, T = 2, 100                                  # 100 frames -> 5 windows of 17 (the tail is dropped by unfold)
    speech = torch.randn(B, T, 1280)
    audio = torch.randn(B, T - 4, 768)             # shorter BEATs track: zero-padded (:83-84)
    y, atts = mod.encode_auditory_feature(speech.cuda(), audio.cuda())
    ref = speech_oracle.encode_auditory_feature(sd, _geom_of(mod.speech_Qformer.config, 0), speech, audio)
    assert y.shape == ref.shape == (B, 5, 512) and atts.shape == (B, 5) and bool(atts.all())
    _metrics_ok(y, ref, "speech qformer module")


@pytest.mark.parametrize("pattern", ["every_second", "sparse"])
def test_audio_pooling_matches_oracle(pattern):
    """tdc_video_b200.audio.pool_audio_per_frame (tdc_avg_pool_tokens kernel) vs the restated
    cambrian_arch.py:1547-1598 (itself pinned to the real reference function in tests/test_driver_pinning.py)."""
    from tdc_video_b200.audio import pool_audio_per_frame
    n_frames = 27
    rs = np.random.RandomState(3)
    if pattern == "every_second":
        seconds, flags = n_frames, [1] * n_frames
    else:
        seconds = 61
        pos = set(np.sort(rs.choice(seconds, size=n_frames, replace=False)).tolist())
        flags = [1 if i in pos else 0 for i in range(seconds)]
    n_win = (seconds + 9) // 10
    windows = []
    for w in range(n_win):
        tlen = min(10, seconds - 10 * w) * 50 - (7 if w == n_win - 1 else 0)
        windows.append(torch.from_numpy(rs.standard_normal((1, tlen, 768)).astype(np.float32)))
    ref = driver_oracle.audio_frames_from_beats(windows, flags, n_frames)
    got = pool_audio_per_frame([w.cuda().bfloat16() for w in windows], flags, n_frames)
    assert got.shape == ref.shape == (n_frames, 50, 768) and got.dtype == torch.bfloat16
    assert torch.allclose(got.float().cpu(), ref, atol=3e-2, rtol=3e-2)
    # frames without audio (zero padding at the end, cambrian_arch.py:1593-1595) are exactly zero in both
    empty = (ref.abs().amax(dim=(1, 2)) == 0)
    assert torch.equal(got.float().cpu().abs().amax(dim=(1, 2)) == 0, empty)


def test_two_engines_on_two_streams():
    """Handles are independent and calls are stream-ordered: two engines with different weights run
    concurrently on two non-default streams and reproduce their serial results bit for bit."""
    from oracle.synth import QFormerGeometry, make_state_dict
    from tdc_video_b200 import QFormerEngine
    geom = QFormerGeometry(hidden=128, heads=2, intermediate=256, layers=2, cross_freq=2, d_enc=64, d_out=96, vocab=0)
    engines, inputs, serial = [], [], []
    for i in range(2):
        eng = QFormerEngine(hidden=128, heads=2, intermediate=256, layers=2, cross_freq=2, d_enc=64, d_out=96)
        eng.load_weights(make_state_dict(geom, 20 + i, with_text=False))
        q = torch.randn(33, 16, 128, device="cuda")
        enc = torch.randn(33, 50, 64, device="cuda").bfloat16()
        engines.append(eng); inputs.append((q, enc)); serial.append(eng.compress(q, enc))
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [None, None]
    for rep in range(3):
        for i in range(2):
            with torch.cuda.stream(streams[i]):
                outs[i] = engines[i].compress(*inputs[i])
    for s in streams:
        s.synchronize()
    for i in range(2):
        assert torch.equal(outs[i], serial[i])
    assert not torch.equal(outs[0], outs[1])
