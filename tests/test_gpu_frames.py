"""The upstream entry (tdc_compress_frames: tower features in, TDC token sequence out) on the GPU against
 (1) committed outputs of the reference's REAL prepare_inputs_labels_for_multimodal run with the GELU-MLP
     projector (tests/golden/towers_*.npz), in both weight modes (fold = 1 / 0);
 (2) the CPU oracle at the shipped widths (hidden 768, 12 heads, d_llm 3584, 1024-wide tower features, audio);
 (3) itself: host streaming == device-resident call, folded == unfolded within the tolerance.
Tolerance as everywhere: per-token cosine >= 0.999, both relative errors <= 2e-2."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import driver_oracle
from oracle import qformer_oracle as oracle
from oracle.make_golden import DRIVER_D, DRIVER_GEOM, driver_audio, driver_tables, driver_weights_mlp

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "towers_*.npz")))


def _ok(test, ref, what):
    m = oracle.parity_metrics(test.float().cpu(), ref)
    print(what, m)
    assert m["min_cos"] >= 0.999 and m["max_abs_over_max_ref"] <= 2e-2 and m["max_tok_rel_l2"] <= 2e-2, (what, m)
    return m


def _compressor(m, w, d_in):
    from tdc_video_b200.compressor import TDCCompressor
    from tdc_video_b200.qformer import QFormerConfig
    g = DRIVER_GEOM
    cfg = QFormerConfig(vocab_size=g.vocab, hidden_size=g.hidden, num_hidden_layers=g.layers,
                        num_attention_heads=g.heads, intermediate_size=g.intermediate,
                        max_position_embeddings=g.max_pos, layer_norm_eps=g.ln_eps,
                        cross_attention_freq=g.cross_freq, encoder_width=DRIVER_D, query_length=m["num_query"])
    comp = TDCCompressor(DRIVER_D, context_token_num=m["num_query"], query_type=m["query_type"], text_input=m["text"],
                         add_static=m["add_static"], audio_input=bool(m.get("audio")), qformer_config=cfg,
                         mm_input_size=d_in)
    sd = {}
    for k, v in w.items():
        if k.startswith(("embeddings.", "encoder.")):
            sd["Qformer.bert." + k] = torch.from_numpy(v)
        elif k.split(".")[0] in ("vision_proj", "query_proj", "frame_seg", "query_tokens", "audio_proj", "mm_projector",
                                 "image_newline"):
            sd[k] = torch.from_numpy(v)
    missing, unexpected = comp.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in k or k.startswith("Qformer.cls.") for k in missing), missing
    return comp.cuda().eval()


@pytest.mark.parametrize("fold", [True, False], ids=["fold", "literal"])
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[7:-4] for p in GOLDEN])
def test_frames_stage_matches_the_reference_driver(path, fold):
    """adapt_segment -> tdc_compress_frames -> assembly, against the real reference function's token sequence."""
    from tdc_video_b200.pipeline import tdc_video_stage
    z = np.load(path)
    m = json.loads(str(z["meta"]))
    w = driver_weights_mlp(m["weight_seed"], m["num_query"])
    sig, dino = driver_tables(m["table_seed"], m["n_frames"])
    n = m["n_frames"]
    audio_kw = {}
    if m.get("audio"):
        windows, flags, _, proj = driver_audio(m["audio_seed"], n, m["audio"])
        w.update(proj)
        audio_kw = dict(audio_windows=[torch.from_numpy(a).cuda() for a in windows], sample_indices=flags)
    feats = torch.from_numpy(np.concatenate([sig, dino], -1)).cuda()          # the input of mm_projector (:1149)
    comp = _compressor(m, w, feats.shape[-1])
    ids = torch.tensor([m["prompt_ids"]], device="cuda")
    seq, selected, bounds = tdc_video_stage(comp, None, torch.from_numpy(dino).cuda(), input_ids=ids,
                                            max_visual_len=m["max_visual_len"], return_segments=True,
                                            tower_features=feats, fold=fold, **audio_kw)
    torch.cuda.synchronize()
    assert bounds.cpu().tolist() == z["segment_frame_indices"].tolist()
    ref = torch.from_numpy(z["visual_tokens"])
    assert tuple(seq.shape) == tuple(ref.shape)
    _ok(seq, ref, f"{os.path.basename(path)} fold={fold}")


def _full_problem(seed, n_frames, chunk, audio=True, T=0, d=3584, d_in=1024):
    """Shipped widths: state dict + tower features + audio tokens + a plan of `chunk`-frame chunks."""
    from tdc_video_b200.synth import QFormerGeometry, make_frontend_state_dict, make_state_dict
    geom = QFormerGeometry(d_enc=d, d_out=d, vocab=30522 if T else 0)
    sd = make_state_dict(geom, seed, stress=2.0, with_text=T > 0)
    sd.update(make_frontend_state_dict(d, d_in, 768 if audio else 0, geom.hidden, seed + 1))
    rs = np.random.RandomState(seed + 2)
    frames = rs.standard_normal((n_frames, 144, d_in)).astype(np.float32)
    aud = (rs.standard_normal((n_frames, 50, 768)) * 0.5).astype(np.float32) if audio else None
    sizes = [chunk] * (n_frames // chunk) + ([n_frames % chunk] if n_frames % chunk else [])
    return geom, sd, frames, aud, sizes


def _oracle_frames(sd, geom, frames, aud, sizes, K, ids=None):
    """CPU oracle of the frames stage on the bf16-rounded inputs the GPU sees."""
    from oracle import frames_oracle
    from tdc_video_b200.compressor import plan_chunks
    p = plan_chunks(sizes, True)
    x = torch.from_numpy(frames).bfloat16().float()
    a = None if aud is None else torch.from_numpy(aud).bfloat16().float()
    return frames_oracle.frames_stage(sd, geom, x, a, p.static_frames, p.chunk_len, K, ids)


def _engine(geom, sd, d_in, audio, T=0):
    from tdc_video_b200 import QFormerEngine
    eng = QFormerEngine(d_enc=geom.d_enc, d_out=geom.d_out, vocab=30522 if T else 0, d_frame_in=d_in,
                        d_audio=768 if audio else 0)
    eng.load_weights(sd)
    return eng


def _plan(sizes):
    from tdc_video_b200.compressor import plan_chunks
    p = plan_chunks(sizes, True)
    i32 = lambda a: torch.from_numpy(a.astype(np.int32))
    return p, i32(p.static_frames), i32(p.row_frames), i32(p.row_chunk)


@pytest.mark.parametrize("fold", [True, False], ids=["fold", "literal"])
def test_frames_entry_full_widths_against_oracle(fold):
    geom, sd, frames, aud, sizes = _full_problem(11, 22, 4)      # 5 chunks of 4 + one of 2 -> 16 rows, 6 key frames
    eng = _engine(geom, sd, 1024, True)
    p, sf, rf, rc = _plan(sizes)
    st, comp = eng.compress_frames(torch.from_numpy(frames).cuda().bfloat16(), sf, rf, rc,
                                   audio=torch.from_numpy(aud).cuda().bfloat16(), fold=fold, out_dtype=torch.float32)
    torch.cuda.synchronize()
    ref_st, ref_comp = _oracle_frames(sd, geom, frames, aud, sizes, 16)
    assert tuple(st.shape) == tuple(ref_st.shape) == (6, 206, 3584)
    assert tuple(comp.shape) == tuple(ref_comp.shape) == (16, 16, 3584)
    _ok(st, ref_st, f"static fold={fold}")
    _ok(comp, ref_comp, f"compressed fold={fold}")
    norms = comp.float().norm(dim=-1)
    assert torch.allclose(norms, torch.ones_like(norms), atol=2e-3)


def test_frames_entry_text_and_ragged_chunks():
    """A shared prompt (text_input mode), chunks of 8 / 1 / 3 frames (a single-frame chunk has no rows)."""
    geom, sd, frames, aud, _ = _full_problem(12, 12, 8, audio=False, T=6)
    sizes = [8, 1, 3]
    eng = _engine(geom, sd, 1024, False, T=6)
    ids = torch.randint(1000, 30000, (1, 6), generator=torch.Generator().manual_seed(3))
    p, sf, rf, rc = _plan(sizes)
    assert p.num_chunks == 3 and p.num_rows == 9
    st, comp = eng.compress_frames(torch.from_numpy(frames).cuda().bfloat16(), sf, rf, rc, input_ids=ids.cuda(),
                                   out_dtype=torch.float32)
    ref_st, ref_comp = _oracle_frames(sd, geom, frames, None, sizes, 16, ids)
    assert tuple(st.shape) == (3, 156, 3584)
    _ok(st, ref_st, "static text")
    _ok(comp, ref_comp, "compressed text")
    # layer-0 de-duplication (embeddings + self-attention block of layer 0 once per chunk, incl. the prompt tokens)
    # does not change a bit
    st2, comp2 = eng.compress_frames(torch.from_numpy(frames).cuda().bfloat16(), sf, rf, rc, input_ids=ids.cuda(),
                                     out_dtype=torch.float32, layer0_dedup=False)
    assert torch.equal(comp, comp2) and torch.equal(st, st2)


def test_frames_small_workspace_and_host_streaming_give_the_same_bits():
    """Row batching inside the call (a workspace for 5 items) and the host-streaming entry (ranges of 2 chunks)
    are invisible: rows are independent."""
    geom, sd, frames, aud, sizes = _full_problem(13, 23, 4)
    eng = _engine(geom, sd, 1024, True)
    p, sf, rf, rc = _plan(sizes)
    f_dev, a_dev = torch.from_numpy(frames).cuda().bfloat16(), torch.from_numpy(aud).cuda().bfloat16()
    st, comp = eng.compress_frames(f_dev, sf, rf, rc, audio=a_dev)
    _, comp_nd = eng.compress_frames(f_dev, sf, rf, rc, audio=a_dev, layer0_dedup=False, want_static=False)
    assert torch.equal(comp, comp_nd)               # layer-0 de-duplication is bit-identical
    small = int(eng.lib.tdc_frames_workspace_bytes(eng._h, p.num_chunks, p.num_rows, 5, 144, 50, 16, 0))
    eng._ws, eng.max_frames_workspace_bytes = None, small
    st2, comp2 = eng.compress_frames(f_dev, sf, rf, rc, audio=a_dev)
    assert eng._ws.numel() == small
    assert torch.equal(st, st2) and torch.equal(comp, comp2)
    eng._ws, eng.max_frames_workspace_bytes = None, 11 << 30
    chunk_start = p.static_frames
    out_host = torch.empty((p.num_rows, 16, 3584), dtype=torch.bfloat16, pin_memory=True)
    st3 = torch.empty_like(st)
    eng.compress_frames_host(torch.from_numpy(frames).bfloat16().pin_memory(), torch.from_numpy(aud).bfloat16().pin_memory(),
                             chunk_start, p.chunk_len, out_host, static_out=st3, chunks_per_batch=2)
    torch.cuda.synchronize()
    assert torch.equal(out_host.cuda(), comp) and torch.equal(st3, st)


def test_frames_entry_errors():
    from tdc_video_b200 import QFormerEngine, TdcError
    from tdc_video_b200.synth import QFormerGeometry, make_state_dict
    geom = QFormerGeometry(hidden=128, heads=2, intermediate=256, layers=2, cross_freq=2, d_enc=64, d_out=64, vocab=0)
    eng = QFormerEngine(hidden=128, heads=2, intermediate=256, layers=2, cross_freq=2, d_enc=64, d_out=64,
                        d_frame_in=32)
    eng.load_weights(make_state_dict(geom, 1, with_text=False))     # no mm_projector / newline / query_proj tensors
    z = torch.zeros(1, dtype=torch.int32)
    with pytest.raises(TdcError):
        eng.compress_frames(torch.zeros(2, 16, 32, device="cuda", dtype=torch.bfloat16), z, z + 1, z)
    with pytest.raises(RuntimeError):
        QFormerEngine(hidden=128, heads=2, intermediate=256, layers=2, cross_freq=2, d_enc=64, d_out=64).compress_frames(
            torch.zeros(2, 16, 32, device="cuda", dtype=torch.bfloat16), z, z + 1, z)


def test_frames_entry_edge_cases():
    """Only single-frame chunks (no rows at all), learned queries with a prompt (layer-0 state computed once in
    total), fp16 outputs, and a non-square token grid rejected loudly."""
    from tdc_video_b200 import TdcError
    geom, sd, frames, aud, _ = _full_problem(14, 6, 1, audio=True, T=5)
    eng = _engine(geom, sd, 1024, True, T=5)
    f_dev, a_dev = torch.from_numpy(frames).cuda().bfloat16(), torch.from_numpy(aud).cuda().bfloat16()
    ids = torch.randint(1000, 30000, (1, 5), generator=torch.Generator().manual_seed(5))
    # (1) six chunks of one frame: key frames pass through, nothing to compress
    p, sf, rf, rc = _plan([1] * 6)
    assert p.num_rows == 0
    st, comp = eng.compress_frames(f_dev, sf, rf, rc, audio=a_dev, input_ids=ids.cuda(), out_dtype=torch.float32)
    ref_st, ref_comp = _oracle_frames(sd, geom, frames, aud, [1] * 6, 16, ids)
    assert comp.shape == (0, 16, 3584) and ref_comp.shape[0] == 0
    _ok(st, ref_st, "key frames only")
    # (2) learned queries + prompt, chunks of 4 and 2 frames, fp16 output
    from oracle import frames_oracle
    p, sf, rf, rc = _plan([4, 2])
    st16, comp16 = eng.compress_frames(f_dev, sf, rf, rc, audio=a_dev, input_ids=ids.cuda(), learned_queries=True,
                                       out_dtype=torch.float16)
    x, a = torch.from_numpy(frames).bfloat16().float(), torch.from_numpy(aud).bfloat16().float()
    ref_st, ref_comp = frames_oracle.frames_stage(sd, geom, x, a, p.static_frames, p.chunk_len, 16, ids,
                                                  learned_queries=True)
    assert comp16.dtype == torch.float16 and tuple(comp16.shape) == (4, 16, 3584)
    _ok(st16, ref_st, "learned/text static fp16")
    _ok(comp16, ref_comp, "learned/text compressed fp16")
    _, comp_nd = eng.compress_frames(f_dev, sf, rf, rc, audio=a_dev, input_ids=ids.cuda(), learned_queries=True,
                                     out_dtype=torch.float16, layer0_dedup=False, want_static=False)
    assert torch.equal(comp16, comp_nd)
    # (3) index arrays are device data the library cannot validate: a bad frame index is clamped, not followed
    rf_bad = rf.clone()
    rf_bad[0] = 10 ** 6
    _, comp_bad = eng.compress_frames(f_dev, sf, rf_bad, rc, audio=a_dev, input_ids=ids.cuda(), learned_queries=True,
                                      out_dtype=torch.float16, want_static=False)
    torch.cuda.synchronize()
    assert torch.equal(comp_bad[1:], comp16[1:]) and bool(torch.isfinite(comp_bad.float()).all())
    # (4) 140 visual tokens are not a square grid
    with pytest.raises(TdcError):
        eng.compress_frames(f_dev[:, :140].contiguous(), sf, rf, rc, audio=a_dev)


def test_several_videos_with_their_own_prompts_in_one_call():
    """compress_videos_from_towers: three clips of different lengths, each with its own question, through ONE
    tdc_compress_frames call (chunk -> prompt map, layer-0 state per chunk) == the per-video calls, bit for bit."""
    from tdc_video_b200.compressor import TDCCompressor
    from tdc_video_b200.qformer import QFormerConfig
    torch.manual_seed(3)
    cfg = QFormerConfig(vocab_size=64, hidden_size=128, num_hidden_layers=4, num_attention_heads=2,
                        intermediate_size=256, max_position_embeddings=16)
    for query_type in ("Avg_pool", "learned"):
        comp = TDCCompressor(96, context_token_num=8, query_type=query_type, audio_input=True, qformer_config=cfg,
                             mm_input_size=40).cuda().eval()
        vids = []
        for n, sizes in ((13, [5, 1, 7]), (9, [9]), (4, [2, 2])):
            vids.append(dict(tower_features=torch.randn(n, 16, 40, device="cuda", dtype=torch.bfloat16),
                             segment_sizes=sizes, input_ids=torch.randint(1, 64, (1, 5), device="cuda"),
                             audio_frames=torch.randn(n, 6, 768, device="cuda", dtype=torch.bfloat16),
                             max_visual_len=None))
        together = comp.compress_videos_from_towers(vids)
        for v, got in zip(vids, together):
            alone = comp.compress_video_from_towers(v["tower_features"], v["segment_sizes"], input_ids=v["input_ids"],
                                                    audio_frames=v["audio_frames"])
            assert torch.equal(got, alone), query_type


def test_key_frames_ready_event_lets_a_side_stream_ship_them_early():
    """tdc_frames_args.static_ready_event is recorded once the key frames' tokens are complete; a side stream that
    waits for it copies them with the copy engines (tdc_peer_copy) while the rows are still being compressed.  The
    copy equals static_out and the compressed tokens are unchanged by the extra event."""
    import ctypes as C
    from tdc_video_b200 import _lib
    geom, sd, frames, aud, sizes = _full_problem(17, 40, 4)
    eng = _engine(geom, sd, 1024, True)
    p, sf, rf, rc = _plan(sizes)
    x, a = torch.from_numpy(frames).cuda().bfloat16(), torch.from_numpy(aud).cuda().bfloat16()
    st0, comp0 = eng.compress_frames(x, sf, rf, rc, audio=a)
    ev = torch.cuda.Event()
    side = torch.cuda.Stream()
    shipped = torch.zeros_like(st0)
    torch.cuda.synchronize()          # (the side stream below is ordered after `ev` only, not after this fill)
    st1, comp1 = eng.compress_frames(x, sf, rf, rc, audio=a, static_ready_event=ev)
    side.wait_event(ev)
    lib = _lib.load_library()
    rc_ = lib.tdc_peer_copy(C.c_void_p(st1.data_ptr()), C.c_void_p(shipped.data_ptr()),
                            st1.numel() * st1.element_size(), C.c_void_p(side.cuda_stream))
    assert rc_ == 0
    side.synchronize()
    assert torch.equal(shipped, st0)
    torch.cuda.synchronize()
    assert torch.equal(st1, st0) and torch.equal(comp1, comp0)
    assert lib.tdc_peer_copy(None, C.c_void_p(shipped.data_ptr()), 16, None) == -1
    assert lib.tdc_multicast_copy(C.c_void_p(st1.data_ptr()), C.c_void_p(shipped.data_ptr() + 8), 32, 0, None) == -1


def test_frames_entry_captures_into_a_cuda_graph():
    """The header's contract: stream-ordered, no host reads of device data -> the whole call replays from a CUDA graph
    with the same bits, also after the inputs changed in place."""
    geom, sd, frames, aud, sizes = _full_problem(23, 36, 4)
    eng = _engine(geom, sd, 1024, True)
    p, sf, rf, rc = _plan(sizes)
    sf, rf, rc = sf.cuda(), rf.cuda(), rc.cuda()
    x, a = torch.from_numpy(frames).cuda().bfloat16(), torch.from_numpy(aud).cuda().bfloat16()
    st0, comp0 = eng.compress_frames(x, sf, rf, rc, audio=a)          # also sizes the workspace
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            st1, comp1 = eng.compress_frames(x, sf, rf, rc, audio=a)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(st1, st0) and torch.equal(comp1, comp0)
    x2 = torch.roll(x, 5, dims=0).contiguous()
    st2, comp2 = eng.compress_frames(x2, sf, rf, rc, audio=a)
    torch.cuda.synchronize()
    x.copy_(x2)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(st1, st2) and torch.equal(comp1, comp2) and not torch.equal(comp2, comp0)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_frames_entry_16_bit_outputs_are_the_rounded_fp32_outputs(dtype):
    """I/O in bf16 and in fp16 (the reference's inference dtype, builder.py:69).  The 16-bit outputs are the
    fp32 outputs rounded once (same kernels, only the final stores differ)."""
    geom, sd, frames, aud, sizes = _full_problem(29, 14, 4)
    eng = _engine(geom, sd, 1024, True)
    p, sf, rf, rc = _plan(sizes)
    x, a = torch.from_numpy(frames).cuda().to(dtype), torch.from_numpy(aud).cuda().to(dtype)
    st32, c32 = eng.compress_frames(x, sf, rf, rc, audio=a, out_dtype=torch.float32)
    st16, c16 = eng.compress_frames(x, sf, rf, rc, audio=a, out_dtype=dtype)
    torch.cuda.synchronize()
    assert st16.dtype == dtype and c16.dtype == dtype
    assert torch.equal(c16, c32.to(dtype))
    assert torch.equal(st16, st32.to(dtype))


def test_outputs_written_in_place_into_caller_buffers():
    """static_into / out_into: the library writes the key-frame and compressed tokens straight into slices of a
    whole-video buffer (no fresh tensors, no extra device copy); bad destinations are rejected."""
    geom, sd, frames, aud, sizes = _full_problem(31, 20, 4)
    eng = _engine(geom, sd, 1024, True)
    p, sf, rf, rc = _plan(sizes)
    x, a = torch.from_numpy(frames).cuda().bfloat16(), torch.from_numpy(aud).cuda().bfloat16()
    st0, c0 = eng.compress_frames(x, sf, rf, rc, audio=a)
    big_s = torch.zeros((st0.shape[0] + 3,) + tuple(st0.shape[1:]), dtype=torch.bfloat16, device="cuda")
    big_c = torch.zeros((c0.shape[0] + 5,) + tuple(c0.shape[1:]), dtype=torch.bfloat16, device="cuda")
    st1, c1 = eng.compress_frames(x, sf, rf, rc, audio=a, static_into=big_s[2:2 + st0.shape[0]],
                                  out_into=big_c[4:4 + c0.shape[0]])
    torch.cuda.synchronize()
    assert st1.data_ptr() == big_s[2:].data_ptr() and c1.data_ptr() == big_c[4:].data_ptr()
    assert torch.equal(big_s[2:2 + st0.shape[0]], st0) and torch.equal(big_c[4:4 + c0.shape[0]], c0)
    assert not big_s[:2].any() and not big_s[2 + st0.shape[0]:].any() and not big_c[:4].any()
    with pytest.raises(ValueError):
        eng.compress_frames(x, sf, rf, rc, audio=a, out_into=big_c)                       # wrong row count
    with pytest.raises(ValueError):
        eng.compress_frames(x, sf, rf, rc, audio=a, static_into=big_s[2:2 + st0.shape[0]].float())   # wrong dtype
