"""Host-side integer logic of the TDC driver (chunk planning, token offsets, budget truncation)
against the literal restatement of the reference loop (oracle/driver_oracle.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import driver_oracle
from oracle.synth import QFormerGeometry, make_state_dict
from tdc_video_b200.compressor import output_layout, plan_chunks, truncation_keep_index

GEOM = QFormerGeometry(hidden=64, heads=1, intermediate=64, layers=1, cross_freq=1, d_enc=32, d_out=32, vocab=0)


def _weights(seed=3):
    rs = np.random.RandomState(seed)
    w = make_state_dict(GEOM, seed, with_text=False)
    w["query_proj.weight"] = (rs.standard_normal((64, 32)) * 0.1).astype(np.float32)
    w["query_proj.bias"] = (rs.standard_normal((64,)) * 0.1).astype(np.float32)
    w["frame_seg"] = rs.standard_normal((32,)).astype(np.float32)
    return w


@pytest.mark.parametrize("segment_sizes", [[1], [8], [9], [3, 1, 17, 8, 2], [0, 5, 0, 16], [25]])
@pytest.mark.parametrize("keep_static", [True, False])
def test_plan_and_layout_match_reference_loop(segment_sizes, keep_static):
    n = sum(segment_sizes)
    Lv, K = 6, 4
    frames = torch.randn(n, Lv, 32)
    seq, chunks, n_calls = driver_oracle.compress_video(_weights(), GEOM, frames, segment_sizes, context_token_num=K,
                                                        add_text=False, keep_static=keep_static, return_chunks=True)
    plan = plan_chunks(segment_sizes, keep_static)
    off, tok, row_off = output_layout(plan, Lv, K, keep_static)
    assert plan.num_chunks == len(chunks)
    assert tok.tolist() == [c.shape[0] for c in chunks]
    assert int(tok.sum()) == seq.shape[0]
    assert n_calls == int((plan.rows_per_chunk > 0).sum())
    assert (plan.rows_per_chunk <= (7 if keep_static else 8)).all()
    # static frames sit where the plan says, frame_seg right behind them
    if keep_static:
        for c, o in enumerate(off.tolist()):
            assert torch.equal(seq[o:o + Lv], frames[plan.static_frames[c]])
    # every frame is either a static frame or a row, exactly once (keep_static) / rows cover all frames
    covered = sorted(plan.row_frames.tolist() + (plan.static_frames.tolist() if keep_static else []))
    assert covered == list(range(n))
    # rows of a chunk follow its static frame contiguously
    for r in range(plan.num_rows):
        c = plan.row_chunk[r]
        assert 0 <= plan.row_frames[r] - plan.static_frames[c] < 8


@pytest.mark.parametrize("budget", [10_000, 61, 40, 17, 5])
def test_budget_truncation_matches_reference(budget):
    segment_sizes = [3, 9, 1, 4]
    Lv, K = 5, 3
    frames = torch.randn(sum(segment_sizes), Lv, 32)
    ref = driver_oracle.compress_video(_weights(), GEOM, frames, segment_sizes, context_token_num=K, add_text=False,
                                       max_visual_len=budget)
    full = driver_oracle.compress_video(_weights(), GEOM, frames, segment_sizes, context_token_num=K, add_text=False)
    plan = plan_chunks(segment_sizes)
    off, tok, _ = output_layout(plan, Lv, K)
    keep = truncation_keep_index(off, tok, budget)
    got = full if keep is None else full[torch.from_numpy(keep)]
    assert got.shape == ref.shape and torch.equal(got, ref)


def test_reference_token_accounting_example():
    """SURVEY §8c probe: 60 frames in 25 segments -> 25 static frames x (156+1) + 35 rows x (16+1) = 4520."""
    sizes = [8, 8, 8, 8, 4] + [1] * 16 + [2, 2, 2, 2]   # 25 segments, 60 frames
    assert sum(sizes) == 60 and len(sizes) == 25
    plan = plan_chunks(sizes)
    _, tok, _ = output_layout(plan, 156, 16)
    assert plan.num_chunks == 25 and plan.num_rows == 35 and int(tok.sum()) == 25 * 157 + 35 * 17 == 4520


def test_append_newline_tokens_and_stage_guards():
    """pipeline.py host-side pieces: the newline column layout (cambrian_arch.py:1269-1281) is pure data movement
    and the composed stage refuses CPU tensors (no fallback)."""
    import torch
    from tdc_video_b200.pipeline import append_newline_tokens, tdc_video_stage
    x = torch.arange(2 * 16 * 3, dtype=torch.float32).view(2, 16, 3)
    nl = torch.tensor([-1.0, -2.0, -3.0])
    y = append_newline_tokens(x, nl)
    ref = torch.cat([x.view(2, 4, 4, 3), nl.view(1, 1, 1, 3).expand(2, 4, 1, 3)], dim=2).flatten(1, 2)
    assert y.shape == (2, 20, 3) and torch.equal(y, ref)
    with pytest.raises(ValueError):
        append_newline_tokens(x[:, :15], nl)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tdc_video_stage(None, torch.zeros(3, 20, 8), torch.zeros(3, 4, 8))
