"""Host-side logic of the upstream entry (no GPU): chunk-range planning of the host-streaming call, the per-range
integer plan handed to tdc_compress_frames, and the FLOP accounting bench.py reports for it."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from tdc_video_b200.compressor import plan_chunks  # noqa: E402
from tdc_video_b200.engine import plan_chunk_ranges, range_plan  # noqa: E402


@pytest.mark.parametrize("head", [False, True])
@pytest.mark.parametrize("n,cb", [(1, 4), (7, 4), (3600, 300), (3600, 256), (100, 1000), (9, 1), (700, 300)])
def test_chunk_ranges_cover_everything_once(n, cb, head):
    b = plan_chunk_ranges(n, cb, taper_head=head)
    if head and n > 2 * cb and cb >= 16:
        assert b[0][1] - b[0][0] == cb // 8
    assert b[0][0] == 0 and b[-1][1] == n
    assert all(x[1] == y[0] for x, y in zip(b, b[1:])) and all(lo < hi for lo, hi in b)
    assert max(hi - lo for lo, hi in b) <= max(1, min(cb, n))
    if n > 2 * cb and cb >= 16:                       # tapered tail: the last range is much smaller than a batch
        assert b[-1][1] - b[-1][0] <= cb // 4


def test_range_plan_matches_the_global_plan():
    sizes = [9, 1, 12, 5, 8, 3]
    p = plan_chunks(sizes, True)
    rows_seen = []
    for a, b in plan_chunk_ranges(p.num_chunks, 3, taper_tail=False):
        f0, f1, st, rf, rck = range_plan(p.static_frames, p.chunk_len, a, b)
        assert f0 == p.static_frames[a] and f1 == p.static_frames[b - 1] + p.chunk_len[b - 1]
        assert np.array_equal(st + f0, p.static_frames[a:b])
        rows_seen.append(rf + f0)
        assert len(rck) == len(rf) and (len(rf) == 0 or (rck.max() < b - a and (np.diff(rck) >= 0).all()))
        assert np.array_equal(np.bincount(rck, minlength=b - a), p.rows_per_chunk[a:b])
    assert np.array_equal(np.concatenate(rows_seen), p.row_frames)


def test_frames_flop_accounting():
    """Model FLOPs = the reference formulation (SURVEY 8d per-row figure + mm_projector / audio_proj on every frame);
    executed FLOPs = the folded path; per video-second at the north-star shapes."""
    w = dict(bench.WORKLOADS["hour_qwen7b"], num_text=0)
    model, executed, kv_exec = bench.frames_flops(w)
    f_row, f_row_kv = bench.flops_per_row(206, 3584, 16, 0, 3584)
    proj = 2 * 144 * (1024 * 3584 + 3584 * 3584)
    audio = 2 * 50 * 768 * 3584
    assert model == 4 * (proj + audio) + 2 * 16 * 3584 * 768 + 3 * f_row
    assert abs(model / 1e9 - 70.35) < 0.01 and abs(executed / 1e9 - 48.26) < 0.01
    assert kv_exec == 3 * (2 * 144 * 3584 + 2 * 50 * 768) * 9216
    assert executed < model and kv_exec < 3 * f_row_kv


def test_range_plans_of_random_videos_tile_the_global_plan():
    """Property: for any segment sizes and any range size (with or without the head / tail tapers), the per-range
    plans handed to tdc_compress_frames are exactly the global plan cut at chunk boundaries."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(1, 30), min_size=1, max_size=25), st.integers(1, 40), st.booleans(), st.booleans())
    def check(sizes, cb, head, tail):
        p = plan_chunks(sizes, True)
        bounds = plan_chunk_ranges(p.num_chunks, cb, taper_tail=tail, taper_head=head)
        assert bounds[0][0] == 0 and bounds[-1][1] == p.num_chunks
        rows, statics, chunk_of_row = [], [], []
        for a, b in bounds:
            f0, f1, stt, rf, rck = range_plan(p.static_frames, p.chunk_len, a, b)
            assert 0 <= f0 < f1 <= sum(sizes) and (stt >= 0).all() and (rf < f1 - f0).all()
            statics.append(stt + f0)
            rows.append(rf + f0)
            chunk_of_row.append(rck + a)
        assert np.array_equal(np.concatenate(statics), p.static_frames)
        assert np.array_equal(np.concatenate(rows), p.row_frames)
        assert np.array_equal(np.concatenate(chunk_of_row), p.row_chunk)

    check()
