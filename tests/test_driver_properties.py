"""Property tests (hypothesis) of the driver's host-side integer logic against the literal restatement of the
reference loop: for ANY segment split / K / budget the planned layout equals the token sequence the
reference assembles (checked on marker tokens, no Q-Former needed).  CPU only."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from tdc_video_b200.compressor import output_layout, plan_chunks, truncation_keep_index
from tdc_video_b200.dist import row_range_of_chunks, shard_chunk_ranges


def _reference_assembly(segment_sizes, Ls, K, keep_static, budget):
    """cambrian_arch.py:1603-1709 with marker tokens: static token = (0, frame, t), frame_seg = (1, -1, -1),
    compressed token = (2, frame, q).  Returns the [tokens, 3] int sequence after truncation."""
    out_chunks, base = [], 0
    for n in segment_sizes:
        for start in range(0, n, 8):
            ln = min(8, n - start)
            f0 = base + start
            static = [(0, f0, t) for t in range(Ls)]
            sep = [(1, -1, -1)]
            if keep_static and ln == 1:
                out_chunks.append(static + sep)
                continue
            rows = range(f0 + 1, f0 + ln) if keep_static else range(f0, f0 + ln)
            body = []
            for f in rows:
                body += [(2, f, q) for q in range(K)] + sep
            out_chunks.append((static + sep + body) if keep_static else body)
        base += n
    total = sum(len(c) for c in out_chunks)
    if budget is not None and total > budget:
        import math
        fr = math.ceil((total - budget) / len(out_chunks))
        out_chunks = [c[:-fr] for c in out_chunks]
    seq = [t for c in out_chunks for t in c]
    if budget is not None:
        seq = seq[:budget]
    return np.asarray(seq, dtype=np.int64).reshape(-1, 3)


def _planned_assembly(segment_sizes, Ls, K, keep_static, budget):
    plan = plan_chunks(segment_sizes, keep_static)
    off, tok, row_off = output_layout(plan, Ls, K, keep_static)
    total = int(tok.sum())
    seq = np.full((total, 3), -7, dtype=np.int64)
    if keep_static:
        for c, o in enumerate(off.tolist()):
            seq[o:o + Ls] = [(0, plan.static_frames[c], t) for t in range(Ls)]
            seq[o + Ls] = (1, -1, -1)
    for r, o in enumerate(row_off.tolist()):
        seq[o:o + K] = [(2, plan.row_frames[r], q) for q in range(K)]
        seq[o + K] = (1, -1, -1)
    keep = truncation_keep_index(off, tok, budget)
    return seq if keep is None else seq[keep]


@settings(max_examples=200, deadline=None)
@given(sizes=st.lists(st.integers(0, 30), min_size=1, max_size=12).filter(lambda s: sum(s) > 0),
       Ls=st.integers(1, 9), K=st.integers(1, 5), keep_static=st.booleans(),
       budget=st.one_of(st.none(), st.integers(1, 400)))
def test_layout_equals_reference_assembly(sizes, Ls, K, keep_static, budget):
    ref = _reference_assembly(sizes, Ls, K, keep_static, budget)
    got = _planned_assembly(sizes, Ls, K, keep_static, budget)
    assert got.shape == ref.shape and (got == ref).all()
    assert not (got == -7).any()          # every planned slot was written exactly by one source


@settings(max_examples=200, deadline=None)
@given(sizes=st.lists(st.integers(0, 40), min_size=1, max_size=20).filter(lambda s: sum(s) > 0),
       world=st.integers(1, 8), keep_static=st.booleans())
def test_sharding_partitions_rows_in_order(sizes, world, keep_static):
    plan = plan_chunks(sizes, keep_static)
    ranges = shard_chunk_ranges(plan.rows_per_chunk, world)
    rows = [row_range_of_chunks(plan.rows_per_chunk, lo, hi) for lo, hi in ranges]
    flat = [r for lo, hi in rows for r in range(lo, hi)]
    assert flat == list(range(plan.num_rows))                      # contiguous, ordered, complete
    for (clo, chi), (rlo, rhi) in zip(ranges, rows):               # a row never leaves its chunk's rank
        assert all(clo <= plan.row_chunk[r] < chi for r in range(rlo, rhi))
