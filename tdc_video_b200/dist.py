"""Multi-GPU plumbing for the TDC path: one process per GPU, `torch.distributed` (NCCL over
NVLink on a B200 box; gloo in the CPU tests).

Rows (dynamic frames) are independent and the weights are replicated, so the path shards
with no data-path collective until the very end: every rank compresses a contiguous range of
chunks (a row and its static frame stay together), then ONE all-gather of the compressed
tokens `[rows, K, d_out]` gives every rank — in particular the rank that runs the LLM — the
ordered sequence.  The reference has no intra-video parallelism at all
(eval shards whole videos over ranks, eval/eval_mlvu.py:129-156).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_chunk_ranges(rows_per_chunk: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Split chunks [0, C) into `world` contiguous ranges with near-equal ROW counts (rows are
    the unit of work; chunks are the unit that must not be split).  Greedy prefix cut at the
    ideal row boundaries; ranges may be empty when there are fewer chunks than ranks."""
    rpc = np.asarray(rows_per_chunk, dtype=np.int64)
    C = len(rpc)
    cum = np.concatenate([[0], np.cumsum(rpc)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        # first chunk boundary whose cumulative row count reaches the target (ties -> earlier)
        c = int(np.searchsorted(cum, target, side="left"))
        c = min(max(c, cuts[-1]), C)
        cuts.append(c)
    cuts.append(C)
    return [(cuts[i], cuts[i + 1]) for i in range(world)]


def row_range_of_chunks(rows_per_chunk: Sequence[int], chunk_lo: int, chunk_hi: int) -> Tuple[int, int]:
    cum = np.concatenate([[0], np.cumsum(np.asarray(rows_per_chunk, dtype=np.int64))])
    return int(cum[chunk_lo]), int(cum[chunk_hi])


def all_gather_rows(local: torch.Tensor, rows_per_rank: Sequence[int], group: Optional[dist.ProcessGroup] = None
                    ) -> torch.Tensor:
    """All-gather variable row counts: pad every rank to max(rows_per_rank) (fixed-size collective,
    in place into the rank-ordered buffer), then drop the padding.  local: [rows_per_rank[rank], ...]."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert len(rows_per_rank) == world and local.shape[0] == rows_per_rank[rank]
    rmax = max(int(r) for r in rows_per_rank)
    tail = tuple(local.shape[1:])
    if rmax == 0:
        return local.new_empty((0,) + tail)
    padded = local
    if local.shape[0] != rmax:
        padded = local.new_zeros((rmax,) + tail)
        padded[: local.shape[0]] = local
    out = local.new_empty((world * rmax,) + tail)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if all(int(r) == rmax for r in rows_per_rank):
        return out
    out = out.view((world, rmax) + tail)
    return torch.cat([out[r, : int(n)] for r, n in enumerate(rows_per_rank)], dim=0)
