"""Multi-rank host logic on CPU: world_size-2 (and 3) gloo process groups exercising the chunk
sharding and the padded all-gather that assembles the ordered compressed-token sequence."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tdc_video_b200.compressor import plan_chunks
from tdc_video_b200.dist import all_gather_rows, row_range_of_chunks, shard_chunk_ranges


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("sizes,world", [([8] * 10, 2), ([3, 1, 17, 8, 2, 30], 3), ([1, 1], 4), ([25], 8)])
def test_shard_ranges_cover_chunks_and_balance_rows(sizes, world):
    plan = plan_chunks(sizes)
    ranges = shard_chunk_ranges(plan.rows_per_chunk, world)
    assert len(ranges) == world and ranges[0][0] == 0 and ranges[-1][1] == plan.num_chunks
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    rows = [row_range_of_chunks(plan.rows_per_chunk, lo, hi) for lo, hi in ranges]
    assert rows[0][0] == 0 and rows[-1][1] == plan.num_rows
    counts = [hi - lo for lo, hi in rows]
    assert sum(counts) == plan.num_rows
    if plan.num_chunks >= 4 * world:   # enough chunks: imbalance is bounded by one chunk (<= 7 rows)
        assert max(counts) - min(counts) <= 7 + 7


def _worker(rank, world, port, rows_per_rank, K, d, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # "compressed tokens" of global row i are filled with i so that the order is checkable
        lo = sum(rows_per_rank[:rank])
        local = torch.arange(lo, lo + rows_per_rank[rank], dtype=torch.float32).view(-1, 1, 1).expand(-1, K, d)
        out = all_gather_rows(local.contiguous(), rows_per_rank)
        ok = out.shape == (sum(rows_per_rank), K, d) and torch.equal(
            out[:, 0, 0], torch.arange(sum(rows_per_rank), dtype=torch.float32))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("rows_per_rank", [[5, 5], [7, 3], [0, 4], [6, 6, 1]])
def test_all_gather_rows_gloo(rows_per_rank):
    world = len(rows_per_rank)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, rows_per_rank, 3, 4, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, True) for r in range(world)]
