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


class MulticastGather:
    """Rank-ordered gather buffer `[world, rows, ...]` in torch symmetric memory with an NVSwitch
    multicast mapping.  Each rank's final kernel stores its rows through `slot_ptr()` (a multicast
    address), which delivers them into the same slot of EVERY rank's buffer — the all-gather is fused
    into the producing kernel instead of being a separate NCCL collective.  `barrier()` (device-side,
    stream-ordered) makes the peers' rows visible before `gathered` is read.

    Raises RuntimeError when symmetric memory / multicast is unavailable (callers fall back to the
    NCCL all-gather of `all_gather_rows`)."""

    def __init__(self, rows: int, tail: Sequence[int], dtype: torch.dtype, device, group: Optional[dist.ProcessGroup] = None):
        import torch.distributed._symmetric_memory as symm

        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.rows, self.tail = int(rows), tuple(int(t) for t in tail)
        self.buf = symm.empty((self.world, self.rows) + self.tail, dtype=dtype, device=device)
        self.handle = symm.rendezvous(self.buf, group.group_name)
        if not getattr(self.handle, "multicast_ptr", 0):
            raise RuntimeError("symmetric memory has no multicast mapping on this system")
        self.row_bytes = self.buf[0, 0].numel() * self.buf.element_size() if self.rows else 0
        self.device = torch.device(device)
        self._side: List[torch.cuda.Stream] = []
        self._side_used = False

    def slot_ptr(self, row0: int = 0) -> int:
        """Multicast address of row `row0` of this rank's slot."""
        return int(self.handle.multicast_ptr) + (self.rank * self.rows + int(row0)) * self.row_bytes

    def put_async(self, src: torch.Tensor, row0: int = 0, after: Optional[torch.cuda.Event] = None,
                  mode: str = "dma", ctas: int = 16) -> None:
        """Ship rows `src` (local, contiguous, the buffer's dtype) into rows [row0, row0 + len) of this rank's slot on
        EVERY rank, on side streams, so that the transfer overlaps whatever the current stream does next.
        mode "dma": one copy-engine transfer per peer (tdc_peer_copy into the peer's mapping of the symmetric buffer) —
        takes no SM, the right tool beside persistent compute kernels; "multimem": tdc_multicast_copy (multimem.st from
        `ctas` CTAs: one outbound copy, replicated by the switch).  `after`: event the side streams wait for (default:
        everything enqueued on the current stream so far).  `barrier()` joins the side streams."""
        from . import _lib
        from .engine import _ptr
        if src.dtype != self.buf.dtype or not src.is_contiguous():
            raise ValueError("put_async needs a contiguous tensor of the buffer's dtype")
        if tuple(src.shape[1:]) != self.tail or row0 < 0 or row0 + src.shape[0] > self.rows:
            raise ValueError("put_async: rows do not fit the slot")
        if mode not in ("dma", "multimem"):
            raise ValueError('mode must be "dma" or "multimem"')
        if not self._side:
            self._side = [torch.cuda.Stream(self.device) for _ in range(min(self.world, 4))]
        if after is None:
            after = torch.cuda.Event()
            after.record(torch.cuda.current_stream(self.device))
        lib = _lib.load_library()
        nbytes = src.numel() * src.element_size()
        offset = (self.rank * self.rows + int(row0)) * self.row_bytes
        with torch.cuda.device(self.device):
            if mode == "multimem":
                st = self._side[0]
                st.wait_event(after)
                src.record_stream(st)
                rc = lib.tdc_multicast_copy(_ptr(src), int(self.handle.multicast_ptr) + offset, nbytes, int(ctas),
                                            st.cuda_stream)
                _lib.check(rc, None, "tdc_multicast_copy")
            else:
                ptrs = self.handle.buffer_ptrs
                for i in range(self.world):
                    peer = (self.rank + i) % self.world      # start with the local copy, then a different peer per rank
                    st = self._side[i % len(self._side)]
                    st.wait_event(after)
                    src.record_stream(st)
                    rc = lib.tdc_peer_copy(_ptr(src), int(ptrs[peer]) + offset, nbytes, st.cuda_stream)
                    _lib.check(rc, None, "tdc_peer_copy")
        self._side_used = True

    def barrier(self) -> None:
        if self._side_used:
            for st in self._side:
                torch.cuda.current_stream(self.device).wait_stream(st)
            self._side_used = False
        self.handle.barrier()

    @property
    def gathered(self) -> torch.Tensor:
        return self.buf.view((self.world * self.rows,) + self.tail)
