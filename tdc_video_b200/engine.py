"""Handle-level wrapper: torch tensors in, libtdc_b200.so calls out.

torch is used for device memory, streams and dtype bookkeeping only; every FLOP on this path
happens in the library's sm_100a kernels.  No CPU fallback: constructing an engine without a
CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Mapping, Optional

import torch

from . import _lib
from ._lib import TdcConfig, TdcFramesArgs, TdcTensor, check

_DTYPES = {torch.bfloat16: _lib.TDC_BF16, torch.float16: _lib.TDC_F16, torch.float32: _lib.TDC_F32}


def _dt(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}; use bfloat16, float16 or float32") from None


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def plan_chunk_ranges(n_chunks: int, chunks_per_batch: int, taper_tail: bool = True, taper_head: bool = False):
    """Ranges of consecutive chunks for the host-streaming entries: full batches, then — because the caller waits
    for the LAST range's compute + read-back after the last byte has arrived — a tapered tail (half of what is left,
    again and again, down to 1/8 of a batch).  `taper_head`: the first ranges grow 1/8, 1/4, 1/2 of a batch, so
    that the copy nothing can overlap with (the first one) is short."""
    bounds, c0 = [], 0
    cb = max(1, min(int(chunks_per_batch), int(n_chunks)))
    if taper_head and cb >= 16:
        for frac in (8, 4, 2):
            if n_chunks - c0 > 2 * cb:
                bounds.append((c0, c0 + cb // frac))
                c0 += cb // frac
    while n_chunks - c0 > cb:
        bounds.append((c0, c0 + cb))
        c0 += cb
    while taper_tail and n_chunks - c0 > max(cb // 8, 1):
        step = (n_chunks - c0) // 2
        bounds.append((c0, c0 + step))
        c0 += step
    if c0 < n_chunks:
        bounds.append((c0, n_chunks))
    return bounds


def range_plan(chunk_start, chunk_len, a: int, b: int, keep_static: bool = True):
    """Integer plan of chunks [a, b) relative to the first frame of the range: (first frame, end frame,
    static_frames [b-a], row_frames [rows], row_chunk [rows]) — int32 numpy arrays for tdc_compress_frames."""
    import numpy as np
    cs = np.asarray(chunk_start, dtype=np.int64)
    cl = np.asarray(chunk_len, dtype=np.int64)
    f0, f1 = int(cs[a]), int(cs[b - 1] + cl[b - 1])
    st = (cs[a:b] - f0).astype(np.int32)
    first = 1 if keep_static else 0
    rows = [np.arange(s + first, s + n, dtype=np.int32) for s, n in zip(st, cl[a:b])]
    rf = np.concatenate(rows) if rows and sum(len(r) for r in rows) else np.zeros(0, np.int32)
    rck = np.repeat(np.arange(b - a, dtype=np.int32), (cl[a:b] - first).astype(np.int64))
    return f0, f1, st, rf, rck


def _raw_event(ev: torch.cuda.Event, device) -> int:
    """cudaEvent_t of a torch event (torch creates it lazily at the first record)."""
    if not ev.cuda_event:
        with torch.cuda.device(device):
            ev.record()
    return int(ev.cuda_event)


class QFormerEngine:
    """Owns one `tdc_handle` (re-packed bf16 weights on one GPU) and a growable workspace."""

    def __init__(self, *, hidden=768, heads=12, intermediate=3072, layers=12, cross_freq=2, d_enc=3584, d_out=0,
                 vocab=0, max_pos=512, ln_eps=1e-12, device=None, gemm_cta_group=0, d_frame_in=0, d_audio=0,
                 max_workspace_bytes: int = 8 << 30):
        if not torch.cuda.is_available():
            raise RuntimeError("tdc_video_b200 needs a CUDA (sm_100a) device: there is no CPU fallback")
        self.lib = _lib.load_library()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.cfg = TdcConfig(hidden, heads, intermediate, layers, cross_freq, d_enc, d_out, vocab, max_pos,
                             float(ln_eps), gemm_cta_group, d_frame_in, d_audio)
        self.geometry = dict(hidden=hidden, heads=heads, intermediate=intermediate, layers=layers,
                             cross_freq=cross_freq, d_enc=d_enc, d_out=d_out, vocab=vocab, max_pos=max_pos,
                             ln_eps=ln_eps, d_frame_in=d_frame_in, d_audio=d_audio)
        env_cap = os.environ.get("TDC_MAX_WORKSPACE_GB")   # dev knob: smaller workspace = smaller internal row batches
        self.max_workspace_bytes = int(float(env_cap) * (1 << 30)) if env_cap else int(max_workspace_bytes)
        # the upstream entry keeps more per row (tower features, gelu(mm_projector.0)): 11 GB = 1800-row batches at
        # the north-star shapes, the same batch the 8 GB cap gives the tokens entry (profiles/r02_frames_row_batch.txt)
        self.max_frames_workspace_bytes = int(float(env_cap) * (1 << 30)) if env_cap else max(int(max_workspace_bytes),
                                                                                              11 << 30)
        self._ws: Optional[torch.Tensor] = None
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.tdc_create(C.byref(self._h), C.byref(self.cfg)), None, "tdc_create")

    # -- lifecycle ------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.tdc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_weights(self, state: Mapping[str, "torch.Tensor"]) -> None:
        """`state`: reference key names relative to `Qformer.bert.` plus `vision_proj.*`
        (tensors or numpy arrays, any of fp32/fp16/bf16; copied to the device if needed)."""
        keep, table = [], []
        for name, value in state.items():
            t = value if isinstance(value, torch.Tensor) else torch.as_tensor(value)
            if name == "query_tokens" and t.dim() == 3:
                t = t[0]  # [1, K, hidden] parameter of the reference (cambrian_arch.py:420-423)
            if not t.dtype.is_floating_point or t.dim() not in (1, 2):
                continue  # e.g. embeddings.position_ids
            t = t.detach().to(self.device).contiguous()
            if t.dtype not in _DTYPES:
                t = t.float()
            keep.append(t)
            shape = (C.c_int64 * 4)(*(list(t.shape) + [0] * (4 - t.dim())))
            table.append(TdcTensor(name.encode(), t.data_ptr(), _dt(t), t.dim(), shape))
        arr = (TdcTensor * len(table))(*table)
        with torch.cuda.device(self.device):
            check(self.lib.tdc_load_weights(self._h, arr, len(table), _stream(self.device)), self._h,
                  "tdc_load_weights")
            torch.cuda.current_stream(self.device).synchronize()  # sources may be freed after return

    # -- workspace ------------------------------------------------------------------------
    def workspace_bytes(self, rows: int, kv_tokens: int, num_query: int, num_text: int) -> int:
        return int(self.lib.tdc_workspace_bytes(self._h, rows, kv_tokens, num_query, num_text))

    def _workspace(self, rows, kv_tokens, num_query, num_text) -> torch.Tensor:
        need = self.workspace_bytes(rows, kv_tokens, num_query, num_text)
        floor = self.workspace_bytes(1, kv_tokens, num_query, num_text)
        want = max(min(need, self.max_workspace_bytes), floor)
        if self._ws is None or self._ws.numel() < want:
            self._ws = None
            self._ws = torch.empty(want, dtype=torch.uint8, device=self.device)
        return self._ws

    # -- the hot path ---------------------------------------------------------------------
    def _run(self, fn, what, query_embeds, enc, input_ids, query_set, text_set, kv_len, out_width, out_tokens,
             out_dtype, out_ptr=None):
        if enc.dim() != 3 or query_embeds.dim() != 3:
            raise ValueError("query_embeds must be [sets, K, hidden] and enc [rows, L, d_enc]")
        rows, L, d_enc = enc.shape
        K = query_embeds.shape[1]
        if d_enc != self.cfg.d_enc or query_embeds.shape[2] != self.cfg.hidden:
            raise ValueError(f"width mismatch: enc {d_enc} vs d_enc {self.cfg.d_enc}, "
                             f"query_embeds {query_embeds.shape[2]} vs hidden {self.cfg.hidden}")
        T = 0 if input_ids is None else int(input_ids.shape[1])
        if query_set is None and query_embeds.shape[0] != rows:
            raise ValueError("query_embeds needs one query set per row unless query_set is given")
        if T > 0 and text_set is None and input_ids.shape[0] != rows:
            raise ValueError("input_ids needs one row per enc row unless text_set is given")
        dev = self.device
        enc = enc.to(dev).contiguous()
        query_embeds = query_embeds.to(dev).contiguous()
        ids = None if T == 0 else input_ids.to(dev, torch.int64).contiguous()
        qs = None if query_set is None else query_set.to(dev, torch.int32).contiguous()
        ts = None if text_set is None else text_set.to(dev, torch.int32).contiguous()
        kl = None if kv_len is None else kv_len.to(dev, torch.int32).contiguous()
        out_dtype = out_dtype or enc.dtype
        if out_ptr is None:
            out = torch.empty((rows, out_tokens(K, T), out_width), dtype=out_dtype, device=dev)
            dst = _ptr(out)
        else:  # caller-owned destination (e.g. a multicast address): nothing to return
            out, dst = None, C.c_void_p(int(out_ptr))
        if rows == 0:
            return out
        ws = self._workspace(rows, L, K, T)
        with torch.cuda.device(dev):
            rc = fn(self._h, _ptr(query_embeds), _dt(query_embeds), _ptr(qs), _ptr(ids), _ptr(ts), _ptr(enc), _dt(enc),
                    _ptr(kl), rows, L, K, T, dst, _DTYPES[out_dtype], _ptr(ws), ws.numel(), _stream(dev))
        check(rc, self._h, what)
        return out

    def forward(self, query_embeds, enc, input_ids=None, *, query_set=None, text_set=None, kv_len=None,
                out_dtype=None) -> torch.Tensor:
        """last_hidden_state [rows, K+T, hidden] (tdc_qformer_forward)."""
        return self._run(self.lib.tdc_qformer_forward, "tdc_qformer_forward", query_embeds, enc, input_ids, query_set,
                         text_set, kv_len, self.cfg.hidden, lambda K, T: K + T, out_dtype)

    def compress(self, query_embeds, enc, input_ids=None, *, query_set=None, text_set=None, kv_len=None,
                 out_dtype=None) -> torch.Tensor:
        """normalize(vision_proj(last_hidden_state[:, :K])) [rows, K, d_out] (tdc_compress)."""
        if self.cfg.d_out <= 0:
            raise RuntimeError("engine was created without d_out: no vision_proj")
        return self._run(self.lib.tdc_compress, "tdc_compress", query_embeds, enc, input_ids, query_set, text_set,
                         kv_len, self.cfg.d_out, lambda K, T: K, out_dtype)

    def compress_multicast(self, query_embeds, enc, multicast_ptr: int, input_ids=None, *, query_set=None,
                           text_set=None, kv_len=None, out_dtype=torch.bfloat16) -> None:
        """`compress` whose [rows, K, d_out] result is stored through an NVSwitch multicast address
        (tdc_compress_multicast): every GPU mapped behind `multicast_ptr` receives the rows.  The
        caller synchronises the group before reading (see dist.MulticastGather)."""
        if self.cfg.d_out <= 0:
            raise RuntimeError("engine was created without d_out: no vision_proj")
        self._run(self.lib.tdc_compress_multicast, "tdc_compress_multicast", query_embeds, enc, input_ids, query_set,
                  text_set, kv_len, self.cfg.d_out, lambda K, T: K, out_dtype, out_ptr=multicast_ptr)

    def compress_host(self, query_embeds: torch.Tensor, enc_host: torch.Tensor, out_host: Optional[torch.Tensor] = None,
                      *, query_set: Optional[torch.Tensor] = None, input_ids: Optional[torch.Tensor] = None,
                      text_set: Optional[torch.Tensor] = None, rows_per_batch: int = 1024,
                      out_device: Optional[torch.Tensor] = None, taper_tail: bool = True) -> torch.Tensor:
        """`compress` for inputs that live in (pinned) HOST memory: enc_host [rows, L, d_enc] is streamed
        to the GPU in row batches on a copy stream while the previous batch computes, and each batch's
        [n, K, d_out] result is copied back to `out_host` on a third stream.  Stream-ordered: the
        result is complete once the current stream is synchronised.  `out_device` [rows, K, d_out], if
        given, additionally keeps the result on the GPU (e.g. as the send buffer of the all-gather)."""
        rows, L, _ = enc_host.shape
        K = query_embeds.shape[1]
        dev = self.device
        if out_host is None:
            out_host = torch.empty((rows, K, self.cfg.d_out), dtype=enc_host.dtype, pin_memory=True)
        if rows == 0:
            return out_host
        rb = max(1, min(rows_per_batch, rows))
        cur = torch.cuda.current_stream(dev)
        if not hasattr(self, "_h2d_stream"):
            self._h2d_stream, self._d2h_stream = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        stage = [torch.empty((rb, L, self.cfg.d_enc), dtype=enc_host.dtype, device=dev) for _ in range(2)]
        h2d_done = [torch.cuda.Event() for _ in range(2)]
        compute_done = [torch.cuda.Event() for _ in range(2)]
        q_dev = query_embeds.to(dev, non_blocking=True)
        qs_dev = None if query_set is None else query_set.to(dev, torch.int32)
        ts_dev = None if text_set is None else text_set.to(dev, torch.int32)
        ids_dev = None if input_ids is None else input_ids.to(dev)
        self._h2d_stream.wait_stream(cur)
        # The stream is transfer-bound (PCIe), so what the caller waits for after the last byte has arrived is the
        # last batch's compute + read-back: taper the tail (rb, ..., rb/2, rb/4, ..., 64-128 rows) to keep that short.
        bounds, r0 = [], 0
        while rows - r0 > rb:
            bounds.append((r0, r0 + rb))
            r0 += rb
        while rows - r0 > 128 and taper_tail:
            step = (rows - r0) // 2
            bounds.append((r0, r0 + step))
            r0 += step
        if r0 < rows:
            bounds.append((r0, rows))
        for i, (r0, r1) in enumerate(bounds):
            b = i % 2
            with torch.cuda.stream(self._h2d_stream):
                if i >= 2:
                    self._h2d_stream.wait_event(compute_done[b])   # staging buffer free again
                stage[b][: r1 - r0].copy_(enc_host[r0:r1], non_blocking=True)
                h2d_done[b].record(self._h2d_stream)
            cur.wait_event(h2d_done[b])
            out_dev = self.compress(q_dev if qs_dev is not None else q_dev[r0:r1], stage[b][: r1 - r0],
                                    ids_dev if (ids_dev is None or ts_dev is not None) else ids_dev[r0:r1],
                                    query_set=None if qs_dev is None else qs_dev[r0:r1],
                                    text_set=None if ts_dev is None else ts_dev[r0:r1], out_dtype=out_host.dtype)
            if out_device is not None:
                out_device[r0:r1].copy_(out_dev)
            compute_done[b].record(cur)
            with torch.cuda.stream(self._d2h_stream):
                self._d2h_stream.wait_event(compute_done[b])
                out_host[r0:r1].copy_(out_dev, non_blocking=True)
                out_dev.record_stream(self._d2h_stream)
        cur.wait_stream(self._d2h_stream)
        for t in stage:
            t.record_stream(self._h2d_stream)
        return out_host

    # -- the upstream entry: from the towers' outputs -----------------------------------------
    def _frames_workspace(self, n_chunks, rows, Tv, Ta, K, T) -> torch.Tensor:
        most = max(rows, n_chunks, 1)
        need = int(self.lib.tdc_frames_workspace_bytes(self._h, n_chunks, rows, most, Tv, Ta, K, T))
        floor = int(self.lib.tdc_frames_workspace_bytes(self._h, n_chunks, rows, 1, Tv, Ta, K, T))
        want = max(min(need, self.max_frames_workspace_bytes), floor)
        if self._ws is None or self._ws.numel() < want:
            self._ws = None
            self._ws = torch.empty(want, dtype=torch.uint8, device=self.device)
        return self._ws

    def compress_frames(self, frames: torch.Tensor, static_frames: torch.Tensor, row_frames: torch.Tensor,
                        row_chunk: torch.Tensor, *, audio: Optional[torch.Tensor] = None,
                        input_ids: Optional[torch.Tensor] = None, num_query: int = 16, learned_queries: bool = False,
                        fold: bool = True, want_static: bool = True, out_dtype=torch.bfloat16,
                        multicast_ptr: Optional[int] = None, layer0_dedup: bool = True,
                        chunk_prompt: Optional[torch.Tensor] = None, static_multicast_ptr: Optional[int] = None,
                        static_ready_event: Optional[torch.cuda.Event] = None,
                        static_into: Optional[torch.Tensor] = None, out_into: Optional[torch.Tensor] = None):
        """The TDC stage from the towers' outputs (tdc_compress_frames): mm_projector, image_newline, audio_proj,
        query build, Q-Former, vision_proj + L2-normalise for all chunks of a video in one call.

        frames [n_frames, Tv, d_frame_in] bf16 (input of mm_projector), audio [n_frames, Ta, d_audio] bf16 or None,
        static_frames [C] / row_frames [R] / row_chunk [R] int32 (see compressor.plan_chunks).
        input_ids [1, T] (one prompt for all rows) or [P, T] with chunk_prompt [C] int32 (prompt of every chunk:
        several videos with their own questions in one call).
        `static_into` / `out_into`: contiguous destinations ([C, side*(side+1)+Ta, d] / [R, K, d], e.g. slices of a
        whole-video buffer) the library writes directly instead of fresh tensors.
        Returns (static_out [C, side*(side+1)+Ta, d] or None, compressed [R, K, d])."""
        if self.cfg.d_frame_in <= 0:
            raise RuntimeError("engine was created without d_frame_in: no upstream entry")
        dev = self.device
        if not frames.is_cuda:
            raise RuntimeError("compress_frames needs CUDA tensors: there is no CPU fallback")
        frames = frames.to(dev, torch.bfloat16).contiguous()
        n_frames, Tv, d_in = frames.shape
        if d_in != self.cfg.d_frame_in:
            raise ValueError(f"frames width {d_in} vs d_frame_in {self.cfg.d_frame_in}")
        Ta = 0
        if audio is not None:
            audio = audio.to(dev, torch.bfloat16).contiguous()
            if audio.shape[0] != n_frames or audio.shape[2] != self.cfg.d_audio:
                raise ValueError("audio must be [n_frames, Ta, d_audio]")
            Ta = int(audio.shape[1])
        sf = static_frames.to(dev, torch.int32).contiguous()
        rf = row_frames.to(dev, torch.int32).contiguous()
        rc_ = row_chunk.to(dev, torch.int32).contiguous()
        C_, R = int(sf.numel()), int(rf.numel())
        T = 0 if input_ids is None else int(input_ids.shape[-1])
        ids = None if T == 0 else input_ids.reshape(-1, T).to(dev, torch.int64).contiguous()
        cp = None
        if T > 0 and chunk_prompt is not None:
            cp = chunk_prompt.to(dev, torch.int32).contiguous()
            if cp.numel() != C_:
                raise ValueError("chunk_prompt needs one entry per chunk")
        elif T > 0 and ids.shape[0] != 1:
            raise ValueError("several prompts need chunk_prompt")
        side = int(round(Tv ** 0.5))
        d = self.cfg.d_out
        def dest(t, shape, what):
            if tuple(t.shape) != shape or t.dtype != out_dtype or t.device != dev or not t.is_contiguous():
                raise ValueError(f"{what} must be a contiguous {shape} {out_dtype} tensor on {dev}")
            return t

        static_out = None
        if static_into is not None and static_multicast_ptr is None:
            static_out = dest(static_into, (C_, side * (side + 1) + Ta, d), "static_into")
        elif want_static and static_multicast_ptr is None:
            static_out = torch.empty((C_, side * (side + 1) + Ta, d), dtype=out_dtype, device=dev)
        static_ptr = None if static_out is None else static_out.data_ptr()
        if static_multicast_ptr is not None:     # the key frames' tokens go to a caller-owned multicast mapping
            static_ptr = int(static_multicast_ptr)
        if multicast_ptr is not None:
            out = None
        elif out_into is not None:
            out = dest(out_into, (R, num_query, d), "out_into")
        else:
            out = torch.empty((R, num_query, d), dtype=out_dtype, device=dev)
        if C_ == 0 and R == 0:
            return static_out, out
        ws = self._frames_workspace(C_, R, Tv, Ta, num_query, T)
        a = TdcFramesArgs(frames.data_ptr(), None if audio is None else audio.data_ptr(), sf.data_ptr(), rf.data_ptr(),
                          rc_.data_ptr(), None if ids is None else ids.data_ptr(), n_frames, C_, R, Tv, Ta, num_query,
                          T, int(learned_queries), int(fold), int(multicast_ptr is not None), _DTYPES[out_dtype],
                          int(not layer0_dedup), static_ptr,
                          int(multicast_ptr) if multicast_ptr is not None else out.data_ptr(),
                          None if cp is None else cp.data_ptr(), 0 if ids is None else int(ids.shape[0]),
                          int(static_multicast_ptr is not None),
                          None if static_ready_event is None else _raw_event(static_ready_event, dev))
        with torch.cuda.device(dev):
            rc = self.lib.tdc_compress_frames(self._h, C.byref(a), _ptr(ws), ws.numel(), _stream(dev))
        check(rc, self._h, "tdc_compress_frames")
        return static_out, out

    def compress_frames_host(self, frames_host: torch.Tensor, audio_host: Optional[torch.Tensor], chunk_start,
                             chunk_len, out_host: Optional[torch.Tensor] = None, *, input_ids=None, num_query: int = 16,
                             learned_queries: bool = False, fold: bool = True, keep_static: bool = True,
                             static_out: Optional[torch.Tensor] = None, chunks_per_batch: int = 256,
                             out_device: Optional[torch.Tensor] = None, taper_head: bool = True) -> torch.Tensor:
        """`compress_frames` for tower outputs that live in (pinned) HOST memory, in video order: contiguous
        ranges of whole chunks are streamed to the GPU on a copy stream while the previous range computes, each
        range's compressed tokens go back to `out_host` on a third stream.  This is the end-to-end entry of the
        path: what crosses PCIe is what the towers produce (1024-wide features + 768-wide audio tokens), not the
        d_llm-wide projected tokens.

        chunk_start / chunk_len: first frame and frame count (1..8) of every chunk, ascending and contiguous per
        range (compressor.plan_chunks); the first frame of a chunk is its key frame, the others are rows.
        `static_out` [C, Ls, d] (device) receives the key frames' pass-through tokens when given."""
        import numpy as np
        dev = self.device
        cs = np.asarray(chunk_start, dtype=np.int64)
        cl = np.asarray(chunk_len, dtype=np.int64)
        Cn = len(cs)
        rows_per_chunk = cl - 1 if keep_static else cl
        row_base = np.concatenate([[0], np.cumsum(rows_per_chunk)])
        R = int(row_base[-1])
        K, d = num_query, self.cfg.d_out
        if out_host is None:
            out_host = torch.empty((R, K, d), dtype=torch.bfloat16, pin_memory=True)
        if Cn == 0:
            return out_host
        cur = torch.cuda.current_stream(dev)
        if not hasattr(self, "_h2d_stream"):
            self._h2d_stream, self._d2h_stream = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        bounds = plan_chunk_ranges(Cn, chunks_per_batch, taper_head=taper_head)
        max_frames = max(int(cs[b - 1] + cl[b - 1] - cs[a]) for a, b in bounds)
        Tv, d_in = frames_host.shape[1], frames_host.shape[2]
        stage_f = [torch.empty((max_frames, Tv, d_in), dtype=torch.bfloat16, device=dev) for _ in range(2)]
        stage_a = None
        if audio_host is not None:
            stage_a = [torch.empty((max_frames,) + tuple(audio_host.shape[1:]), dtype=torch.bfloat16, device=dev)
                       for _ in range(2)]
        h2d_done = [torch.cuda.Event() for _ in range(2)]
        compute_done = [torch.cuda.Event() for _ in range(2)]
        ids_dev = None if input_ids is None else input_ids.to(dev)
        # the integer plans of all ranges go up in ONE small copy before the pipeline starts (a pageable copy per
        # range would block the host behind that range's frame copy)
        plans = [range_plan(cs, cl, a, b, keep_static) for a, b in bounds]
        flat = np.concatenate([np.concatenate(p[2:5]) for p in plans]) if plans else np.zeros(0, np.int32)
        flat_dev = torch.from_numpy(flat.astype(np.int32)).to(dev)
        offs, o = [], 0
        for p in plans:
            n_st, n_rf = len(p[2]), len(p[3])
            offs.append((o, o + n_st, o + n_st + n_rf, o + n_st + 2 * n_rf))
            o += n_st + 2 * n_rf
        # destinations the library can write in place (same dtype, contiguous, on this device); else a copy per range
        fits = lambda t: t is not None and t.dtype == out_host.dtype and t.device == dev and t.is_contiguous()
        direct_static, direct_out = fits(static_out), fits(out_device)
        self._h2d_stream.wait_stream(cur)
        for i, (a, b) in enumerate(bounds):
            sb = i % 2
            f0, f1 = plans[i][0], plans[i][1]
            o0, o1, o2, o3 = offs[i]
            plan_dev = [flat_dev[o0:o1], flat_dev[o1:o2], flat_dev[o2:o3]]
            with torch.cuda.stream(self._h2d_stream):
                if i >= 2:
                    self._h2d_stream.wait_event(compute_done[sb])
                stage_f[sb][: f1 - f0].copy_(frames_host[f0:f1], non_blocking=True)
                if stage_a is not None:
                    stage_a[sb][: f1 - f0].copy_(audio_host[f0:f1], non_blocking=True)
                h2d_done[sb].record(self._h2d_stream)
            cur.wait_event(h2d_done[sb])
            r0, r1 = int(row_base[a]), int(row_base[b])
            s_out, comp = self.compress_frames(stage_f[sb][: f1 - f0], plan_dev[0], plan_dev[1], plan_dev[2],
                                               audio=None if stage_a is None else stage_a[sb][: f1 - f0],
                                               input_ids=ids_dev, num_query=K, learned_queries=learned_queries,
                                               fold=fold, want_static=static_out is not None,
                                               out_dtype=out_host.dtype,
                                               static_into=static_out[a:b] if direct_static else None,
                                               out_into=out_device[r0:r1] if direct_out else None)
            if static_out is not None and not direct_static:
                static_out[a:b].copy_(s_out)
            if out_device is not None and not direct_out:
                out_device[r0:r1].copy_(comp)
            compute_done[sb].record(cur)
            with torch.cuda.stream(self._d2h_stream):
                self._d2h_stream.wait_event(compute_done[sb])
                out_host[r0:r1].copy_(comp, non_blocking=True)
                comp.record_stream(self._d2h_stream)
        cur.wait_stream(self._d2h_stream)
        for t in stage_f + (stage_a or []):
            t.record_stream(self._h2d_stream)
        return out_host

    def proj_norm(self, hidden: torch.Tensor, num_query: int, out_dtype=None) -> torch.Tensor:
        rows, tokens, H = hidden.shape
        hidden = hidden.to(self.device).contiguous()
        out = torch.empty((rows, num_query, self.cfg.d_out), dtype=out_dtype or hidden.dtype, device=self.device)
        if rows == 0:
            return out
        need = rows * num_query * (H * 2 + self.cfg.d_out * 4) + 1024
        ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.tdc_proj_norm(self._h, _ptr(hidden), _dt(hidden), rows, tokens, num_query, _ptr(out),
                                        _dt(out), _ptr(ws), ws.numel(), _stream(self.device))
        check(rc, self._h, "tdc_proj_norm")
        return out

    # -- instrumentation ------------------------------------------------------------------
    def set_profiling(self, enabled: bool) -> None:
        check(self.lib.tdc_set_profiling(self._h, int(enabled)), self._h)

    def reset_profile(self) -> None:
        check(self.lib.tdc_reset_profile(self._h), self._h)

    def profile(self) -> Dict[str, Dict[str, float]]:
        out = {}
        for cls, name in enumerate(_lib.KERNEL_CLASS_NAMES):
            ms, n = C.c_double(), C.c_int64()
            check(self.lib.tdc_get_profile(self._h, cls, C.byref(ms), C.byref(n)), self._h)
            out[name] = {"ms": ms.value, "launches": int(n.value)}
        return out

    def launch_count(self) -> int:
        return int(self.lib.tdc_launch_count(self._h))


# ---- free functions (no handle) -----------------------------------------------------------
def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, *, gelu=False,
           out_dtype=torch.bfloat16, cta_group=0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = x . weight^T + bias on the tcgen05 GEMM (tdc_linear).  x [..., k], weight [n, k].
    `out`: optional preallocated contiguous [m, n] destination (e.g. a row slice of a larger buffer)."""
    lib = _lib.load_library()
    if not x.is_cuda:
        raise RuntimeError("tdc_video_b200.linear needs CUDA tensors: there is no CPU fallback")
    k = x.shape[-1]
    n = weight.shape[0]
    x2 = x.reshape(-1, k).to(torch.bfloat16).contiguous()
    w = weight.to(x.device, torch.bfloat16).contiguous()
    b = None if bias is None else bias.to(x.device, torch.float32).contiguous()
    if out is not None:
        if out.shape != (x2.shape[0], n) or not out.is_contiguous() or out.device != x.device:
            raise ValueError("linear(out=...): need a contiguous [m, n] tensor on the input's device")
        y, out_dtype = out, out.dtype
    else:
        y = torch.empty((x2.shape[0], n), dtype=out_dtype, device=x.device)
    if x2.shape[0] > 0:
        with torch.cuda.device(x.device):
            rc = lib.tdc_linear(_ptr(x2), _ptr(w), _ptr(b), _ptr(y), x2.shape[0], n, k, _DTYPES[out_dtype], int(gelu),
                                cta_group, _stream(x.device))
        check(rc, None, "tdc_linear")
    return y if out is not None else y.reshape(*x.shape[:-1], n)


def avg_pool_tokens(frames: torch.Tensor, num_query: int) -> torch.Tensor:
    """[n, tokens, d] -> [n, K, d] bf16, adaptive average pooling over tokens (tdc_avg_pool_tokens)."""
    lib = _lib.load_library()
    if not frames.is_cuda:
        raise RuntimeError("tdc_video_b200.avg_pool_tokens needs CUDA tensors: there is no CPU fallback")
    frames = frames.contiguous()
    n, tokens, d = frames.shape
    out = torch.empty((n, num_query, d), dtype=torch.bfloat16, device=frames.device)
    with torch.cuda.device(frames.device):
        rc = lib.tdc_avg_pool_tokens(_ptr(frames), _dt(frames), n, tokens, d, num_query, _ptr(out),
                                     _stream(frames.device))
    check(rc, None, "tdc_avg_pool_tokens")
    return out
