"""The TDC driver: everything `prepare_inputs_labels_for_multimodal` does between "frames of one
video are projected" and "compressed token sequence is spliced into the prompt"
(tdc/cambrian_arch.py:1507-1709), batched over ALL chunks of the video instead of the
reference's Python double loop with <= 7 rows per Q-Former call.

`TDCCompressor` owns the same attributes the reference keeps on the model
(cambrian_arch.py:148-150, 180-181, 469-484): `Qformer`, `query_tokens`, `vision_proj`,
`query_proj`, `frame_seg`, optional `audio_proj` — same names, same shapes, so the matching
entries of a reference checkpoint (`model.<name>`) load unchanged.

Data flow per video (n_frames frames already split into DINO segments):
  host ints : segments -> chunks of <= 8 frames -> static frame + dynamic rows  (plan_chunks)
  GPU       : audio_proj GEMM  ->  KV tokens [R, Lv(+La), d]
              avg-pool(static) -> query_proj GEMM -> one query set per chunk    (Avg_pool)
              libtdc tdc_compress: Q-Former + vision_proj + L2-normalise for all R rows
              scatter static / compressed / frame_seg tokens to their final offsets
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch
from torch import nn

from . import dist as tdist
from .engine import avg_pool_tokens, linear
from .qformer import QFormerConfig, TDCQFormer

CHUNK_FRAMES = 8  # cambrian_arch.py:1606


@dataclass
class ChunkPlan:
    """Host-side index plan of one video (pure integers, no device work)."""
    static_frames: np.ndarray   # [C]   frame index of every chunk's key frame
    chunk_len: np.ndarray       # [C]   frames in the chunk (1..8)
    row_frames: np.ndarray      # [R]   frame index of every Q-Former row
    row_chunk: np.ndarray       # [R]   chunk index of every row
    rows_per_chunk: np.ndarray  # [C]

    @property
    def num_chunks(self) -> int:
        return int(self.static_frames.shape[0])

    @property
    def num_rows(self) -> int:
        return int(self.row_frames.shape[0])


def plan_chunks(segment_sizes: Sequence[int], keep_static: bool = True) -> ChunkPlan:
    """segments -> chunks of <= 8 frames (cambrian_arch.py:1603-1628).  With `keep_static` the
    chunk's first frame passes through uncompressed and the others become rows; without it
    every frame of the chunk is a row (and the first frame still provides the queries)."""
    static, clen, rf, rc, rpc = [], [], [], [], []
    base = 0
    for n in segment_sizes:
        n = int(n)
        for start in range(0, n, CHUNK_FRAMES):
            ln = min(CHUNK_FRAMES, n - start)
            c = len(static)
            static.append(base + start)
            clen.append(ln)
            frames = range(base + start + 1, base + start + ln) if keep_static else range(base + start,
                                                                                          base + start + ln)
            rf.extend(frames)
            rc.extend([c] * len(frames))
            rpc.append(len(frames))
        base += n
    i64 = np.int64
    return ChunkPlan(np.asarray(static, i64), np.asarray(clen, i64), np.asarray(rf, i64), np.asarray(rc, i64),
                     np.asarray(rpc, i64))


def output_layout(plan: ChunkPlan, static_tokens: int, num_query: int, keep_static: bool = True, add_sep: bool = True):
    """Token offsets of the assembled sequence (cambrian_arch.py:1617-1623, 1668-1692).
    Returns (chunk_offset [C], chunk_tokens [C], row_offset [R])."""
    sep = 1 if add_sep else 0
    head = (static_tokens + sep) if keep_static else 0
    chunk_tokens = head + plan.rows_per_chunk * (num_query + sep)
    chunk_offset = np.concatenate([[0], np.cumsum(chunk_tokens)[:-1]]) if plan.num_chunks else np.zeros(0, np.int64)
    first_row = np.concatenate([[0], np.cumsum(plan.rows_per_chunk)[:-1]]) if plan.num_chunks else np.zeros(0, np.int64)
    within = np.arange(plan.num_rows) - first_row[plan.row_chunk] if plan.num_rows else np.zeros(0, np.int64)
    row_offset = chunk_offset[plan.row_chunk] + head + within * (num_query + sep) if plan.num_rows else np.zeros(0, np.int64)
    return chunk_offset.astype(np.int64), chunk_tokens.astype(np.int64), row_offset.astype(np.int64)


def truncation_keep_index(chunk_offset: np.ndarray, chunk_tokens: np.ndarray, max_visual_len: int) -> Optional[np.ndarray]:
    """Budget truncation (cambrian_arch.py:1694-1709): when over budget drop the last
    ceil(excess / num_chunks) tokens of EVERY chunk, then hard-clip.  Returns the kept token
    indices, or None when nothing is dropped."""
    total = int(chunk_tokens.sum())
    if max_visual_len is None or total <= max_visual_len:
        return None
    force_remove = math.ceil((total - max_visual_len) / len(chunk_tokens))
    keep = []
    for off, n in zip(chunk_offset.tolist(), chunk_tokens.tolist()):
        kept = max(n - force_remove, 0) if force_remove > 0 else n   # x[:-k] semantics
        keep.append(np.arange(off, off + kept, dtype=np.int64))
    idx = np.concatenate(keep) if keep else np.zeros(0, np.int64)
    return idx[:max_visual_len]


class TDCCompressor(nn.Module):
    """`initialize_compressor` (cambrian_arch.py:469-484) + the TDC block (:1507-1709)."""

    def __init__(self, llm_hidden_size: int, context_token_num: int = 16, query_type: str = "Avg_pool",
                 text_input: bool = True, add_static: bool = True, audio_input: bool = False,
                 qformer_config: Optional[QFormerConfig] = None, with_lm_head: bool = False,
                 mm_input_size: int = 0):
        super().__init__()
        if query_type not in ("Avg_pool", "learned"):
            raise ValueError("query_type must be 'Avg_pool' or 'learned' (cambrian_arch.py:1633-1640)")
        cfg = qformer_config or QFormerConfig()
        cfg.encoder_width = llm_hidden_size          # encoder_width = config.hidden_size (:470, :408)
        cfg.query_length = context_token_num
        self.llm_hidden_size = llm_hidden_size
        self.context_token_num = context_token_num
        self.query_type = query_type
        self.text_input = text_input
        self.add_static = add_static
        self.Qformer = TDCQFormer(cfg, with_lm_head=with_lm_head)
        self.query_tokens = nn.Parameter(torch.zeros(1, context_token_num, cfg.hidden_size))
        self.query_tokens.data.normal_(mean=0.0, std=cfg.initializer_range)
        self.vision_proj = nn.Linear(cfg.hidden_size, llm_hidden_size)
        self.query_proj = nn.Linear(llm_hidden_size, cfg.hidden_size)
        self.frame_seg = nn.Parameter(torch.randn(llm_hidden_size))
        if audio_input:
            self.audio_proj = nn.Linear(768, llm_hidden_size)
        if mm_input_size:
            # the upstream entry also owns the projector and the newline vector, under the reference's names
            # (cambrian_arch.py:65-69 `mm_projector`, :148 `image_newline`)
            from .projector import GeluMLPProjector
            self.mm_projector = GeluMLPProjector(mm_input_size, llm_hidden_size)
            self.image_newline = nn.Parameter(torch.randn(llm_hidden_size) * (1.0 / math.sqrt(llm_hidden_size)))
        self.mm_input_size = mm_input_size
        self.eval()

    # ------------------------------------------------------------------------------------
    def _engine(self):
        if self.mm_input_size:
            return self._frames_engine()   # a superset handle: no rebuild when both entries are used
        extra = {"vision_proj.weight": self.vision_proj.weight, "vision_proj.bias": self.vision_proj.bias}
        key = (self.vision_proj.weight.data_ptr(), self.vision_proj.weight._version, self.vision_proj.bias._version)
        return self.Qformer.bert.engine(d_out=self.llm_hidden_size, extra_state=extra, extra_key=key)

    def _frames_engine(self):
        """One handle with everything the upstream entry needs (tdc_compress_frames)."""
        if not self.mm_input_size:
            raise RuntimeError("TDCCompressor was built without mm_input_size: no mm_projector / image_newline")
        l0, l2 = getattr(self.mm_projector, "0"), getattr(self.mm_projector, "2")
        extra = {"vision_proj.weight": self.vision_proj.weight, "vision_proj.bias": self.vision_proj.bias,
                 "mm_projector.0.weight": l0.weight, "mm_projector.0.bias": l0.bias,
                 "mm_projector.2.weight": l2.weight, "mm_projector.2.bias": l2.bias,
                 "image_newline": self.image_newline, "query_proj.weight": self.query_proj.weight,
                 "query_proj.bias": self.query_proj.bias, "query_tokens": self.query_tokens}
        d_audio = 0
        if hasattr(self, "audio_proj"):
            extra.update({"audio_proj.weight": self.audio_proj.weight, "audio_proj.bias": self.audio_proj.bias})
            d_audio = self.audio_proj.in_features
        key = ("frames",) + tuple((t.data_ptr(), t._version) for t in extra.values())
        return self.Qformer.bert.engine(d_out=self.llm_hidden_size, extra_state=extra, extra_key=key,
                                        d_frame_in=self.mm_input_size, d_audio=d_audio)

    @torch.no_grad()
    def compress_video_from_towers(self, tower_features: torch.Tensor, segment_sizes: Sequence[int],
                                   input_ids: Optional[torch.Tensor] = None,
                                   audio_frames: Optional[torch.Tensor] = None,
                                   max_visual_len: Optional[int] = None, fold: bool = True):
        """`compress_video` from the towers' outputs: tower_features [n_frames, Tv, mm_input_size] is what the
        reference feeds to `mm_projector` (cambrian_arch.py:1149).  The projector, the newline tokens, audio_proj,
        the query build and the Q-Former all run inside ONE library call (tdc_compress_frames); with `fold` the
        dynamic frames never materialise their d_llm-wide tokens.  Returns the same token sequence as
        `compress_video(append_newline(mm_projector(tower_features)), ...)`."""
        if self.training:
            raise RuntimeError("TDCCompressor is inference-only (eval mode)")
        if not tower_features.is_cuda:
            raise RuntimeError("TDCCompressor needs CUDA tensors: there is no CPU fallback")
        if not self.add_static:
            raise NotImplementedError("compress_video_from_towers keeps the key frames (add_static=True)")
        n_frames, Tv, _ = tower_features.shape
        if sum(int(s) for s in segment_sizes) != n_frames:
            raise ValueError("segment_sizes must sum to the number of frames")
        dev = tower_features.device
        dtype = tower_features.dtype if tower_features.dtype in (torch.bfloat16, torch.float16) else torch.float32
        plan = plan_chunks(segment_sizes, True)
        ids = None
        if self.text_input and input_ids is not None and input_ids.numel() > 0:
            ids = input_ids.reshape(1, -1)
        if audio_frames is not None and not hasattr(self, "audio_proj"):
            raise RuntimeError("audio_frames given but the compressor was built with audio_input=False")
        static_tok, comp = self._frames_engine().compress_frames(
            tower_features, torch.from_numpy(plan.static_frames.astype(np.int32)),
            torch.from_numpy(plan.row_frames.astype(np.int32)), torch.from_numpy(plan.row_chunk.astype(np.int32)),
            audio=audio_frames, input_ids=ids, num_query=self.context_token_num,
            learned_queries=self.query_type == "learned", fold=fold, want_static=True, out_dtype=dtype)
        prep = dict(plan=plan, static_tok=static_tok, L=static_tok.shape[1], d=self.llm_hidden_size, dev=dev,
                    dtype=dtype)
        return self._assemble(prep, comp, max_visual_len)

    @torch.no_grad()
    def compress_videos_from_towers(self, videos: Sequence[dict], fold: bool = True) -> List[torch.Tensor]:
        """Several videos through ONE tdc_compress_frames call (the eval loops run many clips concurrently; BASELINE
        config 5): every item is a dict with the arguments of `compress_video_from_towers` — `tower_features`,
        `segment_sizes`, optional `input_ids` (each video its own question), `audio_frames`, `max_visual_len`.
        Videos must agree on having a prompt / audio and on the prompt length (group them otherwise).
        Returns one token sequence per video, equal bit for bit to the per-video calls."""
        if not videos:
            return []
        if self.training:
            raise RuntimeError("TDCCompressor is inference-only (eval mode)")
        if not self.add_static:
            raise NotImplementedError("compress_videos_from_towers keeps the key frames (add_static=True)")
        plans, frame_base, chunk_base = [], [0], [0]
        for v in videos:
            if sum(int(x) for x in v["segment_sizes"]) != v["tower_features"].shape[0]:
                raise ValueError("segment_sizes must sum to the number of frames")
            plans.append(plan_chunks(v["segment_sizes"], True))
            frame_base.append(frame_base[-1] + v["tower_features"].shape[0])
            chunk_base.append(chunk_base[-1] + plans[-1].num_chunks)
        use_text = self.text_input and videos[0].get("input_ids") is not None and videos[0]["input_ids"].numel() > 0
        use_audio = videos[0].get("audio_frames") is not None
        for v in videos:
            if (v.get("audio_frames") is not None) != use_audio:
                raise ValueError("all videos of one call must agree on audio")
            if use_text and (v.get("input_ids") is None or v["input_ids"].numel() != videos[0]["input_ids"].numel()):
                raise ValueError("all videos of one call must have prompts of the same length")
        t0 = videos[0]["tower_features"]
        dev = t0.device
        dtype = t0.dtype if t0.dtype in (torch.bfloat16, torch.float16) else torch.float32
        frames = torch.cat([v["tower_features"] for v in videos], dim=0)
        audio = torch.cat([v["audio_frames"].to(dev) for v in videos], dim=0) if use_audio else None
        i32 = lambda arrs: torch.from_numpy(np.concatenate(arrs).astype(np.int32))
        sf = i32([p.static_frames + fb for p, fb in zip(plans, frame_base)])
        rf = i32([p.row_frames + fb for p, fb in zip(plans, frame_base)])
        rc = i32([p.row_chunk + cb for p, cb in zip(plans, chunk_base)])
        ids = torch.cat([v["input_ids"].reshape(1, -1) for v in videos], dim=0) if use_text else None
        cp = i32([np.full(p.num_chunks, i) for i, p in enumerate(plans)]) if use_text else None
        static_tok, comp = self._frames_engine().compress_frames(
            frames, sf, rf, rc, audio=audio, input_ids=ids, num_query=self.context_token_num,
            learned_queries=self.query_type == "learned", fold=fold, want_static=True, out_dtype=dtype, chunk_prompt=cp)
        outs, r0 = [], 0
        for i, (p, v) in enumerate(zip(plans, videos)):
            prep = dict(plan=p, static_tok=static_tok[chunk_base[i]:chunk_base[i + 1]], L=static_tok.shape[1],
                        d=self.llm_hidden_size, dev=dev, dtype=dtype)
            outs.append(self._assemble(prep, comp[r0:r0 + p.num_rows], v.get("max_visual_len")))
            r0 += p.num_rows
        return outs

    def build_queries(self, static_visual: torch.Tensor):
        """[C, Lv, d] visual-only key frames -> query sets [C, K, hidden] fp32 (cambrian_arch.py:1629-1640).
        The key frame is taken BEFORE the audio tokens are appended (:1609 vs :1614)."""
        if self.query_type == "learned":
            return self.query_tokens.detach().float(), True
        pooled = avg_pool_tokens(static_visual, self.context_token_num)            # [C, K, d] bf16
        q = linear(pooled, self.query_proj.weight, self.query_proj.bias, out_dtype=torch.float32)
        return q, False

    # ---- one video = prepare (KV tokens, queries) -> Q-Former rows -> assemble --------------------
    def _prepare(self, visual_emb_frame, segment_sizes, input_ids, audio_frames):
        if self.training:
            raise RuntimeError("TDCCompressor is inference-only (eval mode)")
        if not visual_emb_frame.is_cuda:
            raise RuntimeError("TDCCompressor needs CUDA tensors: there is no CPU fallback")
        n_frames, Lv, d = visual_emb_frame.shape
        if sum(int(s) for s in segment_sizes) != n_frames:
            raise ValueError("segment_sizes must sum to the number of frames")
        dev, dtype = visual_emb_frame.device, visual_emb_frame.dtype
        plan = plan_chunks(segment_sizes, self.add_static)
        R = plan.num_rows
        static_idx = torch.from_numpy(plan.static_frames).to(dev)
        row_idx = torch.from_numpy(plan.row_frames).to(dev)

        # --- KV tokens of every row: the frame's visual tokens (+ projected audio tokens, :1611-1614)
        La = 0
        audio_tok = None
        if audio_frames is not None:
            if not hasattr(self, "audio_proj"):
                raise RuntimeError("audio_frames given but the compressor was built with audio_input=False")
            La = audio_frames.shape[1]
            audio_tok = linear(audio_frames.to(dev), self.audio_proj.weight, self.audio_proj.bias,
                               out_dtype=torch.bfloat16).to(dtype)                 # [n_frames, La, d]
        L = Lv + La
        if La:
            enc = torch.empty((R, L, d), dtype=dtype, device=dev)
            enc[:, :Lv].copy_(visual_emb_frame.index_select(0, row_idx))
            enc[:, Lv:].copy_(audio_tok.index_select(0, row_idx))
            static_tok = torch.cat([visual_emb_frame.index_select(0, static_idx),
                                    audio_tok.index_select(0, static_idx)], dim=1)  # [C, L, d]
        else:
            enc = visual_emb_frame.index_select(0, row_idx)
            static_tok = visual_emb_frame.index_select(0, static_idx)
        static_visual = static_tok[:, :Lv] if La else static_tok

        # --- queries (one set per chunk) and prompt ids (one set per video)
        q_sets = query_set = ids = None
        if R > 0:
            q_sets, shared = self.build_queries(static_visual.contiguous())
            query_set = torch.zeros(R, dtype=torch.int32) if shared else torch.from_numpy(plan.row_chunk.astype(np.int32))
            if self.text_input and input_ids is not None and input_ids.numel() > 0:
                ids = input_ids.reshape(1, -1)
        return dict(plan=plan, enc=enc, static_tok=static_tok, q_sets=q_sets, query_set=query_set, ids=ids, L=L, d=d,
                    dev=dev, dtype=dtype)

    def _assemble(self, prep, comp, max_visual_len):
        """[static, sep, (K compressed, sep) x rows] per chunk (:1668-1692), then the budget truncation (:1694-1709)."""
        plan, L, d, dev, dtype = prep["plan"], prep["L"], prep["d"], prep["dev"], prep["dtype"]
        K = self.context_token_num
        chunk_off, chunk_tok, row_off = output_layout(plan, L, K, self.add_static, add_sep=True)
        total = int(chunk_tok.sum())
        out = torch.empty((total, d), dtype=dtype, device=dev)
        seg = self.frame_seg.detach().to(dtype)
        sep_pos = []
        if self.add_static and plan.num_chunks > 0:
            idx = (chunk_off[:, None] + np.arange(L)[None, :]).reshape(-1)
            out.index_copy_(0, torch.from_numpy(idx).to(dev), prep["static_tok"].reshape(-1, d))
            sep_pos.append(chunk_off + L)
        if plan.num_rows > 0:
            idx = (row_off[:, None] + np.arange(K)[None, :]).reshape(-1)
            out.index_copy_(0, torch.from_numpy(idx).to(dev), comp.reshape(-1, d))
            sep_pos.append(row_off + K)
        if sep_pos:
            sp = torch.from_numpy(np.concatenate(sep_pos)).to(dev)
            out.index_copy_(0, sp, seg[None, :].expand(sp.numel(), d).contiguous())
        keep = truncation_keep_index(chunk_off, chunk_tok, max_visual_len)
        if keep is not None:
            out = out.index_select(0, torch.from_numpy(keep).to(dev))
        return out

    @torch.no_grad()
    def compress_video(self, visual_emb_frame: torch.Tensor, segment_sizes: Sequence[int],
                       input_ids: Optional[torch.Tensor] = None, audio_frames: Optional[torch.Tensor] = None,
                       max_visual_len: Optional[int] = None, return_parts: bool = False, group=None,
                       shard: bool = False):
        """visual_emb_frame [n_frames, Lv, d] (one video), segment_sizes (frames per DINO segment),
        input_ids [1, T] BERT ids of the prompt (used iff text_input), audio_frames [n_frames, La, 768]
        per-frame BEATs tokens (iff the video has audio).  Returns the token sequence
        `new_visual_emb_frames[:max_visual_len]` of cambrian_arch.py:1694-1709.

        shard=True (inside an initialised torch.distributed job): every rank holds the same inputs,
        compresses a contiguous range of chunks and one all-gather assembles all rows on every rank."""
        prep = self._prepare(visual_emb_frame, segment_sizes, input_ids, audio_frames)
        plan, enc, dtype = prep["plan"], prep["enc"], prep["dtype"]
        R = plan.num_rows
        comp = None
        if R > 0:
            q_sets, query_set, ids = prep["q_sets"], prep["query_set"], prep["ids"]
            text_set = None if ids is None else torch.zeros(R, dtype=torch.int32)
            sharded = shard and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size(group) > 1
            if not sharded:
                comp = self._engine().compress(q_sets, enc, ids, query_set=query_set, text_set=text_set,
                                               out_dtype=dtype)
            else:
                world, rank = torch.distributed.get_world_size(group), torch.distributed.get_rank(group)
                ranges = tdist.shard_chunk_ranges(plan.rows_per_chunk, world)
                row_ranges = [tdist.row_range_of_chunks(plan.rows_per_chunk, lo, hi) for lo, hi in ranges]
                r_lo, r_hi = row_ranges[rank]
                local = self._engine().compress(q_sets, enc[r_lo:r_hi], ids, query_set=query_set[r_lo:r_hi],
                                                text_set=None if text_set is None else text_set[r_lo:r_hi],
                                                out_dtype=dtype)
                comp = tdist.all_gather_rows(local, [hi - lo for lo, hi in row_ranges], group)
        out = self._assemble(prep, comp, max_visual_len)
        if return_parts:
            return out, comp, plan
        return out

    @torch.no_grad()
    def compress_videos(self, videos: Sequence[dict]) -> List[torch.Tensor]:
        """Several videos at once (the eval-style workload: many concurrent clips): the rows of all videos that
        share a KV length and a prompt length go through ONE tdc_compress call, instead of one call per video (and
        one per <= 7 rows in the reference).  Every item is a dict with the arguments of `compress_video`:
        `visual_emb_frame`, `segment_sizes`, optional `input_ids`, `audio_frames`, `max_visual_len`.
        Returns one token sequence per video, equal bit for bit to calling `compress_video` per video."""
        preps = [self._prepare(v["visual_emb_frame"], v["segment_sizes"], v.get("input_ids"), v.get("audio_frames"))
                 for v in videos]
        comps: List[Optional[torch.Tensor]] = [None] * len(videos)
        groups = {}
        for i, p in enumerate(preps):
            if p["plan"].num_rows > 0:
                T = 0 if p["ids"] is None else int(p["ids"].shape[1])
                groups.setdefault((p["L"], T, p["dtype"], p["q_sets"].shape[0] == 1 and self.query_type == "learned"),
                                  []).append(i)
        for (L, T, dtype, shared), members in groups.items():
            encs, qsets, qmaps, tmaps, idss = [], [], [], [], []
            set_base = 0
            for g, i in enumerate(members):
                p = preps[i]
                R = p["plan"].num_rows
                encs.append(p["enc"])
                if shared:                       # learned queries: one set serves every row of every video
                    qmaps.append(torch.zeros(R, dtype=torch.int32))
                else:
                    qsets.append(p["q_sets"])
                    qmaps.append(p["query_set"] + set_base)
                    set_base += p["q_sets"].shape[0]
                if T > 0:
                    idss.append(p["ids"])
                    tmaps.append(torch.full((R,), g, dtype=torch.int32))
            q_all = preps[members[0]]["q_sets"] if shared else torch.cat(qsets, dim=0)
            out = self._engine().compress(q_all, torch.cat(encs, dim=0) if len(encs) > 1 else encs[0],
                                          torch.cat(idss, dim=0) if T > 0 else None, query_set=torch.cat(qmaps),
                                          text_set=torch.cat(tmaps) if T > 0 else None, out_dtype=dtype)
            r0 = 0
            for i in members:
                R = preps[i]["plan"].num_rows
                comps[i] = out[r0:r0 + R]
                r0 += R
        return [self._assemble(p, c, v.get("max_visual_len")) for p, c, v in zip(preps, comps, videos)]
