"""TEST INFRASTRUCTURE ONLY — CPU (fp32 torch) restatement of the TDC stage FROM THE TOWERS' OUTPUTS, i.e. the
oracle of the library's upstream entry `tdc_compress_frames`.  Imported only by tests/, `__graft_entry__.smoke()`
and bench.py's `cpu_baseline` / `--impl reference` legs — never by the product.

Reference order (tdc/cambrian_arch.py):
    :1149-1150   image_features = mm_projector(cat(tower features))          Linear -> GELU(erf) -> Linear (:65-69)
    :1269-1281   image_newline appended to every row of the 12 x 12 token grid  -> [n, 156, d]
    :1611-1614   audio_proj(audio tokens) concatenated to every frame of the chunk -> [n, 206, d]
    :1609, 1629-1640   key frame = first frame of the chunk, VISUAL tokens only; queries =
                 query_proj(adaptive_avg_pool1d(key frame)) (or the learned query_tokens)
    :1653-1667   Qformer.bert(...) on the other frames -> F.normalize(vision_proj(h[:, :K]))
    :1617-1623, 1668-1692   the key frame (with its audio tokens) passes through

Pinning: the composition (make_golden.driver_frames with the GELU-MLP projector + driver_oracle.compress_video)
is checked against committed outputs of the reference's real prepare_inputs_labels_for_multimodal in
tests/test_towers_golden.py; `frames_stage` below is the same arithmetic with all rows of all chunks batched into
one Q-Former call (rows are independent), checked against that composition in tests/test_frames_oracle.py.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn.functional as F

from . import qformer_oracle as qo


def project_frames(sd, tower_features, side: Optional[int] = None) -> torch.Tensor:
    """mm_projector + newline tokens: [n, side*side, d_in] -> [n, side*(side+1), d]  (:1149-1150, :1269-1281)."""
    t = qo._t
    x = t(tower_features)
    xv = qo.gelu_mlp(sd["mm_projector.0.weight"], sd["mm_projector.0.bias"], sd["mm_projector.2.weight"],
                     sd["mm_projector.2.bias"], x)
    n, tokens, d = xv.shape
    side = side or int(round(tokens ** 0.5))
    nl = t(sd["image_newline"]).view(1, 1, 1, -1).expand(n, side, 1, -1)
    return torch.cat([xv.view(n, side, side, d), nl], dim=2).flatten(1, 2)


def frames_stage(sd, geom, tower_features, audio, chunk_start: Sequence[int], chunk_len: Sequence[int],
                 num_query: int, input_ids=None, learned_queries: bool = False):
    """-> (static [C, side*(side+1) + Ta, d]: the key frames as they pass through,
           compressed [R, K, d]: one row per non-key frame, chunk after chunk)."""
    t = qo._t
    fr = project_frames(sd, tower_features)                                     # [n, 156, d]
    full = fr
    if audio is not None:
        full = torch.cat([fr, F.linear(t(audio), t(sd["audio_proj.weight"]), t(sd["audio_proj.bias"]))], dim=1)
    statics, rows, queries = [], [], []
    for c0, ln in zip(chunk_start, chunk_len):
        c0, ln = int(c0), int(ln)
        statics.append(full[c0])
        if ln > 1:
            rows.append(full[c0 + 1:c0 + ln])
            if learned_queries:
                q = t(sd["query_tokens"]).reshape(1, num_query, -1)
            else:
                q = qo.avg_pool_queries(fr[c0][None], num_query)                # key frame: visual + newline only
                q = F.linear(q, t(sd["query_proj.weight"]), t(sd["query_proj.bias"]))
            queries.append(q.expand(ln - 1, -1, -1))
    static = torch.stack(statics) if statics else full[:0]
    if not rows:
        return static, full.new_zeros((0, num_query, full.shape[-1]))
    enc, q = torch.cat(rows), torch.cat(queries)
    ids = None
    if input_ids is not None:
        ids = torch.as_tensor(input_ids, dtype=torch.long).reshape(1, -1).expand(enc.shape[0], -1)
    return static, qo.compress(sd, geom, q, enc, ids)
