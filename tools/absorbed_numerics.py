"""CPU experiment (no GPU): would the weight-absorbed cross-attention of DESIGN §7 item 6 stay inside the parity
tolerance?  Emulates the CUDA path's roundings (bf16 tensor-core operands, fp32 accumulate / LayerNorm / softmax)
in torch for (a) the shipped formulation  K = X Wk^T, V = X Wv^T  and (b) the absorbed one
S_h = (Q_h Wk_h) X^T,  O_h = (P_h X) Wv_h^T + bv_h  with the two [K*heads, d_enc] intermediates rounded to bf16,
and compares both with the fp32 oracle on the full-size stress case.   python tools/absorbed_numerics.py"""
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import qformer_oracle as oracle  # noqa: E402
from tdc_video_b200.synth import QFormerGeometry, make_inputs, make_state_dict  # noqa: E402

bf = lambda x: x.bfloat16().float()


def lin(sd, p, x):      # bf16 operands, fp32 accumulate, fp32 bias
    return F.linear(bf(x), bf(sd[p + ".weight"]), sd[p + ".bias"])


def ln(sd, p, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def attn(q, k, v, heads):
    B, n, H = q.shape
    dh = H // heads
    qh = bf(q).view(B, n, heads, dh).permute(0, 2, 1, 3)
    kh = bf(k).view(B, -1, heads, dh).permute(0, 2, 1, 3)
    vh = bf(v).view(B, -1, heads, dh).permute(0, 2, 1, 3)
    p = bf(torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(dh), -1))
    return (p @ vh).permute(0, 2, 1, 3).reshape(B, n, H)


def cross_absorbed(sd, p, xq, enc, heads):
    B, K, H = xq.shape
    dh = H // heads
    d = enc.shape[-1]
    q = bf(lin(sd, p + ".query", xq)).view(B, K, heads, dh).permute(0, 2, 1, 3)            # [B, h, K, dh]
    wk = bf(sd[p + ".key.weight"]).view(heads, dh, d)
    wv = bf(sd[p + ".value.weight"]).view(heads, dh, d)
    qt = bf(torch.einsum("bhkd,hde->bhke", q, wk))                                          # Q_h Wk_h  [B, h, K, d_enc]
    x = bf(enc)
    s = torch.einsum("bhke,ble->bhkl", qt, x) / math.sqrt(dh)                               # key bias: softmax-invariant
    pr = bf(torch.softmax(s, -1))
    y = bf(torch.einsum("bhkl,ble->bhke", pr, x))                                           # P_h X     [B, h, K, d_enc]
    o = torch.einsum("bhke,hde->bhkd", y, wv) + sd[p + ".value.bias"].view(1, heads, 1, dh)
    return o.permute(0, 2, 1, 3).reshape(B, K, H)


def forward(sd, geom, q, enc, ids, absorbed):
    eps = geom.ln_eps
    B, K, H = q.shape
    T = 0 if ids is None else ids.shape[1]
    x = q
    if T:
        text = F.embedding(ids, sd["embeddings.word_embeddings.weight"]) + sd["embeddings.position_embeddings.weight"][:T][None]
        x = torch.cat([q, text], 1)
    x = ln(sd, "embeddings.LayerNorm", x, eps)
    for l in range(geom.layers):
        p = f"encoder.layer.{l}."
        ctx = attn(lin(sd, p + "attention.self.query", x), lin(sd, p + "attention.self.key", x),
                   lin(sd, p + "attention.self.value", x), geom.heads)
        x = ln(sd, p + "attention.output.LayerNorm", lin(sd, p + "attention.output.dense", ctx) + x, eps)
        xq = x[:, :K]
        if l % geom.cross_freq == 0:
            c = p + "crossattention.self"
            if absorbed:
                ctx = cross_absorbed(sd, c, xq, enc, geom.heads)
            else:
                ctx = attn(lin(sd, c + ".query", xq), lin(sd, c + ".key", enc), lin(sd, c + ".value", enc), geom.heads)
            xq = ln(sd, p + "crossattention.output.LayerNorm", lin(sd, p + "crossattention.output.dense", ctx) + xq, eps)
        mid = F.gelu(lin(sd, p + "intermediate_query.dense", xq))
        xq = ln(sd, p + "output_query.LayerNorm", lin(sd, p + "output_query.dense", mid) + xq, eps)
        if T:
            xt = x[:, K:]
            mid = F.gelu(lin(sd, p + "intermediate.dense", xt))
            xt = ln(sd, p + "output.LayerNorm", lin(sd, p + "output.dense", mid) + xt, eps)
            x = torch.cat([xq, xt], 1)
        else:
            x = xq
    return x


def main():
    torch.set_num_threads(os.cpu_count())
    geom = QFormerGeometry(d_enc=3584, d_out=3584)
    for seed, stress, T in ((22, 2.0, 6), (23, 2.0, 0), (24, 0.0, 0)):
        sd_np = make_state_dict(geom, seed, stress=stress)
        inp = make_inputs(geom, seed, 3, 206, 16, T, audio_tokens=50)
        sd = {k: torch.from_numpy(v) for k, v in sd_np.items()}
        q, enc = torch.from_numpy(inp["query_embeds"]), bf(torch.from_numpy(inp["enc"]))
        ids = torch.from_numpy(inp["input_ids"]) if T else None
        with torch.no_grad():
            ref = oracle.qformer_forward(sd, geom, q, enc, ids)
            for name, flag in (("shipped ", False), ("absorbed", True)):
                out = forward(sd, geom, q, enc, ids, flag)
                m = oracle.parity_metrics(out, ref)
                mc = oracle.parity_metrics(oracle.proj_norm(sd, out, 16), oracle.proj_norm(sd, ref, 16))
                print(f"seed {seed} stress {stress} T {T}  {name}: hidden cos {m['min_cos']:.6f} err {m['max_abs_over_max_ref']:.2e}"
                      f" | compressed cos {mc['min_cos']:.6f} err {mc['max_abs_over_max_ref']:.2e}", flush=True)


if __name__ == "__main__":
    main()
