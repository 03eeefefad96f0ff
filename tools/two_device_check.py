"""One process, two GPUs: an engine on cuda:0 and an engine on cuda:1 (the per-device launch state of the
library: shared-memory opt-in, SM count, co-resident cluster count) must both work and agree bit for bit.

    python tools/two_device_check.py          (needs >= 2 visible GPUs)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tdc_video_b200 import QFormerEngine  # noqa: E402
from tdc_video_b200.synth import QFormerGeometry, make_inputs, make_state_dict  # noqa: E402


def main():
    if torch.cuda.device_count() < 2:
        print("two_device_check: needs 2 GPUs, skipped")
        return 0
    geom = QFormerGeometry(d_enc=3584, d_out=3584, vocab=0)   # full geometry: the 230 KB smem GEMM variants
    sd = make_state_dict(geom, 5, with_text=False)
    inp = make_inputs(geom, 6, rows=24, kv_tokens=206, num_query=16, audio_tokens=50)
    outs = []
    # cuda:1 FIRST, then cuda:0, then cuda:1 again: whichever device comes second used to miss its opt-in
    for dev in ("cuda:1", "cuda:0", "cuda:1"):
        eng = QFormerEngine(d_enc=3584, d_out=3584, vocab=0, device=dev)
        eng.load_weights(sd)
        q = torch.from_numpy(inp["query_embeds"]).to(dev)
        enc = torch.from_numpy(inp["enc"]).to(dev).bfloat16()
        out = eng.compress(q, enc, out_dtype=torch.bfloat16)
        torch.cuda.synchronize(dev)
        outs.append(out.cpu())
        print(f"two_device_check: {dev} ok, {eng.launch_count()} launches, finite={bool(torch.isfinite(out.float()).all())}")
    same = torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    print(f"two_device_check: cuda:0 and cuda:1 results identical: {same}")
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
