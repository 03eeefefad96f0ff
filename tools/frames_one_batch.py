"""One 1800-row batch of the upstream entry (600 one-second chunks x 4 frames, north-star widths) — a short target for
ncu captures of individual kernels.   python tools/frames_one_batch.py [chunks]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tdc_video_b200 import QFormerEngine  # noqa: E402
from tdc_video_b200.synth import QFormerGeometry, make_frontend_state_dict, make_state_dict  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 600
geom = QFormerGeometry(d_enc=3584, d_out=3584, vocab=0)
sd = make_state_dict(geom, 1, with_text=False)
sd.update(make_frontend_state_dict(3584, 1024, 768, geom.hidden, 2))
eng = QFormerEngine(d_enc=3584, d_out=3584, vocab=0, d_frame_in=1024, d_audio=768)
eng.load_weights(sd)
frames = torch.randn(S * 4, 144, 1024, device="cuda").bfloat16()
audio = (torch.randn(S * 4, 50, 768, device="cuda") * 0.5).bfloat16()
st = torch.arange(S, dtype=torch.int32) * 4
rf = (st[:, None] + torch.arange(1, 4, dtype=torch.int32)[None]).reshape(-1)
rc = torch.arange(S, dtype=torch.int32).repeat_interleave(3)
for _ in range(2):
    eng.compress_frames(frames, st, rf, rc, audio=audio)
torch.cuda.synchronize()
print("done", eng.launch_count())
