"""The folded K/V projection GEMM of one 1800-row batch of the upstream entry as a stand-alone launch (for ncu):
A = gelu(mm_projector.0) output [1800 x 144, 3584], W = (Wkv . W2) [9216, 3584], bf16 out.   python tools/kv_gemm_shape.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tdc_video_b200 import linear  # noqa: E402

M, N, K = 1800 * 144, 9216, 3584
x = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
b = torch.randn(N, device="cuda")
y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    linear(x, w, b, out=y)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    linear(x, w, b, out=y)
e.record()
torch.cuda.synchronize()
ms = a.elapsed_time(e) / 5
print(f"kv gemm M {M} N {N} K {K}: {ms:.3f} ms = {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s")
