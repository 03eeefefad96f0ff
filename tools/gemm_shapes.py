"""Time the plain tcgen05 GEMM (tdc_linear) on the query-side shapes of one 1800-row batch (M = 28 800 tokens) and on
the K/V shape, CUDA events, L2-cold rotation of the operands.  Dev knobs are read from the environment by the library
(TDC_GEMM_DEBUG=1: drain TMEM but skip the epilogue math + stores).   python tools/gemm_shapes.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tdc_video_b200.engine import linear  # noqa: E402

SHAPES = [("self Q|K|V", 28800, 2304, 768, False), ("FFN up + GELU", 28800, 3072, 768, True),
          ("cross Q", 28800, 768, 768, False), ("K/V visual (1/4 batch)", 64800, 9216, 3584, False)]


def main():
    dev = torch.device("cuda", 0)
    for name, m, n, k, gelu in SHAPES:
        copies = 4                                   # rotate operands so that A is not L2-resident from the last launch
        xs = [torch.randn(m, k, device=dev).bfloat16() for _ in range(copies)]
        w = (torch.randn(n, k, device=dev) * 0.02).bfloat16()
        b = torch.randn(n, device=dev)
        outs = [torch.empty(m, n, dtype=torch.bfloat16, device=dev) for _ in range(copies)]
        for i in range(copies):
            linear(xs[i], w, b, gelu=gelu, out=outs[i])
        torch.cuda.synchronize()
        iters = 40
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            linear(xs[i % copies], w, b, gelu=gelu, out=outs[i % copies])
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        print(f"{name:26s} M {m:6d} N {n:5d} K {k:5d}: {us:8.1f} us  {2.0 * m * n * k / us / 1e6:7.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
