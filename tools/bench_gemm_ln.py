"""Time the fused GEMM + residual + LayerNorm kernel against GEMM -> fp32 -> LayerNorm on the Q-Former's shapes.
    python tools/bench_gemm_ln.py [rows]        (rows of 16 query tokens; default 1800 = one internal batch)"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tdc_video_b200 import _lib  # noqa: E402
from tdc_video_b200.engine import _ptr, _stream  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1800
    lib = _lib.load_library()
    dev = torch.device("cuda")
    M, N = rows * 16, 768
    s = _stream(dev)
    for K in (768, 3072):
        x = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
        bias, gamma, beta = (torch.randn(N, device=dev) for _ in range(3))
        resid = torch.randn(M, N, device=dev)
        pre = torch.empty(M, N, device=dev)
        y32, y16 = torch.empty(M, N, device=dev), torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        z32, z16 = torch.empty(M, N, device=dev), torch.empty(M, N, device=dev, dtype=torch.bfloat16)

        def fused():
            rc = lib.tdc_linear_layernorm(_ptr(x), _ptr(w), _ptr(bias), _ptr(resid), _ptr(gamma), _ptr(beta),
                                          C.c_float(1e-12), _ptr(y32), _ptr(y16), M, N, K, s)
            assert rc == 0, lib.tdc_last_error(None)

        def gemm_only():
            assert lib.tdc_linear(_ptr(x), _ptr(w), _ptr(bias), _ptr(pre), M, N, K, _lib.TDC_F32, 0, 0, s) == 0

        def ln_only():
            assert lib.tdc_layernorm(_ptr(pre), _ptr(resid), 0, _ptr(gamma), _ptr(beta), C.c_float(1e-12), _ptr(z32),
                                     _ptr(z16), M, N, s) == 0

        fused(); gemm_only(); ln_only()
        torch.cuda.synchronize()
        err = (y32 - z32).abs().max().item()
        tf, tg, tl = timeit(fused), timeit(gemm_only), timeit(ln_only)
        fl = 2.0 * M * N * K / 1e9
        print(f"M {M} N {N} K {K}: fused {tf * 1e3:.0f} us ({fl / tf:.0f} TFLOP/s) | gemm {tg * 1e3:.0f} us "
              f"({fl / tg:.0f} TFLOP/s) + layernorm {tl * 1e3:.0f} us = {(tg + tl) * 1e3:.0f} us | max |fused - split| {err:.2e}",
              flush=True)


if __name__ == "__main__":
    main()
