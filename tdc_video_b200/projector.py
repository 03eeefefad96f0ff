"""Drop-in for the reference's GELU-MLP projector (`mm_projector`).

Reference: `nn.Sequential(nn.Linear(d_in, d), nn.GELU(), nn.Linear(d, d))` built at
tdc/cambrian_arch.py:65-69 (SVA variant) and tdc/multimodal_projector/builder.py:40-47
(`mlp2x_gelu`), applied to every frame token at cambrian_arch.py:1149-1150.

`GeluMLPProjector` keeps the Sequential's parameter names (`0.weight`, `0.bias`, `2.weight`, `2.bias`) so
`model.mm_projector.*` checkpoint entries load unchanged; the forward is one `tdc_gelu_mlp` call: two
tcgen05 GEMMs with bias + exact-erf GELU fused into the first epilogue.  CUDA only.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib
from .engine import _ptr, _stream


def gelu_mlp(x: torch.Tensor, w0: torch.Tensor, b0: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor) -> torch.Tensor:
    """y = Linear(w1,b1)(gelu_erf(Linear(w0,b0)(x))) in bf16 (fp32 accumulate) via tdc_gelu_mlp."""
    if not x.is_cuda:
        raise RuntimeError("tdc_video_b200.gelu_mlp needs CUDA tensors: there is no CPU fallback")
    lib = _lib.load_library()
    d_in, d_mid, d_out = x.shape[-1], w0.shape[0], w1.shape[0]
    x2 = x.reshape(-1, d_in).to(torch.bfloat16).contiguous()
    dev = x.device
    w0b, w1b = w0.to(dev, torch.bfloat16).contiguous(), w1.to(dev, torch.bfloat16).contiguous()
    b0f, b1f = b0.to(dev, torch.float32).contiguous(), b1.to(dev, torch.float32).contiguous()
    mid = torch.empty((x2.shape[0], d_mid), dtype=torch.bfloat16, device=dev)
    y = torch.empty((x2.shape[0], d_out), dtype=torch.bfloat16, device=dev)
    if x2.shape[0] > 0:
        with torch.cuda.device(dev):
            rc = lib.tdc_gelu_mlp(_ptr(x2), _ptr(w0b), _ptr(b0f), _ptr(w1b), _ptr(b1f), _ptr(mid), _ptr(y),
                                  x2.shape[0], d_in, d_mid, d_out, _stream(dev))
        _lib.check(rc, None, "tdc_gelu_mlp")
    return y.reshape(*x.shape[:-1], d_out).to(x.dtype if x.dtype in (torch.float16, torch.bfloat16) else torch.bfloat16)


class GeluMLPProjector(nn.Module):
    """`mm_projector`: Linear(d_in -> d) . GELU . Linear(d -> d); parameters live under `0.*` and `2.*`."""

    def __init__(self, d_in: int, d_out: int):
        super().__init__()
        self.add_module("0", nn.Linear(d_in, d_out))
        self.add_module("1", nn.GELU())
        self.add_module("2", nn.Linear(d_out, d_out))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise RuntimeError("GeluMLPProjector is inference-only (eval mode)")
        l0, l2 = getattr(self, "0"), getattr(self, "2")
        return gelu_mlp(x, l0.weight, l0.bias, l2.weight, l2.bias)
