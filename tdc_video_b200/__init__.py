"""tdc_video_b200 — B200-native (sm_100a) implementation of TDC-Video's Temporal Dynamic
Context compression path: Q-Former + vision_proj/L2-normalise + projector, behind the
reference's own module interface.  Hand-written CUDA (tcgen05/TMEM/TMA GEMMs, register-direct
short-query attention, fused row ops) in libtdc_b200.so; torch only for memory, streams and
torch.distributed."""
from ._lib import TdcError, load_library  # noqa: F401
from .engine import QFormerEngine, avg_pool_tokens, linear  # noqa: F401
from .projector import GeluMLPProjector, gelu_mlp  # noqa: F401

__all__ = ["QFormerEngine", "TdcError", "load_library", "linear", "avg_pool_tokens", "GeluMLPProjector", "gelu_mlp"]
