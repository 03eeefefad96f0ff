"""ctypes binding of libtdc_b200.so (C ABI: include/tdc_b200.h).

There is deliberately no fallback: if the shared library is missing it is built with nvcc
(in-tree); if that fails, or a call is made without a Blackwell GPU, a RuntimeError is
raised.  Nothing here imports the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libtdc_b200.so"

TDC_OK = 0
STATUS_NAMES = {0: "TDC_OK", -1: "TDC_EINVAL", -2: "TDC_ECUDA", -3: "TDC_ENOMEM", -4: "TDC_ESTATE",
                -5: "TDC_EWORKSPACE"}
TDC_BF16, TDC_F16, TDC_F32 = 0, 1, 2
K_KV_GEMM, K_QUERY_GEMM, K_ATTENTION, K_ROWOPS, K_FRONTEND, K_COUNT = 0, 1, 2, 3, 4, 5
KERNEL_CLASS_NAMES = ["kv_gemm", "query_gemm", "attention", "rowops", "frontend"]

# every symbol include/tdc_b200.h declares (checked by tests/test_c_abi.py)
EXPORTED_SYMBOLS = [
    "tdc_abi_version", "tdc_create", "tdc_destroy", "tdc_last_error", "tdc_load_weights", "tdc_workspace_bytes",
    "tdc_qformer_forward", "tdc_proj_norm", "tdc_compress", "tdc_compress_multicast", "tdc_frames_workspace_bytes",
    "tdc_compress_frames", "tdc_linear", "tdc_linear_layernorm", "tdc_gelu_mlp", "tdc_avg_pool_tokens",
    "tdc_convert", "tdc_layernorm", "tdc_attention", "tdc_residual_add", "tdc_resize_tokens_bilinear", "tdc_window_rearrange", "tdc_combine_parts", "tdc_multicast_copy", "tdc_peer_copy", "tdc_segment_workspace_bytes", "tdc_segment_boundaries", "tdc_set_profiling", "tdc_get_profile", "tdc_reset_profile", "tdc_launch_count",
]


class TdcConfig(C.Structure):
    _fields_ = [
        ("hidden", C.c_int32), ("heads", C.c_int32), ("intermediate", C.c_int32), ("layers", C.c_int32),
        ("cross_freq", C.c_int32), ("d_enc", C.c_int32), ("d_out", C.c_int32), ("vocab", C.c_int32),
        ("max_pos", C.c_int32), ("ln_eps", C.c_float), ("gemm_cta_group", C.c_int32), ("d_frame_in", C.c_int32),
        ("d_audio", C.c_int32), ("reserved", C.c_int32 * 3),
    ]


class TdcTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("dtype", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int64 * 4)]


class TdcFramesArgs(C.Structure):
    """tdc_frames_args (include/tdc_b200.h)."""
    _fields_ = [
        ("frames", C.c_void_p), ("audio", C.c_void_p), ("static_frames", C.c_void_p), ("row_frames", C.c_void_p),
        ("row_chunk", C.c_void_p), ("input_ids", C.c_void_p),
        ("n_frames", C.c_int32), ("n_chunks", C.c_int32), ("rows", C.c_int32), ("visual_tokens", C.c_int32),
        ("audio_tokens", C.c_int32), ("num_query", C.c_int32), ("num_text", C.c_int32),
        ("learned_queries", C.c_int32), ("fold", C.c_int32), ("multicast", C.c_int32), ("out_dtype", C.c_int32),
        ("no_layer0_dedup", C.c_int32),
        ("static_out", C.c_void_p), ("out", C.c_void_p), ("chunk_prompt", C.c_void_p),
        ("n_prompts", C.c_int32), ("static_multicast", C.c_int32), ("static_ready_event", C.c_void_p),
    ]


_lib: Optional[C.CDLL] = None


def _declare(lib: C.CDLL) -> None:
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    lib.tdc_abi_version.restype = C.c_int
    lib.tdc_create.argtypes = [C.POINTER(vp), C.POINTER(TdcConfig)]
    lib.tdc_destroy.argtypes = [vp]
    lib.tdc_last_error.argtypes = [vp]
    lib.tdc_last_error.restype = C.c_char_p
    lib.tdc_load_weights.argtypes = [vp, C.POINTER(TdcTensor), i32, vp]
    lib.tdc_workspace_bytes.argtypes = [vp, i32, i32, i32, i32]
    lib.tdc_workspace_bytes.restype = sz
    fwd = [vp, vp, i32, vp, vp, vp, vp, i32, vp, i32, i32, i32, i32, vp, i32, vp, sz, vp]
    lib.tdc_qformer_forward.argtypes = fwd
    lib.tdc_compress.argtypes = fwd
    lib.tdc_compress_multicast.argtypes = fwd
    lib.tdc_compress_multicast.restype = C.c_int
    lib.tdc_frames_workspace_bytes.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32]
    lib.tdc_frames_workspace_bytes.restype = sz
    lib.tdc_compress_frames.argtypes = [vp, C.POINTER(TdcFramesArgs), vp, sz, vp]
    lib.tdc_compress_frames.restype = C.c_int
    lib.tdc_proj_norm.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32, vp, sz, vp]
    lib.tdc_linear.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.tdc_linear_layernorm.argtypes = [vp, vp, vp, vp, vp, vp, C.c_float, vp, vp, i32, i32, i32, vp]
    lib.tdc_linear_layernorm.restype = C.c_int
    lib.tdc_gelu_mlp.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.tdc_avg_pool_tokens.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp]
    lib.tdc_convert.argtypes = [vp, i32, vp, i32, i64, vp]
    lib.tdc_layernorm.argtypes = [vp, vp, i32, vp, vp, C.c_float, vp, vp, i64, i32, vp]
    lib.tdc_attention.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, i32, i32, i32, i32, i64, i64, i32, i32, i64, i64,
                                  vp, vp, vp]
    lib.tdc_residual_add.argtypes = [vp, vp, vp, vp, i64, vp]
    lib.tdc_resize_tokens_bilinear.argtypes = [vp, i32, i32, i32, i32, i32, vp, i32, vp]
    lib.tdc_resize_tokens_bilinear.restype = C.c_int
    lib.tdc_window_rearrange.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp]
    lib.tdc_window_rearrange.restype = C.c_int
    lib.tdc_combine_parts.argtypes = [vp, vp, vp, i32, i32, i64, i32, vp, vp]
    lib.tdc_combine_parts.restype = C.c_int
    lib.tdc_multicast_copy.argtypes = [vp, vp, C.c_size_t, i32, vp]
    lib.tdc_multicast_copy.restype = C.c_int
    lib.tdc_peer_copy.argtypes = [vp, vp, C.c_size_t, vp]
    lib.tdc_peer_copy.restype = C.c_int
    for _n in ("tdc_layernorm", "tdc_attention", "tdc_residual_add"):
        getattr(lib, _n).restype = C.c_int
    lib.tdc_segment_workspace_bytes.argtypes = [i32, i64]
    lib.tdc_segment_workspace_bytes.restype = sz
    lib.tdc_segment_boundaries.argtypes = [vp, i32, i32, i64, i32, vp, vp, vp, sz, vp]
    lib.tdc_set_profiling.argtypes = [vp, i32]
    lib.tdc_get_profile.argtypes = [vp, i32, C.POINTER(C.c_double), C.POINTER(i64)]
    lib.tdc_reset_profile.argtypes = [vp]
    lib.tdc_launch_count.argtypes = [vp]
    lib.tdc_launch_count.restype = i64
    for name in ("tdc_create", "tdc_destroy", "tdc_load_weights", "tdc_qformer_forward", "tdc_compress",
                 "tdc_proj_norm", "tdc_linear", "tdc_gelu_mlp", "tdc_avg_pool_tokens", "tdc_convert", "tdc_segment_boundaries",
                 "tdc_set_profiling", "tdc_get_profile", "tdc_reset_profile"):
        getattr(lib, name).restype = C.c_int


def load_library(build_if_missing: bool = True) -> C.CDLL:
    """dlopen libtdc_b200.so (building it in-tree first if needed).  Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if not build_if_missing:
            raise RuntimeError(f"{LIB_PATH} not built; run `python -m tdc_video_b200.build`")
        from .build import build_library
        build_library()
    lib = C.CDLL(str(LIB_PATH))
    _declare(lib)
    if lib.tdc_abi_version() != 2:
        raise RuntimeError("libtdc_b200.so ABI version mismatch; rebuild with `python -m tdc_video_b200.build --force`")
    _lib = lib
    return lib


class TdcError(RuntimeError):
    pass


def check(rc: int, handle=None, what: str = "") -> None:
    if rc == TDC_OK:
        return
    lib = load_library()
    msg = lib.tdc_last_error(handle)
    raise TdcError(f"{what or 'libtdc_b200'} failed: {STATUS_NAMES.get(rc, rc)}: "
                   f"{msg.decode() if msg else ''}")
