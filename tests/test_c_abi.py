"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol
include/tdc_b200.h declares.  No compute calls here (CPU-only container)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from tdc_video_b200.build import build_library
    path = build_library()
    return ctypes.CDLL(str(path))


def _declared_symbols():
    text = (ROOT / "include" / "tdc_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tdc_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tdc_b200.h but not exported"


def test_python_binding_lists_the_same_symbols():
    from tdc_video_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared_symbols()


def test_abi_version_and_null_handling(lib):
    assert lib.tdc_abi_version() == 2
    lib.tdc_last_error.restype = ctypes.c_char_p
    lib.tdc_create.restype = ctypes.c_int
    assert lib.tdc_create(None, None) == -1          # TDC_EINVAL, no crash
    assert b"null" in lib.tdc_last_error(None)
    assert lib.tdc_destroy(None) == 0


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tdc_video_b200 import QFormerEngine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        QFormerEngine(d_enc=64)


def test_product_never_imports_the_oracle():
    for p in (ROOT / "tdc_video_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p
