"""Build libtdc_b200.so (and the C oracle) in-tree with nvcc / gcc.

The library is plain CUDA C++ behind a C ABI (include/tdc_b200.h): no torch headers,
cudart linked statically, the one driver entry point it needs (cuTensorMapEncodeTiled)
resolved at run time — so the .so loads (and exports its symbols) on a box without a
GPU or libcuda, and fails loudly only when a compute call is made.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libtdc_b200.so"
OBJ_DIR = REPO_ROOT / "build" / "obj"

SOURCES = ["gemm_sm100.cu", "gemm_ln_sm100.cu", "attention.cu", "rowops.cu", "segment.cu", "frontend.cu", "tdc_api.cu"]
HEADERS = ["tdc_ptx.cuh", "tdc_gemm.cuh", "tdc_kernels.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    f"-I{REPO_ROOT / 'include'}", f"-I{CSRC}",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: cannot build libtdc_b200.so")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile every kernel for sm_100a and link libtdc_b200.so next to the package."""
    nvcc = _nvcc()
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    common_deps = [CSRC / h for h in HEADERS] + [REPO_ROOT / "include" / "tdc_b200.h", Path(__file__)]

    def compile_one(src: str) -> Path:
        obj = OBJ_DIR / (src + ".o")
        if force or _stale(obj, [CSRC / src] + common_deps):
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), file=sys.stderr)
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
            if verbose:
                print(res.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=4) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    if force or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
               "-o", str(LIB_PATH), *map(str, objs), "-cudart", "static"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
