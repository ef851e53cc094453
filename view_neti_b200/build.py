"""In-tree build of libviewneti_sm100a.so (hand-written CUDA for sm_100a behind the C-ABI in include/viewneti.h).

    python -m view_neti_b200.build [--force] [-v]

nvcc cross-compiles without a GPU; the .so lands in view_neti_b200/lib/ (git-ignored, travels with gpurun).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import List

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
OBJDIR = PKG / "build"
# experiment builds: VN_LIB_SUFFIX=_x VN_CFLAGS="-DVN_EPI_HALVES=1" python -m view_neti_b200.build --force
_SUFFIX = os.environ.get("VN_LIB_SUFFIX", "")
LIB = LIBDIR / f"libviewneti_sm100a{_SUFFIX}.so"
if _SUFFIX:
    OBJDIR = PKG / f"build{_SUFFIX}"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def sources() -> List[Path]:
    return sorted(CSRC.glob("*.cu"))


def _deps_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list((PKG.parent / "include").glob("*.h"))
    return max((h.stat().st_mtime for h in hdrs), default=0.0)


def build(force: bool = False, verbose: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    OBJDIR.mkdir(exist_ok=True)
    srcs = sources()
    if not srcs:
        raise RuntimeError(f"no CUDA sources under {CSRC}")
    hdr_m = _deps_mtime()
    nvcc = _nvcc()

    def compile_one(src: Path) -> Path:
        obj = OBJDIR / (src.stem + ".o")
        if not force and obj.exists() and obj.stat().st_mtime >= max(src.stat().st_mtime, hdr_m):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("VN_CFLAGS", "").split(), "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose and r.stderr:
            print(r.stderr, flush=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    if force or not LIB.exists() or any(o.stat().st_mtime > LIB.stat().st_mtime for o in objs):
        # static cudart (nvcc default) and no libcuda link: the driver entry point for cuTensorMapEncodeTiled is
        # resolved at run time, so the library also loads (symbols only) on a CPU-only box.
        cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
