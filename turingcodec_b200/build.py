"""In-tree build of libhvb.so (the CUDA kernels + C-ABI) for sm_100a.

The reference builds havoc with CMake into a static library that JIT-assembles x86 at start-up
(havoc/CMakeLists.txt, havoc/Jit.h); here every kernel is compiled ahead of time by nvcc into one
shared object next to the sources, so the built library travels with the tree.
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB = CSRC / "libhvb.so"
OBJ = CSRC / "build"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libhvb.so cannot be built (there is no CPU fallback)")


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp"))


def _headers_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((CSRC.parent.parent / "include").glob("*.h"))
    return max((h.stat().st_mtime for h in hdrs), default=0.0)


def _compile(nvcc: str, src: Path, obj: Path, verbose: bool) -> None:
    cmd = [nvcc, *NVCC_FLAGS, "-I", str(CSRC.parent.parent / "include"), "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        sys.stderr.write(res.stderr)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link csrc/libhvb.so.  Incremental by mtime."""
    nvcc = _nvcc()
    OBJ.mkdir(exist_ok=True)
    hdr_time = _headers_mtime()
    jobs = []
    objs = []
    for src in sources():
        obj = OBJ / (src.stem + ".o")
        objs.append(obj)
        stale = force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_time)
        if stale:
            jobs.append((src, obj))
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
            list(pool.map(lambda so: _compile(nvcc, so[0], so[1], verbose), jobs))
    if jobs or not LIB.exists() or LIB.stat().st_mtime < max(o.stat().st_mtime for o in objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xlinker", "-soname=libhvb.so", "-o", str(LIB),
               *map(str, objs)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
