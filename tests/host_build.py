"""A CPU build of libhvb's HOST side over the emulated kernels (test infrastructure, never product code).

tests/host_emu_warp.py runs kernels; this module goes one level up: whole csrc/*.cu files -- context, pictures, pools,
staging and the batch entry points exactly as written -- are compiled with g++ against tests/fake_cuda/fake_cudart.cpp (the
CUDA runtime on host memory), each kernel launch `k<<<grid, block, smem, stream>>>(args)` rewritten textually into a launch of
the same kernel under the warp-level emulator.  The result is a shared library with libhvb.so's C-ABI that
turingcodec_b200/hvb.py can drive, so the Python binding, the argument checks, the uploads and the launch geometry of an
entry point are exercised without a GPU.  Only a subset of the files is built (what the caller names); hvb.py tolerates the
missing entry points."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

import host_emu_warp

ROOT = Path(__file__).resolve().parent.parent
CSRC = host_emu_warp.CSRC


def _matching(text: str, start: int, open_ch: str, close_ch: str) -> int:
    """index just behind the bracket that closes the one at text[start]"""
    depth = 0
    for i in range(start, len(text)):
        depth += text[i] == open_ch
        depth -= text[i] == close_ch
        if depth == 0:
            return i + 1
    raise ValueError("unbalanced brackets")


def _split_top_level(text: str) -> list:
    parts, depth, cur = [], 0, ""
    for ch in text:
        if ch in "(<[":
            depth += 1
        elif ch in ")>]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite_launches(text: str) -> str:
    """kernel<T><<<grid, block[, smem[, stream]]>>>(args);  ->  emuLaunch(grid, block, [&] { kernel<T>(args); });"""
    out, pos = "", 0
    while True:
        at = text.find("<<<", pos)
        if at < 0:
            return out + text[pos:]
        # the kernel name (with an optional template argument list) ends at `at`
        name_start = at
        if text[at - 1] == ">":
            depth, i = 0, at - 1
            while True:
                depth += text[i] == ">"
                depth -= text[i] == "<"
                if depth == 0:
                    break
                i -= 1
            name_start = i
        m = re.search(r"[\w:]+$", text[:name_start])
        name = text[m.start():at]
        cfg_end = text.index(">>>", at)
        cfg = _split_top_level(text[at + 3:cfg_end])
        args_start = text.index("(", cfg_end)
        args_end = _matching(text, args_start, "(", ")")
        assert text[args_end:].lstrip().startswith(";"), "a launch is expected to be a statement"
        out += text[pos:m.start()] + f"emuLaunch({cfg[0]}, {cfg[1]}, [&] {{ {name}{text[args_start:args_end]}; }})"
        pos = args_end


def build(tmp_dir: Path, cu_files: tuple, stubs: str = "", max_grid: int = 2) -> C.CDLL:
    """stubs: C++ definitions of the functions the named files call in files that are left out"""
    if not (host_emu_warp.CUDA_INC / "cuda_runtime.h").exists():
        pytest.skip("CUDA headers not found")
    internal = (CSRC / "hvb_internal.cuh").read_text()
    helpers = internal[internal.index("#ifdef __CUDACC__") + len("#ifdef __CUDACC__"):internal.index("#endif // __CUDACC__")]
    objects = []
    flags = ["-O1", "-fPIC", "-w", "-std=c++17", f"-I{host_emu_warp.CUDA_INC}", f"-I{ROOT / 'include'}", f"-I{CSRC}"]
    for name in cu_files:
        src = rewrite_launches((CSRC / name).read_text())
        assert "<<<" not in src
        # the file's own include of hvb_internal.cuh brings the host declarations; its device helpers are compiled only by
        # nvcc (#ifdef __CUDACC__), so they are appended behind that include together with the emulator's prelude
        include = '#include "hvb_internal.cuh"'
        assert include in src
        prelude = host_emu_warp.PRELUDE.replace('#include "hvb.h"', "").replace('extern "C" void emu_set_schedule(', "static void emu_set_schedule(")
        # the entry points size their grids for 148 SMs; the kernels built here walk their work in grid-stride loops, so a
        # grid of at most `max_grid` blocks computes the same and keeps the emulation fast
        cap = "    gridDim = dim3(grid); blockDim = dim3(block);"
        assert cap in prelude
        prelude = prelude.replace(cap, f"    grid = std::min(grid, {max_grid});\n" + cap)
        src = src.replace(include, include + "\n" + prelude + helpers, 1)
        src = re.sub(r"extern\s+__shared__\s+(__align__\(\d+\)\s+)?(\w[\w\s]*?)\s+(\w+)\[\];", r"\2 *const \3 = reinterpret_cast<\2 *>(emu::sharedArena);", src)
        src = src.replace("__shared__", "static")
        cpp = tmp_dir / (name.replace(".", "_") + ".cpp")
        cpp.write_text(src)
        obj = cpp.with_suffix(".o")
        res = subprocess.run(["g++", *flags, "-c", str(cpp), "-o", str(obj)], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr[-3000:]
        objects.append(str(obj))
    if stubs:
        cpp = tmp_dir / "stubs.cpp"
        cpp.write_text('#include "hvb_internal.cuh"\n' + stubs)
        res = subprocess.run(["g++", *flags, "-c", str(cpp), "-o", str(cpp.with_suffix(".o"))], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr[-3000:]
        objects.append(str(cpp.with_suffix(".o")))
    fake = tmp_dir / "fake_cudart.o"
    res = subprocess.run(["g++", *flags, "-c", str(ROOT / "tests" / "fake_cuda" / "fake_cudart.cpp"), "-o", str(fake)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    lib = tmp_dir / "libhvb_host.so"
    # -Bsymbolic: the library's calls bind to its own (fake) runtime even when the real libcudart is already in the process
    res = subprocess.run(["g++", "-shared", "-Wl,-Bsymbolic", "-o", str(lib), *objects, str(fake)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    return C.CDLL(str(lib))
