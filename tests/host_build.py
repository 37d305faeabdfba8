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
        launcher = "emuLaunchExact" if name.split("<")[0] in EXACT_GRID else "emuLaunch"
        out += text[pos:m.start()] + f"{launcher}({cfg[0]}, {cfg[1]}, [&] {{ {name}{text[args_start:args_end]}; }})"
        pos = args_end


# kernels whose launches size the grid by the work (a thread or block per element, no stride loop): never capped
EXACT_GRID = ("tuOrderKernel", "rdoqBitsKernel", "rdoqLastKernel", "codedResidualTotalKernel")

# the inline-PTX wrappers of the individual files and their emulated replacements (as in tests/emu_context.py)
FILE_PTX = {
    "hvb_metrics.cu": dict(replace={
        "__device__ __forceinline__ void cpAsync8(": "static inline void cpAsync8(uint32_t dst, const void *src) { memcpy(emu::sharedArena + dst, src, 8); }",
        "__device__ __forceinline__ void cpAsync4(": "static inline void cpAsync4(uint32_t dst, const void *src) { memcpy(emu::sharedArena + dst, src, 4); }",
        "__device__ __forceinline__ void cpAsync16(": "static inline void cpAsync16(uint32_t dst, const void *src) { memcpy(emu::sharedArena + dst, src, 16); }",
        "__device__ __forceinline__ void cpAsyncCommit(": "static inline void cpAsyncCommit() {}",
        "template <int PENDING>\n__device__ __forceinline__ void cpAsyncWait(": "template <int PENDING> static inline void cpAsyncWait() {}"}),
    "hvb_intra.cu": dict(mma={"imma16832": False}),
    "hvb_tu.cu": dict(mma={"immaS8U8": False, "immaS8S8": True}),
}

RUNTIME_TEMPLATES = ("#include <mutex>\n"
                     "// kernels are plain functions here; cuda_runtime.h declares this overload for nvcc only\n"
                     "template <class... A> static cudaError_t cudaFuncSetAttribute(void (*)(A...), cudaFuncAttribute, int) { return cudaSuccess; }\n")


def build(tmp_dir: Path, cu_files: tuple, stubs: str = "", max_grid: int = 2, cpp_files: tuple = (), soname: str = "") -> C.CDLL:
    """stubs: C++ definitions of the functions the named files call in files that are left out; cpp_files: plain C++ sources
    of csrc/ to compile along (the havoc table shim); soname: give the library this soname (to stand in for libhvb.so)"""
    if not (host_emu_warp.CUDA_INC / "cuda_runtime.h").exists():
        pytest.skip("CUDA headers not found")
    internal = (CSRC / "hvb_internal.cuh").read_text()
    helpers = internal[internal.index("#ifdef __CUDACC__") + len("#ifdef __CUDACC__"):internal.index("#endif // __CUDACC__")]
    objects = []
    # hvb_unit.cuh without its two inline-PTX wrappers (the prelude emulates them) shadows the real header
    (tmp_dir / "hvb_unit.cuh").write_text("#pragma once\n" + host_emu_warp.unit_header())
    flags = ["-O1", "-fPIC", "-w", "-std=c++17", f"-I{tmp_dir}", f"-I{host_emu_warp.CUDA_INC}", f"-I{ROOT / 'include'}", f"-I{CSRC}"]
    for name in cu_files:
        src = (CSRC / name).read_text()
        ptx = FILE_PTX.get(name, {})
        injected = ""
        for signature, replacement in ptx.get("replace", {}).items():
            src = host_emu_warp._strip_function(src, signature)
            injected += replacement + "\n"
        for wrapper, signed_b in ptx.get("mma", {}).items():
            src = host_emu_warp._strip_function(src, f"__device__ __forceinline__ void {wrapper}(")
            injected += (f"static inline void {wrapper}(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)"
                         f" {{ emuMma16832(c, a0, a1, a2, a3, b0, b1, {'true' if signed_b else 'false'}); }}\n")
        assert "asm volatile" not in src and "asm(" not in src, f"{name}: inline PTX outside the emulated wrappers"
        src = rewrite_launches(src)
        assert "<<<" not in src
        # the file's own include of hvb_internal.cuh brings the host declarations; its device helpers are compiled only by
        # nvcc (#ifdef __CUDACC__), so they are appended behind that include together with the emulator's prelude
        include = '#include "hvb_internal.cuh"'
        assert include in src
        prelude = host_emu_warp.PRELUDE.replace('#include "hvb.h"', "").replace('extern "C" void emu_set_schedule(', "static void emu_set_schedule(")
        # the entry points size their grids for 148 SMs; the kernels built here walk their work in grid-stride loops, so a
        # grid of at most `max_grid` blocks computes the same and keeps the emulation fast
        # launches are serialised per translation unit (the emulator's state is static), whichever host thread issues them
        head = "template <class F> static void emuLaunch(int grid, int block, F kernel)\n{"
        assert head in prelude
        # (one mutex per translation unit, not per instantiation of the launcher: two host threads launching DIFFERENT
        # kernels of one file share the emulator's static state just the same)
        prelude = prelude.replace(head, "static std::mutex &emuSerial() { static std::mutex m; return m; }\n"
                                        "template <class F> static void emuLaunchExact(int grid, int block, F kernel)\n{\n"
                                        "    std::lock_guard<std::mutex> hold(emuSerial());")
        prelude += ("template <class F> static void emuLaunch(int grid, int block, F kernel) "
                    f"{{ emuLaunchExact(std::min(grid, {max_grid}), block, kernel); }}\n")
        src = src.replace(include, include + "\n" + RUNTIME_TEMPLATES + prelude + helpers + injected, 1)
        src = re.sub(r"extern\s+__shared__\s+(__align__\(\d+\)\s+)?(\w[\w\s]*?)\s+(\w+)\[\];", r"\2 *const \3 = reinterpret_cast<\2 *>(emu::sharedArena);", src)
        src = src.replace("__shared__", "static")
        cpp = tmp_dir / (name.replace(".", "_") + ".cpp")
        cpp.write_text(src)
        obj = cpp.with_suffix(".o")
        res = subprocess.run(["g++", *flags, "-c", str(cpp), "-o", str(obj)], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr[-3000:]
        objects.append(str(obj))
    for name in cpp_files:
        obj = tmp_dir / (name.replace(".", "_") + ".o")
        res = subprocess.run(["g++", *flags, "-c", str(CSRC / name), "-o", str(obj)], capture_output=True, text=True)
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
    lib = tmp_dir / (soname or "libhvb_host.so")
    # -Bsymbolic: the library's calls bind to its own (fake) runtime even when the real libcudart is already in the process
    res = subprocess.run(["g++", "-shared", "-Wl,-Bsymbolic", *(["-Wl,-soname=" + soname] if soname else []), "-o", str(lib), *objects, str(fake),
                          "-pthread"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    return C.CDLL(str(lib))
