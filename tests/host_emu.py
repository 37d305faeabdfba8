"""Host emulation of device kernels that use no warp-level primitive (test infrastructure, never product code).

A kernel that walks its work in grid-stride / block-stride loops and synchronises only through block barriers and integer
atomics is, on a grid of ONE thread, an ordinary sequential program.  build() cuts the anonymous namespace (the kernels and
their device helpers, verbatim) out of a csrc/*.cu file and compiles it with g++ against the CUDA headers' host definitions:
__global__ / __device__ / __shared__ / __constant__ become ignored attributes, blockIdx / threadIdx are constants of a
1 x 1 grid, __syncthreads() is a no-op and atomicAdd a plain addition.  The structs the kernels share with the library's
host side are taken from csrc/hvb_internal.cuh textually, so the emulation cannot drift from them.  What emulation cannot
show is device-only behaviour (alignment faults, the launch); the tests/test_gpu_zz_*.py files cover that on a B200."""
import ctypes as C
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "turingcodec_b200" / "csrc"
CUDA_INC = Path("/usr/local/cuda/include")

PRELUDE = r'''
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include "hvb.h"
#define __launch_bounds__(...)
using std::max;
using std::min;
@STRUCTS@
static inline int hvbClip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline void __syncthreads() {}
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); } // one fiber runs at a time; the host side reads after the launch returns
static inline int atomicAdd(int *p, int v) { const int old = *p; *p += v; return old; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { const unsigned long long old = *p; *p += v; return old; }
static const uint3 blockIdx = {0, 0, 0}, threadIdx = {0, 0, 0};
static const dim3 blockDim(1), gridDim(1);
'''


def struct_text(header: str, name: str) -> str:
    """the definition of `struct name { ... };` as csrc/hvb_internal.cuh has it"""
    start = header.index(f"struct {name}\n{{")
    return header[start:header.index("};", start) + 2]


def build(tmp_dir: Path, cu_file: str, structs: list[str], entry: str) -> C.CDLL:
    """-> the emulation library of csrc/<cu_file>: its anonymous namespace + `entry` (extern "C" functions calling the kernels)"""
    if not (CUDA_INC / "cuda_runtime.h").exists():
        pytest.skip("CUDA headers not found")
    src = (CSRC / cu_file).read_text()
    kernels = src[src.index("namespace {"):src.index("} // namespace") + len("} // namespace")]
    for forbidden in ("__shfl", "__syncwarp", "__ballot", "__any_sync", "asm volatile", "asm("):
        assert forbidden not in kernels, f"{cu_file} uses {forbidden}: not emulable on a 1-thread grid"
    internal = (CSRC / "hvb_internal.cuh").read_text()
    text = PRELUDE.replace("@STRUCTS@", "\n".join(struct_text(internal, s) for s in structs)) + kernels + entry
    (tmp_dir / "emu.cpp").write_text(text)
    # -fsanitize=alignment: a misaligned vector access, which x86 tolerates, aborts here as it would fault on the device (the
    # tests give the kernels buffers aligned the way the device's are)
    subprocess.run(["g++", "-O1", "-fPIC", "-shared", "-w", "-std=c++17", "-fsanitize=alignment", "-fno-sanitize-recover=alignment",
                    "-static-libubsan", f"-I{CUDA_INC}", f"-I{ROOT / 'include'}", str(tmp_dir / "emu.cpp"), "-o", str(tmp_dir / "emu.so")],
                   check=True, capture_output=True)
    return C.CDLL(str(tmp_dir / "emu.so"))
