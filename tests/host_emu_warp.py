"""Host emulation of WARP-LEVEL device kernels (test infrastructure, never product code).

tests/host_emu.py runs kernels without warp primitives on a grid of one thread.  Kernels that shuffle, vote, synchronise
warps or issue tensor-core products need their 32 lanes in lock step; here every CUDA thread of a block is a FIBER
(its own stack; a register-push stack switch on x86-64, ucontext elsewhere) on one OS thread, scheduled round-robin, and a collective is a rendezvous of a warp's fibers:

  __shfl_sync / __shfl_xor_sync / __any_sync / __ballot_sync / __syncwarp     exchange through a per-warp buffer
  __syncthreads                                                                rendezvous of the block's fibers
  mma.sync.m16n8k32.s8.u8 (the imma16832 wrapper of csrc/hvb_unit.cuh)         gathers the warp's A / B / C fragments and
                                                                               forms the PTX-defined 16x8x32 product
  atomicAdd, __ldg, __funnelshift_r, __byte_perm, __dp4a, __dp2a_lo/hi, ...    plain host code (one fiber runs at a time)

build() assembles one translation unit: this prelude, the device helpers of csrc/hvb_internal.cuh, csrc/hvb_unit.cuh and
the anonymous namespace of the kernel's .cu file -- all VERBATIM except that the two inline-PTX wrapper bodies (imma16832,
dp4aUS) are replaced by the emulated ones and `__shared__` is mapped to block-wide storage.  Blocks run one after another;
the launch geometry is the caller's.  Device-only behaviour (alignment faults, real scheduling) remains the GPU tests'."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "turingcodec_b200" / "csrc"
CUDA_INC = Path("/usr/local/cuda/include")

PRELUDE = r'''
#include <cuda_runtime.h>
#include <ucontext.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "hvb.h"
#define __launch_bounds__(...)
#ifndef __noinline__
#define __noinline__
#endif
using std::max;
using std::min;

// ---- fibers -------------------------------------------------------------------------------------------------------------
namespace emu {
// A fiber switch is six register pushes and a stack-pointer exchange on x86-64 (swapcontext would add two system calls,
// for the signal mask, to every rendezvous); elsewhere ucontext does it.
#if defined(__x86_64__)
__attribute__((naked, noinline)) static void switchStack(void **saveSp, void *loadSp)
{
    asm volatile("pushq %rbp\n pushq %rbx\n pushq %r12\n pushq %r13\n pushq %r14\n pushq %r15\n"
                 "movq %rsp, (%rdi)\n movq %rsi, %rsp\n"
                 "popq %r15\n popq %r14\n popq %r13\n popq %r12\n popq %rbx\n popq %rbp\n ret\n");
}
struct Fiber { void *sp = nullptr; bool done = false; };
static void *schedulerSp = nullptr;
#else
struct Fiber { ucontext_t ctx; bool done = false; };
static ucontext_t scheduler;
#endif
static std::vector<Fiber> fibers;
static std::vector<std::vector<char>> stacks; // one per thread of a block, kept across blocks and launches
static int current = 0, threads = 0;
static void (*body)() = nullptr;
// a rendezvous is among the lanes of a mask: full warps, or the disjoint lane groups some kernels synchronise separately
struct MaskState { unsigned mask = 0; int arrived = 0; unsigned generation = 0; };
struct WarpState { MaskState masks[16]; uint32_t buf[32]; uint64_t wide[32][10]; };
static std::vector<WarpState> warps;
static int blockArrived = 0; static unsigned blockGeneration = 0;
// the order in which the scheduler visits the fibers: 0 ascending, 1 descending, 2 a fresh pseudo-random permutation per sweep.
// Between two collectives a lane runs undisturbed, so a missing barrier shows as a result that depends on this order.
static int scheduleMode = 0; static unsigned scheduleState = 1;
alignas(16) static unsigned char sharedArena[256 * 1024];

#if defined(__x86_64__)
static void yield() { switchStack(&fibers[current].sp, schedulerSp); }
static void trampoline() { body(); fibers[current].done = true; switchStack(&fibers[current].sp, schedulerSp); __builtin_trap(); }
static void resume(int t) { current = t; switchStack(&schedulerSp, fibers[t].sp); }
static void prepare(Fiber &f, std::vector<char> &stack)
{
    // the first switch to the fiber pops six registers and returns into trampoline with the stack as after a call
    uintptr_t top = (reinterpret_cast<uintptr_t>(stack.data()) + stack.size()) & ~uintptr_t(15);
    void **sp = reinterpret_cast<void **>(top - 16);
    *sp = reinterpret_cast<void *>(&trampoline);
    for (int i = 0; i < 6; ++i) *--sp = nullptr;
    f.sp = sp;
}
#else
static void yield() { swapcontext(&fibers[current].ctx, &scheduler); }
static void trampoline() { body(); fibers[current].done = true; swapcontext(&fibers[current].ctx, &scheduler); }
static void resume(int t) { current = t; swapcontext(&scheduler, &fibers[t].ctx); }
static void prepare(Fiber &f, std::vector<char> &stack)
{
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = stack.data();
    f.ctx.uc_stack.ss_size = stack.size();
    f.ctx.uc_link = &scheduler;
    makecontext(&f.ctx, trampoline, 0);
}
#endif
static void warpRendezvous(unsigned mask = 0xffffffffu)
{
    WarpState &w = warps[current >> 5];
    MaskState *m = nullptr;
    for (MaskState &c : w.masks)
        if (c.mask == mask || c.mask == 0) { m = &c; break; }
    if (!m) { fprintf(stderr, "emu: more than 16 distinct lane masks in one warp\n"); abort(); }
    m->mask = mask;
    const unsigned gen = m->generation;
    if (++m->arrived == __builtin_popcount(mask)) { m->arrived = 0; ++m->generation; }
    else while (m->generation == gen) yield();
}
static void blockRendezvous()
{
    const unsigned gen = blockGeneration;
    if (++blockArrived == threads) { blockArrived = 0; ++blockGeneration; }
    else while (blockGeneration == gen) yield();
}
static uint32_t exchange(unsigned mask, uint32_t v, int src)
{
    WarpState &w = warps[current >> 5];
    w.buf[current & 31] = v;
    warpRendezvous(mask);
    const uint32_t r = w.buf[src & 31];
    warpRendezvous(mask);
    return r;
}
} // namespace emu

static uint3 blockIdx = {0, 0, 0};
static dim3 blockDim(1), gridDim(1);
struct ThreadIdxProxy { struct X { operator unsigned() const { return (unsigned)emu::current; } } x; };
static ThreadIdxProxy threadIdx;

// run `kernel` (a closure over its arguments) for every block of the grid
template <class F> static void emuLaunch(int grid, int block, F kernel)
{
    static F *closure; closure = &kernel;
    emu::body = [] { (*closure)(); };
    gridDim = dim3(grid); blockDim = dim3(block); emu::threads = block;
    for (int b = 0; b < grid; ++b)
    {
        blockIdx.x = b;
        emu::fibers.assign(block, emu::Fiber());
        emu::warps.assign((block + 31) / 32, emu::WarpState());
        emu::blockArrived = 0;
        if ((int)emu::stacks.size() < block) emu::stacks.resize(block);
        for (int t = 0; t < block; ++t)
        {
            emu::Fiber &f = emu::fibers[t];
            if (emu::stacks[t].empty()) emu::stacks[t].resize(256 * 1024);
            emu::prepare(f, emu::stacks[t]);
        }
        std::vector<int> order(block);
        for (int t = 0; t < block; ++t) order[t] = emu::scheduleMode == 1 ? block - 1 - t : t;
        for (bool live = true; live;)
        {
            live = false;
            if (emu::scheduleMode == 2)
                for (int t = block - 1; t > 0; --t)
                {
                    emu::scheduleState = emu::scheduleState * 1664525u + 1013904223u;
                    std::swap(order[t], order[(emu::scheduleState >> 8) % (unsigned)(t + 1)]);
                }
            for (int k = 0; k < block; ++k)
            {
                const int t = order[k];
                if (!emu::fibers[t].done) { live = true; emu::resume(t); }
            }
        }
    }
}

extern "C" void emu_set_schedule(int mode, unsigned seed) { emu::scheduleMode = mode; emu::scheduleState = seed ? seed : 1u; }

// ---- warp and block primitives ------------------------------------------------------------------------------------------
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warpRendezvous(mask); }
static inline void __syncthreads() { emu::blockRendezvous(); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); } // one fiber runs at a time; the host side reads after the launch returns
static inline int __shfl_sync(unsigned mask, int v, int src) { return (int)emu::exchange(mask, (uint32_t)v, src); }
static inline unsigned __shfl_sync(unsigned mask, unsigned v, int src) { return emu::exchange(mask, v, src); }
static inline int __shfl_xor_sync(unsigned mask, int v, int m) { return (int)emu::exchange(mask, (uint32_t)v, (emu::current & 31) ^ m); }
static inline unsigned __shfl_xor_sync(unsigned mask, unsigned v, int m) { return emu::exchange(mask, v, (emu::current & 31) ^ m); }
static inline int __shfl_up_sync(unsigned mask, int v, unsigned delta) { const int lane = emu::current & 31; const int r = (int)emu::exchange(mask, (uint32_t)v, lane >= (int)delta ? lane - (int)delta : lane); return r; }
static inline int __shfl_down_sync(unsigned mask, int v, unsigned delta) { const int lane = emu::current & 31; return (int)emu::exchange(mask, (uint32_t)v, lane + (int)delta < 32 ? lane + (int)delta : lane); }
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
    emu::WarpState &w = emu::warps[emu::current >> 5];
    w.buf[emu::current & 31] = pred != 0;
    emu::warpRendezvous(mask);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i)
        if (mask >> i & 1) r |= w.buf[i] << i;
    emu::warpRendezvous(mask);
    return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
static inline int atomicAdd(int *p, int v) { const int old = *p; *p += v; return old; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { const unsigned long long old = *p; *p += v; return old; }

// shared-memory addresses of the asynchronous-copy wrappers are offsets into the block's arena
static inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)(static_cast<const unsigned char *>(p) - emu::sharedArena); }

// ---- scalar intrinsics ----------------------------------------------------------------------------------------------------
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t sel)
{
    const uint64_t both = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((both >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
static inline int __dp4a(int a, int b, int c) { for (int i = 0; i < 4; ++i) c += (int)(int8_t)(a >> (8 * i)) * (int)(int8_t)(b >> (8 * i)); return c; }
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) { for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xff) * ((b >> (8 * i)) & 0xff); return c; }
static inline int __dp2a_lo(int a, int b, int c) { return c + (int)(int16_t)(a & 0xffff) * (int)(int8_t)(b & 0xff) + (int)(int16_t)(a >> 16) * (int)(int8_t)((b >> 8) & 0xff); }
static inline int __dp2a_hi(int a, int b, int c) { return c + (int)(int16_t)(a & 0xffff) * (int)(int8_t)((b >> 16) & 0xff) + (int)(int16_t)(a >> 16) * (int)(int8_t)((b >> 24) & 0xff); }
static inline unsigned __sad(int a, int b, unsigned c) { return c + (unsigned)std::abs(a - b); }
static inline unsigned __vsadu2(unsigned a, unsigned b) { return (unsigned)std::abs((int)(a & 0xffff) - (int)(b & 0xffff)) + (unsigned)std::abs((int)(a >> 16) - (int)(b >> 16)); }
static inline unsigned __vsadu4(unsigned a, unsigned b) { unsigned r = 0; for (int i = 0; i < 4; ++i) r += (unsigned)std::abs((int)((a >> (8 * i)) & 0xff) - (int)((b >> (8 * i)) & 0xff)); return r; }
static inline unsigned __vabsdiffu2(unsigned a, unsigned b) { return (unsigned)std::abs((int)(a & 0xffff) - (int)(b & 0xffff)) | ((unsigned)std::abs((int)(a >> 16) - (int)(b >> 16)) << 16); }
static inline unsigned __vabsdiffu4(unsigned a, unsigned b) { unsigned r = 0; for (int i = 0; i < 4; ++i) r |= (unsigned)std::abs((int)((a >> (8 * i)) & 0xff) - (int)((b >> (8 * i)) & 0xff)) << (8 * i); return r; }
static inline int __vimin_s32_relu(int a, int b) { return std::max(std::min(a, b), 0); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int __mul24(int a, int b) { return a * b; }
static inline long long __mul64hi(long long a, long long b) { return (long long)(((__int128)a * b) >> 64); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }

// ---- inline-PTX wrappers, emulated ----------------------------------------------------------------------------------------
// mma.sync.aligned.m16n8k32.row.col.s32.s8.{u8|s8}.s32 (PTX ISA, "Matrix Fragments for mma.m16n8k32"): lane = 4 g + t holds
//   A (16 x 32, s8, row):  a_r bytes j = A[g + 8 (r & 1)][4 t + 16 (r >> 1) + j]
//   B (32 x 8,  col):      b_r bytes j = B[4 t + 16 r + j][g]
//   C / D (16 x 8, s32):   c_i        = C[g + 8 (i >> 1)][2 t + (i & 1)]
static inline void emuMma16832(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1, bool signedB)
{
    emu::WarpState &w = emu::warps[emu::current >> 5];
    const int lane = emu::current & 31, g = lane >> 2, t = lane & 3;
    uint64_t *mine = w.wide[lane];
    mine[0] = a0, mine[1] = a1, mine[2] = a2, mine[3] = a3, mine[4] = b0, mine[5] = b1;
    emu::warpRendezvous();
    for (int i = 0; i < 4; ++i)
    {
        const int row = g + 8 * (i >> 1), col = 2 * t + (i & 1);
        int sum = c[i];
        for (int k = 0; k < 32; ++k)
        {
            // A[row][k]: lane (row & 7, (k & 15) >> 2), register (row >> 3) + 2 (k >> 4), byte k & 3
            const uint32_t areg = (uint32_t)w.wide[4 * (row & 7) + ((k & 15) >> 2)][(row >> 3) + 2 * (k >> 4)];
            // B[k][col]: lane (col, (k & 15) >> 2), register k >> 4, byte k & 3
            const uint32_t breg = (uint32_t)w.wide[4 * col + ((k & 15) >> 2)][4 + (k >> 4)];
            const int bval = (int)((breg >> (8 * (k & 3))) & 0xff);
            sum += (int)(int8_t)(areg >> (8 * (k & 3))) * (signedB ? (int)(int8_t)bval : bval);
        }
        c[i] = sum;
    }
    emu::warpRendezvous();
}
namespace hvb_unit {
// dp4a.u32.s32: unsigned bytes of a times signed bytes of b
static inline int dp4aUS(uint32_t a, uint32_t b, int c) { for (int i = 0; i < 4; ++i) c += (int)((a >> (8 * i)) & 0xff) * (int)(int8_t)(b >> (8 * i)); return c; }
static inline void imma16832(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    emuMma16832(c, a0, a1, a2, a3, b0, b1, false);
}
} // namespace hvb_unit
'''


def _strip_function(text: str, signature_start: str) -> str:
    """remove the definition that starts with `signature_start` (up to the closing brace of its body)"""
    at = text.index(signature_start)
    depth, i = 0, text.index("{", at)
    while True:
        depth += text[i] == "{"
        depth -= text[i] == "}"
        i += 1
        if depth == 0:
            break
    return text[:at] + text[i:]


def device_helpers() -> str:
    """csrc/hvb_internal.cuh: the structs the kernels see and the section of device helpers"""
    internal = (CSRC / "hvb_internal.cuh").read_text()
    start = internal.index("struct HvbPlane\n{")
    plane = internal[start:internal.index("};", start) + 2]
    dev = internal[internal.index("#ifdef __CUDACC__") + len("#ifdef __CUDACC__"):internal.index("#endif // __CUDACC__")]
    return plane + "\n" + dev


def unit_header() -> str:
    """csrc/hvb_unit.cuh without its includes and without the two PTX wrappers (emulated in the prelude)"""
    text = (CSRC / "hvb_unit.cuh").read_text()
    text = "\n".join(line for line in text.split("\n") if not line.startswith("#include") and not line.startswith("#pragma once"))
    text = _strip_function(text, "__device__ __forceinline__ int dp4aUS(")
    return _strip_function(text, "__device__ __forceinline__ void imma16832(")


def header_text(name: str) -> str:
    """a csrc/*.cuh header without its include / pragma-once lines (its dependencies are already in the translation unit)"""
    text = (CSRC / name).read_text()
    return "\n".join(line for line in text.split("\n") if not line.startswith("#include") and not line.startswith("#pragma once"))


def build(tmp_dir: Path, cu_file: str, entry: str, strip: tuple = (), use_unit_header: bool = True, mma_wrappers: dict | None = None,
          namespaces: int = 1, extra_headers: tuple = (), replace: dict | None = None, check_alignment: bool = False, prelude_structs: str = "") -> C.CDLL:
    """strip: starts of the host-side definitions inside the kernel file's anonymous namespace (launch helpers with <<< >>>);
    mma_wrappers: the file's own inline-PTX mma wrappers, name -> True when the B operand is signed (s8.s8), replaced by
    the emulated product; namespaces: how many leading anonymous namespaces of the file hold the kernels; replace: other
    inline-PTX wrappers of the file, start of the definition -> emulated definition; check_alignment: build with
    -fsanitize=alignment (a misaligned vector access, which x86 would tolerate, aborts the test as it would fault on the
    device) -- for callers whose buffers are aligned the way the device's are"""
    if not (CUDA_INC / "cuda_runtime.h").exists():
        pytest.skip("CUDA headers not found")
    src = (CSRC / cu_file).read_text()
    end = 0
    for _ in range(namespaces):
        end = src.index("} // namespace", end) + len("} // namespace")
    kernels = src[src.index("namespace {"):end]
    for signature in strip:
        kernels = _strip_function(kernels, signature)
    injected = ""
    for signature, replacement in (replace or {}).items():  # further inline-PTX wrappers: definition start -> emulated definition
        kernels = _strip_function(kernels, signature)
        injected += replacement + "\n"
    for name, signed_b in (mma_wrappers or {}).items():
        kernels = _strip_function(kernels, f"__device__ __forceinline__ void {name}(")
        injected += (f"static inline void {name}(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)"
                     f" {{ emuMma16832(c, a0, a1, a2, a3, b0, b1, {'true' if signed_b else 'false'}); }}\n")
    assert "<<<" not in kernels, "a host-side launch is left in the kernel text: add it to `strip`"
    assert "asm volatile" not in kernels and "asm(" not in kernels, f"{cu_file} has inline PTX outside the emulated wrappers"
    body = (device_helpers() + "\n" + prelude_structs + ("\n" + unit_header() if use_unit_header else "") + "\n" + "\n".join(header_text(h) for h in extra_headers)
            + "\n" + injected + kernels)
    # dynamic shared memory is one block-wide arena; static __shared__ arrays become block-wide statics
    body = re.sub(r"extern\s+__shared__\s+(__align__\(\d+\)\s+)?(\w[\w\s]*?)\s+(\w+)\[\];", r"\2 *const \3 = reinterpret_cast<\2 *>(emu::sharedArena);", body)
    body = body.replace("__shared__", "static")
    (tmp_dir / "emu_warp.cpp").write_text(PRELUDE + body + entry)
    sanitize = ["-fsanitize=alignment", "-fno-sanitize-recover=alignment", "-static-libubsan"] if check_alignment else []
    res = subprocess.run(["g++", "-O1", "-fPIC", "-shared", "-w", "-std=c++17", *sanitize, f"-I{CUDA_INC}", f"-I{ROOT / 'include'}",
                          str(tmp_dir / "emu_warp.cpp"), "-o", str(tmp_dir / "emu_warp.so")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    return C.CDLL(str(tmp_dir / "emu_warp.so"))
