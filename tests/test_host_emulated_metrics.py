"""The batched metric kernels' own source (csrc/hvb_metrics.cu: SAD, SAD4, SSD, and the streaming SATD with its cp.async
staging queue and tensor-core products), executed on the CPU by the warp-level emulator (tests/host_emu_warp.py; the three
asynchronous-copy wrappers become synchronous copies into the block's shared arena), against the oracle: all PU sizes,
candidates reaching into the padding at arbitrary alignment, luma and chroma, 8 and 10 bit -- tests/test_gpu_metrics.py's
comparisons in the CPU-only suite."""
import ctypes as C

import numpy as np
import pytest

import host_emu_warp
import test_gpu_metrics as gm
from test_host_emulated_loopfilter import Plane
from turingcodec_b200 import hvb, synth

ENTRY = r'''
extern "C" void emu_metric(int which, const HvbPlane *planes, const hvb_metric_task *tasks, int n, int32_t *out, int bps, int grid)
{
    if (which == 0)
    {
        if (bps == 1) emuLaunch(grid, kWarpsPerBlock * 32, [&] { sadKernel<uint8_t>(planes, tasks, n, out); });
        else emuLaunch(grid, kWarpsPerBlock * 32, [&] { sadKernel<uint16_t>(planes, tasks, n, out); });
    }
    else if (which == 1)
    {
        if (bps == 1) emuLaunch(grid, kWarpsPerBlock * 32, [&] { ssdKernel<uint8_t>(planes, tasks, n, reinterpret_cast<uint32_t *>(out)); });
        else emuLaunch(grid, kWarpsPerBlock * 32, [&] { ssdKernel<uint16_t>(planes, tasks, n, reinterpret_cast<uint32_t *>(out)); });
    }
    else
    {
        // hvb_satd_batch: the tensor-core kernels for 8-bit blocks tiled 8x8 (streaming, small blocks), then the register-tile kernel for what they left
        int leftover[3] = {0, 0, 0}; // [0]: blocks for satdKernel, [2]: blocks for satdMmaSmallKernel
        if (bps == 1)
        {
            emuLaunch(grid, kWarpsPerBlock * 32, [&] { satdMmaKernel<2>(planes, tasks, n, out, leftover); });
            emuLaunch(grid, kWarpsPerBlock * 32, [&] { satdMmaSmallKernel(planes, tasks, n, out, leftover); });
            emuLaunch(grid, kWarpsPerBlock * 32, [&] { satdKernel<uint8_t>(planes, tasks, n, out, leftover); });
        }
        else
        {
            emuLaunch(grid, kWarpsPerBlock * 32, [&] { satdMma16Kernel<2>(planes, tasks, n, out, leftover); });
            emuLaunch(grid, kWarpsPerBlock * 32, [&] { satdKernel<uint16_t>(planes, tasks, n, out, leftover); });
        }
    }
}
'''
REPLACE = {
    "__device__ __forceinline__ void cpAsync8(": "static inline void cpAsync8(uint32_t dst, const void *src) { memcpy(emu::sharedArena + dst, src, 8); }",
        "__device__ __forceinline__ void cpAsync4(": "static inline void cpAsync4(uint32_t dst, const void *src) { memcpy(emu::sharedArena + dst, src, 4); }",
        "__device__ __forceinline__ void cpAsync16(": "static inline void cpAsync16(uint32_t dst, const void *src) { memcpy(emu::sharedArena + dst, src, 16); }",
    "__device__ __forceinline__ void cpAsyncCommit(": "static inline void cpAsyncCommit() {}",
    "template <int PENDING>\n__device__ __forceinline__ void cpAsyncWait(": "template <int PENDING> static inline void cpAsyncWait() {}",
}


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    return host_emu_warp.build(tmp_path_factory.mktemp("emu_metrics"), "hvb_metrics.cu", ENTRY, strip=("template <typename Task>\nint gridFor(",),
                               extra_headers=("hvb_satd.cuh",), namespaces=2, replace=REPLACE)


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
@pytest.mark.parametrize("chroma", [False, True])
def test_metric_kernels_on_cpu_match_oracle(emu, oracle, bps, bit_depth, chroma):
    dtype = np.uint8 if bps == 1 else np.uint16
    frames = [[p.astype(dtype) for p in synth.frame(i, gm.W, gm.H, bit_depth)] for i in range(2)]
    if bps == 2:
        frames[1][0][::7, ::5] = 1023
    host = [[np.ascontiguousarray(gm.padded(pl, gm.PAD if c == 0 else gm.PAD // 2)) for c, pl in enumerate(f)] for f in frames]
    table = (Plane * 6)()
    for i, pic in enumerate(host):
        for c, a in enumerate(pic):
            pad = gm.PAD if c == 0 else gm.PAD // 2
            table[3 * i + c] = Plane(a.ctypes.data + (pad * a.shape[1] + pad) * a.itemsize, a.shape[1], a.shape[1] - 2 * pad,
                                     a.shape[0] - 2 * pad, pad, 0)
    rng = np.random.default_rng(5 + chroma)
    t = gm.make_tasks(rng, 150, chroma)
    t["b"]["x"][::5] = t["a"]["x"][::5]  # some co-located, aligned pairs: the 128-bit SAD / SSD path and the cp.async SATD path
    for which, name in enumerate(("sad", "ssd", "satd")):
        got = np.full(t.size, -1, np.int32)
        emu.emu_metric(which, table, C.c_void_p(t.ctypes.data), t.size, C.c_void_p(got.ctypes.data), bps, 2)
        for i in range(t.size):
            c = int(t[i]["a"]["cIdx"])
            a, oa, sa = gm.view(host, 0, c, t[i]["a"]["x"], t[i]["a"]["y"])
            b, ob, sb = gm.view(host, 1, c, t[i]["b"]["x"], t[i]["b"]["y"])
            w, h = int(t[i]["w"]), int(t[i]["h"])
            if name == "sad":
                want = oracle.sad(a, oa, sa, b, ob, sb, w, h)
            elif name == "ssd":
                want = oracle.ssd(a, oa, sa, b, ob, sb, w, h)
            else:
                want = oracle.measure_satd(a, oa, sa, b, ob, sb, w, h)
            assert int(got[i]) & 0xffffffff == int(want) & 0xffffffff, (name, i, w, h, c)


@pytest.mark.parametrize("count", [1, 31, 32, 33, 97, 300])
def test_small_block_satd_groups_across_tasks(emu, oracle, count):
    """satdMmaSmallKernel: batches dominated by blocks of one to eight tiles (incl. three tiles in a row), mixed with larger
    and with not-8x8-tiled ones, at batch sizes around the 32-task chunk"""
    frames = [[p.astype(np.uint8) for p in synth.frame(i, gm.W, gm.H, 8)] for i in range(2)]
    host = [[np.ascontiguousarray(gm.padded(pl, gm.PAD if c == 0 else gm.PAD // 2)) for c, pl in enumerate(f)] for f in frames]
    table = (Plane * 6)()
    for i, pic in enumerate(host):
        for c, a in enumerate(pic):
            pad = gm.PAD if c == 0 else gm.PAD // 2
            table[3 * i + c] = Plane(a.ctypes.data + (pad * a.shape[1] + pad) * a.itemsize, a.shape[1], a.shape[1] - 2 * pad,
                                     a.shape[0] - 2 * pad, pad, 0)
    rng = np.random.default_rng(count)
    sizes = [(8, 8)] * 6 + [(16, 8), (8, 16), (16, 16), (24, 8), (8, 24), (32, 8), (64, 8), (32, 16), (16, 32), (32, 32), (64, 64), (12, 16), (8, 4)]
    t = np.zeros(count, hvb.metric_task_t)
    for i in range(count):
        w, h = sizes[rng.integers(len(sizes))] if count > 1 else (8, 8)
        t[i]["w"], t[i]["h"] = w, h
        t[i]["a"]["pic"], t[i]["b"]["pic"] = 0, 1
        t[i]["a"]["x"], t[i]["a"]["y"] = rng.integers(0, (gm.W - w) // 4 + 1) * 4, rng.integers(0, gm.H - h + 1)
        t[i]["b"]["x"] = rng.integers(-gm.PAD + 1, gm.W + gm.PAD - w - 1)
        t[i]["b"]["y"] = rng.integers(-gm.PAD + 1, gm.H + gm.PAD - h - 1)
    if count == 300:
        t["w"][64:128], t["h"][64:128] = 64, 64  # two whole chunks without a small block
    got = np.full(t.size, -1, np.int32)
    emu.emu_metric(2, table, C.c_void_p(t.ctypes.data), t.size, C.c_void_p(got.ctypes.data), 1, 2)
    for i in range(t.size):
        a, oa, sa = gm.view(host, 0, 0, t[i]["a"]["x"], t[i]["a"]["y"])
        b, ob, sb = gm.view(host, 1, 0, t[i]["b"]["x"], t[i]["b"]["y"])
        assert int(got[i]) == oracle.measure_satd(a, oa, sa, b, ob, sb, int(t[i]["w"]), int(t[i]["h"])), (i, int(t[i]["w"]), int(t[i]["h"]))
