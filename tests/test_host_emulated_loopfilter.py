"""The device in-loop-filter kernels' own source (deblocking, SAO application, SAO statistics), executed on the CPU
(tests/host_emu.py: a 1-thread grid, g++ against the CUDA host headers), against the pinned oracle: job decomposition,
vector load / store packing, decisions, filters, undo runs and the statistics sums are checked bit-for-bit without a GPU.
tests/test_gpu_zz_loopfilter.py is the same comparison on a B200."""
import ctypes as C

import numpy as np
import pytest

import host_emu
import test_oracle_pin_loopfilter as pin
from turingcodec_b200 import hvb

ENTRY = r'''
extern "C" void emu_deblock(const HvbPlane *planes, const HvbLoopInfo *info, const hvb_deblock_task *tasks, int n, int bitDepth, int bps)
{
    if (bps == 1) deblockKernel<uint8_t>(planes, info, tasks, n, bitDepth);
    else deblockKernel<uint16_t>(planes, info, tasks, n, bitDepth);
}
extern "C" void emu_sao(const HvbPlane *planes, const HvbLoopInfo *info, const hvb_sao_task *tasks, int n, int bitDepth, int bps)
{
    if (bps == 1) saoKernel<uint8_t>(planes, info, tasks, n, bitDepth);
    else saoKernel<uint16_t>(planes, info, tasks, n, bitDepth);
}
extern "C" void emu_sao_stats(const HvbPlane *planes, const hvb_sao_stats_task *tasks, int n, hvb_sao_stats *out, int bitDepth, int bps)
{
    if (bps == 1) saoStatsKernel<uint8_t>(planes, tasks, n, out, bitDepth);
    else saoStatsKernel<uint16_t>(planes, tasks, n, out, bitDepth);
}
extern "C" int emu_sizeof_plane() { return sizeof(HvbPlane); }
extern "C" int emu_sizeof_loop_info() { return sizeof(HvbLoopInfo); }
'''


class Plane(C.Structure):
    _fields_ = [("base", C.c_void_p), ("stride", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("pad", C.c_int32),
                ("reserved", C.c_int32)]


class LoopInfo(C.Structure):
    _fields_ = [("blocks", C.c_void_p), ("ctus", C.c_void_p), ("sao", C.c_void_p), ("blockStride", C.c_int32), ("blockRows", C.c_int32),
                ("widthInCtbs", C.c_int32), ("ctbLog2", C.c_int32), ("saoCount", C.c_int32), ("reserved", C.c_int32)]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    lib = host_emu.build(tmp_path_factory.mktemp("emu_loopfilter"), "hvb_loopfilter.cu", ["HvbPlane", "HvbLoopInfo"], ENTRY)
    assert lib.emu_sizeof_plane() == C.sizeof(Plane) and lib.emu_sizeof_loop_info() == C.sizeof(LoopInfo)
    return lib


def run_kernel(emu, planes, bps, bit_depth, blocks, ctu, ctbs, tasks):
    """planes: arrays whose rows are 256-byte aligned like device pictures (the kernel's vector accesses assume it)"""
    table = (Plane * 3)(*[Plane(p.ctypes.data, p.strides[0] // p.itemsize, p.shape[1], p.shape[0], 0, 0) for p in planes])
    b = np.zeros(blocks.shape[:2], hvb.deblock_block_t)
    b["data"], b["packedBs"] = blocks[..., 0].view(np.int8), blocks[..., 1]
    c = np.zeros(ctu.shape[0], hvb.deblock_ctu_t)
    c["tc_offset_div2"], c["beta_offset_div2"] = ctu[:, 0], ctu[:, 1]
    info = LoopInfo(b.ctypes.data, c.ctypes.data, None, b.shape[1], b.shape[0], ctbs[0], pin.CTB_LOG2, 0, 0)
    tasks = np.ascontiguousarray(tasks, dtype=hvb.deblock_task_t)
    emu.emu_deblock(table, C.byref(info), C.c_void_p(tasks.ctypes.data), tasks.size, bit_depth, bps)


def aligned_copy(plane):
    """a copy whose rows start on 256-byte boundaries (pitch a multiple of 256 bytes), as hvb pictures are laid out"""
    pitch = -(-plane.shape[1] * plane.itemsize // 256) * 256
    raw = np.zeros(plane.shape[0] * pitch + 256, np.uint8)
    off = (-raw.ctypes.data) % 256
    view = raw[off:off + plane.shape[0] * pitch].view(plane.dtype).reshape(plane.shape[0], pitch // plane.itemsize)[:, :plane.shape[1]]
    view[...] = plane
    return view


def task(edge, region, offsets):
    t = np.zeros(1, hvb.deblock_task_t)
    t["pic"], t["edgeType"] = 0, edge
    t["xBegin"], t["yBegin"], t["xEnd"], t["yEnd"] = region
    t["cbQpOffset"], t["crQpOffset"] = offsets
    return t


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10), (2, 9)])
def test_kernel_source_on_cpu_matches_oracle(emu, oracle, bps, bit_depth):
    rng = np.random.default_rng(400 + bit_depth)
    for trial in range(8):
        planes, blocks, ctu, stride, ctbs = pin.make_case(rng, bps, bit_depth)
        offsets = tuple(int(v) for v in rng.integers(-4, 5, 2))
        want = [p.copy() for p in planes]
        whole = [aligned_copy(p) for p in planes]
        for edge in (0, 1):
            pin.call(oracle.lib.orc_deblock, False, want, bps, bit_depth, blocks, ctu, stride, ctbs, offsets, edge, (0, 0, pin.W, pin.H))
            run_kernel(emu, whole, bps, bit_depth, blocks, ctu, ctbs, task(edge, (0, 0, pin.W, pin.H), offsets))
            for c in range(3):
                assert np.array_equal(whole[c], want[c]), (trial, edge, c)
        # TaskDeblock's per-CTU regions, every vertical region in one batch and every horizontal region in the next
        regional = [aligned_copy(p) for p in planes]
        regions = pin.ctu_regions(ctbs)
        for edge in (0, 1):
            run_kernel(emu, regional, bps, bit_depth, blocks, ctu, ctbs, np.concatenate([task(edge, r[edge], offsets) for r in regions]))
        for c in range(3):
            assert np.array_equal(regional[c], want[c]), (trial, "regions", c)


def sao_records(ctus):
    """test_oracle_pin_loopfilter's ctypes records -> hvb.sao_ctu_t (same 42-byte layout)"""
    out = np.frombuffer(bytes(ctus), dtype=hvb.sao_ctu_t).copy()
    assert out.size == len(ctus)
    return out


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10), (2, 9)])
def test_sao_kernel_source_on_cpu_matches_oracle(emu, oracle, bps, bit_depth):
    rng = np.random.default_rng(600 + bit_depth)
    wc, hc = -(-pin.W >> pin.CTB_LOG2), -(-pin.H >> pin.CTB_LOG2)
    for trial in range(8):
        src, views, blocks, stride, ctus = pin.make_sao_case(rng, bps, bit_depth)
        flags = (1, 1) if trial < 6 else ((1, 0) if trial == 6 else (0, 1))
        want = [a.copy() for a in src]
        pin.sao_call(oracle.lib.orc_sao, False, want, src, bps, bit_depth, blocks, stride, ctus, *flags)
        # pictures 0 (source) and 1 (destination), rows 256-byte aligned, margins as in the padded host copies
        pad = pin.SAO_PAD
        host = [aligned_copy(a) for a in src] + [aligned_copy(a) for a in src]
        table = (Plane * 6)()
        for i, a in enumerate(host):
            w, h = a.shape[1] - 2 * pad, a.shape[0] - 2 * pad
            pitch = a.strides[0] // a.itemsize
            assert (a.ctypes.data + (pad * pitch + pad) * a.itemsize) % 8 == 0  # hvb pictures: sample (0,0) of a row is 256-byte aligned
            table[i] = Plane(a.ctypes.data + (pad * pitch + pad) * a.itemsize, pitch, w, h, pad, 0)
        b = np.zeros(blocks.shape[:2], hvb.deblock_block_t)
        b["data"], b["packedBs"] = blocks[..., 0].view(np.int8), blocks[..., 1]
        rec = sao_records(ctus)
        info = (LoopInfo * 2)()
        info[1] = LoopInfo(b.ctypes.data, None, rec.ctypes.data, b.shape[1], b.shape[0], wc, pin.CTB_LOG2, rec.size, 0)
        # two tasks splitting the CTUs, as two TaskSao rows would
        tasks = np.zeros(2, hvb.sao_task_t)
        tasks["src_pic"], tasks["dst_pic"] = 0, 1
        tasks["ctuBegin"], tasks["ctuEnd"] = (0, wc), (wc, wc * hc)
        tasks["lumaFlag"], tasks["chromaFlag"] = flags
        emu.emu_sao(table, info, C.c_void_p(tasks.ctypes.data), tasks.size, bit_depth, bps)
        for c in range(3):
            assert np.array_equal(host[3 + c][views[c]], want[c][views[c]]), (trial, c)


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10), (2, 9)])
def test_sao_statistics_kernel_source_on_cpu_matches_oracle(emu, oracle, bps, bit_depth):
    """every CTU of a picture, three components: the statistics kernel's sums against the oracle's (and thereby EncSao's)"""
    rng = np.random.default_rng(900 + bit_depth)
    oracle.lib.orc_sao_stats.argtypes = [C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    org3, rec3 = [], []
    for c in range(3):
        w, h = (pin.W, pin.H) if c == 0 else (pin.W // 2, pin.H // 2)
        o, r = pin.stats_case(rng, bps, bit_depth, w - 4, h - 4)  # stats_case adds a margin of 2: exactly w x h
        org3.append(aligned_copy(o))
        rec3.append(aligned_copy(r))
    table = (Plane * 6)(*[Plane(a.ctypes.data, a.strides[0] // a.itemsize, a.shape[1], a.shape[0], 0, 0) for a in org3 + rec3])
    tasks = []
    n = 1 << pin.CTB_LOG2
    for c in range(3):
        nc, w, h = (n, pin.W, pin.H) if c == 0 else (n // 2, pin.W // 2, pin.H // 2)
        for y0 in range(0, h, nc):
            for x0 in range(0, w, nc):
                tasks.append((0, 1, c, x0, y0, min(nc, w - x0), min(nc, h - y0), 0))
    t = np.array(tasks, dtype=hvb.sao_stats_task_t)
    out = np.zeros(t.size, hvb.sao_stats_t)
    emu.emu_sao_stats(table, C.c_void_p(t.ctypes.data), t.size, C.c_void_p(out.ctypes.data), bit_depth, bps)
    for i, task in enumerate(t):
        o, r = org3[task["cIdx"]], rec3[task["cIdx"]]
        want = np.zeros(104, np.int64)
        at = lambda a: a.ctypes.data + (int(task["y0"]) * (a.strides[0] // a.itemsize) + int(task["x0"])) * a.itemsize  # noqa: E731
        oracle.lib.orc_sao_stats(at(o), o.strides[0] // o.itemsize, at(r), r.strides[0] // r.itemsize, int(task["w"]), int(task["h"]),
                                 bit_depth - 8, bps, want.ctypes.data)
        got = np.concatenate([np.stack([out[i]["edgeE"], out[i]["edgeCount"]], axis=1).reshape(-1), out[i]["bandE"], out[i]["bandCount"]])
        assert np.array_equal(got, want), (i, tuple(task))
