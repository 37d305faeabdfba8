"""The intra-complexity kernel's own source (csrc/hvb_preanalysis.cu), executed on the CPU (tests/host_emu.py), against
the oracle (pinned against EstimateIntraComplexity::computeSatd8x8 by tests/test_oracle_pin_preanalysis.py): whole pictures
and sub-regions with their own output offsets, 8- and 16-bit samples."""
import ctypes as C

import numpy as np
import pytest

import host_emu
import test_oracle_pin_preanalysis as pin
from test_host_emulated_loopfilter import Plane, aligned_copy
from turingcodec_b200 import hvb

ENTRY = r'''
extern "C" void emu_intra_complexity(const HvbPlane *planes, const hvb_intra_complexity_task *tasks, int n, int32_t *out, int bps)
{
    if (bps == 1) intraComplexityKernel<uint8_t>(planes, tasks, n, out);
    else intraComplexityKernel<uint16_t>(planes, tasks, n, out);
}
'''


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    return host_emu.build(tmp_path_factory.mktemp("emu_preanalysis"), "hvb_preanalysis.cu", ["HvbPlane"], ENTRY)


def region_tasks(pic, w, h):
    """the whole picture as one task, then its four quadrants as four more, each with its own output range"""
    wb, hb = w // 8, h // 8
    tasks = [(pic, 0, 0, 0, wb, hb, 0)]
    at = wb * hb
    for qy in range(2):
        for qx in range(2):
            x0, y0 = qx * (wb // 2), qy * (hb // 2)
            tw, th = (wb // 2 if qx == 0 else wb - wb // 2), (hb // 2 if qy == 0 else hb - hb // 2)
            tasks.append((pic, 0, 8 * x0, 8 * y0, tw, th, at))
            at += tw * th
    return np.array(tasks, dtype=hvb.intra_complexity_task_t), at


def expected(oracle, pic_array, tasks, count, bps):
    oracle.lib.orc_intra_complexity.argtypes = [C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_int, C.c_void_p]
    want = np.zeros(count, np.int32)
    for t in tasks:
        sub = np.ascontiguousarray(pic_array[t["y0"]:t["y0"] + 8 * t["hBlocks"], t["x0"]:t["x0"] + 8 * t["wBlocks"]])
        out = np.zeros(int(t["wBlocks"]) * int(t["hBlocks"]), np.int32)
        oracle.lib.orc_intra_complexity(sub.ctypes.data, sub.shape[1], sub.shape[1], sub.shape[0], bps, out.ctypes.data)
        want[t["out"]:t["out"] + out.size] = out
    return want


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
def test_intra_complexity_kernel_source_on_cpu_matches_oracle(emu, oracle, bps, bit_depth):
    rng = np.random.default_rng(80 + bit_depth)
    for pic in pin.pictures(rng, bps, bit_depth, 104, 72):
        luma = aligned_copy(pic)
        table = (Plane * 3)(Plane(luma.ctypes.data, luma.strides[0] // luma.itemsize, pic.shape[1], pic.shape[0], 0, 0))
        tasks, count = region_tasks(0, pic.shape[1], pic.shape[0])
        got = np.full(count, -1, np.int32)
        emu.emu_intra_complexity(table, C.c_void_p(tasks.ctypes.data), tasks.size, C.c_void_p(got.ctypes.data), bps)
        assert np.array_equal(got, expected(oracle, pic, tasks, count, bps))


# ---- adaptive-quantisation activity and shot-change detection ------------------------------------------------------------

ENTRY_AQ_SCD = r'''
extern "C" void emu_aq_activity(const HvbPlane *planes, const hvb_aq_layer_task *tasks, int n, long long *out, int bps)
{
    if (bps == 1) aqActivityKernel<uint8_t>(planes, tasks, n, out);
    else aqActivityKernel<uint16_t>(planes, tasks, n, out);
}
extern "C" void emu_scd_histogram(const HvbPlane *planes, const int16_t *pics, int n, int *out, int bps)
{
    if (bps == 1) scdHistogramKernel<uint8_t>(planes, pics, n, out);
    else scdHistogramKernel<uint16_t>(planes, pics, n, out);
}
extern "C" void emu_scd_block_stats(const HvbPlane *planes, const hvb_scd_stats_task *tasks, int n, double *out, int bps)
{
    if (bps == 1) scdBlockStatsKernel<uint8_t>(planes, tasks, n, out);
    else scdBlockStatsKernel<uint16_t>(planes, tasks, n, out);
}
'''


@pytest.fixture(scope="module")
def emu2(tmp_path_factory):
    return host_emu.build(tmp_path_factory.mktemp("emu_preanalysis2"), "hvb_preanalysis.cu", ["HvbPlane"], ENTRY_AQ_SCD)


def aq_tasks(pic, w, h):
    """the four layers of a 64-sample CTU (turing/AdaptiveQuantisation.h:139-142), each with its own output range"""
    tasks, at = [], 0
    for depth in range(4):
        unit = 64 >> depth
        tasks.append((pic, unit, at))
        at += -(-w // unit) * -(-h // unit)
    return np.array(tasks, dtype=hvb.aq_layer_task_t), at


def aq_expected(oracle, picture, tasks, count):
    lib = pin.oracle_aq(oracle)
    want = np.zeros(count, np.int64)
    for t in tasks:
        got, _ = pin.aq_layer(lib, picture, int(t["unit"]))
        want[t["out"]:t["out"] + got.size] = got
    return want


def scd_expected(oracle, picture, margin):
    lib = pin.oracle_aq(oracle)
    h, w = picture.shape
    out = np.zeros(2 * 64, np.float64)
    n = lib.orc_scd_block_stats(picture.ctypes.data, w, w, h, picture.itemsize, margin, out.ctypes.data)
    return out[:2 * n]


def scd_histogram_expected(oracle, picture):
    lib = pin.oracle_aq(oracle)
    h, w = picture.shape
    out = np.zeros(64, np.int32)
    lib.orc_scd_histogram(picture.ctypes.data, w, w, h, picture.itemsize, out.ctypes.data)
    return out


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
def test_aq_and_scd_kernel_source_on_cpu_matches_oracle(emu2, oracle, bps, bit_depth):
    rng = np.random.default_rng(180 + bit_depth)
    for w, h in [(136, 112), (90, 70), (128, 72)]:
        for pic in pin.pictures(rng, bps, bit_depth, w, h):
            luma = aligned_copy(pic)
            table = (Plane * 3)(Plane(luma.ctypes.data, luma.strides[0] // luma.itemsize, w, h, 0, 0))
            tasks, count = aq_tasks(0, w, h)
            got = np.full(count, -7, np.int64)
            emu2.emu_aq_activity(table, C.c_void_p(tasks.ctypes.data), tasks.size, C.c_void_p(got.ctypes.data), bps)
            assert np.array_equal(got, aq_expected(oracle, pic, tasks, count)), (w, h)
            hist = np.zeros(64, np.int32)
            pics = np.zeros(1, np.int16)
            emu2.emu_scd_histogram(table, C.c_void_p(pics.ctypes.data), 1, C.c_void_p(hist.ctypes.data), bps)
            assert np.array_equal(hist, scd_histogram_expected(oracle, pic))
            if (w, h) != (128, 72):
                continue
            stats_tasks = np.array([(0, 1, 0), (0, 2, 72)], dtype=hvb.scd_stats_task_t)
            stats = np.zeros(72 + 32, np.float64)
            emu2.emu_scd_block_stats(table, C.c_void_p(stats_tasks.ctypes.data), 2, C.c_void_p(stats.ctypes.data), bps)
            assert np.array_equal(stats[:72], scd_expected(oracle, pic, 1)) and np.array_equal(stats[72:], scd_expected(oracle, pic, 2))
