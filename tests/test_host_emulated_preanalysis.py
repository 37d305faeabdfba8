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
