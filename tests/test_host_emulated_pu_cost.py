"""The PU-cost kernel's own source (csrc/hvb_pu_cost.cu: shuffles, warp barriers, tensor-core SATD), executed on the CPU
by the warp-level emulator (tests/host_emu_warp.py: a fiber per CUDA thread, collectives as rendezvous, mma.sync emulated
from the PTX fragment layout), against the oracle -- the same comparison tests/test_gpu_pu_cost.py makes on a B200, here
in the CPU-only suite: all 24 PU shapes, uni / bi, vectors far outside the picture, the stored predictions, 8 and 10 bit."""
import ctypes as C

import numpy as np
import pytest

import host_emu_warp
import orc
import test_oracle_pu_cost_pin as pin
from test_host_emulated_loopfilter import Plane
from turingcodec_b200 import hvb

ENTRY = r'''
extern "C" void emu_pu_cost(const HvbPlane *planes, const hvb_pu_cost_task *tasks, int n, int32_t *out, int bitDepth, int bps, int grid)
{
    int cursor = 0;
    if (bps == 1) emuLaunch(grid, 256, [&] { puCostKernel<uint8_t>(planes, tasks, n, out, bitDepth, &cursor); });
    else emuLaunch(grid, 256, [&] { puCostKernel<uint16_t>(planes, tasks, n, out, bitDepth, &cursor); });
}
'''


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    return host_emu_warp.build(tmp_path_factory.mktemp("emu_pu_cost"), "hvb_pu_cost.cu", ENTRY, strip=("template <typename Sample>\nint launch(",))


def plane_table(pictures, bps):
    """HvbPlane records of padded host pictures (list of [Y, Cb, Cr] arrays padded by PAD / PAD // 2)"""
    table = (Plane * (3 * len(pictures)))()
    for i, pic in enumerate(pictures):
        for c, a in enumerate(pic):
            pad = pin.PAD if c == 0 else pin.PAD // 2
            table[3 * i + c] = Plane(a.ctypes.data + (pad * a.shape[1] + pad) * a.itemsize, a.shape[1], a.shape[1] - 2 * pad,
                                     a.shape[0] - 2 * pad, pad, 0)
    return table


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
def test_pu_cost_kernel_source_on_cpu_matches_oracle(emu, oracle, bps, bit_depth):
    rng = np.random.default_rng(3)
    n = 96
    ref_tasks = (pin.RefPuTask * n)(*[pin.make_pu_task(rng, i) for i in range(n)])
    frames = [pin.padded_frame(bps, bit_depth, k) for k in range(3)]
    dst = [np.zeros_like(a) for a in frames[0]]
    table = plane_table(frames + [dst], bps)
    tasks = np.zeros(n, hvb.pu_cost_task_t)
    for i, r in enumerate(ref_tasks):
        tasks[i]["src_pic"], tasks[i]["dst_pic"] = 0, (3 if i % 4 == 0 and not overlaps(ref_tasks, i) else -1)
        tasks[i]["ref_pic"] = (1 if r.predFlag[0] else -1, 2 if r.predFlag[1] else -1)
        tasks[i]["x0"], tasks[i]["y0"], tasks[i]["w"], tasks[i]["h"] = r.x0, r.y0, r.w, r.h
        tasks[i]["mvx"], tasks[i]["mvy"] = (r.mv[0], r.mv[2]), (r.mv[1], r.mv[3])
    got = np.full((n, 3), -1, np.int32)
    emu.emu_pu_cost(table, C.c_void_p(tasks.ctypes.data), n, C.c_void_p(got.ctypes.data), bit_depth, bps, 2)

    planes = [orc.planes3(f, pin.PAD) for f in frames]
    oracle.lib.orc_pu_cost.argtypes = [C.c_void_p] * 3 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    stored = 0
    for i in range(n):
        t = ref_tasks[i]
        o, satd = pin.oracle_task(t, bit_depth), (C.c_int32 * 3)()
        bufs = [np.zeros(64 * 64, frames[0][0].dtype) for _ in range(3)]
        out = (C.c_void_p * 3)(*[b.ctypes.data for b in bufs])
        oracle.lib.orc_pu_cost(planes[0], planes[1], planes[2], C.byref(o), satd, out, bps)
        assert list(got[i]) == list(satd), (i, (t.x0, t.y0, t.w, t.h), tuple(t.predFlag), tuple(t.mv))
        if tasks[i]["dst_pic"] >= 0:
            stored += 1
            for c in range(3):
                sh, pad = int(c > 0), (pin.PAD if c == 0 else pin.PAD // 2)
                w, h = t.w >> sh, t.h >> sh
                y0, x0 = pad + (t.y0 >> sh), pad + (t.x0 >> sh)
                assert np.array_equal(dst[c][y0:y0 + h, x0:x0 + w], bufs[c][:w * h].reshape(h, w)), (i, c)
    assert stored > 5 and (got[:, 1] == 0).sum() > 3


def overlaps(ref_tasks, i):
    """does PU i overlap an earlier PU that stores its prediction (every 4th)?  Then it does not store."""
    a = ref_tasks[i]
    for j in range(0, i, 4):
        b = ref_tasks[j]
        if a.x0 < b.x0 + b.w and b.x0 < a.x0 + a.w and a.y0 < b.y0 + b.h and b.y0 < a.y0 + a.h:
            return True
    return False
