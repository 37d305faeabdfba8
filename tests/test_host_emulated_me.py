"""The uni-directional motion search's three kernels -- their own source (csrc/hvb_me_small.cu: four searches per warp on
8-lane groups with group-masked collectives; csrc/hvb_me.cu: a warp per larger PU; csrc/hvb_me_subpel.cu: shared
interpolation planes and tensor-core SATD) -- executed on the CPU by the warp-level emulator (tests/host_emu_warp.py) and
chained as hvb_me_search_batch chains them, against the oracle: every output field, all 23 PU shapes, MET exits, wavefront
limits, 8 and 10 bit.  The same comparison as tests/test_gpu_me.py::test_me_search_matches_oracle, in the CPU-only suite."""
import ctypes as C
from types import SimpleNamespace

import numpy as np
import pytest

import host_emu_warp
import test_gpu_me as gpu_me
from gpu_common import H, PAD, W
from test_host_emulated_loopfilter import Plane
from turingcodec_b200 import hvb, synth

ENTRY_SMALL = r'''
extern "C" void emu_me_small(const HvbPlane *planes, const hvb_me_task *tasks, int n, hvb_me_result *out, int bps, int grid)
{
    int cursor = 0;
    if (bps == 1) emuLaunch(grid, kWarps * 32, [&] { meSearchSmallKernel<uint8_t>(planes, tasks, n, out, &cursor); });
    else emuLaunch(grid, kWarps * 32, [&] { meSearchSmallKernel<uint16_t>(planes, tasks, n, out, &cursor); });
}
'''
ENTRY_LARGE = r'''
extern "C" void emu_me_large(const HvbPlane *planes, const hvb_me_task *tasks, int n, hvb_me_result *out, int bitDepth, int bps, int grid)
{
    if (bps == 1) emuLaunch(grid, kWarps * 32, [&] { meSearchKernel<uint8_t>(planes, tasks, n, out, bitDepth); });
    else emuLaunch(grid, kWarps * 32, [&] { meSearchKernel<uint16_t>(planes, tasks, n, out, bitDepth); });
}
'''
ENTRY_BI = r'''
extern "C" void emu_me_bi(const HvbPlane *planes, const hvb_me_bi_task *tasks, int n, hvb_me_bi_result *out, int bitDepth, int bps, int grid)
{
    if (bps == 1) emuLaunch(grid, kWarps * 32, [&] { meBiSearchKernel<uint8_t>(planes, tasks, n, out, bitDepth); });
    else emuLaunch(grid, kWarps * 32, [&] { meBiSearchKernel<uint16_t>(planes, tasks, n, out, bitDepth); });
}
'''
ENTRY_SUBPEL = r'''
extern "C" void emu_me_bi_subpel(const HvbPlane *planes, const hvb_me_bi_task *tasks, int n, hvb_me_bi_result *out, int bitDepth, int bps, int grid)
{
    if (bps == 1) emuLaunch(grid, kWarps * 32, [&] { meSubpelKernel<uint8_t, true>(planes, tasks, n, out, bitDepth); });
    else emuLaunch(grid, kWarps * 32, [&] { meSubpelKernel<uint16_t, true>(planes, tasks, n, out, bitDepth); });
}
extern "C" void emu_me_subpel(const HvbPlane *planes, const hvb_me_task *tasks, int n, hvb_me_result *out, int bitDepth, int bps, int grid)
{
    if (bps == 1) emuLaunch(grid, kWarps * 32, [&] { meSubpelKernel<uint8_t, false>(planes, tasks, n, out, bitDepth); });
    else emuLaunch(grid, kWarps * 32, [&] { meSubpelKernel<uint16_t, false>(planes, tasks, n, out, bitDepth); });
}
'''


@pytest.fixture(scope="module")
def kernels(tmp_path_factory):
    d = tmp_path_factory.mktemp("emu_me")
    libs = []
    for name, entry in (("hvb_me_small.cu", ENTRY_SMALL), ("hvb_me.cu", ENTRY_LARGE + ENTRY_BI), ("hvb_me_subpel.cu", ENTRY_SUBPEL)):
        sub = d / name.replace(".", "_")
        sub.mkdir()
        libs.append(host_emu_warp.build(sub, name, entry))
    return libs


def host_scene(bps, bit_depth):
    """what tests/gpu_common.Scene holds on the host side: three padded synthetic pictures, numbered 0..2"""
    dtype = np.uint8 if bps == 1 else np.uint16
    host = []
    for i in range(3):
        f = [p.astype(dtype) for p in synth.frame(i, W, H, bit_depth)]
        if bps == 2 and i == 1:
            f[0][::7, ::5] = (1 << bit_depth) - 1
        host.append([np.ascontiguousarray(np.pad(pl, PAD if c == 0 else PAD // 2, mode="edge")) for c, pl in enumerate(f)])
    return SimpleNamespace(bps=bps, bd=bit_depth, pics=[0, 1, 2], host=host)


def plane_table(scene):
    table = (Plane * 9)()
    for i, pic in enumerate(scene.host):
        for c, a in enumerate(pic):
            pad = PAD if c == 0 else PAD // 2
            table[3 * i + c] = Plane(a.ctypes.data + (pad * a.shape[1] + pad) * a.itemsize, a.shape[1], a.shape[1] - 2 * pad,
                                     a.shape[0] - 2 * pad, pad, 0)
    return table


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
def test_me_bi_search_kernels_on_cpu_match_oracle(kernels, oracle, bps, bit_depth):
    """searchMotionBi: the integer grid (meBiSearchKernel), then both fractional rounds (meSubpelKernel<Sample, true>)"""
    _, large, subpel = kernels
    scene = host_scene(bps, bit_depth)
    table = plane_table(scene)
    tasks, otasks = gpu_me.make_bi_tasks(np.random.default_rng(77), scene, 69)
    got = np.zeros(tasks.size, hvb.me_bi_result_t)
    t_ptr, o_ptr = C.c_void_p(tasks.ctypes.data), C.c_void_p(got.ctypes.data)
    large.emu_me_bi(table, t_ptr, tasks.size, o_ptr, bit_depth, bps, 2)
    subpel.emu_me_bi_subpel(table, t_ptr, tasks.size, o_ptr, bit_depth, bps, 2)
    assert gpu_me.check_bi_results(oracle, scene, otasks, got) > 5


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
def test_me_search_kernels_on_cpu_match_oracle(kernels, oracle, bps, bit_depth):
    small, large, subpel = kernels
    scene = host_scene(bps, bit_depth)
    table = plane_table(scene)
    rng = np.random.default_rng(51)
    n = 92
    tasks = np.zeros(n, hvb.me_task_t)
    for i in range(n):
        tasks[i] = gpu_me.make_task(rng, scene, i)
    got = np.zeros(n, hvb.me_result_t)
    t_ptr, o_ptr = C.c_void_p(tasks.ctypes.data), C.c_void_p(got.ctypes.data)
    small.emu_me_small(table, t_ptr, n, o_ptr, bps, 2)           # PUs up to 8x8
    large.emu_me_large(table, t_ptr, n, o_ptr, bit_depth, bps, 2)  # the larger ones
    subpel.emu_me_subpel(table, t_ptr, n, o_ptr, bit_depth, bps, 2)
    early = refined = 0
    for i in range(n):
        r = gpu_me.oracle_search(oracle, scene, tasks[i])
        g = got[i]
        key = (i, tuple(tasks[i][["x0", "y0", "w", "h"]]))
        assert (int(g["mv"]["x"]), int(g["mv"]["y"])) == tuple(r.mv), key
        assert (int(g["mvd"]["x"]), int(g["mvd"]["y"])) == tuple(r.mvd), key
        assert (int(g["mvInteger"]["x"]), int(g["mvInteger"]["y"])) == tuple(r.mvInteger), key
        assert int(g["mvpFlag"]) == r.mvpFlag and int(g["cost"]) == r.cost, key
        assert int(g["subpelCost"]) == r.subpelCost, key
        assert int(g["nSad"]) == r.nSad, key
        assert (int(g["flags"]) & 1) == r.earlyExit, key
        if not r.earlyExit:
            assert list(g["costMvdZero"]) == list(r.costMvdZero), key
        early += r.earlyExit
        refined += tuple(r.mv) != tuple(r.mvInteger)
    assert early > 2 and refined > 5, (early, refined)
