"""The 35-mode intra SATD sweep's own source (csrc/hvb_intra.cu: intraSweepKernel8, tensor-core SATD, reference-sample
filtering on the device), executed on the CPU by the warp-level emulator (tests/host_emu_warp.py), against the oracle:
every mode of blocks 4x4 .. 32x32 with derived and with explicit filtered neighbours, 8 and 10 bit -- the comparison of
tests/test_gpu_pred_intra.py::test_intra_satd35_sweep, in the CPU-only suite."""
import ctypes as C

import numpy as np
import pytest

import host_emu_warp
from gpu_common import H, PAD, W
from orc import filter_flag, filtered_neighbours
from test_host_emulated_loopfilter import Plane
from test_host_emulated_me import host_scene
from turingcodec_b200 import hvb

ENTRY = r'''
extern "C" void emu_intra_sweep(const HvbPlane *planes, const void *pool, const hvb_intra_sweep_task *tasks, int n, int32_t *out, int bitDepth,
                                int bps, int grid)
{
    if (bps == 1) emuLaunch(grid, kWarps * 32, [&] { intraSweepKernel8<uint8_t>(planes, static_cast<const uint8_t *>(pool), tasks, n, out, bitDepth); });
    else emuLaunch(grid, kWarps * 32, [&] { intraSweepKernel8<uint16_t>(planes, static_cast<const uint16_t *>(pool), tasks, n, out, bitDepth); });
}
'''


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    return host_emu_warp.build(tmp_path_factory.mktemp("emu_intra"), "hvb_intra.cu", ENTRY, strip=("int gridWarps(hvb_context *ctx",),
                               use_unit_header=False, mma_wrappers={"imma16832": False})


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
@pytest.mark.parametrize("derive_filtered", [True, False])
def test_intra_sweep_kernel_on_cpu_matches_oracle(emu, oracle, bps, bit_depth, derive_filtered):
    scene = host_scene(bps, bit_depth)
    dtype = np.uint8 if bps == 1 else np.uint16
    rng = np.random.default_rng(35)
    span = 4 * 32 + 1
    pool = rng.integers(0, 1 << bit_depth, (8, span)).astype(dtype)
    pool[1::4] = (np.linspace(0, (1 << bit_depth) - 1, span)[None, :]).astype(dtype)  # smooth: triggers the strong filter
    pool[2::4] = (1 << bit_depth) - 1 - (pool[2::4] & 3)
    flat = pool.reshape(-1)
    tasks = []
    for log2n in (2, 3, 4, 5):
        n = 1 << log2n
        for k in range(6):
            tasks.append((log2n, (k + log2n) % 8, int(rng.integers(0, (W - n) // 4 + 1)) * 4, int(rng.integers(0, (H - n) // 4 + 1)) * 4))
    t = np.zeros(len(tasks), hvb.intra_sweep_task_t)
    upload = np.concatenate([flat, np.zeros(len(tasks) * span, dtype)])
    want_nb = []
    for i, (log2n, which, x, y) in enumerate(tasks):
        n = 1 << log2n
        corner = which * span + 2 * 32
        u = flat[corner - 2 * n: corner + 2 * n + 1]
        f = filtered_neighbours(u, n, bit_depth, True).astype(dtype)
        fcorner = flat.size + i * span + 2 * 32
        upload[fcorner - 2 * n: fcorner + 2 * n + 1] = f
        want_nb.append((u, f))
        t[i]["src"]["pic"], t[i]["src"]["cIdx"], t[i]["src"]["x"], t[i]["src"]["y"] = 0, 0, x, y
        t[i]["nb_unfiltered"] = corner
        t[i]["nb_filtered"] = -1 if derive_filtered else fcorner
        t[i]["log2n"], t[i]["cIdx"], t[i]["strong_intra_smoothing"] = log2n, 0, 1
    table = (Plane * 9)()
    for i, pic in enumerate(scene.host):
        for c, a in enumerate(pic):
            pad = PAD if c == 0 else PAD // 2
            table[3 * i + c] = Plane(a.ctypes.data + (pad * a.shape[1] + pad) * a.itemsize, a.shape[1], a.shape[1] - 2 * pad,
                                     a.shape[0] - 2 * pad, pad, 0)
    got = np.full((len(tasks), 35), -1, np.int32)
    emu.emu_intra_sweep(table, C.c_void_p(upload.ctypes.data), C.c_void_p(t.ctypes.data), len(tasks), C.c_void_p(got.ctypes.data), bit_depth, bps, 2)
    src = scene.host[0][0]
    for i, (log2n, which, x, y) in enumerate(tasks):
        n = 1 << log2n
        u, f = want_nb[i]
        at, stride = (y + PAD) * src.shape[1] + (x + PAD), src.shape[1]
        for mode in range(35):
            nb = f if filter_flag(0, mode, n) else u
            pred = np.zeros((n, n), dtype)
            oracle.pred_intra(pred, n, np.ascontiguousarray(nb), 2 * n, mode, log2n, bit_depth, int(log2n < 5))
            lt = 2 if log2n == 2 else 3  # PredictIntraLumaBlock (Reconstruct.cpp:683-701): 4x4 tile for 4x4 blocks, 8x8 tiles otherwise
            tn = 1 << lt
            want = sum(oracle.hadamard_satd(src, at + dy * stride + dx, stride, pred, dy * n + dx, n, lt)
                       for dy in range(0, n, tn) for dx in range(0, n, tn))
            assert got[i][mode] == want, (log2n, mode, which)
