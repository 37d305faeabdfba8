"""GPU parity at workload scale: one whole frame pass (every PU search, intra sweep and TU pipeline of
a 640x384 picture, ~55k tasks) through the C-ABI equals the oracle's frame pass bit-for-bit.  This is
bench.py's workload at a size the CPU finishes in seconds, so the measured path is the checked path."""
import ctypes as C
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bit_depth, width, height", [(8, 640, 384), (10, 384, 256)], ids=["8bit", "10bit-u16"])
def test_frame_pass_equals_oracle(bit_depth, width, height):
    """8 bit: BASELINE.json configs[2] at a size the CPU finishes in seconds; 10 bit: the 16-bit sample path of configs[3]."""
    import bench
    from turingcodec_b200 import hvb
    args = SimpleNamespace(width=width, height=height, bit_depth=bit_depth)
    gpu = bench.GpuArm(args, 0)
    gpu.ctx.set_stream(None)
    gpu.step_e2e()  # host-facing path: uploads, three launches, downloads
    levels = gpu.ctx.coeff_download(gpu.fp.coeff_count)

    cpu = bench.CpuArm(args, frames=gpu.frames)
    cpu.oracle.lib.orc_bench_use_port()
    _, _, (o_me, o_intra, o_tu, *_rest) = cpu.run_fraction(1.0)

    for name in ("mv", "mvd", "mvInteger", "mvpFlag", "cost", "costMvdZero", "subpelCost", "nSad", "flags"):
        assert np.array_equal(gpu.h_me[name], o_me[name]), name
    assert np.array_equal(gpu.h_intra, o_intra)
    for name in ("ssd", "ssdPred", "cbf", "status", "sadQuad"):
        bad = np.nonzero((gpu.h_tu[name] != o_tu[name]).reshape(o_tu.size, -1).any(axis=1))[0]
        assert bad.size == 0, (name, bad[:5], gpu.fp.tu[bad[:5]])
    assert np.array_equal(levels, cpu.levels)
    # the pass exercises what it claims: coded and uncoded TUs, early exits and full searches
    assert 0.05 < o_tu["cbf"].mean() < 0.95
    assert 0.05 < (o_me["flags"] & 1).mean() < 0.999
    gpu.ctx.close()


def test_pipelined_frame_pass_equals_blocking():
    """hvb_set_pipelined: the same host-facing pass with copies and kernels overlapped (three steps in flight through the
    ring of staging slots) delivers exactly what the blocking calls deliver."""
    import bench
    args = SimpleNamespace(width=640, height=384)
    gpu = bench.GpuArm(args, 0)
    gpu.ctx.set_stream(None)
    gpu.step_e2e()
    want = {k: getattr(gpu, k).copy() for k in ("h_me", "h_intra", "h_tu")}
    levels = gpu.ctx.coeff_download(gpu.fp.coeff_count)
    for k in ("h_me", "h_intra", "h_tu"):
        getattr(gpu, k).view(np.uint8)[...] = 0xA5
    gpu.ctx.set_pipelined(True)
    for k in range(3):
        gpu.step_e2e(k)  # alternates between the two upload sets: step k+1's copies overlap step k's kernels
    gpu.ctx.sync()
    for k in ("h_me", "h_intra", "h_tu"):
        assert np.array_equal(getattr(gpu, k).view(np.uint8), want[k].view(np.uint8)), k
    gpu.ctx.set_pipelined(False)
    assert np.array_equal(gpu.ctx.coeff_download(gpu.fp.coeff_count), levels)
    gpu.ctx.close()
