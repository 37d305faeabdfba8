"""CPU-only: the bench's frame-pass workload through the oracle port and through the reference's own
(JIT) havoc tables gives identical results -- the CPU baseline arm measures the same computation the
GPU arm is checked against."""
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def test_workload_counts():
    from turingcodec_b200 import synth, workload
    y = synth.frame(0, 128, 64)[0]
    fp = workload.frame_pass(y, 0, 1, (1, 2), tuple(range(3, 9)), n_ctx=4)
    # 5 PUs per CU at 4 depths; partitions 32..4; 2 candidates x (luma + 2 chroma) x CU sizes 32,16,8
    cus = [(128 // s) * (64 // s) for s in (64, 32, 16, 8)]
    assert fp.me.size == 5 * sum(cus)
    assert fp.intra.size == sum((128 // s) * (64 // s) for s in (32, 16, 8, 4))
    assert fp.tu.size == 2 * 3 * sum(cus[1:])
    assert fp.coeff_count == 2 * (128 * 64 + 2 * 64 * 32) * 3
    assert fp.me["limitMax"]["x"].max() <= 128 + 64 and fp.me["limitMin"]["x"].min() >= -64 - 128


def test_reference_tables_equal_port_on_the_bench_workload():
    import orc
    if not orc.have_ref():
        pytest.skip("oracle/_ref not built")
    import bench
    args = SimpleNamespace(width=128, height=64)
    arm = bench.CpuArm(args)
    assert arm.kind == "reference"
    _, _, ref_out = arm.run_fraction(1.0)
    ref_levels = arm.levels.copy()
    arm.oracle.lib.orc_bench_use_port()
    arm.levels[:] = 0
    _, _, port_out = arm.run_fraction(1.0)
    for a, b, name in zip(ref_out[:3], port_out[:3], ("me", "intra", "tu")):
        assert np.array_equal(a, b), name
    assert np.array_equal(ref_levels, arm.levels)
    assert ref_out[0]["nSad"].sum() > 10 * ref_out[0].size  # every search evaluated a non-trivial pattern walk
