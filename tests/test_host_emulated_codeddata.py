"""The coded-residual kernel's own source (csrc/hvb_codeddata.cu), executed on the CPU (tests/host_emu.py), against the
oracle, which tests/test_oracle_pin_codeddata.py pins against the reference's CodedData::storeResidual: a mixed batch of
blocks of every size and scan, all-zero blocks, and a record region too small for the batch."""
import ctypes as C

import numpy as np
import pytest

import host_emu
import test_oracle_pin_codeddata as pin
from turingcodec_b200 import hvb

ENTRY = r'''
extern "C" void emu_coded_residual(int16_t *pool, const hvb_coded_residual_task *tasks, int n, int recordsBase, int capacityWords,
                                   hvb_coded_residual *out)
{
    int cursor = 0;
    codedResidualKernel(pool, reinterpret_cast<uint16_t *>(pool), tasks, n, recordsBase, capacityWords, out, &cursor);
    codedResidualTotalKernel(out, n, recordsBase, capacityWords, &cursor);
}
'''


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    return host_emu.build(tmp_path_factory.mktemp("emu_codeddata"), "hvb_codeddata.cu", [], ENTRY)


def make_batch(rng, count):
    """-> level pool (blocks back to back), tasks, list of (levels, log2n, scanIdx)"""
    blocks, tasks, at = [], np.zeros(count, hvb.coded_residual_task_t), 0
    kinds = ["zero", "single", "sparse", "sparse", "medium", "medium", "dense"]
    for i in range(count):
        log2n, scan = int(rng.integers(2, 6)), int(rng.integers(0, 3))
        lv = pin.make_levels(rng, log2n, kinds[i % len(kinds)])
        tasks[i]["levels"], tasks[i]["log2n"], tasks[i]["scanIdx"] = at, log2n, scan
        blocks.append((lv, log2n, scan))
        at += lv.size
    return np.concatenate([b[0] for b in blocks]), tasks, blocks


def check(oracle, pool, blocks, out, base, capacity):
    """every record equals the oracle's; the records tile the used region without overlap"""
    oracle.lib.orc_coded_residual.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    spans, dropped = [], 0
    for i, (lv, log2n, scan) in enumerate(blocks):
        want = np.zeros(5 + 19 * (1 << (2 * (log2n - 2))), np.uint16)
        n_want = oracle.lib.orc_coded_residual(lv.ctypes.data, log2n, scan, want.ctypes.data)
        if out[i]["words"] == -1:
            dropped += n_want
            continue
        assert out[i]["words"] == n_want, i
        if n_want:
            got = pool[out[i]["offset"]:out[i]["offset"] + n_want].view(np.uint16)
            assert np.array_equal(got, want[:n_want]), (i, log2n, scan)
            spans.append((int(out[i]["offset"]), n_want))
    spans.sort()
    assert all(a + n <= b for (a, n), (b, _) in zip(spans, spans[1:])) and (not spans or spans[0][0] >= base)
    used = sum(n for _, n in spans)
    assert spans == [] or spans[-1][0] + spans[-1][1] <= base + capacity
    return used, dropped


def test_coded_residual_kernel_source_on_cpu_matches_oracle(emu, oracle):
    rng = np.random.default_rng(43)
    levels, tasks, blocks = make_batch(rng, 210)
    base, capacity = levels.size + 16, 200000
    pool = np.concatenate([levels, np.full(16 + capacity, -1, np.int16)])
    out = np.zeros(tasks.size + 1, hvb.coded_residual_t)
    emu.emu_coded_residual(C.c_void_p(pool.ctypes.data), C.c_void_p(tasks.ctypes.data), tasks.size, base, capacity, C.c_void_p(out.ctypes.data))
    used, dropped = check(oracle, pool, blocks, out, base, capacity)
    assert dropped == 0 and out[-1]["offset"] == base + used and out[-1]["words"] == 0
    assert np.array_equal(pool[:levels.size], levels) and (pool[levels.size:base] == -1).all() and (pool[base + used:] == -1).all()


def test_record_region_too_small(emu, oracle):
    rng = np.random.default_rng(44)
    levels, tasks, blocks = make_batch(rng, 60)
    base, capacity = levels.size, 700
    pool = np.concatenate([levels, np.full(capacity + 64, -1, np.int16)])
    out = np.zeros(tasks.size + 1, hvb.coded_residual_t)
    emu.emu_coded_residual(C.c_void_p(pool.ctypes.data), C.c_void_p(tasks.ctypes.data), tasks.size, base, capacity, C.c_void_p(out.ctypes.data))
    used, dropped = check(oracle, pool, blocks, out, base, capacity)
    assert dropped > 0 and (out["words"][:-1] == -1).any()
    assert (pool[base + capacity:] == -1).all()  # nothing written beyond the region
    assert out[-1]["offset"] == base + capacity and out[-1]["words"] > 0
