"""GPU properties of the batched kernels that do not need the oracle: a task's result must not depend on what else is in
the batch.  The second-generation kernels share a warp between several tasks (four PU searches, four 8x8 units from up to
four PUs, 32 RDOQ walks, re-ordered by lastSp) and pad the last group of a batch -- so every prefix of a shuffled,
size-mixed batch has to reproduce the full batch's results, down to batches of 1, 2, 3 and 5 tasks.  Also: the pipelined
host mode falls back to blocking calls for pageable arrays."""
import numpy as np
import pytest

from gpu_common import Scene
from test_gpu_me import make_task
from turingcodec_b200 import hvb, workload

pytestmark = pytest.mark.gpu

PREFIXES = (1, 2, 3, 5, 7, 33, 64)


@pytest.fixture(scope="module", params=[(1, 8), (2, 10)], ids=["u8", "u16-10bit"])
def scene(request):
    s = Scene(*request.param)
    yield s
    s.close()


def test_me_search_is_independent_of_the_batch(scene):
    rng = np.random.default_rng(5)
    n = 150
    tasks = np.zeros(n, hvb.me_task_t)
    for i in range(n):
        tasks[i] = make_task(rng, scene, int(rng.integers(0, 1000)))  # sizes in random order: groups mix PU shapes
    full = scene.ctx.me_search(tasks)
    for k in PREFIXES:
        part = scene.ctx.me_search(np.ascontiguousarray(tasks[:k]))
        assert np.array_equal(part.view(np.uint8), full[:k].view(np.uint8)), k
    # a task alone, from the middle of the batch
    for i in (40, 77, 149):
        one = scene.ctx.me_search(np.ascontiguousarray(tasks[i:i + 1]))
        assert np.array_equal(one.view(np.uint8), full[i:i + 1].view(np.uint8)), i


def test_me_bi_search_is_independent_of_the_batch(scene):
    rng = np.random.default_rng(6)
    n = 90
    uni = np.zeros(n, hvb.me_task_t)
    for i in range(n):
        uni[i] = make_task(rng, scene, int(rng.integers(0, 1000)))
    tasks = np.zeros(n, hvb.me_bi_task_t)
    for name in ("src_pic", "x0", "y0", "w", "h", "mvp", "rateMvpFlag", "limitMin", "limitMax", "halfPel", "quarterPel"):
        tasks[name] = uni[name]
    tasks["ref_pic"], tasks["other_pic"] = scene.pics[1], scene.pics[2]
    tasks["lambda"] = uni["lambda"] // 2
    tasks["mvStart"]["x"], tasks["mvStart"]["y"] = 12 + rng.integers(-9, 10, n), 8 + rng.integers(-9, 10, n)
    tasks["mvOther"]["x"], tasks["mvOther"]["y"] = 24 + rng.integers(-9, 10, n), 16 + rng.integers(-9, 10, n)
    tasks["smallWindow"] = rng.integers(0, 2, n)
    full = scene.ctx.me_bi_search(tasks)
    for k in PREFIXES:
        part = scene.ctx.me_bi_search(np.ascontiguousarray(tasks[:k]))
        assert np.array_equal(part.view(np.uint8), full[:k].view(np.uint8)), k


def test_tu_chain_and_intra_sweep_are_independent_of_the_batch():
    from types import SimpleNamespace

    import bench
    arm = bench.GpuArm(SimpleNamespace(width=256, height=128, double_buffer=False), 0)
    arm.ctx.set_stream(None)
    fp = arm.fp
    rng = np.random.default_rng(7)
    tu = fp.tu[rng.permutation(fp.tu.size)]  # sizes, planes and candidates interleaved
    full = arm.ctx.tu_chain(tu)
    levels = arm.ctx.coeff_download(fp.coeff_count)
    for k in PREFIXES:
        part = arm.ctx.tu_chain(np.ascontiguousarray(tu[:k]))
        assert np.array_equal(part.view(np.uint8), full[:k].view(np.uint8)), k
    assert np.array_equal(arm.ctx.coeff_download(fp.coeff_count), levels)  # re-running prefixes rewrote the same levels
    intra = fp.intra[rng.permutation(fp.intra.size)]
    full = arm.ctx.intra_satd35(intra)
    for k in PREFIXES:
        assert np.array_equal(arm.ctx.intra_satd35(np.ascontiguousarray(intra[:k])), full[:k]), k
    arm.ctx.close()


def test_pipelined_mode_falls_back_for_pageable_arrays(scene):
    rng = np.random.default_rng(8)
    tasks = np.zeros(40, hvb.me_task_t)
    for i in range(40):
        tasks[i] = make_task(rng, scene, i)
    want = scene.ctx.me_search(tasks)
    scene.ctx.set_pipelined(True)
    try:
        got = scene.ctx.me_search(tasks)  # pageable numpy arrays: the call must block and deliver
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
    finally:
        scene.ctx.sync()
        scene.ctx.set_pipelined(False)
