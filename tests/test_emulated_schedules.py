"""A race check without a GPU: under the warp-level emulator a lane runs undisturbed between two collectives, so a missing
__syncwarp / __syncthreads shows as a result that depends on the order in which the lanes are visited.  The frame pass of
smoke(), the PU cost, the bi search, the stand-alone SATD kernels and the in-loop filter kernels are run with the fibers
visited in descending and in pseudo-random order (ascending is what every other emulated test uses) and must stay
bit-exact against the oracle.  (Removing, say, the barrier behind satdMmaSmallKernel's tile map makes this file fail.)"""
import numpy as np
import pytest

import __graft_entry__ as entry
import emu_context
from emu_context import EmuContext, EmuScene


@pytest.fixture(params=[(1, 1), (2, 12345)], ids=["descending", "random"])
def schedule(request):
    emu_context.set_schedule(*request.param)
    yield request.param
    emu_context.set_schedule(0, 1)


def test_frame_pass(schedule):
    entry.smoke_pass(EmuContext(0, 1, 8))


def test_pu_cost_bi_search_and_metrics(schedule, oracle):
    import test_gpu_me as me
    import test_gpu_pu_cost as pu_cost
    scene = EmuScene(1, 8)
    pu_cost.test_pu_cost_stores_the_prediction(scene, oracle)
    tasks, otasks = me.make_bi_tasks(np.random.default_rng(77), scene, 46)
    assert me.check_bi_results(oracle, scene, otasks, scene.ctx.me_bi_search(tasks)) > 3


def test_standalone_satd(schedule, oracle):
    import test_gpu_metrics as gm
    ctx = EmuContext(0, 1, 8)
    from turingcodec_b200 import synth
    frames = [synth.frame(i, gm.W, gm.H, 8) for i in range(2)]
    pics = []
    for f in frames:
        pics.append(ctx.picture_create(gm.W, gm.H, gm.PAD))
        ctx.upload_yuv(pics[-1], *f)
    host = [[gm.padded(np.asarray(pl), gm.PAD if c == 0 else gm.PAD // 2) for c, pl in enumerate(f)] for f in frames]
    gm.test_sad_ssd_satd((ctx, pics, host, 1), oracle, False)


def test_in_loop_filters(schedule, oracle, monkeypatch):
    import test_gpu_zz_loopfilter as lf
    from turingcodec_b200 import hvb
    monkeypatch.setattr(hvb, "Context", EmuContext)
    lf.test_deblock_matches_oracle(oracle, 1, 8)
    lf.test_sao_matches_oracle(oracle, 1, 8)
    lf.test_sao_statistics_match_oracle(oracle, 1, 8)
