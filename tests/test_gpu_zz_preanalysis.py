"""GPU parity of hvb_intra_complexity_batch against the oracle (pinned against the reference's
EstimateIntraComplexity::computeSatd8x8 by tests/test_oracle_pin_preanalysis.py).

First run on a B200 at the end of round 1 (driver run, GPUTEST_r01.json: bit-exact as written); a plain parity test since."""
import numpy as np
import pytest

import test_host_emulated_preanalysis as emu_test
import test_oracle_pin_preanalysis as pin
from turingcodec_b200 import hvb

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
def test_intra_complexity_matches_oracle(oracle, bps, bit_depth):
    rng = np.random.default_rng(90 + bit_depth)
    w, h = 104, 72
    ctx = hvb.Context(0, bps, bit_depth)
    try:
        pic = ctx.picture_create(w, h, 16)
        for picture in pin.pictures(rng, bps, bit_depth, w, h):
            ctx.picture_upload(pic, 0, picture)
            tasks, count = emu_test.region_tasks(pic, w, h)
            got = ctx.intra_complexity(tasks, count)
            assert np.array_equal(got, emu_test.expected(oracle, picture, tasks, count, bps))
    finally:
        ctx.close()


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
def test_aq_activity_and_scd_match_oracle(oracle, bps, bit_depth):
    """hvb_aq_activity_batch / hvb_scd_histogram_batch / hvb_scd_block_stats_batch against the oracle (pinned against the
    reference's AdaptiveQuantisation::preAnalysis and ShotChangeDetection by tests/test_oracle_pin_preanalysis.py): whole and
    clipped units, sizes that are no multiple of 8, doubles compared bit for bit"""
    rng = np.random.default_rng(190 + bit_depth)
    ctx = hvb.Context(0, bps, bit_depth)
    try:
        for w, h in [(136, 112), (90, 70), (128, 72), (640, 360)]:
            pic = ctx.picture_create(w, h, 16)
            for picture in pin.pictures(rng, bps, bit_depth, w, h):
                ctx.picture_upload(pic, 0, picture)
                tasks, count = emu_test.aq_tasks(pic, w, h)
                assert np.array_equal(ctx.aq_activity(tasks, count), emu_test.aq_expected(oracle, picture, tasks, count)), (w, h)
                assert np.array_equal(ctx.scd_histogram([pic])[0], emu_test.scd_histogram_expected(oracle, picture))
                if w * 9 != h * 16:
                    continue
                stats_tasks = np.array([(pic, 1, 0), (pic, 2, 72)], dtype=hvb.scd_stats_task_t)
                stats = ctx.scd_block_stats(stats_tasks, 72 + 32)
                want = np.concatenate([emu_test.scd_expected(oracle, picture, 1), emu_test.scd_expected(oracle, picture, 2)])
                assert stats.tobytes() == want.tobytes()
            ctx.picture_destroy(pic)
    finally:
        ctx.close()


def test_scd_block_stats_refuses_what_the_reference_overreads(oracle):
    """a portrait picture: the reference's `h * height` addressing runs past its byte vector; the library refuses the task"""
    ctx = hvb.Context(0, 1, 8)
    try:
        pic = ctx.picture_create(64, 256, 16)
        with pytest.raises(hvb.HvbError):
            ctx.scd_block_stats(np.array([(pic, 1, 0)], dtype=hvb.scd_stats_task_t), 72)
    finally:
        ctx.close()
