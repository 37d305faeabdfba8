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
