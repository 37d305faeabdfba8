"""GPU parity of hvb_intra_complexity_batch against the oracle (pinned against the reference's
EstimateIntraComplexity::computeSatd8x8 by tests/test_oracle_pin_preanalysis.py).

STATUS: written after round 1's GPU budget was spent; the kernel's own source is bit-exact under host emulation
(tests/test_host_emulated_preanalysis.py) but has not run on a GPU yet.  Sorted last and marked xfail(strict=False) so that
an undiscovered bug cannot mask the verified suite; the marker is to be removed at the first GPU run of round 2."""
import numpy as np
import pytest

import test_host_emulated_preanalysis as emu_test
import test_oracle_pin_preanalysis as pin
from turingcodec_b200 import hvb

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="first GPU run of hvb_intra_complexity_batch (written without GPU access)")]


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
