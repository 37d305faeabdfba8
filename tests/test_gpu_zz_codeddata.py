"""GPU parity of hvb_coded_residual_batch against the oracle (pinned against the reference's
CodedData::storeResidual by tests/test_oracle_pin_codeddata.py): a mixed batch through the coefficient pool.

First run on a B200 at the end of round 1 (driver run, GPUTEST_r01.json: bit-exact as written); a plain parity test since."""
import numpy as np
import pytest

import test_host_emulated_codeddata as emu_test
from turingcodec_b200 import hvb

pytestmark = pytest.mark.gpu


def test_coded_residual_matches_oracle(oracle):
    rng = np.random.default_rng(45)
    levels, tasks, blocks = emu_test.make_batch(rng, 400)
    base, capacity = levels.size + 32, 400000
    ctx = hvb.Context(0, 1, 8)
    try:
        ctx.coeff_upload(levels, 0)
        out = ctx.coded_residual(tasks, base, capacity)
        end = int(out[-1]["offset"])
        assert out[-1]["words"] == 0 and base <= end <= base + capacity
        pool = ctx.coeff_download(end)
        used, dropped = emu_test.check(oracle, pool, blocks, out, base, capacity)
        assert dropped == 0 and end == base + used
        assert np.array_equal(pool[:levels.size], levels)
    finally:
        ctx.close()
