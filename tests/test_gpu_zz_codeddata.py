"""GPU parity of hvb_coded_residual_batch against the oracle (pinned against the reference's
CodedData::storeResidual by tests/test_oracle_pin_codeddata.py): a mixed batch through the coefficient pool.

STATUS: written after round 1's GPU budget was spent; the kernel's own source is bit-exact under host emulation
(tests/test_host_emulated_codeddata.py) but has not run on a GPU yet.  Sorted last and marked xfail(strict=False) so that an
undiscovered bug cannot mask the verified suite; the marker is to be removed at the first GPU run of round 2."""
import numpy as np
import pytest

import test_host_emulated_codeddata as emu_test
from turingcodec_b200 import hvb

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="first GPU run of hvb_coded_residual_batch (written without GPU access)")]


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
