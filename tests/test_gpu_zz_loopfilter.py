"""GPU parity of the deblocking pixel pass (hvb_deblock_batch) against the oracle, which
tests/test_oracle_pin_loopfilter.py pins against the reference's own LoopFilter templates: whole-picture passes (all
vertical edges, then all horizontal edges) and the per-CTU regions of TaskDeblock::run as two batches, 8 and 10 bit.

STATUS: this kernel was written after round 1's GPU budget was spent and has not run on a GPU yet.  The file sorts last
and is marked xfail(strict=False) so that an undiscovered bug cannot mask the verified suite in front of it; the marker
is to be removed at the first GPU run of round 2 (XPASS in the report means the kernel is bit-exact as written)."""
import numpy as np
import pytest

import test_oracle_pin_loopfilter as pin
from turingcodec_b200 import hvb

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="first GPU run of hvb_deblock_batch (written without GPU access)")]


def upload(ctx, pic, planes):
    for c, p in enumerate(planes):
        ctx.picture_upload(pic, c, p)


def records(blocks, ctu):
    b = np.zeros(blocks.shape[:2], hvb.deblock_block_t)
    b["data"], b["packedBs"] = blocks[..., 0].view(np.int8), blocks[..., 1]
    c = np.zeros(ctu.shape[0], hvb.deblock_ctu_t)
    c["tc_offset_div2"], c["beta_offset_div2"] = ctu[:, 0], ctu[:, 1]
    return b, c


def task(pic, edge, region, offsets):
    t = np.zeros(1, hvb.deblock_task_t)
    t["pic"], t["edgeType"] = pic, edge
    t["xBegin"], t["yBegin"], t["xEnd"], t["yEnd"] = region
    t["cbQpOffset"], t["crQpOffset"] = offsets
    return t


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
def test_deblock_matches_oracle(oracle, bps, bit_depth):
    rng = np.random.default_rng(300 + bit_depth)
    ctx = hvb.Context(0, bps, bit_depth)
    try:
        pic = ctx.picture_create(pin.W, pin.H, 16)
        for trial in range(6):
            planes, blocks, ctu, stride, ctbs = pin.make_case(rng, bps, bit_depth)
            offsets = tuple(int(v) for v in rng.integers(-4, 5, 2))
            ctx.deblock_info_upload(pic, *records(blocks, ctu), ctbs[0], ctbs[1], pin.CTB_LOG2)
            # whole-picture passes
            upload(ctx, pic, planes)
            want = [p.copy() for p in planes]
            for edge in (0, 1):
                pin.call(oracle.lib.orc_deblock, False, want, bps, bit_depth, blocks, ctu, stride, ctbs, offsets, edge, (0, 0, pin.W, pin.H))
                ctx.deblock(task(pic, edge, (0, 0, pin.W, pin.H), offsets))
                for c in range(3):
                    got = ctx.picture_download(pic, c, planes[c].shape[1], planes[c].shape[0])
                    assert np.array_equal(got, want[c]), (trial, edge, c)
            # TaskDeblock's per-CTU regions: every vertical region in one batch, every horizontal region in the next
            upload(ctx, pic, planes)
            regions = pin.ctu_regions(ctbs)
            for edge in (0, 1):
                ctx.deblock(np.concatenate([task(pic, edge, r[edge], offsets) for r in regions]))
            for c in range(3):
                got = ctx.picture_download(pic, c, planes[c].shape[1], planes[c].shape[0])
                assert np.array_equal(got, want[c]), (trial, "regions", c)
    finally:
        ctx.close()
