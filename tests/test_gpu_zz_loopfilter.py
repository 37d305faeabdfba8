"""GPU parity of the in-loop filters' pixel passes (hvb_deblock_batch, hvb_sao_batch) against the
oracle, which tests/test_oracle_pin_loopfilter.py pins against the reference's own LoopFilter templates: deblocking as
whole-picture passes (all vertical edges, then all horizontal edges) and as the per-CTU regions of TaskDeblock::run in two
batches; SAO from a copy of the picture into the picture, in two CTU ranges; 8 and 10 bit.

First run on a B200 at the end of round 1 (driver run, GPUTEST_r01.json: bit-exact as written); a plain parity test since."""
import numpy as np
import pytest

import test_oracle_pin_loopfilter as pin
from turingcodec_b200 import hvb

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
def test_sao_matches_oracle(oracle, bps, bit_depth):
    """hvb_sao_batch: SAO of a copy of the deblocked picture into the reconstructed picture, two CTU ranges per picture"""
    rng = np.random.default_rng(700 + bit_depth)
    wc, hc = -(-pin.W >> pin.CTB_LOG2), -(-pin.H >> pin.CTB_LOG2)
    ctx = hvb.Context(0, bps, bit_depth)
    try:
        src_pic, dst_pic = ctx.picture_create(pin.W, pin.H, 16), ctx.picture_create(pin.W, pin.H, 16)
        for trial in range(6):
            src, views, blocks, stride, ctus = pin.make_sao_case(rng, bps, bit_depth)
            flags = (1, 1) if trial < 4 else ((1, 0) if trial == 4 else (0, 1))
            want = [a.copy() for a in src]
            pin.sao_call(oracle.lib.orc_sao, False, want, src, bps, bit_depth, blocks, stride, ctus, *flags)
            visible = [np.ascontiguousarray(a[v]) for a, v in zip(src, views)]
            upload(ctx, src_pic, visible)
            ctx.picture_pad(src_pic)  # the edge classes read one sample beyond the picture before the undo runs discard the result
            ctx.picture_copy(dst_pic, src_pic)  # as TaskSao: the picture and a copy of it, one filtered from the other
            b, _ = records(blocks, np.zeros((wc * hc, 2), np.int8))
            ctx.deblock_info_upload(dst_pic, b, np.zeros(wc * hc, hvb.deblock_ctu_t), wc, hc, pin.CTB_LOG2)
            ctx.sao_info_upload(dst_pic, np.frombuffer(bytes(ctus), dtype=hvb.sao_ctu_t).copy())
            tasks = np.zeros(2, hvb.sao_task_t)
            tasks["src_pic"], tasks["dst_pic"] = src_pic, dst_pic
            tasks["ctuBegin"], tasks["ctuEnd"] = (0, wc), (wc, wc * hc)
            tasks["lumaFlag"], tasks["chromaFlag"] = flags
            ctx.sao(tasks)
            for c in range(3):
                got = ctx.picture_download(dst_pic, c, visible[c].shape[1], visible[c].shape[0])
                assert np.array_equal(got, want[c][views[c]]), (trial, c)
    finally:
        ctx.close()


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
def test_sao_statistics_match_oracle(oracle, bps, bit_depth):
    """hvb_sao_stats_batch over every CTU and component of a picture against the oracle (pinned against EncSao's own functions)"""
    import ctypes as C
    rng = np.random.default_rng(1000 + bit_depth)
    oracle.lib.orc_sao_stats.argtypes = [C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    ctx = hvb.Context(0, bps, bit_depth)
    try:
        org_pic, rec_pic = ctx.picture_create(pin.W, pin.H, 16), ctx.picture_create(pin.W, pin.H, 16)
        org3, rec3 = [], []
        for c in range(3):
            w, h = (pin.W, pin.H) if c == 0 else (pin.W // 2, pin.H // 2)
            o, r = pin.stats_case(rng, bps, bit_depth, w - 4, h - 4)
            org3.append(np.ascontiguousarray(o))
            rec3.append(np.ascontiguousarray(r))
        upload(ctx, org_pic, org3)
        upload(ctx, rec_pic, rec3)
        tasks, n = [], 1 << pin.CTB_LOG2
        for c in range(3):
            nc, w, h = (n, pin.W, pin.H) if c == 0 else (n // 2, pin.W // 2, pin.H // 2)
            for y0 in range(0, h, nc):
                for x0 in range(0, w, nc):
                    tasks.append((org_pic, rec_pic, c, x0, y0, min(nc, w - x0), min(nc, h - y0), 0))
        t = np.array(tasks, dtype=hvb.sao_stats_task_t)
        out = ctx.sao_stats(t)
        for i, task in enumerate(t):
            o, r = org3[task["cIdx"]], rec3[task["cIdx"]]
            want = np.zeros(104, np.int64)
            at = lambda a: a.ctypes.data + (int(task["y0"]) * a.shape[1] + int(task["x0"])) * a.itemsize  # noqa: E731
            oracle.lib.orc_sao_stats(at(o), o.shape[1], at(r), r.shape[1], int(task["w"]), int(task["h"]), bit_depth - 8, bps, want.ctypes.data)
            got = np.concatenate([np.stack([out[i]["edgeE"], out[i]["edgeCount"]], axis=1).reshape(-1), out[i]["bandE"], out[i]["bandCount"]])
            assert np.array_equal(got, want), (i, tuple(task))
    finally:
        ctx.close()
