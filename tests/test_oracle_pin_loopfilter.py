"""Pins the oracle's deblocking pixel pass (oracle/oracle_loopfilter.c) against the UNMODIFIED reference templates --
LoopFilter::Picture::deblock<edgeType> with LumaBlockEdge / ChromaBlockEdge (turing/LoopFilter.h:165-423, :739-777),
driven by oracle/ref_shim_loopfilter.cpp inside oracle/_ref/libhavoc_ref.so -- on blocky pictures with random boundary
strengths, QPs, filter-disable bits (pcm / transquant bypass), slice tc / beta offsets and chroma QP offsets at 8, 9
and 10 bit: whole-picture passes (vertical edges, then horizontal) and the per-CTU regions of TaskDeblock::run
(turing/TaskDeblock.cpp:104-127) must leave identical pictures."""
import ctypes as C

import numpy as np
import pytest

import orc

W, H, CTB_LOG2 = 176, 144, 6  # 3 x 3 CTUs of 64, the last row / column partial


@pytest.fixture(scope="module")
def reflib():
    if not orc.REF_LIB.exists():
        pytest.skip("oracle/_ref/libhavoc_ref.so not built")
    lib = C.CDLL(str(orc.REF_LIB))
    if not hasattr(lib, "ref_deblock"):
        pytest.skip("libhavoc_ref.so predates ref_shim_loopfilter.cpp (make -C oracle ref)")
    return lib


def make_case(rng, bps, bit_depth):
    """-> planes (list of 3 arrays), block records [rows][stride][data, packedBs], CTU offsets, grid stride, CTUs (w, h)"""
    dtype = np.uint8 if bps == 1 else np.uint16
    top = (1 << bit_depth) - 1
    planes = []
    for c in range(3):
        w, h = (W, H) if c == 0 else (W // 2, H // 2)
        n = 8 if c == 0 else 4
        # a common level plus a small step per block (so that edges are real but within reach of tC), and noise of
        # an amplitude drawn per picture: flat pictures take the strong filter, busier ones the normal filter with and
        # without p1 / q1, the busiest leave edges unfiltered (d >= beta)
        step = rng.integers(-6, 7, (h // n + 1, w // n + 1)) << (bit_depth - 8)
        base = int(rng.integers(top // 4, 3 * top // 4)) + np.kron(step, np.ones((n, n), np.int64))[:h, :w]
        amp = int(rng.choice([0, 1, 1, 2, 4])) << (bit_depth - 8)
        noise = rng.integers(-amp, amp + 1, (h, w))
        planes.append(np.clip(base + noise, 0, top).astype(dtype))
    wc, hc = -(-W >> CTB_LOG2), -(-H >> CTB_LOG2)
    stride, rows = ((wc << CTB_LOG2) >> 3) + 1, ((hc << CTB_LOG2) >> 3) + 1  # turing/LoopFilter.h:436-443
    blocks = np.zeros((rows, stride, 2), np.uint8)
    qp = rng.integers(18, 46, (rows, stride))
    disable = rng.random((rows, stride)) < 0.1
    blocks[..., 0] = ((qp << 1) | disable).astype(np.uint8)
    bs = rng.choice([0, 1, 2], (rows, stride, 4), p=[0.3, 0.35, 0.35])  # [vertical pos 0, 1, horizontal pos 0, 1]
    bs[:, 0, 0:2] = 0  # no picture-boundary edges (the encoder never sets them: processCu / neighbour availability)
    bs[0, :, 2:4] = 0
    blocks[..., 1] = (bs[..., 0] | bs[..., 1] << 2 | bs[..., 2] << 4 | bs[..., 3] << 6).astype(np.uint8)
    ctu = rng.integers(-3, 4, (wc * hc, 2)).astype(np.int8)
    return planes, blocks, ctu, stride, (wc, hc)


def test_grid_geometry_is_the_reference_s(reflib):
    stride, rows = C.c_int(), C.c_int()
    wc, hc = -(-W >> CTB_LOG2), -(-H >> CTB_LOG2)
    reflib.ref_deblock_grid(wc, hc, CTB_LOG2, C.byref(stride), C.byref(rows))
    assert (stride.value, rows.value) == (((wc << CTB_LOG2) >> 3) + 1, ((hc << CTB_LOG2) >> 3) + 1)


def ctu_regions(ctbs):
    """the regions TaskDeblock::run filters per CTU (turing/TaskDeblock.cpp:104-127): [(vertical, horizontal)] in raster order"""
    n, out = 1 << CTB_LOG2, []
    for ry in range(ctbs[1]):
        for rx in range(ctbs[0]):
            x0, y0 = rx * n, ry * n
            out.append(((x0 + (8 if rx else 0), y0 + (8 if ry else 0), min(x0 + n + 8, W), min(y0 + n + 8, H)),
                        (x0, y0 + (8 if ry else 0), min(x0 + n, W), min(y0 + n + 8, H))))
    return out


def call(fn, is_ref, planes, bps, bit_depth, blocks, ctu, stride, ctbs, offsets, edge, region):
    ptrs = (C.c_void_p * 3)(*[p.ctypes.data for p in planes])
    strides = (C.c_ssize_t * 3)(*[p.shape[1] for p in planes])
    b, o = blocks.ctypes.data_as(C.c_void_p), ctu.ctypes.data_as(C.c_void_p)
    if is_ref:
        fn(ptrs, strides, bps, bit_depth, bit_depth, ctbs[0], ctbs[1], CTB_LOG2, offsets[0], offsets[1], b, o, edge, *region)
    else:
        fn(ptrs, strides, bps, bit_depth, bit_depth, b, stride, o, ctbs[0], CTB_LOG2, offsets[0], offsets[1], edge, *region)


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10), (2, 9)])
def test_whole_picture_passes_match_reference(reflib, oracle, bps, bit_depth):
    rng = np.random.default_rng(200 + bit_depth)
    changed = strong_like = 0
    for trial in range(12):
        planes, blocks, ctu, stride, ctbs = make_case(rng, bps, bit_depth)
        offsets = tuple(int(v) for v in rng.integers(-4, 5, 2))
        want = [p.copy() for p in planes]
        got = [p.copy() for p in planes]
        for edge in (0, 1):  # H.265 8.7.2: every vertical edge of the picture, then every horizontal edge
            call(reflib.ref_deblock, True, want, bps, bit_depth, blocks, ctu, stride, ctbs, offsets, edge, (0, 0, W, H))
            call(oracle.lib.orc_deblock, False, got, bps, bit_depth, blocks, ctu, stride, ctbs, offsets, edge, (0, 0, W, H))
            for c in range(3):
                assert np.array_equal(got[c], want[c]), (trial, edge, c)
            if edge == 0:  # after the vertical pass, columns 2 mod 8 (q2) can only have moved in the strong filter
                strong_like += int((want[0] != planes[0])[:, 2::8].sum())
        changed += sum(int((w != p).sum()) for w, p in zip(want, planes))
    assert changed > 20000 and strong_like > 200, (changed, strong_like)


def test_ctu_regions_of_task_deblock_match_whole_picture(reflib, oracle):
    """TaskDeblock::run walks the CTUs in raster order, vertical edges of a region shifted by 8 then the horizontal edges
    of the CTU (turing/TaskDeblock.cpp:104-127); the result is the whole-picture two-pass result, for the reference and
    for the oracle alike."""
    rng = np.random.default_rng(77)
    planes, blocks, ctu, stride, ctbs = make_case(rng, 1, 8)
    whole = [p.copy() for p in planes]
    for edge in (0, 1):
        call(oracle.lib.orc_deblock, False, whole, 1, 8, blocks, ctu, stride, ctbs, (1, -2), edge, (0, 0, W, H))
    for fn, is_ref in ((reflib.ref_deblock, True), (oracle.lib.orc_deblock, False)):
        got = [p.copy() for p in planes]
        for ver, hor in ctu_regions(ctbs):
            call(fn, is_ref, got, 1, 8, blocks, ctu, stride, ctbs, (1, -2), 0, ver)
            call(fn, is_ref, got, 1, 8, blocks, ctu, stride, ctbs, (1, -2), 1, hor)
        for c in range(3):
            assert np.array_equal(got[c], whole[c]), (is_ref, c)


# ---- sample adaptive offset ---------------------------------------------------------------------------------------------

class SaoPlane(C.Structure):
    _pack_ = 1
    _fields_ = [("typeIdx", C.c_int8), ("classOrBand", C.c_int8), ("offset", C.c_int16 * 4)]


class SaoCtu(C.Structure):
    _pack_ = 1
    _fields_ = [("left", C.c_int16), ("top", C.c_int16), ("right", C.c_int16), ("bottom", C.c_int16),
                ("topLeft", C.c_uint8), ("topRight", C.c_uint8), ("bottomLeft", C.c_uint8), ("bottomRight", C.c_uint8),
                ("plane", SaoPlane * 3)]


SAO_PAD = 80  # samples of margin around every plane: the reference's whole-CTB passes reach outside a partial CTU


def make_sao_case(rng, bps, bit_depth):
    """-> padded planes (random margins), visible-area slices, block records, CTU records"""
    assert C.sizeof(SaoCtu) == 42
    dtype = np.uint8 if bps == 1 else np.uint16
    top = (1 << bit_depth) - 1
    padded, views = [], []
    for c in range(3):
        w, h = (W, H) if c == 0 else (W // 2, H // 2)
        # smooth texture + noise: every edge category and many bands occur; a few samples at the range limits for the clip
        yy, xx = np.mgrid[0:h + 2 * SAO_PAD, 0:w + 2 * SAO_PAD]
        tex = (top / 2 + top / 4 * np.sin(xx / 7.0) * np.cos(yy / 5.0)).astype(np.int64) + rng.integers(-3, 4, xx.shape) * (1 << (bit_depth - 8))
        tex[rng.random(tex.shape) < 0.01] = top
        tex[rng.random(tex.shape) < 0.01] = 0
        a = np.clip(tex, 0, top).astype(dtype)
        padded.append(a)
        views.append((slice(SAO_PAD, SAO_PAD + h), slice(SAO_PAD, SAO_PAD + w)))
    wc, hc = -(-W >> CTB_LOG2), -(-H >> CTB_LOG2)
    stride, rows = ((wc << CTB_LOG2) >> 3) + 1, ((hc << CTB_LOG2) >> 3) + 1
    blocks = np.zeros((rows, stride, 2), np.uint8)
    blocks[..., 0] = ((rng.integers(20, 40, (rows, stride)) << 1) | (rng.random((rows, stride)) < 0.1)).astype(np.uint8)
    n = 1 << CTB_LOG2
    ctus = (SaoCtu * (wc * hc))()
    for ry in range(hc):
        for rx in range(wc):
            t = ctus[ry * wc + rx]
            x0, y0 = rx * n, ry * n
            # picture limits, or a slice / tile boundary without loop filtering across it (LoopFilter.h:476-534)
            t.left = x0 if (rx and rng.random() < 0.25) else 0
            t.top = y0 if (ry and rng.random() < 0.25) else 0
            t.right = x0 + n if (rx < wc - 1 and rng.random() < 0.25) else W
            t.bottom = y0 + n if (ry < hc - 1 and rng.random() < 0.25) else H
            t.topLeft = int(rx > 0 and ry > 0 and rng.random() < 0.7)
            t.topRight = int(rx < wc - 1 and ry > 0 and rng.random() < 0.7)
            t.bottomLeft = int(rx > 0 and ry < hc - 1 and rng.random() < 0.7)
            t.bottomRight = int(rx < wc - 1 and ry < hc - 1 and rng.random() < 0.7)
            for c in range(3):
                t.plane[c].typeIdx = int(rng.integers(0, 3))
                t.plane[c].classOrBand = int(rng.integers(0, 4)) if t.plane[c].typeIdx == 2 else int(rng.integers(0, 32))
                for k in range(4):
                    t.plane[c].offset[k] = int(rng.integers(-7, 8)) << (bit_depth - min(bit_depth, 10))
    return padded, views, blocks, stride, ctus


def sao_call(fn, is_ref, dst, src, bps, bit_depth, blocks, stride, ctus, luma_flag, chroma_flag):
    def origin(a):
        return a.ctypes.data + (SAO_PAD * a.shape[1] + SAO_PAD) * a.itemsize
    d = (C.c_void_p * 3)(*[origin(a) for a in dst])
    s = (C.c_void_p * 3)(*[origin(a) for a in src])
    strides = (C.c_ssize_t * 3)(*[a.shape[1] for a in src])
    b = blocks.ctypes.data_as(C.c_void_p)
    if is_ref:
        fn(d, s, strides, bps, bit_depth, bit_depth, W, H, CTB_LOG2, b, ctus, luma_flag, chroma_flag)
    else:
        fn(d, s, strides, bps, bit_depth, bit_depth, W, H, CTB_LOG2, b, stride, ctus, luma_flag, chroma_flag)


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10), (2, 9)])
def test_sao_matches_reference(reflib, oracle, bps, bit_depth):
    if not hasattr(reflib, "ref_sao"):
        pytest.skip("libhavoc_ref.so predates ref_sao (make -C oracle ref)")
    rng = np.random.default_rng(500 + bit_depth)
    changed = kept = 0
    for trial in range(10):
        src, views, blocks, stride, ctus = make_sao_case(rng, bps, bit_depth)
        flags = (1, 1) if trial < 8 else ((1, 0) if trial == 8 else (0, 1))  # slice_sao_luma_flag, slice_sao_chroma_flag
        want = [a.copy() for a in src]  # the encoder filters from a copy of the deblocked picture back into it (TaskSao.cpp:96-121)
        got = [a.copy() for a in src]
        sao_call(reflib.ref_sao, True, want, src, bps, bit_depth, blocks, stride, ctus, *flags)
        sao_call(oracle.lib.orc_sao, False, got, src, bps, bit_depth, blocks, stride, ctus, *flags)
        for c in range(3):
            assert np.array_equal(got[c][views[c]], want[c][views[c]]), (trial, c)
            changed += int((want[c][views[c]] != src[c][views[c]]).sum())
            kept += int((want[c][views[c]] == src[c][views[c]]).sum())
    assert changed > 20000 and kept > 100000, (changed, kept)


# ---- SAO statistics (encoder side) ---------------------------------------------------------------------------------------

def stats_case(rng, bps, bit_depth, w, h):
    """original and reconstructed block (reconstruction = original + coding noise), each inside a larger array"""
    dtype = np.uint8 if bps == 1 else np.uint16
    top = (1 << bit_depth) - 1
    yy, xx = np.mgrid[0:h + 4, 0:w + 4]
    org = np.clip(top / 2 + top / 3 * np.sin(xx / 5.0 + rng.random() * 6) * np.cos(yy / 4.0) + rng.integers(-4, 5, xx.shape) * (1 << (bit_depth - 8)), 0, top)
    rec = np.clip(org + rng.integers(-6, 7, xx.shape) * (1 << (bit_depth - 8)), 0, top)
    rec[rng.random(rec.shape) < 0.3] = rec[0, 0]  # flat runs: category 0 and ties occur
    return org.astype(dtype), rec.astype(dtype)


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10), (2, 9)])
def test_sao_statistics_match_reference(reflib, oracle, bps, bit_depth):
    if not hasattr(reflib, "ref_sao_stats"):
        pytest.skip("libhavoc_ref.so predates ref_shim_saostats.cpp (make -C oracle ref)")
    rng = np.random.default_rng(800 + bit_depth)
    for lib_fn in (reflib.ref_sao_stats, oracle.lib.orc_sao_stats):
        lib_fn.argtypes = [C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    seen = np.zeros(104, np.int64)
    for w, h in [(64, 64), (32, 32), (48, 64), (64, 16), (24, 8), (8, 8), (16, 40)] * 3:
        org, rec = stats_case(rng, bps, bit_depth, w, h)
        want, got = np.zeros(104, np.int64), np.zeros(104, np.int64)
        at = lambda a: a.ctypes.data + (2 * a.shape[1] + 2) * a.itemsize  # noqa: E731  (the block starts at (2, 2))
        start_ref = reflib.ref_sao_stats(at(org), org.shape[1], at(rec), rec.shape[1], w, h, bit_depth - 8, bps, want.ctypes.data)
        start = oracle.lib.orc_sao_stats(at(org), org.shape[1], at(rec), rec.shape[1], w, h, bit_depth - 8, bps, got.ctypes.data)
        assert np.array_equal(got, want), (w, h, np.nonzero(got != want)[0])
        assert start == start_ref
        seen += want != 0
    assert (seen[:40].reshape(4, 2, 5)[:, 1] > 0).all()  # every category of every class was populated
