"""Pins the oracle's intra-complexity measure against the UNMODIFIED reference function
EstimateIntraComplexity::computeSatd8x8 (turing/EstimateIntraComplexity.h:55-157, through oracle/ref_shim_preanalysis.cpp):
noise, flat, ramps and range extremes at 8 and 16 bit samples."""
import ctypes as C

import numpy as np
import pytest

import orc


@pytest.fixture(scope="module")
def reflib():
    if not orc.REF_LIB.exists():
        pytest.skip("oracle/_ref/libhavoc_ref.so not built")
    lib = C.CDLL(str(orc.REF_LIB))
    if not hasattr(lib, "ref_intra_complexity_8x8"):
        pytest.skip("libhavoc_ref.so predates ref_shim_preanalysis.cpp (make -C oracle ref)")
    lib.ref_intra_complexity_8x8.argtypes = [C.c_void_p, C.c_ssize_t, C.c_int]
    return lib


def pictures(rng, bps, bit_depth, w=96, h=64):
    dtype = np.uint8 if bps == 1 else np.uint16
    top = (1 << bit_depth) - 1
    yy, xx = np.mgrid[0:h, 0:w]
    yield rng.integers(0, top + 1, (h, w)).astype(dtype)                                   # noise
    yield np.full((h, w), top, dtype)                                                       # flat at the limit
    yield np.clip(xx * top // w + yy, 0, top).astype(dtype)                                 # ramps
    yield np.where((xx // 8 + yy // 8) % 2, top, 0).astype(dtype)                           # block checkerboard
    yield np.where((xx + yy) % 2, top, 0).astype(dtype)                                     # sample checkerboard (largest AC energy)
    yield np.clip(top / 2 + top / 3 * np.sin(xx / 3.0) * np.cos(yy / 5.0), 0, top).astype(dtype)


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10), (2, 16)])
def test_intra_complexity_matches_reference(reflib, oracle, bps, bit_depth):
    oracle.lib.orc_intra_complexity_8x8.argtypes = [C.c_void_p, C.c_ssize_t, C.c_int]
    oracle.lib.orc_intra_complexity.argtypes = [C.c_void_p, C.c_ssize_t, C.c_int, C.c_int, C.c_int, C.c_void_p]
    rng = np.random.default_rng(70 + bit_depth)
    seen = set()
    for pic in pictures(rng, bps, bit_depth):
        h, w = pic.shape
        out = np.zeros((h // 8) * (w // 8), np.int32)
        total = oracle.lib.orc_intra_complexity(pic.ctypes.data, w, w, h, bps, out.ctypes.data)
        k = 0
        for by in range(h // 8):
            for bx in range(w // 8):
                at = pic.ctypes.data + (8 * by * w + 8 * bx) * pic.itemsize
                want = reflib.ref_intra_complexity_8x8(at, w, bps)
                assert out[k] == want == oracle.lib.orc_intra_complexity_8x8(at, w, bps), (by, bx)
                seen.add(int(want))
                k += 1
        assert total == int(out.sum())
    assert 0 in seen and len(seen) > 50


# ---- adaptive quantisation and shot-change detection (turing/AdaptiveQuantisation.h, turing/SCDetection.h) ---------------

def aq_reflib(reflib):
    if not hasattr(reflib, "ref_aq_layer"):
        pytest.skip("libhavoc_ref.so predates the AQ / SCD shim (make -C oracle ref)")
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    reflib.ref_aq_layer.argtypes = [vp, C.c_ssize_t, i, i, i, i, i, i, vp, vp]
    reflib.ref_aq_offset.argtypes = [vp, C.c_ssize_t, i, i, i, i, i, i, i, i, i]
    reflib.ref_scd_likelihood.argtypes = [vp, vp, C.c_ssize_t, i, i, i]
    reflib.ref_scd_likelihood.restype = d
    reflib.ref_scd_histogram.argtypes = [vp, C.c_ssize_t, i, i, i, vp]
    return reflib


def oracle_aq(oracle):
    vp, i = C.c_void_p, C.c_int
    oracle.lib.orc_aq_layer.argtypes = [vp, C.c_ssize_t, i, i, i, i, vp]
    oracle.lib.orc_scd_histogram.argtypes = [vp, C.c_ssize_t, i, i, i, vp]
    oracle.lib.orc_scd_block_stats.argtypes = [vp, C.c_ssize_t, i, i, i, i, vp]
    oracle.lib.orc_scd_likelihood.argtypes = [vp, vp]
    oracle.lib.orc_scd_likelihood.restype = C.c_double
    return oracle.lib


def aq_layer(lib, pic, unit):
    h, w = pic.shape
    out = np.zeros(-(-h // unit) * -(-w // unit), np.int64)
    average = lib.orc_aq_layer(pic.ctypes.data, w, w, h, pic.itemsize, unit, out.ctypes.data)
    return out, average


# picture sizes: whole units, a clipped last row / column of units (4K's 2160 = 33 * 64 + 48), and sizes that are no multiple of 8
AQ_SIZES = [(128, 64), (104, 72), (136, 112), (90, 70)]


# (the reference squares samples in `int`: well defined up to 15-bit samples, which covers every HEVC bit depth)
@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10), (2, 15)])
def test_aq_activity_matches_reference(reflib, oracle, bps, bit_depth):
    ref, lib = aq_reflib(reflib), oracle_aq(oracle)
    rng = np.random.default_rng(170 + bit_depth)
    negative = False
    for w, h in AQ_SIZES:
        for pic in pictures(rng, bps, bit_depth, w, h):
            for depth in range(4):
                unit = 64 >> depth
                got, average = aq_layer(lib, pic, unit)
                want = np.zeros(got.size, np.float64)
                want_average = C.c_double()
                n = ref.ref_aq_layer(pic.ctypes.data, w, w, h, bps, 64, depth, 3, want.ctypes.data, C.byref(want_average))
                assert n == got.size
                assert np.array_equal(1.0 + got.astype(np.float64), want), (w, h, depth)
                assert float(average) == want_average.value
                negative |= bool((got < 0).any())
    assert negative  # the two quirky quadrants do produce "variances" below zero: the quirks are exercised


def test_aq_offset_from_device_style_outputs(reflib, oracle):
    """getAqOffset (turing/AdaptiveQuantisation.h:147-169) evaluated from the integer outputs equals the reference's, i.e. what
    hvb_aq_activity_batch returns is all the encoder's QP offsets need"""
    ref, lib = aq_reflib(reflib), oracle_aq(oracle)
    rng = np.random.default_rng(5)
    w, h = 136, 112
    pic = next(iter(pictures(rng, 1, 8, w, h)))
    import math
    for depth in range(4):
        unit = 64 >> depth
        units, average = aq_layer(lib, pic, unit)
        per_row = -(-w // unit)
        for row, col in [(0, 0), (40, 70), (111, 135), (64, 64)]:
            activity = 1.0 + float(units[(row // unit) * per_row + col // unit])
            scale = math.pow(2.0, 6 / 6.0)
            norm = (scale * activity + average) / (activity + scale * average)
            want = ref.ref_aq_offset(pic.ctypes.data, w, w, h, 1, 64, 3, 6, row, col, depth)
            assert int(math.floor(math.log(norm) / math.log(2.0) * 6.0 + 0.49999)) == want


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
def test_scd_matches_reference(reflib, oracle, bps, bit_depth):
    ref, lib = aq_reflib(reflib), oracle_aq(oracle)
    rng = np.random.default_rng(270 + bit_depth)
    w, h = 128, 72
    pics = list(pictures(rng, bps, bit_depth, w, h))
    for pic in pics:
        got, want = np.zeros(64, np.int32), np.zeros(64, np.int32)
        lib.orc_scd_histogram(pic.ctypes.data, w, w, h, bps, got.ctypes.data)
        ref.ref_scd_histogram(pic.ctypes.data, w, w, h, bps, want.ctypes.data)
        assert np.array_equal(got, want) and got.sum() == w * h
    values = set()
    for prev, cur in zip(pics, pics[1:] + pics[:1]):
        a, b = np.zeros(72, np.float64), np.zeros(32, np.float64)
        assert lib.orc_scd_block_stats(prev.ctypes.data, w, w, h, bps, 1, a.ctypes.data) == 36
        assert lib.orc_scd_block_stats(cur.ctypes.data, w, w, h, bps, 2, b.ctypes.data) == 16
        got = lib.orc_scd_likelihood(a.ctypes.data, b.ctypes.data)
        want = ref.ref_scd_likelihood(prev.ctypes.data, cur.ctypes.data, w, w, h, bps)
        assert (got == want) or (np.isnan(got) and np.isnan(want)), (got, want)   # bit-for-bit: same operations in the same order
        values.add(got)
    assert len(values) >= 4
