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
