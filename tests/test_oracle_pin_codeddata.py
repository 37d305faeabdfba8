"""Pins the oracle's restatement of CodedData::storeResidual (turing/CodedData.h:457-517) -- the record a transform
block's quantised levels become in the encoder's coded-data stream -- against the UNMODIFIED reference function
(oracle/ref_shim_codeddata.cpp inside oracle/_ref/libhavoc_ref.so): every block size, the three scans, sparse / dense /
single-coefficient / all-zero blocks, magnitudes up to the int16 limits."""
import ctypes as C

import numpy as np
import pytest

import orc


@pytest.fixture(scope="module")
def reflib():
    if not orc.REF_LIB.exists():
        pytest.skip("oracle/_ref/libhavoc_ref.so not built")
    lib = C.CDLL(str(orc.REF_LIB))
    if not hasattr(lib, "ref_coded_residual"):
        pytest.skip("libhavoc_ref.so predates ref_shim_codeddata.cpp (make -C oracle ref)")
    lib.ref_coded_residual.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    return lib


def make_levels(rng, log2n, kind):
    """quantised levels of one block, raster order"""
    n = 1 << log2n
    if kind == "zero":
        return np.zeros(n * n, np.int16)
    if kind == "single":
        lv = np.zeros(n * n, np.int16)
        lv[rng.integers(0, n * n)] = rng.choice([-3, -1, 1, 2, 32767, -32768 + 1])
        return lv
    density = {"sparse": 0.03, "medium": 0.2, "dense": 0.9}[kind]
    yy, xx = np.mgrid[0:n, 0:n]
    keep = rng.random((n, n)) < density * np.exp(-(xx + yy) / (n / 2.0)) * 3  # low frequencies are likelier, as after a transform
    mag = np.maximum(1, (rng.exponential(1.5, (n, n))).astype(np.int64))
    mag[rng.random((n, n)) < 0.02] = 32767
    return (np.where(keep, mag * rng.choice([-1, 1], (n, n)), 0)).astype(np.int16).reshape(-1)


def test_coded_residual_matches_reference(reflib, oracle):
    oracle.lib.orc_coded_residual.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    rng = np.random.default_rng(41)
    cases = nonzero = 0
    for log2n in (2, 3, 4, 5):
        cap = 5 + 19 * (1 << (2 * (log2n - 2))) + 8
        for scan in (0, 1, 2):
            for kind in ("zero", "single", "single", "sparse", "sparse", "medium", "medium", "dense"):
                for _ in range(6):
                    lv = make_levels(rng, log2n, kind)
                    want, got = np.zeros(cap, np.uint16), np.zeros(cap, np.uint16)
                    mutable = lv.copy()  # the reference takes a non-const pointer
                    n_ref = reflib.ref_coded_residual(mutable.ctypes.data, log2n, scan, want.ctypes.data, cap)
                    n_orc = oracle.lib.orc_coded_residual(lv.ctypes.data, log2n, scan, got.ctypes.data)
                    assert n_orc == n_ref, (log2n, scan, kind)
                    assert np.array_equal(got[:n_ref], want[:n_ref]), (log2n, scan, kind)
                    cases += 1
                    nonzero += n_ref > 0
    assert cases == 4 * 3 * 8 * 6 and nonzero > 0.8 * cases
