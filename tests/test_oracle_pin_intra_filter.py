"""Pins the reference-sample filtering the intra sweep relies on -- filterFlag (turing/Dsp.h:57-70) and
IntraReferenceSamples::filter (turing/IntraReferenceSamples.h:373-419) -- against the UNMODIFIED reference headers
(oracle/ref_shim_intra.cpp inside oracle/_ref/libhavoc_ref.so): the oracle's C version (used by the frame-pass checker)
and the numpy version (used by the GPU sweep test) must both equal the reference on every mode/size and on noise,
ramps (strong filter) and extremes at 8 and 10 bit."""
import ctypes as C

import numpy as np
import pytest

import orc


@pytest.fixture(scope="module")
def reflib():
    if not orc.REF_LIB.exists():
        pytest.skip("oracle/_ref/libhavoc_ref.so not built")
    lib = C.CDLL(str(orc.REF_LIB))
    if not hasattr(lib, "ref_intra_filter_flag"):
        pytest.skip("libhavoc_ref.so predates ref_shim_intra.cpp (make -C oracle ref)")
    lib.ref_intra_filter_neighbours.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    return lib


def test_filter_flag_rule_matches_reference_table(reflib, oracle):
    for c_idx in (0, 1, 2):
        for n in (4, 8, 16, 32):
            for mode in range(35):
                want = reflib.ref_intra_filter_flag(c_idx, mode, n)
                assert oracle.lib.orc_intra_filter_flag(c_idx, mode, n) == want, (c_idx, n, mode)
                assert int(orc.filter_flag(c_idx, mode, n)) == want, (c_idx, n, mode)


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10), (2, 9)])
def test_neighbour_filter_matches_reference(reflib, oracle, bps, bit_depth):
    rng = np.random.default_rng(61 + bit_depth)
    dtype = np.uint8 if bps == 1 else np.uint16
    top = (1 << bit_depth) - 1
    oracle.lib.orc_intra_filter_neighbours.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    strong_hits = 0
    for n in (4, 8, 16, 32):
        span = 4 * n + 1
        for trial in range(60):
            kind = trial % 4
            if kind == 0:
                u = rng.integers(0, top + 1, span)
            elif kind == 1:  # smooth ramp (+ a little noise on some): the bi-linear strong filter fires for n == 32
                u = np.linspace(rng.integers(0, top // 2), rng.integers(top // 2, top + 1), span) + (trial % 8 == 1) * rng.integers(0, 2, span)
            elif kind == 2:
                u = top - (rng.integers(0, 4, span))
            else:  # flat with one step: strong filter on one side only must not fire
                u = np.full(span, int(rng.integers(0, top + 1)))
                u[: int(rng.integers(1, span))] = int(rng.integers(0, top + 1))
            u = np.clip(u, 0, top).astype(dtype)
            for strong in (0, 1):
                want = np.zeros(span, dtype)
                reflib.ref_intra_filter_neighbours(want.ctypes.data, u.ctypes.data, n, bit_depth, strong, bps)
                u16, f16 = u.astype(np.uint16), np.zeros(span, np.uint16)
                oracle.lib.orc_intra_filter_neighbours(f16.ctypes.data, u16.ctypes.data, n, bit_depth, strong)
                assert np.array_equal(f16, want.astype(np.uint16)), ("oracle C", n, trial, strong)
                got = orc.filtered_neighbours(u, n, bit_depth, bool(strong)).astype(dtype)
                assert np.array_equal(got, want), ("numpy", n, trial, strong)
                if strong and n == 32:
                    plain = np.zeros(span, dtype)
                    reflib.ref_intra_filter_neighbours(plain.ctypes.data, u.ctypes.data, n, bit_depth, 0, bps)
                    strong_hits += not np.array_equal(plain, want)
    assert strong_hits > 5, strong_hits
