"""Pin oracle_rdoq.c against the reference's own Rdoq.cpp (compiled unmodified into oracle/_ref)."""
import numpy as np
import pytest

import orc


def coefficient_block(oracle, rng, log2n, bit_depth, amplitude):
    n = 1 << log2n
    res = np.clip(rng.normal(0, amplitude, (n, n)), -(1 << bit_depth) + 1, (1 << bit_depth) - 1).astype(np.int16)
    if rng.integers(0, 3) == 0:
        res[:, : n // 2] = 0  # structured residual: energy in a few coefficients
    coeffs = np.zeros(n * n, np.int16)
    oracle.transform_fwd(coeffs, res, n, 0, log2n, bit_depth)
    return coeffs


@pytest.mark.parametrize("bit_depth", [8, 10])
def test_rdoq_matches_reference(oracle, ref_c, bit_depth):
    rng = np.random.default_rng(21 + bit_depth)
    checked = nonzero = hidden = 0
    for trial in range(400):
        log2n = int(rng.integers(2, 6))
        c_idx = int(rng.integers(0, 3)) if log2n < 5 else 0
        scan_idx = int(rng.integers(0, 3)) if log2n <= 3 else 0
        qp = int(rng.integers(10, 40))
        lam = 0.57 * 2 ** ((qp - 12) / 3.0) * float(rng.uniform(0.5, 2.0))
        is_intra = int(rng.integers(0, 2))
        sdh = int(rng.integers(0, 2))
        amp = float(rng.choice([2, 6, 20, 60])) * (1 << (bit_depth - 8))
        src = coefficient_block(oracle, rng, log2n, bit_depth, amp)
        ctx = orc.random_rdoq_ctx(rng, lam)
        qscale, qshift, iqscale, _ = orc.quant_params(qp, log2n, bit_depth)
        a = np.zeros_like(src)
        b = np.zeros_like(src)
        args = (src, ctx, qscale, qshift, iqscale, log2n, c_idx, scan_idx, is_intra, sdh, bit_depth)
        ca = orc.ref_rdoq(ref_c, a, *args)
        cb = orc.oracle_rdoq(oracle, b, *args)
        assert np.array_equal(a, b), (trial, log2n, c_idx, scan_idx, qp, is_intra, sdh)
        assert (ca != 0) == (cb != 0)
        checked += 1
        nonzero += int(a.any())
        plain = np.zeros_like(src)
        orc.ref_rdoq(ref_c, plain, *args[:-2], 0, bit_depth)
        hidden += int(sdh and not np.array_equal(plain, a))
    assert nonzero > checked // 3      # the test must exercise coded blocks...
    assert hidden > 5                  # ...and blocks where sign-data hiding changed a level


@pytest.mark.parametrize("bit_depth", [8, 10])
def test_rdoq_decomposes_by_coefficient_group(oracle, bit_depth):
    """The decomposition a warp-parallel RDOQ walk rests on (DESIGN.md roadmap 0(a)): a 4x4 group's level decisions depend on the rest of
    the block only through one carried bit and its right / below neighbours' coded flags, so all groups can be walked for the eight
    combinations independently and selected afterwards.  orc_rdoq_grouped does that and must equal the pinned serial oracle bit for bit."""
    rng = np.random.default_rng(77 + bit_depth)
    coded = dense = carried = 0
    for trial in range(500):
        log2n = int(rng.integers(2, 6))
        c_idx = int(rng.integers(0, 3)) if log2n < 5 else 0
        scan_idx = int(rng.integers(0, 3)) if log2n <= 3 else 0
        qp = int(rng.integers(8, 40))
        lam = 0.57 * 2 ** ((qp - 12) / 3.0) * float(rng.uniform(0.5, 2.0))
        is_intra, sdh = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        amp = float(rng.choice([2, 6, 20, 60, 200])) * (1 << (bit_depth - 8))
        src = coefficient_block(oracle, rng, log2n, bit_depth, amp)
        ctx = orc.random_rdoq_ctx(rng, lam)
        qscale, qshift, iqscale, _ = orc.quant_params(qp, log2n, bit_depth)
        a, b = np.zeros_like(src), np.zeros_like(src)
        args = (src, ctx, qscale, qshift, iqscale, log2n, c_idx, scan_idx, is_intra, sdh, bit_depth)
        ca = orc.oracle_rdoq(oracle, a, *args)
        cb = orc.oracle_rdoq_grouped(oracle, b, *args)
        assert np.array_equal(a, b), (trial, log2n, c_idx, scan_idx, qp, is_intra, sdh)
        assert ca == cb
        coded += int(a.any())
        dense += int(np.count_nonzero(a) > a.size // 4)
        carried += int((np.abs(a) > 1).sum() > 8)  # many levels above 1: the carried bit is set again and again
    assert coded > 250 and dense > 40 and carried > 60


def test_scan_orders(oracle):
    # turing/ScanOrder.h:32-101: up-right diagonal, horizontal, vertical
    L = oracle.lib
    diag4 = [(L.orc_scan_order(2, 0, i, 0), L.orc_scan_order(2, 0, i, 1)) for i in range(16)]
    assert diag4[:6] == [(0, 0), (0, 1), (1, 0), (0, 2), (1, 1), (2, 0)] and diag4[-1] == (3, 3)
    assert [(L.orc_scan_order(1, 1, i, 0), L.orc_scan_order(1, 1, i, 1)) for i in range(4)] == [(0, 0), (1, 0), (0, 1), (1, 1)]
    assert [(L.orc_scan_order(1, 2, i, 0), L.orc_scan_order(1, 2, i, 1)) for i in range(4)] == [(0, 0), (0, 1), (1, 0), (1, 1)]
    assert L.orc_scan_order(0, 0, 0, 0) == 0
