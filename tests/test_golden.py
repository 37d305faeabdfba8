"""CPU-only: the oracle reproduces the committed golden vectors that were generated from the unmodified
reference (tests/golden/make_golden.py, reference C tables = `--asm 0` path).  This is what pins the oracle on
machines that do not have /root/reference (the GPU box)."""
from pathlib import Path

import numpy as np
import pytest

import orc

G = np.load(Path(__file__).resolve().parent / "golden" / "havoc_golden.npz")


@pytest.mark.parametrize("name,dtype,bits", [("u8", np.uint8, 8), ("u16", np.uint16, 10)])
def test_metrics_and_prediction(oracle, name, dtype, bits):
    a, b = G[f"planeA_{name}"], G[f"planeB_{name}"]
    sizes = [tuple(s) for s in G["sizes"]]
    assert [oracle.sad(a, 0, 96, b, 97, 96, w, h) for w, h in sizes] == list(G[f"sad_{name}"])
    assert [oracle.sad4(a, 0, 96, b, [97, 98, 2 * 96 + 1, 3 * 96 + 7], 96, w, h) for w, h in sizes] == G[f"sad4_{name}"].tolist()
    assert [oracle.ssd(a, 0, 96, b, 5, 96, 1 << lg, 1 << lg) for lg in range(2, 7)] == list(G[f"ssd_{name}"])
    assert [oracle.hadamard_satd(a, 0, 96, b, 3, 96, lg) for lg in (1, 2, 3)] == list(G[f"satd_{name}"])
    k = 0
    for taps, fr in ((8, [(x, y) for x in range(4) for y in range(4)]), (4, [(0, 0), (1, 0), (0, 5), (3, 7), (4, 4)])):
        for xf, yf in fr:
            d = np.zeros((16, 96), dtype)
            oracle.pred_uni(d, 0, 96, a, 8 * 96 + 8, 96, 16, 16, xf, yf, bits, taps)
            assert np.array_equal(d[:, :16], G[f"pred_uni_{name}"][k]), (taps, xf, yf)
            k += 1
    d = np.zeros((16, 96), dtype)
    oracle.pred_bi(d, 0, 96, a, 8 * 96 + 8, b, 9 * 96 + 11, 96, 16, 16, 1, 2, 3, 0, bits, 8)
    assert np.array_equal(d[:, :16], G[f"pred_bi_{name}"])
    d = np.zeros((16, 96), dtype)
    oracle.subtract_bi(d, 0, 96, a, 0, 96, b, 0, 96, 16, 16, bits)
    assert np.array_equal(d[:, :16], G[f"subtract_bi_{name}"])


@pytest.mark.parametrize("name,dtype,bits", [("u8", np.uint8, 8), ("u16", np.uint16, 10)])
def test_intra_and_inverse_transform(oracle, name, dtype, bits):
    nb = G[f"intra_nb_{name}"]
    for lg in (2, 3, 4, 5):
        n = 1 << lg
        want = G[f"intra_{name}_{lg}"]
        for c_idx in (0, 1):
            for mode in range(35):
                d = np.zeros((n, n), dtype)
                oracle.pred_intra(d, n, nb, 64, mode, lg, bits, int(c_idx == 0 and lg < 5))
                assert np.array_equal(d, want[c_idx, mode]), (lg, c_idx, mode)
    for lg, tr in ((2, 1), (2, 0), (3, 0), (4, 0), (5, 0)):
        n = 1 << lg
        d = np.zeros((n, n), dtype)
        oracle.inverse_transform_add(d, n, G[f"ita_{name}_{lg}_{tr}_pred"], n, G[f"ita_{name}_{lg}_{tr}_coeffs"], tr, lg, bits)
        assert np.array_equal(d, G[f"ita_{name}_{lg}_{tr}_out"])


def test_forward_transform_quantisation_rdoq(oracle):
    for bits in (8, 10):
        for lg, tr in ((2, 1), (2, 0), (3, 0), (4, 0), (5, 0)):
            n = 1 << lg
            for tag in ("res", "wrap"):
                co = np.zeros(n * n, np.int16)
                oracle.transform_fwd(co, G[f"fwd_{bits}_{lg}_{tr}_{tag}_in"], n, tr, lg, bits)
                assert np.array_equal(co, G[f"fwd_{bits}_{lg}_{tr}_{tag}_out"]), (bits, lg, tr, tag)
    src = G["quant_src"]
    for i, (scale, shift, off) in enumerate(((51, 20, 14), (20560, 22, 10880), (26214, 21, 171 << 7))):
        d = np.zeros(1024, np.int16)
        oracle.quantize(d, src, scale, shift, off)
        assert np.array_equal(d, G[f"quant_{i}"])
    for i, (scale, shift) in enumerate(((51, 4), (52224, 9), (816, 4))):
        d = np.zeros(1024, np.int16)
        oracle.quantize_inverse(d, src, scale, shift)
        assert np.array_equal(d, G[f"dequant_{i}"])
    for i, case in enumerate(G["rdoq_cases"]):
        qs, qsh, iqs, lg, c_idx, scan_idx, intra, sdh, cbf = (int(v) for v in case)
        out = np.zeros_like(G[f"rdoq_{i}_in"])
        c = orc.oracle_rdoq(oracle, out, G[f"rdoq_{i}_in"], G[f"rdoq_{i}_ctx"], qs, qsh, iqs, lg, c_idx, scan_idx, intra, sdh, 8)
        assert np.array_equal(out, G[f"rdoq_{i}_out"]), i
        assert int(c != 0) == cbf
