"""Generate tests/golden/havoc_golden.npz from the UNMODIFIED reference (oracle/_ref/libhavoc_ref.so,
C_REF|C_OPT tables = the `--asm 0` identity path).  Run once in the container that has /root/reference;
the .npz is committed so that the oracle can be checked without the reference (e.g. on the GPU box).

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT / "tests"))
import orc  # noqa: E402

ref = orc.Ref(use_asm=False)
rng = np.random.default_rng(20260925)
g = {}

# SAD / SAD4 / SSD / SATD on a 96x96 pair of planes (u8 and 10-bit u16)
for name, dtype, bits in (("u8", np.uint8, 8), ("u16", np.uint16, 10)):
    a = rng.integers(0, 1 << bits, (96, 96)).astype(dtype)
    b = rng.integers(0, 1 << bits, (96, 96)).astype(dtype)
    g[f"planeA_{name}"], g[f"planeB_{name}"] = a, b
    sizes = [(64, 64), (64, 32), (32, 64), (32, 32), (32, 24), (16, 16), (16, 12), (12, 16), (8, 8), (8, 4), (4, 8)]
    g[f"sad_{name}"] = np.array([ref.sad(a, 0, 96, b, 96 + 1, 96, w, h) for w, h in sizes], np.int64)
    g[f"sad4_{name}"] = np.array([ref.sad4(a, 0, 96, b, [97, 98, 2 * 96 + 1, 3 * 96 + 7], 96, w, h) for w, h in sizes], np.int64)
    g[f"ssd_{name}"] = np.array([ref.ssd(a, 0, 96, b, 5, 96, lg) for lg in range(2, 7)], np.int64)
    g[f"satd_{name}"] = np.array([ref.hadamard_satd(a, 0, 96, b, 3, 96, lg) for lg in (1, 2, 3)], np.int64)
    g["sizes"] = np.array(sizes)

    # inter prediction: every phase for luma, a few for chroma; bi; subtract
    out = []
    for taps, fr in ((8, [(x, y) for x in range(4) for y in range(4)]), (4, [(0, 0), (1, 0), (0, 5), (3, 7), (4, 4)])):
        for xf, yf in fr:
            d = np.zeros((16, 96), dtype)
            assert ref.pred_uni(d, 0, 96, a, 8 * 96 + 8, 96, 16, 16, xf, yf, bits, taps)
            out.append(d[:, :16].copy())
    g[f"pred_uni_{name}"] = np.stack(out)
    d = np.zeros((16, 96), dtype)
    assert ref.pred_bi(d, 0, 96, a, 8 * 96 + 8, b, 9 * 96 + 11, 96, 16, 16, 1, 2, 3, 0, bits, 8)
    g[f"pred_bi_{name}"] = d[:, :16].copy()
    d = np.zeros((16, 96), dtype)
    ref.subtract_bi(d, 0, 96, a, 0, 96, b, 0, 96, 16, 16, bits)
    g[f"subtract_bi_{name}"] = d[:, :16].copy()

    # intra: all modes, sizes 4..32, luma-edge and chroma variants
    nb = rng.integers(0, 1 << bits, 4 * 32 + 1).astype(dtype)
    g[f"intra_nb_{name}"] = nb
    for lg in (2, 3, 4, 5):
        n = 1 << lg
        arr = np.zeros((2, 35, n, n), dtype)
        for c_idx in (0, 1):
            for mode in range(35):
                assert ref.pred_intra(arr[c_idx, mode], n, nb, 64, mode, lg, bits, c_idx)
        g[f"intra_{name}_{lg}"] = arr

    # inverse transform + add
    for lg, tr in ((2, 1), (2, 0), (3, 0), (4, 0), (5, 0)):
        n = 1 << lg
        co = rng.integers(-2000, 2000, n * n).astype(np.int16)
        pr = rng.integers(0, 1 << bits, (n, n)).astype(dtype)
        d = np.zeros((n, n), dtype)
        ref.inverse_transform_add(d, n, pr, n, co, tr, lg, bits)
        g[f"ita_{name}_{lg}_{tr}_coeffs"], g[f"ita_{name}_{lg}_{tr}_pred"], g[f"ita_{name}_{lg}_{tr}_out"] = co, pr, d

# forward transforms (8 and 10 bit), incl. a block that wraps in pass 1
for bits in (8, 10):
    for lg, tr in ((2, 1), (2, 0), (3, 0), (4, 0), (5, 0)):
        n = 1 << lg
        src = rng.integers(-(1 << bits) + 1, 1 << bits, (n, n)).astype(np.int16)
        wrap = np.where(rng.integers(0, 2, (n, n)) > 0, 32767, -32768).astype(np.int16)
        for tag, s in (("res", src), ("wrap", wrap)):
            co = np.zeros(n * n, np.int16)
            ref.transform_fwd(co, s, n, tr, lg, bits)
            g[f"fwd_{bits}_{lg}_{tr}_{tag}_in"], g[f"fwd_{bits}_{lg}_{tr}_{tag}_out"] = s, co

# quantise / inverse
src = (rng.integers(0, 1 << 15, 1024) - rng.integers(0, 1 << 15, 1024)).astype(np.int16)
g["quant_src"] = src
for i, (scale, shift, off) in enumerate(((51, 20, 14), (20560, 22, 10880), (26214, 21, 171 << 7))):
    d = np.zeros(1024, np.int16)
    ref.quantize(d, src, scale, shift, off)
    g[f"quant_{i}"] = d
for i, (scale, shift) in enumerate(((51, 4), (52224, 9), (816, 4))):
    d = np.zeros(1024, np.int16)
    ref.quantize_inverse(d, src, scale, shift)
    g[f"dequant_{i}"] = d

# RDOQ
cases = []
for i in range(24):
    lg = 2 + i % 4
    n = 1 << lg
    qp = [18, 26, 32, 38][i % 4]
    coeffs = (rng.normal(0, [8, 30, 100][i % 3], n * n)).astype(np.int16)
    ctx = orc.random_rdoq_ctx(rng, 0.57 * 2 ** ((qp - 12) / 3.0))
    qs, qsh, iqs, _ = orc.quant_params(qp, lg, 8)
    c_idx, scan_idx = (i // 4) % 3 if lg < 5 else 0, (i // 2) % 3 if lg <= 3 else 0
    out = np.zeros_like(coeffs)
    cbf = orc.ref_rdoq(ref, out, coeffs, ctx, qs, qsh, iqs, lg, c_idx, scan_idx, i % 2, (i // 2) % 2, 8)
    g[f"rdoq_{i}_in"], g[f"rdoq_{i}_ctx"], g[f"rdoq_{i}_out"] = coeffs, ctx, out
    cases.append((qs, qsh, iqs, lg, c_idx, scan_idx, i % 2, (i // 2) % 2, int(cbf != 0)))
g["rdoq_cases"] = np.array(cases)

np.savez_compressed(Path(__file__).resolve().parent / "havoc_golden.npz", **g)
print("wrote", len(g), "arrays")
