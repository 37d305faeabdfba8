"""Pin the C restatement (oracle/liboracle.so) to the unmodified reference havoc library
(oracle/_ref/libhavoc_ref.so, C_REF|C_OPT tables = the `--asm 0` identity path) on the reference
self-test's own input recipes (havoc_test_*).  CPU only."""
import numpy as np
import pytest

PU_SIZES = [(64, 64), (64, 48), (64, 32), (64, 16), (48, 64), (32, 64), (32, 32), (32, 24), (32, 16), (32, 8),
            (24, 32), (16, 64), (16, 32), (16, 16), (16, 12), (16, 8), (16, 4), (12, 16), (8, 32), (8, 16),
            (8, 8), (8, 4), (4, 8)]  # havoc/sad.h:28-51


def planes(rng, dtype, mask, shape=(128, 128)):
    return (rng.integers(0, 1 << 16, size=shape, dtype=np.uint32) & mask).astype(dtype)


@pytest.mark.parametrize("dtype,mask", [(np.uint8, 0xFF), (np.uint16, 0x3FF)])
def test_sad_and_sad4(oracle, ref_c, ref_asm, dtype, mask):
    # havoc/sad.cpp:1112-1136,1231-1261: 128x128 buffers, rand()&0x3ff, unaligned ref &ref[1+1*128], 23 PU sizes
    rng = np.random.default_rng(1)
    src, ref = planes(rng, dtype, mask), planes(rng, dtype, mask)
    for w, h in PU_SIZES:
        want = ref_c.sad(src, 0, 64, ref, 1 + 128, 64, w, h)
        assert oracle.sad(src, 0, 64, ref, 1 + 128, 64, w, h) == want
        assert ref_asm.sad(src, 0, 64, ref, 1 + 128, 64, w, h) == want
        offs = [1 + 128, 2 + 128, 1 + 2 * 128, 7 + 3 * 128]
        want4 = ref_c.sad4(src, 0, 64, ref, offs, 64, w, h)
        assert oracle.sad4(src, 0, 64, ref, offs, 64, w, h) == want4
        assert ref_asm.sad4(src, 0, 64, ref, offs, 64, w, h) == want4
    # every multiple-of-4 rectangle the populate covers (sad.cpp:494-504)
    for w in range(4, 65, 12):
        for h in range(4, 65, 20):
            assert oracle.sad(src, 0, 128, ref, 3, 128, w, h) == ref_c.sad(src, 0, 128, ref, 3, 128, w, h)


@pytest.mark.parametrize("dtype,mask", [(np.uint8, 0xFF), (np.uint16, 0x3FF), (np.uint16, 0xFFFF)])
def test_ssd(oracle, ref_c, dtype, mask):
    # havoc/ssd.cpp:287-318: log2 2..6, stride 2*n; 0xffff exercises the mod-2^32 wrap before the >> 4
    rng = np.random.default_rng(2)
    a, b = planes(rng, dtype, mask), planes(rng, dtype, mask)
    for log2n in range(2, 7):
        n = 1 << log2n
        assert oracle.ssd(a, 0, 2 * n, b, 0, 2 * n, n, n) == ref_c.ssd(a, 0, 2 * n, b, 0, 2 * n, log2n)
    a8, b8 = planes(rng, np.uint8, 0xFF, (4096,)), planes(rng, np.uint8, 0xFF, (4096,))
    assert oracle.ssd_linear(a8, b8) == ref_c.ssd_linear(a8, b8)


@pytest.mark.parametrize("dtype,bits", [(np.uint8, 8), (np.uint16, 10)])
def test_hadamard(oracle, ref_c, dtype, bits):
    # havoc/hadamard.cpp:865-898: near-extreme inputs srcA = max - (rand&7), srcB = rand&7; sizes 8,4,2
    rng = np.random.default_rng(3)
    a = ((1 << bits) - 1 - rng.integers(0, 8, (16, 16))).astype(dtype)
    b = rng.integers(0, 8, (16, 16)).astype(dtype)
    c = rng.integers(0, 1 << bits, (16, 16)).astype(dtype)
    for log2n in (1, 2, 3):
        for x, y in ((a, b), (b, a), (a, c), (c, c)):
            assert oracle.hadamard_satd(x, 0, 16, y, 0, 16, log2n) == ref_c.hadamard_satd(x, 0, 16, y, 0, 16, log2n)


@pytest.mark.parametrize("dtype,bit_depths", [(np.uint8, [8]), (np.uint16, [8, 9, 10])])
def test_pred_uni(oracle, ref_c, dtype, bit_depths):
    # havoc/pred_inter.cpp:1111-1189: bd x taps x frac x 24 partitions, stride 192
    rng = np.random.default_rng(4)
    S = 192
    for bd in bit_depths:
        ref = rng.integers(0, 1 << bd, (80, S)).astype(dtype)
        for taps in (8, 4):
            widths = [64, 48, 32, 24, 16, 12, 8] + ([4] if taps == 8 else [4, 6, 2])
            for w in widths:
                for h in (w, max(4, w // 2)) if taps == 8 else (w, max(2, w // 2)):
                    for xf, yf in ((0, 0), (1, 0), (0, 2), (3, 1), (2, 2)):
                        if taps == 4:
                            xf, yf = xf * 2 + (1 if xf else 0), yf * 2
                        a = np.zeros((64, S), dtype)
                        b = np.zeros((64, S), dtype)
                        ok = ref_c.pred_uni(a, 0, S, ref, 8 * S + 8, S, w, h, xf, yf, bd, taps)
                        if not ok:
                            continue
                        oracle.pred_uni(b, 0, S, ref, 8 * S + 8, S, w, h, xf, yf, bd, taps)
                        assert np.array_equal(a[:h, :w], b[:h, :w]), (bd, taps, w, h, xf, yf)


@pytest.mark.parametrize("dtype,bit_depths", [(np.uint8, [8]), (np.uint16, [8, 9, 10])])
def test_pred_bi_and_subtract(oracle, ref_c, dtype, bit_depths):
    # havoc/pred_inter.cpp:2020-2059 (fracs (0,0,0,0) and (1,2,3,0)) and :2204-2241 (SubtractBi 8..64)
    rng = np.random.default_rng(5)
    S = 192
    for bd in bit_depths:
        r0 = rng.integers(0, 1 << bd, (80, S)).astype(dtype)
        r1 = rng.integers(0, 1 << bd, (80, S)).astype(dtype)
        for taps in (8, 4):
            for w, h in ((64, 64), (32, 16), (16, 16), (8, 8), (8, 4)):
                for f in ((0, 0, 0, 0), (1, 2, 3, 0), (0, 3, 0, 1)):
                    a = np.zeros((64, S), dtype)
                    b = np.zeros((64, S), dtype)
                    ok = ref_c.pred_bi(a, 0, S, r0, 8 * S + 8, r1, 8 * S + 9, S, w, h, *f, bd, taps)
                    assert ok
                    oracle.pred_bi(b, 0, S, r0, 8 * S + 8, r1, 8 * S + 9, S, w, h, *f, bd, taps)
                    assert np.array_equal(a[:h, :w], b[:h, :w]), (bd, taps, w, h, f)
        for n in (8, 16, 32, 64):
            a = np.zeros((64, S), dtype)
            b = np.zeros((64, S), dtype)
            ref_c.subtract_bi(a, 0, S, r0, 0, S, r1, 0, S, n, n, bd)
            oracle.subtract_bi(b, 0, S, r0, 0, S, r1, 0, S, n, n, bd)
            assert np.array_equal(a, b)


@pytest.mark.parametrize("dtype,bit_depths", [(np.uint8, [8]), (np.uint16, [8, 9, 10])])
def test_pred_intra(oracle, ref_c, dtype, bit_depths):
    # havoc/pred_intra.cpp:22096-22116: bd x log2 2..5 x 35 modes (+ edge-filter variants via cIdx 0)
    rng = np.random.default_rng(6)
    for bd in bit_depths:
        for log2n in (2, 3, 4, 5):
            n = 1 << log2n
            for trial in range(3):
                nb = rng.integers(0, 1 << bd, 4 * n + 1 + 8).astype(dtype)
                if trial == 2:
                    nb[:] = (1 << bd) - 1 - (nb & 3)  # near-maximum, exercises the clips
                corner = 2 * n + 4
                for c_idx in (0, 1):
                    for mode in range(35):
                        a = np.zeros((n, n), dtype)
                        b = np.zeros((n, n), dtype)
                        assert ref_c.pred_intra(a, n, nb, corner, mode, log2n, bd, c_idx)
                        edge = int(c_idx == 0 and log2n < 5)
                        oracle.pred_intra(b, n, nb, corner, mode, log2n, bd, edge)
                        assert np.array_equal(a, b), (bd, log2n, c_idx, mode)


def test_transform_forward(oracle, ref_c):
    # havoc/transform.cpp:5355-5377: src in [-256,255], bd 8 & 10, DST4 + DCT 4..32; plus full-range
    # residuals that drive the first pass into the int16 wrap the reference reproduces
    rng = np.random.default_rng(7)
    for bd in (8, 10):
        for tr_type, log2n in ((1, 2), (0, 2), (0, 3), (0, 4), (0, 5)):
            n = 1 << log2n
            for lo, hi in ((-256, 256), (-(1 << bd) + 1, 1 << bd), (-32768, 32768)):
                src = rng.integers(lo, hi, (n, 2 * n)).astype(np.int16)
                if lo == -32768:
                    src[:] = np.where(rng.integers(0, 2, src.shape) > 0, 32767, -32768)
                a = np.zeros(n * n, np.int16)
                b = np.zeros(n * n, np.int16)
                ref_c.transform_fwd(a, src, 2 * n, tr_type, log2n, bd)
                oracle.transform_fwd(b, src, 2 * n, tr_type, log2n, bd)
                assert np.array_equal(a, b), (bd, tr_type, log2n, lo)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
def test_inverse_transform_add(oracle, ref_c, dtype):
    # havoc/transform.cpp:3049-3063: coeffs in [-128,127], bitDepth 8 for both sample types; we add the
    # DST case the reference loop never reaches, 10-bit for u16, and full-range coefficients (clip16)
    rng = np.random.default_rng(8)
    for bd in ([8] if dtype == np.uint8 else [8, 10]):
        for tr_type, log2n in ((1, 2), (0, 2), (0, 3), (0, 4), (0, 5)):
            n = 1 << log2n
            for lo, hi in ((-128, 128), (-32768, 32768)):
                coeffs = rng.integers(lo, hi, n * n).astype(np.int16)
                pred = rng.integers(0, 1 << bd, (n, 2 * n)).astype(dtype)
                a = np.zeros((n, 2 * n), dtype)
                b = np.zeros((n, 2 * n), dtype)
                ref_c.inverse_transform_add(a, 2 * n, pred, 2 * n, coeffs, tr_type, log2n, bd)
                oracle.inverse_transform_add(b, 2 * n, pred, 2 * n, coeffs, tr_type, log2n, bd)
                assert np.array_equal(a, b), (bd, tr_type, log2n, lo)
                ra = np.zeros(n * n, np.int16)
                rb = np.zeros(n * n, np.int16)
                ref_c.inverse_transform(ra, coeffs, tr_type, log2n, bd)
                oracle.inverse_transform(rb, coeffs, tr_type, log2n, bd)
                assert np.array_equal(ra, rb)


def test_quantize(oracle, ref_c):
    # havoc/quantize.cpp:509-532 (src=rand()-rand(), scale 51, shift 20, offset 14), :250-275 (inverse:
    # scale 51 & 52224, shift log2-1), :753-778 (reconstruct)
    rng = np.random.default_rng(9)
    for n in (16, 64, 256, 1024):
        src = (rng.integers(0, 1 << 15, n) - rng.integers(0, 1 << 15, n)).astype(np.int16)
        for scale, shift, offset in ((51, 20, 14), (26214, 21, 171), (16384, 16, 21845), (32767, 27, 32767)):
            a = np.zeros(n, np.int16)
            b = np.zeros(n, np.int16)
            ca = ref_c.quantize(a, src, scale, shift, offset)
            cb = oracle.quantize(b, src, scale, shift, offset)
            assert np.array_equal(a, b) and (ca != 0) == (cb != 0)
        for scale, shift in ((51, 1), (51, 4), (52224, 3), (52224, 9), (64 << 4, 5)):
            a = np.zeros(n, np.int16)
            b = np.zeros(n, np.int16)
            ref_c.quantize_inverse(a, src, scale, shift)
            oracle.quantize_inverse(b, src, scale, shift)
            assert np.array_equal(a, b), (scale, shift)
    for log2n in (2, 3, 4, 5):
        n = 1 << log2n
        pred = rng.integers(0, 256, (n, 2 * n)).astype(np.uint8)
        res = rng.integers(-300, 300, n * n).astype(np.int16)
        a = np.zeros((n, 2 * n), np.uint8)
        b = np.zeros((n, 2 * n), np.uint8)
        ref_c.quantize_reconstruct(a, 2 * n, pred, 2 * n, res, log2n)
        oracle.quantize_reconstruct(b, 2 * n, pred, 2 * n, res, n)
        assert np.array_equal(a, b)


def test_measure_satd_tiling(oracle):
    # turing/Measure.h:96-135: the tile size follows the alignment of (w|h)
    rng = np.random.default_rng(10)
    a = rng.integers(0, 256, (64, 64)).astype(np.uint8)
    b = rng.integers(0, 256, (64, 64)).astype(np.uint8)
    for w, h, log2n in ((64, 64, 3), (16, 8, 3), (12, 16, 2), (8, 4, 2), (4, 8, 2), (6, 8, 1), (2, 4, 1)):
        n = 1 << log2n
        want = sum(oracle.hadamard_satd(a, x + y * 64, 64, b, x + y * 64, 64, log2n)
                   for y in range(0, h, n) for x in range(0, w, n))
        assert oracle.measure_satd(a, 0, 64, b, 0, 64, w, h) == want
