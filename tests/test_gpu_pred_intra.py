"""GPU parity: inter prediction (uni/bi), SubtractBi, fused interpolation+SATD, intra prediction and
the 35-mode SATD sweep, through the C-ABI vs the oracle -- bit-exact."""
import numpy as np
import pytest

from gpu_common import H, PAD, W, Scene, block
from turingcodec_b200 import hvb

pytestmark = pytest.mark.gpu

PU_SIZES = [(64, 64), (64, 32), (32, 64), (32, 32), (32, 24), (32, 8), (24, 32), (16, 64), (16, 16), (16, 12),
            (16, 4), (12, 16), (8, 32), (8, 8), (8, 4), (4, 8)]


@pytest.fixture(scope="module", params=[(1, 8), (2, 10), (2, 8)], ids=["u8", "u16-10bit", "u16-8bit"])
def scene(request):
    s = Scene(*request.param)
    yield s
    s.close()


def test_pred_uni_and_bi(scene, oracle):
    """each task writes its own disjoint 64x64 cell of a scratch picture"""
    rng = np.random.default_rng(31)
    cells = [(cx, cy) for cy in range(0, H - 63, 64) for cx in range(0, W - 63, 64)]
    for c_idx in (0, 1):
        taps = 8 if c_idx == 0 else 4
        frac_mask = 3 if c_idx == 0 else 7
        frac_shift = 2 if c_idx == 0 else 3
        scale = 1 if c_idx == 0 else 2
        t = np.zeros(len(cells), hvb.pred_task_t)
        for i, (cx, cy) in enumerate(cells):
            w, h = PU_SIZES[rng.integers(len(PU_SIZES))]
            w, h = w // scale, h // scale
            block(t[i:i + 1], "dst", scene.scratch[0], c_idx, cx // scale, cy // scale)
            bi = i % 3 == 2
            t[i]["ref_pic"] = (scene.pics[1], scene.pics[2] if bi else -1)
            t[i]["x"], t[i]["y"] = rng.integers(0, W // scale - w + 1), rng.integers(0, H // scale - h + 1)
            t[i]["w"], t[i]["h"] = w, h
            # vectors that reach into the padding, all fractional phases
            t[i]["mvx"] = rng.integers(-70 * (frac_mask + 1), 70 * (frac_mask + 1), 2) // scale
            t[i]["mvy"] = rng.integers(-70 * (frac_mask + 1), 70 * (frac_mask + 1), 2) // scale
            if i % 5 == 0:
                t[i]["mvx"][0] &= ~frac_mask  # integer / half-integer special cases
            if i % 7 == 0:
                t[i]["mvy"][0] &= ~frac_mask
        scene.ctx.pred(t)
        got = scene.download(scene.scratch[0], c_idx)
        for i in range(t.size):
            w, h = int(t[i]["w"]), int(t[i]["h"])
            want = np.zeros((64, 64), scene.dtype)
            refs = []
            for r in range(2):
                if t[i]["ref_pic"][r] < 0:
                    break
                mvx, mvy = int(t[i]["mvx"][r]), int(t[i]["mvy"][r])
                a, off, stride = scene.view(1 + r, c_idx, int(t[i]["x"]) + (mvx >> frac_shift), int(t[i]["y"]) + (mvy >> frac_shift))
                refs.append((a, off, stride, mvx & frac_mask, mvy & frac_mask))
            if len(refs) == 1:
                a, off, stride, xf, yf = refs[0]
                oracle.pred_uni(want, 0, 64, a, off, stride, w, h, xf, yf, scene.bd, taps)
            else:
                (a0, o0, stride, xf0, yf0), (a1, o1, _, xf1, yf1) = refs
                # both references share a stride; the oracle takes two pointers into possibly different arrays
                import ctypes as C
                oracle.lib.orc_pred_bi(C.c_void_p(want.ctypes.data), 64, C.c_void_p(a0.ctypes.data + o0 * a0.itemsize),
                                       C.c_void_p(a1.ctypes.data + o1 * a1.itemsize), stride, w, h, xf0, yf0, xf1, yf1,
                                       scene.bd, taps, a0.itemsize)
            dx, dy = int(t[i]["dst"]["x"]), int(t[i]["dst"]["y"])
            assert np.array_equal(got[dy:dy + h, dx:dx + w], want[:h, :w]), (c_idx, i, w, h, t[i])


def test_subtract_bi(scene, oracle):
    rng = np.random.default_rng(32)
    cells = [(cx, cy) for cy in range(0, H - 63, 64) for cx in range(0, W - 63, 64)]
    t = np.zeros(len(cells), hvb.subtract_bi_task_t)
    for i, (cx, cy) in enumerate(cells):
        w, h = PU_SIZES[rng.integers(len(PU_SIZES))]
        block(t[i:i + 1], "dst", scene.scratch[1], 0, cx, cy)
        block(t[i:i + 1], "pred", scene.pics[1], 0, rng.integers(0, W - w), rng.integers(0, H - h))
        block(t[i:i + 1], "src", scene.pics[0], 0, cx, cy)
        t[i]["w"], t[i]["h"] = w, h
    scene.ctx.subtract_bi(t)
    got = scene.download(scene.scratch[1], 0)
    for i in range(t.size):
        w, h = int(t[i]["w"]), int(t[i]["h"])
        want = np.zeros((64, 64), scene.dtype)
        p, op, sp = scene.view(1, 0, t[i]["pred"]["x"], t[i]["pred"]["y"])
        s, os_, ss = scene.view(0, 0, t[i]["src"]["x"], t[i]["src"]["y"])
        oracle.subtract_bi(want, 0, 64, p, op, sp, s, os_, ss, w, h, scene.bd)
        dx, dy = int(t[i]["dst"]["x"]), int(t[i]["dst"]["y"])
        assert np.array_equal(got[dy:dy + h, dx:dx + w], want[:h, :w])


def test_interp_satd(scene, oracle):
    """costDistortionMv's distortion: interpolate at quarter-pel, SATD against the source block"""
    rng = np.random.default_rng(33)
    n = 400
    t = np.zeros(n, hvb.interp_satd_task_t)
    for i in range(n):
        w, h = PU_SIZES[rng.integers(len(PU_SIZES))]
        block(t[i:i + 1], "src", scene.pics[0], 0, rng.integers(0, (W - w) // 4 + 1) * 4, rng.integers(0, (H - h) // 4 + 1) * 4)
        t[i]["ref_pic"] = scene.pics[1]
        t[i]["w"], t[i]["h"] = w, h
        t[i]["mvx"], t[i]["mvy"] = rng.integers(-300, 300), rng.integers(-300, 300)
    got = scene.ctx.interp_satd(t)
    for i in range(n):
        w, h = int(t[i]["w"]), int(t[i]["h"])
        mvx, mvy = int(t[i]["mvx"]), int(t[i]["mvy"])
        x, y = int(t[i]["src"]["x"]), int(t[i]["src"]["y"])
        a, off, stride = scene.view(1, 0, x + (mvx >> 2), y + (mvy >> 2))
        pred = np.zeros((64, 64), scene.dtype)
        oracle.pred_uni(pred, 0, 64, a, off, stride, w, h, mvx & 3, mvy & 3, scene.bd, 8)
        s, os_, ss = scene.view(0, 0, x, y)
        assert got[i] == oracle.measure_satd(s, os_, ss, pred, 0, 64, w, h), (i, w, h, mvx, mvy)


def neighbour_pool(scene, rng, count):
    """`count` neighbour arrays of 4*32+1 samples each (enough for every size), mixing noise, ramps and extremes"""
    span = 4 * 32 + 1
    pool = rng.integers(0, 1 << scene.bd, (count, span)).astype(scene.dtype)
    pool[1::4] = (np.linspace(0, (1 << scene.bd) - 1, span)[None, :]).astype(scene.dtype)  # smooth: triggers strong filter
    pool[2::4] = (1 << scene.bd) - 1 - (pool[2::4] & 3)
    return pool


def test_intra_pred_all_modes(scene, oracle):
    rng = np.random.default_rng(34)
    pool = neighbour_pool(scene, rng, 8)
    span = pool.shape[1]
    scene.ctx.pool_upload(pool.reshape(-1))
    for c_idx in (0, 1):
        for log2n in (2, 3, 4, 5):
            n = 1 << log2n
            scale = 1 if c_idx == 0 else 2
            cells = [(cx, cy) for cy in range(0, H // scale - n + 1, n) for cx in range(0, W // scale - n + 1, n)][:35 * 4]
            t = np.zeros(len(cells), hvb.intra_task_t)
            for i, (cx, cy) in enumerate(cells):
                block(t[i:i + 1], "dst", scene.scratch[0], c_idx, cx, cy)
                t[i]["nb"] = (i % 8) * span + 2 * 32
                t[i]["log2n"], t[i]["mode"] = log2n, i % 35
                t[i]["edge_flag"] = int(c_idx == 0 and log2n < 5)
            scene.ctx.intra_pred(t)
            got = scene.download(scene.scratch[0], c_idx)
            for i, (cx, cy) in enumerate(cells):
                want = np.zeros((n, n), scene.dtype)
                oracle.pred_intra(want, n, pool.reshape(-1), int(t[i]["nb"]), int(t[i]["mode"]), log2n, scene.bd,
                                  int(t[i]["edge_flag"]))
                assert np.array_equal(got[cy:cy + n, cx:cx + n], want), (c_idx, log2n, int(t[i]["mode"]))


from orc import filter_flag, filtered_neighbours  # noqa: E402  (pinned in test_oracle_pin_intra_filter.py)


@pytest.mark.parametrize("derive_filtered", [True, False])
def test_intra_satd35_sweep(scene, oracle, derive_filtered):
    rng = np.random.default_rng(35)
    pool = neighbour_pool(scene, rng, 8)
    span = pool.shape[1]
    flat = pool.reshape(-1)
    # explicit filtered arrays live behind the unfiltered ones
    filt = np.zeros_like(pool)
    tasks = []
    for log2n in (2, 3, 4, 5):
        n = 1 << log2n
        for k in range(12):
            tasks.append((log2n, k % 8, int(rng.integers(0, (W - n) // 4 + 1)) * 4, int(rng.integers(0, (H - n) // 4 + 1)) * 4))
    t = np.zeros(len(tasks), hvb.intra_sweep_task_t)
    upload = np.concatenate([flat, np.zeros(len(tasks) * span, scene.dtype)])
    want_nb = []
    for i, (log2n, which, x, y) in enumerate(tasks):
        n = 1 << log2n
        corner = which * span + 2 * 32
        u = flat[corner - 2 * n: corner + 2 * n + 1]
        f = filtered_neighbours(u, n, scene.bd, True).astype(scene.dtype)
        fcorner = flat.size + i * span + 2 * 32
        upload[fcorner - 2 * n: fcorner + 2 * n + 1] = f
        want_nb.append((u, f))
        block(t[i:i + 1], "src", scene.pics[0], 0, x, y)
        t[i]["nb_unfiltered"] = corner
        t[i]["nb_filtered"] = -1 if derive_filtered else fcorner
        t[i]["log2n"], t[i]["cIdx"], t[i]["strong_intra_smoothing"] = log2n, 0, 1
    scene.ctx.pool_upload(upload)
    got = scene.ctx.intra_satd35(t)
    for i, (log2n, which, x, y) in enumerate(tasks):
        n = 1 << log2n
        u, f = want_nb[i]
        s, os_, ss = scene.view(0, 0, x, y)
        for mode in range(35):
            nb = f if filter_flag(0, mode, n) else u
            pred = np.zeros((n, n), scene.dtype)
            oracle.pred_intra(pred, n, np.ascontiguousarray(nb), 2 * n, mode, log2n, scene.bd, int(log2n < 5))
            # PredictIntraLumaBlock (Reconstruct.cpp:683-701): 4x4 tile for log2 2, 8x8 tiles otherwise
            lt = 2 if log2n == 2 else 3
            tn = 1 << lt
            want = sum(oracle.hadamard_satd(s, os_ + dy * ss + dx, ss, pred, dy * n + dx, n, lt)
                       for dy in range(0, n, tn) for dx in range(0, n, tn))
            assert got[i][mode] == want, (log2n, mode, which)
