"""GPU parity: the whole uni-directional PU motion search (integer pattern search + sub-pel refinement
with their control flow) on the device vs the oracle's restatement of turing/Search.hpp -- every
output field bit-exact, including the number of SAD evaluations (same path through the search)."""
import ctypes as C

import numpy as np
import pytest

import orc
from gpu_common import H, PAD, W, Scene
from turingcodec_b200 import hvb

pytestmark = pytest.mark.gpu

PU_SIZES = [(64, 64), (64, 32), (32, 64), (32, 32), (32, 16), (16, 32), (16, 16), (16, 8), (8, 16), (8, 8), (8, 4),
            (4, 8), (32, 24), (24, 32), (16, 12), (12, 16), (64, 48), (48, 64), (64, 16), (16, 64), (32, 8), (8, 32),
            (16, 4)]
CTB = 64


@pytest.fixture(scope="module", params=[(1, 8), (2, 10)], ids=["u8", "u16-10bit"])
def scene(request):
    s = Scene(*request.param)
    yield s
    s.close()


def make_task(rng, scene, i):
    w, h = PU_SIZES[i % len(PU_SIZES)]
    x0 = int(rng.integers(0, (W - w) // 4 + 1)) * 4
    y0 = int(rng.integers(0, (H - h) // 4 + 1)) * 4
    t = np.zeros(1, hvb.me_task_t)[0]
    t["src_pic"], t["ref_pic"] = scene.pics[0], scene.pics[1 + i % 2]
    t["x0"], t["y0"], t["w"], t["h"] = x0, y0, w, h
    spread = [2, 12, 40, 160][i % 4]
    truth = np.array([12, 8]) * (1 + i % 2)  # synth.frame moves (+3,+2) samples per frame
    for k in range(2):
        if i % 3 == 1:  # predictors near the true motion: the MET early exits fire
            t["mvp"][k]["x"], t["mvp"][k]["y"] = truth + rng.integers(-2, 3, 2)
        else:
            t["mvp"][k]["x"], t["mvp"][k]["y"] = rng.integers(-spread, spread + 1, 2)
    t["rateMvpFlag"] = rng.integers(20000, 60000, 2)
    t["lambda"] = int(rng.choice([0.05, 0.2, 0.6, 1.5]) * 65536 + 0.5)
    # LimitFullPelMv (Search.hpp:1366-1407)
    t["limitMin"]["x"], t["limitMin"]["y"] = -CTB - x0, -CTB - y0
    t["limitMax"]["x"], t["limitMax"]["y"] = W + CTB - x0 - w, H + CTB - y0 - h
    if i % 5 == 0:  # the wavefront restriction (concurrent frames > 1)
        t["limitMax"]["x"] = min(int(t["limitMax"]["x"]), (x0 // CTB) * CTB + 3 * CTB - x0 - w - 15)
        t["limitMax"]["y"] = min(int(t["limitMax"]["y"]), (y0 // CTB) * CTB + 2 * CTB - y0 - h - 15)
    t["prev2Nx2N"]["x"], t["prev2Nx2N"]["y"] = rng.integers(-12, 13, 2) * 4
    t["smallSearchWindow"] = i % 3 == 0
    t["met"] = i % 2
    t["log2CbSize"] = 3 + (max(w, h) > 8) + (max(w, h) > 16) + (max(w, h) > 32)
    t["usePrev2Nx2N"] = (i // 2) % 2
    t["halfPel"] = 1
    t["quarterPel"] = i % 4 != 3
    return t


def oracle_search(oracle, scene, t):
    ot, r = orc.MeTask(), orc.MeResult()
    ot.x0, ot.y0, ot.w, ot.h = int(t["x0"]), int(t["y0"]), int(t["w"]), int(t["h"])
    for k in range(2):
        ot.mvp[2 * k], ot.mvp[2 * k + 1] = int(t["mvp"][k]["x"]), int(t["mvp"][k]["y"])
        ot.rateMvpFlag[k] = int(t["rateMvpFlag"][k])
    ot.lambda_ = int(t["lambda"])
    ot.limitMin[0], ot.limitMin[1] = int(t["limitMin"]["x"]), int(t["limitMin"]["y"])
    ot.limitMax[0], ot.limitMax[1] = int(t["limitMax"]["x"]), int(t["limitMax"]["y"])
    ot.smallSearchWindow, ot.met, ot.log2CbSize = int(t["smallSearchWindow"]), int(t["met"]), int(t["log2CbSize"])
    ot.usePrev2Nx2N = int(t["usePrev2Nx2N"])
    ot.prev2Nx2N[0], ot.prev2Nx2N[1] = int(t["prev2Nx2N"]["x"]), int(t["prev2Nx2N"]["y"])
    ot.halfPel, ot.quarterPel, ot.bitDepth = int(t["halfPel"]), int(t["quarterPel"]), scene.bd
    src = scene.host[0][0]
    ref = scene.host[int(t["ref_pic"]) - scene.pics[0]][0]
    base = (PAD * src.shape[1] + PAD) * src.itemsize
    oracle.lib.orc_me_search(C.c_void_p(src.ctypes.data + base), src.shape[1], C.c_void_p(ref.ctypes.data + base),
                             ref.shape[1], C.byref(ot), C.byref(r), scene.bps)
    return r


def test_me_search_matches_oracle(scene, oracle):
    rng = np.random.default_rng(51)
    n = 230
    tasks = np.zeros(n, hvb.me_task_t)
    for i in range(n):
        tasks[i] = make_task(rng, scene, i)
    got = scene.ctx.me_search(tasks)
    early = refined = 0
    for i in range(n):
        r = oracle_search(oracle, scene, tasks[i])
        g = got[i]
        key = (i, tuple(tasks[i][["x0", "y0", "w", "h"]]))
        assert (int(g["mv"]["x"]), int(g["mv"]["y"])) == tuple(r.mv), key
        assert (int(g["mvd"]["x"]), int(g["mvd"]["y"])) == tuple(r.mvd), key
        assert (int(g["mvInteger"]["x"]), int(g["mvInteger"]["y"])) == tuple(r.mvInteger), key
        assert int(g["mvpFlag"]) == r.mvpFlag and int(g["cost"]) == r.cost, key
        assert int(g["subpelCost"]) == r.subpelCost, key
        assert int(g["nSad"]) == r.nSad, key
        assert (int(g["flags"]) & 1) == r.earlyExit, key
        if not r.earlyExit:
            assert list(g["costMvdZero"]) == list(r.costMvdZero), key
        early += r.earlyExit
        refined += tuple(r.mv) != tuple(r.mvInteger)
    assert early > 5 and refined > 10, (early, refined)  # both the MET exits and the sub-pel moves are exercised


@pytest.mark.parametrize("jit", [False, True], ids=["ref-c", "ref-jit"])
def test_me_search_matches_reference_templates(scene, jit):
    """The device search against the UNMODIFIED reference templates themselves (oracle/_ref/libsearch_ref.so,
    see test_oracle_search_pin.py), without the oracle in between."""
    import test_oracle_search_pin as pin
    if not pin.LIB.exists():
        pytest.skip("oracle/_ref/libsearch_ref.so not built")
    lib = C.CDLL(str(pin.LIB))
    lib.ref_search_batch.argtypes = [C.POINTER(pin.RefPictures), C.POINTER(pin.RefTask), C.POINTER(pin.RefResult), C.c_int]
    lib.havoc_instruction_set_support.restype = C.c_int
    rng = np.random.default_rng(99)
    n = 460
    for distance in (1, 2):
        src, ref = scene.host[0][0], scene.host[distance][0]
        base = (PAD * src.shape[1] + PAD) * src.itemsize
        pics = pin.RefPictures(src.ctypes.data + base, ref.ctypes.data + base, src.shape[1], ref.shape[1], W, H, PAD,
                               scene.bps, lib.havoc_instruction_set_support() if jit else 3, CTB, int(jit))
        rtasks = (pin.RefTask * n)(*[pin.make_ref_task(rng, i, scene.bd, distance) for i in range(n)])
        want = (pin.RefResult * n)()
        assert lib.ref_search_batch(C.byref(pics), rtasks, want, n) == 0
        tasks = np.zeros(n, hvb.me_task_t)
        for i in range(n):
            o, t = pin.oracle_task(rtasks[i], want[i]), tasks[i]
            t["src_pic"], t["ref_pic"] = scene.pics[0], scene.pics[distance]
            t["x0"], t["y0"], t["w"], t["h"] = o.x0, o.y0, o.w, o.h
            for k in range(2):
                t["mvp"][k]["x"], t["mvp"][k]["y"] = o.mvp[2 * k], o.mvp[2 * k + 1]
            t["rateMvpFlag"] = (o.rateMvpFlag[0], o.rateMvpFlag[1])
            t["lambda"] = o.lambda_
            t["limitMin"]["x"], t["limitMin"]["y"] = o.limitMin[0], o.limitMin[1]
            t["limitMax"]["x"], t["limitMax"]["y"] = o.limitMax[0], o.limitMax[1]
            t["prev2Nx2N"]["x"], t["prev2Nx2N"]["y"] = o.prev2Nx2N[0], o.prev2Nx2N[1]
            t["smallSearchWindow"], t["met"], t["log2CbSize"] = o.smallSearchWindow, o.met, o.log2CbSize
            t["usePrev2Nx2N"], t["halfPel"], t["quarterPel"] = o.usePrev2Nx2N, o.halfPel, o.quarterPel
        got = scene.ctx.me_search(tasks)
        for i in range(n):
            g, r = got[i], want[i]
            key = (distance, i, tuple(tasks[i][["x0", "y0", "w", "h"]]))
            assert (int(g["mv"]["x"]), int(g["mv"]["y"])) == tuple(r.mv), key
            assert (int(g["mvd"]["x"]), int(g["mvd"]["y"])) == tuple(r.mvd), key
            assert (int(g["mvInteger"]["x"]), int(g["mvInteger"]["y"])) == tuple(r.mvInteger), key
            assert int(g["mvpFlag"]) == r.mvpFlag and int(g["cost"]) == r.cost, key
            early = int(g["flags"]) & 1
            if not early:
                assert list(g["costMvdZero"]) == list(r.costMvdZero), key
            changed = tuple(r.prev2Nx2NAfter) != tuple(rtasks[i].prev2Nx2N)
            assert not (changed and (early or rtasks[i].partMode != 0)), key


# ---- searchMotionBi on the device ------------------------------------------------------------------------------

def fill_bi(t, o, src_pic, ref_pic, other_pic):
    """orc.MeBiTask -> hvb_me_bi_task"""
    t["src_pic"], t["ref_pic"], t["other_pic"] = src_pic, ref_pic, other_pic
    t["x0"], t["y0"], t["w"], t["h"] = o.x0, o.y0, o.w, o.h
    for k in range(2):
        t["mvp"][k]["x"], t["mvp"][k]["y"] = o.mvp[2 * k], o.mvp[2 * k + 1]
    t["rateMvpFlag"] = (o.rateMvpFlag[0], o.rateMvpFlag[1])
    t["lambda"] = o.lambda_
    t["limitMin"]["x"], t["limitMin"]["y"] = o.limitMin[0], o.limitMin[1]
    t["limitMax"]["x"], t["limitMax"]["y"] = o.limitMax[0], o.limitMax[1]
    t["mvStart"]["x"], t["mvStart"]["y"] = o.mvStart[0], o.mvStart[1]
    t["mvOther"]["x"], t["mvOther"]["y"] = o.mvOther[0], o.mvOther[1]
    t["smallWindow"], t["halfPel"], t["quarterPel"] = o.smallWindow, o.halfPel, o.quarterPel


def make_bi_tasks(rng, scene, n):
    """-> hvb_me_bi_task array and the matching orc.MeBiTask list (also used by tests/test_host_emulated_me.py)"""
    tasks = np.zeros(n, hvb.me_bi_task_t)
    otasks = []
    for i in range(n):
        w, h = PU_SIZES[i % len(PU_SIZES)]
        o = orc.MeBiTask()
        o.x0 = int(rng.integers(0, (W - w) // 4 + 1)) * 4
        o.y0 = int(rng.integers(0, (H - h) // 4 + 1)) * 4
        if i % 7 == 3:  # picture corners: the clamps and the SAD4-group quirk
            o.x0, o.y0 = (0, 0) if i % 2 else (W - w, H - h)
        o.w, o.h = w, h
        far = i % 5 == 0
        for name, truth in (("mvStart", (12, 8)), ("mvOther", (24, 16))):
            v = np.array(truth) + (rng.integers(-300, 301, 2) if far else rng.integers(-9, 10, 2))
            getattr(o, name)[0], getattr(o, name)[1] = int(v[0]), int(v[1])
        for k in range(4):
            o.mvp[k] = int((12, 8)[k % 2] + rng.integers(-20, 21))
        o.rateMvpFlag[0], o.rateMvpFlag[1] = int(rng.integers(20000, 60000)), int(rng.integers(20000, 60000))
        o.lambda_ = int(rng.choice([0.025, 0.1, 0.3, 0.75]) * 65536 + 0.5)
        o.limitMin[0], o.limitMin[1] = -CTB - o.x0, -CTB - o.y0
        o.limitMax[0], o.limitMax[1] = W + CTB - o.x0 - w, H + CTB - o.y0 - h
        if i % 4 == 0:
            o.limitMax[0] = min(o.limitMax[0], (o.x0 // CTB) * CTB + 3 * CTB - o.x0 - w - 15)
            o.limitMax[1] = min(o.limitMax[1], (o.y0 // CTB) * CTB + 2 * CTB - o.y0 - h - 15)
        o.smallWindow, o.halfPel, o.quarterPel, o.bitDepth = i % 3 == 0, 1, i % 4 != 1, scene.bd
        fill_bi(tasks[i], o, scene.pics[0], scene.pics[1], scene.pics[2])
        otasks.append(o)
    return tasks, otasks


def check_bi_results(oracle, scene, otasks, got):
    """every field of every result against orc_me_bi_search; returns how many vectors the fractional rounds moved"""
    src, ref, other = (scene.host[k][0] for k in range(3))
    base = (PAD * src.shape[1] + PAD) * src.itemsize
    moved = 0
    for i, o in enumerate(otasks):
        r = orc.MeBiResult()
        oracle.lib.orc_me_bi_search(C.c_void_p(src.ctypes.data + base), src.shape[1], C.c_void_p(ref.ctypes.data + base),
                                    ref.shape[1], C.c_void_p(other.ctypes.data + base), other.shape[1], C.byref(o),
                                    C.byref(r), scene.bps)
        g = got[i]
        key = (i, (o.x0, o.y0, o.w, o.h), o.smallWindow, o.quarterPel)
        assert (int(g["mvInteger"]["x"]), int(g["mvInteger"]["y"])) == tuple(r.mvInteger), key
        assert (int(g["mv"]["x"]), int(g["mv"]["y"])) == tuple(r.mv), key
        assert (int(g["mvd"]["x"]), int(g["mvd"]["y"])) == tuple(r.mvd), key
        assert int(g["mvpFlag"]) == r.mvpFlag and int(g["cost"]) == r.cost, key
        assert int(g["nSad"]) == r.nSad, key
        moved += tuple(r.mv) != tuple(r.mvInteger)
    return moved


def test_me_bi_search_matches_oracle(scene, oracle):
    rng = np.random.default_rng(77)
    tasks, otasks = make_bi_tasks(rng, scene, 300)
    got = scene.ctx.me_bi_search(tasks)
    moved = check_bi_results(oracle, scene, otasks, got)
    assert moved > 30, moved


def test_me_bi_search_chain_matches_reference_templates(scene):
    """searchBi's L0 -> L1 chain (Search.hpp:1805-1823) on the device against the unmodified reference template."""
    import test_oracle_search_pin as pin
    if not pin.LIB.exists():
        pytest.skip("oracle/_ref/libsearch_ref.so not built")
    lib = C.CDLL(str(pin.LIB))
    lib.havoc_instruction_set_support.restype = C.c_int
    lib.ref_search_bi_batch.argtypes = [C.POINTER(pin.RefPictures), C.c_void_p, C.c_ssize_t, C.POINTER(pin.RefBiTask),
                                        C.POINTER(pin.RefBiResult), C.c_int]
    rng = np.random.default_rng(123)
    n = 320
    src, ref0, ref1 = (scene.host[k][0] for k in range(3))
    base = (PAD * src.shape[1] + PAD) * src.itemsize
    pics = pin.RefPictures(src.ctypes.data + base, ref0.ctypes.data + base, src.shape[1], ref0.shape[1], W, H, PAD,
                           scene.bps, 3, CTB, 0)
    rtasks = (pin.RefBiTask * n)(*[pin.make_bi_task(rng, i, scene.bd) for i in range(n)])
    want = (pin.RefBiResult * n)()
    assert lib.ref_search_bi_batch(C.byref(pics), ref1.ctypes.data + base, ref1.shape[1], rtasks, want, n) == 0

    mvs = [pin.bi_vectors(t) for t in rtasks]
    mvd = [[(t.mvd[0], t.mvd[1]), (t.mvd[2], t.mvd[3])] for t in rtasks]
    flag = [[t.mvpFlag[0], t.mvpFlag[1]] for t in rtasks]
    for lst in (0, 1):  # one batch per list; the second uses the first's refined vectors, as the encoder would
        idx = [i for i in range(n) if lst == 0 or rtasks[i].chain]
        tasks = np.zeros(len(idx), hvb.me_bi_task_t)
        for k, i in enumerate(idx):
            o = pin.oracle_bi_task(rtasks[i], want[i], lst, mvs[i][lst], mvs[i][1 - lst])
            fill_bi(tasks[k], o, scene.pics[0], scene.pics[1 + lst], scene.pics[2 - lst])
        got = scene.ctx.me_bi_search(tasks)
        for k, i in enumerate(idx):
            mvs[i] = list(mvs[i])
            mvs[i][lst] = (int(got[k]["mv"]["x"]), int(got[k]["mv"]["y"]))
            mvd[i][lst] = (int(got[k]["mvd"]["x"]), int(got[k]["mvd"]["y"]))
            flag[i][lst] = int(got[k]["mvpFlag"])
    for i in range(n):
        r = want[i]
        key = (i, (rtasks[i].x0, rtasks[i].y0, rtasks[i].w, rtasks[i].h), rtasks[i].speed, rtasks[i].chain)
        assert mvd[i] == [(r.mvd[0], r.mvd[1]), (r.mvd[2], r.mvd[3])], key
        assert flag[i] == [r.mvpFlag[0], r.mvpFlag[1]], key
