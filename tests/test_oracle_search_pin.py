"""Pins oracle_search.c (the restated motion-search control flow) against the UNMODIFIED reference:
oracle/_ref/libsearch_ref.so instantiates the reference's own fullPelMotionEstimation / subPelRefinement templates
(turing/Search.hpp:2064-2357) with a stand-in handler (oracle/ref_shim_search.cpp) and runs them on the same
pictures and the same per-PU state.  Every output the reference produces must match: final mv / mvd, the integer
vector, best cost, mvp flag, costMvdZero and the mvPreviousInteger2Nx2N side effect.

Inputs are given to the reference in ITS terms (context state, reciprocalSqrtLambda, Speed preset, part mode,
concurrent-frames, CTB position); the oracle's task (limits, Q16 lambda, flag rates, switches) is derived here the
way INTEGRATION.md tells a maintainer to derive it, so the derivation is pinned too."""
import ctypes as C

import numpy as np
import pytest

import orc
from turingcodec_b200 import synth

W, H, PAD, CTB = 256, 192, 96, 64
LIB = orc.ORACLE_DIR / "_ref" / "libsearch_ref.so"

PU_SIZES = [(64, 64), (64, 32), (32, 64), (32, 32), (32, 16), (16, 32), (16, 16), (16, 8), (8, 16), (8, 8), (8, 4),
            (4, 8), (32, 24), (24, 32), (16, 12), (12, 16), (64, 48), (48, 64), (64, 16), (16, 64), (32, 8), (8, 32),
            (16, 4), (4, 16)]


class RefTask(C.Structure):
    _fields_ = [("x0", C.c_int), ("y0", C.c_int), ("w", C.c_int), ("h", C.c_int),
                ("cqtX0", C.c_int), ("cqtY0", C.c_int), ("log2CbSize", C.c_int), ("cqtDepth", C.c_int),
                ("partMode", C.c_int), ("mvp", C.c_int16 * 4), ("mvpFlagState", C.c_int),
                ("reciprocalSqrtLambda", C.c_double), ("speed", C.c_int), ("met", C.c_int),
                ("concurrentFrames", C.c_int), ("xCtb", C.c_int), ("yCtb", C.c_int), ("prev2Nx2N", C.c_int16 * 2),
                ("bitDepth", C.c_int), ("refList", C.c_int)]


class RefResult(C.Structure):
    _fields_ = [("mv", C.c_int16 * 2), ("mvd", C.c_int16 * 2), ("mvInteger", C.c_int16 * 2),
                ("prev2Nx2NAfter", C.c_int16 * 2), ("cost", C.c_int64), ("mvpFlag", C.c_int), ("reserved", C.c_int),
                ("costMvdZero", C.c_int64 * 2), ("rateMvpFlag", C.c_int64 * 2), ("lambda_", C.c_int32),
                ("reserved2", C.c_int32)]


class RefPictures(C.Structure):
    _fields_ = [("src", C.c_void_p), ("ref", C.c_void_p), ("strideSrc", C.c_ssize_t), ("strideRef", C.c_ssize_t),
                ("width", C.c_int), ("height", C.c_int), ("pad", C.c_int), ("bps", C.c_int), ("isa", C.c_int),
                ("ctbSize", C.c_int), ("lzcnt", C.c_int)]


@pytest.fixture(scope="module")
def searchref():
    if not LIB.exists():
        pytest.skip("oracle/_ref/libsearch_ref.so not built (make -C oracle searchref, needs /root/reference)")
    lib = C.CDLL(str(LIB))
    lib.ref_search_batch.argtypes = [C.POINTER(RefPictures), C.POINTER(RefTask), C.POINTER(RefResult), C.c_int]
    lib.havoc_instruction_set_support.restype = C.c_int
    return lib


def planes(bps, bit_depth, frame):
    dtype = np.uint8 if bps == 1 else np.uint16
    luma = synth.frame(frame, W, H, bit_depth)[0].astype(dtype)
    if bps == 2 and frame == 1:
        luma[::7, ::5] = (1 << bit_depth) - 1
    return np.pad(luma, PAD, mode="edge")


def make_ref_task(rng, i, bit_depth, frame_distance):
    t = RefTask()
    w, h = PU_SIZES[i % len(PU_SIZES)]
    log2cb = 3 + (max(w, h) > 8) + (max(w, h) > 16) + (max(w, h) > 32)
    cb = 1 << log2cb
    # the CU sits on its own grid inside the picture; the PU is one of its partitions
    t.cqtX0 = int(rng.integers(0, (W - cb) // cb + 1)) * cb
    t.cqtY0 = int(rng.integers(0, (H - cb) // cb + 1)) * cb
    t.log2CbSize = log2cb
    t.x0 = t.cqtX0 + (cb - w if i % 2 else 0)
    t.y0 = t.cqtY0 + (cb - h if i % 2 else 0)
    t.w, t.h = w, h
    t.partMode = 0 if (w == cb and h == cb) else (1 if w == cb else 2 if h == cb else 3)  # 2Nx2N, 2NxN, Nx2N, NxN
    t.cqtDepth = int(rng.integers(0, 2)) if log2cb == 6 else 6 - log2cb
    if log2cb == 6 and i % 3 == 0:
        t.cqtDepth = 0
    truth = np.array([12, 8]) * frame_distance
    spread = [2, 12, 40, 160][i % 4]
    for k in range(2):
        v = truth + rng.integers(-2, 3, 2) if i % 3 == 1 else rng.integers(-spread, spread + 1, 2)
        t.mvp[2 * k], t.mvp[2 * k + 1] = int(v[0]), int(v[1])
    t.mvpFlagState = int(rng.integers(0, 126))
    t.reciprocalSqrtLambda = float(rng.choice([0.05, 0.11, 0.2, 0.37, 0.6, 1.5]))
    t.speed = i % 3
    t.met = (i // 3) % 2
    t.concurrentFrames = 4 if i % 5 == 0 else 1
    t.xCtb, t.yCtb = (t.x0 // CTB) * CTB, (t.y0 // CTB) * CTB
    p = rng.integers(-12, 13, 2) * 4
    t.prev2Nx2N[0], t.prev2Nx2N[1] = int(p[0]), int(p[1])
    t.bitDepth = bit_depth
    t.refList = i % 2
    return t


def oracle_task(t: RefTask, r: RefResult) -> orc.MeTask:
    """the derivation INTEGRATION.md section 2 prescribes for hvb_me_task / orc_me_task"""
    o = orc.MeTask()
    o.x0, o.y0, o.w, o.h = t.x0, t.y0, t.w, t.h
    for k in range(4):
        o.mvp[k] = t.mvp[k]
    o.rateMvpFlag[0], o.rateMvpFlag[1] = r.rateMvpFlag[0], r.rateMvpFlag[1]
    o.lambda_ = r.lambda_
    o.limitMin[0], o.limitMin[1] = -CTB - t.x0, -CTB - t.y0
    o.limitMax[0], o.limitMax[1] = W + CTB - t.x0 - t.w, H + CTB - t.y0 - t.h
    if t.concurrentFrames > 1:
        o.limitMax[0] = min(o.limitMax[0], t.xCtb + 3 * CTB - t.x0 - t.w - 15)
        o.limitMax[1] = min(o.limitMax[1], t.yCtb + 2 * CTB - t.y0 - t.h - 15)
    o.smallSearchWindow = int(t.speed >= 2)          # Speed::useSmallSearchWindow
    o.met = t.met
    o.log2CbSize = t.log2CbSize
    o.usePrev2Nx2N = int(t.partMode != 0 or t.cqtDepth != 0)
    o.prev2Nx2N[0], o.prev2Nx2N[1] = t.prev2Nx2N[0], t.prev2Nx2N[1]
    o.halfPel = 1                                    # Speed::doHalfPelRefinement
    o.quarterPel = int(t.speed <= 1)                 # Speed::doQuarterPelRefinement
    o.bitDepth = t.bitDepth
    return o


CASES = [
    ("u8-c", 1, 8, False, 0),
    ("u8-jit-lzcnt", 1, 8, True, 1),
    ("u16-10bit-c", 2, 10, False, 0),
    ("u16-10bit-jit-lzcnt", 2, 10, True, 1),
]


@pytest.mark.parametrize("tag,bps,bit_depth,jit,lzcnt", CASES, ids=[c[0] for c in CASES])
def test_search_control_flow_matches_reference(searchref, oracle, tag, bps, bit_depth, jit, lzcnt):
    rng = np.random.default_rng(7 + bps + lzcnt)
    n = 480
    for frame_distance in (1, 2):
        src, ref = planes(bps, bit_depth, 0), planes(bps, bit_depth, frame_distance)
        base = (PAD * src.shape[1] + PAD) * src.itemsize
        pics = RefPictures(src.ctypes.data + base, ref.ctypes.data + base, src.shape[1], ref.shape[1], W, H, PAD, bps,
                           searchref.havoc_instruction_set_support() if jit else 3, CTB, lzcnt)
        tasks = (RefTask * n)(*[make_ref_task(rng, i, bit_depth, frame_distance) for i in range(n)])
        results = (RefResult * n)()
        assert searchref.ref_search_batch(C.byref(pics), tasks, results, n) == 0

        early = refined = moved = 0
        for i in range(n):
            t, r = tasks[i], results[i]
            o, g = oracle_task(t, r), orc.MeResult()
            oracle.lib.orc_me_search(C.c_void_p(src.ctypes.data + base), src.shape[1],
                                     C.c_void_p(ref.ctypes.data + base), ref.shape[1], C.byref(o), C.byref(g), bps)
            key = (tag, frame_distance, i, (t.x0, t.y0, t.w, t.h), t.speed, t.met)
            assert tuple(g.mvInteger) == tuple(r.mvInteger), key
            assert g.cost == r.cost and g.mvpFlag == r.mvpFlag, key
            assert tuple(g.mv) == tuple(r.mv) and tuple(g.mvd) == tuple(r.mvd), key
            # side effects: costMvdZero entries are written as the predictors are visited; the previous-2Nx2N vector
            # is replaced only when a 2Nx2N search runs to completion (Search.hpp:2332-2335)
            written = [z for z in r.costMvdZero if z != 0]
            assert list(g.costMvdZero)[:len(written)] == written, key
            if not g.earlyExit:
                assert len(written) == 2, key
            want_prev = tuple(r.mvInteger) if (t.partMode == 0 and not g.earlyExit) else tuple(t.prev2Nx2N)
            assert tuple(r.prev2Nx2NAfter) == want_prev, key
            early += g.earlyExit
            refined += tuple(g.mv) != tuple(g.mvInteger)
            moved += tuple(g.mvInteger) != (0, 0)
        assert early > 10 and refined > 20 and moved > 100, (early, refined, moved)


# ---- searchMotionBi (Search.hpp:1498-1653) ------------------------------------------------------------------

class RefBiTask(C.Structure):
    _fields_ = [("x0", C.c_int), ("y0", C.c_int), ("w", C.c_int), ("h", C.c_int),
                ("mvp", C.c_int16 * 8), ("mvd", C.c_int16 * 4), ("mvpFlag", C.c_int * 2), ("mvpFlagState", C.c_int),
                ("reciprocalSqrtLambda", C.c_double), ("speed", C.c_int), ("concurrentFrames", C.c_int),
                ("xCtb", C.c_int), ("yCtb", C.c_int), ("bitDepth", C.c_int), ("chain", C.c_int)]


class RefBiResult(C.Structure):
    _fields_ = [("mvd", C.c_int16 * 4), ("mvpFlag", C.c_int * 2), ("rateMvpFlag", C.c_int64 * 2),
                ("lambdaHalf", C.c_int32), ("reserved", C.c_int32)]


def make_bi_task(rng, i, bit_depth):
    t = RefBiTask()
    w, h = PU_SIZES[i % len(PU_SIZES)]
    if w + h == 12:  # bi-prediction is not allowed for 8x4 / 4x8 (Search.hpp:1872)
        w, h = 8, 8
    t.x0 = int(rng.integers(0, (W - w) // 4 + 1)) * 4
    t.y0 = int(rng.integers(0, (H - h) // 4 + 1)) * 4
    t.w, t.h = w, h
    if i % 7 == 3:  # hug a picture corner so that the LimitFullPelMv clamps (and their SAD4 grouping quirk) engage
        t.x0, t.y0 = (0, 0) if i % 2 else (W - w, H - h)
    # list 0 points one frame back, list 1 two frames back (synth motion is +3,+2 samples per frame)
    for lst, truth in ((0, np.array([12, 8])), (1, np.array([24, 16]))):
        spread = [1, 3, 9, 30][i % 4]
        for k in range(2):
            v = truth + rng.integers(-spread, spread + 1, 2)
            t.mvp[lst * 4 + k * 2], t.mvp[lst * 4 + k * 2 + 1] = int(v[0]), int(v[1])
        d = rng.integers(-6, 7, 2) if i % 5 else rng.integers(-300, 301, 2)  # a few far outside: clamping
        t.mvd[lst * 2], t.mvd[lst * 2 + 1] = int(d[0]), int(d[1])
        t.mvpFlag[lst] = int(rng.integers(0, 2))
    t.mvpFlagState = int(rng.integers(0, 126))
    t.reciprocalSqrtLambda = float(rng.choice([0.05, 0.11, 0.2, 0.37, 0.6, 1.5]))
    t.speed = i % 3
    t.concurrentFrames = 4 if i % 5 == 0 else 1
    t.xCtb, t.yCtb = (t.x0 // CTB) * CTB, (t.y0 // CTB) * CTB
    t.bitDepth = bit_depth
    t.chain = i % 2
    return t


def oracle_bi_task(t: RefBiTask, r: RefBiResult, lst: int, mv_this, mv_other) -> orc.MeBiTask:
    o = orc.MeBiTask()
    o.x0, o.y0, o.w, o.h = t.x0, t.y0, t.w, t.h
    o.mvOther[0], o.mvOther[1] = mv_other
    o.mvStart[0], o.mvStart[1] = mv_this
    for k in range(4):
        o.mvp[k] = t.mvp[lst * 4 + k]
    o.rateMvpFlag[0], o.rateMvpFlag[1] = r.rateMvpFlag[0], r.rateMvpFlag[1]
    o.lambda_ = r.lambdaHalf
    o.limitMin[0], o.limitMin[1] = -CTB - t.x0, -CTB - t.y0
    o.limitMax[0], o.limitMax[1] = W + CTB - t.x0 - t.w, H + CTB - t.y0 - t.h
    if t.concurrentFrames > 1:
        o.limitMax[0] = min(o.limitMax[0], t.xCtb + 3 * CTB - t.x0 - t.w - 15)
        o.limitMax[1] = min(o.limitMax[1], t.yCtb + 2 * CTB - t.y0 - t.h - 15)
    o.smallWindow = int(t.speed >= 2)   # Speed::useBiSmallSearchWindow
    o.halfPel, o.quarterPel = 1, int(t.speed <= 1)
    o.bitDepth = t.bitDepth
    return o


def s16(v):
    return ((int(v) + 0x8000) & 0xFFFF) - 0x8000


def bi_vectors(t: RefBiTask):
    """puData.mv(X) = mvp[X][flag] + mvd (turing/Mvp.h:781-799), int16 arithmetic"""
    return [(s16(t.mvp[l * 4 + t.mvpFlag[l] * 2] + t.mvd[l * 2]), s16(t.mvp[l * 4 + t.mvpFlag[l] * 2 + 1] + t.mvd[l * 2 + 1]))
            for l in range(2)]


@pytest.mark.parametrize("tag,bps,bit_depth,jit,lzcnt", CASES, ids=[c[0] for c in CASES])
def test_bi_search_matches_reference(searchref, oracle, tag, bps, bit_depth, jit, lzcnt):
    searchref.ref_search_bi_batch.argtypes = [C.POINTER(RefPictures), C.c_void_p, C.c_ssize_t, C.POINTER(RefBiTask),
                                              C.POINTER(RefBiResult), C.c_int]
    rng = np.random.default_rng(17 + bps + lzcnt)
    n = 360
    src, ref0, ref1 = planes(bps, bit_depth, 0), planes(bps, bit_depth, 1), planes(bps, bit_depth, 2)
    base = (PAD * src.shape[1] + PAD) * src.itemsize
    pics = RefPictures(src.ctypes.data + base, ref0.ctypes.data + base, src.shape[1], ref0.shape[1], W, H, PAD, bps,
                       searchref.havoc_instruction_set_support() if jit else 3, CTB, lzcnt)
    tasks = (RefBiTask * n)(*[make_bi_task(rng, i, bit_depth) for i in range(n)])
    results = (RefBiResult * n)()
    assert searchref.ref_search_bi_batch(C.byref(pics), ref1.ctypes.data + base, ref1.shape[1], tasks, results, n) == 0

    refs = [ref0, ref1]
    moved = 0
    for i in range(n):
        t, r = tasks[i], results[i]
        mv = bi_vectors(t)
        want_mvd = [(t.mvd[0], t.mvd[1]), (t.mvd[2], t.mvd[3])]
        want_flag = [t.mvpFlag[0], t.mvpFlag[1]]
        for lst in ((0, 1) if t.chain else (0,)):
            o, g = oracle_bi_task(t, r, lst, mv[lst], mv[1 - lst]), orc.MeBiResult()
            oracle.lib.orc_me_bi_search(C.c_void_p(src.ctypes.data + base), src.shape[1],
                                        C.c_void_p(refs[lst].ctypes.data + base), refs[lst].shape[1],
                                        C.c_void_p(refs[1 - lst].ctypes.data + base), refs[1 - lst].shape[1],
                                        C.byref(o), C.byref(g), bps)
            moved += tuple(g.mv) != mv[lst]
            mv[lst] = tuple(g.mv)
            want_mvd[lst], want_flag[lst] = tuple(g.mvd), g.mvpFlag
        key = (tag, i, (t.x0, t.y0, t.w, t.h), t.speed, t.chain)
        assert [(r.mvd[0], r.mvd[1]), (r.mvd[2], r.mvd[3])] == want_mvd, key
        assert [r.mvpFlag[0], r.mvpFlag[1]] == want_flag, key
    assert moved > 100, moved
