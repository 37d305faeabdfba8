"""GPU parity: measurePuCost's distortion (predictInter uni/bi + SATD of Y, Cb, Cr, turing/Search.hpp:1668-1682)
on the device vs the oracle and vs the unmodified reference workers (predictUni / predictBi / measureSatd), incl. the
stored prediction samples, vectors far outside the picture (clipMvLumaComponent) and the chroma-skip rule."""
import ctypes as C

import numpy as np
import pytest

import orc
import test_oracle_pu_cost_pin as pin
from gpu_common import H, PAD, W, Scene
from turingcodec_b200 import hvb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[(1, 8), (2, 10)], ids=["u8", "u16-10bit"])
def scene(request):
    s = Scene(*request.param)
    yield s
    s.close()


def to_hvb(tasks, scene, dst_pic):
    t = np.zeros(len(tasks), hvb.pu_cost_task_t)
    for i, r in enumerate(tasks):
        t[i]["src_pic"], t[i]["dst_pic"] = scene.pics[0], dst_pic
        t[i]["ref_pic"] = (scene.pics[1] if r.predFlag[0] else -1, scene.pics[2] if r.predFlag[1] else -1)
        t[i]["x0"], t[i]["y0"], t[i]["w"], t[i]["h"] = r.x0, r.y0, r.w, r.h
        t[i]["mvx"], t[i]["mvy"] = (r.mv[0], r.mv[2]), (r.mv[1], r.mv[3])
    return t


def test_pu_cost_matches_oracle_and_reference(scene, oracle):
    rng = np.random.default_rng(3)
    n = 330
    tasks = (pin.RefPuTask * n)(*[pin.make_pu_task(rng, i) for i in range(n)])
    got = scene.ctx.pu_cost(to_hvb(tasks, scene, -1))
    assert got.shape == (n, 3)

    planes = [orc.planes3(scene.host[k], PAD) for k in range(3)]
    oracle.lib.orc_pu_cost.argtypes = [C.c_void_p] * 3 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    for i in range(n):
        o, satd = pin.oracle_task(tasks[i], scene.bd), (C.c_int32 * 3)()
        oracle.lib.orc_pu_cost(planes[0], planes[1], planes[2], C.byref(o), satd, None, scene.bps)
        assert list(got[i]) == list(satd), (i, (o.x0, o.y0, o.w, o.h), tuple(o.predFlag), tuple(o.mv))
    assert (got[:, 1] == 0).sum() > 10 and (got[:, 0] > 0).all()

    if pin.LIB.exists():  # and straight against the reference's own predictUni / predictBi / measureSatd
        want, _ = pin.run_reference(C.CDLL(str(pin.LIB)), scene.host, scene.bps, scene.bd, False, tasks, want_pred=False)
        assert np.array_equal(got, want)


def test_pu_cost_stores_the_prediction(scene, oracle):
    """dst_pic >= 0: non-overlapping PUs leave predictInter's samples in the destination picture"""
    rng = np.random.default_rng(4)
    cells = [(x, y) for y in range(0, H, 64) for x in range(0, W, 64)]
    tasks = (pin.RefPuTask * len(cells))()
    for i, (x, y) in enumerate(cells):
        t = pin.make_pu_task(rng, i)
        t.x0, t.y0 = x, y  # one PU per 64x64 cell, anchored at the cell origin
        tasks[i] = t
    dst = scene.scratch[0]
    got = scene.ctx.pu_cost(to_hvb(tasks, scene, dst))
    planes = [orc.planes3(scene.host[k], PAD) for k in range(3)]
    oracle.lib.orc_pu_cost.argtypes = [C.c_void_p] * 3 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    pics = [scene.download(dst, c) for c in range(3)]
    for i, t in enumerate(tasks):
        o, satd = pin.oracle_task(t, scene.bd), (C.c_int32 * 3)()
        bufs = [np.zeros(64 * 64, scene.dtype) for _ in range(3)]
        out = (C.c_void_p * 3)(*[b.ctypes.data for b in bufs])
        oracle.lib.orc_pu_cost(planes[0], planes[1], planes[2], C.byref(o), satd, out, scene.bps)
        assert list(got[i]) == list(satd), i
        for c in range(3):
            sh = int(c > 0)
            w, h = t.w >> sh, t.h >> sh
            block = pics[c][(t.y0 >> sh):(t.y0 >> sh) + h, (t.x0 >> sh):(t.x0 >> sh) + w]
            assert np.array_equal(block, bufs[c][:w * h].reshape(h, w)), (i, c)
