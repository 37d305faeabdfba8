"""Pins the oracle's restatement of measurePuCost's distortion (predictInter + SATD of Y/Cb/Cr,
turing/Search.hpp:1668-1682) against the UNMODIFIED reference workers: predictUni / predictBi (turing/Dsp.h:769-864,
incl. clipMvLumaComponent and the chroma origin/vector derivation) and measureSatd (turing/Measure.h:96-135), called
directly by oracle/ref_shim_search.cpp.  SATDs and the predicted samples themselves must be equal."""
import ctypes as C

import numpy as np
import pytest

import orc
from turingcodec_b200 import synth

W, H, PAD = 256, 192, 96
LIB = orc.ORACLE_DIR / "_ref" / "libsearch_ref.so"
PU_SIZES = [(64, 64), (64, 32), (32, 64), (32, 32), (32, 16), (16, 32), (16, 16), (16, 8), (8, 16), (8, 8), (8, 4),
            (4, 8), (32, 24), (24, 32), (16, 12), (12, 16), (64, 48), (48, 64), (64, 16), (16, 64), (32, 8), (8, 32),
            (16, 4), (4, 16)]


class RefPuPlanes(C.Structure):
    _fields_ = [("p", C.c_void_p * 3), ("stride", C.c_ssize_t * 3)]


class RefPuTask(C.Structure):
    _fields_ = [("x0", C.c_int), ("y0", C.c_int), ("w", C.c_int), ("h", C.c_int), ("predFlag", C.c_int * 2),
                ("mv", C.c_int16 * 4)]


def padded_frame(bps, bit_depth, index):
    dtype = np.uint8 if bps == 1 else np.uint16
    f = [p.astype(dtype) for p in synth.frame(index, W, H, bit_depth)]
    if bps == 2 and index == 1:
        f[0][::7, ::5] = (1 << bit_depth) - 1
        f[1][::5, ::3] = (1 << bit_depth) - 1
    return [np.ascontiguousarray(np.pad(p, PAD if c == 0 else PAD // 2, mode="edge")) for c, p in enumerate(f)]


def make_pu_task(rng, i):
    t = RefPuTask()
    w, h = PU_SIZES[i % len(PU_SIZES)]
    t.x0 = int(rng.integers(0, (W - w) // 4 + 1)) * 4
    t.y0 = int(rng.integers(0, (H - h) // 4 + 1)) * 4
    t.w, t.h = w, h
    mode = i % 3  # L0 only, L1 only, bi (never for 8x4 / 4x8)
    if mode == 2 and w + h == 12:
        mode = 0
    t.predFlag[0], t.predFlag[1] = int(mode != 1), int(mode != 0)
    for l in range(2):
        if i % 6 == 5:  # far outside: clipMvLumaComponent pins the block just beyond the picture edge
            v = rng.choice([-1, 1], 2) * rng.integers(600, 2000, 2)
        else:
            v = np.array([12, 8]) * (l + 1) + rng.integers(-40, 41, 2)
        t.mv[2 * l], t.mv[2 * l + 1] = int(v[0]), int(v[1])
    return t


def ref_planes(frames):
    out = (RefPuPlanes * 3)()
    for k, f in enumerate(frames):
        for c, a in enumerate(f):
            pd = PAD if c == 0 else PAD // 2
            out[k].p[c] = a.ctypes.data + (pd * a.shape[1] + pd) * a.itemsize
            out[k].stride[c] = a.shape[1]
    return out


def run_reference(lib, frames, bps, bit_depth, jit, tasks, want_pred=True):
    n = len(tasks)
    lib.ref_pu_cost_batch.argtypes = [C.POINTER(RefPuPlanes)] + [C.c_int] * 7 + [C.POINTER(RefPuTask), C.c_void_p, C.c_void_p, C.c_int]
    lib.havoc_instruction_set_support.restype = C.c_int
    satd = np.zeros((n, 3), np.int32)
    pred = np.zeros((n, 3, 64 * 64), np.uint8 if bps == 1 else np.uint16)
    rc = lib.ref_pu_cost_batch(ref_planes(frames), W, H, PAD, bps, lib.havoc_instruction_set_support() if jit else 3,
                               bit_depth, bit_depth, tasks, satd.ctypes.data, pred.ctypes.data if want_pred else None, n)
    assert rc == 0
    return satd, pred


def oracle_task(t: RefPuTask, bit_depth) -> orc.PuCostTask:
    o = orc.PuCostTask()
    o.x0, o.y0, o.w, o.h = t.x0, t.y0, t.w, t.h
    o.predFlag[0], o.predFlag[1] = t.predFlag[0], t.predFlag[1]
    for k in range(4):
        o.mv[k] = t.mv[k]
    o.picWidth, o.picHeight, o.bitDepthY, o.bitDepthC = W, H, bit_depth, bit_depth
    return o


@pytest.mark.parametrize("bps,bit_depth,jit", [(1, 8, False), (1, 8, True), (2, 10, False), (2, 10, True), (2, 9, False)])
def test_pu_cost_matches_reference(oracle, bps, bit_depth, jit):
    if not LIB.exists():
        pytest.skip("oracle/_ref/libsearch_ref.so not built (make -C oracle searchref, needs /root/reference)")
    lib = C.CDLL(str(LIB))
    rng = np.random.default_rng(5 + bit_depth + jit)
    frames = [padded_frame(bps, bit_depth, k) for k in range(3)]
    n = 360
    tasks = (RefPuTask * n)(*[make_pu_task(rng, i) for i in range(n)])
    want, want_pred = run_reference(lib, frames, bps, bit_depth, jit, tasks)
    planes = [orc.planes3(f, PAD) for f in frames]
    oracle.lib.orc_pu_cost.argtypes = [C.c_void_p] * 3 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    clipped = 0
    for i in range(n):
        t = tasks[i]
        o = oracle_task(t, bit_depth)
        satd = (C.c_int32 * 3)()
        bufs = [np.zeros(64 * 64, want_pred.dtype) for _ in range(3)]
        out = (C.c_void_p * 3)(*[b.ctypes.data for b in bufs])
        oracle.lib.orc_pu_cost(planes[0], planes[1], planes[2], C.byref(o), satd, out, bps)
        key = (i, (t.x0, t.y0, t.w, t.h), tuple(t.predFlag), tuple(t.mv))
        for c in range(3):
            count = (t.w >> (c > 0)) * (t.h >> (c > 0))
            assert np.array_equal(bufs[c][:count], want_pred[i, c, :count]), (key, c)
        assert list(satd) == list(want[i]), key
        clipped += any(abs(v) > 500 for v in t.mv)
    assert clipped > 20 and (want[:, 1] == 0).sum() > 10  # the edge clamp and the chroma-skip rule both occur
