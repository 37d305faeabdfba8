"""ctypes faces of the two CPU checkers used by the tests (never by the product):

* ``Oracle``  -- oracle/liboracle.so, our own C restatement (oracle/oracle_*.c)
* ``Ref``     -- oracle/_ref/libhavoc_ref.so, the unmodified reference havoc library behind
                 oracle/ref_shim.cpp (present wherever oracle/Makefile `ref` was run)

Both take numpy arrays; strides are in samples, as in the reference.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
ORACLE_LIB = ORACLE_DIR / "liboracle.so"
REF_LIB = ORACLE_DIR / "_ref" / "libhavoc_ref.so"

vp = C.c_void_p
ip = C.c_ssize_t


def _ptr(a: np.ndarray, offset_elems: int = 0):
    return C.c_void_p(a.ctypes.data + offset_elems * a.itemsize)


def build_oracle() -> Path:
    srcs = list(ORACLE_DIR.glob("oracle_*.c")) + [ORACLE_DIR / "oracle.h"]
    if not ORACLE_LIB.exists() or ORACLE_LIB.stat().st_mtime < max(s.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(ORACLE_DIR), "port"], check=True, capture_output=True)
    return ORACLE_LIB


class MeTask(C.Structure):
    _fields_ = [
        ("x0", C.c_int), ("y0", C.c_int), ("w", C.c_int), ("h", C.c_int),
        ("mvp", C.c_int16 * 4),
        ("rateMvpFlag", C.c_int64 * 2),
        ("lambda_", C.c_int32),
        ("limitMin", C.c_int16 * 2), ("limitMax", C.c_int16 * 2),
        ("smallSearchWindow", C.c_int), ("met", C.c_int), ("log2CbSize", C.c_int),
        ("usePrev2Nx2N", C.c_int), ("prev2Nx2N", C.c_int16 * 2),
        ("halfPel", C.c_int), ("quarterPel", C.c_int), ("bitDepth", C.c_int),
    ]


class MeResult(C.Structure):
    _fields_ = [
        ("mv", C.c_int16 * 2), ("mvd", C.c_int16 * 2), ("mvInteger", C.c_int16 * 2), ("earlyExit", C.c_int),
        ("cost", C.c_int64), ("mvpFlag", C.c_int),
        ("costMvdZero", C.c_int64 * 2), ("subpelCost", C.c_int64), ("nSad", C.c_int),
    ]


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(str(build_oracle()))
        L = self.lib
        L.orc_sad.restype = C.c_int
        L.orc_sad.argtypes = [vp, ip, vp, ip, C.c_int, C.c_int, C.c_int]
        L.orc_sad_multiref4.argtypes = [vp, ip, C.POINTER(vp), ip, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int]
        L.orc_ssd.restype = C.c_uint32
        L.orc_ssd.argtypes = [vp, ip, vp, ip, C.c_int, C.c_int, C.c_int]
        L.orc_ssd_linear.argtypes = [vp, vp, C.c_int]
        L.orc_hadamard_satd.argtypes = [vp, ip, vp, ip, C.c_int, C.c_int]
        L.orc_measure_satd.argtypes = [vp, ip, vp, ip, C.c_int, C.c_int, C.c_int]
        L.orc_pred_uni.argtypes = [vp, ip, vp, ip] + [C.c_int] * 7
        L.orc_pred_bi.argtypes = [vp, ip, vp, vp, ip] + [C.c_int] * 9
        L.orc_subtract_bi.argtypes = [vp, ip, vp, ip, vp, ip] + [C.c_int] * 4
        L.orc_pred_intra.argtypes = [vp, ip, vp] + [C.c_int] * 5
        L.orc_transform_fwd.argtypes = [vp, vp, ip, C.c_int, C.c_int, C.c_int]
        L.orc_inverse_transform_add.argtypes = [vp, ip, vp, ip, vp] + [C.c_int] * 4
        L.orc_inverse_transform.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
        L.orc_quantize.argtypes = [vp, vp] + [C.c_int] * 4
        L.orc_quantize_inverse.argtypes = [vp, vp] + [C.c_int] * 3
        L.orc_quantize_reconstruct.argtypes = [vp, ip, vp, ip, vp, C.c_int]
        if hasattr(L, "orc_me_search"):
            L.orc_me_search.argtypes = [vp, ip, vp, ip, C.POINTER(MeTask), C.POINTER(MeResult), C.c_int]
            L.orc_rate_of_mvd.restype = C.c_int64
            L.orc_rate_of_mvd.argtypes = [C.c_int, C.c_int]

    # plane views: (array, offset in samples, stride in samples)
    def sad(self, a, oa, sa, b, ob, sb, w, h):
        return self.lib.orc_sad(_ptr(a, oa), sa, _ptr(b, ob), sb, w, h, a.itemsize)

    def sad4(self, a, oa, sa, b, obs, sb, w, h):
        refs = (vp * 4)(*[b.ctypes.data + o * b.itemsize for o in obs])
        out = (C.c_int * 4)()
        self.lib.orc_sad_multiref4(_ptr(a, oa), sa, refs, sb, out, w, h, a.itemsize)
        return list(out)

    def ssd(self, a, oa, sa, b, ob, sb, w, h):
        return self.lib.orc_ssd(_ptr(a, oa), sa, _ptr(b, ob), sb, w, h, a.itemsize)

    def ssd_linear(self, a, b):
        return self.lib.orc_ssd_linear(_ptr(a), _ptr(b), a.size)

    def hadamard_satd(self, a, oa, sa, b, ob, sb, log2n):
        return self.lib.orc_hadamard_satd(_ptr(a, oa), sa, _ptr(b, ob), sb, log2n, a.itemsize)

    def measure_satd(self, a, oa, sa, b, ob, sb, w, h):
        return self.lib.orc_measure_satd(_ptr(a, oa), sa, _ptr(b, ob), sb, w, h, a.itemsize)

    def pred_uni(self, dst, od, sd, ref, orf, sr, w, h, xf, yf, bit_depth, taps):
        self.lib.orc_pred_uni(_ptr(dst, od), sd, _ptr(ref, orf), sr, w, h, xf, yf, bit_depth, taps, ref.itemsize)

    def pred_bi(self, dst, od, sd, ref0, o0, ref1, o1, sr, w, h, xf0, yf0, xf1, yf1, bit_depth, taps):
        self.lib.orc_pred_bi(_ptr(dst, od), sd, _ptr(ref0, o0), _ptr(ref1, o1), sr, w, h, xf0, yf0, xf1, yf1,
                             bit_depth, taps, ref0.itemsize)

    def subtract_bi(self, dst, od, sd, pred, op, sp, src, os_, ss, w, h, bit_depth):
        self.lib.orc_subtract_bi(_ptr(dst, od), sd, _ptr(pred, op), sp, _ptr(src, os_), ss, w, h, bit_depth, src.itemsize)

    def pred_intra(self, dst, sd, neighbours, nb_index, mode, log2n, bit_depth, edge_flag):
        """nb_index = index of p(-1,-1) in `neighbours`; the reference pointer is one past it."""
        self.lib.orc_pred_intra(_ptr(dst), sd, _ptr(neighbours, nb_index + 1), mode, log2n, bit_depth, edge_flag,
                                neighbours.itemsize)

    def transform_fwd(self, coeffs, src, stride, tr_type, log2n, bit_depth):
        self.lib.orc_transform_fwd(_ptr(coeffs), _ptr(src), stride, tr_type, log2n, bit_depth)

    def inverse_transform_add(self, dst, sd, pred, sp, coeffs, tr_type, log2n, bit_depth):
        self.lib.orc_inverse_transform_add(_ptr(dst), sd, _ptr(pred), sp, _ptr(coeffs), tr_type, log2n, bit_depth,
                                           dst.itemsize)

    def inverse_transform(self, res, coeffs, tr_type, log2n, bit_depth):
        self.lib.orc_inverse_transform(_ptr(res), _ptr(coeffs), tr_type, log2n, bit_depth)

    def quantize(self, dst, src, scale, shift, offset):
        return self.lib.orc_quantize(_ptr(dst), _ptr(src), scale, shift, offset, src.size)

    def quantize_inverse(self, dst, src, scale, shift):
        self.lib.orc_quantize_inverse(_ptr(dst), _ptr(src), scale, shift, src.size)

    def quantize_reconstruct(self, rec, sr, pred, sp, res, n):
        self.lib.orc_quantize_reconstruct(_ptr(rec), sr, _ptr(pred), sp, _ptr(res), n)


class Ref:
    """The reference's own tables.  use_asm=False -> C_REF|C_OPT (the identity oracle)."""

    def __init__(self, use_asm: bool = False):
        if not REF_LIB.exists():
            raise FileNotFoundError(str(REF_LIB))
        self.lib = C.CDLL(str(REF_LIB))
        L = self.lib
        L.ref_create.restype = vp
        L.ref_create.argtypes = [C.c_int]
        L.ref_destroy.argtypes = [vp]
        L.ref_sad.argtypes = [vp, vp, ip, vp, ip, C.c_int, C.c_int, C.c_int]
        L.ref_sad_multiref4.argtypes = [vp, vp, ip, C.POINTER(vp), ip, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int]
        L.ref_ssd.restype = C.c_uint32
        L.ref_ssd.argtypes = [vp, vp, ip, vp, ip, C.c_int, C.c_int]
        L.ref_ssd_linear.argtypes = [vp, vp, vp, C.c_int]
        L.ref_hadamard_satd.argtypes = [vp, vp, ip, vp, ip, C.c_int, C.c_int]
        L.ref_pred_uni.argtypes = [vp, vp, ip, vp, ip] + [C.c_int] * 7
        L.ref_pred_bi.argtypes = [vp, vp, ip, vp, vp, ip] + [C.c_int] * 9
        L.ref_subtract_bi.argtypes = [vp, vp, ip, vp, ip, vp, ip] + [C.c_int] * 4
        L.ref_pred_intra.argtypes = [vp, vp, ip, vp] + [C.c_int] * 5
        L.ref_transform_fwd.argtypes = [vp, vp, vp, ip, C.c_int, C.c_int, C.c_int]
        L.ref_inverse_transform_add.argtypes = [vp, vp, ip, vp, ip, vp] + [C.c_int] * 4
        L.ref_inverse_transform.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int]
        L.ref_quantize.argtypes = [vp, vp, vp] + [C.c_int] * 4
        L.ref_quantize_inverse.argtypes = [vp, vp, vp] + [C.c_int] * 3
        L.ref_quantize_reconstruct.argtypes = [vp, vp, ip, vp, ip, vp, C.c_int]
        self.h = L.ref_create(1 if use_asm else 0)

    def __del__(self):
        try:
            self.lib.ref_destroy(self.h)
        except Exception:
            pass

    def sad(self, a, oa, sa, b, ob, sb, w, h):
        return self.lib.ref_sad(self.h, _ptr(a, oa), sa, _ptr(b, ob), sb, w, h, a.itemsize)

    def sad4(self, a, oa, sa, b, obs, sb, w, h):
        refs = (vp * 4)(*[b.ctypes.data + o * b.itemsize for o in obs])
        out = (C.c_int * 4)()
        self.lib.ref_sad_multiref4(self.h, _ptr(a, oa), sa, refs, sb, out, w, h, a.itemsize)
        return list(out)

    def ssd(self, a, oa, sa, b, ob, sb, log2n):
        return self.lib.ref_ssd(self.h, _ptr(a, oa), sa, _ptr(b, ob), sb, log2n, a.itemsize)

    def ssd_linear(self, a, b):
        return self.lib.ref_ssd_linear(self.h, _ptr(a), _ptr(b), a.size)

    def hadamard_satd(self, a, oa, sa, b, ob, sb, log2n):
        return self.lib.ref_hadamard_satd(self.h, _ptr(a, oa), sa, _ptr(b, ob), sb, log2n, a.itemsize)

    def pred_uni(self, dst, od, sd, ref, orf, sr, w, h, xf, yf, bit_depth, taps):
        return self.lib.ref_pred_uni(self.h, _ptr(dst, od), sd, _ptr(ref, orf), sr, w, h, xf, yf, bit_depth, taps,
                                     ref.itemsize)

    def pred_bi(self, dst, od, sd, ref0, o0, ref1, o1, sr, w, h, xf0, yf0, xf1, yf1, bit_depth, taps):
        return self.lib.ref_pred_bi(self.h, _ptr(dst, od), sd, _ptr(ref0, o0), _ptr(ref1, o1), sr, w, h, xf0, yf0,
                                    xf1, yf1, bit_depth, taps, ref0.itemsize)

    def subtract_bi(self, dst, od, sd, pred, op, sp, src, os_, ss, w, h, bit_depth):
        self.lib.ref_subtract_bi(self.h, _ptr(dst, od), sd, _ptr(pred, op), sp, _ptr(src, os_), ss, w, h, bit_depth,
                                 src.itemsize)

    def pred_intra(self, dst, sd, neighbours, nb_index, mode, log2n, bit_depth, c_idx):
        return self.lib.ref_pred_intra(self.h, _ptr(dst), sd, _ptr(neighbours, nb_index + 1), mode, log2n, bit_depth,
                                       c_idx, neighbours.itemsize)

    def transform_fwd(self, coeffs, src, stride, tr_type, log2n, bit_depth):
        self.lib.ref_transform_fwd(self.h, _ptr(coeffs), _ptr(src), stride, tr_type, log2n, bit_depth)

    def inverse_transform_add(self, dst, sd, pred, sp, coeffs, tr_type, log2n, bit_depth):
        self.lib.ref_inverse_transform_add(self.h, _ptr(dst), sd, _ptr(pred), sp, _ptr(coeffs), tr_type, log2n,
                                           bit_depth, dst.itemsize)

    def inverse_transform(self, res, coeffs, tr_type, log2n, bit_depth):
        self.lib.ref_inverse_transform(self.h, _ptr(res), _ptr(coeffs), tr_type, log2n, bit_depth)

    def quantize(self, dst, src, scale, shift, offset):
        return self.lib.ref_quantize(self.h, _ptr(dst), _ptr(src), scale, shift, offset, src.size)

    def quantize_inverse(self, dst, src, scale, shift):
        self.lib.ref_quantize_inverse(self.h, _ptr(dst), _ptr(src), scale, shift, src.size)

    def quantize_reconstruct(self, rec, sr, pred, sp, res, log2n):
        self.lib.ref_quantize_reconstruct(self.h, _ptr(rec), sr, _ptr(pred), sp, _ptr(res), log2n)


def have_ref() -> bool:
    return REF_LIB.exists()


class RdoqCtx(C.Structure):
    _fields_ = [("sig_coeff_flag", C.c_uint8 * 44), ("greater1_flag", C.c_uint8 * 24), ("greater2_flag", C.c_uint8 * 6),
                ("coded_sub_block_flag", C.c_uint8 * 4), ("last_x_prefix", C.c_uint8 * 18),
                ("last_y_prefix", C.c_uint8 * 18), ("cbf_luma", C.c_uint8 * 2), ("cbf_cbcr", C.c_uint8 * 5),
                ("rqt_root_cbf", C.c_uint8 * 1), ("reserved", C.c_uint8 * 6), ("lambda_", C.c_double)]


assert C.sizeof(RdoqCtx) == 136

QUANT_SCALE = [26214, 23302, 20560, 18396, 16384, 14564]  # turing/QpState.h scaleLookup
LEVEL_SCALE = [40, 45, 51, 57, 64, 72]                    # turing/QpState.h:85


def quant_params(qp: int, log2n: int, bit_depth: int):
    """(qscale, qshift, iqscale, iqshift) as the TU pipeline derives them (Reconstruct.cpp:779-786)."""
    qscale = QUANT_SCALE[qp % 6]
    qshift = 29 - bit_depth + qp // 6 - log2n
    iqscale = LEVEL_SCALE[qp % 6] << (qp // 6)
    iqshift = log2n - 1 + bit_depth - 8
    return qscale, qshift, iqscale, iqshift


def random_rdoq_ctx(rng, lam: float) -> np.ndarray:
    """136-byte context snapshot with random legal CABAC states (0..125)."""
    raw = rng.integers(0, 126, 136).astype(np.uint8)
    raw[128:136] = np.frombuffer(np.float64(lam).tobytes(), np.uint8)
    return raw


def _rdoq(fn, dst, src, ctx_bytes, qscale, qshift, iqscale, log2n, c_idx, scan_idx, is_intra, sdh, bit_depth):
    fn.argtypes = [vp, vp, vp] + [C.c_int] * 9
    return fn(_ptr(dst), _ptr(src), _ptr(ctx_bytes), qscale, qshift, iqscale, log2n, c_idx, scan_idx, is_intra,
              sdh, bit_depth)


def oracle_rdoq(oracle: "Oracle", *a):
    return _rdoq(oracle.lib.orc_rdoq, *a)


def oracle_rdoq_grouped(oracle: "Oracle", *a):
    """orc_rdoq with stage 1 decomposed by coefficient group (every (carry, right, below) variant, then a selection pass)"""
    return _rdoq(oracle.lib.orc_rdoq_grouped, *a)


def ref_rdoq(ref: "Ref", *a):
    return _rdoq(ref.lib.ref_rdoq, *a)


# ---- reference-sample filtering of the intra sweep (numpy restatement; pinned against the reference headers
# and the oracle's C version in tests/test_oracle_pin_intra_filter.py) ----

def filter_flag(c_idx, mode, n):
    """turing/Dsp.h:57-70"""
    lookup = [0b111000, 0, 0b111000] + [0b110000] * 6 + [0b100000, 0, 0b100000] + [0b110000] * 6 + [0b111000] + \
             [0b110000] * 6 + [0b100000, 0, 0b100000] + [0b110000] * 6 + [0b111000]
    return c_idx == 0 and bool(lookup[mode] & n)


def filtered_neighbours(u, n, bd, strong_enabled):
    """turing/IntraReferenceSamples.h:373-419 on an array whose corner is at index 2n"""
    c = 2 * n
    u = u.astype(np.int64)
    f = u.copy()
    top = lambda x: u[c + 1 + x]
    left = lambda y: u[c - 1 - y]
    strong = strong_enabled and n == 32 and abs(u[c] + top(63) - 2 * top(31)) < (1 << (bd - 5)) and \
        abs(u[c] + left(63) - 2 * left(31)) < (1 << (bd - 5))
    if strong:
        for k in range(63):
            f[c - 1 - k] = ((63 - k) * u[c] + (k + 1) * left(63) + 32) >> 6
            f[c + 1 + k] = ((63 - k) * u[c] + (k + 1) * top(63) + 32) >> 6
    else:
        f[1:4 * n] = (u[0:4 * n - 1] + 2 * u[1:4 * n] + u[2:4 * n + 1] + 2) >> 2
    return f


class MeBiTask(C.Structure):
    _fields_ = [
        ("x0", C.c_int), ("y0", C.c_int), ("w", C.c_int), ("h", C.c_int),
        ("mvOther", C.c_int16 * 2), ("mvStart", C.c_int16 * 2), ("mvp", C.c_int16 * 4),
        ("rateMvpFlag", C.c_int64 * 2), ("lambda_", C.c_int32),
        ("limitMin", C.c_int16 * 2), ("limitMax", C.c_int16 * 2),
        ("smallWindow", C.c_int), ("halfPel", C.c_int), ("quarterPel", C.c_int), ("bitDepth", C.c_int),
    ]


class MeBiResult(C.Structure):
    _fields_ = [("mv", C.c_int16 * 2), ("mvd", C.c_int16 * 2), ("mvInteger", C.c_int16 * 2), ("mvpFlag", C.c_int),
                ("cost", C.c_int64), ("nSad", C.c_int)]


class Plane(C.Structure):
    _fields_ = [("p", C.c_void_p), ("stride", C.c_ssize_t)]


class PuCostTask(C.Structure):
    _fields_ = [("x0", C.c_int), ("y0", C.c_int), ("w", C.c_int), ("h", C.c_int), ("predFlag", C.c_int * 2),
                ("mv", C.c_int16 * 4), ("picWidth", C.c_int), ("picHeight", C.c_int), ("bitDepthY", C.c_int),
                ("bitDepthC", C.c_int)]


def planes3(padded, pad):
    """(Plane * 3) over three edge-padded numpy planes (luma padded by `pad`, chroma by pad // 2)"""
    out = (Plane * 3)()
    for c, a in enumerate(padded):
        pd = pad if c == 0 else pad // 2
        out[c].p = a.ctypes.data + (pd * a.shape[1] + pd) * a.itemsize
        out[c].stride = a.shape[1]
    return out
