"""Host-side mirror of include/hvb.h: a thin ctypes binding over csrc/libhvb.so.

The reference's host side is C++ calling havoc function tables (turing/StateFunctionTables.h); the
C++ table shim over this ABI is csrc/havoc_b200.cpp.  This module is the same ABI for Python callers
(tests, bench): numpy structured arrays mirror the task structs one-to-one, and every call either
takes host arrays (HOST: staged through pinned memory inside the library) or raw device pointers
(DEVICE: e.g. ``torch.Tensor.data_ptr()``).

There is no CPU fallback: if the CUDA library is missing or no B200 is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "csrc" / "libhvb.so"

HOST, DEVICE = 0, 1

block_t = np.dtype([("pic", "<i2"), ("cIdx", "<i2"), ("x", "<i2"), ("y", "<i2")])
metric_task_t = np.dtype([("a", block_t), ("b", block_t), ("w", "<i2"), ("h", "<i2"), ("reserved", "<i4")])
sad4_task_t = np.dtype([("src", block_t), ("ref_pic", "<i2"), ("ref_cIdx", "<i2"), ("w", "<i2"), ("h", "<i2"),
                        ("rx", "<i2", 4), ("ry", "<i2", 4)])
pred_task_t = np.dtype([("dst", block_t), ("ref_pic", "<i2", 2), ("x", "<i2"), ("y", "<i2"), ("w", "<i2"),
                        ("h", "<i2"), ("mvx", "<i2", 2), ("mvy", "<i2", 2), ("reserved", "<i4")])
subtract_bi_task_t = np.dtype([("dst", block_t), ("pred", block_t), ("src", block_t), ("w", "<i2"), ("h", "<i2"),
                               ("reserved", "<i4")])
interp_satd_task_t = np.dtype([("src", block_t), ("ref_pic", "<i2"), ("reserved0", "<i2"), ("w", "<i2"),
                               ("h", "<i2"), ("mvx", "<i2"), ("mvy", "<i2")])
intra_task_t = np.dtype([("dst", block_t), ("nb", "<i4"), ("log2n", "i1"), ("mode", "i1"), ("edge_flag", "i1"),
                         ("reserved", "i1")])
intra_sweep_task_t = np.dtype([("src", block_t), ("nb_unfiltered", "<i4"), ("nb_filtered", "<i4"),
                               ("log2n", "i1"), ("cIdx", "i1"), ("strong_intra_smoothing", "i1"), ("reserved", "i1", 5)])
transform_task_t = np.dtype([("src", "<i4"), ("dst", "<i4"), ("src_stride", "<i4"), ("log2n", "i1"),
                             ("trType", "i1"), ("reserved", "<i2")])
quant_task_t = np.dtype([("src", "<i4"), ("dst", "<i4"), ("n", "<i4"), ("scale", "<i4"), ("shift", "<i4"),
                         ("offset", "<i4")])
ita_task_t = np.dtype([("dst", block_t), ("pred", block_t), ("coeffs", "<i4"), ("log2n", "i1"), ("trType", "i1"),
                       ("reserved", "<i2")])
tu_task_t = np.dtype([("src", block_t), ("pred", block_t), ("rec", block_t), ("levels", "<i4"), ("log2n", "i1"),
                      ("trType", "i1"), ("cIdx", "i1"), ("flags", "i1"), ("qscale", "<i4"), ("qshift", "<i4"),
                      ("qoffset", "<i4"), ("iqscale", "<i4"), ("iqshift", "<i4"), ("scanIdx", "i1"),
                      ("reserved", "i1", 3), ("rdoq_ctx", "<i4")])
tu_result_t = np.dtype([("ssd", "<u4"), ("ssdPred", "<u4"), ("cbf", "<i4"), ("status", "<i4"), ("sadQuad", "<u4", (4,))])
rdoq_ctx_t = np.dtype([("sig_coeff_flag", "u1", 44), ("greater1_flag", "u1", 24), ("greater2_flag", "u1", 6),
                       ("coded_sub_block_flag", "u1", 4), ("last_x_prefix", "u1", 18), ("last_y_prefix", "u1", 18),
                       ("cbf_luma", "u1", 2), ("cbf_cbcr", "u1", 5), ("rqt_root_cbf", "u1", 1),
                       ("reserved", "u1", 6), ("lambda", "<f8")])
rdoq_task_t = np.dtype([("src", "<i4"), ("dst", "<i4"), ("qscale", "<i4"), ("qshift", "<i4"), ("iqscale", "<i4"),
                        ("log2n", "i1"), ("cIdx", "i1"), ("scanIdx", "i1"), ("flags", "i1"), ("rdoq_ctx", "<i4")])
mv_t = np.dtype([("x", "<i2"), ("y", "<i2")])
me_task_t = np.dtype([("src_pic", "<i2"), ("ref_pic", "<i2"), ("x0", "<i2"), ("y0", "<i2"), ("w", "<i2"),
                      ("h", "<i2"), ("mvp", mv_t, 2), ("rateMvpFlag", "<i8", 2), ("lambda", "<i4"),
                      ("limitMin", mv_t), ("limitMax", mv_t), ("prev2Nx2N", mv_t), ("smallSearchWindow", "u1"),
                      ("met", "u1"), ("log2CbSize", "u1"), ("usePrev2Nx2N", "u1"), ("halfPel", "u1"),
                      ("quarterPel", "u1"), ("reserved", "u1", 2)], align=True)
pu_cost_task_t = np.dtype([("src_pic", "<i2"), ("dst_pic", "<i2"), ("ref_pic", "<i2", 2), ("x0", "<i2"), ("y0", "<i2"),
                           ("w", "<i2"), ("h", "<i2"), ("mvx", "<i2", 2), ("mvy", "<i2", 2)], align=True)
deblock_block_t = np.dtype([("data", "i1"), ("packedBs", "u1")], align=True)
deblock_ctu_t = np.dtype([("tc_offset_div2", "i1"), ("beta_offset_div2", "i1")], align=True)
deblock_task_t = np.dtype([("pic", "<i2"), ("edgeType", "<i2"), ("xBegin", "<i2"), ("yBegin", "<i2"), ("xEnd", "<i2"),
                           ("yEnd", "<i2"), ("cbQpOffset", "<i2"), ("crQpOffset", "<i2")], align=True)
sao_ctu_t = np.dtype([("left", "<i2"), ("top", "<i2"), ("right", "<i2"), ("bottom", "<i2"), ("topLeft", "u1"), ("topRight", "u1"),
                      ("bottomLeft", "u1"), ("bottomRight", "u1"),
                      ("plane", np.dtype([("typeIdx", "i1"), ("classOrBand", "i1"), ("offset", "<i2", 4)], align=True), 3)], align=True)
sao_task_t = np.dtype([("src_pic", "<i2"), ("dst_pic", "<i2"), ("ctuBegin", "<i2"), ("ctuEnd", "<i2"), ("lumaFlag", "u1"),
                       ("chromaFlag", "u1"), ("reserved", "<i2")], align=True)
sao_stats_task_t = np.dtype([("org_pic", "<i2"), ("rec_pic", "<i2"), ("cIdx", "<i2"), ("x0", "<i2"), ("y0", "<i2"), ("w", "<i2"),
                             ("h", "<i2"), ("reserved", "<i2")], align=True)
sao_stats_t = np.dtype([("edgeE", "<i4", (4, 5)), ("edgeCount", "<i4", (4, 5)), ("bandE", "<i4", 32), ("bandCount", "<i4", 32)], align=True)
coded_residual_task_t = np.dtype([("levels", "<i4"), ("log2n", "i1"), ("scanIdx", "i1"), ("reserved", "<i2")], align=True)
coded_residual_t = np.dtype([("offset", "<i4"), ("words", "<i4")], align=True)
intra_complexity_task_t = np.dtype([("pic", "<i2"), ("reserved", "<i2"), ("x0", "<i2"), ("y0", "<i2"), ("wBlocks", "<i2"),
                                    ("hBlocks", "<i2"), ("out", "<i4")], align=True)
aq_layer_task_t = np.dtype([("pic", "<i2"), ("unit", "<i2"), ("out", "<i4")], align=True)
scd_stats_task_t = np.dtype([("pic", "<i2"), ("margin", "<i2"), ("out", "<i4")], align=True)
me_bi_task_t = np.dtype([("src_pic", "<i2"), ("ref_pic", "<i2"), ("x0", "<i2"), ("y0", "<i2"), ("w", "<i2"),
                         ("h", "<i2"), ("mvp", mv_t, 2), ("other_pic", "<i2"), ("reserved0", "<i2"),
                         ("rateMvpFlag", "<i8", 2), ("lambda", "<i4"), ("limitMin", mv_t), ("limitMax", mv_t),
                         ("mvStart", mv_t), ("mvOther", mv_t), ("smallWindow", "u1"), ("halfPel", "u1"),
                         ("quarterPel", "u1"), ("reserved1", "u1")], align=True)
me_bi_result_t = np.dtype([("mv", mv_t), ("mvd", mv_t), ("mvInteger", mv_t), ("mvpFlag", "<i4"), ("cost", "<i8"),
                           ("nSad", "<i4"), ("reserved", "<i4")], align=True)
me_result_t = np.dtype([("mv", mv_t), ("mvd", mv_t), ("mvInteger", mv_t), ("mvpFlag", "<i4"), ("cost", "<i8"),
                        ("costMvdZero", "<i8", 2), ("subpelCost", "<i8"), ("nSad", "<i4"), ("flags", "<i4")],
                       align=True)

_SIZES = {
    "block": (block_t, 8), "metric": (metric_task_t, 24), "sad4": (sad4_task_t, 32), "pred": (pred_task_t, 32),
    "subtract_bi": (subtract_bi_task_t, 32), "interp_satd": (interp_satd_task_t, 20), "intra": (intra_task_t, 16),
    "intra_sweep": (intra_sweep_task_t, 24), "transform": (transform_task_t, 16), "quant": (quant_task_t, 24),
    "ita": (ita_task_t, 24), "tu": (tu_task_t, 60), "tu_result": (tu_result_t, 32), "rdoq_ctx": (rdoq_ctx_t, 136),
    "rdoq": (rdoq_task_t, 28), "me": (me_task_t, 64), "me_result": (me_result_t, 56),
    "pu_cost": (pu_cost_task_t, 24), "me_bi": (me_bi_task_t, 64), "me_bi_result": (me_bi_result_t, 32),
}
for _name, (_dt, _size) in _SIZES.items():
    assert _dt.itemsize == _size, (_name, _dt.itemsize, _size)


class HvbError(RuntimeError):
    pass


_lib = None


def load_library() -> C.CDLL:
    """dlopen csrc/libhvb.so; raises if it has not been built (``python -m turingcodec_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise HvbError(f"{LIB_PATH} is missing: build it with __graft_entry__.build(); there is no CPU fallback")
    lib = C.CDLL(str(LIB_PATH))
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    lib.hvb_create.argtypes = [i32, i32, i32, C.POINTER(vp)]
    lib.hvb_destroy.argtypes = [vp]
    lib.hvb_destroy.restype = None
    lib.hvb_last_error.argtypes = [vp]
    lib.hvb_last_error.restype = C.c_char_p
    lib.hvb_set_stream.argtypes = [vp, vp]
    lib.hvb_sync.argtypes = [vp]
    lib.hvb_set_pipelined.argtypes = [vp, C.c_int]
    lib.hvb_launch_count.argtypes = [vp]
    lib.hvb_launch_count.restype = i64
    lib.hvb_device_ok.argtypes = [i32]
    lib.hvb_picture_create.argtypes = [vp, i32, i32, i32, C.POINTER(i32)]
    lib.hvb_picture_destroy.argtypes = [vp, i32]
    lib.hvb_picture_upload.argtypes = [vp, i32, i32, vp, C.c_ssize_t, i32, i32]
    lib.hvb_picture_download.argtypes = [vp, i32, i32, vp, C.c_ssize_t, i32, i32]
    lib.hvb_picture_pad.argtypes = [vp, i32]
    if hasattr(lib, "hvb_picture_copy"):
        lib.hvb_picture_copy.argtypes = [vp, i32, i32]
    lib.hvb_picture_upload_rect.argtypes = [vp, i32, i32, vp, C.c_ssize_t, i32, i32, i32, i32]
    lib.hvb_picture_download_rect.argtypes = [vp, i32, i32, vp, C.c_ssize_t, i32, i32, i32, i32]
    lib.hvb_picture_plane.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(C.c_ssize_t)]
    lib.hvb_pool_upload.argtypes = [vp, vp, C.c_size_t, C.c_size_t]
    lib.hvb_coeff_upload.argtypes = [vp, vp, C.c_size_t, C.c_size_t]
    lib.hvb_coeff_download.argtypes = [vp, vp, C.c_size_t, C.c_size_t]
    lib.hvb_rdoq_contexts_upload.argtypes = [vp, vp, i32, i32]
    for name in ("hvb_sad_batch", "hvb_ssd_batch", "hvb_satd_batch", "hvb_sad4_batch", "hvb_interp_satd_batch",
                 "hvb_intra_satd35_batch", "hvb_quantize_batch", "hvb_tu_chain_batch", "hvb_rdoq_batch",
                 "hvb_me_search_batch", "hvb_me_bi_search_batch", "hvb_pu_cost_batch", "hvb_sao_stats_batch"):
        if hasattr(lib, name):
            getattr(lib, name).argtypes = [vp, vp, i32, vp, i32]
    if hasattr(lib, "hvb_intra_complexity_batch"):
        lib.hvb_intra_complexity_batch.argtypes = [vp, vp, i32, vp, i32, i32]
    if hasattr(lib, "hvb_aq_activity_batch"):
        lib.hvb_aq_activity_batch.argtypes = [vp, vp, i32, vp, i32, i32]
        lib.hvb_scd_histogram_batch.argtypes = [vp, vp, i32, vp, i32]
        lib.hvb_scd_block_stats_batch.argtypes = [vp, vp, i32, vp, i32, i32]
    if hasattr(lib, "hvb_coded_residual_batch"):
        lib.hvb_coded_residual_batch.argtypes = [vp, vp, i32, i32, i32, vp, i32]
    if hasattr(lib, "hvb_deblock_info_upload"):
        lib.hvb_deblock_info_upload.argtypes = [vp, i32, vp, i32, i32, vp, i32, i32, i32]
        lib.hvb_sao_info_upload.argtypes = [vp, i32, vp, i32]
    for name in ("hvb_pred_batch", "hvb_subtract_bi_batch", "hvb_intra_pred_batch", "hvb_transform_fwd_batch",
                 "hvb_transform_inv_batch", "hvb_quantize_inverse_batch", "hvb_inverse_transform_add_batch", "hvb_deblock_batch",
                 "hvb_sao_batch"):
        if hasattr(lib, name):
            getattr(lib, name).argtypes = [vp, vp, i32, i32]
    # entry points of the submission queue's fast paths (hvb_encoder.cpp): completion flags, page-locked arrays the kernels address
    # themselves, one device allocation for a pool of pictures
    for name, types in (("hvb_host_alloc", [vp, C.c_size_t, C.POINTER(vp)]), ("hvb_host_free", [vp, vp]), ("hvb_signal", [vp, vp, i32]),
                        ("hvb_poll", [vp]), ("hvb_picture_reserve", [vp, i32, i32, i32, i32]), ("hvb_coeff_pool_wrap", [vp, vp, C.c_size_t]),
                        ("hvb_rdoq_contexts_wrap", [vp, vp, i32]), ("hvb_set_tu_fused_max", [vp, i32])):
        if hasattr(lib, name):
            getattr(lib, name).argtypes = types
    _lib = lib
    return lib


def exported_symbols() -> list[str]:
    """Every function include/hvb.h declares (used by the CPU-side ABI test)."""
    import re
    text = (Path(__file__).resolve().parent.parent / "include" / "hvb.h").read_text()
    return sorted(set(re.findall(r"\b(hvb_[a-z0-9_]+)\s*\(", text)))


def _as_ptr(x):
    """numpy array -> host pointer; int -> raw (device) pointer."""
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return C.c_void_p(x.ctypes.data)
    return C.c_void_p(int(x))


class Context:
    """One hvb_context (one CUDA stream, one host thread)."""

    def __init__(self, device: int = 0, bytes_per_sample: int = 1, bit_depth: int = 8):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.hvb_create(device, bytes_per_sample, bit_depth, C.byref(self.h))
        if rc != 0:
            raise HvbError(f"hvb_create failed with status {rc} (no usable sm_100 device or bad arguments); "
                           "this library has no CPU fallback")
        self.bps = bytes_per_sample
        self.bit_depth = bit_depth
        self.sample_dtype = np.uint8 if bytes_per_sample == 1 else np.uint16

    def close(self):
        if self.h:
            self.lib.hvb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise HvbError(f"{what}: status {rc}: {self.lib.hvb_last_error(self.h).decode()}")

    # -- plumbing ---------------------------------------------------------------------------
    def set_stream(self, cuda_stream: int | None):
        self._check(self.lib.hvb_set_stream(self.h, C.c_void_p(cuda_stream or 0)), "hvb_set_stream")

    def sync(self):
        self._check(self.lib.hvb_sync(self.h), "hvb_sync")

    def set_tma(self, on: bool):
        self._check(self.lib.hvb_set_tma(self.h, 1 if on else 0), "hvb_set_tma")

    def host_alloc(self, nbytes: int) -> int:
        """page-locked, device-addressable host memory (address as int); free with host_free"""
        p = C.c_void_p()
        self._check(self.lib.hvb_host_alloc(self.h, nbytes, C.byref(p)), "hvb_host_alloc")
        return p.value

    def host_free(self, address: int):
        self._check(self.lib.hvb_host_free(self.h, C.c_void_p(address)), "hvb_host_free")

    def signal(self, flag_address: int, value: int):
        """after everything enqueued so far has completed, `value` is stored to the int32 at flag_address (memory from host_alloc)"""
        self._check(self.lib.hvb_signal(self.h, C.c_void_p(flag_address), int(value)), "hvb_signal")

    def poll(self) -> int:
        return self.lib.hvb_poll(self.h)

    def picture_reserve(self, width: int, height: int, pad: int, count: int):
        """one device allocation for the planes of the next `count` pictures of this geometry"""
        self._check(self.lib.hvb_picture_reserve(self.h, width, height, pad, count), "hvb_picture_reserve")

    def coeff_pool_wrap(self, address: int | None, count: int = 0):
        self._check(self.lib.hvb_coeff_pool_wrap(self.h, C.c_void_p(address or 0), count), "hvb_coeff_pool_wrap")

    def rdoq_contexts_wrap(self, address: int | None, count: int = 0):
        self._check(self.lib.hvb_rdoq_contexts_wrap(self.h, C.c_void_p(address or 0), count), "hvb_rdoq_contexts_wrap")

    def set_tu_fused_max(self, blocks: int):
        """batches of at most `blocks` transform blocks take the one-launch form of tu_chain (0: always the staged kernels)"""
        self._check(self.lib.hvb_set_tu_fused_max(self.h, int(blocks)), "hvb_set_tu_fused_max")

    def set_pipelined(self, on: bool):
        """HOST calls on page-locked arrays only enqueue (copies overlap the kernels); results are valid after sync()."""
        self._check(self.lib.hvb_set_pipelined(self.h, int(bool(on))), "hvb_set_pipelined")

    @property
    def launch_count(self) -> int:
        return int(self.lib.hvb_launch_count(self.h))

    # -- pictures ---------------------------------------------------------------------------
    def picture_create(self, width: int, height: int, pad: int = 96) -> int:
        pic = C.c_int(-1)
        self._check(self.lib.hvb_picture_create(self.h, width, height, pad, C.byref(pic)), "hvb_picture_create")
        return pic.value

    def picture_destroy(self, pic: int):
        self._check(self.lib.hvb_picture_destroy(self.h, pic), "hvb_picture_destroy")

    def picture_upload(self, pic: int, c_idx: int, plane: np.ndarray, y0: int = 0, rows: int | None = None):
        plane = np.ascontiguousarray(plane, dtype=self.sample_dtype)
        rows = plane.shape[0] - y0 if rows is None else rows
        self._check(self.lib.hvb_picture_upload(self.h, pic, c_idx, _as_ptr(plane), plane.shape[1], y0, rows),
                    "hvb_picture_upload")

    def picture_download(self, pic: int, c_idx: int, width: int, height: int) -> np.ndarray:
        out = np.zeros((height, width), self.sample_dtype)
        self._check(self.lib.hvb_picture_download(self.h, pic, c_idx, _as_ptr(out), width, 0, height),
                    "hvb_picture_download")
        return out

    def picture_copy(self, dst_pic: int, src_pic: int):
        self._check(self.lib.hvb_picture_copy(self.h, dst_pic, src_pic), "hvb_picture_copy")

    def picture_pad(self, pic: int):
        self._check(self.lib.hvb_picture_pad(self.h, pic), "hvb_picture_pad")

    def picture_plane(self, pic: int, c_idx: int) -> tuple[int, int]:
        ptr, stride = C.c_void_p(), C.c_ssize_t()
        self._check(self.lib.hvb_picture_plane(self.h, pic, c_idx, C.byref(ptr), C.byref(stride)), "hvb_picture_plane")
        return int(ptr.value), int(stride.value)

    def upload_yuv(self, pic: int, y: np.ndarray, u: np.ndarray, v: np.ndarray, pad: bool = True):
        for c, p in enumerate((y, u, v)):
            self.picture_upload(pic, c, p)
        if pad:
            self.picture_pad(pic)

    # -- pools -------------------------------------------------------------------------------
    def pool_upload(self, samples: np.ndarray, offset: int = 0):
        samples = np.ascontiguousarray(samples, dtype=self.sample_dtype)
        self._check(self.lib.hvb_pool_upload(self.h, _as_ptr(samples), samples.size, offset), "hvb_pool_upload")

    def coeff_upload(self, data: np.ndarray, offset: int = 0):
        data = np.ascontiguousarray(data, dtype=np.int16)
        self._check(self.lib.hvb_coeff_upload(self.h, _as_ptr(data), data.size, offset), "hvb_coeff_upload")

    def coeff_download(self, count: int, offset: int = 0) -> np.ndarray:
        out = np.zeros(count, np.int16)
        self._check(self.lib.hvb_coeff_download(self.h, _as_ptr(out), count, offset), "hvb_coeff_download")
        return out

    def rdoq_contexts_upload(self, snapshots: np.ndarray, first: int = 0):
        snapshots = np.ascontiguousarray(snapshots, dtype=rdoq_ctx_t)
        self._check(self.lib.hvb_rdoq_contexts_upload(self.h, _as_ptr(snapshots), snapshots.size, first),
                    "hvb_rdoq_contexts_upload")

    # -- batched calls -----------------------------------------------------------------------
    def _with_out(self, fn_name, tasks, n, out, out_dtype, out_shape, mem):
        fn = getattr(self.lib, fn_name)
        if mem == HOST:
            n = tasks.size
            if out is None:
                out = np.zeros(out_shape(n), out_dtype)
        self._check(fn(self.h, _as_ptr(tasks), n, _as_ptr(out), mem), fn_name)
        return out

    def _no_out(self, fn_name, tasks, n, mem):
        if mem == HOST:
            n = tasks.size
        self._check(getattr(self.lib, fn_name)(self.h, _as_ptr(tasks), n, mem), fn_name)

    def sad(self, tasks, n=None, out=None, mem=HOST):
        return self._with_out("hvb_sad_batch", tasks, n, out, np.int32, lambda k: (k,), mem)

    def ssd(self, tasks, n=None, out=None, mem=HOST):
        return self._with_out("hvb_ssd_batch", tasks, n, out, np.uint32, lambda k: (k,), mem)

    def satd(self, tasks, n=None, out=None, mem=HOST):
        return self._with_out("hvb_satd_batch", tasks, n, out, np.int32, lambda k: (k,), mem)

    def sad4(self, tasks, n=None, out=None, mem=HOST):
        return self._with_out("hvb_sad4_batch", tasks, n, out, np.int32, lambda k: (k, 4), mem)

    def pred(self, tasks, n=None, mem=HOST):
        self._no_out("hvb_pred_batch", tasks, n, mem)

    def subtract_bi(self, tasks, n=None, mem=HOST):
        self._no_out("hvb_subtract_bi_batch", tasks, n, mem)

    def interp_satd(self, tasks, n=None, out=None, mem=HOST):
        return self._with_out("hvb_interp_satd_batch", tasks, n, out, np.int32, lambda k: (k,), mem)

    def intra_pred(self, tasks, n=None, mem=HOST):
        self._no_out("hvb_intra_pred_batch", tasks, n, mem)

    def intra_satd35(self, tasks, n=None, out=None, mem=HOST):
        return self._with_out("hvb_intra_satd35_batch", tasks, n, out, np.int32, lambda k: (k, 35), mem)

    def transform_fwd(self, tasks, n=None, mem=HOST):
        self._no_out("hvb_transform_fwd_batch", tasks, n, mem)

    def transform_inv(self, tasks, n=None, mem=HOST):
        self._no_out("hvb_transform_inv_batch", tasks, n, mem)

    def quantize(self, tasks, n=None, out=None, mem=HOST):
        return self._with_out("hvb_quantize_batch", tasks, n, out, np.int32, lambda k: (k,), mem)

    def quantize_inverse(self, tasks, n=None, mem=HOST):
        self._no_out("hvb_quantize_inverse_batch", tasks, n, mem)

    def inverse_transform_add(self, tasks, n=None, mem=HOST):
        self._no_out("hvb_inverse_transform_add_batch", tasks, n, mem)

    def tu_chain(self, tasks, n=None, out=None, mem=HOST):
        return self._with_out("hvb_tu_chain_batch", tasks, n, out, tu_result_t, lambda k: (k,), mem)

    def rdoq(self, tasks, n=None, out=None, mem=HOST):
        return self._with_out("hvb_rdoq_batch", tasks, n, out, np.int32, lambda k: (k,), mem)

    def me_search(self, tasks, n=None, out=None, mem=HOST):
        return self._with_out("hvb_me_search_batch", tasks, n, out, me_result_t, lambda k: (k,), mem)

    def pu_cost(self, tasks, n=None, out=None, mem=HOST):
        """-> int32 [n][3]: SATD of Y, Cb, Cr of each PU's inter prediction"""
        return self._with_out("hvb_pu_cost_batch", tasks, n, out, np.int32, lambda k: (k, 3), mem)

    def deblock_info_upload(self, pic: int, blocks: np.ndarray, ctus: np.ndarray, pic_width_in_ctbs: int, pic_height_in_ctbs: int,
                            ctb_log2: int):
        """blocks: [rows][stride] of deblock_block_t (the reference's 8x8 grid), ctus: [n] of deblock_ctu_t"""
        blocks = np.ascontiguousarray(blocks, dtype=deblock_block_t)
        ctus = np.ascontiguousarray(ctus, dtype=deblock_ctu_t)
        self._check(self.lib.hvb_deblock_info_upload(self.h, pic, _as_ptr(blocks), blocks.shape[1], blocks.shape[0], _as_ptr(ctus),
                                                     pic_width_in_ctbs, pic_height_in_ctbs, ctb_log2), "hvb_deblock_info_upload")

    def deblock(self, tasks, n=None, mem=HOST):
        self._no_out("hvb_deblock_batch", tasks, n, mem)

    def sao_info_upload(self, pic: int, ctus: np.ndarray):
        ctus = np.ascontiguousarray(ctus, dtype=sao_ctu_t)
        self._check(self.lib.hvb_sao_info_upload(self.h, pic, _as_ptr(ctus), ctus.size), "hvb_sao_info_upload")

    def intra_complexity(self, tasks, out_count: int) -> np.ndarray:
        """-> int32 [out_count]: the 8x8 blocks' AC Hadamard energies, placed by each task's `out` offset (host tasks)"""
        tasks = np.ascontiguousarray(tasks, dtype=intra_complexity_task_t)
        out = np.zeros(out_count, np.int32)
        self._check(self.lib.hvb_intra_complexity_batch(self.h, _as_ptr(tasks), tasks.size, _as_ptr(out), out_count, HOST),
                    "hvb_intra_complexity_batch")
        return out

    def aq_activity(self, tasks, out_count: int) -> np.ndarray:
        """-> int64 [out_count]: per unit of each layer task the smallest quadrant variance (AdaptiveQuantisation's activity - 1)"""
        tasks = np.ascontiguousarray(tasks, dtype=aq_layer_task_t)
        out = np.zeros(out_count, np.int64)
        self._check(self.lib.hvb_aq_activity_batch(self.h, _as_ptr(tasks), tasks.size, _as_ptr(out), out_count, HOST), "hvb_aq_activity_batch")
        return out

    def scd_histogram(self, pics) -> np.ndarray:
        """-> int32 [n, 64]: ShotChangeDetection's luma histogram of each picture"""
        pics = np.ascontiguousarray(pics, dtype=np.int16)
        out = np.zeros((pics.size, 64), np.int32)
        self._check(self.lib.hvb_scd_histogram_batch(self.h, _as_ptr(pics), pics.size, _as_ptr(out), HOST), "hvb_scd_histogram_batch")
        return out

    def scd_block_stats(self, tasks, out_count: int) -> np.ndarray:
        """-> float64 [out_count]: (avg, var) pairs of getLikelihood's block grids, placed by each task's `out`"""
        tasks = np.ascontiguousarray(tasks, dtype=scd_stats_task_t)
        out = np.zeros(out_count, np.float64)
        self._check(self.lib.hvb_scd_block_stats_batch(self.h, _as_ptr(tasks), tasks.size, _as_ptr(out), out_count, HOST),
                    "hvb_scd_block_stats_batch")
        return out

    def coded_residual(self, tasks, records_base: int, capacity_words: int) -> np.ndarray:
        """-> coded_residual_t [n + 1] (host tasks): per block (offset, words), last entry (end of the used region, overflow)"""
        tasks = np.ascontiguousarray(tasks, dtype=coded_residual_task_t)
        out = np.zeros(tasks.size + 1, coded_residual_t)
        self._check(self.lib.hvb_coded_residual_batch(self.h, _as_ptr(tasks), tasks.size, records_base, capacity_words, _as_ptr(out), HOST),
                    "hvb_coded_residual_batch")
        return out

    def sao_stats(self, tasks, n=None, out=None, mem=HOST):
        return self._with_out("hvb_sao_stats_batch", tasks, n, out, sao_stats_t, lambda k: (k,), mem)

    def sao(self, tasks, n=None, mem=HOST):
        self._no_out("hvb_sao_batch", tasks, n, mem)

    def me_bi_search(self, tasks, n=None, out=None, mem=HOST):
        return self._with_out("hvb_me_bi_search_batch", tasks, n, out, me_bi_result_t, lambda k: (k,), mem)
