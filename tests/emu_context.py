"""An hvb.Context look-alike whose batch calls run the library's kernels under the warp-level host emulator
(tests/host_emu_warp.py) on host memory -- test infrastructure, never product code.

With it the GPU parity tests themselves (tests/test_gpu_*.py) run in the CPU-only suite: tests/test_emulated_gpu_suite.py
hands their test functions an EmuScene instead of a gpu_common.Scene.  Pictures are edge-padded host arrays, the sample /
coefficient / context pools numpy arrays; every batch call launches the same kernels in the same order as the entry points
of csrc/*.cu do (a small grid: the kernels walk their work in grid-stride loops or pull it from a cursor)."""
import ctypes as C
import tempfile
from pathlib import Path

import numpy as np

import host_emu_warp
from turingcodec_b200 import hvb, synth

GRID = 2


class Plane(C.Structure):
    _fields_ = [("base", C.c_void_p), ("stride", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("pad", C.c_int32),
                ("reserved", C.c_int32)]


def _both(kernel_call: str) -> str:
    """`KERNEL<Sample>(args)` launched for the context's sample type; grid and block are named in the call text"""
    return ("if (bps == 1) { typedef uint8_t Sample; " + kernel_call + " } else { typedef uint16_t Sample; " + kernel_call + " }")


ENTRIES = {
    "hvb_metrics.cu": dict(
        strip=("template <typename Task>\nint gridFor(",), extra_headers=("hvb_satd.cuh",), namespaces=2,
        replace={
            "__device__ __forceinline__ void cpAsync8(": "static inline void cpAsync8(uint32_t dst, const void *src) { memcpy(emu::sharedArena + dst, src, 8); }",
        "__device__ __forceinline__ void cpAsync4(": "static inline void cpAsync4(uint32_t dst, const void *src) { memcpy(emu::sharedArena + dst, src, 4); }",
        "__device__ __forceinline__ void cpAsync16(": "static inline void cpAsync16(uint32_t dst, const void *src) { memcpy(emu::sharedArena + dst, src, 16); }",
            "__device__ __forceinline__ void cpAsyncCommit(": "static inline void cpAsyncCommit() {}",
            "template <int PENDING>\n__device__ __forceinline__ void cpAsyncWait(": "template <int PENDING> static inline void cpAsyncWait() {}",
        },
        entry=r'''
extern "C" void emu_sad(const HvbPlane *planes, const hvb_metric_task *tasks, int n, int32_t *out, int bps, int grid)
{ ''' + _both("emuLaunch(grid, kWarpsPerBlock * 32, [&] { sadKernel<Sample>(planes, tasks, n, out); });") + r''' }
extern "C" void emu_sad4(const HvbPlane *planes, const hvb_sad4_task *tasks, int n, int32_t *out, int bps, int grid)
{ ''' + _both("emuLaunch(grid, kWarpsPerBlock * 32, [&] { sad4Kernel<Sample>(planes, tasks, n, out); });") + r''' }
extern "C" void emu_ssd(const HvbPlane *planes, const hvb_metric_task *tasks, int n, uint32_t *out, int bps, int grid)
{ ''' + _both("emuLaunch(grid, kWarpsPerBlock * 32, [&] { ssdKernel<Sample>(planes, tasks, n, out); });") + r''' }
extern "C" void emu_satd(const HvbPlane *planes, const hvb_metric_task *tasks, int n, int32_t *out, int bps, int grid)
{
    int leftover[3] = {0, 0, 0}; // [0]: blocks for satdKernel, [2]: blocks for satdMmaSmallKernel
    if (bps == 1)
    {
        emuLaunch(grid, kWarpsPerBlock * 32, [&] { satdMmaKernel<2>(planes, tasks, n, out, leftover); });
        emuLaunch(grid, kWarpsPerBlock * 32, [&] { satdMmaSmallKernel(planes, tasks, n, out, leftover); });
        emuLaunch(grid, kWarpsPerBlock * 32, [&] { satdKernel<uint8_t>(planes, tasks, n, out, leftover); });
    }
    else
        {
            emuLaunch(grid, kWarpsPerBlock * 32, [&] { satdMma16Kernel<2>(planes, tasks, n, out, leftover); });
            emuLaunch(grid, kWarpsPerBlock * 32, [&] { satdKernel<uint16_t>(planes, tasks, n, out, leftover); });
        }
}
'''),
    "hvb_pred.cu": dict(
        strip=("int gridWarps(hvb_context *ctx",), extra_headers=("hvb_satd.cuh", "hvb_interp.cuh"), use_unit_header=False,
        entry=r'''
extern "C" void emu_pred(const HvbPlane *planes, const hvb_pred_task *tasks, int n, int bitDepth, int bps, int grid)
{ ''' + _both("emuLaunch(grid, kWarps * 32, [&] { predKernel<Sample>(planes, tasks, n, bitDepth); });") + r''' }
extern "C" void emu_interp_satd(const HvbPlane *planes, const hvb_interp_satd_task *tasks, int n, int32_t *out, int bitDepth, int bps, int grid)
{ ''' + _both("emuLaunch(grid, kWarps * 32, [&] { interpSatdKernel<Sample>(planes, tasks, n, out, bitDepth); });") + r''' }
extern "C" void emu_subtract_bi(const HvbPlane *planes, const hvb_subtract_bi_task *tasks, int n, int bitDepth, int bps, int grid)
{ ''' + _both("emuLaunch(grid, 256, [&] { subtractBiKernel<Sample>(planes, tasks, n, bitDepth); });") + r''' }
'''),
    "hvb_intra.cu": dict(
        strip=("int gridWarps(hvb_context *ctx",), use_unit_header=False, mma_wrappers={"imma16832": False},
        entry=r'''
extern "C" void emu_intra_pred(const HvbPlane *planes, const void *pool, const hvb_intra_task *tasks, int n, int bitDepth, int bps, int grid)
{ ''' + _both("emuLaunch(grid, kWarps * 32, [&] { intraPredKernel<Sample>(planes, static_cast<const Sample *>(pool), tasks, n, bitDepth); });") + r''' }
extern "C" void emu_intra_sweep(const HvbPlane *planes, const void *pool, const hvb_intra_sweep_task *tasks, int n, int32_t *out, int bitDepth,
                                int bps, int grid)
{ ''' + _both("emuLaunch(grid, kWarps * 32, [&] { intraSweepKernel8<Sample>(planes, static_cast<const Sample *>(pool), tasks, n, out, bitDepth); });") + r''' }
'''),
    "hvb_tu.cu": dict(
        use_unit_header=False, mma_wrappers={"immaS8U8": False, "immaS8S8": True}, extra_headers=("hvb_rdoq.cuh",),
        entry=r'''
struct RdoqTables
{
    std::vector<int2> bits;
    std::vector<int> last;
    RdoqTables(const hvb_rdoq_ctx *rdoqCtx, int nCtx)
    {
        // what initRdoqTables and hvbLaunchRdoqBits set up on the device
        for (int log2 = 2; log2 <= 5; ++log2)
            for (int scanIdx = 0; scanIdx < 3; ++scanIdx)
                for (int sp = 0; sp < (1 << (2 * log2)); ++sp)
                    hvb_rdoq::gScanTable[((log2 - 2) * 3 + scanIdx) * 1024 + sp] = (short)hvb_rdoq::scanToRaster(log2, scanIdx, sp);
        const int bytes = nCtx * (int)sizeof(hvb_rdoq_ctx), entries = nCtx * hvb_rdoq::kLastTabPerCtx;
        bits.resize(bytes + 1);
        last.resize(entries + 1);
        if (nCtx)
        {
            emuLaunch((bytes + 255) / 256, 256, [&] { rdoqBitsKernel(rdoqCtx, bits.data(), bytes); });
            emuLaunch((entries + 255) / 256, 256, [&] { rdoqLastKernel(rdoqCtx, last.data(), nCtx); });
        }
    }
};
extern "C" void emu_transform(int16_t *pool, const hvb_transform_task *tasks, int n, int bitDepth, int inverse, int grid)
{
    emuLaunch(grid, kWarps * 32, [&] { transformKernel(pool, tasks, n, bitDepth, inverse); });
}
extern "C" void emu_quant(int16_t *pool, const hvb_quant_task *tasks, int n, int32_t *cbf, int inverse, int grid)
{
    emuLaunch(grid, 256, [&] { quantKernel(pool, tasks, n, cbf, inverse); });
}
extern "C" void emu_ita(const HvbPlane *planes, const int16_t *pool, const hvb_ita_task *tasks, int n, int bitDepth, int bps, int grid)
{ ''' + _both("emuLaunch(grid, kWarps * 32, [&] { itaKernel<Sample>(planes, pool, tasks, n, bitDepth); });") + r''' }
extern "C" void emu_rdoq(int16_t *pool, int poolCount, const hvb_rdoq_ctx *rdoqCtx, int nCtx, const hvb_rdoq_task *tasks, int n, int32_t *cbf,
                         int bitDepth, int grid)
{
    RdoqTables tables(rdoqCtx, nCtx);
    std::vector<HvbCoefRec> recs(poolCount);
    std::vector<HvbRdoqMid> mids(n);
    const int chunks = (n + 1023) / 1024;
    std::vector<int> compact(n), chunkBase(chunks + 1);
    emuLaunch(chunks, 1024, [&] { scanLocalKernel(RdoqCount{tasks}, n, compact.data(), chunkBase.data()); });
    emuLaunch(1, 1024, [&] { scanBlocksKernel(chunkBase.data(), chunks); });
    emuLaunch(grid, kWarps * 32, [&] { rdoqPrepassKernel(pool, rdoqCtx, tasks, n, mids.data(), cbf, bitDepth); });
    emuLaunch(std::min((n + 127) / 128, grid), 128, [&] { rdoqThreadKernel(pool, recs.data(), rdoqCtx, tasks, n, mids.data(), cbf, bitDepth,
                                                                           tables.bits.data(), tables.last.data(), compact.data(), chunkBase.data()); });
}
template <typename Sample>
static void chain(const HvbPlane *planes, int16_t *pool, int poolCount, const hvb_rdoq_ctx *rdoqCtx, int nCtx, const hvb_tu_task *tasks, int n,
                  hvb_tu_result *out, int bitDepth, int grid)
{
    RdoqTables tables(rdoqCtx, nCtx);
    std::vector<int16_t> coefTmp(poolCount);
    std::vector<HvbCoefRec> recs(poolCount);
    std::vector<HvbRdoqMid> mids(n);
    std::vector<int> buckets(64 + n, 0);
    int *order = buckets.data() + 64;
    const int chunks = (n + 1023) / 1024;
    std::vector<int> compact(n), chunkBase(chunks + 1);
    emuLaunch(chunks, 1024, [&] { scanLocalKernel(TuCount{tasks}, n, compact.data(), chunkBase.data()); });
    emuLaunch(1, 1024, [&] { scanBlocksKernel(chunkBase.data(), chunks); });
    emuLaunch(grid, kWarps * 32, [&] { tuFrontKernel<Sample>(planes, pool, coefTmp.data(), rdoqCtx, tasks, n, out, mids.data(), buckets.data(), bitDepth, compact.data(), chunkBase.data(), nCtx, (unsigned)poolCount); });
    emuLaunch((n + 255) / 256, 256, [&] { tuOrderKernel(mids.data(), n, buckets.data(), buckets.data() + kRdoqBuckets, order); });
    emuLaunch(std::min((n + 127) / 128, grid), 128, [&] { tuRdoqKernel(pool, coefTmp.data(), recs.data(), rdoqCtx, tasks, n, out, mids.data(),
                                                                       buckets.data(), order, bitDepth, tables.bits.data(), tables.last.data(), compact.data(),
                                                                       chunkBase.data()); });
    emuLaunch(grid, kWarps * 32, [&] { tuBackKernel<Sample>(planes, pool, tasks, n, out, bitDepth); });
}
template <typename Sample>
static void fusedChain(const HvbPlane *planes, int16_t *pool, int poolCount, const hvb_rdoq_ctx *rdoqCtx, int nCtx, const hvb_tu_task *tasks, int n,
                       hvb_tu_result *out, int bitDepth, int grid)
{
    emuLaunch(std::min(n, grid), 32, [&] { tuFusedKernel<Sample>(planes, pool, rdoqCtx, tasks, n, out, bitDepth, nCtx, (unsigned)poolCount); });
}
extern "C" void emu_tu_chain_fused(const HvbPlane *planes, int16_t *pool, int poolCount, const hvb_rdoq_ctx *rdoqCtx, int nCtx, const hvb_tu_task *tasks,
                                   int n, hvb_tu_result *out, int bitDepth, int bps, int grid)
{
    if (bps == 1) fusedChain<uint8_t>(planes, pool, poolCount, rdoqCtx, nCtx, tasks, n, out, bitDepth, grid);
    else fusedChain<uint16_t>(planes, pool, poolCount, rdoqCtx, nCtx, tasks, n, out, bitDepth, grid);
}
extern "C" void emu_tu_chain(const HvbPlane *planes, int16_t *pool, int poolCount, const hvb_rdoq_ctx *rdoqCtx, int nCtx, const hvb_tu_task *tasks,
                             int n, hvb_tu_result *out, int bitDepth, int bps, int grid)
{
    if (bps == 1) chain<uint8_t>(planes, pool, poolCount, rdoqCtx, nCtx, tasks, n, out, bitDepth, grid);
    else chain<uint16_t>(planes, pool, poolCount, rdoqCtx, nCtx, tasks, n, out, bitDepth, grid);
}
'''),
    "hvb_me_small.cu": dict(entry=r'''
extern "C" void emu_me_small(const HvbPlane *planes, const hvb_me_task *tasks, int n, hvb_me_result *out, int bps, int grid)
{
    int cursor = 0;
    ''' + _both("emuLaunch(grid, kWarps * 32, [&] { meSearchSmallKernel<Sample>(planes, tasks, n, out, &cursor); });") + r'''
}
'''),
    "hvb_me.cu": dict(entry=r'''
extern "C" void emu_me_large(const HvbPlane *planes, const hvb_me_task *tasks, int n, hvb_me_result *out, int bitDepth, int bps, int grid)
{ ''' + _both("emuLaunch(grid, kWarps * 32, [&] { meSearchKernel<Sample>(planes, tasks, n, out, bitDepth); });") + r''' }
extern "C" void emu_me_bi(const HvbPlane *planes, const hvb_me_bi_task *tasks, int n, hvb_me_bi_result *out, int bitDepth, int bps, int grid)
{ ''' + _both("emuLaunch(grid, kWarps * 32, [&] { meBiSearchKernel<Sample>(planes, tasks, n, out, bitDepth); });") + r''' }
'''),
    "hvb_me_subpel.cu": dict(entry=r'''
extern "C" void emu_me_subpel(const HvbPlane *planes, const hvb_me_task *tasks, int n, hvb_me_result *out, int bitDepth, int bps, int grid)
{ ''' + _both("emuLaunch(grid, kWarps * 32, [&] { meSubpelKernel<Sample, false>(planes, tasks, n, out, bitDepth); });") + r''' }
extern "C" void emu_me_bi_subpel(const HvbPlane *planes, const hvb_me_bi_task *tasks, int n, hvb_me_bi_result *out, int bitDepth, int bps, int grid)
{ ''' + _both("emuLaunch(grid, kWarps * 32, [&] { meSubpelKernel<Sample, true>(planes, tasks, n, out, bitDepth); });") + r''' }
'''),
    "hvb_pu_cost.cu": dict(
        strip=("template <typename Sample>\nint launch(",),
        entry=r'''
extern "C" void emu_pu_cost(const HvbPlane *planes, const hvb_pu_cost_task *tasks, int n, int32_t *out, int bitDepth, int bps, int grid)
{
    int cursor = 0;
    ''' + _both("emuLaunch(grid, 256, [&] { puCostKernel<Sample>(planes, tasks, n, out, bitDepth, &cursor); });") + r'''
}
'''),
}

ENTRIES["hvb_loopfilter.cu"] = dict(
    use_unit_header=False, loop_info=True,
    entry=r'''
extern "C" void emu_deblock(const HvbPlane *planes, const HvbLoopInfo *info, const hvb_deblock_task *tasks, int n, int bitDepth, int bps, int grid)
{ ''' + _both("emuLaunch(grid, 256, [&] { deblockKernel<Sample>(planes, info, tasks, n, bitDepth); });") + r''' }
extern "C" void emu_sao(const HvbPlane *planes, const HvbLoopInfo *info, const hvb_sao_task *tasks, int n, int bitDepth, int bps, int grid)
{ ''' + _both("emuLaunch(grid, 256, [&] { saoKernel<Sample>(planes, info, tasks, n, bitDepth); });") + r''' }
extern "C" void emu_sao_stats(const HvbPlane *planes, const hvb_sao_stats_task *tasks, int n, hvb_sao_stats *out, int bitDepth, int bps, int grid)
{ ''' + _both("emuLaunch(std::min(n, grid), 256, [&] { saoStatsKernel<Sample>(planes, tasks, n, out, bitDepth); });") + r''' }
''')
ENTRIES["hvb_codeddata.cu"] = dict(
    use_unit_header=False,
    entry=r'''
extern "C" void emu_coded_residual(int16_t *pool, const hvb_coded_residual_task *tasks, int n, int recordsBase, int capacityWords,
                                   hvb_coded_residual *out, int grid)
{
    int cursor = 0;
    emuLaunch(grid, 128, [&] { codedResidualKernel(pool, reinterpret_cast<uint16_t *>(pool), tasks, n, recordsBase, capacityWords, out, &cursor); });
    emuLaunch(1, 32, [&] { codedResidualTotalKernel(out, n, recordsBase, capacityWords, &cursor); });
}
''')
ENTRIES["hvb_preanalysis.cu"] = dict(
    use_unit_header=False,
    entry=r'''
extern "C" void emu_intra_complexity(const HvbPlane *planes, const hvb_intra_complexity_task *tasks, int n, int32_t *out, int bps, int grid)
{ ''' + _both("emuLaunch(grid, 128, [&] { intraComplexityKernel<Sample>(planes, tasks, n, out); });") + r''' }
''')


class LoopInfo(C.Structure):
    _fields_ = [("blocks", C.c_void_p), ("ctus", C.c_void_p), ("sao", C.c_void_p), ("blockStride", C.c_int32), ("blockRows", C.c_int32),
                ("widthInCtbs", C.c_int32), ("ctbLog2", C.c_int32), ("saoCount", C.c_int32), ("reserved", C.c_int32)]


_LIBS: dict = {}
_TMP = None


def kernels_of(cu_file: str) -> C.CDLL:
    """the emulation library of one csrc/*.cu file, built once per process"""
    global _TMP
    if cu_file not in _LIBS:
        if _TMP is None:
            _TMP = tempfile.TemporaryDirectory(prefix="hvb_emu_")
        d = Path(_TMP.name) / cu_file.replace(".", "_")
        d.mkdir()
        spec = dict(ENTRIES[cu_file])
        if spec.pop("loop_info", False):  # the kernels of this file also see HvbLoopInfo (csrc/hvb_internal.cuh)
            internal = (host_emu_warp.CSRC / "hvb_internal.cuh").read_text()
            start = internal.index("struct HvbLoopInfo\n{")
            spec["prelude_structs"] = internal[start:internal.index("};", start) + 2]
        _LIBS[cu_file] = host_emu_warp.build(d, cu_file, spec.pop("entry"), check_alignment=True, **spec)
        _LIBS[cu_file].emu_set_schedule(*_SCHEDULE)
    return _LIBS[cu_file]


def set_schedule(mode: int, seed: int = 1):
    """fiber visiting order of every emulation library built so far and from now on: 0 ascending, 1 descending, 2 random"""
    global _SCHEDULE
    _SCHEDULE = (mode, seed)
    for lib in _LIBS.values():
        lib.emu_set_schedule(mode, seed)


_SCHEDULE = (0, 1)


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


class EmuContext:
    """the subset of hvb.Context the parity tests use, on emulated kernels (host task arrays only)"""

    def __init__(self, device: int = 0, bytes_per_sample: int = 1, bit_depth: int = 8):
        self.bps, self.bit_depth = bytes_per_sample, bit_depth
        self.sample_dtype = np.uint8 if bytes_per_sample == 1 else np.uint16
        self.pictures: list = []  # per picture: three plane records (see picture_create)
        self.sample_pool = np.zeros(0, self.sample_dtype)
        self.coeff_pool = np.zeros(0, np.int16)
        self.rdoq_ctx = np.zeros(0, hvb.rdoq_ctx_t)
        self.loop_info: dict = {}
        self.launch_count = 0

    # -- pictures ----------------------------------------------------------------------------
    # laid out as hvb_picture_create does (csrc/hvb_context.cu): row pitch a multiple of 256 bytes, sample (0,0) of every
    # row 256-byte aligned, `pad` valid samples on every side plus slack -- the kernels' vector accesses rely on it, and the
    # emulation libraries are built with -fsanitize=alignment so that a misaligned one aborts here as it would fault there
    def picture_create(self, width: int, height: int, pad: int = 96) -> int:
        planes = []
        for c in range(3):
            w, h, pd = (width, height, pad) if c == 0 else (width // 2, height // 2, pad // 2)
            unit = 256 // self.bps
            pad_left = -(-pd // unit) * unit
            stride = -(-(pad_left + w + pd + 16) // unit) * unit
            rows = h + 2 * pd + 2
            raw = np.zeros(stride * rows * self.bps + 256, np.uint8)
            off = (-raw.ctypes.data) % 256
            full = raw[off:off + stride * rows * self.bps].view(self.sample_dtype).reshape(rows, stride)
            planes.append(dict(full=full, ox=pad_left, oy=pd, w=w, h=h, pad=pd, keep=raw))
        self.pictures.append(planes)
        return len(self.pictures) - 1

    def _visible(self, pic, c):
        p = self.pictures[pic][c]
        return p["full"][p["oy"]:p["oy"] + p["h"], p["ox"]:p["ox"] + p["w"]]

    def picture_upload(self, pic: int, c_idx: int, plane: np.ndarray, y0: int = 0, rows=None):
        rows = plane.shape[0] - y0 if rows is None else rows
        self._visible(pic, c_idx)[y0:y0 + rows, :plane.shape[1]] = plane[:rows]

    def picture_download(self, pic: int, c_idx: int, width: int, height: int) -> np.ndarray:
        return self._visible(pic, c_idx)[:height, :width].copy()

    def picture_pad(self, pic: int):
        for c in range(3):
            p = self.pictures[pic][c]
            if p["pad"]:
                full = p["full"]
                full[...] = np.pad(self._visible(pic, c), ((p["oy"], full.shape[0] - p["oy"] - p["h"]), (p["ox"], full.shape[1] - p["ox"] - p["w"])),
                                   mode="edge")

    def padded(self, pic: int, c_idx: int) -> np.ndarray:
        """a plain copy of the plane with exactly `pad` samples on every side (what np.pad(visible, pad, 'edge') gives after picture_pad)"""
        p = self.pictures[pic][c_idx]
        return np.ascontiguousarray(p["full"][p["oy"] - p["pad"]:p["oy"] + p["h"] + p["pad"], p["ox"] - p["pad"]:p["ox"] + p["w"] + p["pad"]])

    def upload_yuv(self, pic: int, y, u, v, pad: bool = True):
        for c, p in enumerate((y, u, v)):
            self.picture_upload(pic, c, p)
        if pad:
            self.picture_pad(pic)

    def _planes(self):
        table = (Plane * (3 * max(len(self.pictures), 1)))()
        for i, pic in enumerate(self.pictures):
            for c, p in enumerate(pic):
                full = p["full"]
                table[3 * i + c] = Plane(full.ctypes.data + (p["oy"] * full.shape[1] + p["ox"]) * full.itemsize, full.shape[1], p["w"], p["h"],
                                         p["pad"], 0)
        return table

    # -- pools -------------------------------------------------------------------------------
    @staticmethod
    def _aligned(count, dtype):
        """`count` zeroed elements starting on a 256-byte boundary, as cudaMalloc returns them"""
        itemsize = np.dtype(dtype).itemsize
        raw = np.zeros(count * itemsize + 256, np.uint8)
        off = (-raw.ctypes.data) % 256
        return raw[off:off + count * itemsize].view(dtype)

    @classmethod
    def _store(cls, pool, data, offset, dtype):
        data = np.ascontiguousarray(data, dtype=dtype).reshape(-1)
        if pool.size < offset + data.size:
            grown = cls._aligned(offset + data.size, dtype)
            grown[:pool.size] = pool
            pool = grown
        pool[offset:offset + data.size] = data
        return pool

    def pool_upload(self, samples, offset: int = 0):
        self.sample_pool = self._store(self.sample_pool, samples, offset, self.sample_dtype)

    def coeff_upload(self, data, offset: int = 0):
        self.coeff_pool = self._store(self.coeff_pool, data, offset, np.int16)

    def coeff_download(self, count: int, offset: int = 0) -> np.ndarray:
        return self.coeff_pool[offset:offset + count].copy()

    def rdoq_contexts_upload(self, snapshots, first: int = 0):
        snapshots = np.ascontiguousarray(snapshots, dtype=hvb.rdoq_ctx_t).reshape(-1)
        if self.rdoq_ctx.size < first + snapshots.size:
            grown = self._aligned(first + snapshots.size, hvb.rdoq_ctx_t)
            grown[:self.rdoq_ctx.size] = self.rdoq_ctx
            self.rdoq_ctx = grown
        self.rdoq_ctx[first:first + snapshots.size] = snapshots

    # -- batched calls -----------------------------------------------------------------------
    def _tasks(self, tasks, dtype):
        tasks = np.ascontiguousarray(tasks, dtype=dtype).reshape(-1)
        staged = self._aligned(max(tasks.size, 1), dtype)  # the staging buffer of hvbStageIn is 256-byte aligned
        staged[:tasks.size] = tasks
        return staged[:tasks.size]

    def _metric(self, fn, tasks, dtype, out_dtype, per_task):
        t = self._tasks(tasks, dtype)
        out = np.zeros(t.size * per_task, out_dtype)
        getattr(kernels_of("hvb_metrics.cu"), fn)(self._planes(), _ptr(t), t.size, _ptr(out), self.bps, GRID)
        return out if per_task == 1 else out.reshape(t.size, per_task)

    def sad(self, tasks, **_):
        return self._metric("emu_sad", tasks, hvb.metric_task_t, np.int32, 1)

    def ssd(self, tasks, **_):
        return self._metric("emu_ssd", tasks, hvb.metric_task_t, np.uint32, 1)

    def satd(self, tasks, **_):
        return self._metric("emu_satd", tasks, hvb.metric_task_t, np.int32, 1)

    def sad4(self, tasks, **_):
        return self._metric("emu_sad4", tasks, hvb.sad4_task_t, np.int32, 4)

    def pred(self, tasks, **_):
        t = self._tasks(tasks, hvb.pred_task_t)
        kernels_of("hvb_pred.cu").emu_pred(self._planes(), _ptr(t), t.size, self.bit_depth, self.bps, GRID)

    def subtract_bi(self, tasks, **_):
        t = self._tasks(tasks, hvb.subtract_bi_task_t)
        kernels_of("hvb_pred.cu").emu_subtract_bi(self._planes(), _ptr(t), t.size, self.bit_depth, self.bps, GRID)

    def interp_satd(self, tasks, **_):
        t = self._tasks(tasks, hvb.interp_satd_task_t)
        out = np.zeros(t.size, np.int32)
        kernels_of("hvb_pred.cu").emu_interp_satd(self._planes(), _ptr(t), t.size, _ptr(out), self.bit_depth, self.bps, GRID)
        return out

    def intra_pred(self, tasks, **_):
        t = self._tasks(tasks, hvb.intra_task_t)
        kernels_of("hvb_intra.cu").emu_intra_pred(self._planes(), _ptr(self.sample_pool), _ptr(t), t.size, self.bit_depth, self.bps, GRID)

    def intra_satd35(self, tasks, **_):
        t = self._tasks(tasks, hvb.intra_sweep_task_t)
        out = np.zeros((t.size, 35), np.int32)
        kernels_of("hvb_intra.cu").emu_intra_sweep(self._planes(), _ptr(self.sample_pool), _ptr(t), t.size, _ptr(out), self.bit_depth, self.bps, GRID)
        return out

    def transform_fwd(self, tasks, **_):
        t = self._tasks(tasks, hvb.transform_task_t)
        kernels_of("hvb_tu.cu").emu_transform(_ptr(self.coeff_pool), _ptr(t), t.size, self.bit_depth, 0, GRID)

    def transform_inv(self, tasks, **_):
        t = self._tasks(tasks, hvb.transform_task_t)
        kernels_of("hvb_tu.cu").emu_transform(_ptr(self.coeff_pool), _ptr(t), t.size, self.bit_depth, 1, GRID)

    def quantize(self, tasks, **_):
        t = self._tasks(tasks, hvb.quant_task_t)
        cbf = np.zeros(t.size, np.int32)
        kernels_of("hvb_tu.cu").emu_quant(_ptr(self.coeff_pool), _ptr(t), t.size, _ptr(cbf), 0, GRID)
        return cbf

    def quantize_inverse(self, tasks, **_):
        t = self._tasks(tasks, hvb.quant_task_t)
        kernels_of("hvb_tu.cu").emu_quant(_ptr(self.coeff_pool), _ptr(t), t.size, None, 1, GRID)

    def inverse_transform_add(self, tasks, **_):
        t = self._tasks(tasks, hvb.ita_task_t)
        kernels_of("hvb_tu.cu").emu_ita(self._planes(), _ptr(self.coeff_pool), _ptr(t), t.size, self.bit_depth, self.bps, GRID)

    def rdoq(self, tasks, **_):
        t = self._tasks(tasks, hvb.rdoq_task_t)
        cbf = np.zeros(t.size, np.int32)
        kernels_of("hvb_tu.cu").emu_rdoq(_ptr(self.coeff_pool), self.coeff_pool.size, _ptr(self.rdoq_ctx), self.rdoq_ctx.size, _ptr(t), t.size,
                                         _ptr(cbf), self.bit_depth, GRID)
        return cbf

    def set_tu_fused_max(self, blocks: int):
        self.tu_fused_max = int(blocks)

    def tu_chain(self, tasks, **_):
        t = self._tasks(tasks, hvb.tu_task_t)
        out = np.zeros(t.size, hvb.tu_result_t)
        lib = kernels_of("hvb_tu.cu")
        entry = lib.emu_tu_chain_fused if t.size <= getattr(self, "tu_fused_max", 256) else lib.emu_tu_chain  # as hvb_tu_chain_batch chooses
        entry(self._planes(), _ptr(self.coeff_pool), self.coeff_pool.size, _ptr(self.rdoq_ctx), self.rdoq_ctx.size,
              _ptr(t), t.size, _ptr(out), self.bit_depth, self.bps, GRID)
        return out

    def me_search(self, tasks, **_):
        t = self._tasks(tasks, hvb.me_task_t)
        out = np.zeros(t.size, hvb.me_result_t)
        planes = self._planes()
        kernels_of("hvb_me_small.cu").emu_me_small(planes, _ptr(t), t.size, _ptr(out), self.bps, GRID)
        kernels_of("hvb_me.cu").emu_me_large(planes, _ptr(t), t.size, _ptr(out), self.bit_depth, self.bps, GRID)
        kernels_of("hvb_me_subpel.cu").emu_me_subpel(planes, _ptr(t), t.size, _ptr(out), self.bit_depth, self.bps, GRID)
        return out

    def me_bi_search(self, tasks, **_):
        t = self._tasks(tasks, hvb.me_bi_task_t)
        out = np.zeros(t.size, hvb.me_bi_result_t)
        planes = self._planes()
        kernels_of("hvb_me.cu").emu_me_bi(planes, _ptr(t), t.size, _ptr(out), self.bit_depth, self.bps, GRID)
        kernels_of("hvb_me_subpel.cu").emu_me_bi_subpel(planes, _ptr(t), t.size, _ptr(out), self.bit_depth, self.bps, GRID)
        return out

    def pu_cost(self, tasks, **_):
        t = self._tasks(tasks, hvb.pu_cost_task_t)
        out = np.zeros((t.size, 3), np.int32)
        kernels_of("hvb_pu_cost.cu").emu_pu_cost(self._planes(), _ptr(t), t.size, _ptr(out), self.bit_depth, self.bps, GRID)
        return out

    # -- in-loop filters, coded data, pre-analysis ---------------------------------
    def deblock_info_upload(self, pic, blocks, ctus, pic_width_in_ctbs, pic_height_in_ctbs, ctb_log2):
        blocks = np.ascontiguousarray(blocks, dtype=hvb.deblock_block_t)
        b = self._aligned(blocks.size, hvb.deblock_block_t)
        b[:] = blocks.reshape(-1)
        ctus = np.ascontiguousarray(ctus, dtype=hvb.deblock_ctu_t).reshape(-1)
        c = self._aligned(ctus.size, hvb.deblock_ctu_t)
        c[:] = ctus
        info = self.loop_info.setdefault(pic, {})
        info.update(blocks=b, ctus=c, blockStride=blocks.shape[1], blockRows=blocks.shape[0], widthInCtbs=pic_width_in_ctbs, ctbLog2=ctb_log2)

    def sao_info_upload(self, pic, ctus):
        assert pic in self.loop_info and "blocks" in self.loop_info[pic], "hvb_sao_info_upload before hvb_deblock_info_upload"
        ctus = np.ascontiguousarray(ctus, dtype=hvb.sao_ctu_t).reshape(-1)
        rec = self._aligned(ctus.size, hvb.sao_ctu_t)
        rec[:] = ctus
        self.loop_info[pic].update(sao=rec)

    def _loop_table(self):
        table = (LoopInfo * max(len(self.pictures), 1))()
        for pic, info in self.loop_info.items():
            sao = info.get("sao")
            table[pic] = LoopInfo(info["blocks"].ctypes.data, info["ctus"].ctypes.data, sao.ctypes.data if sao is not None else None,
                                  info["blockStride"], info["blockRows"], info["widthInCtbs"], info["ctbLog2"], sao.size if sao is not None else 0, 0)
        return table

    def deblock(self, tasks, **_):
        t = self._tasks(tasks, hvb.deblock_task_t)
        kernels_of("hvb_loopfilter.cu").emu_deblock(self._planes(), self._loop_table(), _ptr(t), t.size, self.bit_depth, self.bps, GRID)

    def sao(self, tasks, **_):
        t = self._tasks(tasks, hvb.sao_task_t)
        kernels_of("hvb_loopfilter.cu").emu_sao(self._planes(), self._loop_table(), _ptr(t), t.size, self.bit_depth, self.bps, GRID)

    def sao_stats(self, tasks, **_):
        t = self._tasks(tasks, hvb.sao_stats_task_t)
        out = np.zeros(t.size, hvb.sao_stats_t)
        kernels_of("hvb_loopfilter.cu").emu_sao_stats(self._planes(), _ptr(t), t.size, _ptr(out), self.bit_depth, self.bps, GRID)
        return out

    def picture_copy(self, dst_pic, src_pic):
        for c in range(3):
            self.pictures[dst_pic][c]["full"][...] = self.pictures[src_pic][c]["full"]

    def coded_residual(self, tasks, records_base, capacity_words):
        t = self._tasks(tasks, hvb.coded_residual_task_t)
        if self.coeff_pool.size < records_base + capacity_words:
            grown = self._aligned(records_base + capacity_words, np.int16)
            grown[:self.coeff_pool.size] = self.coeff_pool
            self.coeff_pool = grown
        out = np.zeros(t.size + 1, hvb.coded_residual_t)
        kernels_of("hvb_codeddata.cu").emu_coded_residual(_ptr(self.coeff_pool), _ptr(t), t.size, records_base, capacity_words, _ptr(out), GRID)
        return out

    def intra_complexity(self, tasks, out_count):
        t = self._tasks(tasks, hvb.intra_complexity_task_t)
        out = np.zeros(out_count, np.int32)
        kernels_of("hvb_preanalysis.cu").emu_intra_complexity(self._planes(), _ptr(t), t.size, _ptr(out), self.bps, GRID)
        return out

    def sync(self):
        pass

    def close(self):
        pass


class EmuScene:
    """gpu_common.Scene on an EmuContext: three synthetic pictures and two scratch pictures"""

    def __init__(self, bps: int, bit_depth: int, width: int = 256, height: int = 192, pad: int = 96):
        self.bps, self.bd = bps, bit_depth
        self.dtype = np.uint8 if bps == 1 else np.uint16
        self.w, self.h, self.pad = width, height, pad
        self.ctx = EmuContext(0, bps, bit_depth)
        self.pics, self.host = [], []
        for i in range(3):
            f = [p.astype(self.dtype) for p in synth.frame(i, width, height, bit_depth)]
            if bps == 2 and i == 1:
                f[0][::7, ::5] = (1 << bit_depth) - 1
            pic = self.ctx.picture_create(width, height, pad)
            self.ctx.upload_yuv(pic, *f)
            self.pics.append(pic)
            self.host.append([self.ctx.padded(pic, c) for c in range(3)])
        self.scratch = [self.ctx.picture_create(width, height, pad) for _ in range(2)]

    def view(self, pic_index: int, c_idx: int, x: int, y: int):
        pad = self.pad if c_idx == 0 else self.pad // 2
        a = self.host[pic_index][c_idx]
        return a, (int(y) + pad) * a.shape[1] + (int(x) + pad), a.shape[1]

    def download(self, pic: int, c_idx: int) -> np.ndarray:
        w, h = (self.w, self.h) if c_idx == 0 else (self.w // 2, self.h // 2)
        return self.ctx.picture_download(pic, c_idx, w, h)

    def close(self):
        self.ctx.close()
