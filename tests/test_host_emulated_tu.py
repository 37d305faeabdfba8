"""The TU pipeline's own source (csrc/hvb_tu.cu + csrc/hvb_rdoq.cuh: residual, tensor-core DCT, RDOQ pre-pass and the
bucket ordering, the serial RDOQ walk with sign-data hiding, dequantisation, IDCT, reconstruction, SSD), executed on the CPU
by the warp-level emulator (tests/host_emu_warp.py) and chained as hvb_tu_chain_batch chains it -- context bit tables, front,
order, RDOQ, back -- against the oracle: levels, reconstruction, SSDs and cbf of every block, all transform sizes and the
DST, plain quantisation and RDOQ, 8 and 10 bit.  The comparison of tests/test_gpu_tu.py::test_tu_chain in the CPU-only suite."""
import ctypes as C

import numpy as np
import pytest

import host_emu_warp
import orc
import test_gpu_tu as gpu_tu
from gpu_common import H, PAD, W
from test_host_emulated_loopfilter import Plane
from test_host_emulated_me import host_scene
from turingcodec_b200 import hvb

ENTRY = r'''
template <typename Sample>
static void chain(const HvbPlane *planes, int16_t *pool, int poolCount, const hvb_rdoq_ctx *rdoqCtx, int nCtx, const hvb_tu_task *tasks, int n,
                  hvb_tu_result *out, int bitDepth)
{
    // what initRdoqTables / hvbLaunchRdoqBits / chainScratch set up on the device
    for (int log2 = 2; log2 <= 5; ++log2)
        for (int scanIdx = 0; scanIdx < 3; ++scanIdx)
            for (int sp = 0; sp < (1 << (2 * log2)); ++sp)
                hvb_rdoq::gScanTable[((log2 - 2) * 3 + scanIdx) * 1024 + sp] = (short)hvb_rdoq::scanToRaster(log2, scanIdx, sp);
    const int bytes = nCtx * (int)sizeof(hvb_rdoq_ctx), entries = nCtx * hvb_rdoq::kLastTabPerCtx;
    std::vector<int2> bits(bytes);
    std::vector<int> last(entries);
    emuLaunch((bytes + 255) / 256, 256, [&] { rdoqBitsKernel(rdoqCtx, bits.data(), bytes); });
    emuLaunch((entries + 255) / 256, 256, [&] { rdoqLastKernel(rdoqCtx, last.data(), nCtx); });
    std::vector<int16_t> coefTmp(poolCount);
    std::vector<HvbCoefRec> recs(poolCount);
    std::vector<HvbRdoqMid> mids(n);
    std::vector<int> buckets(64 + n, 0);
    int *order = buckets.data() + 64;
    const int chunks = (n + 1023) / 1024;
    std::vector<int> compact(n), chunkBase(chunks + 1);
    emuLaunch(chunks, 1024, [&] { scanLocalKernel(TuCount{tasks}, n, compact.data(), chunkBase.data()); });
    emuLaunch(1, 1024, [&] { scanBlocksKernel(chunkBase.data(), chunks); });
    const int gridW = std::min((n + kWarps - 1) / kWarps, 3), gridT = std::min((n + 127) / 128, 2);
    emuLaunch(gridW, kWarps * 32, [&] { tuFrontKernel<Sample>(planes, pool, coefTmp.data(), rdoqCtx, tasks, n, out, mids.data(), buckets.data(), bitDepth, compact.data(), chunkBase.data(), nCtx, (unsigned)poolCount); });
    emuLaunch((n + 255) / 256, 256, [&] { tuOrderKernel(mids.data(), n, buckets.data(), buckets.data() + kRdoqBuckets, order); });
    emuLaunch(gridT, 128, [&] { tuRdoqKernel(pool, coefTmp.data(), recs.data(), rdoqCtx, tasks, n, out, mids.data(), buckets.data(), order, bitDepth,
                                             bits.data(), last.data(), compact.data(), chunkBase.data()); });
    emuLaunch(gridW, kWarps * 32, [&] { tuBackKernel<Sample>(planes, pool, tasks, n, out, bitDepth); });
}
extern "C" void emu_tu_chain(const HvbPlane *planes, int16_t *pool, int poolCount, const hvb_rdoq_ctx *rdoqCtx, int nCtx, const hvb_tu_task *tasks,
                             int n, hvb_tu_result *out, int bitDepth, int bps)
{
    if (bps == 1) chain<uint8_t>(planes, pool, poolCount, rdoqCtx, nCtx, tasks, n, out, bitDepth);
    else chain<uint16_t>(planes, pool, poolCount, rdoqCtx, nCtx, tasks, n, out, bitDepth);
}
'''


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    return host_emu_warp.build(tmp_path_factory.mktemp("emu_tu"), "hvb_tu.cu", ENTRY, use_unit_header=False,
                               mma_wrappers={"immaS8U8": False, "immaS8S8": True}, extra_headers=("hvb_rdoq.cuh",))


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
@pytest.mark.parametrize("use_rdoq", [False, True])
def test_tu_chain_kernels_on_cpu_match_oracle(emu, oracle, bps, bit_depth, use_rdoq):
    scene = host_scene(bps, bit_depth)
    dtype = np.uint8 if bps == 1 else np.uint16
    rec_pic = [np.zeros_like(a) for a in scene.host[0]]
    pictures = scene.host + [rec_pic]
    table = (Plane * (3 * len(pictures)))()
    for i, pic in enumerate(pictures):
        for c, a in enumerate(pic):
            pad = PAD if c == 0 else PAD // 2
            table[3 * i + c] = Plane(a.ctypes.data + (pad * a.shape[1] + pad) * a.itemsize, a.shape[1], a.shape[1] - 2 * pad,
                                     a.shape[0] - 2 * pad, pad, 0)
    rng = np.random.default_rng(44 + use_rdoq)
    cells = [(cx, cy) for cy in range(0, H - 31, 32) for cx in range(0, W - 31, 32)][:40]
    qps = [18, 22, 27, 32, 37, 42]
    ctxs = np.stack([orc.random_rdoq_ctx(rng, 0.57 * 2 ** ((qp - 12) / 3.0)) for qp in qps])
    snapshots = np.ascontiguousarray(ctxs.view(hvb.rdoq_ctx_t).reshape(-1))
    t = np.zeros(len(cells), hvb.tu_task_t)
    want, offset = [], 0
    fake = type("S", (), {"bd": bit_depth, "dtype": dtype})()
    for i, (cx, cy) in enumerate(cells):
        tr, log2n = gpu_tu.SHAPES[i % len(gpu_tu.SHAPES)]
        n = 1 << log2n
        k = i % len(qps)
        is_intra, sdh = int(i % 2), int((i // 2) % 2)
        scan_idx = int(rng.integers(0, 3)) if log2n <= 3 else 0
        px, py = min(max(cx + int(rng.integers(-3, 4)), 0), W - n), min(max(cy + int(rng.integers(-3, 4)), 0), H - n)
        for name, pic, x, y in (("src", 0, cx, cy), ("pred", 1, px, py), ("rec", 3, cx, cy)):
            t[i][name]["pic"], t[i][name]["cIdx"], t[i][name]["x"], t[i][name]["y"] = pic, 0, x, y
        src = scene.host[0][0][PAD + cy:PAD + cy + n, PAD + cx:PAD + cx + n]
        pred = scene.host[1][0][PAD + py:PAD + py + n, PAD + px:PAD + px + n]
        levels, rec, ssd, ssd_pred, cbf, q = gpu_tu.oracle_tu_chain(oracle, fake, src, pred, tr, log2n, qps[k], 0, use_rdoq, ctxs[k], scan_idx,
                                                                    is_intra, sdh)
        t[i]["levels"], t[i]["log2n"], t[i]["trType"], t[i]["cIdx"] = offset, log2n, tr, 0
        t[i]["flags"] = (1 if use_rdoq else 0) | (is_intra << 1) | (sdh << 2)
        t[i]["qscale"], t[i]["qshift"], t[i]["qoffset"], t[i]["iqscale"], t[i]["iqshift"] = q
        t[i]["scanIdx"], t[i]["rdoq_ctx"] = scan_idx, k
        want.append((levels, rec, ssd, ssd_pred, cbf, offset, n, gpu_tu.sad_quadrants(src, pred)))
        offset += n * n
    pool = np.zeros(len(cells) * 1024, np.int16)
    out = np.zeros(len(cells), hvb.tu_result_t)
    emu.emu_tu_chain(table, C.c_void_p(pool.ctypes.data), pool.size, C.c_void_p(snapshots.ctypes.data), len(qps), C.c_void_p(t.ctypes.data),
                     len(cells), C.c_void_p(out.ctypes.data), bit_depth, bps)
    coded = 0
    got_rec = rec_pic[0][PAD:PAD + H, PAD:PAD + W]
    for i, (cx, cy) in enumerate(cells):
        levels, rec, ssd, ssd_pred, cbf, off, n, quad = want[i]
        assert out[i]["sadQuad"].tolist() == quad and int(out[i]["status"]) == 0, (i, n)
        assert np.array_equal(pool[off:off + n * n], levels), (i, n, use_rdoq)
        assert np.array_equal(got_rec[cy:cy + n, cx:cx + n], rec), (i, n)
        assert int(out[i]["ssd"]) == ssd and int(out[i]["ssdPred"]) == ssd_pred and int(out[i]["cbf"]) == cbf, (i, n)
        coded += cbf
    assert coded > len(cells) // 4
