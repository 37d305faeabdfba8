"""GPU parity: forward/inverse transforms, (de)quantisation, inverse-transform-add, RDOQ and the fused
TU pipeline, through the C-ABI vs the oracle -- bit-exact."""
import numpy as np
import pytest

import orc
from gpu_common import H, W, Scene, block
from turingcodec_b200 import hvb

pytestmark = pytest.mark.gpu

SHAPES = [(1, 2), (0, 2), (0, 3), (0, 4), (0, 5)]  # (trType, log2n)


@pytest.fixture(scope="module", params=[(1, 8), (2, 10)], ids=["u8", "u16-10bit"])
def scene(request):
    s = Scene(*request.param)
    yield s
    s.close()


def test_forward_and_inverse_transform(scene, oracle):
    rng = np.random.default_rng(41)
    bd = scene.bd
    blocks, offset = [], 0
    t = np.zeros(60, hvb.transform_task_t)
    pool = []
    for i in range(t.size):
        tr, log2n = SHAPES[i % len(SHAPES)]
        n = 1 << log2n
        stride = n + (i % 3) * 8
        kind = i % 4
        if kind == 0:
            res = rng.integers(-256, 256, (n, stride))
        elif kind == 1:
            res = rng.integers(-(1 << bd) + 1, 1 << bd, (n, stride))
        elif kind == 2:
            res = np.where(rng.integers(0, 2, (n, stride)) > 0, 32767, -32768)  # drives the int16 wrap of pass 1
        else:
            res = rng.integers(-32768, 32768, (n, stride))
        res = res.astype(np.int16)
        t[i]["src"], t[i]["dst"], t[i]["src_stride"] = offset, offset + n * stride, stride
        t[i]["log2n"], t[i]["trType"] = log2n, tr
        pool += [res.reshape(-1), np.zeros(n * n, np.int16)]
        blocks.append((res, offset + n * stride, stride))
        offset += n * stride + n * n
    scene.ctx.coeff_upload(np.concatenate(pool))
    scene.ctx.transform_fwd(t)
    got = scene.ctx.coeff_download(offset)
    coeff_sets = []
    for i, (res, dst, stride) in enumerate(blocks):
        tr, log2n = int(t[i]["trType"]), int(t[i]["log2n"])
        n = 1 << log2n
        want = np.zeros(n * n, np.int16)
        oracle.transform_fwd(want, res, stride, tr, log2n, bd)
        assert np.array_equal(got[dst:dst + n * n], want), (i, tr, log2n)
        coeff_sets.append(want)

    # inverse (residual only) of those coefficients, in place
    ti = np.zeros(t.size, hvb.transform_task_t)
    for i in range(t.size):
        ti[i]["src"] = ti[i]["dst"] = t[i]["dst"]
        ti[i]["log2n"], ti[i]["trType"] = t[i]["log2n"], t[i]["trType"]
    scene.ctx.transform_inv(ti)
    got = scene.ctx.coeff_download(offset)
    for i, c in enumerate(coeff_sets):
        tr, log2n = int(t[i]["trType"]), int(t[i]["log2n"])
        n = 1 << log2n
        want = np.zeros(n * n, np.int16)
        oracle.inverse_transform(want, c, tr, log2n, bd)
        assert np.array_equal(got[int(t[i]["dst"]):int(t[i]["dst"]) + n * n], want), (i, tr, log2n)


def test_quantize_and_inverse(scene, oracle):
    rng = np.random.default_rng(42)
    sizes = [16, 64, 256, 1024]
    params = [(51, 20, 14), (26214, 21, 171), (16384, 16, 21845), (32767, 27, 32767), (20560, 22, 10880)]
    t = np.zeros(40, hvb.quant_task_t)
    pool, offset, srcs = [], 0, []
    for i in range(t.size):
        n = sizes[i % 4]
        src = (rng.integers(0, 1 << 15, n) - rng.integers(0, 1 << 15, n)).astype(np.int16)
        if i % 9 == 0:
            src[:] = 0
        scale, shift, off = params[i % len(params)]
        t[i] = (offset, offset + n, n, scale, shift, off)
        pool += [src, np.zeros(n, np.int16)]
        srcs.append(src)
        offset += 2 * n
    scene.ctx.coeff_upload(np.concatenate(pool))
    cbf = scene.ctx.quantize(t)
    got = scene.ctx.coeff_download(offset)
    for i, src in enumerate(srcs):
        n = src.size
        want = np.zeros(n, np.int16)
        c = oracle.quantize(want, src, int(t[i]["scale"]), int(t[i]["shift"]), int(t[i]["offset"]))
        assert np.array_equal(got[int(t[i]["dst"]):int(t[i]["dst"]) + n], want)
        assert bool(cbf[i]) == (c != 0)
    # inverse: scale 51 / 52224 (quantize.cpp:250-275) and a realistic pair
    inv = [(51, 1), (51, 4), (52224, 3), (52224, 9), (816, 4)]
    for i in range(t.size):
        t[i]["scale"], t[i]["shift"] = inv[i % len(inv)]
    scene.ctx.coeff_upload(np.concatenate(pool))
    scene.ctx.quantize_inverse(t)
    got = scene.ctx.coeff_download(offset)
    for i, src in enumerate(srcs):
        want = np.zeros(src.size, np.int16)
        oracle.quantize_inverse(want, src, int(t[i]["scale"]), int(t[i]["shift"]))
        assert np.array_equal(got[int(t[i]["dst"]):int(t[i]["dst"]) + src.size], want)


def test_inverse_transform_add(scene, oracle):
    rng = np.random.default_rng(43)
    bd = scene.bd
    cells = [(cx, cy) for cy in range(0, H - 31, 32) for cx in range(0, W - 31, 32)]
    t = np.zeros(len(cells), hvb.ita_task_t)
    pool, offset, coeffs = [], 0, []
    for i, (cx, cy) in enumerate(cells):
        tr, log2n = SHAPES[i % len(SHAPES)]
        n = 1 << log2n
        c = rng.integers(-128, 128, n * n) if i % 3 else rng.integers(-32768, 32768, n * n)
        c = c.astype(np.int16)
        block(t[i:i + 1], "dst", scene.scratch[0], 0, cx, cy)
        block(t[i:i + 1], "pred", scene.pics[1], 0, int(rng.integers(0, W - n)), int(rng.integers(0, H - n)))
        t[i]["coeffs"], t[i]["log2n"], t[i]["trType"] = offset, log2n, tr
        pool.append(c)
        coeffs.append(c)
        offset += n * n
    scene.ctx.coeff_upload(np.concatenate(pool))
    scene.ctx.inverse_transform_add(t)
    got = scene.download(scene.scratch[0], 0)
    for i, (cx, cy) in enumerate(cells):
        tr, log2n = int(t[i]["trType"]), int(t[i]["log2n"])
        n = 1 << log2n
        p, op, sp = scene.view(1, 0, t[i]["pred"]["x"], t[i]["pred"]["y"])
        pred = np.ascontiguousarray(p.reshape(-1)[op:op + n * sp].reshape(n, sp)[:, :n])
        want = np.zeros((n, n), scene.dtype)
        oracle.inverse_transform_add(want, n, pred, n, coeffs[i], tr, log2n, bd)
        assert np.array_equal(got[cy:cy + n, cx:cx + n], want), (i, tr, log2n)


def oracle_tu_chain(oracle, scene, src, pred, tr, log2n, qp, c_idx, use_rdoq, ctx_bytes, scan_idx, is_intra, sdh):
    """Reconstruct.cpp:731-857 restated with oracle primitives; returns (levels, rec, ssd, ssdPred, cbf)"""
    bd = scene.bd
    n = 1 << log2n
    res = (src.astype(np.int32) - pred.astype(np.int32)).astype(np.int16)
    coeffs = np.zeros(n * n, np.int16)
    oracle.transform_fwd(coeffs, np.ascontiguousarray(res), n, tr, log2n, bd)
    qscale, qshift, iqscale, iqshift = orc.quant_params(qp, log2n, bd)
    qoffset = (171 if is_intra else 85) << 7  # (171|85) << (shift-9) >> (shift-16)
    levels = np.zeros(n * n, np.int16)
    if use_rdoq:
        cbf = orc.oracle_rdoq(oracle, levels, coeffs, ctx_bytes, qscale, qshift, iqscale, log2n, c_idx, scan_idx, is_intra, sdh, bd)
    else:
        cbf = oracle.quantize(levels, coeffs, qscale, qshift, qoffset)
    deq = np.zeros(n * n, np.int16)
    oracle.quantize_inverse(deq, levels, iqscale, iqshift)
    rec = np.zeros((n, n), scene.dtype)
    oracle.inverse_transform_add(rec, n, np.ascontiguousarray(pred), n, deq, tr, log2n, bd)
    ssd = oracle.ssd(np.ascontiguousarray(src), 0, n, rec, 0, n, n, n)
    ssd_pred = oracle.ssd(np.ascontiguousarray(src), 0, n, np.ascontiguousarray(pred), 0, n, n, n)
    return levels, rec, ssd, ssd_pred, int(cbf != 0), (qscale, qshift, qoffset, iqscale, iqshift)


def sad_quadrants(src, pred):
    """Candidate::sadResidueQuad of one block (Reconstruct.cpp:1268-1287): sum |src - pred| per quadrant, [2*yHalf + xHalf]"""
    d = np.abs(src.astype(np.int64) - pred.astype(np.int64))
    h = d.shape[0] // 2
    return [int(d[:h, :h].sum()), int(d[:h, h:].sum()), int(d[h:, :h].sum()), int(d[h:, h:].sum())]


@pytest.mark.parametrize("fused", [False, True], ids=["staged", "fused"])
@pytest.mark.parametrize("use_rdoq", [False, True])
def test_tu_chain(scene, oracle, use_rdoq, fused):
    """both forms of hvb_tu_chain_batch: the staged kernels (whole frames) and the one-launch form (a warp per block, small batches)"""
    scene.ctx.set_tu_fused_max(1 << 20 if fused else 0)
    rng = np.random.default_rng(44 + use_rdoq)
    cells = [(cx, cy) for cy in range(0, H - 31, 32) for cx in range(0, W - 31, 32)]
    n_ctx = 6
    ctxs = np.stack([orc.random_rdoq_ctx(rng, 0.57 * 2 ** ((qp - 12) / 3.0)) for qp in (18, 22, 27, 32, 37, 42)])
    scene.ctx.rdoq_contexts_upload(ctxs.view(hvb.rdoq_ctx_t).reshape(-1))
    qps = [18, 22, 27, 32, 37, 42]
    t = np.zeros(len(cells), hvb.tu_task_t)
    want = []
    offset = 0
    scene.ctx.coeff_upload(np.zeros(len(cells) * 1024, np.int16))
    for i, (cx, cy) in enumerate(cells):
        tr, log2n = SHAPES[i % len(SHAPES)]
        n = 1 << log2n
        k = i % n_ctx
        is_intra, sdh = int(i % 2), int((i // 2) % 2)
        scan_idx = int(rng.integers(0, 3)) if log2n <= 3 else 0
        # the prediction is the co-located block of the next frame, shifted a little: realistic residuals
        px, py = min(max(cx + int(rng.integers(-3, 4)), 0), W - n), min(max(cy + int(rng.integers(-3, 4)), 0), H - n)
        block(t[i:i + 1], "src", scene.pics[0], 0, cx, cy)
        block(t[i:i + 1], "pred", scene.pics[1], 0, px, py)
        block(t[i:i + 1], "rec", scene.scratch[1], 0, cx, cy)
        s, os_, ss = scene.view(0, 0, cx, cy)
        p, op, sp = scene.view(1, 0, px, py)
        src = s.reshape(-1)[os_:os_ + n * ss].reshape(n, ss)[:, :n]
        pred = p.reshape(-1)[op:op + n * sp].reshape(n, sp)[:, :n]
        levels, rec, ssd, ssd_pred, cbf, q = oracle_tu_chain(oracle, scene, src, pred, tr, log2n, qps[k], 0, use_rdoq,
                                                              ctxs[k], scan_idx, is_intra, sdh)
        t[i]["levels"], t[i]["log2n"], t[i]["trType"], t[i]["cIdx"] = offset, log2n, tr, 0
        t[i]["flags"] = (1 if use_rdoq else 0) | (is_intra << 1) | (sdh << 2)
        t[i]["qscale"], t[i]["qshift"], t[i]["qoffset"], t[i]["iqscale"], t[i]["iqshift"] = q
        t[i]["scanIdx"], t[i]["rdoq_ctx"] = scan_idx, k
        want.append((levels, rec, ssd, ssd_pred, cbf, offset, n, sad_quadrants(src, pred)))
        offset += n * n
    out = scene.ctx.tu_chain(t)
    got_levels = scene.ctx.coeff_download(offset)
    got_rec = scene.download(scene.scratch[1], 0)
    coded = 0
    for i, (cx, cy) in enumerate(cells):
        levels, rec, ssd, ssd_pred, cbf, off, n, quad = want[i]
        assert np.array_equal(got_levels[off:off + n * n], levels), (i, n, use_rdoq)
        assert np.array_equal(got_rec[cy:cy + n, cx:cx + n], rec), (i, n)
        assert int(out[i]["ssd"]) == ssd and int(out[i]["ssdPred"]) == ssd_pred and int(out[i]["cbf"]) == cbf, (i, n)
        assert out[i]["sadQuad"].tolist() == quad and int(out[i]["status"]) == 0, (i, n)
        coded += cbf
    assert coded > len(cells) // 4


def test_rdoq_standalone(scene, oracle):
    rng = np.random.default_rng(46)
    bd = scene.bd
    n_tasks = 120
    ctxs = np.stack([orc.random_rdoq_ctx(rng, 0.57 * 2 ** ((qp - 12) / 3.0)) for qp in (16, 24, 30, 36)])
    scene.ctx.rdoq_contexts_upload(ctxs.view(hvb.rdoq_ctx_t).reshape(-1))
    qps = [16, 24, 30, 36]
    t = np.zeros(n_tasks, hvb.rdoq_task_t)
    pool, offset, cases = [], 0, []
    for i in range(n_tasks):
        log2n = 2 + i % 4
        n = 1 << log2n
        c_idx = int(rng.integers(0, 3)) if log2n < 5 else 0
        scan_idx = int(rng.integers(0, 3)) if log2n <= 3 else 0
        k = i % 4
        amp = float(rng.choice([2, 6, 20, 60])) * (1 << (bd - 8))
        res = np.clip(rng.normal(0, amp, (n, n)), -(1 << bd) + 1, (1 << bd) - 1).astype(np.int16)
        coeffs = np.zeros(n * n, np.int16)
        oracle.transform_fwd(coeffs, res, n, 0, log2n, bd)
        qscale, qshift, iqscale, _ = orc.quant_params(qps[k], log2n, bd)
        is_intra, sdh = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        t[i] = (offset, offset + n * n, qscale, qshift, iqscale, log2n, c_idx, scan_idx, (is_intra << 1) | (sdh << 2), k)
        pool += [coeffs, np.zeros(n * n, np.int16)]
        cases.append((coeffs, ctxs[k], qscale, qshift, iqscale, log2n, c_idx, scan_idx, is_intra, sdh))
        offset += 2 * n * n
    scene.ctx.coeff_upload(np.concatenate(pool))
    cbf = scene.ctx.rdoq(t)
    got = scene.ctx.coeff_download(offset)
    for i, (coeffs, ctx, *rest) in enumerate(cases):
        want = np.zeros_like(coeffs)
        c = orc.oracle_rdoq(oracle, want, coeffs, ctx, *rest, bd)
        d = int(t[i]["dst"])
        assert np.array_equal(got[d:d + coeffs.size], want), (i, rest)
        assert bool(cbf[i]) == (c != 0)
