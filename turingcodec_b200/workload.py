"""The per-frame hot-path workload that bench.py times and the tests replay at small sizes.

One "frame pass" is what a batch-speculative encoder would send to the device for one inter picture
at the reference's `--speed medium` settings (turing/Speed.h:32-198: full +-64 window, MET on, 1/4-pel
on, RDOQ + SDH on, RQT off, SMP at every CU size, AMP off, min CU 8):

  loops A+B  one uni-directional motion search per PU: every CU of the quadtree (64..8) x
             {2Nx2N, 2NxN, Nx2N} (turing/Search.hpp:1029-1084, :2064-2357)
  loop  D    one 35-mode intra SATD sweep per partition 32..4 (Search.hpp:39-267)
  loop  C    the TU pipeline with RDOQ for two candidates (one inter, one intra flavoured) of every CU,
             luma 32..8 and both chroma planes 16..4 (turing/Reconstruct.cpp:180-356, :731-857)

Only CUs that lie completely inside the picture are enumerated (the reference forces splits on the
bottom CTU row, Search.hpp:810-812).  Everything the reference would read from encoder state (AMVP
predictors, lambda, CABAC context snapshots, intra neighbours) is synthesised deterministically.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import hvb

CTB = 64
QP = 26  # the reference's default (turing/encode.cpp:92-153)


def lambdas(qp: int = QP):
    """(lambda, Q16 reciprocal-sqrt lambda) as the encoder derives them for a P picture:
    lambda = 0.57 * 2^((qp-12)/3); the ME cost uses 1/sqrt(lambda) in Q16 (FixedPoint.h:50-58)."""
    lam = 0.57 * 2.0 ** ((qp - 12) / 3.0)
    return lam, int((1.0 / np.sqrt(lam)) * 65536 + 0.5)


@dataclass
class FramePass:
    me: np.ndarray            # hvb.me_task_t
    intra: np.ndarray         # hvb.intra_sweep_task_t
    tu: np.ndarray            # hvb.tu_task_t
    neighbours: np.ndarray    # sample pool for the intra sweeps
    rdoq_ctx: np.ndarray      # hvb.rdoq_ctx_t snapshots
    coeff_count: int          # int16 elements the TU pipeline writes (levels)

    @property
    def units(self):
        return {"pu_searches": int(self.me.size), "intra_partitions": int(self.intra.size), "tus": int(self.tu.size)}


def _cu_grid(width, height, size):
    nx, ny = width // size, height // size
    xs, ys = np.meshgrid(np.arange(nx) * size, np.arange(ny) * size)
    return xs.reshape(-1), ys.reshape(-1)


def me_tasks(width, height, src_pic, ref_pic, seed=7, concurrent_frames=True) -> np.ndarray:
    rng = np.random.default_rng(seed)
    _, lam_q16 = lambdas()
    parts = []
    for depth in range(4):
        cu = CTB >> depth
        xs, ys = _cu_grid(width, height, cu)
        for (dx, dy, w, h, is2Nx2N) in ((0, 0, cu, cu, 1), (0, 0, cu, cu // 2, 0), (0, cu // 2, cu, cu // 2, 0),
                                         (0, 0, cu // 2, cu, 0), (cu // 2, 0, cu // 2, cu, 0)):
            n = xs.size
            t = np.zeros(n, hvb.me_task_t)
            t["src_pic"], t["ref_pic"] = src_pic, ref_pic
            t["x0"], t["y0"], t["w"], t["h"] = xs + dx, ys + dy, w, h
            # AMVP predictors: the true global motion (+3,+2 samples/frame, synth.frame) in quarter-samples,
            # jittered the way neighbouring PUs disagree
            for k in range(2):
                t["mvp"][:, k]["x"] = 12 + rng.integers(-6, 7, n)
                t["mvp"][:, k]["y"] = 8 + rng.integers(-6, 7, n)
            t["rateMvpFlag"][:, 0], t["rateMvpFlag"][:, 1] = 52429, 78643  # ~0.8 / 1.2 bits in Q16
            t["lambda"] = lam_q16
            # LimitFullPelMv (Search.hpp:1366-1407), including the concurrent-frames wavefront clamp
            t["limitMin"]["x"], t["limitMin"]["y"] = -CTB - t["x0"], -CTB - t["y0"]
            max_x = width + CTB - t["x0"].astype(np.int32) - w
            max_y = height + CTB - t["y0"].astype(np.int32) - h
            if concurrent_frames:
                x_ctb = (t["x0"].astype(np.int32) // CTB) * CTB
                y_ctb = (t["y0"].astype(np.int32) // CTB) * CTB
                max_x = np.minimum(max_x, x_ctb + 3 * CTB - t["x0"] - w - 15)
                max_y = np.minimum(max_y, y_ctb + 2 * CTB - t["y0"] - h - 15)
            t["limitMax"]["x"], t["limitMax"]["y"] = max_x, max_y
            t["prev2Nx2N"]["x"], t["prev2Nx2N"]["y"] = 12, 8
            t["smallSearchWindow"], t["met"], t["log2CbSize"] = 0, 1, 6 - depth
            t["usePrev2Nx2N"] = 0 if (is2Nx2N and depth == 0) else 1
            t["halfPel"], t["quarterPel"] = 1, 1
            parts.append(t)
    return np.concatenate(parts)


def intra_tasks(y_plane: np.ndarray, src_pic: int):
    """35-mode sweep tasks for every partition 32..4 with neighbours taken from the source picture
    (the encoder takes them from the reconstruction; the arithmetic is the same)."""
    height, width = y_plane.shape
    padded = np.pad(y_plane, ((1, 64), (1, 64)), mode="edge")  # (-1,-1) is index (0,0)
    tasks, pools, base = [], [], 0
    for log2n in (5, 4, 3, 2):
        n = 1 << log2n
        xs, ys = _cu_grid(width, height, n)
        count = xs.size
        span = 4 * n + 1
        k = np.arange(span)
        # array index k <-> neighbour: k < 2n: left column p(-1, 2n-1-k); k == 2n: corner; k > 2n: top p(k-2n-1, -1)
        col = np.where(k <= 2 * n, 0, k - 2 * n)                 # padded x (offset +1 already)
        row = np.where(k < 2 * n, 2 * n - k, 0)                  # padded y
        nb = padded[(ys[:, None] + row[None, :]).clip(0, height + 64), (xs[:, None] + col[None, :]).clip(0, width + 64)]
        t = np.zeros(count, hvb.intra_sweep_task_t)
        t["src"]["pic"], t["src"]["cIdx"], t["src"]["x"], t["src"]["y"] = src_pic, 0, xs, ys
        t["nb_unfiltered"] = base + np.arange(count) * span + 2 * n
        t["nb_filtered"] = -1
        t["log2n"], t["cIdx"], t["strong_intra_smoothing"] = log2n, 0, 1
        tasks.append(t)
        pools.append(nb.reshape(-1))
        base += count * span
    return np.concatenate(tasks), np.concatenate(pools)


def tu_tasks(width, height, src_pic, pred_pics, rec_pics, n_ctx, bit_depth=8, seed=9):
    assert len(rec_pics) == 3 * len(pred_pics)
    rng = np.random.default_rng(seed)
    tasks, offset = [], 0
    for cand, pred_pic in enumerate(pred_pics):
        is_intra = cand % 2
        for c_idx in (0, 1, 2):
            scale = 1 if c_idx == 0 else 2
            for k, log2cu in enumerate((5, 4, 3)):  # CU 64 uses four 32x32 TUs, already covered by the 32 grid
                # every (candidate, CU size) reconstructs into its own picture: the encoder keeps one
                # reconstruction per candidate (ReconstructionCache.h), they never alias
                rec_pic = rec_pics[cand * 3 + k]
                log2n = log2cu if c_idx == 0 else log2cu - 1
                n = 1 << log2n
                xs, ys = _cu_grid(width // scale, height // scale, n)
                count = xs.size
                t = np.zeros(count, hvb.tu_task_t)
                for name, pic in (("src", src_pic), ("pred", pred_pic), ("rec", rec_pic)):
                    t[name]["pic"], t[name]["cIdx"], t[name]["x"], t[name]["y"] = pic, c_idx, xs, ys
                # the prediction is motion compensated: candidate k predicts from the picture k+1 frames away,
                # displaced by the true global motion of synth.frame (+3,+2 luma samples per frame), so the
                # residual is what an encoder sees after ME (noise + the independently moving centre layer)
                t["pred"]["x"] += 3 * (cand + 1) // scale
                t["pred"]["y"] += 2 * (cand + 1) // scale
                t["levels"] = offset + np.arange(count) * n * n
                t["log2n"], t["trType"], t["cIdx"] = log2n, 0, c_idx
                t["flags"] = 1 | (is_intra << 1) | 4  # RDOQ + SDH (medium preset)
                # Reconstruct.cpp:779-786 (qp 26, chroma qp mapped equal for this synthetic pass)
                t["qscale"] = [26214, 23302, 20560, 18396, 16384, 14564][QP % 6]
                t["qshift"] = 29 - bit_depth + QP // 6 - log2n
                t["qoffset"] = (171 if is_intra else 85) << 7
                t["iqscale"] = [40, 45, 51, 57, 64, 72][QP % 6] << (QP // 6)
                t["iqshift"] = log2n - 1 + bit_depth - 8
                t["scanIdx"] = 0
                t["rdoq_ctx"] = ((ys * scale // CTB) * 61 + xs * scale // CTB) % n_ctx
                tasks.append(t)
                offset += count * n * n
    return np.concatenate(tasks), offset


def rdoq_contexts(n_ctx: int, seed=11) -> np.ndarray:
    rng = np.random.default_rng(seed)
    lam, _ = lambdas()
    c = np.zeros(n_ctx, hvb.rdoq_ctx_t)
    raw = c.view(np.uint8).reshape(n_ctx, -1)
    raw[:, :128] = rng.integers(0, 126, (n_ctx, 128))
    c["lambda"] = lam
    return c


def frame_pass(y_plane: np.ndarray, src_pic: int, ref_pic: int, pred_pics, rec_pics, n_ctx: int = 64, bit_depth: int = 8) -> FramePass:
    height, width = y_plane.shape
    intra, pool = intra_tasks(y_plane, src_pic)
    tu, coeff_count = tu_tasks(width, height, src_pic, pred_pics, rec_pics, n_ctx, bit_depth=bit_depth)
    return FramePass(me=me_tasks(width, height, src_pic, ref_pic), intra=intra, tu=tu, neighbours=pool,
                     rdoq_ctx=rdoq_contexts(n_ctx), coeff_count=coeff_count)


def algorithmic_bytes(fp: FramePass, n_sad: int, bps: int = 1) -> dict:
    """SURVEY.md section 8(d) per-unit formulas summed over one frame pass (B = bytes/sample)."""
    w, h = fp.me["w"].astype(np.int64), fp.me["h"].astype(np.int64)
    me_fixed = int((w * h * bps).sum() + 17 * ((w + 7) * (h + 7) * bps + w * h * bps).sum())
    n = (1 << fp.intra["log2n"].astype(np.int64))
    intra = int(((4 * n + 1) * bps + n * n * bps + 35 * 4).sum())
    m = (1 << fp.tu["log2n"].astype(np.int64))
    tu = int((3 * m * m * bps + 2 * m * m + 16).sum())
    return {"me_fixed": me_fixed, "me_per_sad_sample": bps, "intra": intra, "tu": tu}
