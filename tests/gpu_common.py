"""Shared scaffolding of the GPU parity tests: a context with a few synthetic pictures uploaded,
plus edge-padded host copies that the oracle reads with the same coordinates."""
from __future__ import annotations

import numpy as np

from turingcodec_b200 import hvb, synth

W, H, PAD = 256, 192, 96


class Scene:
    def __init__(self, bps: int, bit_depth: int, n_pictures: int = 3, scratch_pictures: int = 2):
        self.bps, self.bd = bps, bit_depth
        self.dtype = np.uint8 if bps == 1 else np.uint16
        self.ctx = hvb.Context(0, bps, bit_depth)
        self.pics, self.host = [], []
        for i in range(n_pictures):
            f = [p.astype(self.dtype) for p in synth.frame(i, W, H, bit_depth)]
            if bps == 2 and i == 1:
                f[0][::7, ::5] = (1 << bit_depth) - 1  # extremes
            pic = self.ctx.picture_create(W, H, PAD)
            self.ctx.upload_yuv(pic, *f)
            self.pics.append(pic)
            self.host.append([np.pad(pl, PAD if c == 0 else PAD // 2, mode="edge") for c, pl in enumerate(f)])
        self.scratch = [self.ctx.picture_create(W, H, PAD) for _ in range(scratch_pictures)]

    def view(self, pic_index: int, c_idx: int, x: int, y: int):
        """(array, offset, stride) of sample (x, y) of uploaded picture `pic_index` in the padded host copy"""
        pad = PAD if c_idx == 0 else PAD // 2
        a = self.host[pic_index][c_idx]
        return a, (int(y) + pad) * a.shape[1] + (int(x) + pad), a.shape[1]

    def download(self, pic: int, c_idx: int) -> np.ndarray:
        w, h = (W, H) if c_idx == 0 else (W // 2, H // 2)
        return self.ctx.picture_download(pic, c_idx, w, h)

    def close(self):
        self.ctx.close()


def block(t, name, pic, c_idx, x, y):
    t[name]["pic"], t[name]["cIdx"], t[name]["x"], t[name]["y"] = pic, c_idx, x, y


# the unmodified reference encoder built by oracle/Makefile `encoder` (test infrastructure)
from pathlib import Path as _Path
REFERENCE_ENCODER = _Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "turing_ref"
