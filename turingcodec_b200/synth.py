"""Synthetic planar 4:2:0 frames of the shape SURVEY.md section 8(d) prescribes: a seeded
low-pass-filtered noise texture translated by (+3i, +2i) per frame with a second layer moving
(-2i, +i) in the centre third.  Used by tests and bench (no datasets exist offline)."""
from __future__ import annotations

import numpy as np


def _texture(rng: np.random.Generator, h: int, w: int, bit_depth: int) -> np.ndarray:
    t = rng.integers(0, 1 << bit_depth, (h, w)).astype(np.float32)
    for _ in range(2):  # separable 5-tap box blur, twice: cheap low-pass
        t = (np.roll(t, 2, 1) + np.roll(t, 1, 1) + t + np.roll(t, -1, 1) + np.roll(t, -2, 1)) / 5
        t = (np.roll(t, 2, 0) + np.roll(t, 1, 0) + t + np.roll(t, -1, 0) + np.roll(t, -2, 0)) / 5
    lo, hi = t.min(), t.max()
    t = (t - lo) / max(hi - lo, 1e-6) * ((1 << bit_depth) - 1)
    return t


def frame(index: int, width: int, height: int, bit_depth: int = 8, seed: int = 1234):
    """Returns (Y, U, V) numpy planes of frame `index` (uint8, or uint16 for bit_depth > 8)."""
    dtype = np.uint8 if bit_depth == 8 else np.uint16
    rng = np.random.default_rng(seed)
    base = _texture(rng, height, width, bit_depth)
    layer = _texture(rng, height, width, bit_depth)
    noise = np.random.default_rng(seed + 1 + index).integers(-2, 3, (height, width))
    y = np.roll(base, (2 * index, 3 * index), (0, 1)).copy()
    y0, y1, x0, x1 = height // 3, 2 * height // 3, width // 3, 2 * width // 3
    moved = np.roll(layer, (index, -2 * index), (0, 1))
    y[y0:y1, x0:x1] = moved[y0:y1, x0:x1]
    y = np.clip(y + noise, 0, (1 << bit_depth) - 1).astype(dtype)
    u = y[0::2, 0::2].copy()
    v = ((1 << bit_depth) - 1 - y[1::2, 1::2]).astype(dtype)
    return y, u, v
