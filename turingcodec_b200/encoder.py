"""Driving the batched B200 build of the reference encoder (integration/_build/turing_b200_batched: the reference's encoder
with its hot loops on libhvb.so, and turing_b200_segments, its IDR-segment-parallel driver; integration/) -- and, for the
callers that compare with it, any other `turing` binary handed in (tests and bench.py pass the unmodified reference build).

All take the reference's own command line (turing/encode.cpp:61-234); this module writes the synthetic clips of SURVEY.md
section 8(d), runs `turing encode`, and reports what the reference's golden-hash test reports (turing/signature.cpp:103-190):
md5 of the bitstream and of the reconstruction -- plus wall-clock frames per second and the submission queue's counters.
Used by bench.py, tests/test_gpu_batched_encoder.py and tools/."""
from __future__ import annotations

import hashlib
import json
import os
import re
import subprocess
import time
from pathlib import Path

import numpy as np

from . import synth

ROOT = Path(__file__).resolve().parent.parent
BATCHED = ROOT / "integration" / "_build" / "turing_b200_batched"
SEGMENTS = ROOT / "integration" / "_build" / "turing_b200_segments"
LIB_DIR = ROOT / "turingcodec_b200" / "csrc"

# The configuration bit-identity is defined on.  With SAO on, `--speed medium` is not a deterministic function of its input
# in the reference itself: the SAO decision reads reconstructed samples while other pool threads still filter them
# (default: the live picture, turing/EncSao.h:302-304; `--sao-slow-mode`: the deblocked copy), and the same binary gives
# different bitstreams from run to run (profiles/r02a_asm0_asm1_experiment.txt; with --sao-slow-mode the divergence needs
# more than one SOP to show).  `--speed medium --no-sao` -- every other medium tool, RDOQ, sign hiding, deblocking, WPP and the
# four concurrent frames on -- gives one bitstream for --asm 0 and --asm 1, any thread count, every run: SURVEY.md section
# 6(b)'s fallback identity configuration.
MEDIUM = ["--speed", "medium", "--no-sao"]

def write_clip(path: Path, width: int, height: int, frames: int, bit_depth: int = 8, seed: int = 1234) -> Path:
    """planar I420 (little-endian 16-bit samples above 8 bit), the content turingcodec_b200.synth.frame defines"""
    dtype = np.uint8 if bit_depth == 8 else np.uint16
    rng = np.random.default_rng(seed)
    base = synth._texture(rng, height, width, bit_depth)
    layer = synth._texture(rng, height, width, bit_depth)
    y0, y1, x0, x1 = height // 3, 2 * height // 3, width // 3, 2 * width // 3
    top = (1 << bit_depth) - 1
    with open(path, "wb") as f:
        for index in range(frames):
            noise = np.random.default_rng(seed + 1 + index).integers(-2, 3, (height, width))
            y = np.roll(base, (2 * index, 3 * index), (0, 1)).copy()
            y[y0:y1, x0:x1] = np.roll(layer, (index, -2 * index), (0, 1))[y0:y1, x0:x1]
            y = np.clip(y + noise, 0, top).astype(dtype)
            f.write(y.tobytes())
            f.write(y[0::2, 0::2].tobytes())
            f.write((top - y[1::2, 1::2]).astype(dtype).tobytes())
    return path


def md5_file(path: Path) -> str:
    h = hashlib.md5()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 22), b""):
            h.update(chunk)
    return h.hexdigest()


def encode(binary: Path, clip: Path, width: int, height: int, frames: int, options, out_dir: Path, tag: str, threads: int | None = None,
           dump_reconstruction: bool = True, env: dict | None = None, timeout: float = 3600, frame_rate: int = 30) -> dict:
    """one `turing encode` run; returns wall seconds, fps, md5s and (batched build) the submission queue's counters"""
    bit, rec = out_dir / f"{tag}.bit", out_dir / f"{tag}.yuv"
    cmd = [str(binary)] + ([] if binary.name == SEGMENTS.name else ["encode"])
    cmd += ["--input-res", f"{width}x{height}", "--frame-rate", str(frame_rate), "--frames", str(frames), "-o", str(bit)]
    if dump_reconstruction and binary.name != SEGMENTS.name:  # (each instance of the segment driver would write the same file)
        cmd += ["--dump-pictures", str(rec)]
    else:
        dump_reconstruction = False
    if threads is not None:
        cmd += ["--threads", str(threads)]
    cmd += [*options, str(clip)]
    run_env = dict(os.environ, HVB_STATS="1")
    run_env["LD_LIBRARY_PATH"] = str(LIB_DIR) + ":" + run_env.get("LD_LIBRARY_PATH", "")
    if env:
        run_env.update(env)
    t0 = time.perf_counter()
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=run_env)
    wall = time.perf_counter() - t0
    if res.returncode != 0:
        raise RuntimeError(f"{binary.name} failed ({res.returncode}): {res.stdout[-1500:]}\n{res.stderr[-3000:]}")
    out = {"binary": binary.name, "wall_s": wall, "fps": frames / wall, "frames": frames, "bitstream_bytes": bit.stat().st_size,
           "bitstream_md5": md5_file(bit), "cmd": " ".join(cmd[1:-1])}
    bit.unlink()
    if dump_reconstruction:
        out["reconstruction_md5"] = md5_file(rec)
        rec.unlink()
    m = re.search(r"hvbenc stats: (\{.*\})", res.stderr)
    if m:
        out["queue"] = json.loads(m.group(1))
    # the encoder's own clock (turing/Encoder.cpp:357-358): wall time of the encode without process start-up and file hashing
    m = re.search(r"([\d.]+)s wall", res.stdout)
    if m:
        out["encoder_wall_s"] = float(m.group(1))
    # the same footer's host time (boost cpu_timer: "Ns wall, Ns user + Ns system"): user + system over wall = host cores kept busy
    m = re.search(r"([\d.]+)s user \+ ([\d.]+)s system", res.stdout)
    if m:
        out["host_user_s"], out["host_system_s"] = float(m.group(1)), float(m.group(2))
    return out
