"""Encode one synthetic clip with the reference (--asm 1 for speed, --asm 0 for the identity md5) and with the batched B200
build; print one JSON object per run.  usage: python tools/encode_compare.py WxH frames [--threads N] [--no-asm0] [--bit-depth 10] [--opts "encoder options"]"""
import argparse
import json
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from turingcodec_b200 import encoder  # noqa: E402

REFERENCE_ENCODER = ROOT / "oracle" / "_ref" / "turing_ref"

p = argparse.ArgumentParser()
p.add_argument("res")
p.add_argument("frames", type=int)
p.add_argument("--threads", default="64")
p.add_argument("--no-asm0", action="store_true")
p.add_argument("--no-asm1", action="store_true")
p.add_argument("--bit-depth", type=int, default=8)
p.add_argument("--segments", default="", help="also run turing_b200_segments with these --parallel-segments values (needs --segment N in --opts)")
p.add_argument("--env", default="", help="comma-separated NAME=VALUE for the batched run")
p.add_argument("--opts", default="", help="encoder options, one string (default: the medium identity configuration)")
a = p.parse_args()
w, h = map(int, a.res.split("x"))
opts = a.opts.split() or encoder.MEDIUM
with tempfile.TemporaryDirectory(dir="/dev/shm" if Path("/dev/shm").exists() else None) as tmp:
    tmp = Path(tmp)
    clip = encoder.write_clip(tmp / "clip.yuv", w, h, a.frames, a.bit_depth)
    runs = {}
    if not a.no_asm1:
        runs["ref_asm1"] = encoder.encode(REFERENCE_ENCODER, clip, w, h, a.frames, ["--asm", "1", *opts], tmp, "ref1")
        print(json.dumps({"run": "ref_asm1", **runs["ref_asm1"]}), flush=True)
    if not a.no_asm0:
        runs["ref_asm0"] = encoder.encode(REFERENCE_ENCODER, clip, w, h, a.frames, ["--asm", "0", *opts], tmp, "ref0")
        print(json.dumps({"run": "ref_asm0", **runs["ref_asm0"]}), flush=True)
    env = dict(kv.split("=", 1) for kv in a.env.split(",") if kv)
    for t in [t for t in a.threads.split(",") if t and not a.segments]:
        r = encoder.encode(encoder.BATCHED, clip, w, h, a.frames, opts, tmp, f"b{t}", threads=int(t), env=env)
        ident = None
        if "ref_asm0" in runs:
            ident = (r["bitstream_md5"], r["reconstruction_md5"]) == (runs["ref_asm0"]["bitstream_md5"], runs["ref_asm0"]["reconstruction_md5"])
        print(json.dumps({"run": f"batched_threads{t}", "identical_to_asm0": ident, **r}), flush=True)
    for par in [v for v in a.segments.split(",") if v]:
        for t in a.threads.split(","):
            r = encoder.encode(encoder.SEGMENTS, clip, w, h, a.frames, ["--parallel-segments", par, *opts], tmp, f"s{par}_{t}", threads=int(t), env=env)
            ident = r["bitstream_md5"] == runs["ref_asm0"]["bitstream_md5"] if "ref_asm0" in runs else None
            print(json.dumps({"run": f"segments{par}_threads{t}", "identical_to_asm0": ident, **r}), flush=True)
