"""GPU drop-in test: the UNMODIFIED reference encoder, linked against libhvb.so's havoc_b200 shim instead of
the reference havoc library (oracle/Makefile `encoder`), must produce the same HEVC bitstream and the same
reconstruction as the reference built with its own C havoc path (`--asm 0`, the identity oracle of SURVEY.md
section 6) -- the reference's golden-hash methodology (turing/signature.cpp:103-190) with the B200 build as the
subject.  Every pixel primitive of these encodes (SAD, SATD, interpolation, intra, DCT, quantisation, ...) runs
on the GPU, one call at a time."""
import hashlib
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF = ROOT / "oracle" / "_ref" / "turing_ref"
B200 = ROOT / "oracle" / "_ref" / "turing_b200"

pytestmark = pytest.mark.gpu


def write_clip(path, width, height, frames, bit_depth=8):
    from turingcodec_b200 import synth
    with open(path, "wb") as f:
        for i in range(frames):
            for plane in synth.frame(i, width, height, bit_depth):
                f.write(plane.tobytes())


def encode(binary, clip, out_dir, tag, width, height, frames, options):
    bit, rec = out_dir / f"{tag}.bit", out_dir / f"{tag}.yuv"
    cmd = [str(binary), "encode", "--input-res", f"{width}x{height}", "--frame-rate", "24", "--frames", str(frames),
           "--threads", "2", "-o", str(bit), "--dump-pictures", str(rec), *options, str(clip)]
    env = dict(os.environ, LD_LIBRARY_PATH=str(ROOT / "turingcodec_b200" / "csrc") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    return hashlib.md5(bit.read_bytes()).hexdigest(), hashlib.md5(rec.read_bytes()).hexdigest(), bit.stat().st_size


# the option sets of the reference's own signature test (turing/signature.cpp:228-237) that exercise this path
CASES = [
    ("fast", 128, 64, 3, ["--speed", "fast"]),
    ("slow-nosao", 64, 64, 2, ["--no-sao"]),
    ("medium", 64, 64, 2, ["--speed", "medium", "--no-wpp", "--concurrent-frames", "1"]),
    ("internal10", 64, 64, 2, ["--speed", "fast", "--bit-depth", "8", "--internal-bit-depth", "10"]),
]


@pytest.mark.parametrize("tag,width,height,frames,options", CASES, ids=[c[0] for c in CASES])
def test_encoder_on_b200_primitives_matches_reference(tmp_path, tag, width, height, frames, options):
    if not (REF.exists() and B200.exists()):
        pytest.skip("oracle/_ref/turing_ref / turing_b200 not built (make -C oracle encoder, needs /root/reference)")
    clip = tmp_path / "clip.yuv"
    write_clip(clip, width, height, frames)
    want = encode(REF, clip, tmp_path, "ref", width, height, frames, ["--asm", "0", *options])
    got = encode(B200, clip, tmp_path, "b200", width, height, frames, options)
    assert want[2] > 100  # a real bitstream came out
    assert got == want, (tag, got, want)


def decode(binary, bitstream, out_yuv, lib_dir=None):
    """`turing decode` (turing/decode.cpp): the reference decoder; with turing_b200 its reconstruction -- intra prediction,
    dequantisation, inverse transform + add, inter prediction (turing/Decode.h:396-505) -- runs through the same table shim"""
    env = dict(os.environ, LD_LIBRARY_PATH=str(lib_dir or ROOT / "turingcodec_b200" / "csrc") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    res = subprocess.run([str(binary), "decode", "-o", str(out_yuv), str(bitstream)], capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    return hashlib.md5(out_yuv.read_bytes()).hexdigest()


# decoder reuse: streams with uni- and bi-predicted PUs, intra CUs, every transform size; 8 bit and 10 bit internal
DECODE_CASES = [
    ("fast", 128, 64, 3, ["--speed", "fast"]),
    ("slow-nosao", 64, 64, 3, ["--no-sao"]),
    ("internal10", 64, 64, 2, ["--speed", "fast", "--bit-depth", "8", "--internal-bit-depth", "10"]),
]


@pytest.mark.parametrize("tag,width,height,frames,options", DECODE_CASES, ids=[c[0] for c in DECODE_CASES])
def test_decoder_on_b200_primitives_matches_reference(tmp_path, tag, width, height, frames, options):
    """the reference's signature test decodes what it encoded and compares (turing/signature.cpp:150-190): here the decoder's
    reconstruction runs on the B200 build and must equal both the reference decoder's output and the encoder's reconstruction"""
    if not (REF.exists() and B200.exists()):
        pytest.skip("oracle/_ref/turing_ref / turing_b200 not built (make -C oracle encoder, needs /root/reference)")
    clip = tmp_path / "clip.yuv"
    write_clip(clip, width, height, frames)
    _, rec_md5, size = encode(REF, clip, tmp_path, "ref", width, height, frames, ["--asm", "0", *options])
    assert size > 100
    want = decode(REF, tmp_path / "ref.bit", tmp_path / "dec_ref.yuv")
    got = decode(B200, tmp_path / "ref.bit", tmp_path / "dec_b200.yuv")
    assert got == want, tag
    if "--internal-bit-depth" not in options:  # (the dump of a 10-bit-internal encode is 16-bit, the decoder writes the same)
        assert got == rec_md5
