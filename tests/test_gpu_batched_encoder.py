"""GPU identity test of the batched encoder (integration/_build/turing_b200_batched: the reference encoder with its motion
search, PU cost, intra sweep and transform blocks on libhvb.so through the submission queue of include/hvb_encoder.h): for
each option set it must produce the HEVC bitstream and the reconstruction of the reference built with its own C havoc path
(`--asm 0`) -- the reference's golden-hash methodology (turing/signature.cpp:103-190) with the batched
build as the subject.  More worker threads than cores on purpose: the workers block on the device, the batches form from
what is in flight."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

pytestmark = pytest.mark.gpu

CASES = [
    # BASELINE.json configs[0] (plumbing case): 640x360, --speed fast; frames limited to keep the GPU suite short
    ("fast-640x360", 640, 360, 10, 8, ["--speed", "fast"]),
    # the identity configuration of configs[1] / [2] at a size the suite affords: every medium tool, WPP, 4 concurrent frames
    ("medium-640x360", 640, 360, 17, 8, ["--speed", "medium", "--no-sao"]),
    # SAO on (statistics from the deblocked copy): short enough that the reference's own SAO race does not show
    ("medium-saoslow-352x288", 352, 288, 5, 8, ["--speed", "medium", "--sao-slow-mode"]),
    # configs[3]: 16-bit sample path, slow preset (RQT, AMP, no early-outs)
    ("slow-10bit-320x192", 320, 192, 3, 10, ["--speed", "slow", "--bit-depth", "10"]),
    ("slow-internal10-256x128", 256, 128, 3, 8, ["--speed", "slow", "--bit-depth", "8", "--internal-bit-depth", "10"]),
]


@pytest.mark.parametrize("tag,width,height,frames,bit_depth,options", CASES, ids=[c[0] for c in CASES])
def test_batched_encoder_matches_reference_asm0(tmp_path, tag, width, height, frames, bit_depth, options):
    import gpu_common
    from turingcodec_b200 import encoder
    if not (gpu_common.REFERENCE_ENCODER.exists() and encoder.BATCHED.exists()):
        pytest.skip("turing_ref / turing_b200_batched not built (make -C oracle encoder && make -C integration; needs /root/reference)")
    clip = encoder.write_clip(tmp_path / "clip.yuv", width, height, frames, bit_depth)
    want = encoder.encode(gpu_common.REFERENCE_ENCODER, clip, width, height, frames, ["--asm", "0", *options], tmp_path, "ref", frame_rate=24)
    got = encoder.encode(encoder.BATCHED, clip, width, height, frames, options, tmp_path, "batched", threads=48, frame_rate=24)
    assert want["bitstream_bytes"] > 100
    assert (got["bitstream_md5"], got["reconstruction_md5"]) == (want["bitstream_md5"], want["reconstruction_md5"]), (tag, got, want)
    q = got["queue"]
    # the device did the work: every hooked loop went through the queue
    assert q["kernel_launches"] > 0 and all(q[k]["tasks"] > 0 for k in ("me", "pu_cost", "intra_sweep", "tu_chain")), q
