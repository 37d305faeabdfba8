"""The drop-in claim, checked without a GPU: the UNMODIFIED reference encoder (oracle/_ref/turing_b200: the reference's own
objects, dynamically linked against libhvb.so) is run against a CPU build of this library -- every csrc file as written,
the CUDA runtime replaced by tests/fake_cuda, every kernel launch executed by the warp-level emulator (tests/host_build.py)
-- and must produce the same HEVC bitstream and the same reconstruction as the reference built with its own C havoc path
(`--asm 0`), the methodology of turing/signature.cpp:103-190 and of tests/test_gpu_dropin.py.  Every pixel primitive of the
encode (SAD, SATD, interpolation, intra prediction, transforms, quantisation ...) goes through include/havoc_b200.h's table
shim, the batched C-ABI, the host staging code and the kernels' source."""
import hashlib
import os
import subprocess

import pytest

import host_build
import test_gpu_dropin as dropin


@pytest.fixture(scope="module")
def cpu_libhvb(tmp_path_factory):
    d = tmp_path_factory.mktemp("libhvb_cpu")
    files = tuple(sorted(p.name for p in host_build.CSRC.glob("*.cu") if not p.name.endswith("_tma.cu")))  # (TMA has no host emulation)
    host_build.build(d, files, cpp_files=("havoc_b200.cpp",), soname="libhvb.so")
    return d


@pytest.mark.parametrize("tag,width,height,frames,options", dropin.CASES, ids=[c[0] for c in dropin.CASES])
def test_reference_encoder_on_the_emulated_library_matches_reference(cpu_libhvb, tmp_path, tag, width, height, frames, options):
    """the option sets of tests/test_gpu_dropin.py (the reference's own signature test, turing/signature.cpp:228-237)"""
    if not (dropin.REF.exists() and dropin.B200.exists()):
        pytest.skip("oracle/_ref/turing_ref / turing_b200 not built (make -C oracle encoder, needs /root/reference)")
    clip = tmp_path / "clip.yuv"
    dropin.write_clip(clip, width, height, frames)

    def encode(binary, tag, extra, lib_dir):
        bit, rec = tmp_path / f"{tag}.bit", tmp_path / f"{tag}.yuv"
        cmd = [str(binary), "encode", "--input-res", f"{width}x{height}", "--frame-rate", "24", "--frames", str(frames), "--threads", "2",
               "-o", str(bit), "--dump-pictures", str(rec), *extra, *options, str(clip)]
        env = dict(os.environ)
        if lib_dir:
            env["LD_LIBRARY_PATH"] = str(lib_dir)  # libhvb.so resolves to the CPU build
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, env=env)
        assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
        return hashlib.md5(bit.read_bytes()).hexdigest(), hashlib.md5(rec.read_bytes()).hexdigest(), bit.stat().st_size

    want = encode(dropin.REF, "ref", ["--asm", "0"], None)
    got = encode(dropin.B200, "emulated", [], cpu_libhvb)
    assert want[2] > 100  # a real bitstream came out
    assert got == want, tag


@pytest.mark.parametrize("tag,width,height,frames,options", dropin.DECODE_CASES[:2], ids=[c[0] for c in dropin.DECODE_CASES[:2]])
def test_reference_decoder_on_the_emulated_library_matches_reference(cpu_libhvb, tmp_path, tag, width, height, frames, options):
    """decoder reuse without a GPU: the reference decoder's reconstruction through the table shim and
    the kernels' source, against the reference decoder on its own C path"""
    if not (dropin.REF.exists() and dropin.B200.exists()):
        pytest.skip("oracle/_ref/turing_ref / turing_b200 not built (make -C oracle encoder, needs /root/reference)")
    clip = tmp_path / "clip.yuv"
    dropin.write_clip(clip, width, height, frames)
    dropin.encode(dropin.REF, clip, tmp_path, "ref", width, height, frames, ["--asm", "0", *options])
    want = dropin.decode(dropin.REF, tmp_path / "ref.bit", tmp_path / "dec_ref.yuv")
    got = dropin.decode(dropin.B200, tmp_path / "ref.bit", tmp_path / "dec_emulated.yuv", lib_dir=cpu_libhvb)
    assert got == want, tag
