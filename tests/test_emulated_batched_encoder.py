"""The batched encoder without a GPU: integration/_build/turing_b200_batched (the reference encoder with its motion search, PU
cost, intra sweep and transform blocks posted to the submission queue of include/hvb_encoder.h) run against a CPU build of
libhvb.so -- every csrc file as written, hvb_encoder.cpp included, the CUDA runtime replaced by tests/fake_cuda, every kernel
launch executed by the warp emulator (tests/host_build.py) -- must write the bitstream and the reconstruction of the reference's
`--asm 0` run.  Several engines and more pool threads than one, so the queue's hand-over, the picture pool shared between
contexts and the per-thread memo are exercised; the segment-parallel driver's output must equal the reference's `--segment` run."""
import hashlib
import os
import subprocess

import pytest

import gpu_common
import host_build
from turingcodec_b200 import encoder


@pytest.fixture(scope="module")
def cpu_libhvb(tmp_path_factory):
    if not (gpu_common.REFERENCE_ENCODER.exists() and encoder.BATCHED.exists() and encoder.SEGMENTS.exists()):
        pytest.skip("turing_ref / turing_b200_batched not built (make -C oracle encoder && make -C integration; needs /root/reference)")
    d = tmp_path_factory.mktemp("libhvb_cpu_batched")
    files = tuple(sorted(p.name for p in host_build.CSRC.glob("*.cu") if not p.name.endswith("_tma.cu")))  # (TMA has no host emulation)
    host_build.build(d, files, cpp_files=("havoc_b200.cpp", "hvb_encoder.cpp"), soname="libhvb.so")
    return d


def run(cmd, lib_dir=None, **extra_env):
    env = dict(os.environ, **extra_env)
    if lib_dir:
        env["LD_LIBRARY_PATH"] = str(lib_dir)  # libhvb.so resolves to the CPU build
    res = subprocess.run([str(c) for c in cmd], capture_output=True, text=True, timeout=1500, env=env)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-2500:]
    return res


CASES = [("fast", 128, 64, 3, ["--speed", "fast"]), ("medium", 128, 64, 3, encoder.MEDIUM)]


@pytest.mark.parametrize("tag,width,height,frames,options", CASES, ids=[c[0] for c in CASES])
def test_batched_encoder_on_the_emulated_library_matches_reference(cpu_libhvb, tmp_path, tag, width, height, frames, options):
    clip = encoder.write_clip(tmp_path / "clip.yuv", width, height, frames)
    common = ["--input-res", f"{width}x{height}", "--frame-rate", "24", "--frames", frames, *options]
    run([gpu_common.REFERENCE_ENCODER, "encode", "--asm", "0", "-o", tmp_path / "ref.bit", "--dump-pictures", tmp_path / "ref.yuv", *common, clip])
    res = run([encoder.BATCHED, "encode", "--threads", "3", "-o", tmp_path / "b.bit", "--dump-pictures", tmp_path / "b.yuv", *common, clip], cpu_libhvb,
              HVB_ENGINES="3", HVB_STATS="1")
    for name in ("bit", "yuv"):
        assert hashlib.md5((tmp_path / f"b.{name}").read_bytes()).digest() == hashlib.md5((tmp_path / f"ref.{name}").read_bytes()).digest(), (tag, name)
    assert '"me": {"tasks": 0' not in res.stderr and "hvbenc stats" in res.stderr  # the queue did the work


MODES = {
    "thread_per_row": dict(HVB_FIBERS="0"),                                  # the reference's pool loop: a hand-over sleeps on a condition variable
    "service_threads": dict(HVB_SERVICE_THREADS="2"),                        # no dispatcher / completion threads: two threads sweep the engines
    "stream_queries": dict(HVB_DONE_FLAG="0"),                               # completion by hvb_poll instead of the flag the stream writes
    "engines_by_kind": dict(HVB_ENGINES="12", HVB_ENGINE_SHARES="1,1,1,1,1,7"),
    "transform_blocks_only": dict(HVB_HOOKS="48", HVB_INTRA_TU_MIN_LOG2="2", HVB_TU_MIN_LOG2="3"),  # bench.py's hooks, every block size
}


@pytest.mark.parametrize("mode", sorted(MODES))
def test_queue_modes_give_the_same_bitstream(cpu_libhvb, tmp_path, mode):
    """every way the submission queue and the pool can be run (fiber schedulers are the default of the other tests) writes the reference's stream"""
    width, height, frames = 128, 64, 2
    clip = encoder.write_clip(tmp_path / "clip.yuv", width, height, frames)
    common = ["--input-res", f"{width}x{height}", "--frame-rate", "24", "--frames", frames, *encoder.MEDIUM]
    run([gpu_common.REFERENCE_ENCODER, "encode", "--asm", "0", "-o", tmp_path / "ref.bit", *common, clip])
    env = dict(HVB_ENGINES="3", HVB_STATS="1")
    env.update(MODES[mode])
    res = run([encoder.BATCHED, "encode", "--threads", "3", "-o", tmp_path / "b.bit", *common, clip], cpu_libhvb, **env)
    assert (tmp_path / "b.bit").read_bytes() == (tmp_path / "ref.bit").read_bytes(), mode
    assert "hvbenc stats" in res.stderr and '"tu_chain": {"tasks": 0' not in res.stderr


def test_segment_parallel_driver_on_the_emulated_library_matches_reference_segment_run(cpu_libhvb, tmp_path):
    width, height, frames, seg = 64, 64, 7, 3
    clip = encoder.write_clip(tmp_path / "clip.yuv", width, height, frames)
    common = ["--input-res", f"{width}x{height}", "--frame-rate", "24", "--frames", frames, "--segment", seg, "--speed", "fast", "--threads", "2"]
    run([gpu_common.REFERENCE_ENCODER, "encode", "--asm", "0", "-o", tmp_path / "ref.bit", *common, clip])
    run([encoder.SEGMENTS, "--parallel-segments", "3", "-o", tmp_path / "s.bit", *common, clip], cpu_libhvb, HVB_ENGINES="2")
    assert (tmp_path / "s.bit").read_bytes() == (tmp_path / "ref.bit").read_bytes()
