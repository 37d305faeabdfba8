"""CPU-only, 2 gloo ranks: the N>1 host logic -- segment sharding covers every segment exactly once, the
concatenation order reassembles the stream, and the bench's max-over-ranks timing reduction works."""
import os
import socket
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, seg_len, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from turingcodec_b200 import sharding
    mine = sharding.frames_for_rank(n_frames, seg_len, rank, world)
    # every rank "encodes" its segments: the payload is the list of frame numbers
    payload = [list(r) for r in mine]
    gathered = [None] * world
    dist.all_gather_object(gathered, payload)
    # timing reduction as in bench.py: max over ranks
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        n_segments = (n_frames + seg_len - 1) // seg_len
        stream = []
        for r, i in sharding.concatenation_order(n_segments, world):
            stream += gathered[r][i]
        out.put((stream, float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing_reduction():
    world, n_frames, seg_len = 2, 37, 8
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, seg_len, out)) for r in range(world)]
    for p in procs:
        p.start()
    stream, tmax = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert stream == list(range(n_frames))  # every frame once, in order
    assert tmax == 11.0


def test_sharding_properties():
    from turingcodec_b200 import sharding
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 8, 33):
            seen = sorted(s for r in range(world) for s in sharding.segments_for_rank(n, r, world))
            assert seen == list(range(n))


# ---- the real thing on the CPU: two ranks encode alternate IDR segments, rank 0 reassembles the reference's stream ----------

def _encode_worker(rank, world, port, clip, width, height, frames, seg, work):
    import subprocess
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(ROOT / "tests"))
    import gpu_common
    from turingcodec_b200 import sharding
    for k, rng in zip(sharding.segments_for_rank((frames + seg - 1) // seg, rank, world), sharding.frames_for_rank(frames, seg, rank, world)):
        # what integration/segments_main.cpp hands each encoder instance, with the reference encoder standing in on the CPU
        cmd = [str(gpu_common.REFERENCE_ENCODER), "encode", "--input-res", f"{width}x{height}", "--frame-rate", "30", "--speed", "fast", "--threads", "2",
               "--segment", str(seg), "--seek", str(rng.start), "--frames", str(len(rng)), "-o", f"{work}/out.seg{k}", clip]
        subprocess.run(cmd, check=True, capture_output=True)
    dist.barrier()
    if rank == 0:
        n = (frames + seg - 1) // seg
        sharding.concat_segments([f"{work}/out.seg{k}" for k in range(n)], f"{work}/out.bit")
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reassemble_the_reference_segment_stream(tmp_path):
    import subprocess

    import pytest

    import gpu_common
    from turingcodec_b200 import encoder
    if not gpu_common.REFERENCE_ENCODER.exists():
        pytest.skip("oracle/_ref/turing_ref not built")
    width, height, frames, seg = 64, 64, 11, 4
    clip = encoder.write_clip(tmp_path / "clip.yuv", width, height, frames)
    subprocess.run([str(gpu_common.REFERENCE_ENCODER), "encode", "--input-res", f"{width}x{height}", "--frame-rate", "30", "--speed", "fast", "--threads", "2",
                    "--segment", str(seg), "--frames", str(frames), "-o", str(tmp_path / "whole.bit"), str(clip)], check=True, capture_output=True)
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_encode_worker, args=(r, 2, port, str(clip), width, height, frames, seg, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert (tmp_path / "out.bit").read_bytes() == (tmp_path / "whole.bit").read_bytes()
