"""Frame / segment sharding across the GPUs of one box.

The hot path has no cross-frame reduction, so multi-GPU is "replicas only": independent units are IDR-delimited
segments (`--segment N`, turing/InputQueue.cpp:229-233, :270-286); GPU g takes segments g, g+world, ... and
the host concatenates the bitstreams in segment order (concat_segments: byte for byte the stream the reference writes
for `--segment N` in one run).  No data-path collective exists; torch.distributed is used only for the timing barrier.
Used by bench.py (N > 1) and integration/segments_main.cpp's --segment-rank / --segment-ranks."""
from __future__ import annotations

import re
from pathlib import Path


def segments_for_rank(n_segments: int, rank: int, world: int) -> list[int]:
    """round-robin: the segments rank `rank` encodes, in presentation order"""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_segments, world))


def frames_for_rank(n_frames: int, segment_len: int, rank: int, world: int) -> list[range]:
    """frame ranges (one per segment) of rank `rank`; the last segment may be short"""
    n_segments = (n_frames + segment_len - 1) // segment_len
    return [range(s * segment_len, min((s + 1) * segment_len, n_frames)) for s in segments_for_rank(n_segments, rank, world)]


def concatenation_order(n_segments: int, world: int) -> list[tuple[int, int]]:
    """(rank, index within that rank's list) for every segment in output order"""
    return [(s % world, s // world) for s in range(n_segments)]


def concat_segments(parts, out_path):
    """(integration/segments_main.cpp appendSegment does the same for a one-rank job) segment 0 as written, later segments without VPS / SPS / PPS / prefix SEI, the
    first NAL unit kept carrying a four-byte start code -- the reference's own `--segment` stream (used when ranks shard a job)"""
    with open(out_path, "wb") as out:
        for k, part in enumerate(parts):
            b = Path(part).read_bytes()
            if k == 0:
                out.write(b)
                continue
            starts = [m.start() for m in re.finditer(b"\x00\x00\x01", b)]
            wrote = False
            for i, p in enumerate(starts):
                begin = p - 1 if p > 0 and b[p - 1] == 0 else p
                end = starts[i + 1] if i + 1 < len(starts) else len(b)
                if i + 1 < len(starts) and b[end - 1] == 0:
                    end -= 1
                if ((b[p + 3] >> 1) & 0x3F) in (32, 33, 34, 39):
                    continue
                if not wrote and begin == p:
                    out.write(b"\x00")
                wrote = True
                out.write(b[begin:end])
