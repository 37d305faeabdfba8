"""Frame / segment sharding across the GPUs of one box.

The hot path has no cross-frame reduction, so multi-GPU is "replicas only": independent units are IDR-delimited
segments (`--segment N`, turing/InputQueue.cpp:229-233, :270-286); GPU g takes segments g, g+world, ... and
the host concatenates the bitstreams in segment order.  No data-path collective exists; torch.distributed is
used only for the timing barrier and for gathering per-rank byte counts."""
from __future__ import annotations


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
