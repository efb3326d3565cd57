"""Multi-GPU plumbing for the path: the work shards into independent units --
channels first, then frame ranges with a read-only halo -- so there is no data-path collective.
The only exchange is one broadcast of the coefficient block (window or FIR taps) from rank 0 at
setup.  One process per GPU; torch.distributed carries the broadcast (NCCL on GPUs, gloo in the
CPU tests)."""
from __future__ import annotations

from typing import List, NamedTuple, Tuple


class ChannelShard(NamedTuple):
    start: int
    count: int


class FrameShard(NamedTuple):
    frame_start: int   # first frame this rank computes
    frame_count: int
    sample_start: int  # first input sample it must hold (in padded coordinates)
    sample_count: int  # (frame_count - 1) * hop + frame_length: includes the read-only halo


def shard_channels(channels: int, world: int, rank: int) -> ChannelShard:
    """Contiguous blocks of ceil(C / G) channels per rank (the last ranks may get fewer / none)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    per = -(-channels // world)
    start = min(rank * per, channels)
    return ChannelShard(start, min(per, channels - start))


def shard_frames(num_frames: int, frame_length: int, hop: int, world: int, rank: int) -> FrameShard:
    """Split the time axis of one channel by frame range: rank r takes frames
    [r*M/G, (r+1)*M/G) and reads samples [m0*hop, (m1-1)*hop + N) -- an (N - hop)-sample halo that
    is only read, never exchanged (inputs are scattered with halo from the host)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    m0 = (num_frames * rank) // world
    m1 = (num_frames * (rank + 1)) // world
    cnt = m1 - m0
    if cnt <= 0:
        return FrameShard(m0, 0, m0 * hop, 0)
    return FrameShard(m0, cnt, m0 * hop, (cnt - 1) * hop + frame_length)


def fir_shard(length: int, num_taps: int, world: int, rank: int) -> Tuple[int, int, int, int]:
    """FIR (:full indexing): rank r produces full-convolution outputs [o0, o1) and must hold input
    samples [o0 - (K-1), o1) clipped to the signal -- a (K-1)-sample read-only halo."""
    total = length + num_taps - 1
    o0 = (total * rank) // world
    o1 = (total * (rank + 1)) // world
    s0 = max(o0 - (num_taps - 1), 0)
    s1 = min(o1, length)
    return o0, o1, s0, max(s1, s0)


def all_channel_shards(channels: int, world: int) -> List[ChannelShard]:
    return [shard_channels(channels, world, r) for r in range(world)]


def broadcast_coeffs(tensor, src: int = 0):
    """The path's single collective: rank `src` owns the window / taps, everyone else receives.
    No-op when torch.distributed is not initialised (single GPU)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(tensor, src=src)
    return tensor
