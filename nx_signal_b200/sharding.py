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


# ---- the C ABI's own NCCL communicator ---------------------------------------------------------
# nxs_bcast_coeffs_dev takes an ncclComm_t.  One process per GPU: rank 0 creates an ncclUniqueId,
# the id travels over the already-initialised torch.distributed group (any backend), and every rank
# joins with ncclCommInitRank on torch's bundled libnccl -- the library itself has no link-time
# NCCL dependency.

def _libnccl():
    import ctypes as C
    import glob
    import os

    import torch  # noqa: F401  (loads its bundled libnccl into the process)

    for name in ("libnccl.so.2", "libnccl.so"):
        try:
            return C.CDLL(name, mode=C.RTLD_GLOBAL)
        except OSError:
            continue
    import nvidia.nccl  # type: ignore

    path = glob.glob(os.path.join(os.path.dirname(nvidia.nccl.__file__), "lib", "libnccl.so*"))[0]
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


class NcclComm:
    """ncclComm_t of this rank, created through ncclGetUniqueId / ncclCommInitRank."""

    def __init__(self, rank: int, world: int, device: int):
        import ctypes as C

        import torch
        import torch.distributed as dist

        class UniqueId(C.Structure):
            _fields_ = [("internal", C.c_byte * 128)]

        self._nccl = _libnccl()
        uid = UniqueId()
        if rank == 0:
            rc = self._nccl.ncclGetUniqueId(C.byref(uid))
            if rc != 0:
                raise RuntimeError(f"ncclGetUniqueId failed: {rc}")
        box = [bytes(uid.internal) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        C.memmove(C.byref(uid), box[0], 128)
        self._nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
        self.comm = C.c_void_p()
        with torch.cuda.device(device):
            rc = self._nccl.ncclCommInitRank(C.byref(self.comm), world, uid, rank)
        if rc != 0:
            raise RuntimeError(f"ncclCommInitRank failed: {rc}")
        self.rank, self.world = rank, world

    def destroy(self):
        import ctypes as C

        if self.comm:
            self._nccl.ncclCommDestroy.argtypes = [C.c_void_p]
            self._nccl.ncclCommDestroy(self.comm)
            self.comm = None


def broadcast_coeffs_c_abi(comm: "NcclComm", tensor, device: int, src: int = 0):
    """The path's single collective through the C ABI (nxs_bcast_coeffs_dev): `tensor` is a CUDA f32
    tensor on every rank; after the call all ranks hold rank `src`'s values."""
    import ctypes as C

    import torch

    from . import _lib

    ctx = _lib.context(device)
    st = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
    rc = _lib.lib().nxs_bcast_coeffs_dev(ctx, comm.comm, C.c_void_p(tensor.data_ptr()), tensor.numel(), src, st)
    _lib.check(rc, ctx, "nxs_bcast_coeffs_dev")
    return tensor
