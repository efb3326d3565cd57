"""Host-side multi-GPU logic on CPU (gloo, world_size 2): channel / frame sharding covers the
work exactly once, halo reads reproduce the unsharded frames bit for bit (oracle as the
compute), and the coefficient broadcast delivers rank 0's window."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nx_signal_b200 import sharding
from oracle import nxsignal_oracle as o
from tests.util import synth


@pytest.mark.parametrize("C,world", [(8, 1), (8, 2), (8, 8), (1024, 8), (5, 8), (7, 2), (0, 4)])
def test_channel_shards_partition(C, world):
    shards = sharding.all_channel_shards(C, world)
    covered = []
    for s in shards:
        covered += list(range(s.start, s.start + s.count))
    assert covered == list(range(C))
    assert max(s.count for s in shards) == -(-C // world) if C else True


@pytest.mark.parametrize("L,N,hop,world", [(48000, 1024, 256, 2), (48000, 1024, 256, 8), (5000, 256, 100, 3), (2000, 1024, 256, 8)])
def test_frame_shards_with_halo_reproduce_full_stft(L, N, hop, world):
    x = synth((L,), 9)
    w = o.hann(N)
    full, _, _ = o.stft_fast(x, w, overlap_length=N - hop, fft_length=N, sampling_rate=48000)
    M = full.shape[0]
    parts = []
    for r in range(world):
        fs = sharding.shard_frames(M, N, hop, world, r)
        if fs.frame_count == 0:
            continue
        seg = x[fs.sample_start:fs.sample_start + fs.sample_count]  # halo included, never exchanged
        z, _, _ = o.stft_fast(seg, w, overlap_length=N - hop, fft_length=N, sampling_rate=48000)
        assert z.shape[0] == fs.frame_count
        parts.append(z)
    np.testing.assert_array_equal(np.concatenate(parts, axis=0), full)


def test_fir_shards_cover_output():
    L, K, world = 10_000, 257, 4
    x = synth((L,), 3).astype(np.float64)
    h = o.firwin(K, [0.3]).astype(np.float64)
    full = np.convolve(x, h)
    out = np.empty_like(full)
    for r in range(world):
        o0, o1, s0, s1 = sharding.fir_shard(L, K, world, r)
        local = np.convolve(x[s0:s1], h)  # local full conv: index j <-> global s0 + j
        out[o0:o1] = local[o0 - s0:o1 - s0]
    np.testing.assert_allclose(out, full, atol=1e-12)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N = 1024
        w = torch.from_numpy(o.hann(N)) if rank == 0 else torch.zeros(N)
        sharding.broadcast_coeffs(w, src=0)
        ok_bcast = bool(np.array_equal(w.numpy(), o.hann(N)))
        # each rank transforms its channel shard with the broadcast window; rank 0 gathers to compare
        C, L, hop = 6, 9000, 256
        x = synth((C, L), 77)
        sh = sharding.shard_channels(C, world, rank)
        z, _, _ = o.stft_fast(x[sh.start:sh.start + sh.count], w.numpy(), overlap_length=N - hop, fft_length=N,
                              sampling_rate=48000)
        gathered = [None] * world
        dist.all_gather_object(gathered, (sh.start, z))
        if rank == 0:
            full, _, _ = o.stft_fast(x, o.hann(N), overlap_length=N - hop, fft_length=N, sampling_rate=48000)
            got = np.concatenate([p for _, p in sorted(gathered, key=lambda t: t[0])], axis=0)
            q.put((ok_bcast, bool(np.array_equal(got, full))))
        else:
            q.put((ok_bcast, True))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_broadcast_and_channel_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(a and b for a, b in res), res


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_bench_shards_tile_the_baseline_configs(world):
    """bench.py's multi-GPU legs: cfg3's 1024 channels split into whole 64-channel generation blocks, cfg2's 8
    channels by channel, and cfg2's 112 497 frames by frame range -- every unit owned exactly once, and a frame
    shard's sample span is exactly what its frames read."""
    from nx_signal_b200 import sharding

    covered = []
    for r in range(world):
        sh = sharding.shard_channels(1024, world, r)
        assert sh.start % 64 == 0 and sh.count % 64 == 0 and sh.count == 1024 // world
        covered += list(range(sh.start, sh.start + sh.count))
    assert covered == list(range(1024))
    covered = []
    for r in range(world):
        sh = sharding.shard_channels(8, world, r)
        covered += list(range(sh.start, sh.start + sh.count))
    assert covered == list(range(8))
    L, N, hop = 48000 * 600, 1024, 256
    M = (L - N) // hop + 1
    nxt = 0
    for r in range(world):
        fs = sharding.shard_frames(M, N, hop, world, r)
        assert fs.frame_start == nxt and fs.frame_count > 0
        assert fs.sample_start == fs.frame_start * hop
        assert fs.sample_count == (fs.frame_count - 1) * hop + N and fs.sample_start + fs.sample_count <= L
        assert (fs.sample_count - N) // hop + 1 == fs.frame_count  # stft of the span yields exactly the shard's frames
        nxt += fs.frame_count
    assert nxt == M
