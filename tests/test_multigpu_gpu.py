"""Two GPUs of one box, one process: the path's single collective through the C ABI
(nxs_bcast_coeffs_dev over an NCCL communicator) and the sharding rules on real devices --
channel shards and frame-range / FIR shards with their read-only halos reproduce the one-GPU
result bit for bit.  Skipped with fewer than two devices (run: gpurun --gpus 2)."""
import ctypes as C

import numpy as np
import pytest

import nx_signal_b200 as nx
from nx_signal_b200 import _arrays as A
from nx_signal_b200 import _lib, sharding
from oracle import nxsignal_oracle as o
from tests.util import synth

pytestmark = pytest.mark.gpu


def _two_devices():
    import torch

    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


needs2 = pytest.mark.skipif(not _two_devices(), reason="needs two CUDA devices")


def _nccl():
    import torch  # noqa: F401  (loads torch's bundled libnccl into the process)

    for name in ("libnccl.so.2", "libnccl.so"):
        try:
            return C.CDLL(name, mode=C.RTLD_GLOBAL)
        except OSError:
            continue
    import glob
    import os

    import nvidia.nccl  # type: ignore

    path = glob.glob(os.path.join(os.path.dirname(nvidia.nccl.__file__), "lib", "libnccl.so*"))[0]
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


@needs2
def test_coefficient_broadcast_through_the_c_abi():
    import torch

    nccl = _nccl()
    comms = (C.c_void_p * 2)()
    devs = (C.c_int * 2)(0, 1)
    assert nccl.ncclCommInitAll(comms, 2, devs) == 0
    w = nx.windows.hann(1024)
    bufs = [torch.from_numpy(w).to("cuda:0"), torch.zeros(1024, device="cuda:1")]
    ctxs = [_lib.context(0), _lib.context(1)]
    for d in (0, 1):
        torch.cuda.synchronize(d)
    assert nccl.ncclGroupStart() == 0
    for d in (0, 1):
        with torch.cuda.device(d):
            rc = _lib.lib().nxs_bcast_coeffs_dev(ctxs[d], comms[d], A.ptr(bufs[d]), 1024, 0,
                                                 C.c_void_p(torch.cuda.current_stream(d).cuda_stream))
            _lib.check(rc, ctxs[d], "bcast")
    assert nccl.ncclGroupEnd() == 0
    for d in (0, 1):
        torch.cuda.synchronize(d)
    np.testing.assert_array_equal(bufs[1].cpu().numpy(), w)
    for d in (0, 1):
        nccl.ncclCommDestroy(C.c_void_p(comms[d]))


@needs2
def test_channel_and_frame_shards_equal_single_gpu_bitwise():
    import torch

    Cn, L, N, hop = 6, 400_000, 1024, 256
    x = synth((Cn, L), 91)
    w = o.hann(N)
    kw = dict(overlap_length=N - hop, fft_length=N, sampling_rate=48000)
    full, _, _ = nx.stft(torch.from_numpy(x).to("cuda:0"), torch.from_numpy(w).to("cuda:0"), **kw)
    full = torch.view_as_real(full).cpu().numpy()
    # channel shards: rank r on device r
    parts = []
    for r in range(2):
        sh = sharding.shard_channels(Cn, 2, r)
        dev = f"cuda:{r}"
        z, _, _ = nx.stft(torch.from_numpy(x[sh.start:sh.start + sh.count]).to(dev), torch.from_numpy(w).to(dev), **kw)
        parts.append(torch.view_as_real(z).cpu().numpy())
    np.testing.assert_array_equal(np.concatenate(parts, axis=0), full)
    # frame-range shards of one channel, each with its (N - hop)-sample read-only halo
    M = full.shape[1]
    parts = []
    for r in range(2):
        fs = sharding.shard_frames(M, N, hop, 2, r)
        dev = f"cuda:{r}"
        seg = np.ascontiguousarray(x[0, fs.sample_start:fs.sample_start + fs.sample_count])
        z, _, _ = nx.stft(torch.from_numpy(seg).to(dev), torch.from_numpy(w).to(dev), **kw)
        assert z.shape[0] == fs.frame_count
        parts.append(torch.view_as_real(z).cpu().numpy())
    np.testing.assert_array_equal(np.concatenate(parts, axis=0), full[0])


@needs2
def test_fir_shards_with_halo_equal_single_gpu():
    import torch

    L, K = 1_000_000, 2049
    x = synth((1, L), 92)
    taps = nx.filters.firwin(K, [6000], sampling_rate=48000)
    conv = nx.convolution
    full = conv.convolve(torch.from_numpy(x).to("cuda:0"), torch.from_numpy(taps).to("cuda:0")[None, :], mode="full",
                         method="fft").cpu().numpy()
    parts = []
    for r in range(2):
        o0, o1, s0, s1 = sharding.fir_shard(L, K, 2, r)
        dev = f"cuda:{r}"
        seg = torch.from_numpy(np.ascontiguousarray(x[:, s0:s1])).to(dev)
        y = conv.convolve(seg, torch.from_numpy(taps).to(dev)[None, :], mode="full", method="fft").cpu().numpy()
        parts.append(y[:, o0 - s0:o1 - s0])  # the shard's own outputs; the first K-1 of a later shard are halo warm-up
    got = np.concatenate(parts, axis=1)
    assert got.shape == full.shape
    # block boundaries fall differently inside a shard, so fp32 rounding differs in the last bits
    assert np.abs(got - full).max() <= 2e-6 * np.abs(full).max()
