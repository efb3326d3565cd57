"""The "_host" entries (what a NIF calls) against the "_dev" entries: every host entry is a chunked
H2D | kernels | D2H pipeline (csrc/nxs_hostio.cuh) and must return exactly what the device entry
computes -- for pinned and for pageable caller memory, for every transfer mode of the STFT host call,
for padded rows, and when calls on one context alternate between CUDA streams."""
import ctypes as C

import numpy as np
import pytest

import nx_signal_b200 as nx
from nx_signal_b200 import _arrays as A
from nx_signal_b200 import _lib
from nx_signal_b200 import convolution as conv
from oracle import nxsignal_oracle as o
from tests.util import TOL, frame_rel_err, synth

pytestmark = pytest.mark.gpu
FS = 48000


def _bits(a):
    import torch

    if A.is_torch(a):
        a = (torch.view_as_real(a) if a.is_complex() else a).cpu().numpy()
    a = np.ascontiguousarray(a)
    return a.view(np.int32).ravel() if a.dtype.itemsize % 4 == 0 else a.ravel()


def _pinned(arr):
    import torch

    t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
    return t


@pytest.mark.parametrize("mode", [-1, 0, 1, 3])
@pytest.mark.parametrize("pinned_in,pinned_out", [(True, True), (False, False), (True, False), (False, True)])
def test_stft_host_entry_equals_device_entry_for_every_memory_kind_and_mode(mode, pinned_in, pinned_out):
    """several chunks (12 channels), several slabs per chunk, frames not a multiple of the 64-frame work items"""
    import torch

    Cn, L, N, H = 12, 2_400_003, 1024, 256  # 4 channel chunks of 3: the mixed mode applies
    x = synth((Cn, L), 311)
    w = o.hann(N)
    M = (L - N) // H + 1
    zd, _, _ = nx.stft(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), overlap_length=N - H, fft_length=N, sampling_rate=FS)
    xin = _pinned(x) if pinned_in else x.copy()
    zout = torch.empty((Cn, M, N), dtype=torch.complex64, pin_memory=True) if pinned_out else np.empty((Cn, M, N), np.complex64)
    ctx = _lib.context(0)
    _lib.set_host_mode(mode)
    try:
        for _ in range(3 if mode == -1 else 1):  # auto: both automatic modes get explored
            if A.is_torch(zout):
                zout.zero_()
            else:
                zout[...] = 0
            rc = _lib.lib().nxs_stft_f32_host(ctx, A.ptr(xin), Cn, L, L, w.ctypes.data, N, H, N, _lib.PAD_VALID, 0, 0,
                                              _lib.SCALE_NONE, float(FS), A.ptr(zout))
            _lib.check(rc, ctx, "stft(host)")
            info = _lib.host_mode()
            assert info["input_staged"] == (not pinned_in)
            if not pinned_out:
                assert info["result"] == "onesided_d2h+pinned_ring_unstage"
            elif mode >= 0:
                assert info["result"] == _lib.HOST_MODES[mode]
            np.testing.assert_array_equal(_bits(zout), _bits(zd))
    finally:
        _lib.set_host_mode(-1)


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("Cn,L,x_ld,N,H", [(1, 48_000, 48_000, 1024, 256), (3, 20_000, 20_040, 512, 128), (2, 5_000, 5_000, 120, 40)])
def test_stft_small_call_path_equals_device_entry(pinned, Cn, L, x_ld, N, H):
    """calls whose input and result fit 4 MiB (BASELINE configs[0] is one) take the single-stream path
    (csrc/nxs_hostio.cu): same bits as the device entry, padded input rows, pinned or pageable memory"""
    import torch

    xfull = synth((Cn, x_ld), 411)
    x = xfull[:, :L]
    w = o.hann(N)
    M = (L - N) // H + 1
    zd, _, _ = nx.stft(torch.from_numpy(np.ascontiguousarray(x)).cuda(), torch.from_numpy(w).cuda(), overlap_length=N - H,
                       fft_length=N, sampling_rate=FS)
    xin = _pinned(xfull) if pinned else xfull
    zout = torch.zeros((Cn, M, N), dtype=torch.complex64, pin_memory=True) if pinned else np.zeros((Cn, M, N), np.complex64)
    ctx = _lib.context(0)
    for _ in range(2):
        rc = _lib.lib().nxs_stft_f32_host(ctx, A.ptr(xin), Cn, L, x_ld, w.ctypes.data, N, H, N, _lib.PAD_VALID, 0, 0,
                                          _lib.SCALE_NONE, float(FS), A.ptr(zout))
        _lib.check(rc, ctx, "stft(host, small)")
        assert _lib.host_mode()["result"] == ("small_call_single_stream_full_d2h" if pinned else "onesided_d2h+pinned_ring_unstage")
        np.testing.assert_array_equal(_bits(zout), _bits(zd))


@pytest.mark.parametrize("pinned", [True, False])
def test_small_calls_of_the_other_host_entries_equal_the_device_entries(pinned):
    """host_pipeline's one-chunk path (everything on the compute stream): ISTFT and FIR on a short signal"""
    import torch

    rng = np.random.default_rng(412)
    M, N, H = 184, 1024, 256
    z = (rng.standard_normal((2, M, N)) + 1j * rng.standard_normal((2, M, N))).astype(np.complex64)
    w = o.hann(N)
    yd = nx.istft(torch.from_numpy(z).cuda(), torch.from_numpy(w).cuda(), overlap_length=N - H, fft_length=N, sampling_rate=FS)
    out_len = M * H + N - H
    zin = _pinned(z) if pinned else z
    y = torch.zeros((2, out_len), dtype=torch.complex64, pin_memory=True) if pinned else np.zeros((2, out_len), np.complex64)
    ctx = _lib.context(0)
    _lib.check(_lib.lib().nxs_istft_c64_host(ctx, A.ptr(zin), 2, M, N, w.ctypes.data, N, H, N, 0, float(FS), A.ptr(y)), ctx, "istft")
    np.testing.assert_array_equal(_bits(y), _bits(yd))
    Cn, L, K = 5, 30_000, 129
    x = synth((Cn, L), 413)
    taps = synth((1, K), 414)[0]
    for mode in ("full", "same", "valid"):
        out_len = {"full": L + K - 1, "same": L, "valid": L - K + 1}[mode]
        yd = conv.convolve(torch.from_numpy(x).cuda(), torch.from_numpy(taps).cuda()[None, :], mode=mode, method="fft")
        xin = _pinned(x) if pinned else x
        yh = torch.zeros((Cn, out_len), dtype=torch.float32).pin_memory() if pinned else np.zeros((Cn, out_len), np.float32)
        rc = _lib.lib().nxs_fir_f32_host(ctx, A.ptr(xin), Cn, L, L, taps.ctypes.data, K, _lib.MODE[mode], A.ptr(yh), out_len)
        _lib.check(rc, ctx, "fir(host, small)")
        np.testing.assert_array_equal(_bits(yh), _bits(yd))


def test_stft_host_generic_length_on_pageable_memory():
    """a non-power-of-two fft_length has no mirror mode: full rows travel, through the ring when pageable"""
    import torch

    x = synth((3, 20_000), 312)
    w = o.hann(100)
    z_np, _, _ = nx.stft(x, w, overlap_length=60, fft_length=120, sampling_rate=FS)
    zd, _, _ = nx.stft(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), overlap_length=60, fft_length=120, sampling_rate=FS)
    np.testing.assert_array_equal(_bits(z_np), _bits(zd))
    zo, _, _ = o.stft_fast(x, w, overlap_length=60, fft_length=120, sampling_rate=FS)
    assert frame_rel_err(z_np, zo) <= TOL


@pytest.mark.parametrize("pinned", [True, False])
def test_istft_host_pipeline_equals_device_entry(pinned):
    import torch

    Cn, M, N, H = 10, 3001, 1024, 256
    rng = np.random.default_rng(313)
    z = (rng.standard_normal((Cn, M, N)) + 1j * rng.standard_normal((Cn, M, N))).astype(np.complex64)
    w = o.hann(N)
    kw = dict(overlap_length=N - H, fft_length=N, sampling_rate=FS)
    yd = nx.istft(torch.from_numpy(z).cuda(), torch.from_numpy(w).cuda(), **kw)
    zin = _pinned(z) if pinned else z
    out_len = M * H + N - H
    y = torch.empty((Cn, out_len), dtype=torch.complex64, pin_memory=True) if pinned else np.empty((Cn, out_len), np.complex64)
    ctx = _lib.context(0)
    rc = _lib.lib().nxs_istft_c64_host(ctx, A.ptr(zin), Cn, M, N, w.ctypes.data, N, H, N, 0, float(FS), A.ptr(y))
    _lib.check(rc, ctx, "istft(host)")
    np.testing.assert_array_equal(_bits(y), _bits(yd))


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("mode", ["full", "same", "valid"])
def test_fir_host_pipeline_equals_device_entry_and_respects_row_padding(pinned, mode):
    """x_ld > length and y_ld > out_len: the gap bytes of the caller's result rows must stay untouched"""
    import torch

    Cn, L, K, x_ld = 9, 500_000, 2049, 500_008
    x = np.zeros((Cn, x_ld), np.float32)
    x[:, :L] = synth((Cn, L), 314)
    taps = nx.filters.firwin(K, [6000], sampling_rate=FS)
    out_len = {"full": L + K - 1, "same": L, "valid": L - K + 1}[mode]
    y_ld = out_len + 5
    yd = conv.convolve(torch.from_numpy(x[:, :L].copy()).cuda(), torch.from_numpy(taps).cuda()[None, :], mode=mode, method="fft")
    xin = _pinned(x) if pinned else x
    y = torch.full((Cn, y_ld), 7.0, dtype=torch.float32).pin_memory() if pinned else np.full((Cn, y_ld), 7.0, np.float32)
    ctx = _lib.context(0)
    rc = _lib.lib().nxs_fir_f32_host(ctx, A.ptr(xin), Cn, L, x_ld, taps.ctypes.data, K, _lib.MODE[mode], A.ptr(y), y_ld)
    _lib.check(rc, ctx, "fir(host)")
    got = y.numpy() if A.is_torch(y) else y
    np.testing.assert_array_equal(_bits(got[:, :out_len]), _bits(yd))
    assert np.all(got[:, out_len:] == 7.0)


def test_framing_host_pipelines_equal_device_entries():
    import torch

    x = synth((7, 300_001), 315)
    f_np = nx.as_windowed(x, window_length=400, stride=160, padding="reflect")
    f_dev = nx.as_windowed(torch.from_numpy(x).cuda(), window_length=400, stride=160, padding="reflect")
    np.testing.assert_array_equal(_bits(f_np), _bits(f_dev))
    t = synth((7, 900, 512), 316)
    s_np = nx.overlap_and_add(t, overlap_length=384)
    s_dev = nx.overlap_and_add(torch.from_numpy(t).cuda(), overlap_length=384)
    np.testing.assert_array_equal(_bits(s_np), _bits(s_dev))


def test_calls_alternating_between_streams_on_one_context_do_not_clobber_each_other():
    """the prepared window / scratch are per context; a call on stream B right after an
    asynchronous call on stream A (different window, different scaling) must not overwrite what A's
    kernels are still reading.  The context orders B behind A with an event."""
    import torch

    Cn, L, N, H = 8, 4_000_000, 1024, 256
    x = torch.from_numpy(synth((Cn, L), 317)).cuda()
    w1 = torch.from_numpy(o.hann(N)).cuda()
    w2 = torch.from_numpy(o.hamming(N)).cuda()
    kw = dict(overlap_length=N - H, fft_length=N, sampling_rate=FS)
    ref1, _, _ = nx.stft(x, w1, scaling="spectrum", **kw)
    ref2, _, _ = nx.stft(x, w2, scaling="psd", **kw)
    z = nx.istft(ref1, w1, scaling="spectrum", **kw)  # leaves other state in the context's scratch
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(5):
        with torch.cuda.stream(s1):
            a, _, _ = nx.stft(x, w1, scaling="spectrum", **kw)
        with torch.cuda.stream(s2):
            b, _, _ = nx.stft(x, w2, scaling="psd", **kw)
        with torch.cuda.stream(s1):
            y = nx.istft(a, w1, scaling="spectrum", **kw)
        torch.cuda.synchronize()
        assert torch.equal(torch.view_as_real(a), torch.view_as_real(ref1))
        assert torch.equal(torch.view_as_real(b), torch.view_as_real(ref2))
        assert torch.equal(torch.view_as_real(y), torch.view_as_real(z))


@pytest.mark.parametrize("where", ["cuda", "host"])
def test_stft_of_complex_data(where):
    """the reference's graph is a complex transform for complex data (lib/nx_signal.ex:101-102)"""
    import torch

    rng = np.random.default_rng(318)
    x = (synth((3, 20_000), 318) + 1j * rng.standard_normal((3, 20_000))).astype(np.complex64)
    w = o.hann(512)
    kw = dict(overlap_length=384, fft_length=512, sampling_rate=FS, window_padding="reflect")
    zo, to, fo = o.stft_fast(x, w, **kw)
    if where == "cuda":
        z, t, f = nx.stft(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), **kw)
        z, t, f = z.cpu().numpy(), t.cpu().numpy(), f.cpu().numpy()
    else:
        z, t, f = nx.stft(x, w, **kw)
    assert z.shape == zo.shape and z.dtype == np.complex64
    assert frame_rel_err(z, zo) <= TOL
    np.testing.assert_array_equal(t, to)
    np.testing.assert_array_equal(f, fo)
    # a complex signal's spectrum has no mirror symmetry: the two halves differ
    assert np.abs(z[..., 1:] - np.conj(z[..., :0:-1])).max() > 1.0


def test_convolve_rank_4_with_batch_like_leading_axes():
    """convolution.ex:95-211 is rank-generic; size-1 leading axes of in2 are batch axes of in1"""
    import torch

    rng = np.random.default_rng(319)
    a = rng.standard_normal((2, 3, 5, 40)).astype(np.float32)
    b = rng.standard_normal((1, 1, 2, 7)).astype(np.float32)
    for mode in ("full", "same", "valid"):
        want = o.convolve(a, b, mode=mode)
        got = conv.convolve(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), mode=mode).cpu().numpy()
        assert got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-5 * np.abs(want).max())
        got_h = conv.convolve(a, b, mode=mode)
        np.testing.assert_allclose(got_h, want, rtol=0, atol=1e-5 * np.abs(want).max())
    with pytest.raises(NotImplementedError):
        conv.convolve(a, rng.standard_normal((2, 1, 2, 7)).astype(np.float32))


def test_complex_convolve_with_operands_in_different_places():
    """in1 on the device, in2 a numpy array (complex operands)"""
    import torch

    rng = np.random.default_rng(320)
    a = (rng.standard_normal((4, 30)) + 1j * rng.standard_normal((4, 30))).astype(np.complex64)
    b = (rng.standard_normal((1, 5)) + 1j * rng.standard_normal((1, 5))).astype(np.complex64)
    want = o.convolve(a, b, mode="full")
    got = conv.convolve(torch.from_numpy(a).cuda(), b, mode="full")
    assert got.is_cuda
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=1e-5 * np.abs(want).max())
    got2 = conv.convolve(a, torch.from_numpy(b).cuda(), mode="full")
    np.testing.assert_allclose(np.asarray(got2), want, rtol=0, atol=1e-5 * np.abs(want).max())


def test_wiener_returns_the_input_type():
    """filters.ex:108-110: computed in f64, cast back to the type of `t` (integers truncate)"""
    t = (np.arange(36).reshape(6, 6) % 7).astype(np.int32)
    got = nx.Filters.wiener(t, kernel_size=3)
    want = o.wiener(t, kernel_size=3)
    assert got.dtype == np.int32
    assert np.asarray(want).dtype == np.int32
    np.testing.assert_array_equal(got, want)
