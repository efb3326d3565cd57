"""Parity at BASELINE.json's FULL sizes (cfg2, cfg3's per-GPU shard, cfg4, cfg5), where the oracle
cannot run the whole tensor: size-independent properties over the whole result (linearity, Parseval,
Hermitian symmetry, round trip, unit DC gain) plus oracle spot checks on frames / windows drawn from
the beginning, the middle and the very end of the tensors.  Tolerance: 1e-5 relative (north_star)."""
import numpy as np
import pytest

import nx_signal_b200 as nx
from nx_signal_b200 import convolution as conv
from oracle import nxsignal_oracle as o
from tests.util import TOL

pytestmark = pytest.mark.gpu

FS = 48000


def _need(gb):
    import torch

    free, _ = torch.cuda.mem_get_info()
    if free < gb * 2 ** 30:
        pytest.skip(f"needs {gb} GiB of free device memory")


def _signal(C, L, seed):
    """0.25 N(0,1) + two tones, generated on the device."""
    import torch

    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(C, L, device="cuda", generator=g) * 0.25
    t = torch.arange(L, device="cuda", dtype=torch.float32) / FS
    x += 0.5 * torch.sin(2 * np.pi * 440.0 * t) + 0.5 * torch.sin(2 * np.pi * 3000.0 * t)
    return x


def _spot_frames(x, z, w, N, H, picks):
    """frames (c, m) of the device result against the oracle on the same samples"""
    worst = 0.0
    for c, m in picks:
        seg = x[c, m * H: m * H + N].cpu().numpy()
        zo, _, _ = o.stft_fast(seg[None, :], w, overlap_length=N - H, fft_length=N, sampling_rate=FS)
        got = z[c, m].cpu().numpy()
        worst = max(worst, float(np.abs(got - zo[0, 0]).max() / np.abs(zo[0, 0]).max()))
    return worst


def test_cfg2_full_size_stft():
    """8 ch x 600 s @ 48 kHz, hann(1024), hop 256 -> 899 976 frames (7.37 GB)."""
    import torch

    _need(40)
    C, L, N, H = 8, FS * 600, 1024, 256
    M = (L - N) // H + 1
    w_np = nx.windows.hann(N)
    w = torch.from_numpy(w_np).cuda()
    kw = dict(overlap_length=N - H, fft_length=N, sampling_rate=FS)
    a = _signal(C, L, 1002)
    za, times, freqs = nx.stft(a, w, **kw)
    assert za.shape == (C, M, N) and M == 112_497
    # Hermitian symmetry over the whole tensor, channel by channel (bounded temporaries)
    for c in range(C):
        herm = (za[c, :, 1:] - za[c, :, 1:].flip(-1).conj()).abs().amax(dim=-1) / za[c].abs().amax(dim=-1)
        assert float(herm.max()) < 1e-6
    # Parseval on every frame of two channels
    for c in (0, C - 1):
        frames = a[c].unfold(-1, N, H) * w
        e_t = (frames.double() ** 2).sum(-1) * N
        e_f = (za[c].abs().double() ** 2).sum(-1)
        assert float(((e_t - e_f).abs() / e_t).max()) < 1e-5
        del frames
    # oracle spot checks: first, middle, last frames of the first and last channel
    picks = [(0, 0), (0, 1), (0, M // 2), (0, M - 1), (C - 1, 0), (C - 1, M // 3), (C - 1, M - 2), (C - 1, M - 1)]
    assert _spot_frames(a, za, w_np, N, H, picks) <= TOL
    # linearity at full size
    b = _signal(C, L, 77)
    zb, _, _ = nx.stft(b, w, **kw)
    zab, _, _ = nx.stft(2.0 * a - 3.0 * b, w, **kw)
    zab -= 2.0 * za
    zab += 3.0 * zb
    scale = float(za.abs().amax())
    assert float(zab.abs().amax()) / scale < 5e-6
    np.testing.assert_array_equal(times.cpu().numpy(), o.stft_times(N, FS, M))
    np.testing.assert_array_equal(freqs.cpu().numpy(), o.fft_frequencies(FS, N))


def test_cfg3_shard_full_size_stft():
    """One GPU's share of cfg3: 128 ch x 60 s, hann(4096), hop 1024 -> 359 552 frames (11.8 GB)."""
    import torch

    _need(30)
    C, L, N, H = 128, FS * 60, 4096, 1024
    M = (L - N) // H + 1
    w_np = nx.windows.hann(N)
    w = torch.from_numpy(w_np).cuda()
    x = _signal(C, L, 1003)
    z, _, _ = nx.stft(x, w, overlap_length=N - H, fft_length=N, sampling_rate=FS)
    assert z.shape == (C, M, N) and M == 2809
    picks = [(0, 0), (0, M - 1), (63, M // 2), (127, 0), (127, M - 1)]
    assert _spot_frames(x, z, w_np, N, H, picks) <= TOL
    for c in (0, 64, 127):
        herm = (z[c, :, 1:] - z[c, :, 1:].flip(-1).conj()).abs().amax(dim=-1) / z[c].abs().amax(dim=-1)
        assert float(herm.max()) < 1e-6
        frames = x[c].unfold(-1, N, H) * w
        e_t = (frames.double() ** 2).sum(-1) * N
        e_f = (z[c].abs().double() ** 2).sum(-1)
        assert float(((e_t - e_f).abs() / e_t).max()) < 1e-5


def test_cfg4_full_size_fir():
    """64 ch x 600 s, firwin(2049) lowpass at 6 kHz, mode :same (7.37 GB in, 7.37 GB out)."""
    import torch
    from scipy.signal import oaconvolve

    _need(60)
    C, L, K = 64, FS * 600, 2049
    taps_np = nx.filters.firwin(K, [6000], sampling_rate=FS)
    taps = torch.from_numpy(taps_np).cuda()
    a = _signal(C, L, 1004)
    ya = conv.convolve(a, taps[None, :], mode="same", method="fft")
    assert ya.shape == (C, L)
    # double-precision spot checks: head, an interior window, the tail, on three channels
    s = (K - 1) // 2
    for c in (0, 31, 63):
        for lo, hi in ((0, 30_000), (L // 2 - 15_000, L // 2 + 15_000), (L - 30_000, L)):
            xlo, xhi = max(lo - K, 0), min(hi + K, L)
            seg = a[c, xlo:xhi].cpu().numpy().astype(np.float64)
            full = oaconvolve(seg, taps_np.astype(np.float64), mode="full")
            want = full[s + (lo - xlo): s + (lo - xlo) + (hi - lo)]
            # head / tail windows see the zero padding of the real boundary only where xlo / xhi are the row ends
            got = ya[c, lo:hi].cpu().numpy()
            assert np.abs(got - want).max() / np.abs(want).max() <= TOL, (c, lo)
    # the 440 Hz and 3 kHz tones pass, the band above 6 kHz is rejected: the output of a constant is the constant
    ones = torch.ones(1, 200_000, device="cuda")
    yo = conv.convolve(ones, taps[None, :], mode="same", method="fft")
    assert float((yo[0, K:-K] - 1.0).abs().max()) < 1e-5
    # linearity at full size
    b = _signal(C, L, 78)
    yb = conv.convolve(b, taps[None, :], mode="same", method="fft")
    b *= -2.0
    b += a
    yab = conv.convolve(b, taps[None, :], mode="same", method="fft")
    yab -= ya
    yab += 2.0 * yb
    assert float(yab.abs().amax()) / float(ya.abs().amax()) < 5e-6


@pytest.mark.parametrize("onesided", [False, True])
def test_cfg5_full_size_round_trip(onesided):
    """32 ch x 60 s: istft(stft(x)) reproduces x (fp32 error <= 1e-5) away from the first / last window."""
    import torch

    _need(20)
    C, L, N, H = 32, FS * 60, 1024, 256
    w_np = nx.windows.hann(N)
    w = torch.from_numpy(w_np).cuda()
    kw = dict(overlap_length=N - H, fft_length=N, sampling_rate=FS)
    x = _signal(C, L, 1005)
    z, _, _ = nx.stft(x, w, onesided=onesided, **kw)
    y = nx.istft(z, w, onesided=onesided, **kw)
    n = y.shape[-1]
    assert n == ((L - N) // H) * H + N
    yr = y if onesided else y.real
    err = (yr[:, N:n - N] - x[:, N:n - N]).abs().amax()
    assert float(err) <= 1e-5 * float(x.abs().amax())
    if not onesided:
        assert float(y.imag[:, N:n - N].abs().amax()) <= 1e-5 * float(x.abs().amax())
    # oracle spot check of the inverse on a slice of frames of the last channel (interior samples)
    m0, m1 = 5000, 5300
    zs = z[C - 1, m0:m1].cpu().numpy()
    if onesided:
        zs = np.concatenate([zs, np.conj(zs[:, N // 2 - 1:0:-1])], axis=-1)
    yo = o.istft_fast(zs, w_np, **kw)
    lo, hi = N, (m1 - m0) * H - N
    got = yr[C - 1, m0 * H + lo: m0 * H + hi].cpu().numpy()
    assert np.abs(got - yo.real[lo:hi]).max() / np.abs(yo).max() <= TOL


# ---------------------------------------------------------------------------------------------
# Whole-channel oracle parity at the BASELINE sizes: every frame / sample of at least one full channel
# of each config against the oracle (chunked, so the f64 temporaries stay bounded), in the plain
# per-frame / per-channel max-norm bound of 1e-5 -- not a handful of spot frames.
# ---------------------------------------------------------------------------------------------
def _whole_channel_stft_err(x_c, z_c, w_np, N, H, chunk=8192):
    """max over ALL frames of one channel of max|gpu - oracle| / max|oracle| (oracle = stft_fast, chunked)."""
    M = z_c.shape[0]
    worst = 0.0
    for m0 in range(0, M, chunk):
        m1 = min(M, m0 + chunk)
        seg = x_c[m0 * H: (m1 - 1) * H + N].cpu().numpy()
        zo, _, _ = o.stft_fast(seg, w_np, overlap_length=N - H, fft_length=N, sampling_rate=FS)
        got = z_c[m0:m1].cpu().numpy()
        assert zo.shape == got.shape
        err = np.abs(got - zo).max(axis=-1) / np.abs(zo).max(axis=-1)
        worst = max(worst, float(err.max()))
    return worst


def test_cfg2_whole_channels_against_the_oracle():
    """cfg2: all 112 497 frames of the first and the last channel vs the oracle."""
    import torch

    _need(40)
    C, L, N, H = 8, FS * 600, 1024, 256
    w_np = nx.windows.hann(N)
    x = _signal(C, L, 1002)
    z, _, _ = nx.stft(x, torch.from_numpy(w_np).cuda(), overlap_length=N - H, fft_length=N, sampling_rate=FS)
    for c in (0, C - 1):
        assert _whole_channel_stft_err(x[c], z[c], w_np, N, H) <= TOL, c


def test_cfg3_whole_channels_against_the_oracle():
    """cfg3 shard: all 2 809 frames of three channels (first, middle, last of the 128) vs the oracle."""
    import torch

    _need(30)
    C, L, N, H = 128, FS * 60, 4096, 1024
    w_np = nx.windows.hann(N)
    x = _signal(C, L, 1003)
    z, _, _ = nx.stft(x, torch.from_numpy(w_np).cuda(), overlap_length=N - H, fft_length=N, sampling_rate=FS)
    for c in (0, 64, 127):
        assert _whole_channel_stft_err(x[c], z[c], w_np, N, H, chunk=1024) <= TOL, c


def test_cfg4_whole_channel_against_f64_convolution():
    """cfg4: all 28.8 M output samples of two channels vs scipy's oaconvolve in f64 (mode :same)."""
    import torch
    from scipy.signal import oaconvolve

    _need(60)
    C, L, K = 64, FS * 600, 2049
    taps_np = nx.filters.firwin(K, [6000], sampling_rate=FS)
    a = _signal(C, L, 1004)
    y = conv.convolve(a, torch.from_numpy(taps_np).cuda()[None, :], mode="same", method="fft")
    s = (K - 1) // 2
    for c in (0, C - 1):
        full = oaconvolve(a[c].cpu().numpy().astype(np.float64), taps_np.astype(np.float64), mode="full")
        want = full[s: s + L]
        got = y[c].cpu().numpy()
        assert np.abs(got - want).max() / np.abs(want).max() <= TOL, c


def test_cfg5_whole_channel_istft_including_edges():
    """cfg5: istft of a whole channel (11 247 frames) vs the oracle on EVERY output sample -- the first / last
    window included, where the reference divides by an overlap-added window energy that tends to zero and the
    f64 edge kernel does the work (plain 1e-5 bound, per channel max-norm)."""
    import torch

    _need(20)
    C, L, N, H = 32, FS * 60, 1024, 256
    w_np = nx.windows.hann(N)
    w = torch.from_numpy(w_np).cuda()
    kw = dict(overlap_length=N - H, fft_length=N, sampling_rate=FS)
    x = _signal(C, L, 1005)
    z, _, _ = nx.stft(x, w, **kw)
    y = nx.istft(z, w, **kw)
    for c in (0, C - 1):
        yo = o.istft_fast(z[c].cpu().numpy(), w_np, **kw)
        got = y[c].cpu().numpy()
        assert got.shape == yo.shape
        scale = np.abs(yo).max()
        assert np.abs(got - yo).max() / scale <= TOL, c
        # and separately on the edge samples alone, against their own scale
        for sl in (slice(0, N), slice(-N, None)):
            assert np.abs(got[sl] - yo[sl]).max() / np.abs(yo[sl]).max() <= TOL, (c, sl)
