"""GPU parity of the log-mel epilogue (NxSignal.stft_to_mel/3, lib/nx_signal.ex:486-513) against
the oracle.  Outputs are O(1) after the (x + 4) / 4 map; tolerance 1e-5 absolute == relative."""
import numpy as np
import pytest

import nx_signal_b200 as nx
from oracle import nxsignal_oracle as o
from tests.util import TOL, synth

pytestmark = pytest.mark.gpu


def test_doctest_stft_to_mel():  # lib/nx_signal.ex:465-483
    kw = dict(overlap_length=2, fft_length=16, sampling_rate=8.0e3, window_padding="reflect")
    z, _, _ = nx.stft(np.arange(10, dtype=np.int32), nx.windows.hann(4), **kw)
    mel = nx.stft_to_mel(z, 8.0e3, fft_length=16, mel_bins=4)
    want = np.array([[0.29005307, 0.17422175, 0.18422472, 0.09807998],
                     [0.6093881, 0.5647397, 0.43538243, 0.086352706],
                     [0.75841033, 0.70850146, 0.5636921, 0.17911881],
                     [0.8461772, 0.7952491, 0.64707625, 0.25204098],
                     [0.9085489, 0.85726047, 0.70786566, 0.30867678],
                     [0.9085489, 0.85726047, 0.70786566, 0.30867678]], np.float32)
    assert mel.shape == (6, 4) and mel.dtype == np.float32
    np.testing.assert_allclose(mel, want, atol=TOL, rtol=0)


@pytest.mark.parametrize("nfft,mels,sr", [(1024, 128, 48000), (512, 80, 16000), (400, 40, 22050), (2048, 128, 44100),
                                           (256, 33, 8000), (8192, 256, 48000)])
def test_stft_to_mel_vs_oracle(nfft, mels, sr):
    N = min(nfft, 1024) if nfft != 400 else 400
    x = synth((2, 30 * nfft), nfft + mels, fs=float(sr))
    w = o.hann(N)
    z, _, _ = o.stft_fast(x, w, overlap_length=N // 2, fft_length=nfft, sampling_rate=sr)
    got = nx.stft_to_mel(z, sr, fft_length=nfft, mel_bins=mels)
    want = np.stack([o.stft_to_mel(z[c], sr, nfft, mels) for c in range(z.shape[0])])  # per-entry reduce_max
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, atol=TOL, rtol=0)


def test_dynamic_range_clamp_is_per_channel_and_active():
    """A loud and a quiet channel: each is clamped against its own maximum - 8 (log10 units), and the
    quiet bins of the loud channel do hit the floor."""
    rng = np.random.default_rng(3)
    nfft, mels, sr = 512, 64, 16000
    z = (rng.standard_normal((2, 50, nfft)) + 1j * rng.standard_normal((2, 50, nfft))).astype(np.complex64)
    z[0, :, : nfft // 8] *= 1e6   # > 8 decades of power between the low and high mel bins of channel 0
    z[1] *= 1e-3
    got = nx.stft_to_mel(z, sr, fft_length=nfft, mel_bins=mels)
    want = np.stack([o.stft_to_mel(z[c], sr, nfft, mels) for c in range(2)])
    np.testing.assert_allclose(got, want, atol=TOL, rtol=0)
    floor0 = got[0].max() - 2.0  # (max - 8 + 4) / 4 == max' - 2 after the affine map
    assert np.isclose(got[0].min(), floor0, atol=1e-6) and (got[0] == got[0].min()).sum() > 10


def test_onesided_spectrum_and_device_tensors_at_scale():
    """cfg2-style chain on the device: one-sided STFT -> log-mel, equal to the two-sided chain."""
    import torch

    C, L, nfft, hop = 4, 48000 * 20, 1024, 256
    x = torch.from_numpy(synth((C, L), 9)).cuda()
    w = torch.from_numpy(o.hann(nfft)).cuda()
    kw = dict(overlap_length=nfft - hop, fft_length=nfft, sampling_rate=48000)
    z2, _, _ = nx.stft(x, w, **kw)
    z1, _, _ = nx.stft(x, w, onesided=True, **kw)
    m2 = nx.stft_to_mel(z2, 48000, fft_length=nfft, mel_bins=128)
    m1 = nx.stft_to_mel(z1, 48000, fft_length=nfft, mel_bins=128)
    assert m1.is_cuda and m1.shape == (C, z2.shape[1], 128)
    assert torch.equal(m1, m2)
    want = o.stft_to_mel(z2[1, :300].cpu().numpy(), 48000, nfft, 128)
    # the oracle's maximum is over 300 frames only: compare where neither side is clamped
    got = m1[1, :300].cpu().numpy()
    free = (want > want.min() + 1e-3) & (got > got.min() + 1e-3)
    assert free.mean() > 0.9
    np.testing.assert_allclose(got[free], want[free], atol=TOL, rtol=0)


def test_errors():
    z = np.zeros((5, 16), np.complex64)
    with pytest.raises(nx.NxSignalArgumentError, match="fft_length"):
        nx.stft_to_mel(z, 8000.0)
    with pytest.raises(nx.NxSignalArgumentError, match="fewer than"):
        nx.stft_to_mel(z, 8000.0, fft_length=64, mel_bins=4)


# ---- fused STFT -> log-mel (the spectrum never leaves the SM) ------------------------------------
@pytest.mark.parametrize("nfft,hop,mels,padding,scaling", [
    (1024, 256, 128, "valid", None), (1024, 256, 80, "reflect", "spectrum"), (512, 128, 64, "valid", None),
    (2048, 512, 128, "same", "psd"), (4096, 1024, 128, "valid", None), (8192, 2048, 256, "valid", None),
    (1024, 1024, 40, "valid", None), (1024, 64, 300, "valid", None),
    (1024, 250, 128, "valid", None),   # hop % 4 != 0: served by chaining the two entries
    (400, 160, 40, "valid", None),     # generic length: chained
])
def test_fused_stft_mel_equals_chain_and_oracle(nfft, hop, mels, padding, scaling):
    import torch

    sr = 48000
    x = synth((3, 40 * nfft + 4 * 33), 5 + nfft + hop)
    w = o.hann(nfft)
    kw = dict(overlap_length=nfft - hop, fft_length=nfft, sampling_rate=sr, window_padding=padding, scaling=scaling)
    xd, wd = torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda()
    fused = nx.stft_mel(xd, wd, mel_bins=mels, **kw)
    z, _, _ = nx.stft(xd, wd, **kw)
    chain = nx.stft_to_mel(z, sr, fft_length=nfft, mel_bins=mels)
    assert fused.shape == chain.shape == (3, z.shape[1], mels)
    # same arithmetic in both paths up to the summation order inside a filter
    assert float((fused - chain).abs().max()) <= 2e-6
    zo, _, _ = o.stft_fast(x, w, **kw)
    want = np.stack([o.stft_to_mel(zo[c], sr, nfft, mels) for c in range(3)])
    np.testing.assert_allclose(fused.cpu().numpy(), want, atol=TOL, rtol=0)


@pytest.mark.parametrize("nfft,hop,mels,sr", [(1024, 256, 80, 16000), (1024, 256, 128, 16000), (512, 128, 128, 8000),
                                              (2048, 512, 128, 22050), (4096, 1024, 64, 11025), (1024, 256, 20, 16000),
                                              (512, 256, 200, 16000)])
def test_fused_stft_mel_full_band_banks(nfft, hop, mels, sr):
    """Sampling rates at which the filters cover every FFT bin (at 48 kHz only the lowest third is
    touched): every thread of the bin-major epilogue owns weights, filters span many threads."""
    import torch

    x = synth((2, 50 * nfft + 77), nfft + mels, fs=sr)
    w = o.hann(nfft)
    kw = dict(overlap_length=nfft - hop, fft_length=nfft, sampling_rate=sr)
    xd, wd = torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda()
    fused = nx.stft_mel(xd, wd, mel_bins=mels, **kw)
    z, _, _ = nx.stft(xd, wd, **kw)
    chain = nx.stft_to_mel(z, sr, fft_length=nfft, mel_bins=mels)
    assert float((fused - chain).abs().max()) <= 2e-6
    zo, _, _ = o.stft_fast(x, w, **kw)
    want = np.stack([o.stft_to_mel(zo[c], sr, nfft, mels) for c in range(2)])
    np.testing.assert_allclose(fused.cpu().numpy(), want, atol=TOL, rtol=0)


@pytest.mark.parametrize("nfft,hop,mels,padding", [(1024, 256, 128, "valid"), (512, 128, 40, "reflect"), (1024, 250, 80, "valid"),
                                                   (400, 160, 40, "valid"), (1024, 64, 300, "same")])
def test_stft_mel_host_entry_equals_device_entry(nfft, hop, mels, padding):
    """numpy in / numpy out through nxs_stft_mel_f32_host: the fused kernel where it applies, the
    chained kernels (inside the C call) elsewhere; same values as the device entries."""
    import torch

    sr = 16000
    x = synth((3, 30 * nfft + 19), nfft + hop + mels, fs=sr)
    w = o.hann(nfft)
    kw = dict(overlap_length=nfft - hop, fft_length=nfft, sampling_rate=sr, window_padding=padding, mel_bins=mels)
    host = nx.stft_mel(x, w, **kw)
    assert isinstance(host, np.ndarray) and host.dtype == np.float32
    dev = nx.stft_mel(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), **kw).cpu().numpy()
    assert host.shape == dev.shape
    assert np.abs(host - dev).max() <= 2e-6
    kw.pop("mel_bins")
    zo, _, _ = o.stft_fast(x, w, **kw) if nfft & (nfft - 1) == 0 else o.stft(x, w, **kw)
    want = np.stack([o.stft_to_mel(zo[c], sr, nfft, mels) for c in range(3)])
    np.testing.assert_allclose(host, want, atol=TOL, rtol=0)


def test_fused_stft_mel_at_scale_channels_have_own_maximum():
    import torch

    C, L, nfft, hop = 6, 48000 * 30, 1024, 256
    x = torch.from_numpy(synth((C, L), 21)).cuda()
    x[1] *= 1e-3   # a quiet channel: its clamp floor must come from its own maximum
    x[4, : L // 2] = 0  # half silent: those frames sit on the floor
    w = torch.from_numpy(o.hann(nfft)).cuda()
    kw = dict(overlap_length=nfft - hop, fft_length=nfft, sampling_rate=48000)
    fused = nx.stft_mel(x, w, mel_bins=128, **kw)
    z, _, _ = nx.stft(x, w, onesided=True, **kw)
    chain = nx.stft_to_mel(z, 48000, fft_length=nfft, mel_bins=128)
    assert float((fused - chain).abs().max()) <= 2e-6
    assert float(fused[4, : 1000].max()) == float(fused[4].min())
