"""Every STFT kernel variant selectable by NXS_STFT_VARIANT gives the same results as the default
(parity of tuning variants; the default is what ships).  Variants 9 / 10 / 14 / 15 switch the packed fp32x2
arithmetic of the FFT engine (Plan::PK) off or on against the default of their size."""
import numpy as np
import pytest

import nx_signal_b200 as nx
from oracle import nxsignal_oracle as o
from tests.util import TOL, frame_rel_err, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nfft,variant", [(64, "0"), (64, "16"), (128, "0"), (128, "9"), (128, "18"), (256, "0"), (256, "9"),
                                          (256, "16"), (256, "17"), (256, "18"), (512, "0"), (512, "9")] + [(1024, v) for v in "012345678"] + [(1024, "14")] +
                         [(2048, v) for v in ["0", "1", "2", "3", "9"]] +
                         [(4096, v) for v in ["0", "1", "2", "3", "4", "6", "7", "8", "9", "10"]] +
                         [(8192, v) for v in ["0", "1", "15"]])
@pytest.mark.parametrize("padding", ["valid", "reflect"])
def test_variant_parity(nfft, variant, padding, monkeypatch):
    monkeypatch.setenv("NXS_STFT_VARIANT", variant)
    x = synth((3, 40 * nfft + 808), 31)
    w = o.hann(nfft)
    kw = dict(overlap_length=nfft - nfft // 4, fft_length=nfft, sampling_rate=48000, window_padding=padding)
    z, _, _ = nx.stft(x, w, **kw)
    zo, _, _ = o.stft_fast(x, w, **kw)
    assert frame_rel_err(z, zo) <= TOL


@pytest.mark.parametrize("nfft", [512, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("hop_num,hop_den", [(1, 1), (1, 2), (3, 4), (1, 16)])
def test_staged_hops(nfft, hop_num, hop_den):
    """Per-group staging serves any hop that keeps frame starts 16-byte aligned, including no overlap."""
    hop = nfft * hop_num // hop_den
    x = synth((2, 23 * nfft + 4 * 77), 41 + nfft)
    w = o.hann(nfft)
    kw = dict(overlap_length=nfft - hop, fft_length=nfft, sampling_rate=48000)
    z, _, _ = nx.stft(x, w, **kw)
    zo, _, _ = o.stft_fast(x, w, **kw)
    assert frame_rel_err(z, zo) <= TOL


@pytest.mark.parametrize("nfft", [512, 1024, 2048, 4096])
@pytest.mark.parametrize("hop", [250, 441, 333, 6, 1])
@pytest.mark.parametrize("padding", ["valid", "same"])
def test_staged_any_hop_and_padding_offset(nfft, hop, padding):
    """Frames that start at any sample offset are still TMA-staged: the copy starts at the 16-byte
    boundary below the frame (odd offsets read the stage with 4-byte loads).  `:same` padding makes
    pad_lo odd (nfft/2 - 1), shifting every frame start."""
    L = 30 * nfft + 123 if hop > 6 else 3 * nfft + 57
    x = synth((2, L + (-L) % 4), 51 + nfft + hop)   # row stride a multiple of 4 samples -> staged path
    w = o.hann(nfft)
    kw = dict(overlap_length=nfft - hop, fft_length=nfft, sampling_rate=48000, window_padding=padding)
    z, _, _ = nx.stft(x, w, **kw)
    zo, _, _ = o.stft_fast(x, w, **kw)
    assert z.shape == zo.shape
    assert frame_rel_err(z, zo) <= TOL
