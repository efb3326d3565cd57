"""Every STFT kernel variant selectable by NXS_STFT_VARIANT gives the same results as the default
(parity of tuning variants; the default is what ships)."""
import os

import numpy as np
import pytest

import nx_signal_b200 as nx
from oracle import nxsignal_oracle as o
from tests.util import TOL, frame_rel_err, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", ["0", "1", "2", "3"])
@pytest.mark.parametrize("padding", ["valid", "reflect"])
def test_variant_parity(variant, padding, monkeypatch):
    monkeypatch.setenv("NXS_STFT_VARIANT", variant)
    x = synth((3, 50_000), 31)
    w = o.hann(1024)
    kw = dict(overlap_length=768, fft_length=1024, sampling_rate=48000, window_padding=padding)
    z, _, _ = nx.stft(x, w, **kw)
    zo, _, _ = o.stft_fast(x, w, **kw)
    assert frame_rel_err(z, zo) <= TOL
