"""The oracle's C port (timed as the CPU baseline) equals the numpy oracle (pinned to the
reference's vectors) to within one f32 ulp of the frame maximum -- libm vs numpy cos/sin
may differ in the last double bit."""
import numpy as np
import pytest

from oracle import c_port
from oracle import nxsignal_oracle as o
from tests.util import synth


@pytest.mark.parametrize("nfft,N,hop", [(1024, 1024, 256), (256, 200, 50), (16, 4, 2), (10, 10, 3), (15, 12, 5)])
@pytest.mark.parametrize("scaling", [None, "psd"])
def test_c_port_matches_numpy_oracle(nfft, N, hop, scaling):
    x = synth((2, 5000), 17)
    w = o.hann(N)
    z = c_port.stft(x, w, hop, nfft, scaling=scaling, sampling_rate=8000)
    zo, _, _ = o.stft(x, w, overlap_length=N - hop, fft_length=nfft, sampling_rate=8000, scaling=scaling)
    assert z.shape == zo.shape
    assert np.abs(z - zo).max() <= 1.2e-7 * np.abs(zo).max()
    assert (z == zo).mean() > 0.99


def test_c_port_reflect_doctest_shape():  # lib/nx_signal.ex:465-471
    x = np.arange(10, dtype=np.float32)[None, :]
    z = c_port.stft(x, o.hann(4), 2, 16, pad_lo=2, pad_hi=2, reflect=True, sampling_rate=8000)
    zo, _, _ = o.stft(x, o.hann(4), overlap_length=2, fft_length=16, sampling_rate=8000, window_padding="reflect")
    np.testing.assert_allclose(z, zo, atol=1e-6)
