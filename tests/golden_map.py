"""(file, line) of every hot-path doctest vector in tests/golden/reference_doctests.json -> the call
that must reproduce it, written once against an implementation module `m` that exposes the
reference's heads (the oracle, or nx_signal_b200 for the GPU path)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def records():
    with open(os.path.join(HERE, "golden", "reference_doctests.json")) as f:
        return json.load(f)["records"]


def to_array(rec):
    typ = rec["type"]
    v = rec["values"]
    if typ.startswith("c"):
        a = np.asarray(v, dtype=np.float64)
        return (a[..., 0] + 1j * a[..., 1]).astype(np.complex64)
    dt = {"f32": np.float32, "s32": np.int32, "s64": np.int64, "f64": np.float64}[typ]
    return np.asarray(v, dtype=dt)


OUT_OF_SCOPE = {("lib/nx_signal/filters.ex", 71)}  # Filters.wiener (out of scope)


def in_scope(rec):
    if (rec["file"], rec["line"]) in OUT_OF_SCOPE:
        return False
    if rec["file"].endswith("waveforms.ex"):
        return rec["line"] == 445  # only sinc is on the path (firwin)
    return True


def _stft_rect(m):
    return m.stft(np.arange(4, dtype=np.int32), m.windows.rectangular(2), overlap_length=1, fft_length=2, sampling_rate=400)


def _roundtrip(m, scaling, dtype):
    t = np.array([10, 10, 1, 0, 10, 10, 2, 20], dtype=dtype)
    w = m.windows.hann(4)
    kw = dict(sampling_rate=1, fft_length=4, scaling=scaling)
    z, _, _ = m.stft(t, w, **kw)
    return m.istft(z, w, **kw)


def _mel_doctest(m):
    z, _, _ = m.stft(np.arange(10, dtype=np.int32), m.windows.hann(4), overlap_length=2, fft_length=16,
                     sampling_rate=8.0e3, window_padding="reflect")
    return m.stft_to_mel(z, 8.0e3, fft_length=16, mel_bins=4)


# lib/nx_signal.ex:667: Nx.tensor([[[[0, 1, 2, 3], [4, 5, 6, 7]]], [[[10, ...], [14, ...]]]]) |> Nx.vectorize(x: 2, y: 1)
_OLA_VEC = np.array([[[[0, 1, 2, 3], [4, 5, 6, 7]]], [[[10, 11, 12, 13], [14, 15, 16, 17]]]], dtype=np.int32)

W = "lib/nx_signal/windows.ex"
S = "lib/nx_signal.ex"
C = "lib/nx_signal/convolution.ex"

# value: (callable(m) -> array, kind) ; kind "exact" = bit-for-bit in the oracle, integer-exact on the GPU
CALLS = {
    (S, 48): lambda m: _stft_rect(m)[0],
    (S, 57): lambda m: _stft_rect(m)[1],
    (S, 62): lambda m: _stft_rect(m)[2],
    (S, 148): lambda m: m.fft_frequencies(1.6e4, fft_length=10),
    (S, 183): lambda m: m.as_windowed(np.array([0, 1, 2, 3, 4, 10, 11, 12], np.int32), window_length=4),
    (S, 195): lambda m: m.as_windowed(np.array([0, 1, 2, 3, 4, 10, 11, 12], np.int32), window_length=3),
    (S, 208): lambda m: m.as_windowed(np.array([0, 1, 2, 3, 4, 10, 11], np.int32), window_length=2, stride=2,
                                       padding=[(0, 3)]),
    (S, 221): lambda m: m.as_windowed(np.arange(7, dtype=np.int32), window_length=6, padding="reflect", stride=1),
    (S, 236): lambda m: m.as_windowed(np.arange(10, dtype=np.int32), window_length=6, padding="reflect", stride=2),
    (S, 385): lambda m: m.mel_filters(10, 5, 8.0e3),
    (S, 473): _mel_doctest,
    (S, 551): lambda m: np.asarray(_roundtrip(m, None, np.int32)),
    (S, 565): lambda m: np.asarray(_roundtrip(m, "spectrum", np.int32)),
    (S, 576): lambda m: np.asarray(_roundtrip(m, "psd", np.float32)),
    (S, 657): lambda m: m.overlap_and_add(np.arange(12, dtype=np.int32).reshape(3, 4), overlap_length=0),
    (S, 663): lambda m: m.overlap_and_add(np.arange(12, dtype=np.int32).reshape(3, 4), overlap_length=3),
    (S, 670): lambda m: m.overlap_and_add(_OLA_VEC, overlap_length=3),
    (W, 21): lambda m: m.windows.rectangular(5),
    (W, 27): lambda m: m.windows.rectangular(5),
    (W, 51): lambda m: m.windows.bartlett(3),
    (W, 92): lambda m: m.windows.triangular(3),
    (W, 142): lambda m: m.windows.blackman(5, is_periodic=False),
    (W, 148): lambda m: m.windows.blackman(5, is_periodic=True),
    (W, 154): lambda m: m.windows.blackman(6, is_periodic=True),
    (W, 214): lambda m: m.windows.hamming(5, is_periodic=True),
    (W, 219): lambda m: m.windows.hamming(5, is_periodic=False),
    (W, 267): lambda m: m.windows.hann(5, is_periodic=False),
    (W, 272): lambda m: m.windows.hann(5, is_periodic=True),
    (W, 323): lambda m: m.windows.kaiser(4, beta=12.0, is_periodic=True),
    (W, 329): lambda m: m.windows.kaiser(5, beta=12.0, is_periodic=True),
    (W, 335): lambda m: m.windows.kaiser(4, beta=12.0, is_periodic=False),
    (C, 33): lambda m: m.convolution.convolve(np.array([1, 2, 3]), np.array([3, 4, 5])),
    (C, 82): lambda m: m.convolution.correlate(np.array([1, 2, 3]), np.array([3, 4, 5])),
    (C, 247): lambda m: m.convolution.fftconvolve(np.array([1, 2, 3]), np.array([3, 4, 5])),
    ("lib/nx_signal/waveforms.ex", 445): lambda m: m.sinc(np.array([0, 0.25, 1], np.float32)),
}

