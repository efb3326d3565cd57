"""Pins oracle/nxsignal_oracle.py against every doctest / known-answer vector the
reference holds for the STFT / ISTFT / windows / FIR path.

Doctest vectors are printed f32 values (shortest round-trip repr), so parsing the
printed decimal as f32 recovers the exact bits: comparisons are bit-exact unless
stated otherwise.
"""
import numpy as np
import pytest

from oracle import nxsignal_oracle as o

F32 = np.float32


def f32(xs):
    return np.array(xs, dtype=np.float64).astype(F32)


def bit_equal(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    # +0.0 == -0.0 is acceptable: the reference prints both as 0.0 / -0.0 explicitly
    assert np.array_equal(a, b), f"\n{a!r}\n!=\n{b!r}"


# ---- windows (lib/nx_signal/windows.ex) -------------------------------------
def test_rectangular():  # windows.ex:20-30
    bit_equal(o.rectangular(5), np.ones(5, dtype=np.int64))
    bit_equal(o.rectangular(5, dtype=F32), np.ones(5, dtype=F32))


def test_bartlett():  # windows.ex:50-54
    bit_equal(o.bartlett(3), f32([0.0, 0.6666667, 0.6666666]))


def test_triangular():  # windows.ex:91-95
    bit_equal(o.triangular(3), f32([0.5, 1.0, 0.5]))


def test_blackman():  # windows.ex:141-157
    bit_equal(
        o.blackman(5, is_periodic=False),
        f32([-1.4901161e-8, 0.34000003, 0.99999994, 0.34000003, -1.4901161e-8]),
    )
    bit_equal(
        o.blackman(5, is_periodic=True),
        f32([-1.4901161e-8, 0.20077012, 0.84922993, 0.84922993, 0.20077012]),
    )
    bit_equal(
        o.blackman(6, is_periodic=True), f32([-1.4901161e-8, 0.13, 0.63, 0.99999994, 0.63, 0.13])
    )


def test_hamming():  # windows.ex:213-222
    bit_equal(
        o.hamming(5, is_periodic=True), f32([0.08000001, 0.3978522, 0.9121479, 0.9121478, 0.39785212])
    )
    bit_equal(o.hamming(5, is_periodic=False), f32([0.08000001, 0.54, 1.0, 0.54, 0.08000001]))


def test_hann():  # windows.ex:266-275
    bit_equal(o.hann(5, is_periodic=False), f32([0.0, 0.5, 1.0, 0.5, 0.0]))
    bit_equal(
        o.hann(5, is_periodic=True), f32([0.0, 0.34549153, 0.90450853, 0.9045085, 0.34549144])
    )


def test_kaiser():  # windows.ex:322-338
    bit_equal(o.kaiser(4, beta=12.0, is_periodic=True), f32([5.277619e-5, 0.21566667, 1.0, 0.21566667]))
    bit_equal(
        o.kaiser(5, beta=12.0, is_periodic=True),
        f32([5.277619e-5, 0.10171464, 0.792937, 0.792937, 0.10171464]),
    )
    bit_equal(
        o.kaiser(4, beta=12.0, is_periodic=False), f32([5.277619e-5, 0.5188395, 0.51883906, 5.277619e-5])
    )


# ---- stft / framing / frequencies (lib/nx_signal.ex) ---------------------------
def test_stft_doctest():  # lib/nx_signal.ex:46-65
    z, t, f = o.stft(
        np.arange(4, dtype=np.int32), o.rectangular(2), overlap_length=1, fft_length=2, sampling_rate=400
    )
    bit_equal(z, np.array([[1, -1], [3, -1], [5, -1]], dtype=np.complex64))
    bit_equal(t, f32([0.0025, 0.005, 0.0075]))
    bit_equal(f, f32([0.0, 200.0]))


def test_fft_frequencies():  # lib/nx_signal.ex:147-151
    bit_equal(
        o.fft_frequencies(1.6e4, 10),
        f32([0.0, 1.6e3, 3.2e3, 4.8e3, 6.4e3, 8e3, 9.6e3, 1.12e4, 1.28e4, 1.44e4]),
    )


def test_as_windowed_doctests():  # lib/nx_signal.ex:182-246
    x = np.array([0, 1, 2, 3, 4, 10, 11, 12], dtype=np.int32)
    bit_equal(
        o.as_windowed(x, 4),
        np.array([[0, 1, 2, 3], [1, 2, 3, 4], [2, 3, 4, 10], [3, 4, 10, 11], [4, 10, 11, 12]], dtype=np.int32),
    )
    bit_equal(
        o.as_windowed(x, 3),
        np.array(
            [[0, 1, 2], [1, 2, 3], [2, 3, 4], [3, 4, 10], [4, 10, 11], [10, 11, 12]], dtype=np.int32
        ),
    )
    bit_equal(
        o.as_windowed(np.array([0, 1, 2, 3, 4, 10, 11], dtype=np.int32), 2, stride=2, padding=[(0, 3)]),
        np.array([[0, 1], [2, 3], [4, 10], [11, 0], [0, 0]], dtype=np.int32),
    )
    bit_equal(
        o.as_windowed(np.arange(7, dtype=np.int32), 6, padding="reflect", stride=1),
        np.array(
            [
                [3, 2, 1, 0, 1, 2],
                [2, 1, 0, 1, 2, 3],
                [1, 0, 1, 2, 3, 4],
                [0, 1, 2, 3, 4, 5],
                [1, 2, 3, 4, 5, 6],
                [2, 3, 4, 5, 6, 5],
                [3, 4, 5, 6, 5, 4],
                [4, 5, 6, 5, 4, 3],
            ],
            dtype=np.int32,
        ),
    )
    bit_equal(
        o.as_windowed(np.arange(10, dtype=np.int32), 6, padding="reflect", stride=2),
        np.array(
            [
                [3, 2, 1, 0, 1, 2],
                [1, 0, 1, 2, 3, 4],
                [1, 2, 3, 4, 5, 6],
                [3, 4, 5, 6, 7, 8],
                [5, 6, 7, 8, 9, 8],
                [7, 8, 9, 8, 7, 6],
            ],
            dtype=np.int32,
        ),
    )


def test_as_windowed_errors():  # lib/nx_signal.ex:282-284, 319-329
    with pytest.raises(ValueError, match="expected an integer >= 1"):
        o.as_windowed(np.arange(4), 2, stride=0)
    with pytest.raises(ValueError, match="invalid padding mode"):
        o.as_windowed(np.arange(4), 2, padding="zeros")
    with pytest.raises(ValueError, match="padding must be a list"):
        o.as_windowed(np.arange(4), 2, padding=[(0.5, 1)])


def test_stft_reflect_nfft16_via_mel():  # lib/nx_signal.ex:465-483
    z, _, _ = o.stft(
        np.arange(10, dtype=np.int32),
        o.hann(4),
        overlap_length=2,
        fft_length=16,
        sampling_rate=8.0e3,
        window_padding="reflect",
    )
    assert z.shape == (6, 16)
    mel = o.stft_to_mel(z, 8.0e3, 16, 4)
    want = f32(
        [
            [0.29005307, 0.17422175, 0.18422472, 0.09807998],
            [0.6093881, 0.5647397, 0.43538243, 0.086352706],
            [0.75841033, 0.70850146, 0.5636921, 0.17911881],
            [0.8461772, 0.7952491, 0.64707625, 0.25204098],
            [0.9085489, 0.85726047, 0.70786566, 0.30867678],
            [0.9085489, 0.85726047, 0.70786566, 0.30867678],
        ]
    )
    bit_equal(mel, want)


def test_mel_filters_row():  # lib/nx_signal.ex:384-394 (pins Nx.linspace's f32 op order)
    m = o.mel_filters(10, 5, 8.0e3)
    want = f32([0.0, 0.0, 0.0, 0.0, 7.329034e-5, 2.3422057e-4, 3.8295105e-4, 2.871204e-4, 1.9128979e-4, 9.545916e-5])
    bit_equal(m[4], want)
    bit_equal(m[0][:3], f32([0.0, 8.129208e-4, 0.0]))
    # row 3 sits in the log-spaced region (outside the hot path): tolerance only
    np.testing.assert_allclose(m[3], f32([0.0, 0.0, 0.0, 4.035892e-4, 5.276656e-4, 2.574124e-4, 0.0, 0.0, 0.0, 0.0]), rtol=1e-6)


# ---- istft / overlap_and_add ---------------------------------------------------
@pytest.mark.parametrize(
    "scaling,want",
    [
        (None, [0, 10, 1, 0, 10, 10, 2, 20]),
        ("spectrum", [0, 10, 1, 0, 10, 10, 2, 20]),
    ],
)
def test_istft_roundtrip_int(scaling, want):  # lib/nx_signal.ex:545-568
    t = np.array([10, 10, 1, 0, 10, 10, 2, 20], dtype=np.int32)
    w = o.hann(4)
    z, _, _ = o.stft(t, w, sampling_rate=1, fft_length=4, scaling=scaling)
    r = o.istft(z, w, sampling_rate=1, fft_length=4, scaling=scaling)
    assert r.dtype == np.complex64
    # Nx.as_type(c64 -> s32) takes the real part and truncates
    bit_equal(np.trunc(r.real).astype(np.int32), np.array(want, dtype=np.int32))


def test_istft_roundtrip_psd():  # lib/nx_signal.ex:570-579
    t = np.array([10, 10, 1, 0, 10, 10, 2, 20], dtype=F32)
    w = o.hann(4)
    z, _, _ = o.stft(t, w, sampling_rate=1, fft_length=4, scaling="psd")
    r = o.istft(z, w, sampling_rate=1, fft_length=4, scaling="psd")
    bit_equal(r.real, f32([0.0, 10.0, 0.99999994, -2.1900146e-7, 10.0, 10.0, 2.0000002, 20.0]))


def test_overlap_and_add_doctests():  # lib/nx_signal.ex:656-681
    x = np.arange(12, dtype=np.int32).reshape(3, 4)
    bit_equal(o.overlap_and_add(x, 0), np.arange(12, dtype=np.int32))
    bit_equal(o.overlap_and_add(x, 3), np.array([0, 5, 15, 18, 17, 11], dtype=np.int32))
    t = np.array(
        [[[[0, 1, 2, 3], [4, 5, 6, 7]]], [[[10, 11, 12, 13], [14, 15, 16, 17]]]], dtype=np.int32
    )
    bit_equal(
        o.overlap_and_add(t, 3), np.array([[[0, 5, 7, 9, 7]], [[10, 25, 27, 29, 17]]], dtype=np.int32)
    )
    with pytest.raises(ValueError, match="overlap_length must be a number less than"):
        o.overlap_and_add(x, 4)


# ---- firwin (test/nx_signal/filters_test.exs:245-416) --------------------------
FIRWIN_CASES = [
    (dict(num_taps=5, cutoff=[0.3]),
     [0.020103708268285354, 0.23086668180542194, 0.4980592198525855, 0.23086668180542194, 0.020103708268285354], 1e-5),
    (dict(num_taps=7, cutoff=[0.4], pass_zero=False),
     [0.004998140998601554, -0.02905169455437149, -0.23351680322070983, 0.6010660646645265,
      -0.2335168032207099, -0.02905169455437152, 0.004998140998601554], 1e-5),
    (dict(num_taps=9, cutoff=[0.2, 0.6], pass_zero=False, window="hann"),
     [0.0, -0.034265228115753485, -0.17548320982592003, 0.14143709641554006, 0.5732069654682745,
      0.14143709641554006, -0.17548320982592003, -0.034265228115753485, 0.0], 1e-5),
    (dict(num_taps=11, cutoff=[0.3, 0.7], window="blackman"),
     [0.0, -0.004174601858029537, 0.0, 0.17126025417159732, 0.0, 0.6658286953728643, 0.0,
      0.17126025417159732, 0.0, -0.004174601858029537, 0.0], 1e-5),
    (dict(num_taps=7, cutoff=[0.5], window=("kaiser", 5.0)),
     [-0.003951274147023466, 0.0, 0.25034887446528337, 0.5072047993634803, 0.25034887446528337, 0.0,
      -0.003951274147023466], 1e-3),
    (dict(num_taps=7, cutoff=[0.4], window="rectangular"),
     [-0.058404528708691714, 0.08760679306303756, 0.28350153764274655, 0.37459239600581506,
      0.28350153764274655, 0.08760679306303756, -0.058404528708691714], 1e-5),
    (dict(num_taps=5, cutoff=[0.3], scale=False),
     [0.012109227658250522, 0.13905977799613067, 0.3, 0.13905977799613067, 0.012109227658250522], 1e-5),
    (dict(num_taps=5, cutoff=[1000], sampling_rate=8000),
     [0.024553834015016568, 0.23438946423798604, 0.48211340349399473, 0.23438946423798604,
      0.024553834015016568], 1e-5),
]


@pytest.mark.parametrize("kw,want,atol", FIRWIN_CASES)
def test_firwin_known_answers(kw, want, atol):
    h = o.firwin(**kw)
    assert h.dtype == F32
    # the reference's own tolerance: assert_all_close(atol given, rtol default 1e-4)
    np.testing.assert_allclose(h, np.array(want), atol=atol, rtol=1e-4)


def test_firwin_cross_check_scipy():
    from scipy.signal import firwin as sp_firwin

    h = o.firwin(2049, [6000], sampling_rate=48000)
    ref = sp_firwin(2049, 6000, fs=48000)
    assert np.abs(h - ref).max() < 2e-7


def test_firwin_errors():  # filters_test.exs:396-416
    with pytest.raises(ValueError, match="cutoff must be strictly between 0 and Nyquist"):
        o.firwin(5, [1.0])
    with pytest.raises(ValueError, match="cutoff must be strictly between 0 and Nyquist"):
        o.firwin(5, [0.0])
    with pytest.raises(ValueError, match="odd number of taps"):
        o.firwin(6, [0.4], pass_zero=False)
    with pytest.raises(ValueError, match="unknown window"):
        o.firwin(5, [0.3], window="bogus")
    with pytest.raises(ValueError, match="cutoff must be a list"):
        o.firwin(5, 0.3)


def test_sinc_doctest():  # waveforms.ex:439-445
    bit_equal(o.sinc(f32([0, 0.25, 1])), f32([1.0, 0.9003163, -2.7827534e-8]))


# ---- convolution (lib/nx_signal/convolution.ex, test/nx_signal/convolutions_test.exs)
def close(a, b, atol=1e-4, rtol=1e-4):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), atol=atol, rtol=rtol)


def test_convolve_doctests():  # convolution.ex:32-36, 81-85, 246-250
    bit_equal(o.convolve([1, 2, 3], [3, 4, 5]), f32([3.0, 10.0, 22.0, 22.0, 15.0]))
    bit_equal(o.correlate([1, 2, 3], [3, 4, 5]), f32([5.0, 14.0, 26.0, 18.0, 9.0]))
    bit_equal(o.fftconvolve([1, 2, 3], [3, 4, 5]), f32([3.0000007, 10.0, 22.0, 22.0, 15.0]))


def test_convolve_basic_same():  # convolutions_test.exs:7-35
    c = o.convolve(np.ones(100, F32), np.ones(3, F32))[2:-2]
    close(c, np.full(98, 3.0))
    bit_equal(o.convolve([3, 4, 5, 6, 5, 4], [1, 2, 3]), f32([3, 10, 22, 28, 32, 32, 23, 12]))
    bit_equal(o.convolve([3, 4, 5], [1, 2, 3, 4], mode="same"), f32([10, 22, 34]))
    bit_equal(o.convolve([3, 4, 5], [1, 2, 3], mode="same"), f32([10, 22, 22]))


def test_convolve_complex_and_scalars():  # convolutions_test.exs:37-63,145-150
    c = o.convolve(np.array([1 + 1j, 2 + 1j, 3 + 1j]), np.array([1 + 1j, 2 + 1j]))
    bit_equal(c, np.array([2j, 2 + 6j, 5 + 8j, 5 + 5j], dtype=np.complex64))
    bit_equal(o.convolve(np.array(1289), np.array(4567)), np.array(1289 * 4567, dtype=F32))
    bit_equal(o.convolve(np.array([1 + 1j]), np.array([3 + 4j])), np.array([-1 + 7j], dtype=np.complex64))
    bit_equal(o.convolve([4967], [3920]), f32([4967 * 3920]))


def test_convolve_broadcastable():  # convolutions_test.exs:95-143
    a = np.arange(27).reshape(3, 3, 3)
    b = np.arange(3).reshape(1, 1, 3)
    e1 = np.array(
        [
            [[0, 0, 1, 4, 4], [0, 3, 10, 13, 10], [0, 6, 19, 22, 16]],
            [[0, 9, 28, 31, 22], [0, 12, 37, 40, 28], [0, 15, 46, 49, 34]],
            [[0, 18, 55, 58, 40], [0, 21, 64, 67, 46], [0, 24, 73, 76, 52]],
        ]
    )
    close(o.convolve(a, b, method="direct"), e1)
    close(o.convolve(a, b, method="fft"), e1)
    b2 = b.reshape(1, 3, 1)
    e2 = np.array(
        [
            [[0, 0, 0], [0, 1, 2], [3, 6, 9], [12, 15, 18], [12, 14, 16]],
            [[0, 0, 0], [9, 10, 11], [30, 33, 36], [39, 42, 45], [30, 32, 34]],
            [[0, 0, 0], [18, 19, 20], [57, 60, 63], [66, 69, 72], [48, 50, 52]],
        ]
    )
    close(o.convolve(a, b2, method="direct"), e2)
    close(o.convolve(a, b2, method="fft"), e2)
    b3 = b.reshape(3, 1, 1)
    e3 = np.array(
        [
            [[0, 0, 0], [0, 0, 0], [0, 0, 0]],
            [[0, 1, 2], [3, 4, 5], [6, 7, 8]],
            [[9, 12, 15], [18, 21, 24], [27, 30, 33]],
            [[36, 39, 42], [45, 48, 51], [54, 57, 60]],
            [[36, 38, 40], [42, 44, 46], [48, 50, 52]],
        ]
    )
    close(o.convolve(a, b3, method="direct"), e3)
    close(o.convolve(a, b3, method="fft"), e3)


def test_convolve_2d():  # convolutions_test.exs:152-162, 444-453
    c = o.convolve([[1, 2, 3], [3, 4, 5]], [[2, 3, 4], [4, 5, 6]])
    bit_equal(c, f32([[2, 7, 16, 17, 12], [10, 30, 62, 58, 38], [12, 31, 58, 49, 30]]))
    e = [[2, 3, 4, 5, 6, 7, 8], [4, 5, 6, 7, 8, 9, 10]]
    f = [[1, 2, 3], [3, 4, 5]]
    h = f32([[62, 80, 98, 116, 134]])
    bit_equal(o.convolve(e, f, mode="valid"), h)
    bit_equal(o.convolve(f, e, mode="valid"), h)


def test_convolve_input_swapping():  # convolutions_test.exs:164-290 (structure; values by definition)
    small = np.arange(8).reshape(2, 2, 2)
    big = 1j * np.arange(27).reshape(3, 3, 3) + np.arange(27)[::-1].reshape(3, 3, 3)
    full = o.convolve(small, big, mode="full")
    assert full.shape == (4, 4, 4) and full.dtype == np.complex64
    # spot values from the reference's table
    assert full[0, 0, 1] == 26 and full[1, 1, 1] == 632 + 96j and full[3, 3, 3] == 182j
    bit_equal(o.convolve(big, small, mode="full"), full)
    bit_equal(o.convolve(small, big, mode="same"), full[1:3, 1:3, 1:3])
    bit_equal(o.convolve(big, small, mode="same"), full[0:3, 0:3, 0:3])
    bit_equal(o.convolve(small, big, mode="valid"), full[1:3, 1:3, 1:3])
    bit_equal(o.convolve(big, small, mode="valid"), full[1:3, 1:3, 1:3])


def test_convolve_valid_same_modes():  # convolutions_test.exs:337-368
    a = [1, 2, 3, 6, 5, 3]
    b = [2, 3, 4, 5, 3, 4, 2, 2, 1]
    bit_equal(o.convolve(a, b, mode="valid"), f32([70, 78, 73, 65]))
    bit_equal(o.convolve(b, a, mode="valid"), f32([70, 78, 73, 65]))
    a = np.array([1 + 5j, 2 - 1j, 3 + 0j])
    b = np.array([2 - 3j, 1 + 0j])
    e = np.array([2 - 3j, 8 - 10j], dtype=np.complex64)
    bit_equal(o.convolve(a, b, mode="valid"), e)
    bit_equal(o.convolve(b, a, mode="valid"), e)
    c = o.convolve([1, 2, 3, 3, 1, 2], [1, 4, 3, 4, 5, 6, 7, 4, 3, 2, 1, 1, 3], mode="same")
    bit_equal(c, f32([57, 61, 63, 57, 45, 36]))


def test_convolve_errors():  # convolutions_test.exs:292-335, 370-390, 418-442
    a, b = [3, 4, 5], [1, 2, 3]
    with pytest.raises(ValueError, match=r"expected mode to be one of \[:full, :same, :valid\], got: :spam"):
        o.convolve(a, b, mode="spam")
    with pytest.raises(ValueError, match="got: :eggs"):
        o.convolve(a, b, mode="eggs", method="fft")
    with pytest.raises(ValueError, match=r"expected method to be one of \[:direct, :fft\], got: :bacon"):
        o.convolve(a, b, mode="full", method="bacon")
    x = np.arange(1, 7).reshape(2, 3)
    y = np.arange(-6, 0).reshape(3, 2)
    with pytest.raises(ValueError):
        o.convolve(x, y, mode="valid")
    with pytest.raises(ValueError):
        o.convolve(y, x, mode="valid")
    for m in ("direct", "fft"):
        with pytest.raises(ValueError):
            o.convolve(np.array([1]), np.array(2), method=m)
        with pytest.raises(ValueError):
            o.convolve(np.array(1), np.array([2]), method=m)
    with pytest.raises(ValueError):
        o.convolve(np.array([1]), np.array([[2]]))


def test_dont_complexify():  # convolutions_test.exs:392-416
    a = np.array([1, 2, 3])
    b = np.array([4, 5, 6])
    for t1 in (F32, np.complex64):
        for t2 in (F32, np.complex64):
            d = o.convolve(a.astype(t1), b.astype(t2), method="direct")
            f = o.convolve(a.astype(t1), b.astype(t2), method="fft")
            close(d, f)
            want = np.complex64 if np.complex64 in (t1, t2) else F32
            assert d.dtype == want and f.dtype == want


def test_fft_method_cases():  # convolutions_test.exs:455-559
    a = np.array([1, 2, 3])
    close(o.convolve(a, a, method="fft"), [1, 4, 10, 12, 9.0])
    ac = np.array([1 + 1j, 2 + 2j, 3 + 3j])
    close(o.convolve(ac, ac, method="fft"), [2j, 8j, 20j, 24j, 18j])
    a2 = np.array([[1, 2, 3], [4, 5, 6]])
    close(o.convolve(a2, a2, method="fft"), [[1, 4, 10, 12, 9], [8, 26, 56, 54, 36], [16, 40, 73, 60, 36]])
    c2 = np.array([[1 + 2j, 3 + 4j, 5 + 6j], [2 + 1j, 4 + 3j, 6 + 5j]])
    e = np.array(
        [
            [-3 + 4j, -10 + 20j, -21 + 56j, -18 + 76j, -11 + 60j],
            [10j, 44j, 118j, 156j, 122j],
            [3 + 4j, 10 + 20j, 21 + 56j, 18 + 76j, 11 + 60j],
        ]
    )
    close(o.convolve(c2, c2, method="fft"), e)
    b = np.array([3, 3, 5, 6, 8, 7, 9, 0, 1])
    close(o.convolve(a, b, method="fft", mode="same"), [35.0, 41.0, 47.0])
    close(o.convolve(b, a, method="fft", mode="same"), [9.0, 20.0, 25.0, 35.0, 41.0, 47.0, 39.0, 28.0, 2.0])
    a3 = np.array([3, 2, 1])
    e = [24.0, 31.0, 41.0, 43.0, 49.0, 25.0, 12.0]
    close(o.convolve(a3, b, method="fft", mode="valid"), e)
    close(o.convolve(b, a3, method="fft", mode="valid"), e)


def test_correlate_rank1():  # convolutions_test.exs:563-592
    a = o.nx_linspace(0, 3, 4)
    b = o.nx_linspace(1, 2, 2)
    y = np.array([0, 2, 5, 8, 3.0])
    close(o.correlate(a, b, mode="valid"), y[1:4])
    close(o.correlate(b, a, mode="valid"), y[1:4][::-1])
    close(o.correlate(a, b, mode="same"), y[:-1])
    close(o.correlate(a, b, mode="full"), y)


def test_fft_nd_cases():  # convolutions_test.exs:65-93, transforms_test.exs:5-42
    a = np.array([[1, 2, 3], [4, 5, 6]])
    c = o.nx_fft(o.nx_fft(a, 2, axis=0), 3, axis=1)
    close(c, [[21, -3 + 1.732j, -3 - 1.732j], [-9, 0, 0]])
    c = o.nx_fft(o.nx_fft(a, 3, axis=0), 3, axis=1)
    z = [
        [2.1e1, -3 + 1.732j, -3 - 1.732j],
        [-1.5 - 12.99j, -1.11e-16 + 1.732j, -1.5 + 0.866j],
        [-1.5 + 12.99j, -1.5 - 0.866j, -1.11e-16 - 1.732j],
    ]
    close(c, z)
    e = np.array([[1, 0], [0, 1]])
    bit_equal(o.nx_fft(o.nx_fft(e, axis=0), axis=1), np.array([[2, 0], [0, 2]], dtype=np.complex64))
    bit_equal(
        o.nx_ifft(o.nx_ifft(np.array([[2, 0], [0, 2]]), axis=0), axis=1),
        np.array([[1, 0], [0, 1]], dtype=np.complex64),
    )


# ---- the fast (pocketfft, complex128) variants equal the literal restatement -----
def test_fast_variants_match_literal():
    rng = np.random.default_rng(7)
    x = rng.standard_normal(4096).astype(F32)
    w = o.hann(256)
    z1, t1, f1 = o.stft(x, w, overlap_length=192, fft_length=256, sampling_rate=48000)
    z2, t2, f2 = o.stft_fast(x, w, overlap_length=192, fft_length=256, sampling_rate=48000)
    scale = np.abs(z1).max()
    assert np.abs(z1 - z2).max() / scale < 2e-7
    y1 = o.istft(z1, w, overlap_length=192, fft_length=256)
    y2 = o.istft_fast(z1, w, overlap_length=192, fft_length=256)
    assert np.abs(y1 - y2).max() / np.abs(y1).max() < 2e-7
