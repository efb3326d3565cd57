"""GPU parity of FIR filtering (overlap-save kernel) and the general convolution kernel against
the oracle, through the reference-shaped API (Convolution.convolve / correlate / fftconvolve).
FIR tolerance: max|gpu - oracle| / max|oracle| <= 1e-5 per channel; small integer-valued cases
from the reference's tests are exact."""
import numpy as np
import pytest

import nx_signal_b200 as nx
from oracle import nxsignal_oracle as o
from tests.util import TOL, synth

pytestmark = pytest.mark.gpu
conv = nx.convolution


def rel(got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    den = np.abs(want).max(axis=-1)
    return float((np.abs(got - want).max(axis=-1) / np.where(den > 0, den, 1.0)).max())


def ref_conv(x, taps, mode):
    """double-precision linear convolution of every row, rounded once to f32, mode by the
    reference's offsets (convolution.ex:300-329)."""
    from scipy.signal import oaconvolve

    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    full = oaconvolve(x, np.asarray(taps, dtype=np.float64)[None, :], mode="full", axes=-1)
    L, K = x.shape[-1], len(taps)
    if mode == "full":
        return full.astype(np.float32)
    if mode == "same":
        s = (K - 1) // 2
        return full[:, s:s + L].astype(np.float32)
    s = min(L, K) - 1
    return full[:, s:s + abs(L - K) + 1].astype(np.float32)


# ---- reference's own small vectors through the GPU (test/nx_signal/convolutions_test.exs) --------
def test_reference_small_vectors():
    eq = np.testing.assert_array_equal
    f = lambda v: np.array(v, dtype=np.float32)
    eq(conv.convolve([1, 2, 3], [3, 4, 5]), f([3, 10, 22, 22, 15]))                  # convolution.ex:32-36
    eq(conv.correlate([1, 2, 3], [3, 4, 5]), f([5, 14, 26, 18, 9]))                  # :81-85
    np.testing.assert_allclose(conv.fftconvolve([1, 2, 3], [3, 4, 5]), f([3, 10, 22, 22, 15]), atol=1e-5)
    eq(conv.convolve([3, 4, 5, 6, 5, 4], [1, 2, 3]), f([3, 10, 22, 28, 32, 32, 23, 12]))
    eq(conv.convolve([3, 4, 5], [1, 2, 3, 4], mode="same"), f([10, 22, 34]))
    eq(conv.convolve([3, 4, 5], [1, 2, 3], mode="same"), f([10, 22, 22]))
    c = conv.convolve(np.array([1 + 1j, 2 + 1j, 3 + 1j]), np.array([1 + 1j, 2 + 1j]))
    eq(c, np.array([2j, 2 + 6j, 5 + 8j, 5 + 5j], dtype=np.complex64))
    eq(conv.convolve(np.array(1289), np.array(4567)), np.array(1289 * 4567, dtype=np.float32))
    eq(conv.convolve([[1, 2, 3], [3, 4, 5]], [[2, 3, 4], [4, 5, 6]]),
       f([[2, 7, 16, 17, 12], [10, 30, 62, 58, 38], [12, 31, 58, 49, 30]]))
    a, b = [1, 2, 3, 6, 5, 3], [2, 3, 4, 5, 3, 4, 2, 2, 1]
    eq(conv.convolve(a, b, mode="valid"), f([70, 78, 73, 65]))
    eq(conv.convolve(b, a, mode="valid"), f([70, 78, 73, 65]))
    eq(conv.convolve([1, 2, 3, 3, 1, 2], [1, 4, 3, 4, 5, 6, 7, 4, 3, 2, 1, 1, 3], mode="same"), f([57, 61, 63, 57, 45, 36]))
    e = [[2, 3, 4, 5, 6, 7, 8], [4, 5, 6, 7, 8, 9, 10]]
    g = [[1, 2, 3], [3, 4, 5]]
    eq(conv.convolve(e, g, mode="valid"), f([[62, 80, 98, 116, 134]]))
    eq(conv.convolve(g, e, mode="valid"), f([[62, 80, 98, 116, 134]]))


def test_broadcastable_batched():  # convolutions_test.exs:95-143: {3,3,3} * {1,1,3} etc.
    a = np.arange(27).reshape(3, 3, 3)
    for shape in [(1, 1, 3), (1, 3, 1), (3, 1, 1)]:
        b = np.arange(3).reshape(shape)
        want = o.convolve(a, b)
        for method in ("direct", "fft"):
            np.testing.assert_array_equal(conv.convolve(a, b, method=method), want)


def test_input_swapping_complex_3d():  # convolutions_test.exs:164-290
    small = np.arange(8).reshape(2, 2, 2)
    big = 1j * np.arange(27).reshape(3, 3, 3) + np.arange(27)[::-1].reshape(3, 3, 3)
    for mode in ("full", "same", "valid"):
        np.testing.assert_array_equal(conv.convolve(small, big, mode=mode), o.convolve(small, big, mode=mode))
        np.testing.assert_array_equal(conv.convolve(big, small, mode=mode), o.convolve(big, small, mode=mode))


def test_correlate_rank1():  # convolutions_test.exs:563-592
    a = np.array([0, 1, 2, 3], np.float32)
    b = np.array([1, 2], np.float32)
    y = np.array([0, 2, 5, 8, 3], np.float32)
    np.testing.assert_array_equal(conv.correlate(a, b, mode="full"), y)
    np.testing.assert_array_equal(conv.correlate(a, b, mode="valid"), y[1:4])
    np.testing.assert_array_equal(conv.correlate(b, a, mode="valid"), y[1:4][::-1])
    np.testing.assert_array_equal(conv.correlate(a, b, mode="same"), y[:-1])


def test_dont_complexify():  # convolutions_test.exs:392-416
    a, b = np.array([1, 2, 3]), np.array([4, 5, 6])
    for t1 in (np.float32, np.complex64):
        for t2 in (np.float32, np.complex64):
            d = conv.convolve(a.astype(t1), b.astype(t2), method="direct")
            f = conv.convolve(a.astype(t1), b.astype(t2), method="fft")
            want = np.complex64 if np.complex64 in (t1, t2) else np.float32
            assert d.dtype == want and f.dtype == want
            np.testing.assert_allclose(d, f, atol=1e-5)


# ---- the overlap-save kernel --------------------------------------------------------------------
@pytest.mark.parametrize("K", [16, 33, 129, 130, 257, 513, 514, 1000, 2049, 3585])
@pytest.mark.parametrize("mode", ["full", "same", "valid"])
def test_fir_overlap_save_vs_double(K, mode):
    x = synth((3, 40_000 + K), 40 + K)
    rng = np.random.default_rng(K)
    taps = (rng.standard_normal(K) / np.sqrt(K)).astype(np.float32)
    y = conv.convolve(x, taps[None, :], mode=mode, method="fft")
    want = ref_conv(x, taps, mode)
    assert y.dtype == np.float32
    assert rel(y, want) <= TOL


@pytest.mark.parametrize("K,L", [(5, 1000), (15, 77), (4000, 20_000), (64, 64), (64, 65), (300, 301)])
def test_fir_direct_and_tight_lengths(K, L):
    x = synth((2, L), K + L)
    taps = o.firwin(K if K % 2 else K + 1, [0.2])[:K]
    for mode in ("full", "same", "valid"):
        y = conv.convolve(x, taps[None, :], mode=mode)
        assert rel(y, ref_conv(x, taps, mode)) <= TOL


def test_fir_unit_impulse_returns_taps():
    taps = nx.filters.firwin(2049, [6000], sampling_rate=48000)
    x = np.zeros((2, 10_000), np.float32)
    x[0, 0] = 1.0
    x[1, 5000] = 2.0
    y = conv.convolve(x, taps[None, :], mode="full", method="fft")
    assert np.abs(y[0, :2049] - taps).max() <= 1e-6
    assert np.abs(y[1, 5000:7049] - 2 * taps).max() <= 2e-6
    assert np.abs(y[0, 2049:]).max() <= 1e-6


def test_cfg4_reduced_and_direct_equals_fft_oracle():
    """BASELINE config 4 at reduced length: 2049-tap firwin lowpass, mode :same; and the literal
    oracle (the reference's two methods) agrees on a short prefix."""
    taps = nx.filters.firwin(2049, [6000], sampling_rate=48000)
    x = synth((4, 48000 * 2), 1004)
    y = conv.convolve(x, taps[None, :], mode="same", method="fft")
    assert y.shape == x.shape
    assert rel(y, o.fir_same_f64(x, taps)) <= TOL
    xs = x[:1, :3000]
    lit = o.convolve(xs, taps[None, :], mode="same", method="direct")
    got = conv.convolve(xs, taps[None, :], mode="same", method="direct")
    assert rel(got, lit) <= TOL


def test_fir_linearity_and_device_entry_at_scale():
    import torch

    taps = torch.from_numpy(nx.filters.firwin(2049, [6000], sampling_rate=48000)).cuda()
    g = torch.Generator(device="cuda").manual_seed(4)
    a = torch.randn(8, 2_000_000, device="cuda", generator=g)
    b = torch.randn(8, 2_000_000, device="cuda", generator=g)
    ya = conv.convolve(a, taps[None, :], mode="same", method="fft")
    yb = conv.convolve(b, taps[None, :], mode="same", method="fft")
    yab = conv.convolve(a - 2 * b, taps[None, :], mode="same", method="fft")
    assert ya.is_cuda and ya.shape == a.shape
    assert float((yab - (ya - 2 * yb)).abs().max() / yab.abs().max()) <= 5e-6
    # spot check against double on one channel's window
    want = ref_conv(a[0:1, :50_000].cpu().numpy(), taps.cpu().numpy(), "same")[:, :40_000]
    assert rel(ya[0:1, :40_000].cpu().numpy(), want) <= TOL


# ---- per-group (TMA-staged) overlap-save kernel: alignment cases, variants, old-kernel agreement ----
@pytest.mark.parametrize("K", [130, 513, 514, 2049, 3585])
@pytest.mark.parametrize("L", [65_536, 65_537, 50_002])
@pytest.mark.parametrize("variant", ["0", "1", "3", "4", "7"])
def test_fir_pg_alignment_and_variants(K, L, variant, monkeypatch):
    """L % 4 == 0 rows are staged by TMA (spans start at arbitrary sample offsets: the kernel floors
    them to 16 bytes), other row strides take the per-thread load path of the same kernel."""
    monkeypatch.setenv("NXS_FIR_VARIANT", variant)
    x = synth((3, L), K + L)
    rng = np.random.default_rng(K + 1)
    taps = (rng.standard_normal(K) / np.sqrt(K)).astype(np.float32)
    for mode in ("same", "full", "valid"):
        y = conv.convolve(x, taps[None, :], mode=mode, method="fft")
        assert rel(y, ref_conv(x, taps, mode)) <= TOL


def test_fir_pg_matches_previous_kernel(monkeypatch):
    import torch

    taps = torch.from_numpy(nx.filters.firwin(2049, [6000], sampling_rate=48000)).cuda()
    x = torch.randn(5, 1_000_000, device="cuda", generator=torch.Generator(device="cuda").manual_seed(9))
    y1 = conv.convolve(x, taps[None, :], mode="same", method="fft")
    torch.cuda.synchronize()
    monkeypatch.setenv("NXS_FIR_NO_PG", "1")
    y2 = conv.convolve(x, taps[None, :], mode="same", method="fft")
    torch.cuda.synchronize()
    assert float((y1 - y2).abs().max() / y2.abs().max()) <= 2e-6


@pytest.mark.parametrize("K", [2731, 2732, 3585, 3586, 4097, 4100, 6000, 10_000])
@pytest.mark.parametrize("mode", ["full", "same", "valid"])
def test_long_filters_by_partition(K, mode):
    """K > 3585: the taps are cut into runs of 2049 and the partial convolutions accumulated
    (y[n] += (x * h_p)[n - 2049 p]); 4100 leaves a 2-tap last partition."""
    x = synth((2, 30_000 + 4 * (K % 7)), 60 + K)
    rng = np.random.default_rng(K)
    taps = (rng.standard_normal(K) / np.sqrt(K)).astype(np.float32)
    y = conv.convolve(x, taps[None, :], mode=mode, method="fft")
    assert rel(y, ref_conv(x, taps, mode)) <= TOL


def test_long_filter_longer_than_signal():
    x = synth((2, 3000), 5)
    taps = (np.random.default_rng(1).standard_normal(9001) / 95).astype(np.float32)
    for mode in ("full", "same"):  # :valid needs one operand to cover the other in every dimension (convolution.ex:131-134)
        y = conv.convolve(x, taps[None, :], mode=mode, method="fft")
        assert rel(y, ref_conv(x, taps, mode)) <= TOL
