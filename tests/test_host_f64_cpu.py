"""`type: :f64` of the host-side closed forms: the reference computes windows, firwin and
fft_frequencies in the requested type (windows.ex:58,161,226,279,342; filters.ex:153;
lib/nx_signal.ex:155-165), so the C ABI has f64 entries (nxs_window_f64, nxs_firwin_f64,
nxs_fft_frequencies_ex) evaluated in double.  Checked against the oracle's float_type(F64)
restatement (the reference holds no f64 vector for these heads: unpinned, restatement vs
restatement) and against the f32 values, which must be the f64 values up to f32 rounding noise."""
import numpy as np
import pytest

import nx_signal_b200 as nx
from oracle import nxsignal_oracle as o

WINDOWS = [("bartlett", {}), ("triangular", {}), ("blackman", {"is_periodic": True}), ("blackman", {"is_periodic": False}),
           ("hamming", {"is_periodic": True}), ("hamming", {"is_periodic": False}), ("hann", {"is_periodic": True}),
           ("hann", {"is_periodic": False}), ("kaiser", {"beta": 12.0, "is_periodic": True}),
           ("kaiser", {"beta": 5.0, "is_periodic": False})]


@pytest.mark.parametrize("name,kw", WINDOWS)
@pytest.mark.parametrize("n", [4, 5, 64, 1023])
def test_f64_windows_are_computed_in_double(name, kw, n):
    got = getattr(nx.windows, name)(n, type="f64", **kw)
    assert got.dtype == np.float64
    with o.float_type(o.F64):
        want = getattr(o, name)(n, **kw)
    assert want.dtype == np.float64
    np.testing.assert_allclose(got, want, rtol=0, atol=4 * np.finfo(np.float64).eps)
    # not the f32 values cast up: the f64 window carries more than f32's 24 bits ...
    f32 = getattr(nx.windows, name)(n, **kw)
    assert f32.dtype == np.float32
    if name not in ("bartlett", "triangular"):  # bartlett / triangular values are often exact in both types
        assert np.any(got != f32.astype(np.float64))
    # ... and agrees with them to f32 precision (kaiser's series amplifies the rounding a little)
    np.testing.assert_allclose(got, f32, rtol=0, atol=2e-6 if name == "kaiser" else 3e-7)


def test_rectangular_default_type_is_s64():
    w = nx.windows.rectangular(5)
    assert w.dtype == np.int64 and w.tolist() == [1] * 5
    assert nx.windows.rectangular(5, type="f64").dtype == np.float64


@pytest.mark.parametrize("kw", [dict(num_taps=51, cutoff=[0.3]), dict(num_taps=101, cutoff=[0.2, 0.5], pass_zero=False, window="hann"),
                                dict(num_taps=65, cutoff=[6000.0], sampling_rate=48000.0, window=("kaiser", 8.0)),
                                dict(num_taps=33, cutoff=[0.4], scale=False, window="blackman")])
def test_f64_firwin_is_computed_in_double(kw):
    got = nx.filters.firwin(type="f64", **kw)
    assert got.dtype == np.float64
    with o.float_type(o.F64):
        want = o.firwin(**kw)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-15)
    f32 = nx.filters.firwin(**kw)
    assert f32.dtype == np.float32 and np.any(got != f32.astype(np.float64))
    np.testing.assert_allclose(got, f32, rtol=0, atol=5e-7)
    # an independent check of the f64 values: scipy's firwin is the reference's own source of test vectors
    from scipy import signal

    win = kw.get("window", "hamming")
    sp = signal.firwin(kw["num_taps"], kw["cutoff"], window=win, pass_zero=kw.get("pass_zero", True), scale=kw.get("scale", True),
                       fs=kw.get("sampling_rate", 2.0))
    np.testing.assert_allclose(got, sp, rtol=0, atol=2e-6)  # the graph keeps f32 scalars (cutoffs, pi) and a truncated I0 series: ~1e-6, not 1e-16


@pytest.mark.parametrize("endpoint", [False, True])
@pytest.mark.parametrize("type_", ["f32", "f64"])
def test_fft_frequencies_forwards_endpoint_and_type(endpoint, type_):
    got = nx.fft_frequencies(1.6e4, 10, type=type_, endpoint=endpoint)
    with o.float_type(o.F64 if type_ == "f64" else o.F32):
        want = o.fft_frequencies(1.6e4, 10, endpoint=endpoint)
    assert got.dtype == (np.float64 if type_ == "f64" else np.float32)
    np.testing.assert_array_equal(got, want)
    # the doctest's last bin (lib/nx_signal.ex:147-151) without endpoint; with it the step is the f32 value of
    # 16000 / 9, whose f64 product with 9 is not rounded back to 16000
    assert abs(got[-1] - (1.6e4 if endpoint else 1.44e4)) <= (1e-3 if type_ == "f64" else 0)
    with pytest.raises(NotImplementedError):
        nx.fft_frequencies(1.6e4, 10, type="bf16")
