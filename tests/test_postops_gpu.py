"""GPU parity of the spectrogram-adjacent operators (C ABI nxs_median_f32_*,
nxs_wiener_*, nxs_argrelextrema_f32_*) against the oracle and the reference's own vectors.
median and argrel* are selection / index work: bit-exact.  wiener is f64 arithmetic with a
global mean: compared at rtol 1e-12 (f64) / exactly after the f32 rounding up to 1 ulp."""
import numpy as np
import pytest

import nx_signal_b200 as nx
from oracle import nxsignal_oracle as o
from tests import postops_vectors as V

pytestmark = pytest.mark.gpu


# ---- median -------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["MEDIAN_1D", "MEDIAN_2D", "MEDIAN_3D_K331", "MEDIAN_3D_K333"])
def test_median_reference_vectors(case):
    t, ks, want = getattr(V, case)
    got = nx.Filters.median(t, ks)
    assert got.dtype == np.float32
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("shape,ks", [((1000,), (5,)), ((1000,), (8,)), ((64, 300), (1, 31)), ((64, 300), (17, 1)),
                                      ((64, 300), (3, 3)), ((64, 300), (4, 6)), ((5, 40, 50), (2, 3, 4)),
                                      ((7, 9), (7, 9)), ((3, 200), (1, 101)), ((12,), (1,))])
def test_median_vs_oracle_bit_exact(shape, ks):
    rng = np.random.default_rng(sum(shape) + sum(ks))
    t = rng.standard_normal(shape).astype(np.float32)
    t.reshape(-1)[::7] = 0.5  # ties
    np.testing.assert_array_equal(nx.Filters.median(t, ks), o.median(t, ks))


@pytest.mark.parametrize("shape,ks", [((7, 100), (1, 2)), ((7, 100), (1, 3)), ((7, 101), (1, 4)), ((7, 101), (1, 5)),
                                      ((7, 64), (1, 9)), ((7, 65), (1, 10)), ((7, 100), (1, 17)), ((7, 100), (1, 18)),
                                      ((7, 100), (1, 33)), ((7, 100), (1, 34)), ((7, 100), (1, 65)), ((7, 100), (1, 100)),
                                      ((100, 7), (17, 1)), ((101, 7), (18, 1)), ((5, 31, 6), (1, 31, 1)),
                                      ((5, 30, 6), (1, 7, 1)), ((40, 3, 6), (9, 1, 1)), ((64,), (64,)), ((3,), (2,))])
def test_median_one_axis_windows_bit_exact(shape, ks):
    """The shared-core kernel: odd / even windows, odd / even axis lengths, every axis, clamped tails."""
    rng = np.random.default_rng(sum(shape) * 31 + sum(ks))
    t = rng.integers(-4, 5, size=shape).astype(np.float32)  # heavy ties
    t += (rng.standard_normal(shape) * (rng.random(shape) < 0.5)).astype(np.float32)
    np.testing.assert_array_equal(nx.Filters.median(t, ks), o.median(t, ks))


def test_median_large_window_and_rank_counting_kernel(monkeypatch):
    """Windows above 64 elements use the rank-counting kernel; NXS_MEDIAN_NO_NET forces it for small ones."""
    rng = np.random.default_rng(8)
    t = rng.standard_normal((9, 400)).astype(np.float32)
    np.testing.assert_array_equal(nx.Filters.median(t, (1, 129)), o.median(t, (1, 129)))
    np.testing.assert_array_equal(nx.Filters.median(t, (9, 10)), o.median(t, (9, 10)))
    monkeypatch.setenv("NXS_MEDIAN_NO_AXIS", "1")  # one-axis windows through the general network kernel
    np.testing.assert_array_equal(nx.Filters.median(t, (1, 17)), o.median(t, (1, 17)))
    np.testing.assert_array_equal(nx.Filters.median(t, (4, 1)), o.median(t, (4, 1)))
    monkeypatch.setenv("NXS_MEDIAN_NO_NET", "1")
    np.testing.assert_array_equal(nx.Filters.median(t, (3, 5)), o.median(t, (3, 5)))
    np.testing.assert_array_equal(nx.Filters.median(t, (1, 4)), o.median(t, (1, 4)))


def test_median_on_a_spectrogram_stays_on_the_device():
    """HPSS-style use: median of |z| along time and along frequency, on CUDA tensors."""
    import torch

    rng = np.random.default_rng(3)
    x = torch.from_numpy(rng.standard_normal((1, 30_000)).astype(np.float32)).cuda()
    w = torch.from_numpy(o.hann(512)).cuda()
    z, _, _ = nx.stft(x, w, overlap_length=384, fft_length=512, onesided=True)
    mag = z.abs()[0].contiguous()  # [frames][bins]
    h = nx.Filters.median(mag, (17, 1))
    p = nx.Filters.median(mag, (1, 17))
    assert h.is_cuda and p.is_cuda and h.shape == mag.shape
    m = mag.cpu().numpy()
    np.testing.assert_array_equal(h.cpu().numpy(), o.median(m, (17, 1)))
    np.testing.assert_array_equal(p.cpu().numpy(), o.median(m, (1, 17)))


def test_median_errors():  # filters_test.exs:99-117
    with pytest.raises(nx.NxSignalArgumentError, match="kernel shape must be of the same rank as the tensor"):
        nx.Filters.median(np.arange(10), (5, 5))
    with pytest.raises(nx.NxSignalArgumentError, match="kernel shape must be of the same rank as the tensor"):
        nx.Filters.median(np.arange(25).reshape(5, 5), (5, 5, 5))
    with pytest.raises(nx.NxSignalArgumentError):
        nx.Filters.median(np.arange(10), (11,))


# ---- wiener -------------------------------------------------------------------------------------
def test_wiener_reference_vectors():
    np.testing.assert_allclose(nx.Filters.wiener(V.WIENER_IM, kernel_size=(3, 3)), V.WIENER_EST_F64, rtol=1e-14)
    np.testing.assert_allclose(nx.Filters.wiener(V.WIENER_IM, kernel_size=3), V.WIENER_EST_F64, rtol=1e-14)
    got32 = nx.Filters.wiener(V.WIENER_IM.astype(np.float32), kernel_size=(3, 3))
    assert got32.dtype == np.float32
    np.testing.assert_array_equal(got32, V.WIENER_EST_F32)
    np.testing.assert_allclose(nx.Filters.wiener(V.WIENER_IM, kernel_size=(3, 3), noise=10), V.WIENER_N10_F64, rtol=1e-14)
    np.testing.assert_array_equal(nx.Filters.wiener(V.WIENER_IM.astype(np.float32), kernel_size=(3, 3), noise=10),
                                  V.WIENER_N10_F32)
    np.testing.assert_array_equal(nx.Filters.wiener(V.WIENER_IM, kernel_size=(3, 3), noise=0), V.WIENER_IM)
    t, ks, nz, want = V.WIENER_DOC
    np.testing.assert_array_equal(nx.Filters.wiener(t, kernel_size=ks, noise=nz), want)


@pytest.mark.parametrize("shape,ks,noise", [((500,), 5, None), ((500,), (4,), 0.3), ((40, 70), (3, 5), None),
                                            ((40, 70), 3, 0.05), ((6, 20, 30), (2, 3, 3), None), ((300, 513), (1, 9), None)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_wiener_vs_oracle(shape, ks, noise, dtype):
    rng = np.random.default_rng(sum(shape))
    t = (rng.standard_normal(shape) + 2.0).astype(dtype)
    got = nx.Filters.wiener(t, kernel_size=ks, noise=noise)
    want = o.wiener(t, ks, noise)
    assert got.dtype == dtype
    if dtype == np.float64:
        np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-13)
    else:  # the same f64 value rounded once; the global mean's summation order may move 1 ulp
        np.testing.assert_allclose(got, want, rtol=2.0 ** -22, atol=0)


def test_wiener_cuda_tensor():
    import torch

    rng = np.random.default_rng(9)
    t = rng.standard_normal((128, 257)).astype(np.float32) ** 2
    got = nx.Filters.wiener(torch.from_numpy(t).cuda(), kernel_size=(5, 5))
    assert got.is_cuda and got.dtype == torch.float32
    np.testing.assert_allclose(got.cpu().numpy(), o.wiener(t, (5, 5)), rtol=2.0 ** -22, atol=0)


# ---- argrelmin / argrelmax / argrelextrema ------------------------------------------------------
@pytest.mark.parametrize("i", range(len(V.PEAKS)))
def test_argrel_reference_vectors(i):
    cmp, x, kw, rows, count = V.PEAKS[i]
    r = nx.PeakFinding.argrelextrema(x, cmp, **kw)
    assert int(r["valid_indices"]) == count
    idx = r["indices"]
    assert idx.dtype == np.int32 and idx.shape == (x.size, x.ndim)
    np.testing.assert_array_equal(idx[:count], np.asarray(rows, dtype=np.int32))
    assert (idx[count:] == -1).all()
    if cmp == "less":
        r2 = nx.PeakFinding.argrelmin(x, **kw)
    else:
        r2 = nx.PeakFinding.argrelmax(x, **kw)
    np.testing.assert_array_equal(r2["indices"], idx)


@pytest.mark.parametrize("shape,axis,order", [((5000,), 0, 1), ((5000,), 0, 4), ((37, 129), 0, 2), ((37, 129), 1, 1),
                                              ((6, 50, 33), 1, 3), ((6, 50, 33), 2, 1), ((3, 4, 5, 6), 2, 1),
                                              ((1,), 0, 1), ((2, 1500), -1, 700)])
@pytest.mark.parametrize("cmp", ["less", "greater", "less_equal", "greater_equal"])
def test_argrel_vs_oracle_bit_exact(shape, axis, order, cmp):
    rng = np.random.default_rng(sum(shape) + order)
    x = rng.integers(-5, 6, size=shape).astype(np.float32)  # many ties: strict and non-strict differ
    r = nx.PeakFinding.argrelextrema(x, cmp, axis=axis, order=order)
    idx, valid = o.argrelextrema(x, cmp, axis=axis % x.ndim, order=order)
    assert int(r["valid_indices"]) == valid
    np.testing.assert_array_equal(r["indices"], idx)


def test_argrelmax_along_frequency_of_a_device_spectrogram():
    import torch

    rng = np.random.default_rng(4)
    t = np.arange(48_000) / 48000.0
    x = (np.sin(2 * np.pi * 1000 * t) + 0.5 * np.sin(2 * np.pi * 5000 * t) + 0.01 * rng.standard_normal(t.size)).astype(np.float32)
    w = torch.from_numpy(o.hann(1024)).cuda()
    z, _, _ = nx.stft(torch.from_numpy(x[None]).cuda(), w, overlap_length=768, fft_length=1024, onesided=True)
    mag = z.abs()[0].contiguous()
    r = nx.PeakFinding.argrelmax(mag, axis=1, order=8)
    assert r["indices"].is_cuda
    idx, valid = o.argrelmax(mag.cpu().numpy(), axis=1, order=8)
    assert int(r["valid_indices"].cpu()) == valid
    np.testing.assert_array_equal(r["indices"].cpu().numpy(), idx)
    # the two tones are peaks of every frame: bins round(1000 / 46.875) = 21 and round(5000 / 46.875) = 107
    rows = idx[:valid]
    assert {21, 107} <= set(rows[rows[:, 0] == 10][:, 1].tolist())


def test_argrel_errors():
    with pytest.raises(NotImplementedError):
        nx.PeakFinding.argrelextrema(np.arange(4), lambda a, b: a > b)
    with pytest.raises(nx.NxSignalArgumentError):
        nx.PeakFinding.argrelmin(np.arange(4), axis=3)
