"""GPU tests of the framing / overlap-add entry points against the reference's doctests."""
import numpy as np
import pytest

import nx_signal_b200 as nx
from oracle import nxsignal_oracle as o

pytestmark = pytest.mark.gpu


def test_as_windowed_doctests():  # lib/nx_signal.ex:182-246
    x = np.array([0, 1, 2, 3, 4, 10, 11, 12], dtype=np.int32)
    np.testing.assert_array_equal(nx.as_windowed(x, window_length=4), o.as_windowed(x, 4))
    np.testing.assert_array_equal(nx.as_windowed(x, window_length=3), o.as_windowed(x, 3))
    y = np.array([0, 1, 2, 3, 4, 10, 11], dtype=np.int32)
    got = nx.as_windowed(y, window_length=2, stride=2, padding=[(0, 3)])
    np.testing.assert_array_equal(got, np.array([[0, 1], [2, 3], [4, 10], [11, 0], [0, 0]], dtype=np.int32))
    assert got.dtype == np.int32
    np.testing.assert_array_equal(nx.as_windowed(np.arange(7, dtype=np.int32), window_length=6, padding="reflect", stride=1),
                                  o.as_windowed(np.arange(7, dtype=np.int32), 6, 1, "reflect"))
    np.testing.assert_array_equal(nx.as_windowed(np.arange(10, dtype=np.int32), window_length=6, padding="reflect", stride=2),
                                  o.as_windowed(np.arange(10, dtype=np.int32), 6, 2, "reflect"))


@pytest.mark.parametrize("dtype", [np.float32, np.int32, np.complex64, np.int64, np.float64])
@pytest.mark.parametrize("padding", ["valid", "same", "reflect", [(3, 11)]])
def test_as_windowed_types_and_padding(dtype, padding):
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((3, 2, 777)) * 100).astype(dtype)
    got = nx.as_windowed(x, window_length=64, stride=17, padding=padding)
    want = o.as_windowed(x, 64, 17, padding)
    assert got.dtype == x.dtype
    np.testing.assert_array_equal(got, want)


def test_overlap_and_add_doctests():  # lib/nx_signal.ex:656-681
    x = np.arange(12, dtype=np.int32).reshape(3, 4)
    np.testing.assert_array_equal(nx.overlap_and_add(x, overlap_length=0), np.arange(12, dtype=np.int32))
    np.testing.assert_array_equal(nx.overlap_and_add(x, overlap_length=3), np.array([0, 5, 15, 18, 17, 11], dtype=np.int32))
    t = np.array([[[[0, 1, 2, 3], [4, 5, 6, 7]]], [[[10, 11, 12, 13], [14, 15, 16, 17]]]], dtype=np.int32)
    np.testing.assert_array_equal(nx.overlap_and_add(t, overlap_length=3),
                                  np.array([[[0, 5, 7, 9, 7]], [[10, 25, 27, 29, 17]]], dtype=np.int32))


@pytest.mark.parametrize("cplx", [False, True])
def test_overlap_and_add_random(cplx):
    rng = np.random.default_rng(1)
    t = rng.standard_normal((2, 3, 50, 128)).astype(np.float32)
    if cplx:
        t = (t + 1j * rng.standard_normal(t.shape)).astype(np.complex64)
    for ov in (0, 1, 64, 96, 127):
        got = nx.overlap_and_add(t, overlap_length=ov)
        want = o.overlap_and_add(t, ov)
        assert got.dtype == t.dtype and got.shape == want.shape
        assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()


def test_device_tensors():
    import torch

    x = torch.arange(100, dtype=torch.float32, device="cuda").reshape(2, 50)
    fr = nx.as_windowed(x, window_length=8, stride=3)
    assert fr.is_cuda
    np.testing.assert_array_equal(fr.cpu().numpy(), o.as_windowed(x.cpu().numpy(), 8, 3))
    y = nx.overlap_and_add(fr, overlap_length=5)
    np.testing.assert_allclose(y.cpu().numpy(), o.overlap_and_add(fr.cpu().numpy(), 5), rtol=1e-6)


@pytest.mark.parametrize("padding", ["valid", "reflect", [(8, 12)], [(3, 5)], "same"])
@pytest.mark.parametrize("N,stride,L", [(64, 16, 1000), (64, 16, 1003), (8, 4, 64), (1024, 256, 9000), (64, 12, 1000)])
def test_as_windowed_vector_path_and_its_fallbacks(padding, N, stride, L):
    """f32 with N, stride, lo and the row length multiples of 4 takes the 128-bit kernel (padding groups
    handled per sample); anything else the scalar kernel -- same result either way."""
    rng = np.random.default_rng(N + stride + L)
    x = rng.standard_normal((3, L)).astype(np.float32)
    got = nx.as_windowed(x, window_length=N, stride=stride, padding=padding)
    np.testing.assert_array_equal(got, o.as_windowed(x, N, stride, padding))
