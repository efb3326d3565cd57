"""The oracle's median / wiener / argrel* restatements reproduce every vector the reference's
own tests and doctests hold for them (tests/postops_vectors.py cites the lines)."""
import numpy as np
import pytest

from oracle import nxsignal_oracle as o
from tests import postops_vectors as V


@pytest.mark.parametrize("case", ["MEDIAN_1D", "MEDIAN_2D", "MEDIAN_3D_K331", "MEDIAN_3D_K333"])
def test_median_vectors(case):
    t, ks, want = getattr(V, case)
    got = o.median(t, ks)
    assert got.dtype == np.float32
    np.testing.assert_array_equal(got, want)


def test_median_rank_error():  # filters_test.exs:99-117
    with pytest.raises(ValueError, match="kernel shape must be of the same rank as the tensor"):
        o.median(np.arange(10), (5, 5))
    with pytest.raises(ValueError, match="kernel shape must be of the same rank as the tensor"):
        o.median(np.arange(25).reshape(5, 5), (5, 5, 5))


def test_median_even_window_averages_the_middle_pair():
    np.testing.assert_array_equal(o.median(np.array([1.0, 4.0, 2.0, 8.0], dtype=np.float32), (2,)),
                                  np.array([2.5, 3.0, 5.0, 5.0], dtype=np.float32))


def test_wiener_vectors_bit_exact():
    np.testing.assert_array_equal(o.wiener(V.WIENER_IM, (3, 3)), V.WIENER_EST_F64)
    np.testing.assert_array_equal(o.wiener(V.WIENER_IM, 3), V.WIENER_EST_F64)
    np.testing.assert_array_equal(o.wiener(V.WIENER_IM.astype(np.float32), (3, 3)), V.WIENER_EST_F32)
    np.testing.assert_array_equal(o.wiener(V.WIENER_IM, (3, 3), noise=10), V.WIENER_N10_F64)
    np.testing.assert_array_equal(o.wiener(V.WIENER_IM.astype(np.float32), (3, 3), noise=10), V.WIENER_N10_F32)
    np.testing.assert_array_equal(o.wiener(V.WIENER_IM, (3, 3), noise=0), V.WIENER_IM)
    t, ks, nz, want = V.WIENER_DOC
    np.testing.assert_array_equal(o.wiener(t, ks, noise=nz), want)


@pytest.mark.parametrize("i", range(len(V.PEAKS)))
def test_argrel_vectors(i):
    cmp, x, kw, rows, count = V.PEAKS[i]
    idx, valid = o.argrelextrema(x, cmp, **kw)
    assert valid == count
    assert idx.dtype == np.int32 and idx.shape == (x.size, x.ndim)
    np.testing.assert_array_equal(idx[:count], np.asarray(rows, dtype=np.int32))
    assert (idx[count:] == -1).all()


def test_argrelextrema_non_strict_comparator():  # the doctest's custom comparator, restricted to >=
    x = np.array([0, 1, 1, 0, 2, 2, 2, 0], dtype=np.int32)
    idx, valid = o.argrelextrema(x, "greater_equal")
    np.testing.assert_array_equal(idx[:valid, 0], [1, 2, 4, 5, 6])  # edges compare with themselves and their one neighbour


# ---- independent cross-checks against scipy (the reference's docs name scipy as the model) --------
@pytest.mark.parametrize("shape,ks,noise", [((200,), 5, None), ((40, 50), (3, 5), None), ((40, 50), 3, 0.2),
                                            ((6, 20, 30), (3, 3, 3), None), ((31, 17), (4, 2), None)])
def test_wiener_matches_scipy(shape, ks, noise):
    from scipy.signal import wiener as sp_wiener

    rng = np.random.default_rng(sum(shape))
    t = rng.standard_normal(shape) + 1.0
    want = sp_wiener(t, ks, noise)
    got = o.wiener(t, ks if isinstance(ks, tuple) else int(ks), noise)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("shape,axis,order", [((300,), 0, 1), ((300,), 0, 5), ((20, 31), 0, 2), ((20, 31), 1, 1),
                                              ((4, 9, 11), 1, 2)])
@pytest.mark.parametrize("cmp", ["less", "greater", "less_equal", "greater_equal"])
def test_argrelextrema_matches_scipy(shape, axis, order, cmp):
    from scipy.signal import argrelextrema as sp_arg

    rng = np.random.default_rng(sum(shape) + order)
    x = rng.integers(-4, 5, size=shape)
    fn = {"less": np.less, "greater": np.greater, "less_equal": np.less_equal, "greater_equal": np.greater_equal}[cmp]
    want = np.stack(sp_arg(x, fn, axis=axis, order=order, mode="clip"), axis=-1)
    idx, valid = o.argrelextrema(x, cmp, axis=axis, order=order)
    assert valid == want.shape[0]
    np.testing.assert_array_equal(idx[:valid], want.astype(np.int32))


def test_median_equals_a_sliding_numpy_median_inside_the_tensor():
    """Away from the clamped tail the reference's window [i, i + k) is numpy's sliding window."""
    rng = np.random.default_rng(2)
    t = rng.standard_normal((5, 64)).astype(np.float32)
    k = 7
    win = np.lib.stride_tricks.sliding_window_view(t, k, axis=1)
    want = np.median(win.astype(np.float64), axis=-1).astype(np.float32)
    got = o.median(t, (1, k))
    np.testing.assert_array_equal(got[:, : 64 - k + 1], want)
    assert (got[:, 64 - k + 1:] == want[:, -1:]).all()  # clamped starts repeat the last full window
