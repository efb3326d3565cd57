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
