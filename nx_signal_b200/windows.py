"""NxSignal.Windows -- window generators (lib/nx_signal/windows.ex).

Same function heads and options as the reference (``is_periodic``, ``type``, ``beta``,
``eps``); values come from the C ABI's ``nxs_window_f32`` which reproduces
Nx.BinaryBackend's per-op f32 rounding bit for bit, or -- ``type="f64"`` -- from ``nxs_window_f64``,
the same graph evaluated in double."""
from __future__ import annotations

import numpy as np

from . import _lib


def _is_f64(type):
    return type in ("f64", np.float64) or (not isinstance(type, str) and np.dtype(type) == np.float64)


def _gen(kind, n, is_periodic=True, beta=12.0, eps=1.0e-7, type="f32"):
    if not isinstance(n, (int, np.integer)):
        raise _lib.NxSignalArgumentError(f"expected an integer window length, got: {n!r}")
    if _is_f64(type):  # the reference computes in the requested type (windows.ex:58,161,226,279,342)
        out = np.empty(int(n), dtype=np.float64)
        fn = _lib.lib().nxs_window_f64
    else:
        out = np.empty(int(n), dtype=np.float32)
        fn = _lib.lib().nxs_window_f32
    _lib.check(fn(_lib.WIN[kind], int(n), int(bool(is_periodic)), float(beta), float(eps), out.ctypes.data),
               what=f"Windows.{kind}")
    if type in ("f32", np.float32) or _is_f64(type):
        return out
    return out.astype(np.dtype(type) if not isinstance(type, str) else {"s64": np.int64, "s32": np.int32, "f16": np.float16}[type])


def rectangular(n, type="s64"):
    """windows.ex:33-36 (default type s64)."""
    return _gen("rectangular", n, type=type)


def bartlett(n, type="f32", name=None):
    """windows.ex:57-76."""
    return _gen("bartlett", n, type=type)


def triangular(n, type="f32", name=None):
    """windows.ex:98-127."""
    return _gen("triangular", n, type=type)


def blackman(n, is_periodic=True, type="f32", name=None):
    """windows.ex:160-199."""
    return _gen("blackman", n, is_periodic, type=type)


def hamming(n, is_periodic=True, type="f32", name=None):
    """windows.ex:225-252."""
    return _gen("hamming", n, is_periodic, type=type)


def hann(n, is_periodic=True, type="f32", name=None):
    """windows.ex:278-305."""
    return _gen("hann", n, is_periodic, type=type)


def kaiser(n, beta=12.0, eps=1.0e-7, is_periodic=True, type="f32", name=None):
    """windows.ex:341-386."""
    return _gen("kaiser", n, is_periodic, beta, eps, type=type)
