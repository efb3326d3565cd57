"""NxSignal.Filters.firwin (lib/nx_signal/filters.ex:147-279): FIR design by the window
method.  ``median`` / ``wiener`` are outside the accelerated path."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def firwin(num_taps, cutoff, window="hamming", pass_zero=True, scale=True, sampling_rate=2.0, type="f32"):
    """filters.ex:147-221.  ``window``: 'hamming' | 'hann' | 'blackman' | 'bartlett' |
    'rectangular' | ('kaiser', beta).  Raises the reference's ArgumentErrors."""
    if not isinstance(cutoff, (list, tuple)):
        raise _lib.NxSignalArgumentError(f"cutoff must be a list of frequencies, got: {cutoff!r}")
    nyq = sampling_rate / 2.0
    cl = sorted(c / nyq for c in cutoff)
    if cl[0] <= 0.0:
        raise _lib.NxSignalArgumentError(
            f"cutoff must be strictly between 0 and Nyquist (exclusive), got: {cl[0] * nyq}")
    if cl[-1] >= 1.0:
        raise _lib.NxSignalArgumentError(
            f"cutoff must be strictly between 0 and Nyquist (exclusive), got: {cl[-1] * nyq}")
    even = len(cl) % 2 == 0
    if ((pass_zero and even) or (not pass_zero and not even)) and num_taps % 2 == 0:
        raise _lib.NxSignalArgumentError(
            "a filter with non-zero gain at Nyquist (e.g. highpass) requires "
            f"an odd number of taps, got: {num_taps}")
    beta = 0.0
    if isinstance(window, tuple) and len(window) == 2 and window[0] == "kaiser":
        kind, beta = "kaiser", float(window[1])
    elif window in ("hamming", "hann", "blackman", "bartlett", "rectangular"):
        kind = window
    else:
        raise _lib.NxSignalArgumentError(
            f"unknown window {window!r}, supported: "
            ":hamming, :hann, :blackman, :bartlett, :rectangular, {:kaiser, beta}")
    cuts = (C.c_double * len(cutoff))(*[float(c) for c in cutoff])
    out = np.empty(int(num_taps), dtype=np.float32)
    _lib.check(
        _lib.lib().nxs_firwin_f32(int(num_taps), cuts, len(cutoff), _lib.WIN[kind], beta, int(bool(pass_zero)),
                                  int(bool(scale)), float(sampling_rate), out.ctypes.data),
        what="Filters.firwin")
    return out if type in ("f32", np.float32) else out.astype(np.float64)
