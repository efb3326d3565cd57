"""NxSignal.Filters (lib/nx_signal/filters.ex): ``firwin`` (FIR design by the window method,
:147-279, host arithmetic) and the spectrogram-adjacent ``median`` (:17-56) / ``wiener``
(:80-110, :281-303) filters on the device (csrc/nxs_post.cu)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _arrays as A
from . import _lib


def _i64(vals):
    return (C.c_int64 * max(len(vals), 1))(*[int(v) for v in vals])


def median(t, kernel_shape):
    """filters.ex:17-56: out[i] = median of the ``kernel_shape`` window that starts at i (start
    clamped so the window stays inside the tensor), as f32.  Rank <= 3."""
    x = A.to_real_f32(t, "t")
    ks = tuple(kernel_shape) if isinstance(kernel_shape, (tuple, list)) else None
    if ks is None or len(ks) != x.ndim:
        raise _lib.NxSignalArgumentError("kernel shape must be of the same rank as the tensor")
    if x.ndim > 3:
        raise NotImplementedError("median: rank > 3 is not supported by this backend")
    shape = tuple(int(s) for s in x.shape)
    if any(k < 1 or k > d for k, d in zip(ks, shape)):
        raise _lib.NxSignalArgumentError(f"kernel shape {tuple(ks)} does not fit inside the tensor shape {shape}")
    out = A.empty_like_kind(x, shape, "f32")
    ctx = _lib.context(A.device_index(x))
    if A.is_cuda(x):
        rc = _lib.lib().nxs_median_f32_dev(ctx, A.ptr(x), x.ndim, _i64(shape), _i64(ks), A.ptr(out), A.stream_of(x))
    else:
        rc = _lib.lib().nxs_median_f32_host(ctx, A.ptr(x), x.ndim, _i64(shape), _i64(ks), A.ptr(out))
    _lib.check(rc, ctx, "Filters.median")
    return out


def wiener(t, kernel_size=3, noise=None):
    """filters.ex:80-110: Wiener filter with a ``kernel_size`` (int or tuple) local window, computed in
    f64 like the reference; the result has t's float type (f32 unless t is f64).  ``noise`` = None
    estimates the noise power as the mean local variance.  Rank <= 3."""
    is_t = A.is_torch(t)
    if is_t:
        import torch

        f64 = t.dtype == torch.float64
        x = t.contiguous() if f64 else A.to_real_f32(t, "t")
    else:
        a = np.asarray(t)
        f64 = a.dtype == np.float64
        x = np.ascontiguousarray(a) if f64 else A.to_real_f32(a, "t")
    if isinstance(kernel_size, (int, np.integer)):
        ks = (int(kernel_size),) * x.ndim
    elif isinstance(kernel_size, tuple):
        ks = kernel_size
    else:
        raise _lib.NxSignalArgumentError("kernel_size must be an integer or tuple")
    if len(ks) != x.ndim:
        raise _lib.NxSignalArgumentError("kernel_size must have the tensor's rank")
    if x.ndim > 3:
        raise NotImplementedError("wiener: rank > 3 is not supported by this backend")
    shape = tuple(int(s) for s in x.shape)
    out = A.empty_like_kind(x, shape, np.float64 if f64 else "f32")
    ctx = _lib.context(A.device_index(x))
    has_noise, nz = (0, 0.0) if noise is None else (1, float(noise))
    if A.is_cuda(x):
        rc = _lib.lib().nxs_wiener_dev(ctx, A.ptr(x), int(f64), x.ndim, _i64(shape), _i64(ks), has_noise, nz, A.ptr(out),
                                       A.stream_of(x))
    else:
        rc = _lib.lib().nxs_wiener_host(ctx, A.ptr(x), int(f64), x.ndim, _i64(shape), _i64(ks), has_noise, nz, A.ptr(out))
    _lib.check(rc, ctx, "Filters.wiener")
    # the reference casts the f64 result back to the INPUT's type (filters.ex:108-110): integers truncate
    if is_t:
        return out if t.dtype.is_floating_point and out.dtype == t.dtype or t.dtype == torch.float64 else out.to(t.dtype)
    src = np.asarray(t).dtype
    return out if src == out.dtype else out.astype(src)


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
    f64 = type in ("f64", np.float64)
    if not f64 and type not in ("f32", np.float32):
        raise NotImplementedError(f"firwin: type {type!r} is not supported by this backend (f32, f64)")
    out = np.empty(int(num_taps), dtype=np.float64 if f64 else np.float32)
    fn = _lib.lib().nxs_firwin_f64 if f64 else _lib.lib().nxs_firwin_f32  # computed in the requested type (filters.ex:153)
    _lib.check(fn(int(num_taps), cuts, len(cutoff), _lib.WIN[kind], beta, int(bool(pass_zero)), int(bool(scale)),
                  float(sampling_rate), out.ctypes.data), what="Filters.firwin")
    return out
