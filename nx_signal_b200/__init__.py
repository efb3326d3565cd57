"""nx_signal_b200 -- B200-native backend for NxSignal's STFT / ISTFT / windows / FIR path.

Host-side mirror of the reference's public function heads (elixir-nx/nx_signal v0.3.0,
``lib/nx_signal.ex``): same names, option names, defaults, return shapes and error
conditions, over the C ABI in ``include/nxsignal_b200.h``.  Batch dimensions of ``data``
stand in for Nx vectorised axes.  numpy arrays use the host entry points, CUDA torch
tensors the device entry points (asynchronous on torch's current stream).

There is no CPU fallback: without the built shared library or without a CUDA device
every compute call raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _arrays as A
from . import _lib
from . import convolution, filters, peak_finding, windows
from ._lib import NxSignalArgumentError

Windows = windows
Filters = filters
Convolution = convolution
PeakFinding = peak_finding

__all__ = [
    "stft", "istft", "as_windowed", "overlap_and_add", "fft_frequencies", "stft_times", "mel_filters", "stft_to_mel", "stft_mel",
    "Windows", "Filters", "Convolution", "PeakFinding", "windows", "filters", "convolution", "peak_finding",
    "NxSignalArgumentError",
]

_SCALING = {None: _lib.SCALE_NONE, "spectrum": _lib.SCALE_SPECTRUM, "psd": _lib.SCALE_PSD}


def _next_pow2(n):
    p = 1
    while p < n:
        p *= 2
    return p


def _scaling_code(scaling):
    if scaling not in _SCALING:
        raise NxSignalArgumentError(
            f"invalid :scaling, expected one of :spectrum, :psd or nil, got: {scaling!r}")
    return _SCALING[scaling]


def _padding_code(padding):
    """lib/nx_signal.ex:250-255, 303-331 -> (mode, lo, hi)."""
    if isinstance(padding, str):
        if padding == "valid":
            return _lib.PAD_VALID, 0, 0
        if padding == "same":
            return _lib.PAD_SAME, 0, 0
        if padding == "reflect":
            return _lib.PAD_REFLECT, 0, 0
        raise NxSignalArgumentError(
            "invalid padding mode specified, padding must be one of :valid, :same, "
            f"or a padding configuration, got: {padding!r}")
    if isinstance(padding, (list, tuple)) and len(padding) == 1:
        pair = padding[0]
        if (isinstance(pair, (list, tuple)) and len(pair) == 2
                and all(isinstance(v, (int, np.integer)) for v in pair)):
            return _lib.PAD_EXPLICIT, int(pair[0]), int(pair[1])
    raise NxSignalArgumentError(
        "padding must be a list of {high, low} tuples, where each element is an integer. "
        f"Got: {padding!r}")


def _num_frames(length, window_length, stride, mode, lo, hi):
    m = C.c_int64()
    _lib.check(_lib.lib().nxs_num_frames(length, window_length, stride, mode, lo, hi, C.byref(m)),
               what="as_windowed")
    return m.value


def fft_frequencies(sampling_rate, fft_length, type="f32", name="frequencies", endpoint=False):
    """NxSignal.fft_frequencies/2 (lib/nx_signal.ex:154-166): [fft_length] of `type` ("f32" | "f64");
    `endpoint` is forwarded to the reference's Nx.linspace (divide by n - 1 instead of n)."""
    f64 = type in ("f64", np.float64)
    if not f64 and type not in ("f32", np.float32):
        raise NotImplementedError(f"fft_frequencies: type {type!r} is not supported by this backend (f32, f64)")
    out = np.empty(int(fft_length), dtype=np.float64 if f64 else np.float32)
    _lib.check(_lib.lib().nxs_fft_frequencies_ex(float(sampling_rate), int(fft_length), int(bool(endpoint)), int(f64),
                                                 out.ctypes.data), what="fft_frequencies")
    return out


def stft_times(frame_length, sampling_rate, num_frames):
    """Frame times of stft/3 (lib/nx_signal.ex:108-111)."""
    out = np.empty(int(num_frames), dtype=np.float32)
    _lib.check(_lib.lib().nxs_stft_times_f32(int(frame_length), float(sampling_rate), int(num_frames),
                                             out.ctypes.data), what="stft")
    return out


def stft(data, window, overlap_length=None, fft_length="power_of_two", window_padding="valid",
         sampling_rate=100, scaling=None, onesided=False, **ignored):
    """NxSignal.stft/3 (lib/nx_signal.ex:68-130).

    data [..., L], window [N] -> (z c64 [..., M, fft_length], times f32 [M], frequencies f32
    [fft_length]).  Defaults as the reference: overlap_length = N // 2, fft_length =
    next power of two >= N, window_padding = 'valid', sampling_rate = 100 (sic, :77).
    The reference's unused ``:window`` option is accepted and ignored (:74).

    ``onesided=True`` is an opt-in extension (not in the reference): CUDA tensors
    only, z holds bins 0 .. fft_length // 2 ([..., M, fft_length // 2 + 1]), frequencies
    likewise; the dropped bins are conj(z[..., fft_length - k])."""
    for k in ignored:
        if k != "window":
            raise NxSignalArgumentError(f"unknown keys [{k!r}] in options")
    if sampling_rate is None:
        raise NxSignalArgumentError("missing sampling_rate option")
    scale = _scaling_code(scaling)
    # complex `data` is a complex transform in the reference's graph (Nx.multiply -> Nx.fft, :101-102)
    cplx = (data.is_complex() if A.is_torch(data) else np.iscomplexobj(np.asarray(data)))
    x = A.to_c64(data) if cplx else A.to_real_f32(data, "data")
    w = A.like_device(x, A.to_real_f32(window, "window"))
    if w.ndim != 1:
        raise NxSignalArgumentError("window must be a rank-1 tensor")
    N = int(w.shape[0])
    if overlap_length is None:
        overlap_length = N // 2
    hop = N - int(overlap_length)
    if hop < 1:
        raise NxSignalArgumentError(f"expected an integer >= 1 or a list of integers, got: {hop!r}")
    nfft = _next_pow2(N) if fft_length == "power_of_two" else int(fft_length)
    mode, lo, hi = _padding_code(window_padding)
    L = int(x.shape[-1])
    batch_shape = tuple(x.shape[:-1])
    Cn = int(np.prod(batch_shape, dtype=np.int64)) if batch_shape else 1
    M = _num_frames(L, N, hop, mode, lo, hi)
    if cplx:
        if onesided:
            raise NxSignalArgumentError("onesided=True needs real data (a complex signal's spectrum has no mirror symmetry)")
        z = A.empty_like_kind(x, batch_shape + (M, nfft), "c64")
        if M > 0 and Cn > 0:
            ctx = _lib.context(A.device_index(x))
            args = (Cn, L, L, A.ptr(w), N, hop, nfft, mode, lo, hi, scale, float(sampling_rate), A.ptr(z))
            if A.is_cuda(x):
                rc = _lib.lib().nxs_stft_c64_dev(ctx, A.ptr(x), *args, A.stream_of(x))
            else:
                rc = _lib.lib().nxs_stft_c64_host(ctx, A.ptr(x), *args)
            _lib.check(rc, ctx, "stft")
        return z, A.from_host(x, stft_times(N, sampling_rate, M)), A.from_host(x, fft_frequencies(sampling_rate, nfft))
    if onesided and not A.is_cuda(x):
        raise NotImplementedError("onesided=True takes CUDA tensors (the host entry already moves the one-sided "
                                  "form over PCIe and returns the reference's two-sided result)")
    nout = nfft // 2 + 1 if onesided else nfft
    z = A.empty_like_kind(x, batch_shape + (M, nout), "c64")
    if M > 0 and Cn > 0:
        dev = A.device_index(x)
        ctx = _lib.context(dev)
        if onesided:
            rc = _lib.lib().nxs_stft_onesided_f32_dev(ctx, A.ptr(x), Cn, L, L, A.ptr(w), N, hop, nfft, mode, lo,
                                                      hi, scale, float(sampling_rate), A.ptr(z), nout,
                                                      A.stream_of(x))
        elif A.is_cuda(x):
            rc = _lib.lib().nxs_stft_f32_dev(ctx, A.ptr(x), Cn, L, L, A.ptr(w), N, hop, nfft, mode, lo, hi,
                                             scale, float(sampling_rate), A.ptr(z), A.stream_of(x))
        else:
            rc = _lib.lib().nxs_stft_f32_host(ctx, A.ptr(x), Cn, L, L, A.ptr(w), N, hop, nfft, mode, lo, hi,
                                              scale, float(sampling_rate), A.ptr(z))
        _lib.check(rc, ctx, "stft")
    times = A.from_host(x, stft_times(N, sampling_rate, M))
    freqs = A.from_host(x, fft_frequencies(sampling_rate, nfft)[:nout])
    return z, times, freqs


def mel_filters(fft_length, mel_bins, sampling_rate, max_mel=3016, mel_frequency_spacing=200 / 3, type="f32"):
    """NxSignal.mel_filters/4 (lib/nx_signal.ex:397-445): f32 [mel_bins, fft_length]."""
    out = np.empty((int(mel_bins), int(fft_length)), dtype=np.float32)
    _lib.check(_lib.lib().nxs_mel_filters_f32(int(fft_length), int(mel_bins), float(sampling_rate), float(max_mel),
                                              float(mel_frequency_spacing), out.ctypes.data), what="mel_filters")
    return out


def stft_to_mel(z, sampling_rate, fft_length=None, mel_bins=128, max_mel=3016, mel_frequency_spacing=200 / 3,
                type="f32"):
    """NxSignal.stft_to_mel/3 (lib/nx_signal.ex:486-513): z c64 [..., M, K] -> f32 [..., M, mel_bins].

    Only bins 0 .. fft_length // 2 - 1 of z are used (:496), so K may be fft_length (the
    reference's two-sided spectrum) or fft_length // 2 + 1 (``stft(..., onesided=True)``).  Leading
    axes stand in for Nx vectorised axes: the dynamic-range maximum (:512) is taken per entry."""
    if fft_length is None:
        raise NxSignalArgumentError("missing :fft_length option")
    zz = A.to_c64(z)
    if zz.ndim < 2:
        raise NxSignalArgumentError("z must have at least the [frames, frequencies] axes")
    M, K = int(zz.shape[-2]), int(zz.shape[-1])
    nfft = int(fft_length)
    if K < nfft // 2:
        raise NxSignalArgumentError(f"z has {K} frequency bins, fewer than fft_length / 2 = {nfft // 2}")
    batch_shape = tuple(zz.shape[:-2])
    Cn = int(np.prod(batch_shape, dtype=np.int64)) if batch_shape else 1
    out = A.empty_like_kind(zz, batch_shape + (M, int(mel_bins)), "f32")
    if Cn > 0 and M > 0:
        ctx = _lib.context(A.device_index(zz))
        args = (Cn, M, K, nfft, int(mel_bins), float(sampling_rate), float(max_mel), float(mel_frequency_spacing))
        if A.is_cuda(zz):
            rc = _lib.lib().nxs_stft_to_mel_f32_dev(ctx, A.ptr(zz), *args, A.ptr(out), A.stream_of(zz))
        else:
            rc = _lib.lib().nxs_stft_to_mel_f32_host(ctx, A.ptr(zz), *args, A.ptr(out))
        _lib.check(rc, ctx, "stft_to_mel")
    return out


def stft_mel(data, window, overlap_length=None, fft_length="power_of_two", window_padding="valid",
             sampling_rate=100, scaling=None, mel_bins=128, max_mel=3016, mel_frequency_spacing=200 / 3):
    """``stft_to_mel(stft(data, window, ...)[0], sampling_rate, ...)`` in one pass (extension,
    SURVEY 8f rank 1): f32 [..., M, mel_bins].  The fused kernel keeps each frame's spectrum on chip;
    configurations it does not serve chain the two device entries instead.  numpy arrays go through
    the host entry (H2D of the signal, D2H of the mel tensor only)."""
    scale = _scaling_code(scaling)
    x = A.to_real_f32(data, "data")
    w = A.like_device(x, A.to_real_f32(window, "window"))
    N = int(w.shape[0])
    if overlap_length is None:
        overlap_length = N // 2
    hop = N - int(overlap_length)
    if hop < 1:
        raise NxSignalArgumentError(f"expected an integer >= 1 or a list of integers, got: {hop!r}")
    nfft = _next_pow2(N) if fft_length == "power_of_two" else int(fft_length)
    mode, lo, hi = _padding_code(window_padding)
    L = int(x.shape[-1])
    batch_shape = tuple(x.shape[:-1])
    Cn = int(np.prod(batch_shape, dtype=np.int64)) if batch_shape else 1
    M = _num_frames(L, N, hop, mode, lo, hi)
    out = A.empty_like_kind(x, batch_shape + (M, int(mel_bins)), "f32")
    if M > 0 and Cn > 0:
        ctx = _lib.context(A.device_index(x))
        if not A.is_cuda(x):
            rc = _lib.lib().nxs_stft_mel_f32_host(ctx, A.ptr(x), Cn, L, L, A.ptr(w), N, hop, nfft, mode, lo, hi, scale,
                                                  float(sampling_rate), int(mel_bins), float(max_mel),
                                                  float(mel_frequency_spacing), A.ptr(out))
            _lib.check(rc, ctx, "stft_mel")
            return out
        rc = _lib.lib().nxs_stft_mel_f32_dev(ctx, A.ptr(x), Cn, L, L, A.ptr(w), N, hop, nfft, mode, lo, hi, scale,
                                             float(sampling_rate), int(mel_bins), float(max_mel),
                                             float(mel_frequency_spacing), A.ptr(out), A.stream_of(x))
        if rc == _lib.NXS_EUNSUPPORTED:
            z, _, _ = stft(x, w, overlap_length=overlap_length, fft_length=nfft, window_padding=window_padding,
                           sampling_rate=sampling_rate, scaling=scaling, onesided=nfft % 2 == 0)
            return stft_to_mel(z, sampling_rate, fft_length=nfft, mel_bins=mel_bins, max_mel=max_mel,
                               mel_frequency_spacing=mel_frequency_spacing)
        _lib.check(rc, ctx, "stft_mel")
    return out


def istft(data, window, fft_length=None, overlap_length=None, scaling=None, sampling_rate=1000, onesided=False):
    """NxSignal.istft/3 (lib/nx_signal.ex:582-638).

    data c64 [..., M, K], window [N] -> c64 [..., M*hop + N - hop].  fft_length defaults to
    the next power of two >= K; it must equal N (the reference's `frames * window`).

    ``onesided=True`` is an opt-in extension (not in the reference), the counterpart
    of ``stft(..., onesided=True)``: data holds bins 0 .. fft_length // 2 (K = fft_length // 2 + 1,
    fft_length defaults to len(window)) and the result is the REAL f32 signal
    Re(istft(ext(data))), ext = conjugate-mirror extension."""
    scale = _scaling_code(scaling)
    if scaling == "psd" and sampling_rate is None:
        raise NxSignalArgumentError(":sampling_rate is mandatory if scaling is :psd")
    if sampling_rate is None:
        sampling_rate = 1000
    z = A.to_c64(data)
    w = A.like_device(z, A.to_real_f32(window, "window"))
    N = int(w.shape[0])
    K = int(z.shape[-1])
    M = int(z.shape[-2])
    if onesided and fft_length in (None, "power_of_two"):
        nfft = N
    else:
        nfft = _next_pow2(K) if fft_length in (None, "power_of_two") else int(fft_length)
    if overlap_length is None:
        overlap_length = N // 2
    hop = N - int(overlap_length)
    if onesided and (nfft % 2 or K < nfft // 2 + 1):
        raise NxSignalArgumentError(
            f"onesided istft needs an even fft_length and fft_length // 2 + 1 = {nfft // 2 + 1} bins, got {K}")
    if nfft != N:
        raise NxSignalArgumentError(
            f"cannot broadcast frames of length {nfft} with window of length {N} "
            "(istft requires fft_length == length(window))")
    if int(overlap_length) >= N:
        raise NxSignalArgumentError(
            f"overlap_length must be a number less than the window size {N}, got: {N}")
    batch_shape = tuple(z.shape[:-2])
    Cn = int(np.prod(batch_shape, dtype=np.int64)) if batch_shape else 1
    out_len = M * hop + (N - hop)
    y = A.empty_like_kind(z, batch_shape + (out_len,), "f32" if onesided else "c64")
    if Cn > 0:
        ctx = _lib.context(A.device_index(z))
        if onesided:
            if A.is_cuda(z):
                rc = _lib.lib().nxs_istft_c2r_f32_dev(ctx, A.ptr(z), Cn, M, K, A.ptr(w), N, hop, nfft, scale,
                                                      float(sampling_rate), A.ptr(y), A.stream_of(z))
            else:
                rc = _lib.lib().nxs_istft_c2r_f32_host(ctx, A.ptr(z), Cn, M, K, A.ptr(w), N, hop, nfft, scale,
                                                       float(sampling_rate), A.ptr(y))
        elif A.is_cuda(z):
            rc = _lib.lib().nxs_istft_c64_dev(ctx, A.ptr(z), Cn, M, K, A.ptr(w), N, hop, nfft, scale,
                                              float(sampling_rate), A.ptr(y), A.stream_of(z))
        else:
            rc = _lib.lib().nxs_istft_c64_host(ctx, A.ptr(z), Cn, M, K, A.ptr(w), N, hop, nfft, scale,
                                               float(sampling_rate), A.ptr(y))
        _lib.check(rc, ctx, "istft")
    return y


def _elem8(x):
    if A.is_torch(x):
        return x.element_size() == 8
    return x.dtype.itemsize == 8


def as_windowed(tensor, window_length=None, stride=1, padding="valid"):
    """NxSignal.as_windowed/2 (lib/nx_signal.ex:249-364): [..., L] -> [..., M, window_length],
    element type preserved (4- or 8-byte elements)."""
    if window_length is None:
        raise NxSignalArgumentError("missing :window_length option")
    if not (isinstance(stride, (int, np.integer)) and stride >= 1):
        raise NxSignalArgumentError(f"expected an integer >= 1 or a list of integers, got: {stride!r}")
    mode, lo, hi = _padding_code(padding)
    x = tensor if A.is_torch(tensor) else np.ascontiguousarray(np.asarray(tensor))
    if A.is_torch(x):
        x = x.contiguous()
    es = x.element_size() if A.is_torch(x) else x.dtype.itemsize
    if es not in (4, 8):
        raise NotImplementedError("as_windowed supports 4- and 8-byte element types")
    L = int(x.shape[-1])
    batch_shape = tuple(x.shape[:-1])
    Cn = int(np.prod(batch_shape, dtype=np.int64)) if batch_shape else 1
    M = _num_frames(L, int(window_length), int(stride), mode, lo, hi)
    if A.is_cuda(x):
        import torch

        out = torch.empty(batch_shape + (M, int(window_length)), dtype=x.dtype, device=x.device)
    elif A.is_torch(x):
        raise NotImplementedError("CPU torch tensors: pass numpy arrays or CUDA tensors")
    else:
        out = np.empty(batch_shape + (M, int(window_length)), dtype=x.dtype)
    if M > 0 and Cn > 0:
        ctx = _lib.context(A.device_index(x))
        if A.is_cuda(x):
            rc = _lib.lib().nxs_as_windowed_dev(ctx, A.ptr(x), es, Cn, L, L, int(window_length), int(stride),
                                                mode, lo, hi, A.ptr(out), A.stream_of(x))
        else:
            rc = _lib.lib().nxs_as_windowed_host(ctx, A.ptr(x), es, Cn, L, L, int(window_length), int(stride),
                                                 mode, lo, hi, A.ptr(out))
        _lib.check(rc, ctx, "as_windowed")
    return out


def overlap_and_add(tensor, overlap_length=None, type=None):
    """NxSignal.overlap_and_add/2 (lib/nx_signal.ex:684-735): [..., M, N] -> [..., M*hop + overlap].
    The element type is preserved (integers are summed in f32, exact below 2**24)."""
    if overlap_length is None:
        raise NxSignalArgumentError("missing :overlap_length option")
    t = tensor if A.is_torch(tensor) else np.asarray(tensor)
    M, N = int(t.shape[-2]), int(t.shape[-1])
    if int(overlap_length) >= N:
        raise NxSignalArgumentError(
            f"overlap_length must be a number less than the window size {N}, got: {N}")
    if A.is_torch(t):
        cplx = t.is_complex()
        in_dtype = t.dtype
        work = A.to_c64(t) if cplx else A.to_real_f32(t)
    else:
        cplx = np.iscomplexobj(t)
        in_dtype = t.dtype
        work = A.to_c64(t) if cplx else A.to_real_f32(t)
    batch_shape = tuple(t.shape[:-2])
    B = int(np.prod(batch_shape, dtype=np.int64)) if batch_shape else 1
    hop = N - int(overlap_length)
    out = A.empty_like_kind(work, batch_shape + (M * hop + int(overlap_length),), "c64" if cplx else "f32")
    if B > 0:
        ctx = _lib.context(A.device_index(work))
        l = _lib.lib()
        if A.is_cuda(work):
            fn = l.nxs_overlap_and_add_c64_dev if cplx else l.nxs_overlap_and_add_f32_dev
            rc = fn(ctx, A.ptr(work), B, M, N, int(overlap_length), A.ptr(out), A.stream_of(work))
        else:
            fn = l.nxs_overlap_and_add_c64_host if cplx else l.nxs_overlap_and_add_f32_host
            rc = fn(ctx, A.ptr(work), B, M, N, int(overlap_length), A.ptr(out))
        _lib.check(rc, ctx, "overlap_and_add")
    target = type if type is not None else in_dtype
    if A.is_torch(out):
        return out if out.dtype == target else (out.round().to(target) if not target.is_floating_point and not target.is_complex else out.to(target))
    target = np.dtype(target)
    if target == out.dtype:
        return out
    if np.issubdtype(target, np.integer):
        return np.rint(out).astype(target)
    return out.astype(target)
