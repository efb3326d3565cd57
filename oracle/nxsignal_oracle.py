"""CPU oracle for the NxSignal STFT / ISTFT / windows / FIR hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``nx_signal_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker
or the timed CPU baseline, never as the product path.

What it is
----------
A literal numpy restatement of the reference's Elixir ``defn`` graphs
(elixir-nx/nx_signal v0.3.0 @ dcf5b81) *together with* the semantics of the
backend that executes them by default, ``Nx.BinaryBackend`` from the
un-vendored hex dependency ``nx 0.11.0`` (``/root/reference/mix.lock:10``;
complex arithmetic from ``complex 0.6.0``, ``mix.lock:2``).  Nx's source is not
in ``/root/reference``, so its published algorithm is restated here:

  * every Nx op computes element-wise in IEEE double (Elixir floats /
    ``Complex`` structs of doubles) and rounds ONCE to the op's output type
    (f32 / c64); float literals are first rounded to f32 tensors;
  * ``Nx.fft`` / ``Nx.ifft``: zero-pad / truncate to ``length``; recursive
    radix-2 decimation in time while n is even (base cases n = 1, 2), twiddle
    ``exp(-+ i * (2*pi/n) * k)`` as (cos, sin) in double; naive O(n^2) DFT for
    odd n; ``ifft`` divides by n; ``|re|, |im| <= eps`` (1e-10) snapped to 0;
    one rounding to c64;
  * ``Nx.indexed_add`` / ``Nx.sum`` / ``Nx.dot`` / ``Nx.conv`` accumulate in
    double and round once.

Pinning
-------
``tests/test_oracle_golden.py`` checks this file against every doctest /
known-answer vector the reference holds for the path:
``lib/nx_signal.ex:46-65,147-151,182-246,465-483,545-579,656-681``,
``lib/nx_signal/windows.ex:20-338``, ``lib/nx_signal/convolution.ex:32-36,
81-85,246-250``, ``test/nx_signal/filters_test.exs:248-416``,
``test/nx_signal/convolutions_test.exs``.  STFT/ISTFT at nfft >= 32 is NOT
pinned by any reference vector (the reference has none); there the oracle
itself is the anchor.  ``median`` / ``wiener`` / ``argrel*`` (end of file) are
pinned by ``tests/test_postops_oracle.py`` to
``test/nx_signal/filters_test.exs:5-245`` and the doctests at
``lib/nx_signal/filters.ex:68-78``, ``lib/nx_signal/peak_finding.ex:36-128,
158-250`` (bit-exact, including the f64 wiener digits).

Each function cites the reference file:line it follows.
"""
from __future__ import annotations

import contextlib
import math
from typing import Optional, Sequence, Tuple, Union

import numpy as np

F32 = np.float32
F64 = np.float64
C64 = np.complex64
C128 = np.complex128

PI = math.pi


# ---------------------------------------------------------------------------
# Nx.BinaryBackend element-wise semantics: compute in double, round once.
# ---------------------------------------------------------------------------
# The float type the windows / firwin / fft_frequencies graphs run in: f32 by default; `float_type(F64)`
# restates the same graphs for `type: :f64` (windows.ex:58,161,226,279,342; filters.ex:153), where every op is
# plain double arithmetic.  UNPINNED: the reference holds no f64 vector for these functions.
_FT = F32


@contextlib.contextmanager
def float_type(t):
    global _FT
    old, _FT = _FT, t
    try:
        yield
    finally:
        _FT = old


def _f(x):
    """Round to the graph's float type (the single rounding every Nx op performs; f32 unless float_type(F64))."""
    return np.asarray(x, dtype=F64).astype(_FT)


def _f32(x):
    """Round to f32 whatever the graph's type (numbers entering a defn are f32 scalars)."""
    return np.asarray(x, dtype=F64).astype(F32)


def _d(x):
    return np.asarray(x, dtype=F64)


def _c(x):
    """Round a complex128 array to c64 component-wise."""
    return np.asarray(x, dtype=C128).astype(C64)


def _mul(a, b):
    return _f(_d(a) * _d(b))


def _div(a, b):
    return _f(_d(a) / _d(b))


def _add(a, b):
    return _f(_d(a) + _d(b))


def _sub(a, b):
    return _f(_d(a) - _d(b))


def _cos(a):
    return _f(np.cos(_d(a)))


def _sin(a):
    return _f(np.sin(_d(a)))


def _lit(x):
    """A float literal in a defn becomes a scalar of the tensor it meets (f32 by default)."""
    return _FT(x)


def _iota(n):
    return np.arange(n, dtype=_FT)


def nx_linspace(start, stop, n, endpoint=True):
    """Nx.linspace for f32 as tensor ops: start/stop become f32 tensors,
    step = f32(f32(stop - start) / divisor), out = f32(f32(iota * step) + start).

    Pinned by the stft doctest times (lib/nx_signal.ex:56-60), the
    fft_frequencies doctest (:147-151) and the mel doctests (:384-394, :465-483).
    """
    start = _tensor_scalar(start)
    stop = _tensor_scalar(stop)
    div = (n - 1) if endpoint else n
    diff = _f32(_d(stop) - _d(start))  # start / stop are f32 scalars whatever `type:` is, so the step is an f32 value
    with np.errstate(divide="ignore", invalid="ignore"):
        step = _f32(_d(diff) / div)
        return _add(_mul(_iota(n), step), start)


def _tensor_scalar(x):
    """A number crossing into a defn as an argument: ints stay exact (s32),
    floats become f32."""
    if isinstance(x, (int, np.integer)):
        return F64(x)
    return F64(F32(x))


# ---------------------------------------------------------------------------
# Windows (lib/nx_signal/windows.ex)
# ---------------------------------------------------------------------------
def rectangular(n: int, dtype=np.int64):
    """windows.ex:33-36 (default type s64)."""
    return np.ones(n, dtype=dtype)


def bartlett(n: int):
    """windows.ex:57-76: concat(left*2/n, 2 - right*2/n)."""
    n_on_2 = n // 2
    left_size = n_on_2 + n % 2
    left_idx = _iota(left_size)
    right_idx = _add(_iota(n_on_2), left_size)
    left = _div(_mul(left_idx, 2), n)
    right = _sub(2, _div(_mul(right_idx, 2), n))
    return np.concatenate([left, right]).astype(_FT)


def triangular(n: int):
    """windows.ex:98-127."""
    n_on_2 = (n + 1) // 2
    idx = _add(_iota(n_on_2), 1)
    if n % 2 == 1:
        left = _div(_mul(idx, 2), n + 1)
        return np.concatenate([left, left[::-1][1:]]).astype(_FT)
    left = _div(_sub(_mul(2, idx), 1), n)
    return np.concatenate([left, left[::-1]]).astype(_FT)


def blackman(n: int, is_periodic: bool = True):
    """windows.ex:160-199: 0.42 - 0.5 cos(2 pi n/(l-1)) + 0.08 cos(4 pi n/(l-1)),
    built on the left half and mirrored."""
    l = n + 1 if is_periodic else n
    m = -(-l // 2)
    k = _iota(m)
    # 2 * @pi and 4 * @pi are folded by Elixir in double, then become f32 scalars
    a1 = _div(_mul(_lit(2 * PI), k), l - 1)
    a2 = _div(_mul(_lit(4 * PI), k), l - 1)
    left = _add(_sub(_lit(0.42), _mul(_lit(0.5), _cos(a1))), _mul(_lit(0.08), _cos(a2)))
    if l % 2 == 0:
        w = np.concatenate([left, left[::-1]])
    else:
        w = np.concatenate([left, left[::-1][1:]])
    if is_periodic:
        w = w[:-1]
    return w.astype(_FT)


def hamming(n: int, is_periodic: bool = True):
    """windows.ex:225-252: 0.54 - 0.46 cos(2 pi n/(l-1))."""
    l = n + 1 if is_periodic else n
    k = _iota(l)
    a = _div(_mul(_lit(2 * PI), k), l - 1)
    w = _sub(_lit(0.54), _mul(_lit(0.46), _cos(a)))
    return (w[: l - 1] if is_periodic else w).astype(_FT)


def hann(n: int, is_periodic: bool = True):
    """windows.ex:278-305: 0.5 * (1 - cos(2 pi n/(l-1)))."""
    l = n + 1 if is_periodic else n
    k = _iota(l)
    a = _div(_mul(_lit(2 * PI), k), l - 1)
    w = _mul(_lit(0.5), _sub(1, _cos(a)))
    return (w[: l - 1] if is_periodic else w).astype(_FT)


def _kaiser_i0(x):
    """windows.ex:371-386: truncated series below 3.75, asymptotic form above."""
    ax = _f(np.abs(_d(x)))

    def p(e):
        return _f(_d(ax) ** e)

    small = _add(
        _add(_add(_add(1, _div(p(2), 4)), _div(p(4), 64)), _div(p(6), 2304)),
        _div(p(8), 147456),
    )
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        ex = _f(np.exp(_d(ax)))
        den = _f(np.sqrt(_d(_mul(F32(2 * PI), ax))))  # 2 * Nx.Constants.pi(): pi is an f32 tensor by default
        poly = _add(1, _add(_div(1, _mul(8, ax)), _div(9, _mul(128, p(2)))))
        large = _mul(_div(ex, den), poly)
    return np.where(ax < _FT(3.75), small, large).astype(_FT)


def kaiser(n: int, beta: float = 12.0, eps: float = 1.0e-7, is_periodic: bool = True):
    """windows.ex:341-369."""
    l = n + 1 if is_periodic else n
    ratio = nx_linspace(-1, 1, l, endpoint=True)
    sqrt_arg = np.maximum(_sub(1, _f(_d(ratio) ** 2)), _lit(eps)).astype(_FT)
    r = _mul(_lit(beta), _f(np.sqrt(_d(sqrt_arg))))
    with float_type(F32):  # beta is a number: Nx.abs(beta) is an f32 scalar, so I0(beta) is f32 for any `type`
        i0b = _kaiser_i0(np.array([F32(beta)], dtype=F32))
    w = _div(_kaiser_i0(r), i0b)
    return (w[:n] if is_periodic else w).astype(_FT)


# ---------------------------------------------------------------------------
# Nx.fft / Nx.ifft on Nx.BinaryBackend (nx 0.11.0, not vendored)
# ---------------------------------------------------------------------------
def _fft_rec(x: np.ndarray, sign: float) -> np.ndarray:
    """Recursive radix-2 DIT over the last axis, complex128, batch-vectorised."""
    n = x.shape[-1]
    if n <= 1:
        return x
    if n == 2:
        a = x[..., 0:1]
        b = x[..., 1:2]
        return np.concatenate([a + b, a - b], axis=-1)
    if n % 2 == 1:
        # naive DFT: X[k] = sum_j x[j] * exp(sign * 2 pi i j k / n)
        j = np.arange(n, dtype=F64)
        ang = sign * 2.0 * PI * np.outer(j, j) / n
        tw = np.cos(ang) + 1j * np.sin(ang)
        out = np.zeros(x.shape, dtype=C128)
        for jj in range(n):  # sequential accumulation order j = 0..n-1
            out = out + _cmul(x[..., jj : jj + 1], tw[jj])
        return out
    even = _fft_rec(x[..., 0::2], sign)
    odd = _fft_rec(x[..., 1::2], sign)
    k = np.arange(n // 2, dtype=F64)
    t = sign * 2.0 * PI / n
    ang = t * k
    tw = np.cos(ang) + 1j * np.sin(ang)
    bias = _cmul(tw, odd)
    return np.concatenate([even + bias, even - bias], axis=-1)


def _cmul(a, b):
    """complex 0.6.0 multiply: (ac - bd) + (ad + bc) i with plain double ops
    (numpy's complex multiply may use FMA; restate explicitly)."""
    ar, ai = np.real(a), np.imag(a)
    br, bi = np.real(b), np.imag(b)
    return (ar * br - ai * bi) + 1j * (ar * bi + ai * br)


def _fit_length(x: np.ndarray, n: int) -> np.ndarray:
    cur = x.shape[-1]
    if cur == n:
        return x
    if cur > n:
        return x[..., :n]
    pad = [(0, 0)] * (x.ndim - 1) + [(0, n - cur)]
    return np.pad(x, pad)


def _next_pow2(n: int) -> int:
    p = 1
    while p < n:
        p *= 2
    return p


def nx_fft(x, length: Union[int, str, None] = None, eps: float = 1.0e-10, axis: int = -1):
    """Nx.fft on BinaryBackend -> c64."""
    return _nx_fft_impl(x, length, eps, axis, inverse=False)


def nx_ifft(x, length: Union[int, str, None] = None, eps: float = 1.0e-10, axis: int = -1):
    """Nx.ifft on BinaryBackend -> c64."""
    return _nx_fft_impl(x, length, eps, axis, inverse=True)


def _nx_fft_impl(x, length, eps, axis, inverse):
    x = np.asarray(x)
    x = np.moveaxis(x, axis, -1)
    if length is None:
        n = x.shape[-1]
    elif length == "power_of_two":
        n = _next_pow2(x.shape[-1])
    else:
        n = int(length)
    z = _fit_length(x.astype(C128), n)
    if inverse:
        out = _fft_rec(z, +1.0) / n
    else:
        out = _fft_rec(z, -1.0)
    re = np.real(out).copy()
    im = np.imag(out).copy()
    re[np.abs(re) <= eps] = 0.0
    im[np.abs(im) <= eps] = 0.0
    out = (re + 1j * im).astype(C64)
    return np.moveaxis(out, -1, axis)


# ---------------------------------------------------------------------------
# Framing (lib/nx_signal.ex:249-364)
# ---------------------------------------------------------------------------
PaddingT = Union[str, Sequence[Tuple[int, int]]]


def _padding_config(length: int, window_length: int, padding: PaddingT) -> Tuple[int, int]:
    """lib/nx_signal.ex:303-331."""
    if isinstance(padding, str):
        if padding == "valid":
            return (0, 0)
        if padding == "same":
            total = max(length - 1 + window_length - length, 0)
            return (total // 2, -(-total // 2))
        raise ValueError(
            "invalid padding mode specified, padding must be one of :valid, :same, "
            f"or a padding configuration, got: {padding!r}"
        )
    if isinstance(padding, (list, tuple)) and len(padding) == 1:
        lo, hi = padding[0]
        if isinstance(lo, (int, np.integer)) and isinstance(hi, (int, np.integer)):
            return (int(lo), int(hi))
    raise ValueError(
        "padding must be a list of {high, low} tuples, where each element is an integer. "
        f"Got: {padding!r}"
    )


def num_frames(length: int, window_length: int, stride: int, padding: PaddingT = "valid") -> int:
    """Shape rule of Nx.window_max as used at lib/nx_signal.ex:289-298."""
    if padding == "reflect":
        lo = hi = window_length // 2
    else:
        lo, hi = _padding_config(length, window_length, padding)
    padded = length + lo + hi
    if padded < window_length:
        return 0
    return (padded - window_length) // stride + 1


def as_windowed(x, window_length: int, stride: int = 1, padding: PaddingT = "valid"):
    """lib/nx_signal.ex:249-364.  1-D input (or [..., L]: batch stands in for
    Nx vectorised axes); returns [..., M, window_length] of the input dtype."""
    x = np.asarray(x)
    if not (isinstance(stride, (int, np.integer)) and stride >= 1):
        raise ValueError(f"expected an integer >= 1 or a list of integers, got: {stride!r}")
    L = x.shape[-1]
    if padding == "reflect":
        half = window_length // 2
        pad = [(0, 0)] * (x.ndim - 1) + [(half, half)]
        xp = np.pad(x, pad, mode="reflect")  # Nx.reflect == numpy "reflect" (:348-349)
    else:
        lo, hi = _padding_config(L, window_length, padding)
        pad = [(0, 0)] * (x.ndim - 1) + [(lo, hi)]
        xp = np.pad(x, pad)  # Nx.pad with 0 (:338)
    M = num_frames(L, window_length, stride, padding)
    idx = (np.arange(M) * stride)[:, None] + np.arange(window_length)[None, :]
    return xp[..., idx]


def fft_frequencies(sampling_rate, fft_length: int, endpoint: bool = False):
    """lib/nx_signal.ex:154-166: step = sr / nfft (f32 scalars); linspace(0, step * nfft, n: nfft,
    endpoint:, type:) -- the iota * step + start part runs in the graph's float type."""
    sr = _tensor_scalar(sampling_rate)
    step = _f32(_d(sr) / fft_length)
    return nx_linspace(0, _f32(_d(step) * fft_length), fft_length, endpoint=endpoint)


def stft_times(frame_length: int, sampling_rate, num_frames: int):
    """lib/nx_signal.ex:108-111: time_step = N / (2 sr); linspace(time_step,
    time_step * M, n: M); sampling_rate is a defn argument (tensor)."""
    sr = _tensor_scalar(sampling_rate)
    time_step = _div(frame_length, _mul(2, sr))
    last_frame = _mul(time_step, num_frames)
    return nx_linspace(time_step, last_frame, num_frames)


# ---------------------------------------------------------------------------
# STFT / ISTFT (lib/nx_signal.ex:68-130, 582-638, 684-735)
# ---------------------------------------------------------------------------
def _to_f32_or_c64(x):
    x = np.asarray(x)
    if np.iscomplexobj(x):
        return x.astype(C64)
    return x.astype(F64).astype(F32) if x.dtype != F32 else x


def stft(
    data,
    window,
    overlap_length: Optional[int] = None,
    fft_length: Union[int, str] = "power_of_two",
    window_padding: PaddingT = "valid",
    sampling_rate: float = 100,
    scaling: Optional[str] = None,
    fft=None,
):
    """lib/nx_signal.ex:68-130.  data [..., L]; returns (z [..., M, nfft] c64,
    times [M] f32, frequencies [nfft] f32).  ``fft`` lets the large-shape tests
    swap the literal recursive FFT for pocketfft in complex128 (equal to within
    1 f32 ulp for power-of-two nfft)."""
    if sampling_rate is None:
        raise ValueError("missing sampling_rate option")
    if scaling not in (None, "spectrum", "psd"):
        raise ValueError(
            f"invalid :scaling, expected one of :spectrum, :psd or nil, got: {scaling!r}"
        )
    window = np.asarray(window)
    data = np.asarray(data)
    N = window.shape[0]
    if overlap_length is None:
        overlap_length = N // 2
    frames = as_windowed(data, N, N - overlap_length, window_padding)
    # Nx.multiply: int x int stays int (then Nx.fft casts); else f32
    if np.issubdtype(frames.dtype, np.integer) and np.issubdtype(window.dtype, np.integer):
        windowed = frames.astype(np.int64) * window.astype(np.int64)
    elif np.iscomplexobj(frames) or np.iscomplexobj(window):
        # complex data: Nx.multiply(c64, f32) is a complex product in double, rounded once to c64 (:101)
        windowed = _c(np.asarray(frames, dtype=C128) * np.asarray(window, dtype=C128))
    else:
        windowed = _mul(frames, window)
    if fft_length == "power_of_two":
        nfft = _next_pow2(N)
    else:
        nfft = int(fft_length)
    spectrum = (fft or nx_fft)(windowed, nfft)
    M = spectrum.shape[-2]
    freqs = fft_frequencies(sampling_rate, nfft)
    times = stft_times(N, sampling_rate, M)
    wf = _to_f32_or_c64(window)
    if scaling == "spectrum":
        s = _f(np.sum(_d(wf)))
        out = _c(_d(spectrum.real) / _d(s) + 1j * (_d(spectrum.imag) / _d(s)))
    elif scaling == "psd":
        s2 = _f(np.sum(_d(_f(_d(wf) ** 2))))
        den = _f(np.sqrt(_d(_mul(_tensor_scalar(sampling_rate), s2))))
        out = _c(_d(spectrum.real) / _d(den) + 1j * (_d(spectrum.imag) / _d(den)))
    else:
        out = spectrum
    return out, times, freqs


def overlap_and_add(t, overlap_length: int, dtype=None):
    """lib/nx_signal.ex:684-735: [..., M, N] -> [..., M*hop + overlap] with one
    indexed_add (double accumulation, one rounding)."""
    t = np.asarray(t)
    M, N = t.shape[-2], t.shape[-1]
    if overlap_length >= N:
        raise ValueError(
            f"overlap_length must be a number less than the window size {N}, got: {N}"
        )
    hop = N - overlap_length
    out_len = M * hop + overlap_length
    cplx = np.iscomplexobj(t)
    acc = np.zeros(t.shape[:-2] + (out_len,), dtype=C128 if cplx else F64)
    wide = t.astype(C128 if cplx else F64)
    for m in range(M):
        acc[..., m * hop : m * hop + N] += wide[..., m, :]
    out_dtype = dtype or t.dtype
    if np.issubdtype(np.dtype(out_dtype), np.integer):
        return acc.astype(out_dtype)
    return acc.astype(out_dtype)


def istft(
    z,
    window,
    fft_length: Union[int, str, None] = None,
    overlap_length: Optional[int] = None,
    scaling: Optional[str] = None,
    sampling_rate: Optional[float] = 1000,
    ifft=None,
):
    """lib/nx_signal.ex:582-638.  z [..., M, nfft] -> c64 [..., M*hop + N - hop]."""
    z = np.asarray(z)
    window = _to_f32_or_c64(window)
    if scaling == "psd" and sampling_rate is None:
        raise ValueError(":sampling_rate is mandatory if scaling is :psd")
    if scaling not in (None, "spectrum", "psd"):
        raise ValueError(
            f"invalid :scaling, expected one of :spectrum, :psd or nil, got: {scaling!r}"
        )
    if fft_length is None:
        fft_length = "power_of_two"
    if overlap_length is None:
        overlap_length = window.size // 2
    frames = (ifft or nx_ifft)(z, fft_length)
    if scaling == "spectrum":
        s = _d(_f(np.sum(_d(window))))
        frames = _c(_d(frames.real) * s + 1j * (_d(frames.imag) * s))
    elif scaling == "psd":
        s2 = _f(np.sum(_d(_f(_d(window) ** 2))))
        s = _d(_f(np.sqrt(_d(_mul(_tensor_scalar(sampling_rate), s2)))))
        frames = _c(_d(frames.real) * s + 1j * (_d(frames.imag) * s))
    if frames.shape[-1] != window.shape[0]:
        raise ValueError(
            f"cannot broadcast frames of length {frames.shape[-1]} with window of length "
            f"{window.shape[0]} (istft requires fft_length == length(window))"
        )
    w = _d(window)
    fw = _c(_d(frames.real) * w + 1j * (_d(frames.imag) * w))
    result = overlap_and_add(fw, overlap_length)
    w2 = _f(_d(_f(np.abs(_d(window)))) ** 2)
    norm = overlap_and_add(np.broadcast_to(w2, z.shape[:-1] + (w2.shape[0],)), overlap_length)
    norm = np.where(norm > F32(1.0e-10), norm, F32(1.0)).astype(F32)
    out = _c(_d(result.real) / _d(norm) + 1j * (_d(result.imag) / _d(norm)))
    return out


# ---------------------------------------------------------------------------
# FIR design (lib/nx_signal/filters.ex:147-279, waveforms.ex:451-457)
# ---------------------------------------------------------------------------
def sinc(t):
    """waveforms.ex:451-457: t*pi; select(t == 0, 1, sin(t)/t)."""
    t = _mul(_f(t), _lit(PI))
    with np.errstate(divide="ignore", invalid="ignore"):
        s = _div(_sin(t), t)
    return np.where(t == 0, _FT(1.0), s).astype(_FT)


def firwin(
    num_taps: int,
    cutoff,
    window: Union[str, Tuple[str, float]] = "hamming",
    pass_zero: bool = True,
    scale: bool = True,
    sampling_rate: float = 2.0,
):
    """filters.ex:147-252 (f32 output)."""
    nyq = sampling_rate / 2.0
    if not isinstance(cutoff, (list, tuple)):
        raise ValueError(f"cutoff must be a list of frequencies, got: {cutoff!r}")
    cl = sorted(c / nyq for c in cutoff)
    if cl[0] <= 0.0:
        raise ValueError(
            f"cutoff must be strictly between 0 and Nyquist (exclusive), got: {cl[0] * nyq}"
        )
    if cl[-1] >= 1.0:
        raise ValueError(
            f"cutoff must be strictly between 0 and Nyquist (exclusive), got: {cl[-1] * nyq}"
        )
    even_cuts = len(cl) % 2 == 0
    nyquist_gain = (pass_zero and even_cuts) or ((not pass_zero) and (not even_cuts))
    if nyquist_gain and num_taps % 2 == 0:
        raise ValueError(
            "a filter with non-zero gain at Nyquist (e.g. highpass) requires "
            f"an odd number of taps, got: {num_taps}"
        )
    m = (num_taps - 1) / 2.0
    alpha = _sub(_iota(num_taps), _lit(m))
    freqs = [0.0] + cl + [1.0]
    h = np.zeros(num_taps, dtype=_FT)
    for i in range(len(freqs) - 1):
        take = (i % 2 == 0) if pass_zero else (i % 2 == 1)
        if not take:
            continue
        a, b = freqs[i], freqs[i + 1]
        ca = _mul(_lit(a), sinc(_mul(_lit(a), alpha)))
        cb = _mul(_lit(b), sinc(_mul(_lit(b), alpha)))
        h = _sub(_add(h, cb), ca)
    w = _firwin_window(num_taps, window)
    h = _mul(h, w)
    if not scale:
        return h
    if pass_zero:
        sf = 0.0
    elif len(cl) == 1:
        sf = 1.0
    else:
        sf = (cl[0] + cl[1]) / 2.0
    c = _cos(_mul(alpha, _lit(PI * sf)))
    s = _f(np.abs(np.sum(_d(h) * _d(c))))
    return _div(h, s)


def _firwin_window(num_taps, window):
    """filters.ex:254-279."""
    if window == "hamming":
        return hamming(num_taps, is_periodic=False)
    if window == "hann":
        return hann(num_taps, is_periodic=False)
    if window == "blackman":
        return blackman(num_taps, is_periodic=False)
    if window == "bartlett":
        return bartlett(num_taps)
    if window == "rectangular":
        return rectangular(num_taps, dtype=_FT)
    if isinstance(window, tuple) and len(window) == 2 and window[0] == "kaiser":
        return kaiser(num_taps, beta=window[1], is_periodic=False)
    raise ValueError(
        f"unknown window {window!r}, supported: "
        ":hamming, :hann, :blackman, :bartlett, :rectangular, {:kaiser, beta}"
    )


# ---------------------------------------------------------------------------
# Convolution (lib/nx_signal/convolution.ex)
# ---------------------------------------------------------------------------
def _result_type(a, b):
    return C64 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else F32


def _check_mode_method(mode, method):
    if mode not in ("full", "same", "valid"):
        raise ValueError(f"expected mode to be one of [:full, :same, :valid], got: :{mode}")
    if method not in ("direct", "fft"):
        raise ValueError(f"expected method to be one of [:direct, :fft], got: :{method}")


def convolve(in1, in2, mode: str = "full", method: str = "direct"):
    """convolution.ex:38-58."""
    _check_mode_method(mode, method)
    if method == "direct":
        return direct_convolve(in1, in2, mode)
    return fftconvolve(in1, in2, mode)


def correlate(in1, in2, mode: str = "full", method: str = "direct"):
    """convolution.ex:87-93: convolve with reversed (conjugated) kernel."""
    in2 = np.asarray(in2)
    rev = in2[tuple(slice(None, None, -1) for _ in range(in2.ndim))]
    if np.iscomplexobj(in2):
        rev = np.conj(rev)
    return convolve(in1, rev, mode=mode, method=method)


def direct_convolve(in1, in2, mode: str = "full"):
    """convolution.ex:95-223: Nx.conv(volume, reverse(kernel)) with mode padding;
    double accumulation, one rounding (f32 / c64)."""
    a = np.asarray(in1)
    b = np.asarray(in2)
    if a.ndim == 0 and b.ndim == 0:
        rank = 0
    elif a.ndim == 0 or b.ndim == 0:
        raise ValueError(f"Incompatible ranks: {{{a.ndim}, {b.ndim}}}")
    elif a.ndim == b.ndim:
        rank = a.ndim
    else:
        raise ValueError(
            "NxSignal.convolve/3 requires both inputs to have the same rank or one of them "
            f"to be a scalar, got {a.ndim} and {b.ndim}"
        )
    if mode == "valid":
        ok1 = all(i >= j for i, j in zip(a.shape, b.shape))
        ok2 = all(i <= j for i, j in zip(a.shape, b.shape))
        if ok1:
            pass
        elif ok2:
            a, b = b, a
        else:
            raise ValueError(
                "For :valid mode, one must be at least as large as the other in every dimension"
            )
    out_t = _result_type(a, b)
    wide = C128 if out_t is C64 else F64
    a = a.reshape((1,) * max(1 - a.ndim, 0) + a.shape).astype(wide)
    b = b.reshape((1,) * max(1 - b.ndim, 0) + b.shape).astype(wide)
    nd = a.ndim
    # full convolution by shifted accumulation (exact same sum set as Nx.conv)
    full_shape = tuple(i + j - 1 for i, j in zip(a.shape, b.shape))
    full = np.zeros(full_shape, dtype=wide)
    for idx in np.ndindex(*b.shape):
        sl = tuple(slice(i, i + s) for i, s in zip(idx, a.shape))
        full[sl] += a * b[idx]
    if mode == "full":
        out = full
    elif mode == "same":
        # pad_left = (k-1) - (k-1)//2 -> output n uses full index n + (k-1)//2
        sl = tuple(slice((k - 1) // 2, (k - 1) // 2 + n) for n, k in zip(a.shape, b.shape))
        out = full[sl]
    else:
        sl = tuple(slice(k - 1, n) for n, k in zip(a.shape, b.shape))
        out = full[sl]
    out = out.astype(out_t)
    if rank == 0:
        return out.reshape(())
    return out.reshape(out.shape[nd - rank :]) if rank < nd else out


def fftconvolve(in1, in2, mode: str = "full"):
    """convolution.ex:252-329: full-length FFT per axis where both dims != 1."""
    a = np.asarray(in1)
    b = np.asarray(in2)
    if a.ndim != b.ndim or a.ndim == 0:
        raise ValueError("Rank of in1 and in2 must be equal.")
    if mode not in ("full", "same", "valid"):
        raise ValueError(f"expected mode to be one of [:full, :same, :valid], got: :{mode}")
    s1, s2 = list(a.shape), list(b.shape)
    lengths = [x + y - 1 for x, y in zip(s1, s2)]
    axes = [i for i in range(a.ndim) if s1[i] != 1 and s2[i] != 1]
    sp1 = _to_f32_or_c64(a)
    sp2 = _to_f32_or_c64(b)
    for ax in axes:  # transforms.ex:5-12
        sp1 = nx_fft(sp1, lengths[ax], axis=ax)
        sp2 = nx_fft(sp2, lengths[ax], axis=ax)
    sp1 = np.asarray(sp1, dtype=C64)
    sp2 = np.asarray(sp2, dtype=C64)
    c = _c(_cmul(sp1.astype(C128), sp2.astype(C128)))
    out = c
    for ax in axes:  # transforms.ex:14-21
        out = nx_ifft(out, None, axis=ax)
    if _result_type(a, b) is F32:
        out = out.real.astype(F32)
    if mode == "full":
        return out
    if mode == "same":
        return _centered(out, s1)
    ok1 = all(x >= y for x, y in zip(s1, s2))
    ok2 = all(y >= x for x, y in zip(s1, s2))
    if ok1:
        big, small = s1, s2
    elif ok2:
        big, small = s2, s1
    else:
        raise ValueError(
            "For 'valid' mode, one must be at least as large as the other in every dimension."
        )
    return _centered(out, [x - y + 1 for x, y in zip(big, small)])


def _centered(out, new_shape):
    """convolution.ex:319-329."""
    sl = tuple(
        slice((cur - new) // 2, (cur - new) // 2 + new) for cur, new in zip(out.shape, new_shape)
    )
    return out[sl]


# ---------------------------------------------------------------------------
# Fast large-shape variants (pocketfft in complex128 + one rounding).
# For power-of-two nfft these equal the literal restatement to within 1 f32 ulp
# (checked in tests/test_oracle_golden.py); used where the recursive FFT would
# take minutes.
# ---------------------------------------------------------------------------
def _fast_fft(x, n, eps=1.0e-10):
    import scipy.fft as sfft

    out = sfft.fft(np.asarray(x).astype(C128), n=n, axis=-1)
    re = out.real
    im = out.imag
    re[np.abs(re) <= eps] = 0.0
    im[np.abs(im) <= eps] = 0.0
    return out.astype(C64)


def _fast_ifft(x, n, eps=1.0e-10):
    import scipy.fft as sfft

    x = np.asarray(x)
    if n == "power_of_two":
        n = _next_pow2(x.shape[-1])
    out = sfft.ifft(x.astype(C128), n=n, axis=-1)
    re = out.real
    im = out.imag
    re[np.abs(re) <= eps] = 0.0
    im[np.abs(im) <= eps] = 0.0
    return out.astype(C64)


def stft_fast(data, window, **kw):
    return stft(data, window, fft=_fast_fft, **kw)


def istft_fast(z, window, **kw):
    return istft(z, window, ifft=_fast_ifft, **kw)


def fir_same_f64(x, taps):
    """convolve(x, taps, mode: :same) for x [C, L], taps [K] in double via
    scipy.signal.oaconvolve, rounded once to f32 (the value the reference's
    full-length FFT path approximates; used for large FIR shapes)."""
    from scipy.signal import oaconvolve

    x = np.atleast_2d(np.asarray(x, dtype=F64))
    t = np.asarray(taps, dtype=F64)[None, :]
    full = oaconvolve(x, t, mode="full", axes=-1)
    K = t.shape[-1]
    L = x.shape[-1]
    start = (full.shape[-1] - L) // 2
    return full[..., start : start + L].astype(F32)


# ---------------------------------------------------------------------------
# Mel (lib/nx_signal.ex:397-517).  Outside the hot path (SURVEY 8f "next");
# restated only because its doctest is the one reference vector that pins a
# reflect-padded STFT with fft_length > frame_length (lib/nx_signal.ex:465-483).
# ---------------------------------------------------------------------------
def mel_filters(fft_length, mel_bins, sampling_rate, max_mel=3016, mel_frequency_spacing=200 / 3):
    """lib/nx_signal.ex:397-445."""
    f_sp = mel_frequency_spacing
    fftfreqs = fft_frequencies(sampling_rate, fft_length)
    mels = nx_linspace(0, max_mel / f_sp, mel_bins + 2)
    freqs = _mul(_lit(f_sp), mels)
    min_log_hz = 1000
    min_log_mel = min_log_hz / f_sp
    logstep = _div(_f(np.log(_d(_lit(6.4)))), 27)
    log_t = mels >= _lit(min_log_mel)
    e = _f(np.exp(_d(_mul(logstep, _sub(mels, _lit(min_log_mel))))))
    mel_f = np.where(log_t, _mul(min_log_hz, e), freqs).astype(F32)
    fdiff = _sub(mel_f[1:], mel_f[:-1])[:, None]
    ramps = _sub(mel_f[:, None], fftfreqs[None, :])
    lower = _div(_f(-_d(ramps[:mel_bins])), fdiff[:mel_bins])
    upper = _div(ramps[2 : mel_bins + 2], fdiff[1 : mel_bins + 1])
    weights = np.maximum(F32(0), np.minimum(lower, upper)).astype(F32)
    # Nx.max(0, x) is :erlang.max(0.0, x), which returns its first argument on a tie: +0.0 for x = -0.0
    # (numpy returns the second); the doctest prints 0.0 at [0][0] (lib/nx_signal.ex:388)
    weights = np.where(weights == 0, F32(0), weights).astype(F32)
    enorm = _div(2.0, _sub(mel_f[2 : mel_bins + 2], mel_f[:mel_bins]))
    return _mul(weights, enorm[:, None])


def stft_to_mel(z, sampling_rate, fft_length, mel_bins=128, **kw):
    """lib/nx_signal.ex:486-513."""
    z = np.asarray(z, dtype=C64)
    mag = _f(np.hypot(_d(z.real), _d(z.imag)))
    mag = _f(_d(mag) ** 2)
    filters = mel_filters(fft_length, mel_bins, sampling_rate, **kw)
    half = fft_length // 2
    mel_spec = _f(_d(mag[..., :half]) @ _d(filters[:, :half]).T)
    clipped = np.maximum(mel_spec, F32(1.0e-10))
    log_spec = _div(_f(np.log(_d(clipped))), _f(np.log(_d(F32(10)))))
    log_spec = np.maximum(log_spec, _sub(log_spec.max(), 8)).astype(F32)
    return _div(_add(log_spec, 4), 4)


# ---------------------------------------------------------------------------
# Spectrogram-adjacent ops: Filters.median / Filters.wiener
# (lib/nx_signal/filters.ex:17-110, 281-303) and PeakFinding.argrel*
# (lib/nx_signal/peak_finding.ex:131-391).  Pinned by the reference's own test
# vectors (test/nx_signal/filters_test.exs:5-245, peak_finding_test.exs, the
# doctests at filters.ex:68-78 and peak_finding.ex:36-128, 296-331).
# ---------------------------------------------------------------------------
def median(t, kernel_shape):
    """filters.ex:17-56.  out[i] = Nx.median of the window that STARTS at i (Nx.slice clamps the
    start so the window stays inside the tensor: start = min(i, dim - k) per axis), as f32."""
    t = np.asarray(t)
    kernel_shape = tuple(int(k) for k in kernel_shape)
    if t.ndim != len(kernel_shape):
        raise ValueError("kernel shape must be of the same rank as the tensor")
    tf = t.astype(F64)
    out = np.empty(t.shape, dtype=F32)
    n = int(np.prod(kernel_shape))
    for idx in np.ndindex(*t.shape):
        sl = tuple(slice(min(i, d - k), min(i, d - k) + k) for i, d, k in zip(idx, t.shape, kernel_shape))
        win = np.sort(tf[sl].reshape(-1))
        if n % 2:
            out[idx] = F32(win[n // 2])
        else:
            out[idx] = F32((win[n // 2 - 1] + win[n // 2]) / 2.0)
    return out


def _local_sum_same(a, kernel_size):
    """correlate(a, ones(kernel_size), mode: :same) in double: the window of output i covers
    [i - (k - 1) + (k - 1)//2, i + (k - 1)//2] per axis, zeros outside (convolution.ex:95-211)."""
    a = np.asarray(a, dtype=F64)
    full_shape = tuple(n + k - 1 for n, k in zip(a.shape, kernel_size))
    full = np.zeros(full_shape, dtype=F64)
    for idx in np.ndindex(*kernel_size):
        sl = tuple(slice(i, i + s) for i, s in zip(idx, a.shape))
        full[sl] += a
    sl = tuple(slice((k - 1) // 2, (k - 1) // 2 + n) for n, k in zip(a.shape, kernel_size))
    return full[sl]


def wiener(t, kernel_size=3, noise=None):
    """filters.ex:80-110, 281-303: computed in f64 (Nx.as_type(:f64)), cast back to t's type."""
    t = np.asarray(t)
    out_t = t.dtype  # |> Nx.as_type(Nx.type(t)), filters.ex:108-110: integer inputs get the truncated result back
    if isinstance(kernel_size, (int, np.integer)):
        kernel_size = (int(kernel_size),) * t.ndim
    elif not isinstance(kernel_size, tuple):
        raise ValueError("kernel_size must be an integer or tuple")
    size = float(np.prod(kernel_size))
    x = t.astype(F64)
    l_mean = _local_sum_same(x, kernel_size) / size
    l_var = _local_sum_same(x ** 2, kernel_size) / size - l_mean ** 2
    nz = F64(np.mean(l_var)) if noise is None else F64(noise)
    with np.errstate(divide="ignore", invalid="ignore"):
        res = (x - l_mean) * (1.0 - nz / l_var)
    return np.where(l_var < nz, l_mean, res + l_mean).astype(out_t)


def argrelextrema(data, comparator="greater", axis=0, order=1):
    """peak_finding.ex:339-391.  comparator in {'less', 'greater', 'less_equal', 'greater_equal'}
    (the reference takes any arity-2 function).  Returns (indices s32 {n, rank} with -1 rows after
    the valid ones, valid_indices)."""
    data = np.asarray(data)
    cmp = {"less": np.less, "greater": np.greater, "less_equal": np.less_equal, "greater_equal": np.greater_equal}[comparator]
    n = data.shape[axis]
    locs = np.arange(n)
    results = np.ones(data.shape, dtype=bool)
    shift = 1
    while shift < order + 1:
        plus = np.take(data, np.clip(locs + shift, 0, n - 1), axis=axis)
        minus = np.take(data, np.clip(locs - shift, 0, n - 1), axis=axis)
        results &= cmp(data, plus)
        results &= cmp(data, minus)
        if not results.any():
            break
        shift += 1
    flat = results.reshape(-1)
    idx = np.stack(np.unravel_index(np.arange(flat.size), data.shape), axis=-1).astype(np.int32)
    valid = idx[flat]
    out = np.full((flat.size, data.ndim), -1, dtype=np.int32)
    out[: valid.shape[0]] = valid
    return out, int(flat.sum())


def argrelmin(data, axis=0, order=1):
    return argrelextrema(data, "less", axis=axis, order=order)


def argrelmax(data, axis=0, order=1):
    return argrelextrema(data, "greater", axis=axis, order=order)
