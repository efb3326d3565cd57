"""NxSignal.Convolution (lib/nx_signal/convolution.ex): convolve / correlate / fftconvolve.

The batched FIR form -- ``x {C, L}`` with ``h {1, K}`` (or both rank-1) -- runs the
overlap-save FFT kernel (``nxs_fir_f32``); every other shape (N-d, complex, scalars) runs
the general direct kernel (``nxs_convolve_nd``).  ``method`` is accepted for API parity:
both of the reference's methods compute the same values, the GPU picks by shape."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _arrays as A
from . import _lib

# below this many multiply-adds per output sample the direct kernel wins
_FIR_MIN_TAPS = 32


def _check(mode, method):
    if mode not in ("full", "same", "valid"):
        raise _lib.NxSignalArgumentError(
            f"expected mode to be one of [:full, :same, :valid], got: :{mode}")
    if method not in ("direct", "fft"):
        raise _lib.NxSignalArgumentError(
            f"expected method to be one of [:direct, :fft], got: :{method}")


def _is_complex(x):
    return x.is_complex() if A.is_torch(x) else np.iscomplexobj(x)


def _ranks(a, b, fft):
    ra, rb = a.ndim, b.ndim
    if fft:
        if ra != rb or ra == 0:
            raise _lib.NxSignalArgumentError("Rank of in1 and in2 must be equal.")
        return ra
    if ra == 0 and rb == 0:
        return 0
    if ra == 0 or rb == 0:
        raise _lib.NxSignalArgumentError(f"Incompatible ranks: {{{ra}, {rb}}}")
    if ra != rb:
        raise _lib.NxSignalArgumentError(
            "NxSignal.convolve/3 requires both inputs to have the same rank or one of them "
            f"to be a scalar, got {ra} and {rb}")
    return ra


def convolve(in1, in2, mode="full", method="direct"):
    """convolution.ex:38-58."""
    _check(mode, method)
    a = in1 if A.is_torch(in1) else np.asarray(in1)
    b = in2 if A.is_torch(in2) else np.asarray(in2)
    rank = _ranks(a, b, method == "fft")
    if rank > 3:
        return _batched_high_rank(a, b, mode, method, rank)
    sa = (1,) * (3 - a.ndim) + tuple(int(s) for s in a.shape)
    sb = (1,) * (3 - b.ndim) + tuple(int(s) for s in b.shape)
    if mode == "valid":
        ok1 = all(i >= j for i, j in zip(sa, sb))
        ok2 = all(i <= j for i, j in zip(sa, sb))
        if not (ok1 or ok2):
            raise _lib.NxSignalArgumentError(
                "For :valid mode, one must be at least as large as the other in every dimension")
    cplx = _is_complex(a) or _is_complex(b)
    # batched FIR: x {.., C, L} (*) h {.., 1, K}, real
    if (not cplx and sb[0] == 1 and sb[1] == 1 and sb[2] >= _FIR_MIN_TAPS and sa[2] >= sb[2]
            and A.is_cuda(a) == A.is_cuda(b)):
        return _fir(a, b, sa, sb, mode, rank)
    return _direct(a, b, sa, sb, cplx, mode, rank)


def _batched_high_rank(a, b, mode, method, rank):
    """rank > 3 (the reference's direct_convolve is rank-generic, convolution.ex:95-211): served when every
    axis of in2 before its last three has size 1 -- those axes are batch axes of in1 (a size-1 kernel axis
    leaves the axis unchanged in all three modes), so the leading axes fold and each slice is a 3-d problem."""
    lead_a, lead_b = tuple(a.shape[:-3]), tuple(b.shape[:-3])
    if any(d != 1 for d in lead_b):
        raise NotImplementedError("convolve: rank > 3 needs size-1 leading axes on in2 (true > 3-d kernels are not "
                                  "supported by this backend)")
    b3 = b.reshape(tuple(b.shape[-3:]))
    a4 = a.reshape((-1,) + tuple(a.shape[-3:]))
    outs = [convolve(a4[i], b3, mode=mode, method=method) for i in range(a4.shape[0])]
    if A.is_torch(outs[0]):
        import torch

        y = torch.stack(outs, dim=0)
    else:
        y = np.stack(outs, axis=0)
    return y.reshape(lead_a + tuple(y.shape[1:]))


def correlate(in1, in2, mode="full", method="direct"):
    """convolution.ex:87-93: convolve with the reversed (conjugated) second operand."""
    if A.is_torch(in2):
        import torch

        b = torch.flip(in2, dims=list(range(in2.ndim)))
        if in2.is_complex():
            b = b.conj().resolve_conj()
    else:
        b = np.asarray(in2)
        b = b[tuple(slice(None, None, -1) for _ in range(b.ndim))]
        if np.iscomplexobj(b):
            b = np.conj(b)
    return convolve(in1, b, mode=mode, method=method)


def fftconvolve(in1, in2, mode="full", method="direct"):
    """convolution.ex:252-298 (same values as convolve; equal ranks required)."""
    return convolve(in1, in2, mode=mode, method="fft")


def _out_shape(sa, sb, mode):
    if mode == "full":
        return tuple(i + j - 1 for i, j in zip(sa, sb))
    if mode == "same":
        return tuple(sa)
    return tuple(abs(i - j) + 1 for i, j in zip(sa, sb))


def _fir(a, b, sa, sb, mode, rank):
    x = A.to_real_f32(a).reshape(sa[0] * sa[1], sa[2])
    taps = A.like_device(x, A.to_real_f32(b).reshape(-1))
    Cn, L, K = x.shape[0], x.shape[1], int(taps.shape[0])
    os_ = _out_shape(sa, sb, mode)
    y = A.empty_like_kind(x, (Cn, os_[2]), "f32")
    ctx = _lib.context(A.device_index(x))
    l = _lib.lib()
    if A.is_cuda(x):
        rc = l.nxs_fir_f32_dev(ctx, A.ptr(x), Cn, L, L, A.ptr(taps), K, _lib.MODE[mode], A.ptr(y), os_[2],
                               A.stream_of(x))
    else:
        rc = l.nxs_fir_f32_host(ctx, A.ptr(x), Cn, L, L, A.ptr(taps), K, _lib.MODE[mode], A.ptr(y), os_[2])
    _lib.check(rc, ctx, "convolve")
    return y.reshape(os_[3 - rank:] if rank else ())


def _direct(a, b, sa, sb, cplx, mode, rank):
    if cplx:
        xa, xb = A.to_c64(a), A.to_c64(b)
    else:
        xa, xb = A.to_real_f32(a), A.to_real_f32(b)
    if A.is_cuda(xa) != A.is_cuda(xb) or A.is_torch(xa) != A.is_torch(xb):
        xb = A.move_like(xa, xb)  # operands of either family (numpy / torch, host / device) end up next to in1
    os_ = _out_shape(sa, sb, mode)
    out = A.empty_like_kind(xa, os_, "c64" if cplx else "f32")
    ctx = _lib.context(A.device_index(xa))
    l = _lib.lib()
    s1 = (C.c_int64 * 3)(*sa)
    s2 = (C.c_int64 * 3)(*sb)
    if A.is_cuda(xa):
        rc = l.nxs_convolve_nd_dev(ctx, A.ptr(xa), s1, A.ptr(xb), s2, int(cplx), _lib.MODE[mode], A.ptr(out),
                                   A.stream_of(xa))
    else:
        rc = l.nxs_convolve_nd_host(ctx, A.ptr(xa), s1, A.ptr(xb), s2, int(cplx), _lib.MODE[mode], A.ptr(out))
    _lib.check(rc, ctx, "convolve")
    return out.reshape(os_[3 - rank:] if rank else ())
