"""NxSignal.PeakFinding (lib/nx_signal/peak_finding.ex:131-391): argrelmin / argrelmax /
argrelextrema on the device (csrc/nxs_post.cu).

Returns the reference's map as a dict: ``indices`` s32 ``{n, rank}`` (the extrema's multi-indices
in row-major order, then rows of -1) and ``valid_indices`` (their count).  The reference takes an
arbitrary comparator function; here it is one of 'less', 'greater', 'less_equal',
'greater_equal' (``argrelmin`` = 'less', ``argrelmax`` = 'greater').  Data is compared as f32."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _arrays as A
from . import _lib


def argrelextrema(data, comparator, axis=0, order=1):
    """peak_finding.ex:339-391."""
    if comparator not in _lib.CMP:
        raise NotImplementedError(
            f"argrelextrema: comparator must be one of {sorted(_lib.CMP)}, got {comparator!r} "
            "(arbitrary comparator functions are not supported by this backend)")
    x = A.to_real_f32(data, "data")
    rank = x.ndim
    if rank < 1 or rank > 8:
        raise NotImplementedError("argrelextrema: rank must be between 1 and 8")
    if not -rank <= int(axis) < rank:
        raise _lib.NxSignalArgumentError(f"given axis ({axis}) invalid for shape with rank {rank}")
    axis = int(axis) % rank
    if int(order) < 0:
        raise _lib.NxSignalArgumentError(f"order must be a non-negative integer, got: {order!r}")
    shape = tuple(int(s) for s in x.shape)
    total = int(np.prod(shape, dtype=np.int64))
    if int(order) == 0:
        # the reference's while loop (peak_finding.ex:354-363) runs no comparison when order = 0: every element
        # stays marked, so `nonzero` lists all multi-indices in row-major order
        idx = np.stack(np.unravel_index(np.arange(total), shape), axis=-1).astype(np.int32).reshape(total, rank)
        if A.is_cuda(x):
            import torch

            return {"indices": torch.from_numpy(idx).to(x.device),
                    "valid_indices": torch.tensor(total, dtype=torch.int64, device=x.device)}
        return {"indices": idx, "valid_indices": np.int64(total)}
    shp = (C.c_int64 * rank)(*shape)
    ctx = _lib.context(A.device_index(x))
    if A.is_cuda(x):
        import torch

        idx = torch.empty((total, rank), dtype=torch.int32, device=x.device)
        cnt = torch.zeros((), dtype=torch.int64, device=x.device)
        rc = _lib.lib().nxs_argrelextrema_f32_dev(ctx, A.ptr(x), rank, shp, axis, int(order), _lib.CMP[comparator],
                                                  A.ptr(idx), A.ptr(cnt), A.stream_of(x))
        _lib.check(rc, ctx, "PeakFinding.argrelextrema")
        return {"indices": idx, "valid_indices": cnt}
    idx = np.empty((total, rank), dtype=np.int32)
    cnt = C.c_int64(0)
    rc = _lib.lib().nxs_argrelextrema_f32_host(ctx, A.ptr(x), rank, shp, axis, int(order), _lib.CMP[comparator],
                                               A.ptr(idx), C.byref(cnt))
    _lib.check(rc, ctx, "PeakFinding.argrelextrema")
    return {"indices": idx, "valid_indices": np.int64(cnt.value)}  # one dtype on both paths (int64, like the device entry)


def argrelmin(data, axis=0, order=1):
    """peak_finding.ex:131-135: relative minima (&Nx.less/2)."""
    return argrelextrema(data, "less", axis=axis, order=order)


def argrelmax(data, axis=0, order=1):
    """peak_finding.ex:251-255: relative maxima (&Nx.greater/2)."""
    return argrelextrema(data, "greater", axis=axis, order=order)
