"""Buffer plumbing between callers' arrays and the C ABI: numpy arrays go through the
``_host`` entry points, CUDA torch tensors through the ``_dev`` ones (torch is only the
owner of device memory and streams here)."""
from __future__ import annotations

import ctypes as C

import numpy as np


def is_torch(x):
    return hasattr(x, "data_ptr") and hasattr(x, "device")


def is_cuda(x):
    return is_torch(x) and x.device.type == "cuda"


def ptr(x):
    if x is None:
        return None
    if is_torch(x):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(x.ctypes.data)


def to_real_f32(x, name="data"):
    """ints / floats -> contiguous f32 (the reference promotes the same way, lib/nx_signal.ex:46-55)."""
    if is_torch(x):
        import torch

        if x.is_complex():
            raise NotImplementedError(f"{name}: complex input is not supported by this backend")
        return x.to(torch.float32).contiguous()
    a = np.asarray(x)
    if np.iscomplexobj(a):
        raise NotImplementedError(f"{name}: complex input is not supported by this backend")
    return np.ascontiguousarray(a, dtype=np.float32)


def to_c64(x):
    if is_torch(x):
        import torch

        return x.to(torch.complex64).contiguous()
    return np.ascontiguousarray(np.asarray(x), dtype=np.complex64)


def like_device(x, arr):
    """Move a small host array next to x (window / taps)."""
    if is_cuda(x):
        import torch

        if is_torch(arr):
            return arr.to(device=x.device, dtype=torch.float32).contiguous()
        return torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(x.device)
    if is_torch(arr):
        arr = arr.detach().cpu().numpy()
    return np.ascontiguousarray(arr, dtype=np.float32)


def move_like(x, arr):
    """arr (numpy or torch, any dtype) moved next to x, dtype unchanged."""
    if is_cuda(x):
        import torch

        if is_torch(arr):
            return arr.to(x.device).contiguous()
        return torch.from_numpy(np.ascontiguousarray(arr)).to(x.device)
    if is_torch(arr):
        return np.ascontiguousarray(arr.detach().cpu().numpy())
    return np.ascontiguousarray(arr)


def empty_like_kind(x, shape, dtype):
    """dtype in {'f32','c64', np dtype}; allocates where x lives."""
    if is_cuda(x):
        import torch

        td = {"f32": torch.float32, "c64": torch.complex64}.get(dtype, None)
        if td is None:
            td = getattr(torch, np.dtype(dtype).name)
        return torch.empty(shape, dtype=td, device=x.device)
    nd = {"f32": np.float32, "c64": np.complex64}.get(dtype, dtype)
    return np.empty(shape, dtype=nd)


def from_host(x, arr):
    """Return a small host-computed numpy array in the caller's array family."""
    if is_torch(x):
        import torch

        return torch.from_numpy(arr).to(x.device)
    return arr


def stream_of(x):
    import torch

    return C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)


def device_index(x):
    if is_cuda(x):
        idx = x.device.index
        if idx is None:
            import torch

            idx = torch.cuda.current_device()
        return idx
    return 0
