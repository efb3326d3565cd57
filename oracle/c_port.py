"""ctypes wrapper of oracle/libnxs_oracle.so (test infrastructure; see nxs_oracle.c)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libnxs_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            import subprocess

            subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
        l = C.CDLL(_PATH)
        l.nxs_oracle_threads.restype = C.c_int
        l.nxs_oracle_stft_f32.restype = C.c_int
        l.nxs_oracle_stft_f32.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                          C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_int,
                                          C.c_double, C.c_void_p, C.c_int]
        _lib = l
    return _lib


def threads():
    return int(lib().nxs_oracle_threads())


def stft(x, window, hop, nfft, pad_lo=0, pad_hi=0, reflect=False, scaling=None, sampling_rate=100.0, nthreads=0,
         out=None):
    """x [C, L] f32 -> z [C, M, nfft] c64 (valid / zero / reflect padding given as lo, hi).
    `out`: a preallocated C-contiguous complex64 [C, M, nfft] array to write into (timed loops reuse it
    so that page faults of a fresh result are not charged to the transform)."""
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float32)
    w = np.ascontiguousarray(window, dtype=np.float32)
    Cn, L = x.shape
    N = w.shape[0]
    padded = L + pad_lo + pad_hi
    M = 0 if padded < N else (padded - N) // hop + 1
    if out is None:
        z = np.empty((Cn, M, nfft), dtype=np.complex64)
    else:
        z = out
        assert z.shape == (Cn, M, nfft) and z.dtype == np.complex64 and z.flags.c_contiguous
    sc = {None: 0, "spectrum": 1, "psd": 2}[scaling]
    rc = lib().nxs_oracle_stft_f32(x.ctypes.data, Cn, L, L, w.ctypes.data, N, hop, nfft, pad_lo, int(reflect), M,
                                   sc, float(sampling_rate), z.ctypes.data, int(nthreads))
    assert rc == 0
    return z
