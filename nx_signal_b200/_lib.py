"""ctypes binding of libnxsignal_b200.so (C ABI in include/nxsignal_b200.h).

There is no CPU implementation behind this module: if the shared library is missing the
import of any compute entry fails loudly, and if there is no CUDA device
``nxs_ctx_create`` returns NXS_ENODEVICE which is raised as RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libnxsignal_b200.so")

NXS_OK = 0
NXS_EINVAL, NXS_ESHAPE, NXS_EUNSUPPORTED, NXS_ECUDA, NXS_ENCCL, NXS_ENOMEM, NXS_ENODEVICE = (
    -1, -2, -3, -4, -5, -6, -7)

PAD_VALID, PAD_SAME, PAD_REFLECT, PAD_EXPLICIT = 0, 1, 2, 3
SCALE_NONE, SCALE_SPECTRUM, SCALE_PSD = 0, 1, 2
WIN = {"rectangular": 0, "bartlett": 1, "triangular": 2, "blackman": 3, "hamming": 4, "hann": 5, "kaiser": 6}
MODE = {"full": 0, "same": 1, "valid": 2}
CMP = {"less": 0, "greater": 1, "less_equal": 2, "greater_equal": 3}

i64, f64, i32 = C.c_int64, C.c_double, C.c_int
vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/nxsignal_b200.h declares
SIGNATURES = {
    "nxs_abi_version": (i32, []),
    "nxs_build_info": (C.c_char_p, []),
    "nxs_strerror": (C.c_char_p, [i32]),
    "nxs_device_count": (i32, []),
    "nxs_ctx_create": (i32, [i32, C.POINTER(vp)]),
    "nxs_ctx_destroy": (i32, [vp]),
    "nxs_last_error": (C.c_char_p, [vp]),
    "nxs_ctx_synchronize": (i32, [vp]),
    "nxs_ctx_launch_count": (C.c_uint64, [vp]),
    "nxs_ctx_profile": (i32, [vp, i32]),
    "nxs_ctx_profile_read": (i32, [vp, C.POINTER(f64), C.POINTER(i64)]),
    "nxs_ctx_host_timeline": (i32, [vp, C.POINTER(f64 * 4)]),
    "nxs_ctx_set_host_mode": (i32, [vp, i32]),
    "nxs_ctx_host_mode": (i32, [vp, C.POINTER(i32)]),
    "nxs_window_f32": (i32, [i32, i64, i32, f64, f64, vp]),
    "nxs_window_f64": (i32, [i32, i64, i32, f64, f64, vp]),
    "nxs_firwin_f64": (i32, [i64, C.POINTER(f64), i32, i32, f64, i32, i32, f64, vp]),
    "nxs_fft_frequencies_ex": (i32, [f64, i64, i32, i32, vp]),
    "nxs_firwin_f32": (i32, [i64, C.POINTER(f64), i32, i32, f64, i32, i32, f64, vp]),
    "nxs_fft_frequencies_f32": (i32, [f64, i64, vp]),
    "nxs_mel_filters_f32": (i32, [i64, i64, f64, f64, f64, vp]),
    "nxs_stft_to_mel_f32_dev": (i32, [vp, vp, i64, i64, i64, i64, i64, f64, f64, f64, vp, vp]),
    "nxs_stft_to_mel_f32_host": (i32, [vp, vp, i64, i64, i64, i64, i64, f64, f64, f64, vp]),
    "nxs_stft_mel_f32_dev": (i32, [vp, vp, i64, i64, i64, vp, i64, i64, i64, i32, i64, i64, i32, f64, i64, f64, f64, vp, vp]),
    "nxs_stft_mel_f32_host": (i32, [vp, vp, i64, i64, i64, vp, i64, i64, i64, i32, i64, i64, i32, f64, i64, f64, f64, vp]),
    "nxs_stft_times_f32": (i32, [i64, f64, i64, vp]),
    "nxs_num_frames": (i32, [i64, i64, i64, i32, i64, i64, C.POINTER(i64)]),
    "nxs_stft_f32_dev": (i32, [vp, vp, i64, i64, i64, vp, i64, i64, i64, i32, i64, i64, i32, f64, vp, vp]),
    "nxs_stft_c64_dev": (i32, [vp, vp, i64, i64, i64, vp, i64, i64, i64, i32, i64, i64, i32, f64, vp, vp]),
    "nxs_stft_c64_host": (i32, [vp, vp, i64, i64, i64, vp, i64, i64, i64, i32, i64, i64, i32, f64, vp]),
    "nxs_stft_onesided_f32_dev": (i32, [vp, vp, i64, i64, i64, vp, i64, i64, i64, i32, i64, i64, i32, f64, vp, i64, vp]),
    "nxs_stft_f32_host": (i32, [vp, vp, i64, i64, i64, vp, i64, i64, i64, i32, i64, i64, i32, f64, vp]),
    "nxs_istft_c64_dev": (i32, [vp, vp, i64, i64, i64, vp, i64, i64, i64, i32, f64, vp, vp]),
    "nxs_istft_c64_host": (i32, [vp, vp, i64, i64, i64, vp, i64, i64, i64, i32, f64, vp]),
    "nxs_istft_c2r_f32_dev": (i32, [vp, vp, i64, i64, i64, vp, i64, i64, i64, i32, f64, vp, vp]),
    "nxs_istft_c2r_f32_host": (i32, [vp, vp, i64, i64, i64, vp, i64, i64, i64, i32, f64, vp]),
    "nxs_median_f32_dev": (i32, [vp, vp, i32, C.POINTER(i64), C.POINTER(i64), vp, vp]),
    "nxs_median_f32_host": (i32, [vp, vp, i32, C.POINTER(i64), C.POINTER(i64), vp]),
    "nxs_wiener_dev": (i32, [vp, vp, i32, i32, C.POINTER(i64), C.POINTER(i64), i32, f64, vp, vp]),
    "nxs_wiener_host": (i32, [vp, vp, i32, i32, C.POINTER(i64), C.POINTER(i64), i32, f64, vp]),
    "nxs_argrelextrema_f32_dev": (i32, [vp, vp, i32, C.POINTER(i64), i32, i32, i32, vp, vp, vp]),
    "nxs_argrelextrema_f32_host": (i32, [vp, vp, i32, C.POINTER(i64), i32, i32, i32, vp, C.POINTER(i64)]),
    "nxs_as_windowed_dev": (i32, [vp, vp, i32, i64, i64, i64, i64, i64, i32, i64, i64, vp, vp]),
    "nxs_as_windowed_host": (i32, [vp, vp, i32, i64, i64, i64, i64, i64, i32, i64, i64, vp]),
    "nxs_overlap_and_add_f32_dev": (i32, [vp, vp, i64, i64, i64, i64, vp, vp]),
    "nxs_overlap_and_add_c64_dev": (i32, [vp, vp, i64, i64, i64, i64, vp, vp]),
    "nxs_overlap_and_add_f32_host": (i32, [vp, vp, i64, i64, i64, i64, vp]),
    "nxs_overlap_and_add_c64_host": (i32, [vp, vp, i64, i64, i64, i64, vp]),
    "nxs_fir_out_len": (i32, [i64, i64, i32, C.POINTER(i64)]),
    "nxs_fir_f32_dev": (i32, [vp, vp, i64, i64, i64, vp, i64, i32, vp, i64, vp]),
    "nxs_fir_f32_host": (i32, [vp, vp, i64, i64, i64, vp, i64, i32, vp, i64]),
    "nxs_convolve_nd_dev": (i32, [vp, vp, C.POINTER(i64), vp, C.POINTER(i64), i32, i32, vp, vp]),
    "nxs_convolve_nd_host": (i32, [vp, vp, C.POINTER(i64), vp, C.POINTER(i64), i32, i32, vp]),
    "nxs_bcast_coeffs_dev": (i32, [vp, vp, vp, i64, i32, vp]),
}

_lib = None
_lock = threading.Lock()
_ctxs = {}


class NxSignalArgumentError(ValueError):
    """The reference's only error convention is ``ArgumentError``; this is its Python face."""


def lib():
    """Loads the shared library (built by ``__graft_entry__.build()`` / ``make -C nx_signal_b200/csrc``)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nx_signal_b200 has no CPU fallback)")
                l = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(l, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = l
    return _lib


def strerror(code):
    return lib().nxs_strerror(code).decode()


def check(code, ctx=None, what=""):
    if code == NXS_OK:
        return
    msg = strerror(code)
    if ctx is not None and code in (NXS_ECUDA, NXS_ENCCL, NXS_ENOMEM):
        detail = lib().nxs_last_error(ctx).decode()
        if detail:
            msg += f" ({detail})"
    if what:
        msg = f"{what}: {msg}"
    if code in (NXS_EINVAL, NXS_ESHAPE):
        raise NxSignalArgumentError(msg)
    raise RuntimeError(msg)


def context(device=0):
    """One context per (thread, device); created lazily, needs a CUDA device."""
    key = (threading.get_ident(), int(device))
    ctx = _ctxs.get(key)
    if ctx is None:
        h = vp()
        check(lib().nxs_ctx_create(int(device), C.byref(h)), what="nxs_ctx_create")
        ctx = h
        _ctxs[key] = ctx
    return ctx


def launch_count(device=0):
    return int(lib().nxs_ctx_launch_count(context(device)))


def profile(enable, device=0):
    check(lib().nxs_ctx_profile(context(device), int(bool(enable))))


def profile_read(device=0):
    """(summed dominant-kernel milliseconds, launches) since the last read."""
    ms, n = f64(), i64()
    check(lib().nxs_ctx_profile_read(context(device), C.byref(ms), C.byref(n)), context(device))
    return ms.value, n.value


def synchronize(device=0):
    check(lib().nxs_ctx_synchronize(context(device)), context(device))


def host_timeline(device=0):
    """Phases of the last nxs_stft_f32_host call (seconds since its start): enqueued, first slab
    landed, last slab landed, mirror done."""
    t = (f64 * 4)()
    check(lib().nxs_ctx_host_timeline(context(device), C.byref(t)))
    return [float(v) for v in t]


HOST_MODES = {0: "full_d2h", 1: "onesided_d2h+host_mirror", 2: "onesided_d2h+pinned_ring_unstage",
              3: "mixed: 3 of 4 chunks onesided_d2h+host_mirror, 1 of 4 full_d2h",
              4: "small_call_single_stream_full_d2h"}


def host_mode(device=0):
    """How the last nxs_stft_f32_host call moved its result (see nxs_ctx_host_mode) and whether the
    input went through the pinned input ring."""
    m = i32()
    check(lib().nxs_ctx_host_mode(context(device), C.byref(m)))
    return {"result": HOST_MODES.get(m.value & 15, str(m.value & 15)), "input_staged": bool(m.value & 16)}


def set_host_mode(mode, device=0):
    """-1: the context picks the cheapest mode from its own measurements (default); 0 / 1 / 3: pin it."""
    check(lib().nxs_ctx_set_host_mode(context(device), int(mode)))


def build_info():
    """The stamp compiled into the loaded library and the same hash computed now from the sources in the tree
    (None when they are not there): {"lib": ..., "src_sha256_of_tree": ..., "lib_matches_tree": bool}."""
    import glob
    import hashlib

    stamp = lib().nxs_build_info().decode()
    csrc = os.path.join(_HERE, "csrc")
    files = sorted(os.path.basename(p) for pat in ("*.cu", "*.cuh", "*.h") for p in glob.glob(os.path.join(csrc, pat)))
    files += [f for f in ("nxs_host.cpp", "nxs_hostpool.cpp")]
    files = sorted(set(files))
    tree = None
    try:
        h = hashlib.sha256()
        for f in files:
            h.update(open(os.path.join(csrc, f), "rb").read())
        h.update(open(os.path.join(_HERE, "..", "include", "nxsignal_b200.h"), "rb").read())
        tree = h.hexdigest()[:16]
    except OSError:
        pass
    lib_hash = stamp.split("src_sha256=")[1].split()[0] if "src_sha256=" in stamp else None
    return {"lib": stamp, "src_sha256_of_tree": tree, "lib_matches_tree": (tree == lib_hash) if tree else None,
            "path": os.path.relpath(LIB_PATH, os.path.join(_HERE, ".."))}
