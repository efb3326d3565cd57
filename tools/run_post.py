"""Times median / wiener / argrelmax on a magnitude spectrogram [frames][bins] resident on the device.
usage: run_post.py [frames] [bins] [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import nx_signal_b200 as nx
from nx_signal_b200 import _lib, _arrays as A
F = int(sys.argv[1]) if len(sys.argv) > 1 else 359904
B = int(sys.argv[2]) if len(sys.argv) > 2 else 513
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda", 0)
mag = torch.randn(F, B, device=dev).abs_()
out = torch.empty_like(mag)
ctx = _lib.context(0); lib = _lib.lib(); s = A.stream_of(mag)
shape = (C.c_int64 * 2)(F, B)
n = F * B

ONLY = os.environ.get("RUN_POST_ONLY", "")

def timed(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    _lib.profile(True); _lib.profile_read()
    for _ in range(iters): fn()
    ms, k = _lib.profile_read()
    _lib.profile(False)
    return ms / iters

for ks in [] if ONLY and ONLY != "median" else [(1, 3), (1, 9), (1, 17), (1, 31), (17, 1), (3, 3), (5, 5)]:
    k = (C.c_int64 * 2)(*ks)
    ms = timed(lambda: _lib.check(lib.nxs_median_f32_dev(ctx, A.ptr(mag), 2, shape, k, A.ptr(out), s), ctx))
    print(f"MEDIAN {F}x{B} window {ks}: {ms:.3f} ms  {8*n/(ms*1e-3)/1e9:.1f} GB/s algorithmic (4 B in + 4 B out per element)  {n/(ms*1e-3)/1e9:.2f} Gelem/s")
for ks in [] if ONLY and ONLY != "wiener" else [(3, 3), (5, 5), (1, 9)]:
    k = (C.c_int64 * 2)(*ks)
    ms = timed(lambda: _lib.check(lib.nxs_wiener_dev(ctx, A.ptr(mag), 0, 2, shape, k, 0, 0.0, A.ptr(out), s), ctx))
    print(f"WIENER {F}x{B} window {ks} (noise estimated): {ms:.3f} ms  {8*n/(ms*1e-3)/1e9:.1f} GB/s algorithmic  {n/(ms*1e-3)/1e9:.2f} Gelem/s")
idx = torch.empty(n, 2, dtype=torch.int32, device=dev)
cnt = torch.zeros((), dtype=torch.int64, device=dev)
for axis, order in [] if ONLY and ONLY != "argrel" else [(1, 1), (1, 8), (0, 1)]:
    ms = timed(lambda: _lib.check(lib.nxs_argrelextrema_f32_dev(ctx, A.ptr(mag), 2, shape, axis, order, 1, A.ptr(idx), A.ptr(cnt), s), ctx))
    print(f"ARGRELMAX {F}x{B} axis {axis} order {order}: {ms:.3f} ms  {12*n/(ms*1e-3)/1e9:.1f} GB/s algorithmic (4 B in + 8 B of indices out per element)  {n/(ms*1e-3)/1e9:.2f} Gelem/s  valid={int(cnt)}")
