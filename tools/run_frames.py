"""Times the standalone framing ops on device-resident tensors: as_windowed (cfg2's framing: 4x expansion)
and overlap_and_add (cfg5's frames, f32 and c64).  usage: run_frames.py [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nx_signal_b200 import _lib, _arrays as A
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda", 0)
ctx = _lib.context(0); lib = _lib.lib()

def timed(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

# as_windowed: 2 ch x 600 s, window 1024, stride 256 (a quarter of cfg2: 0.92 GB out)
C, L, N, H = 2, 48000 * 600, 1024, 256
M = (L - N) // H + 1
x = torch.randn(C, L, device=dev)
out = torch.empty(C, M, N, device=dev)
s = A.stream_of(x)
ms = timed(lambda: _lib.check(lib.nxs_as_windowed_dev(ctx, A.ptr(x), 4, C, L, L, N, H, 0, 0, 0, A.ptr(out), s), ctx))
b = 4 * C * L + 4 * C * M * N
print(f"AS_WINDOWED C={C} L={L} N={N} stride={H} frames={C*M}: {ms:.3f} ms  {b/(ms*1e-3)/1e9:.1f} GB/s algorithmic (x in once, frames out)")
del x, out
# overlap_and_add: cfg5's frame tensor, 32 ch x 11247 frames x 1024, overlap 768
C, M = 32, 11247
t = torch.randn(C, M, N, device=dev)
y = torch.empty(C, M * H + N - H, device=dev)
ms = timed(lambda: _lib.check(lib.nxs_overlap_and_add_f32_dev(ctx, A.ptr(t), C, M, N, N - H, A.ptr(y), A.stream_of(t)), ctx))
b = 4 * C * M * N + 4 * y.numel()
print(f"OVERLAP_AND_ADD f32 C={C} M={M} N={N} hop={H}: {ms:.3f} ms  {b/(ms*1e-3)/1e9:.1f} GB/s algorithmic")
del t, y
t = torch.randn(C, M, N, 2, device=dev)
y = torch.empty(C, M * H + N - H, 2, device=dev)
ms = timed(lambda: _lib.check(lib.nxs_overlap_and_add_c64_dev(ctx, A.ptr(t), C, M, N, N - H, A.ptr(y), A.stream_of(t)), ctx))
b = 8 * C * M * N + 8 * C * (M * H + N - H)
print(f"OVERLAP_AND_ADD c64 C={C} M={M} N={N} hop={H}: {ms:.3f} ms  {b/(ms*1e-3)/1e9:.1f} GB/s algorithmic")
