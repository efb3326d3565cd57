"""Times nxs_stft_f32_host (pinned host buffers, H2D + kernel + D2H + host mirror) on cfg2.
usage: python tools/run_e2e.py [channels] [seconds] [iters]   (env: NXS_HOST_THREADS, NXS_HOST_NO_MIRROR)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nx_signal_b200 as nx
from nx_signal_b200 import _lib, _arrays as A

C = int(sys.argv[1]) if len(sys.argv) > 1 else 8
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 600
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
nfft, hop = 1024, 256
L = int(48000 * secs)
M = (L - nfft) // hop + 1
xh = torch.randn(C, L).pin_memory()
zh = torch.empty((C, M, nfft), dtype=torch.complex64).pin_memory()
w = nx.windows.hann(nfft)
ctx = _lib.context(0); lib = _lib.lib()
def step():
    _lib.check(lib.nxs_stft_f32_host(ctx, A.ptr(xh), C, L, L, w.ctypes.data, nfft, hop, nfft, 0, 0, 0, 0, 48000.0, A.ptr(zh)), ctx)
step()
ts = []
for _ in range(iters):
    t = time.perf_counter(); step(); ts.append(time.perf_counter() - t)
best, mean = min(ts), sum(ts) / len(ts)
print(f"threads={os.environ.get('NXS_HOST_THREADS','default')} mirror={'off' if os.environ.get('NXS_HOST_NO_MIRROR') else 'on'} "
      f"C={C} frames={C*M}: mean {mean*1e3:.1f} ms best {best*1e3:.1f} ms  {C*M/mean/1e6:.2f} Mframes/s  "
      f"host result {zh.numel()*8/mean/1e9:.1f} GB/s  timeline(ms) enq/first/last/done = "
      + "/".join(f"{1e3*t:.1f}" for t in _lib.host_timeline(0)))
