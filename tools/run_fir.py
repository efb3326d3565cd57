"""Times FIR overlap-save. usage: run_fir.py [channels] [seconds] [taps] [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nx_signal_b200 as nx
from nx_signal_b200 import _lib, _arrays as A
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 600
K = int(sys.argv[3]) if len(sys.argv) > 3 else 2049
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
L = int(48000 * secs)
dev = torch.device("cuda", 0)
x = torch.randn(C, L, device=dev)
taps = torch.from_numpy(nx.filters.firwin(K, [6000], sampling_rate=48000)).to(dev)
y = torch.empty_like(x)
ctx = _lib.context(0); lib = _lib.lib()
def step():
    _lib.check(lib.nxs_fir_f32_dev(ctx, A.ptr(x), C, L, L, A.ptr(taps), K, 1, A.ptr(y), L, A.stream_of(x)), ctx)
for _ in range(2): step()
torch.cuda.synchronize()
_lib.profile(True); _lib.profile_read()
for _ in range(iters): step()
ms, n = _lib.profile_read()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters): step()
e1.record(); torch.cuda.synchronize()
call_ms = e0.elapsed_time(e1) / iters
passes = n // iters
ms, n = call_ms, 1  # whole call (all partitions of a long filter)
algo = 8 * C * L
print(f"FIR C={C} L={L} K={K} ({passes} pass{'es' if passes != 1 else ''}): call {ms/n:.4f} ms  {algo/(ms/n*1e-3)/1e9:.1f} GB/s algorithmic  {C*L/(ms/n*1e-3)/1e9:.2f} Gsamples/s")
