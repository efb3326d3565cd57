"""Runs the fused STFT a few times on a synthetic shape (for ncu / quick timing under gpurun).
usage: python tools/run_stft.py [channels] [seconds] [nfft] [hop] [iters]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nx_signal_b200 as nx
from nx_signal_b200 import _lib, _arrays as A

C = int(sys.argv[1]) if len(sys.argv) > 1 else 8
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 600
nfft = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
hop = int(sys.argv[4]) if len(sys.argv) > 4 else 256
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 5
L = int(48000 * secs)
M = (L - nfft) // hop + 1
dev = torch.device("cuda", 0)
x = torch.randn(C, L, device=dev)
w = torch.from_numpy(nx.windows.hann(nfft)).to(dev)
z = torch.empty((C, M, nfft), dtype=torch.complex64, device=dev)
ctx = _lib.context(0); lib = _lib.lib()
def step():
    _lib.check(lib.nxs_stft_f32_dev(ctx, A.ptr(x), C, L, L, A.ptr(w), nfft, hop, nfft, 0, 0, 0, 0, 48000.0, A.ptr(z), A.stream_of(x)), ctx)
for _ in range(2): step()
torch.cuda.synchronize()
_lib.profile(True); _lib.profile_read()
for _ in range(iters): step()
ms, n = _lib.profile_read()
algo = 4 * C * L + 8 * C * M * nfft
per = ms / iters  # kernel time per CALL: a call over more than ~16 GB is several launches (launch_stft walks channel blocks)
print(f"C={C} L={L} nfft={nfft} hop={hop} frames={C*M}: kernel {per:.4f} ms  {algo/(per*1e-3)/1e9:.1f} GB/s algorithmic  {C*M/(per*1e-3)/1e6:.1f} Mframes/s"
      + (f"  ({n // iters} launches per call)" if n != iters else ""))
