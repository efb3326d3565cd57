"""Times the c2r ISTFT (one-sided input, real output) on a synthetic shape.
usage: run_istft_c2r.py [channels] [seconds] [nfft] [hop] [iters] [z_ld]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nx_signal_b200 as nx
from nx_signal_b200 import _lib, _arrays as A
C = int(sys.argv[1]) if len(sys.argv) > 1 else 32
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 60
nfft = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
hop = int(sys.argv[4]) if len(sys.argv) > 4 else 256
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 5
K = nfft // 2 + 1
z_ld = int(sys.argv[6]) if len(sys.argv) > 6 else K
L = int(48000 * secs); M = (L - nfft) // hop + 1
dev = torch.device("cuda", 0)
z = torch.randn(C, M, z_ld, 2, device=dev)
w = torch.from_numpy(nx.windows.hann(nfft)).to(dev)
out_len = M * hop + nfft - hop
y = torch.empty((C, out_len), device=dev)
ctx = _lib.context(0); lib = _lib.lib()
def step():
    _lib.check(lib.nxs_istft_c2r_f32_dev(ctx, A.ptr(z), C, M, z_ld, A.ptr(w), nfft, hop, nfft, 0, 48000.0, A.ptr(y), A.stream_of(z)), ctx)
for _ in range(2): step()
torch.cuda.synchronize()
_lib.profile(True); _lib.profile_read()
for _ in range(iters): step()
ms, n = _lib.profile_read()
algo = 8 * C * M * K + 4 * C * out_len
print(f"ISTFT-C2R C={C} M={M} nfft={nfft} hop={hop} z_ld={z_ld} frames={C*M}: kernel {ms/n:.4f} ms  {algo/(ms/n*1e-3)/1e9:.1f} GB/s algorithmic  {C*M/(ms/n*1e-3)/1e6:.1f} Mframes/s")
