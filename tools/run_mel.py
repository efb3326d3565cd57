"""Times stft_to_mel on cfg2's spectrum shape. usage: run_mel.py [channels] [seconds] [nfft] [hop] [mels] [onesided] [sampling_rate]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nx_signal_b200 as nx
from nx_signal_b200 import _lib, _arrays as A
C = int(sys.argv[1]) if len(sys.argv) > 1 else 8
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 600
nfft = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
hop = int(sys.argv[4]) if len(sys.argv) > 4 else 256
mels = int(sys.argv[5]) if len(sys.argv) > 5 else 128
onesided = int(sys.argv[6]) if len(sys.argv) > 6 else 0
sr = float(sys.argv[7]) if len(sys.argv) > 7 else 48000.0
L = int(48000 * secs); M = (L - nfft) // hop + 1
K = nfft // 2 + 1 if onesided else nfft
dev = torch.device("cuda", 0)
z = torch.randn(C, M, K, 2, device=dev)
out = torch.empty(C, M, mels, device=dev)
ctx = _lib.context(0); lib = _lib.lib()
def step():
    _lib.check(lib.nxs_stft_to_mel_f32_dev(ctx, A.ptr(z), C, M, K, nfft, mels, sr, 3016.0, 200 / 3, A.ptr(out), A.stream_of(z)), ctx)
for _ in range(2): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
algo = 8 * C * M * (nfft // 2) + 4 * C * M * mels
print(f"MEL sr={sr:.0f} C={C} M={M} nfft={nfft} mels={mels} K={K}: {ms:.4f} ms per call  {algo/(ms*1e-3)/1e9:.1f} GB/s algorithmic  {C*M/(ms*1e-3)/1e6:.1f} Mframes/s")

# fused stft -> mel on the same workload
x = torch.randn(C, L, device=dev)
w = torch.from_numpy(nx.windows.hann(nfft)).to(dev)
def fstep():
    _lib.check(lib.nxs_stft_mel_f32_dev(ctx, A.ptr(x), C, L, L, A.ptr(w), nfft, hop, nfft, 0, 0, 0, 0, sr, mels, 3016.0, 200 / 3, A.ptr(out), A.stream_of(x)), ctx)
for _ in range(2): fstep()
torch.cuda.synchronize()
e0.record()
for _ in range(10): fstep()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
algo = 4 * C * L + 4 * C * M * mels
print(f"FUSED STFT+MEL sr={sr:.0f} C={C} M={M} nfft={nfft} hop={hop} mels={mels}: {ms:.4f} ms per call  {C*M/(ms*1e-3)/1e6:.1f} Mframes/s  ({algo/(ms*1e-3)/1e9:.1f} GB/s algorithmic: x in, mel out)")
