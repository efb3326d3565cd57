"""One nfft-4096 STFT call over C channels against the same work as C/128 calls of 128 channels each (views of the same tensors).
usage: run_split_calls.py [channels]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nx_signal_b200 as nx
from nx_signal_b200 import _lib, _arrays as A
C = int(sys.argv[1]) if len(sys.argv) > 1 else 512
L, nfft, hop = 2880000, 4096, 1024
M = (L - nfft) // hop + 1
dev = torch.device("cuda", 0); ctx = _lib.context(0); lib = _lib.lib()
x = torch.randn(C, L, device=dev); w = torch.from_numpy(nx.windows.hann(nfft)).to(dev)
z = torch.empty((C, M, nfft, 2), device=dev)
def call(c0, n):
    _lib.check(lib.nxs_stft_f32_dev(ctx, A.ptr(x[c0:]), n, L, L, A.ptr(w), nfft, hop, nfft, _lib.PAD_VALID, 0, 0, _lib.SCALE_NONE,
                                    48000.0, A.ptr(z[c0:]), A.stream_of(x)), ctx)
def timed(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
one = timed(lambda: call(0, C))
split = timed(lambda: [call(c0, 128) for c0 in range(0, C, 128)])
algo = 4 * C * L + 8 * C * M * nfft
print(f"nfft 4096, {C} ch x 60 s ({algo / 1e9:.0f} GB per pass): one call {one:.3f} ms ({algo / one / 1e6:.0f} GB/s)   {C // 128} calls of 128 ch {split:.3f} ms ({algo / split / 1e6:.0f} GB/s)")
