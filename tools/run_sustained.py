"""Sustained (power-capped) against burst kernel time: runs one device-resident call back to back for `secs` seconds and
reports the mean CUDA-event time of the first 10 launches (burst) and of the launches of the last third (sustained),
with the SM clock / power / throttle reasons NVML shows at the end of the run.
usage: run_sustained.py stft|istft|fir  <args as tools/run_{stft,istft,fir}.py>  [secs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nx_signal_b200 as nx
from nx_signal_b200 import _lib, _arrays as A

op = sys.argv[1]
C, secs_sig, n1, n2 = int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
run_secs = float(sys.argv[6]) if len(sys.argv) > 6 else 3.0
dev = torch.device("cuda", 0)
ctx = _lib.context(0); lib = _lib.lib()
L = int(48000 * secs_sig)
if op == "stft":
    nfft, hop = n1, n2
    M = (L - nfft) // hop + 1
    x = torch.randn(C, L, device=dev); w = torch.from_numpy(nx.windows.hann(nfft)).to(dev)
    z = torch.empty((C, M, nfft, 2), device=dev)
    def step():
        _lib.check(lib.nxs_stft_f32_dev(ctx, A.ptr(x), C, L, L, A.ptr(w), nfft, hop, nfft, _lib.PAD_VALID, 0, 0, _lib.SCALE_NONE,
                                        48000.0, A.ptr(z), A.stream_of(x)), ctx)
    label = f"STFT C={C} nfft={nfft} hop={hop}"
elif op == "istft":
    nfft, hop = n1, n2
    M = (L - nfft) // hop + 1
    z = torch.randn(C, M, nfft, 2, device=dev); w = torch.from_numpy(nx.windows.hann(nfft)).to(dev)
    y = torch.empty((C, M * hop + nfft - hop, 2), device=dev)
    def step():
        _lib.check(lib.nxs_istft_c64_dev(ctx, A.ptr(z), C, M, nfft, A.ptr(w), nfft, hop, nfft, 0, 48000.0, A.ptr(y), A.stream_of(z)), ctx)
    label = f"ISTFT C={C} nfft={nfft} hop={hop}"
else:
    K = n1
    x = torch.randn(C, L, device=dev); taps = torch.randn(K, device=dev) / K ** 0.5
    y = torch.empty((C, L), device=dev)
    def step():
        _lib.check(lib.nxs_fir_f32_dev(ctx, A.ptr(x), C, L, L, A.ptr(taps), K, _lib.MODE["same"], A.ptr(y), L, A.stream_of(x)), ctx)
    label = f"FIR C={C} K={K}"
for _ in range(3): step()
torch.cuda.synchronize()
time.sleep(1.0)  # start from an idle power state
evs = []
t0 = time.perf_counter()
while time.perf_counter() - t0 < run_secs:
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
info = ""
try:
    import pynvml
    pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
    info = f"  [at the end: SM {pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)} MHz, {pynvml.nvmlDeviceGetPowerUsage(h) / 1000:.0f} W, reasons 0x{pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h):x}]"
except Exception as e:
    info = f"  [nvml: {e!r}]"
ms = [a.elapsed_time(b) for a, b in evs]
n = len(ms)
print(f"{label}: burst (first 10) {sum(ms[:10]) / 10:.4f} ms   sustained (last third of {n} launches over {run_secs:.0f} s) {sum(ms[2 * n // 3:]) / (n - 2 * n // 3):.4f} ms{info}")
