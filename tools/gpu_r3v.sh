#!/bin/bash
# r03v: sub-warp groups synchronise over their own lane mask: racecheck / memcheck on the kernels with 8- and 16-thread groups, parity, timings
OUT=gpurun_out/r03v; mkdir -p $OUT
cat > /tmp/rc_subwarp.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import nx_signal_b200 as nx
from oracle import nxsignal_oracle as o
rng = np.random.default_rng(0)
for hop in (64, 32, 16):
    z = (rng.standard_normal((3, 700, 128)) + 1j * rng.standard_normal((3, 700, 128))).astype(np.complex64); w = o.hann(128)
    y = nx.istft(torch.from_numpy(z).cuda(), torch.from_numpy(w).cuda(), overlap_length=128 - hop, fft_length=128)
    yo = o.istft_fast(z, w, overlap_length=128 - hop, fft_length=128)
    e = np.abs(y.cpu().numpy() - yo).max() / np.abs(yo).max(); print("istft 128 /", hop, "rel err %.2e" % e); assert e < 1e-4
for nfft, hop, pad in [(256, 64, "valid"), (256, 100, "reflect"), (128, 32, "valid"), (128, 50, "same"), (512, 128, "valid")]:
    x = rng.standard_normal((2, 40 * nfft)).astype(np.float32); w = o.hann(nfft)
    kw = dict(overlap_length=nfft - hop, fft_length=nfft, sampling_rate=48000, window_padding=pad)
    z, _, _ = nx.stft(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), **kw)
    zo, _, _ = o.stft_fast(x, w, **kw)
    e = np.abs(z.cpu().numpy() - zo).max() / np.abs(zo).max(); print("stft", nfft, hop, pad, "rel err %.2e" % e); assert e < 1e-5
PY
timeout 900 compute-sanitizer --tool racecheck python /tmp/rc_subwarp.py > $OUT/racecheck.log 2>&1; tail -3 $OUT/racecheck.log
timeout 900 compute-sanitizer --tool memcheck python /tmp/rc_subwarp.py > $OUT/memcheck.log 2>&1; tail -2 $OUT/memcheck.log
timeout 900 python -m pytest tests/test_istft_gpu.py tests/test_stft_gpu.py tests/test_stft_variants_gpu.py -m gpu -q > $OUT/pytest.log 2>&1; echo "tests: $(tail -1 $OUT/pytest.log)"
{ timeout 120 python tools/run_istft.py 64 60 128 32 10; timeout 120 python tools/run_stft.py 8 600 256 64 10; timeout 120 python tools/run_stft.py 8 600 128 32 10; } > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
